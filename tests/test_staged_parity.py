"""GPU parity, staged path: every method CalcResidual calls (reference src/solver.cpp:50-223) is run through the
C ABI (hf_dev_eles_op / hf_dev_int_inters_op / hf_dev_bdy_inters_op) and its output array is compared with the
dump the UNMODIFIED reference writes after the same method (oracle/ref_dump.cpp, stagewise mode), then whole time
steps are compared.  The staged kernels are compiled without FMA contraction and keep the reference's summation
order, so on the B200 they reproduce the reference CPU build bit for bit (measured: error exactly 0 on every array of
every case; 1 ulp where the Sutherland pow() is evaluated).  Tolerance here: 1e-14 relative, far inside the 1e-12 of
BASELINE.json's north_star; the residual norm is a reduction in a different order (1e-13)."""
import numpy as np
import pytest

import util

TOL = 1e-14

EULER_IC = dict(u_c_ic=1., v_c_ic=1., w_c_ic=0., p_c_ic=1., rho_c_ic=1.)

CASES = {
    # name: (mesh kind, n, mesh kwargs, input overrides)
    "hex_p2_ns_hllc_rk34": ("hex", 4, {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5)),
    "hex_p3_ns_rusanov_rk45": ("hex", 3, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, dt=1e-5)),
    "hex_p2_euler_roem_rk24": ("hex", 4, dict(lengths=(20.,) * 3, origin=(-10.,) * 3),
                               dict(order=2, adv_type=1, riemann_solve_type=2, viscous=0, dt=1e-3, ic_form=0, test_case=1, dx_cyclic=20.,
                                    dy_cyclic=20., dz_cyclic=20., **EULER_IC)),
    "hex_p1_ns_sutherland_euler": ("hex", 5, {}, dict(order=1, adv_type=0, riemann_solve_type=3, viscous=1, fix_vis=0, dt=1e-5)),
    "hex_p2_warped_ns_rk414": ("hex", 4, dict(warp=0.15), dict(order=2, adv_type=4, riemann_solve_type=3, viscous=1, dt=1e-5)),
    "hex_p4_ns_hllc_rk34": ("hex", 3, {}, dict(order=4, adv_type=2, riemann_solve_type=3, viscous=1, dt=5e-6)),
    "quad_p3_euler_vortex_rk45": ("quad", 8, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                    dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, **EULER_IC)),
    "quad_p2_ns_hllc_rk34": ("quad", 6, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                             dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5, dz_cyclic=None)),
    # CFL-based time steps: calc_time_step + eles::calc_dt_local (reference src/solver.cpp:484-549, src/eles.cpp:1267-1356)
    "quad_p2_ns_cfl_global_dt": ("quad", 6, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.), warp=0.2),
                                 dict(order=2, adv_type=3, riemann_solve_type=0, viscous=1, dt_type=1, CFL=0.4, dt=None, dz_cyclic=None)),
    "hex_p2_ns_cfl_local_dt": ("hex", 3, dict(warp=0.2), dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt_type=2, CFL=0.4, dt=None)),
    # boundary conditions (reference src/bdy_inters.cpp:213-1189): staged kernels only
    "quad_p2_ns_walls_char_out": ("quad", (8, 6), dict(lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"}),
                                  dict(order=2, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=1e-4, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
                                       ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17, T_free_stream=300.,
                                       L_free_stream=1., dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
                                       bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0.,
                                       bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
                                       bc_Top_type="adiabat_wall", bc_Top_u=20.)),
    "quad_p3_euler_slip_supin_supout": ("quad", (6, 5), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall"}),
                                        dict(order=3, adv_type=2, riemann_solve_type=3, viscous=0, ic_form=1, dt=1e-5, dx_cyclic=None, dy_cyclic=None,
                                             dz_cyclic=None, bc_Cyclic_type=None, u_c_ic=600., v_c_ic=20., w_c_ic=0., p_c_ic=100000., rho_c_ic=1.2,
                                             bc_In_type="sup_in", bc_In_p_static=101000., bc_In_mach=1.8, bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0.,
                                             bc_Out_type="sup_out", bc_Wall_type="slip_wall", calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    # the three remaining boundary kinds: density / velocity inlet with free pressure (sub_in_simp), dual-consistent slip wall
    # (slip_wall_dual: mirrored normal velocity) -- reference src/bdy_inters.cpp:374-394, 976-995 -- inviscid and viscous
    "quad_p3_euler_subinsimp_slipdual": ("quad", (6, 5), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall"}),
                                         dict(order=3, adv_type=2, riemann_solve_type=3, viscous=0, ic_form=1, dt=1e-5, dx_cyclic=None, dy_cyclic=None,
                                              dz_cyclic=None, bc_Cyclic_type=None, u_c_ic=100., v_c_ic=4., w_c_ic=0., p_c_ic=100000., rho_c_ic=1.2,
                                              bc_In_type="sub_in_simp", bc_In_rho=1.21, bc_In_u=102., bc_In_v=3., bc_In_w=0.,
                                              bc_Out_type="sub_out_simp", bc_Out_p_static=99500., bc_Wall_type="slip_wall_dual", calc_force=1, monitor_cp_freq=100000,
                                              area_ref=1.0)),
    "hex_p2_ns_subinsimp_slipdual": ("hex", (3, 2, 3), dict(lengths=(1.5, 1., 1.5), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic",
                                                                                       "z-": "Wall", "z+": "Wall"}),
                                     dict(order=2, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=1e-5, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
                                          ny_c_ic=0., nz_c_ic=0.05, T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17, T_free_stream=300.,
                                          L_free_stream=1., dx_cyclic=None, dy_cyclic=1., dz_cyclic=None, bc_In_type="sub_in_simp", bc_In_rho=1.18,
                                          bc_In_u=105., bc_In_v=0., bc_In_w=4., bc_Out_type="sub_out_simp", bc_Out_p_static=100000.,
                                          bc_Wall_type="slip_wall_dual", calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    "hex_p2_ns_wall_char_periodic": ("hex", (3, 3, 4), dict(lengths=(1., 1., 2.), bcs={"x-": "Cyclic", "x+": "Cyclic", "y-": "Cyclic", "y+": "Cyclic",
                                                                                   "z-": "Wall", "z+": "Far"}),
                                     dict(order=2, adv_type=1, riemann_solve_type=2, viscous=1, ic_form=1, dt=1e-4, Mach_c_ic=0.2, nx_c_ic=1., ny_c_ic=0.,
                                          nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.2, rho_free_stream=1.17, T_free_stream=300.,
                                          L_free_stream=1., dx_cyclic=1., dy_cyclic=1., dz_cyclic=None, bc_Wall_type="adiabat_wall",
                                          bc_Far_type="sub_out_char", bc_Far_p_static=100500.)),
    # simplex / prism element types (dense Dubiner-basis operators) and mixed meshes: BASELINE configs 2 and 4
    "tri_p3_ns_rusanov_rk34": ("tri", 4, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                               dict(order=3, adv_type=2, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None)),
    "tri_p2_euler_vortex_hllc_rk45": ("tri", 6, {}, dict(order=2, adv_type=3, riemann_solve_type=3, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                       dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, **EULER_IC)),
    "mixed_tri_quad_p3_ns_rusanov_walls": ("mixed", (8, 6), dict(lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"}),
                                           dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=5e-5, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
                                                ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17,
                                                T_free_stream=300., L_free_stream=1., dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
                                                bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0.,
                                                bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
                                                bc_Top_type="adiabat_wall", bc_Top_u=20.)),
    # curved six-node triangles with walls, supersonic inflow, Persson sensor + filter, global CFL time step: the shipped cylinder case's ingredients
    # (reference testcases/navier-stokes/cylinder/input_cylinder_visc) on a generated mesh
    "tri6_p3_ns_curved_supin_wall_cfl_shockcap": ("tri6", (6, 5), dict(lengths=(3., 2.), origin=(0., 0.), curve=0.08,
                                                                      bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall"}),
                                                  dict(order=3, adv_type=2, riemann_solve_type=0, viscous=1, ic_form=1, dt_type=1, CFL=0.5, dt=None, fix_vis=0,
                                                       ldg_tau=0.5, Mach_c_ic=1.1, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17723946,
                                                       Mach_free_stream=1.0, rho_free_stream=1.17723946, T_free_stream=300., L_free_stream=1.,
                                                       dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None, bc_In_type="sup_in",
                                                       bc_In_p_static=101325., bc_In_mach=1.1, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0.,
                                                       bc_Out_type="sup_out", bc_Wall_type="isotherm_wall", bc_Wall_T_static=300., shock_cap=1, s0=1e-24,
                                                       calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    "tet_p3_ns_roem_rk34": ("tet", 2, {}, dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5)),
    "tet_p2_ns_hllc_cfl_local_dt": ("tet", 2, dict(warp=0.15), dict(order=2, adv_type=3, riemann_solve_type=3, viscous=1, dt_type=2, CFL=0.3, dt=None)),
    "pri_p3_ns_roem_rk45": ("pri", 2, {}, dict(order=3, adv_type=3, riemann_solve_type=2, viscous=1, dt=1e-5)),
    "hexpri_p2_ns_roem_rk34": ("hexpri", (2, 2, 4), {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=2e-5)),
    "pritet_p3_ns_roem_rk34": ("pritet", (2, 4, 2), {}, dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5)),
    # polynomial de-aliasing by over-integration (eles::evaluate_invFlux_over_int, reference src/eles.cpp:1480-1545): BASELINE config 4
    "hex_p3_ns_roem_overint": ("hex", 3, dict(warp=0.1), dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5, over_int=1, over_int_order=5)),
    "hexpri_p2_ns_roem_overint": ("hexpri", (2, 2, 4), {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=2e-5, over_int=1, over_int_order=4)),
    "pritet_p2_ns_roem_overint": ("pritet", (2, 4, 2), {}, dict(order=2, adv_type=3, riemann_solve_type=2, viscous=1, dt=1e-5, over_int=1, over_int_order=4)),
    "mixed_tri_quad_p3_euler_overint": ("mixed", 6, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                            dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, over_int=1, over_int_order=5, **EULER_IC)),
    # LES, eddy-viscosity sub-grid models (eles::calc_sgsf_upts, extrapolate_sgsFlux; reference src/eles.cpp:2395-2646, 2817-2914): BASELINE config 5
    "hex_p3_les_wale_rk34": ("hex", 3, dict(warp=0.1), dict(order=3, adv_type=2, riemann_solve_type=3, viscous=1, dt=1e-5, LES=1, SGS_model=1, C_s=0.325,
                                                         filter_ratio=2.0)),
    "pritet_p2_les_wale_rk34": ("pritet", (2, 4, 2), {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5, LES=1, SGS_model=1, C_s=0.325,
                                                            filter_ratio=1.5)),
    "mixed_tri_quad_p3_les_smagorinsky_walls": ("mixed", (8, 6), dict(lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"}),
                                                dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=5e-5, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
                                                     ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17,
                                                     T_free_stream=300., L_free_stream=1., dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
                                                     bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0.,
                                                     bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
                                                     bc_Top_type="adiabat_wall", bc_Top_u=20., LES=1, SGS_model=0, C_s=0.1, filter_ratio=2.0)),
    # filter-based models: WALE-similarity (2), spectral vanishing viscosity (3), similarity (4); eles::calc_sgs_terms at the first stage
    "hex_p3_les_wsm_vasilyev": ("hex", 3, dict(warp=0.1), dict(order=3, adv_type=2, riemann_solve_type=3, viscous=1, dt=1e-5, LES=1, SGS_model=2, C_s=0.325,
                                                            filter_ratio=2.0, filter_type=0)),
    "mixed_tri_quad_p3_les_similarity_gaussian": ("mixed", 4, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                                                  dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None, LES=1, SGS_model=4, C_s=0.3,
                                                       filter_ratio=2.0, filter_type=1)),
    "tet_p2_les_svv_modal": ("tet", 2, {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5, LES=1, SGS_model=3, C_s=0.3, filter_ratio=2.0,
                                                filter_type=2)),
    # wall-modelled LES (calc_wall_stress, reference src/wall_model_funcs.cpp:13-118; bdy_inters.cpp:1095-1131): Werner-Wengle on an isothermal wall,
    # the compressible log law on an adiabatic wall
    "mixed_tri_quad_p3_les_wale_werner_wengle": ("mixed", (8, 6), dict(lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"}),
                                                 dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=5e-5, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
                                                      ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17,
                                                      T_free_stream=300., L_free_stream=1., dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
                                                      bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0.,
                                                      bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
                                                      bc_Wall_use_wm=1, bc_Top_type="adiabat_wall", bc_Top_u=20., LES=1, SGS_model=1, C_s=0.325,
                                                      filter_ratio=2.0, wall_model=1)),
    "hex_p2_les_wale_loglaw_wall": ("hex", (3, 3, 4), dict(lengths=(1., 1., 2.), bcs={"x-": "Cyclic", "x+": "Cyclic", "y-": "Cyclic", "y+": "Cyclic",
                                                                                 "z-": "Wall", "z+": "Far"}),
                                    dict(order=2, adv_type=1, riemann_solve_type=2, viscous=1, ic_form=1, dt=1e-4, Mach_c_ic=0.2, nx_c_ic=1., ny_c_ic=0.,
                                         nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.2, rho_free_stream=1.17, T_free_stream=300.,
                                         L_free_stream=1., dx_cyclic=1., dy_cyclic=1., dz_cyclic=None, bc_Wall_type="adiabat_wall", bc_Wall_use_wm=1,
                                         bc_Far_type="sub_out_char", bc_Far_p_static=100500., LES=1, SGS_model=1, C_s=0.325, filter_ratio=2.0, wall_model=2)),
    # BASELINE config 5 in one case: supersonic flow over a wall-modelled wall on a mixed hex / prism mesh, HLLC, LES (WALE), shock capturing,
    # supersonic inlet / outlet, characteristic far field
    "config5_hexpri_p2_les_wm_shockcap_hllc": ("hexpri", (3, 2, 4), dict(lengths=(1.5, 1., 2.), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic",
                                                                                               "z-": "Wall", "z+": "Far"}),
                                               dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, ic_form=1, dt=2e-6, fix_vis=0, Mach_c_ic=1.8, nx_c_ic=1., ny_c_ic=0.,
                                                    nz_c_ic=0.02, T_c_ic=290., rho_c_ic=1.2, Mach_free_stream=1.8, rho_free_stream=1.2, T_free_stream=290., L_free_stream=1.,
                                                    dx_cyclic=None, dy_cyclic=1., dz_cyclic=None, bc_In_type="sup_in", bc_In_p_static=101000., bc_In_mach=1.8,
                                                    bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0., bc_In_nz=0., bc_Out_type="sup_out",
                                                    bc_Wall_type="adiabat_wall", bc_Wall_use_wm=1, bc_Far_type="char", bc_Far_p_static=101000., bc_Far_mach=1.8,
                                                    bc_Far_T_static=290., bc_Far_nx=1., bc_Far_ny=0., bc_Far_nz=0., LES=1, SGS_model=1, C_s=0.325, filter_ratio=2.0,
                                                    wall_model=1, shock_cap=1, s0=1e-9, expf_cutoff=1, calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    "hexpri_p2_les_smagorinsky_periodic": ("hexpri", (2, 2, 4), {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5, LES=1, SGS_model=0,
                                                                       C_s=0.1, filter_ratio=2.0)),
    # Persson sensor + exponential modal filter after every stage (eles::shock_capture, reference src/eles.cpp:2918-2959): BASELINE config 5;
    # s0 is set inside the range of the sensor values of these smooth fields so that some elements are filtered and some are not
    "mixed_tri_quad_p3_euler_shockcap": ("mixed", 6, {}, dict(order=3, adv_type=3, riemann_solve_type=3, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                             dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, shock_cap=1, s0=1e-7, expf_cutoff=1, **EULER_IC)),
    "hex_p2_euler_hllc_shockcap": ("hex", 4, dict(lengths=(20.,) * 3, origin=(-10.,) * 3),
                                   dict(order=2, adv_type=2, riemann_solve_type=3, viscous=0, dt=1e-3, ic_form=0, test_case=1, dx_cyclic=20.,
                                        dy_cyclic=20., dz_cyclic=20., shock_cap=1, s0=1e-7, **EULER_IC)),
    "hexpri_p2_ns_hllc_shockcap": ("hexpri", (2, 2, 4), {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5, shock_cap=1, s0=5e-6,
                                                              shock_det_field=1)),
    "pritet_p2_ns_hllc_shockcap": ("pritet", (2, 4, 2), {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=1e-5, shock_cap=1, s0=1.5e-6)),
}


def make_mesh(meshgen, kind, mesh, n, mkw):
    if kind == "hex":
        meshgen.hex_box(mesh, n, **mkw)
    elif kind == "quad":
        meshgen.quad_box(mesh, n, **mkw)
    elif kind in ("tri", "mixed", "tri6"):
        meshgen.mixed_box_2d(mesh, n, kind=kind, **mkw)
    else:
        meshgen.mixed_box_3d(mesh, n, kind=kind, **mkw)


def make_case(tmp_path, meshgen, name):
    kind, n, mkw, opts = CASES[name]
    mesh = str(tmp_path / (name + ".neu"))
    make_mesh(meshgen, kind, mesh, n, mkw)
    inp = str(tmp_path / ("input_" + name))
    meshgen.write_input(inp, name + ".neu", **opts)
    return inp


def check(name, got, ref, tol=TOL, scale_by=None):
    """scale_by: for difference quantities (delta_disu_fpts = u_c - u_l) the error is measured against the scale of
    the operands, not of the (cancelling) difference."""
    err = util.rel_err(got, ref) if scale_by is None else np.abs(got - ref).max() / np.abs(scale_by).max()
    assert err <= tol, "%s: relative error %.3e > %.1e" % (name, err, tol)
    return err


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_methods_one_by_one(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, name)
    ref = util.run_reference(inp, 1, stagewise=True)
    visc = CASES[name][3]["viscous"]
    with hb.Run(inp) as run:
        run.set_mode(False)
        types = run.ele_types()
        pre = "step0.stage0."
        if "dt_type" in CASES[name][3]:
            dt = run.calc_time_step()  # calc_time_step precedes the RK loop (reference src/HiFiLES.cpp:199)
            assert abs(dt / ref["step0.dt_time"][0] - 1.0) < 1e-14

        # the SGS models evaluate pow(x, 1.25 / 1.5 / 2.5): device and host libm differ by an ulp there, and every array
        # downstream of the SGS flux (divergence = derivative of the flux) amplifies it: 1e-12 from that point on
        late = 1e-12 if CASES[name][3].get("LES") else TOL

        def each(op, arr, key, tol=TOL):
            for t in types:
                run.eles_op(t, op)
            for t in types:
                check(key + "." + t + "." + arr, run.download(t, arr), ref[pre + key + "." + t + "." + arr], tol)

        if CASES[name][3].get("LES") and CASES[name][3]["SGS_model"] >= 2:
            each("calc_sgs_terms", "disuf_upts", "s01_calc_sgs_terms")
            if CASES[name][3]["SGS_model"] != 3:
                for t in types:
                    nd = run.download(t, "Le").shape[2]
                    key = pre + "s01_calc_sgs_terms." + t
                    # the Leonard tensors are differences of nearly equal products: judged against the products' scale
                    check("Lu " + t, run.download(t, "Lu"), ref[key + ".Lu"], 1e-13, scale_by=np.ones(1))
                    check("Le " + t, run.download(t, "Le"), ref[key + ".Le"][:, :, :nd], 1e-13, scale_by=np.abs(ref[key + ".Le"][:, :, :nd]).max() + np.ones(1))
        each("extrapolate_solution", "disu_fpts", "s02_extrapolate_solution")
        if visc:
            each("calculate_gradient", "grad_disu_upts", "s04_calculate_gradient")
        each("evaluate_invFlux_over_int" if CASES[name][3].get("over_int") else "evaluate_invFlux", "tdisf_upts", "s05_evaluate_invFlux")
        for it in range(3):
            run.int_inters_op(it, 0)
        for it in range(3):
            run.bdy_inters_op(it, 0)
        for t in types:
            check("norm_tconf_fpts inv " + t, run.download(t, "norm_tconf_fpts"), ref[pre + "s09_common_invFlux." + t + ".norm_tconf_fpts"])
            if visc:
                check("delta_disu_fpts " + t, run.download(t, "delta_disu_fpts"), ref[pre + "s09_common_invFlux." + t + ".delta_disu_fpts"],
                      scale_by=ref[pre + "s02_extrapolate_solution." + t + ".disu_fpts"])
        if visc:
            for t in types:
                run.eles_op(t, "correct_gradient")
            for t in types:
                check("grad_disu_upts " + t, run.download(t, "grad_disu_upts"), ref[pre + "s11_correct_gradient." + t + ".grad_disu_upts"])
                check("grad_disu_fpts " + t, run.download(t, "grad_disu_fpts"), ref[pre + "s11_correct_gradient." + t + ".grad_disu_fpts"])
            each("evaluate_viscFlux", "tdisf_upts", "s13_evaluate_viscFlux", late)
            if CASES[name][3].get("LES"):
                for t in types:
                    check("sgsf_upts " + t, run.download(t, "sgsf_upts"), ref[pre + "s13_evaluate_viscFlux." + t + ".sgsf_upts"], late)
                each("extrapolate_sgsFlux", "sgsf_fpts", "s14_extrapolate_sgsFlux", late)
        each("extrapolate_totalFlux", "norm_tdisf_fpts", "s15_extrapolate_totalFlux", late)
        each("calculate_divergence", "div_tconf_upts", "s16_calculate_divergence", late)
        if visc:
            for it in range(3):
                run.int_inters_op(it, 1)
            for it in range(3):
                run.bdy_inters_op(it, 1)
            for t in types:
                check("norm_tconf_fpts visc " + t, run.download(t, "norm_tconf_fpts"), ref[pre + "s17_common_viscFlux." + t + ".norm_tconf_fpts"], late)
        each("calculate_corrected_divergence", "div_tconf_upts", "s18_corrected_divergence", late)
        run.advance_solution(0)  # AdvanceSolution, then shock_capture when it is on (reference src/HiFiLES.cpp:209-217)
        for t in types:
            check("advanced " + t, run.download(t, "disu_upts"), ref["step0.stage0.advanced." + t + ".disu_upts"], late)
        if CASES[name][3].get("shock_cap"):
            flagged = total = 0
            for t in types:
                sensor, want = run.download(t, "sensor"), ref["step0.stage0.advanced." + t + ".sensor"]
                check("sensor " + t, sensor, want, 1e-11)
                assert np.array_equal(sensor >= CASES[name][3]["s0"], want >= CASES[name][3]["s0"])
                flagged += int((want >= CASES[name][3]["s0"]).sum())
                total += want.size
            assert 0 < flagged < total, "the case must filter some elements and leave others alone (%d of %d)" % (flagged, total)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_time_steps_reference_call_sequence(tmp_path, hb, meshgen, name):
    """CalcResidual + AdvanceSolution through the host mirror's reference-named methods, 3 steps."""
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=True)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(n_steps, fused=False)
        if "dt_type" in CASES[name][3]:
            ref_dt = ref["step%d.dt_time" % (n_steps - 1)][0]
            assert abs(run.scalar("dt") / ref_dt - 1.0) < 1e-14, (run.scalar("dt"), ref_dt)
        hist = run.norm_residual()
        loose = 1e-12 if CASES[name][3].get("LES") else 1e-13
        check("residual norm", hist, ref["history.norm_residual"][:, -1], loose)
        for t in run.ele_types():
            # after 3 steps the 1-ulp differences of pow() (Sutherland's law, characteristic BCs, SGS models) have propagated
            check("final disu_upts " + t, run.download(t, "disu_upts"), ref["final." + t + ".disu_upts"], loose)
            check("final div_tconf_upts " + t, run.download(t, "div_tconf_upts"), ref["final." + t + ".div_tconf_upts"], loose)
        assert run.launch_count() > 0
