"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference solver (oracle/_ref/ref_dump,
compiled by oracle/build_ref.sh from /root/reference) on small generated meshes.  Run in the build container:
    python tests/golden/make_golden.py
Each fixture holds the reference's own setup arrays (operators, metrics, connectivity, initial condition) and its
results (residual after the first stage, solution after the first stage and after n steps, residual-norm history), so
the numpy restatement (oracle/hifiles_oracle.py) and the host setup can be pinned without the reference binary."""
import os
import pathlib
import sys
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import conftest  # noqa: E402
import util  # noqa: E402
from test_staged_parity import EULER_IC  # noqa: E402

GOLDEN = {
    "hex3_p2_ns_hllc_rk34": ("hex", 3, {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5), 2),
    "hex2_p3_ns_rusanov_rk45": ("hex", 2, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, dt=1e-5), 2),
    "hex3_p1_ns_roem_sutherland_rk24": ("hex", 3, {}, dict(order=1, adv_type=1, riemann_solve_type=2, viscous=1, fix_vis=0, dt=2e-5), 2),
    "quad4_p3_euler_vortex_hllc_rk45": ("quad", 4, {}, dict(order=3, adv_type=3, riemann_solve_type=3, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                          dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, **EULER_IC), 3),
    "quad4_p2_ns_rusanov_euler": ("quad", 4, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                                  dict(order=2, adv_type=0, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None), 3),
    "tri3_p3_ns_rusanov_rk34": ("tri", 3, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                                dict(order=3, adv_type=2, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None), 2),
    "tet1_p2_ns_roem_rk34": ("tet", 1, {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5), 2),
    "pri1_p2_ns_hllc_rk45": ("pri", (1, 2, 1), {}, dict(order=2, adv_type=3, riemann_solve_type=3, viscous=1, dt=1e-5), 2),
}
# boundary interfaces: the eleven Navier-Stokes boundary kinds (reference src/bdy_inters.cpp:340-1008), inviscid and viscous boundary flux
_FREE = dict(ic_form=1, Mach_c_ic=0.3, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0.05, T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17,
             T_free_stream=300., L_free_stream=1.)
_CHANNEL = dict(lengths=(1.5, 1., 1.5), bcs={"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Top"})
_FORCE = dict(calc_force=1, monitor_cp_freq=100000, area_ref=1.0)  # the unmodified reference needs calc_force for any case with an inlet
GOLDEN.update({
    "hexbdy_p2_char_suboutsimp_isotherm_adiabat_hllc": ("hex", (3, 2, 3), _CHANNEL,
        dict(_FREE, order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-6, fix_vis=0, dx_cyclic=None, dy_cyclic=1., dz_cyclic=None, bc_In_type="char",
             bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1., bc_In_ny=0., bc_In_nz=0., bc_Out_type="sub_out_simp", bc_Out_p_static=100000.,
             bc_Wall_type="isotherm_wall", bc_Wall_T_static=310., bc_Top_type="adiabat_wall", bc_Top_u=20., **_FORCE), 2),
    "hexbdy_p2_supin_supout_slip_isotherm_roem_betaneg": ("hex", (3, 2, 3), _CHANNEL,
        dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, ic_form=1, dt=2e-6, dx_cyclic=None, dy_cyclic=1., dz_cyclic=None, Mach_c_ic=1.8, nx_c_ic=1.,
             ny_c_ic=0., nz_c_ic=0.02, T_c_ic=290., rho_c_ic=1.2, Mach_free_stream=1.8, rho_free_stream=1.2, T_free_stream=290., L_free_stream=1., ldg_beta=-0.5,
             ldg_tau=0.1, bc_In_type="sup_in", bc_In_p_static=101000., bc_In_mach=1.8, bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0., bc_In_nz=0.,
             bc_Out_type="sup_out", bc_Wall_type="slip_wall", bc_Top_type="isotherm_wall", bc_Top_T_static=300., **_FORCE), 2),
    "hexbdy_p1_subinchar_suboutchar_adiabat_slipdual": ("hex", (3, 3, 2), dict(lengths=(1.5, 1., 1.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall",
                                                                                                           "z-": "Wall", "z+": "Top"}),
        dict(_FREE, order=1, adv_type=0, riemann_solve_type=3, viscous=1, dt=1e-5, dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
             bc_In_type="sub_in_char", bc_In_p_total=107200., bc_In_T_total=305.4, bc_In_nx=1., bc_In_ny=0., bc_In_nz=0., bc_Out_type="sub_out_char",
             bc_Out_p_static=100500., bc_Wall_type="adiabat_wall", bc_Top_type="slip_wall_dual", **_FORCE), 2),
    # triangles (segment faces, dense operators): supersonic inflow / outflow, isothermal wall, moving adiabatic wall -- the boundary kinds of the
    # shipped cylinder case (BASELINE config 2's origin)
    "tribdy_p2_ns_supin_supout_isotherm_adiabat_rusanov": ("tri", (5, 4), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"}),
        dict(order=2, adv_type=2, riemann_solve_type=0, viscous=1, ic_form=1, dt=2e-6, fix_vis=0, dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
             Mach_c_ic=1.8, nx_c_ic=1., ny_c_ic=0.02, nz_c_ic=0., T_c_ic=290., rho_c_ic=1.2, Mach_free_stream=1.8, rho_free_stream=1.2, T_free_stream=290.,
             L_free_stream=1., bc_In_type="sup_in", bc_In_p_static=101000., bc_In_mach=1.8, bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0., bc_Out_type="sup_out",
             bc_Wall_type="isotherm_wall", bc_Wall_T_static=300., bc_Top_type="adiabat_wall", bc_Top_u=15., **_FORCE), 2),
    "quadbdy_p3_euler_subinsimp_suboutsimp_slipdual": ("quad", (6, 5), dict(lengths=(3., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall"}),
        dict(order=3, adv_type=2, riemann_solve_type=3, viscous=0, ic_form=1, dt=1e-5, dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
             u_c_ic=100., v_c_ic=4., w_c_ic=0., p_c_ic=100000., rho_c_ic=1.2, bc_In_type="sub_in_simp", bc_In_rho=1.21, bc_In_u=102., bc_In_v=3., bc_In_w=0.,
             bc_Out_type="sub_out_simp", bc_Out_p_static=99500., bc_Wall_type="slip_wall_dual", **_FORCE), 2),
})
KINDS = ["hex", "quad", "tri", "tet", "pri"]
KEEP_PREFIX = ("meta", "params", "rk_a", "rk_b", "history.", "final.", "mesh.f2c", "mesh.f2loc_f", "mesh.rot_tag", "mesh.c2v", "mesh.xv",
               "step0.stage0.s18_corrected_divergence", "step0.stage0.advanced", "step0.stage0.s09_common_invFlux", "step0.stage0.s11_correct_gradient",
               "hex.", "quad.", "tri.", "tet.", "pri.", "int_quad.", "int_seg.", "int_tri.", "bdy_quad.", "bdy_seg.", "bdy_tri.")


def main():
    hb = conftest.load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    only = sys.argv[1:]
    for name, (kind, n, mkw, opts, steps) in GOLDEN.items():
        if only and name not in only:
            continue
        work = pathlib.Path(tempfile.mkdtemp())
        mesh = str(work / (name + ".neu"))
        from test_staged_parity import make_mesh
        make_mesh(mg, kind, mesh, n, mkw)
        inp = mg.write_input(str(work / ("input_" + name)), name + ".neu", **opts)
        ref = util.run_reference(inp, steps, stagewise=True)
        keep = {k.replace(".", "__"): v for k, v in ref.items() if k.startswith(KEEP_PREFIX)}
        keep["case__kind"] = np.array([KINDS.index(kind)])
        keep["case__n"] = np.array(n).reshape(-1)
        keep["case__steps"] = np.array([steps])
        keep["case__mesh_text"] = np.frombuffer(open(mesh, "rb").read(), dtype=np.uint8)
        keep["case__input_text"] = np.frombuffer(open(inp, "rb").read(), dtype=np.uint8)
        out = HERE / (name + ".npz")
        np.savez_compressed(out, **keep)
        print(name, "%.0f kB" % (os.path.getsize(out) / 1e3))


if __name__ == "__main__":
    main()
