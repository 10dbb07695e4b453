"""GPU parity, staged path, cases added late in the round.  Curved serendipity elements: 20-node hexahedra (reference src/eles_hexas.cpp:1215-1356) and
8-node quadrilaterals (src/eles_quads.cpp:1037-1130).  The host setup of both is bit-identical to the reference
(tests/test_host_cpu.py); here three time steps of CalcResidual + AdvanceSolution run on the device with those metrics.
Then the characteristic subsonic inlet, plain and with the pressure / temperature ramp.
Written after the round's GPU time was spent: first run on a B200 is the driver's round-end test pass."""
import pytest

import util
from test_staged_parity import check

TWO_PI = 6.2831853071795862

CASES = {
    "hex20_p2_curved_ns_hllc_rk34": ("hex20_box", 3, dict(warp=0.15), dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=1e-5)),
    "quad8_p2_curved_ns_rusanov_rk45": ("quad8_box", 4, dict(lengths=(TWO_PI, TWO_PI), origin=(0., 0.), curve=0.05),
                                        dict(order=2, adv_type=3, riemann_solve_type=0, viscous=1, dt=1e-5, dz_cyclic=None)),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CASES))
def test_time_steps_on_curved_elements(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    gen, n, mkw, opts = CASES[name]
    getattr(meshgen, gen)(str(tmp_path / (name + ".neu")), n, **mkw)
    inp = meshgen.write_input(str(tmp_path / ("input_" + name)), name + ".neu", **opts)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=True)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(n_steps, fused=False)
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], 1e-13)
        for t in run.ele_types():
            check("final disu_upts " + t, run.download(t, "disu_upts"), ref["final." + t + ".disu_upts"], 1e-13)
            check("final div_tconf_upts " + t, run.download(t, "div_tconf_upts"), ref["final." + t + ".div_tconf_upts"], 1e-13)
        assert run.launch_count() > 0


# Characteristic subsonic inlet with total pressure / temperature, plain and ramped over the time steps (reference
# src/bdy_inters.cpp:471-585; the ramp advances once per step, src/HiFiLES.cpp:224-225).  The unmodified reference needs
# calc_force for any case with an inlet.
CHANNEL = dict(order=2, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=1e-4, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1.,
               ny_c_ic=0., nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17, T_free_stream=300.,
               L_free_stream=1., dx_cyclic=None, dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None,
               bc_In_type="sub_in_char", bc_In_p_total=107200., bc_In_T_total=305.4, bc_In_nx=1., bc_In_ny=0.,
               bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
               bc_Top_type="adiabat_wall", bc_Top_u=20., calc_force=1, monitor_cp_freq=100000, area_ref=1.0)
INLET_CASES = {
    "quad_p2_ns_sub_in_char": CHANNEL,
    "quad_p2_ns_sub_in_char_ramped": dict(CHANNEL, bc_In_pressure_ramp=1, bc_In_p_ramp_coeff=0.2, bc_In_T_ramp_coeff=0.25,
                                          bc_In_p_total_old=104000., bc_In_T_total_old=303.),
    # T_ramp_coeff < 0: total temperature from the isentropic relation across the interface (src/bdy_inters.cpp:500-501)
    "quad_p2_ns_sub_in_char_ramped_isentropic": dict(CHANNEL, bc_In_pressure_ramp=1, bc_In_p_ramp_coeff=0.2, bc_In_T_ramp_coeff=-1.,
                                                     bc_In_p_total_old=104000.),
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(INLET_CASES))
def test_time_steps_with_total_pressure_inlet(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    meshgen.quad_box(str(tmp_path / "m.neu"), (8, 6), lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"})
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", **INLET_CASES[name])
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(n_steps, fused=False)
        # the inlet state goes through pow(): device and host libm differ by an ulp there
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], 1e-12)
        check("final disu_upts", run.download("quad", "disu_upts"), ref["final.quad.disu_upts"], 1e-12)
        check("final div_tconf_upts", run.download("quad", "div_tconf_upts"), ref["final.quad.div_tconf_upts"], 1e-12)


# Scalar advection-diffusion test equation (equation 1, one field): Lax-Friedrichs common flux (reference
# src/inters.cpp:535-557), LDG for the diffusive part, periodic meshes.  Host setup is bit-identical on the CPU
# (tests/test_host_cpu.py); the device physics for one field (hf_physics.cuh, NF == 1) has its first B200 run here.
from test_host_cpu import ADVECTION_DIFFUSION, ADVECTION_DIFFUSION_CASES


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quad_sine_single", "quad_sine_group_warped", "hex_sine_single"])
def test_time_steps_of_the_advection_diffusion_equation(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    gen, n, mkw, opts = ADVECTION_DIFFUSION_CASES[name]
    getattr(meshgen, gen)(str(tmp_path / "m.neu"), n, **mkw)
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", **dict(ADVECTION_DIFFUSION, **opts))
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(n_steps, fused=False)
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], 1e-13)
        for t in run.ele_types():
            check("final disu_upts " + t, run.download(t, "disu_upts"), ref["final." + t + ".disu_upts"], 1e-13)
            check("final div_tconf_upts " + t, run.download(t, "div_tconf_upts"), ref["final." + t + ".div_tconf_upts"], 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("gen,n,mkw", [("quad_box", (5, 4), dict(lengths=(2., 2.), origin=(-1., -1.), bcs={"x-": "Wall", "x+": "Wall", "y-": "Wall", "y+": "Wall"})),
                                       ("hex_box", (3, 2, 3), dict(lengths=(2., 2., 2.), origin=(-1., -1., -1.),
                                                                  bcs={"x-": "Wall", "x+": "Wall", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Wall"}))])
def test_advection_diffusion_with_dirichlet_walls(tmp_path, hb, meshgen, gen, n, mkw):
    """ad_wall: the trivial Dirichlet boundary of the scalar test equation (u_r = 0, reference src/bdy_inters.cpp:1010-1019) on all / some sides"""
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    getattr(meshgen, gen)(str(tmp_path / "m.neu"), n, **mkw)
    three_d = gen == "hex_box"
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", **dict(ADVECTION_DIFFUSION, ic_form=2, order=2, dx_cyclic=None, dy_cyclic=2. if three_d else None,
                                                                    dz_cyclic=None, bc_Cyclic_type="cyclic" if three_d else None, bc_Wall_type="ad_wall"))
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(n_steps, fused=False)
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], 1e-13)
        for t in run.ele_types():
            check("final disu_upts " + t, run.download(t, "disu_upts"), ref["final." + t + ".disu_upts"], 1e-13)
            check("final div_tconf_upts " + t, run.download(t, "div_tconf_upts"), ref["final." + t + ".div_tconf_upts"], 1e-13)


@pytest.mark.gpu
@pytest.mark.parametrize("test_case", [2, 3])
def test_advection_diffusion_error_file_matches_reference_binary(tmp_path, hb, meshgen, test_case):
    """test_case 2 / 3: error of the solution and of its gradient against the decaying sine waves, integrated over the volume
    cubature and appended to error.dat (reference src/eles.cpp:5076-5276, src/output.cpp:2052-2160) by the command-line driver."""
    import os
    import subprocess
    import numpy as np
    from test_driver_parity import REF, OURS
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    lines = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        meshgen.quad_box(str(d / "m.neu"), (6, 6), lengths=(2., 2.), origin=(-1., -1.))
        meshgen.write_input(str(d / "input"), "m.neu", **dict(ADVECTION_DIFFUSION, ic_form=test_case, test_case=test_case, dx_cyclic=2., dy_cyclic=2.,
                                                              dz_cyclic=None, n_steps=5, monitor_res_freq=5))
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines[who] = open(d / "error.dat").read().strip().splitlines()
    assert len(lines["ref"]) == len(lines["ours"])
    for la, lb in zip(lines["ref"], lines["ours"]):
        a, b = [t.strip() for t in la.split(",") if t.strip()], [t.strip() for t in lb.split(",") if t.strip()]
        assert len(a) == len(b)
        for x, y in zip(a, b):
            try:
                fx, fy = float(x), float(y)
            except ValueError:
                assert x == y
                continue
            assert abs(fx - fy) <= 2e-6 * max(abs(fx), 1e-300)   # printed with 7 significant digits
