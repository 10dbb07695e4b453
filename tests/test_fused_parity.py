"""GPU parity, fused path (the product's fast path for affine hexahedra): whole time steps through
hf_dev_run_steps / hf_dev_rk_stage are compared with the UNMODIFIED reference (oracle/_ref/ref_dump) and with the
staged path, which is bit-exact against the reference (test_staged_parity.py).  The fused kernels use FMA contraction
and a sum-factorised operation order, so they differ from the reference at rounding level; bar: 1e-12 relative on the
solution (BASELINE.json north_star)."""
import numpy as np
import pytest

import util
from test_staged_parity import CASES, EULER_IC, make_case, check

import os

# fused-only cases: other LDG parameters (beta = -0.5 flips the owner of every flux-point pair, tau != 0 adds the penalty
# term; beta = 0.25 has no single owner and must run the two-sided generation-6 kernels; a sheared mesh is affine with
# normals that are not axis aligned)
CASES.update({
    "hex_p2_ns_hllc_betaneg_tau": ("hex", 4, {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5, ldg_beta=-0.5, ldg_tau=0.1)),
    # two-sided generation-6 kernels (beta = 0.25 has no single owner) and RoeM with viscosity, on the vortex field of the
    # Euler case (box with exactly representable geometry)
    "hex_p2_ns_roem_vortex_beta025": ("hex", 4, dict(lengths=(20.,) * 3, origin=(-10.,) * 3),
                                      dict(order=2, adv_type=3, riemann_solve_type=2, viscous=1, dt=1e-3, ic_form=0, test_case=1, dx_cyclic=20.,
                                           dy_cyclic=20., dz_cyclic=20., ldg_beta=0.25, ldg_tau=0.05, **EULER_IC)),
    "hex_p2_ns_roem_vortex": ("hex", 4, dict(lengths=(20.,) * 3, origin=(-10.,) * 3),
                              dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-3, ic_form=0, test_case=1, dx_cyclic=20.,
                                   dy_cyclic=20., dz_cyclic=20., **EULER_IC)),
    "hex_p3_ns_rusanov_beta025": ("hex", 3, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, dt=1e-5, ldg_beta=0.25, ldg_tau=0.05)),
    # RoeM on the Taylor-Green box: see test_roem_on_rounding_level_normal_mach
    "hex_p3_ns_roem_tgv": ("hex", 3, dict(origin=(0.3, 0.2, 0.1)), dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5)),
})
FUSED_CASES = ["hex_p2_ns_hllc_rk34", "hex_p3_ns_rusanov_rk45", "hex_p2_euler_roem_rk24", "hex_p1_ns_sutherland_euler", "hex_p4_ns_hllc_rk34",
               "hex_p2_euler_hllc_shockcap",  # fused stages + the shock-capturing kernel after each of them
               "hex_p2_ns_hllc_betaneg_tau", "hex_p2_ns_roem_vortex_beta025", "hex_p2_ns_roem_vortex", "hex_p3_ns_rusanov_beta025"]
TOL = 1e-12


def expected_variant(name):
    o = CASES[name][3]
    if not o["viscous"]:
        return "generation 6 (inviscid"
    if abs(o.get("ldg_beta", 0.5)) != 0.5:
        return "generation 6 (two-sided"
    return "generation 9" if o["order"] in (2, 3, 4) else "generation 7"  # generation 9 is the default at P = 2, 3, 4 (hf_fused.cu)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FUSED_CASES)
def test_fused_steps_vs_reference(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        assert run.fused_status() == "available", run.fused_status()
        assert run.fused_variant().startswith(expected_variant(name)), run.fused_variant()
        n0 = run.launch_count()
        run.run(n_steps, fused=True)
        assert run.launch_count() > n0
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], TOL)
        check("final disu_upts", run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], TOL)
        # the pointwise residual is a derivative of fluxes that agree to 1 ulp: rounding is amplified by ~(P+1)^2/h, so
        # this extra array check (not part of the north-star bar, which names the residual *history*) is looser
        check("final div_tconf_upts", run.download("hex", "div_tconf_upts"), ref["final.hex.div_tconf_upts"], 5e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FUSED_CASES)
def test_fused_residual_only_vs_staged(tmp_path, hb, meshgen, name):
    """CalcResidual alone (no update) through both kernel families on the same state."""
    inp = make_case(tmp_path, meshgen, name)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.calc_residual(0)
        staged = run.download("hex", "div_tconf_upts")
        run.set_mode(True)
        run.calc_residual(0)
        fused = run.download("hex", "div_tconf_upts")
        check("div_tconf_upts fused vs staged", fused, staged, 5e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hex_p2_ns_hllc_rk34", "hex_p4_ns_hllc_rk34", "hex_p2_ns_hllc_betaneg_tau"])
def test_two_sided_kernels_on_one_sided_cases(tmp_path, hb, meshgen, name, monkeypatch):
    """The generation-6 (two-sided LDG) kernels stay the path for |beta| != 0.5: keep them checked on the cases that
    normally run generation 7, and check that the two generations agree to rounding."""
    inp = make_case(tmp_path, meshgen, name)
    with hb.Run(inp) as run:
        assert run.fused_variant().startswith(("generation 7", "generation 9"))
        run.run(2, fused=True)
        u7 = run.download("hex", "disu_upts")
    monkeypatch.setenv("HF_FUSED_GEN6", "1")
    with hb.Run(inp) as run:
        assert run.fused_variant().startswith("generation 6 (two-sided")
        run.run(2, fused=True)
        u6 = run.download("hex", "disu_upts")
    check("generation 7 vs 6", u7, u6, 1e-13)
    if util.have_reference():
        ref = util.run_reference(inp, 2, stagewise=False)
        check("generation 6 vs reference", u6, ref["final.hex.disu_upts"], TOL)


@pytest.mark.gpu
def test_roem_on_rounding_level_normal_mach(tmp_path, hb, meshgen):
    """RoeM's f = |Ma_n|^h factor (reference src/inters.cpp:400-404, with the `Ma_n != 0 ? pow : 1` switch) is discontinuous at
    Ma_n = 0: f jumps from exactly 1 at Ma_n = 0 to 1 - 40 h at Ma_n = 1e-17.  The reference's normals differ inside a face in the
    last bit (of the flux points of one face some carry an exact 0 in a component, others 1e-16), so on the Taylor-Green box (w = 0,
    flow along the z faces) WHERE the normal Mach number is exactly zero depends on the normal of the individual flux point.  Round 1's
    fused kernels used one normal per face and landed 2e-7 .. 2e-6 from the reference (HLLC / Rusanov on the same case: 7e-16);
    they now hand RoeM the reference's normal of every flux point (hf_fused_prepare, `nlf`) and agree to 1e-12 like every other
    path.  The last block shows that it was this switch and not ill-conditioning: a last-bit change of a tenth of the initial
    field moves the reference's own arithmetic (the staged kernels, bit-identical) by 1e-15 only."""
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, "hex_p3_ns_roem_tgv")
    ref = util.run_reference(inp, 3, stagewise=False)
    with hb.Run(inp) as run:
        run.set_mode(False)
        run.run(3, fused=False)
        check("staged", run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], 1e-14)
    for env in ({}, {"HF_FUSED_GEN6": "1"}, {"HF_FUSED_GEN7": "1"}, {"HF_FUSED_GEN9": "1"}):
        for k in ("HF_FUSED_GEN6", "HF_FUSED_GEN7", "HF_FUSED_GEN9"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            with hb.Run(inp) as run:
                assert run.fused_status() == "available"
                run.run(3, fused=True)
                check("fused " + run.fused_variant(), run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], TOL)
        finally:
            for k in env:
                os.environ.pop(k, None)
    with hb.Run(inp) as run:
        run.set_mode(False)
        u0 = run.download("hex", "disu_upts")
        rng = np.random.default_rng(7)
        up = np.where(rng.random(u0.shape) < 0.1, np.nextafter(u0, np.inf), u0)
        run.upload("hex", "disu_upts", up)
        run.run(3, fused=False)
        check("reference arithmetic, last-bit perturbed input", run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], TOL)


@pytest.mark.gpu
def test_warped_mesh_falls_back_to_staged(tmp_path, hb, meshgen):
    inp = make_case(tmp_path, meshgen, "hex_p2_warped_ns_rk414")
    with hb.Run(inp) as run:
        assert "affine" in run.fused_status()
        ref = util.run_reference(inp, 1, stagewise=False) if util.have_reference() else None
        run.run(1, fused=True)  # must silently use the staged kernels, not fail and not use the fused ones
        if ref is not None:
            check("final disu_upts", run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], 1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hex_p1_ns_sutherland_euler", "hex_p2_ns_hllc_rk34", "hex_p2_ns_hllc_betaneg_tau", "hex_p3_ns_rusanov_rk45", "hex_p4_ns_hllc_rk34"])
def test_generation9_at_every_order_vs_reference_and_generation7(tmp_path, hb, meshgen, name, monkeypatch):
    """k_face9 + k_resid9 (hf_fused9.cuh) are the default at P = 2, 3, 4; HF_FUSED_GEN9=1 turns them on at any order.  Three time
    steps against the unmodified reference, and against generation 7 (HF_FUSED_GEN7=1) which they must reproduce to rounding."""
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    monkeypatch.setenv("HF_FUSED_GEN9", "1")
    with hb.Run(inp) as run:
        assert run.fused_variant().startswith("generation 9"), run.fused_variant()
        run.run(n_steps, fused=True)
        u9, r9, d9 = run.download("hex", "disu_upts"), run.norm_residual(), run.download("hex", "div_tconf_upts")
    monkeypatch.delenv("HF_FUSED_GEN9")
    monkeypatch.setenv("HF_FUSED_GEN7", "1")
    with hb.Run(inp) as run:
        assert run.fused_variant().startswith("generation 7"), run.fused_variant()
        run.run(n_steps, fused=True)
        u7 = run.download("hex", "disu_upts")
    check("generation 9 vs 7", u9, u7, 1e-13)
    if util.have_reference():
        ref = util.run_reference(inp, n_steps, stagewise=False)
        check("residual norm", r9, ref["history.norm_residual"][:, -1], TOL)
        check("final disu_upts", u9, ref["final.hex.disu_upts"], TOL)
        check("final div_tconf_upts", d9, ref["final.hex.div_tconf_upts"], 5e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["generation 9", "generation 7", "staged"])
def test_nan_residual_stops_the_run(tmp_path, hb, meshgen, mode, monkeypatch):
    """The reference scans div_tconf_upts for NaN after every stage and aborts ("Residual is NaN ...", src/eles.cpp:1781-1795); the
    device kernels raise a flag that hf_dev_run_steps turns into the same failure."""
    inp = make_case(tmp_path, meshgen, "hex_p4_ns_hllc_rk34")
    if mode == "generation 7":
        monkeypatch.setenv("HF_FUSED_GEN7", "1")
    with hb.Run(inp) as run:
        if mode == "staged":
            run.set_mode(False)
        else:
            assert run.fused_variant().startswith(mode)
        u = run.download("hex", "disu_upts")
        u[3, 5, 0] = np.nan
        run.upload("hex", "disu_upts", u)
        with pytest.raises(hb.HiFiLESError, match="Residual is NaN"):
            run.run(1, fused=mode != "staged")


@pytest.mark.gpu
def test_two_phase_upload_equals_synchronous_upload(tmp_path, hb, meshgen):
    """hf_dev_upload_begin / hf_dev_upload_commit (the copy overlaps the previous step's kernels) leave the same state behind as
    hf_dev_upload = eles::cp_disu_upts_cpu_gpu: two steps from the same uploaded field, bit for bit; a second begin without a commit
    is refused."""
    inp = make_case(tmp_path, meshgen, "hex_p4_ns_hllc_rk34")
    with hb.Run(inp) as run:
        u0 = run.download("hex", "disu_upts")
        run.run(1, fused=True)
        run.upload("hex", "disu_upts", u0)
        run.run(2, fused=True)
        a = run.download("hex", "disu_upts")
        keep = run.upload_begin("hex", "disu_upts", u0)  # travels while the device is idle or busy: ordered by the commit
        with pytest.raises(hb.HiFiLESError, match="not committed"):
            run.upload_begin("hex", "disu_upts", u0)
        run.upload_commit("hex")
        keep2 = run.upload_begin("hex", "disu_upts", a)  # the next upload may start while the steps below run
        run.run(2, fused=True)
        b = run.download("hex", "disu_upts")
        run.upload_commit("hex")
        c = run.download("hex", "disu_upts")
        del keep, keep2
    assert np.array_equal(a, b)
    assert np.array_equal(a, c)


# ---- boundary faces inside the sum-factorised kernels (generation 9; hf_fused9.cuh, k_face9<..., BDY>) ---------------------------------------
# Affine hexahedra with walls / inlets / outlets / far fields: the face kernel evaluates the ghost state (bdy_inters::set_boundary_conditions,
# reference src/bdy_inters.cpp:340-1019), the boundary's LDG common solution (u_c = u_r), Riemann flux and viscous boundary flux
# (:213-338, :1024-1090, set_boundary_gradients :1138-1189) for the element that owns the face.  Every boundary kind of the Navier-Stokes
# equations, each order P = 1..4, all three Riemann solvers, both signs of beta.
_FREE = dict(ic_form=1, Mach_c_ic=0.3, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0.05, T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17,
             T_free_stream=300., L_free_stream=1.)
_CHANNEL_BCS = {"x-": "In", "x+": "Out", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Top"}
BDY_CASES = {
    # adiabatic wall + characteristic outflow, RoeM, RK33 (the blocked-kernel case of tests/test_elem_parity.py, now on the fused path)
    "hex_p2_ns_wall_char_periodic": None,
    # density / velocity inlet, pressure outlet, dual-consistent slip walls, Rusanov, RK45, Sutherland viscosity
    "hex_p2_ns_subinsimp_slipdual": None,
    # BASELINE config 3's order: Riemann-invariant far field (char) inlet, pressure outlet, isothermal wall below, moving adiabatic wall above, HLLC, RK34
    "hex_p4_ns_char_isotherm_adiabat_hllc": ("hex", (3, 2, 3), dict(lengths=(1.5, 1., 1.5), bcs=_CHANNEL_BCS),
                                             dict(_FREE, order=4, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-6, fix_vis=0, dx_cyclic=None, dy_cyclic=1.,
                                                  dz_cyclic=None, bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300., bc_In_nx=1.,
                                                  bc_In_ny=0., bc_In_nz=0., bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall",
                                                  bc_Wall_T_static=310., bc_Top_type="adiabat_wall", bc_Top_u=20., calc_force=1, monitor_cp_freq=100000,
                                                  area_ref=1.0)),
    # supersonic inlet / outlet, slip wall (no viscous boundary flux) and isothermal wall, RoeM, negative beta with a penalty
    "hex_p3_ns_supin_supout_slip_roem_betaneg": ("hex", (3, 2, 3), dict(lengths=(1.5, 1., 1.5), bcs=_CHANNEL_BCS),
                                                 dict(order=3, adv_type=2, riemann_solve_type=2, viscous=1, ic_form=1, dt=2e-6, dx_cyclic=None, dy_cyclic=1.,
                                                      dz_cyclic=None, Mach_c_ic=1.8, nx_c_ic=1., ny_c_ic=0., nz_c_ic=0.02, T_c_ic=290., rho_c_ic=1.2,
                                                      Mach_free_stream=1.8, rho_free_stream=1.2, T_free_stream=290., L_free_stream=1., ldg_beta=-0.5, ldg_tau=0.1,
                                                      bc_In_type="sup_in", bc_In_p_static=101000., bc_In_mach=1.8, bc_In_T_static=290., bc_In_nx=1., bc_In_ny=0.,
                                                      bc_In_nz=0., bc_Out_type="sup_out", bc_Wall_type="slip_wall", bc_Top_type="isotherm_wall", bc_Top_T_static=300.,
                                                      calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
    # total-pressure inlet (sub_in_char), characteristic outlet, walls on four sides (no periodic direction: corner elements with three
    # boundary faces), P = 1, forward Euler
    "hex_p1_ns_subinchar_suboutchar_walls": ("hex", (3, 3, 2), dict(lengths=(1.5, 1., 1.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Wall", "z-": "Wall",
                                                                                                 "z+": "Top"}),
                                             dict(_FREE, order=1, adv_type=0, riemann_solve_type=3, viscous=1, dt=1e-5, dx_cyclic=None, dy_cyclic=None, dz_cyclic=None,
                                                  bc_Cyclic_type=None, bc_In_type="sub_in_char", bc_In_p_total=107200., bc_In_T_total=305.4, bc_In_nx=1., bc_In_ny=0.,
                                                  bc_In_nz=0., bc_Out_type="sub_out_char", bc_Out_p_static=100500., bc_Wall_type="adiabat_wall",
                                                  bc_Top_type="slip_wall_dual", calc_force=1, monitor_cp_freq=100000, area_ref=1.0)),
}
for _k, _v in BDY_CASES.items():
    if _v is not None:
        CASES[_k] = _v


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(BDY_CASES))
def test_fused_kernels_with_boundary_faces_vs_reference(tmp_path, hb, meshgen, name):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    with hb.Run(inp) as run:
        assert run.fused_status() == "available", run.fused_status()
        assert run.fused_variant().startswith("generation 9"), run.fused_variant()
        n0 = run.launch_count()
        run.run(n_steps, fused=True)
        launches = run.launch_count() - n0
        res = run.norm_residual()
        u = run.download("hex", "disu_upts")
        div = run.download("hex", "div_tconf_upts")
    # two kernels per RK stage (+ the face values of the initial field + the residual norm), no interface kernels
    n_rk = {0: 1, 1: 3, 2: 4, 3: 5, 4: 14}[CASES[name][3]["adv_type"]]
    assert launches <= 3 * n_rk * n_steps + 10, launches
    check("residual norm", res, ref["history.norm_residual"][:, -1], TOL)
    check("final disu_upts", u, ref["final.hex.disu_upts"], TOL)
    check("final div_tconf_upts", div, ref["final.hex.div_tconf_upts"], 5e-11)


@pytest.mark.gpu
def test_fused_boundary_faces_residual_vs_staged(tmp_path, hb, meshgen):
    """CalcResidual alone on a mesh with boundary faces through both kernel families on the same state; and HF_FUSED_BDY=0 keeps such a
    mesh on the blocked element kernels."""
    inp = make_case(tmp_path, meshgen, "hex_p4_ns_char_isotherm_adiabat_hllc")
    with hb.Run(inp) as run:
        assert run.fused_status() == "available", run.fused_status()
        run.set_mode(False)
        run.calc_residual(0)
        staged = run.download("hex", "div_tconf_upts")
        run.set_mode(True)
        run.calc_residual(0)
        fused = run.download("hex", "div_tconf_upts")
    check("div_tconf_upts fused vs staged", fused, staged, 5e-11)
    os.environ["HF_FUSED_BDY"] = "0"
    try:
        with hb.Run(inp) as run:
            assert run.fused_status() != "available" and run.elem_status() == "available", (run.fused_status(), run.elem_status())
    finally:
        del os.environ["HF_FUSED_BDY"]
