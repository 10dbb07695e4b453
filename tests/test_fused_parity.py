"""GPU parity, fused path (the product's fast path for affine hexahedra): whole time steps through
hf_dev_run_steps / hf_dev_rk_stage are compared with the UNMODIFIED reference (oracle/_ref/ref_dump) and with the
staged path, which is bit-exact against the reference (test_staged_parity.py).  The fused kernels use FMA contraction
and a sum-factorised operation order, so they differ from the reference at rounding level; bar: 1e-12 relative on the
solution (BASELINE.json north_star)."""
import numpy as np
import pytest

import util
from test_staged_parity import CASES, make_case, check

FUSED_CASES = ["hex_p2_ns_hllc_rk34", "hex_p3_ns_rusanov_rk45", "hex_p2_euler_roem_rk24", "hex_p1_ns_sutherland_euler", "hex_p4_ns_hllc_rk34",
               "hex_p2_euler_hllc_shockcap"]  # the last one: fused stages + the shock-capturing kernel after each of them
TOL = 1e-12


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
def test_warped_mesh_falls_back_to_staged(tmp_path, hb, meshgen):
    inp = make_case(tmp_path, meshgen, "hex_p2_warped_ns_rk414")
    with hb.Run(inp) as run:
        assert "affine" in run.fused_status()
        ref = util.run_reference(inp, 1, stagewise=False) if util.have_reference() else None
        run.run(1, fused=True)  # must silently use the staged kernels, not fail and not use the fused ones
        if ref is not None:
            check("final disu_upts", run.download("hex", "disu_upts"), ref["final.hex.disu_upts"], 1e-14)
