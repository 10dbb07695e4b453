"""GPU parity, staged path, on curved serendipity elements: 20-node hexahedra (reference src/eles_hexas.cpp:1215-1356) and
8-node quadrilaterals (src/eles_quads.cpp:1037-1130).  The host setup of both is bit-identical to the reference
(tests/test_host_cpu.py); here three time steps of CalcResidual + AdvanceSolution run on the device with those metrics.
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
