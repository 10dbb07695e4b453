"""GPU parity of the fast mode on element types with dense operators (triangles, tetrahedra, prisms, mixed meshes): the operator
products run as FP64 tensor-core tiles (k_op_dense_mma, hf_device.cu) instead of the thread-per-output kernel of the bit-exact
yardstick.  The tensor core accumulates a tile in its own order with fused products, so the result differs from the reference in the
last bits; bar: 1e-12 relative on solution and residual-norm history after three time steps (BASELINE.json north_star), against the
unmodified reference CPU solver."""
import pytest

import util
from test_staged_parity import make_case, check

CASES = ["tri_p3_ns_rusanov_rk34", "tet_p3_ns_roem_rk34", "pri_p3_ns_roem_rk45", "pritet_p3_ns_roem_rk34", "pritet_p2_ns_roem_overint",
         "mixed_tri_quad_p3_ns_rusanov_walls"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_fast_mode_steps_vs_reference(tmp_path, hb, meshgen, name, monkeypatch):
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    monkeypatch.setenv("HF_NO_ELEM", "1")  # without the blocked element kernels (test_elem_parity.py) the fast mode is: staged kernels + tensor-core operator products
    with hb.Run(inp) as run:
        run.run(n_steps, fused=True)
        fast = {t: run.download(t, "disu_upts") for t in run.ele_types()}
        check("residual norm", run.norm_residual(), ref["history.norm_residual"][:, -1], 1e-12)
        for t in run.ele_types():
            check("final disu_upts " + t, fast[t], ref["final." + t + ".disu_upts"], 1e-12)
    # the same steps without the tensor-core kernel must land on the bit-exact result: the fast mode changes nothing else
    monkeypatch.setenv("HF_NO_DMMA", "1")
    with hb.Run(inp) as run:
        run.run(n_steps, fused=True)
        for t in run.ele_types():
            check("thread-per-output kernel " + t, run.download(t, "disu_upts"), ref["final." + t + ".disu_upts"], 1e-14)
