"""Per-field errors of the fused path against the reference after n steps (printed, no assertions)."""
import pathlib, sys, tempfile
import numpy as np
sys.path.insert(0, str(pathlib.Path(__file__).parent))
import conftest, util  # noqa
import test_staged_parity as T  # noqa
from test_fused_parity import FUSED_CASES  # noqa

hb = conftest.load_package()
import importlib
mg = importlib.import_module("hifiles_solver_b200.meshgen")
n_steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for name in FUSED_CASES:
    tmp = pathlib.Path(tempfile.mkdtemp())
    inp = T.make_case(tmp, mg, name)
    ref = util.run_reference(inp, n_steps, stagewise=False)
    out = {}
    for mode in ("fused", "staged"):
        with hb.Run(inp) as run:
            run.set_mode(mode == "fused")
            run.run(n_steps, fused=True)
            out[mode] = (run.download("hex", "disu_upts"), run.download("hex", "div_tconf_upts"), run.norm_residual())
    print("== %s  (%d steps)" % (name, n_steps))
    for arr, idx in (("disu_upts", 0), ("div_tconf_upts", 1)):
        r = ref["final.hex." + arr]
        for k in range(5):
            sc = np.abs(r[:, :, k]).max()
            ef = np.abs(out["fused"][idx][:, :, k] - r[:, :, k]).max()
            es = np.abs(out["staged"][idx][:, :, k] - r[:, :, k]).max()
            print("   %-16s field %d  scale %.3e  fused abs %.3e rel %.3e | staged abs %.3e" % (arr, k, sc, ef, ef / sc if sc else 0, es))
    print("   residual norm rel: fused %.3e staged %.3e" % (np.abs(out["fused"][2] / ref["history.norm_residual"][:, -1] - 1).max(),
                                                          np.abs(out["staged"][2] / ref["history.norm_residual"][:, -1] - 1).max()))
