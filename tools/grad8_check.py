"""Compare the experimental generation-8 gradient kernel (HF_FUSED_GRAD8=1) with generation 7 on small Taylor-Green cases.
usage: python tools/grad8_check.py   (spawns one process per kernel generation: the switch is read once per process)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run_case(order, n, out):
    import importlib
    import conftest
    hb = conftest.load_package()
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    w = tempfile.mkdtemp()
    mg.hex_box(os.path.join(w, "m.neu"), n)
    inp = mg.write_input(os.path.join(w, "input"), "m.neu", order=order, viscous=1, adv_type=2, dt=1e-5, riemann_solve_type=3)
    with hb.Run(inp) as run:
        assert run.fused_variant().startswith("generation 7"), run.fused_variant()
        run.run(2, fused=True)
        np.save(out, run.download("hex", "disu_upts"))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
        sys.exit(0)
    worst = 0.
    for order, n in ((2, 4), (4, 3), (1, 4)):
        res = {}
        for gen in ("7", "8"):
            out = os.path.join(tempfile.mkdtemp(), "u.npy")
            env = dict(os.environ)
            env.pop("HF_FUSED_GRAD8", None)
            if gen == "8":
                env["HF_FUSED_GRAD8"] = "1"
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(order), str(n), out], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                print("order %d: generation %s FAILED: %s" % (order, gen, (r.stdout + r.stderr)[-400:]))
                sys.exit(1)
            res[gen] = np.load(out)
        import util
        err = util.rel_err(res["8"], res["7"])
        worst = max(worst, err)
        print("order %d n %d: generation 8 vs 7 max rel err %.3e" % (order, n, err), flush=True)
    sys.exit(0 if worst < 1e-12 else 1)
