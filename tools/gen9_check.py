"""Compare the generation-9 fused kernels (k_face9 + k_resid9) with generation 7 and with the staged kernels on small
Taylor-Green cases.  usage: python tools/gen9_check.py [steps]
(one process per generation: the switch is read when a context is set up, and a process keeps the launch attributes it set)"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = [(4, 3, {}), (4, 4, dict(ldg_beta=-0.5, ldg_tau=0.1)), (2, 4, {}), (3, 3, {}), (1, 4, {}), (5, 2, {})]


def run_case(idx, steps, gen, out):
    import importlib
    import conftest
    hb = conftest.load_package()
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    order, n, extra = CASES[idx]
    w = tempfile.mkdtemp()
    mg.hex_box(os.path.join(w, "m.neu"), n)
    inp = mg.write_input(os.path.join(w, "input"), "m.neu", order=order, viscous=1, adv_type=2, dt=1e-5, riemann_solve_type=3, **extra)
    with hb.Run(inp) as run:
        if gen == "staged":
            run.set_mode(False)
        else:
            assert run.fused_variant().startswith("generation " + gen), run.fused_variant()
        run.calc_residual(0)
        div = run.download("hex", "div_tconf_upts")
        run.run(steps, fused=True)
        np.savez(out, u=run.download("hex", "disu_upts"), div=div)


if __name__ == "__main__":
    if len(sys.argv) > 3:
        run_case(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4])
        sys.exit(0)
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    import util
    worst = 0.
    for idx, (order, n, extra) in enumerate(CASES):
        res = {}
        for gen in ("staged", "7", "9"):
            out = os.path.join(tempfile.mkdtemp(), "r.npz")
            env = dict(os.environ)
            env.pop("HF_FUSED_GEN7", None)
            env.pop("HF_FUSED_GEN9", None)
            env["HF_FUSED_GEN7" if gen == "7" else "HF_FUSED_GEN9"] = "1"
            r = subprocess.run([sys.executable, os.path.abspath(__file__), str(idx), str(steps), gen, out], env=env, capture_output=True, text=True)
            if r.returncode != 0:
                print("order %d: %s FAILED: %s" % (order, gen, (r.stdout + r.stderr)[-1500:]))
                sys.exit(1)
            res[gen] = np.load(out)
        line = "order %d n %d %s:" % (order, n, extra)
        for a, b in (("9", "7"), ("9", "staged"), ("7", "staged")):
            eu, ed = util.rel_err(res[a]["u"], res[b]["u"]), util.rel_err(res[a]["div"], res[b]["div"])
            line += "  %s vs %s: u %.2e residual %.2e" % (a, b, eu, ed)
            if (a, b) == ("9", "staged"):
                worst = max(worst, eu)
        print(line, flush=True)
    sys.exit(0 if worst < 1e-12 else 1)
