"""Parity over a real horizon on the named configuration: the Taylor-Green vortex on 15^3 hexahedra (the size of the
mesh the reference ships, testcases/navier-stokes/Taylor_Green_vortex), P = 4, HLLC + LDG, SSP-RK34, 50 time steps.

Two halves, because /root/reference does not exist on the GPU box:
  python tools/horizon_parity.py reference <dir>   (build container, CPU, ~12 min) runs the UNMODIFIED reference binary
        (oracle/_ref/HiFiLES_ref), 50 steps with monitor_res_freq = 1 and ASCII restart files every 25 steps, and packs its
        history.plt and the restart solutions into <dir>/horizon_ref.npz   (oracle/_ref/horizon/ travels to the box, git-ignored)
  python tools/horizon_parity.py compare <dir> [out.txt]   (GPU box) runs the same case through the fused kernels, step by step,
        and prints the relative error of the residual-norm history (every step) and of the solution (steps 25 and 50)."""
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))

N, ORDER, STEPS, DUMP = 15, 4, 50, 25


def make(workdir):
    import importlib
    import conftest
    hb = conftest.load_package()
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    os.makedirs(workdir, exist_ok=True)
    mg.hex_box(os.path.join(workdir, "tgv15.neu"), N)
    dt = 1.440389e-5 * (3.0 / (2 * ORDER + 1))  # the shipped input's time step (P = 1) scaled to P = 4
    inp = mg.write_input(os.path.join(workdir, "input"), "tgv15.neu", order=ORDER, adv_type=2, dt=dt, riemann_solve_type=3, viscous=1, n_steps=STEPS,
                         monitor_res_freq=1, plot_freq=1000000, restart_dump_freq=DUMP)
    return hb, inp


def read_restart(path, nu, nf):
    txt = open(path).read()
    body = txt[txt.index("data") + 4:].split()
    n_eles = N ** 3
    vals = np.array(body[:n_eles * (1 + nu * nf)], dtype=np.float64).reshape(n_eles, 1 + nu * nf)
    gid = vals[:, 0].astype(np.int64)
    u = np.zeros((n_eles, nu, nf))
    u[gid] = vals[:, 1:].reshape(n_eles, nu, nf)
    return u


def reference(d):
    import util
    hb, inp = make(d)
    env = dict(os.environ, HIFILES_HOME=util.REF_DIR)
    if not os.path.exists(os.path.join(d, "Rest_%09d_p0000.dat" % STEPS)):
        r = subprocess.run([os.path.join(util.REF_DIR, "HiFiLES_ref"), "input"], cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:]
    rows = [l for l in open(os.path.join(d, "history.plt")).read().splitlines() if l and l[0].isdigit()]
    hist = np.array([[float(x) for x in l.split(",")[1:6]] for l in rows])  # log10 of the residual norm, five fields
    nu = (ORDER + 1) ** 3
    out = {"history_log10": hist}
    for s in range(DUMP, STEPS + 1, DUMP):
        out["u_%d" % s] = read_restart(os.path.join(d, "Rest_%09d_p0000.dat" % s), nu, 5)
    np.savez_compressed(os.path.join(d, "horizon_ref.npz"), **out)
    print("reference: %d history rows, solutions at steps %s -> %s" % (len(rows), list(range(DUMP, STEPS + 1, DUMP)), os.path.join(d, "horizon_ref.npz")))


def compare(d, out_path=None):
    import tempfile
    ref = np.load(os.path.join(d, "horizon_ref.npz"))
    hb, inp = make(tempfile.mkdtemp(prefix="hf_horizon_"))
    lines = []
    with hb.Run(inp) as run:
        lines.append("# TGV %d^3 hex P=%d, HLLC + LDG(beta 0.5), SSP-RK34, %d steps; kernels: %s; reference: unmodified HiFiLES CPU binary" % (N, ORDER, STEPS, run.fused_variant()))
        lines.append("# step   max rel err of the residual-norm history (five fields, against the reference's history.plt)   solution vs the reference's ASCII restart file")
        worst_h, worst_u = 0., 0.
        for s in range(1, STEPS + 1):
            run.run(1, fused=True)
            res = np.asarray(run.norm_residual())
            refres = 10.0 ** ref["history_log10"][s - 1]
            eh = float(np.abs(res - refres).max() / np.abs(refres).max())
            worst_h = max(worst_h, eh)
            line = "%4d   %.3e" % (s, eh)
            if "u_%d" % s in ref.files:
                u = run.download("hex", "disu_upts").transpose(1, 0, 2)
                gid = run.host_array("hex.ele2global_ele")
                got = np.zeros_like(u)
                got[gid] = u
                r = ref["u_%d" % s]
                sc = np.abs(r).reshape(-1, 5).max(0)
                sc[1:4] = sc[1:4].max()
                eu = float((np.abs(got - r).reshape(-1, 5).max(0) / sc).max())
                worst_u = max(worst_u, eu)
                line += "   %.3e" % eu
            lines.append(line)
        lines.append("# worst: history %.3e, solution %.3e (bar 1e-12; the restart file and history.plt carry 15 significant digits)" % (worst_h, worst_u))
    text = "\n".join(lines)
    print(text)
    if out_path:
        open(out_path, "w").write(text + "\n")
    return 0 if (worst_h <= 1e-12 and worst_u <= 1e-12) else 1


if __name__ == "__main__":
    if sys.argv[1] == "reference":
        reference(sys.argv[2])
    else:
        sys.exit(compare(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None))
