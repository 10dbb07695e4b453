"""Where does the fused RoeM result leave the staged (= reference, bit for bit) one?  Fused vs staged after 3 steps for combinations of
Riemann solver, box origin, viscosity, kernel generation.  usage (GPU box): python tools/roem_probe.py"""
import os, sys, tempfile, importlib, itertools
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import conftest, util
hb = conftest.load_package()
mg = importlib.import_module("hifiles_solver_b200.meshgen")

def case(work, name, n, order, riemann, origin, viscous, **extra):
    d = os.path.join(work, name); os.makedirs(d, exist_ok=True)
    mg.hex_box(os.path.join(d, "m.neu"), n, origin=origin)
    return mg.write_input(os.path.join(d, "input"), "m.neu", order=order, adv_type=2, riemann_solve_type=riemann, viscous=viscous, dt=1e-5, **extra)

def run(inp, fused, env=None):
    for k in ("HF_FUSED_GEN6", "HF_FUSED_GEN7", "HF_FUSED_GEN9", "HF_FUSED_ROEM"):
        os.environ.pop(k, None)
    for k, v in (env or {}).items():
        os.environ[k] = v
    with hb.Run(inp) as r:
        if not fused:
            r.set_mode(False)
        var = (r.fused_variant() if r.fused_status() == "available" else "blocked element kernels") if fused else "staged"
        r.run(3, fused=fused)
        return r.download("hex", "disu_upts"), var

work = tempfile.mkdtemp(prefix="roem_probe_")
i = 0
for riemann, origin, visc, order in itertools.product((2, 3), ((0., 0., 0.), (0.3, 0.2, 0.1)), (1,), (2, 3)):
    i += 1
    inp = case(work, "c%d" % i, 3, order, riemann, origin, visc)
    a, _ = run(inp, False)
    for env in ({}, {"HF_FUSED_GEN6": "1"}, {"HF_NO_FUSED": "1"}):
        b, var = run(inp, True, env)
        err = util.rel_err(b, a)
        d = np.abs(b - a)
        k = np.unravel_index(np.argmax(d), d.shape)
        print("riemann %d origin %s viscous %d P=%d %-28s fused vs staged %.3e  (largest at point %d, element %d, field %d)" % (riemann, origin, visc, order, var[:28], err, k[0], k[1], k[2]), flush=True)
