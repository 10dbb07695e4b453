"""Generates the golden fixtures tests/golden/*.npz by running the UNMODIFIED reference solver (oracle/_ref/ref_dump,
compiled by oracle/build_ref.sh from /root/reference) on small generated meshes.  Run in the build container:
    python tests/golden/make_golden.py
Each fixture holds the reference's own setup arrays (operators, metrics, connectivity, initial condition) and its
results (residual after the first stage, solution after the first stage and after n steps, residual-norm history), so
the numpy restatement (oracle/hifiles_oracle.py) and the host setup can be pinned without the reference binary."""
import os
import pathlib
import sys
import tempfile

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
import conftest  # noqa: E402
import util  # noqa: E402
from test_staged_parity import EULER_IC  # noqa: E402

GOLDEN = {
    "hex3_p2_ns_hllc_rk34": ("hex", 3, {}, dict(order=2, adv_type=2, riemann_solve_type=3, viscous=1, dt=2e-5), 2),
    "hex2_p3_ns_rusanov_rk45": ("hex", 2, {}, dict(order=3, adv_type=3, riemann_solve_type=0, viscous=1, dt=1e-5), 2),
    "hex3_p1_ns_roem_sutherland_rk24": ("hex", 3, {}, dict(order=1, adv_type=1, riemann_solve_type=2, viscous=1, fix_vis=0, dt=2e-5), 2),
    "quad4_p3_euler_vortex_hllc_rk45": ("quad", 4, {}, dict(order=3, adv_type=3, riemann_solve_type=3, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                                                          dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, **EULER_IC), 3),
    "quad4_p2_ns_rusanov_euler": ("quad", 4, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                                  dict(order=2, adv_type=0, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None), 3),
    "tri3_p3_ns_rusanov_rk34": ("tri", 3, dict(lengths=(6.2831853071795862,) * 2, origin=(0., 0.)),
                                dict(order=3, adv_type=2, riemann_solve_type=0, viscous=1, dt=2e-5, dz_cyclic=None), 2),
    "tet1_p2_ns_roem_rk34": ("tet", 1, {}, dict(order=2, adv_type=2, riemann_solve_type=2, viscous=1, dt=1e-5), 2),
    "pri1_p2_ns_hllc_rk45": ("pri", (1, 2, 1), {}, dict(order=2, adv_type=3, riemann_solve_type=3, viscous=1, dt=1e-5), 2),
}
KINDS = ["hex", "quad", "tri", "tet", "pri"]
KEEP_PREFIX = ("meta", "params", "rk_a", "rk_b", "history.", "final.", "mesh.f2c", "mesh.f2loc_f", "mesh.rot_tag", "mesh.c2v", "mesh.xv",
               "step0.stage0.s18_corrected_divergence", "step0.stage0.advanced", "step0.stage0.s09_common_invFlux", "step0.stage0.s11_correct_gradient",
               "hex.", "quad.", "tri.", "tet.", "pri.", "int_quad.", "int_seg.", "int_tri.")


def main():
    hb = conftest.load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    only = sys.argv[1:]
    for name, (kind, n, mkw, opts, steps) in GOLDEN.items():
        if only and name not in only:
            continue
        work = pathlib.Path(tempfile.mkdtemp())
        mesh = str(work / (name + ".neu"))
        from test_staged_parity import make_mesh
        make_mesh(mg, kind, mesh, n, mkw)
        inp = mg.write_input(str(work / ("input_" + name)), name + ".neu", **opts)
        ref = util.run_reference(inp, steps, stagewise=True)
        keep = {k.replace(".", "__"): v for k, v in ref.items() if k.startswith(KEEP_PREFIX)}
        keep["case__kind"] = np.array([KINDS.index(kind)])
        keep["case__n"] = np.array([n])
        keep["case__steps"] = np.array([steps])
        keep["case__mesh_text"] = np.frombuffer(open(mesh, "rb").read(), dtype=np.uint8)
        keep["case__input_text"] = np.frombuffer(open(inp, "rb").read(), dtype=np.uint8)
        out = HERE / (name + ".npz")
        np.savez_compressed(out, **keep)
        print(name, "%.0f kB" % (os.path.getsize(out) / 1e3))


if __name__ == "__main__":
    main()
