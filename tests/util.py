"""Shared helpers of the parity tests: build a case (mesh + input), run the compiled reference on it
(oracle/_ref/ref_dump, built by oracle/build_ref.sh from /root/reference and shipped to the GPU box as a binary), read
its dump."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from hfd import read_hfd  # noqa: E402

REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DUMP = os.path.join(REF_DIR, "ref_dump")


def have_reference():
    return os.path.exists(REF_DUMP) and os.path.exists(os.path.join(REF_DIR, "data", "JacobiGQ.bin"))


def run_reference(input_file, n_steps, stagewise=True, cwd=None):
    """Run the unmodified reference CPU solver (instrumented driver oracle/ref_dump.cpp) and return its dump."""
    cwd = cwd or os.path.dirname(os.path.abspath(input_file))
    out = os.path.join(cwd, os.path.basename(input_file) + ".ref.hfd")
    env = dict(os.environ, HIFILES_HOME=REF_DIR)
    r = subprocess.run([REF_DUMP, os.path.basename(input_file), out, str(n_steps), "1" if stagewise else "0"], cwd=cwd, env=env,
                       capture_output=True, text=True, timeout=1800)
    if r.returncode != 0 or not os.path.exists(out):
        raise RuntimeError("reference run failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    return read_hfd(out)


def field_scales(b):
    """Per-field scale for the parity measure: max |b| of the field, with the momentum components sharing one scale
    (momentum is one vector quantity: rho*w of a flow in the x-y plane is rounding noise of rho*u, rho*v and must be
    judged against |rho u|, not against itself).  b is (pt, ele, field[, dim]) or (field,)."""
    b = np.asarray(b)
    nf = b.shape[2] if b.ndim >= 3 else b.shape[0]
    take = (lambda k: b[:, :, k]) if b.ndim >= 3 else (lambda k: b[k])
    sc = np.array([np.abs(take(k)).max() for k in range(nf)])
    if nf >= 4:  # Euler / Navier-Stokes: [rho, rho u (n_dims components), E]
        sc[1:nf - 1] = sc[1:nf - 1].max()
    return sc


def rel_err(a, b):
    """Relative error used for the 1e-12 parity bar (FP64, north_star): per field, max |a-b| / scale(field)."""
    a = np.asarray(a); b = np.asarray(b)
    # oracle/ref_dump.cpp (adims) drops trailing singleton dimensions, e.g. n_fields = 1 of the scalar test equation
    while a.ndim > b.ndim and a.shape[-1] == 1:
        a = a[..., 0]
    while b.ndim > a.ndim and b.shape[-1] == 1:
        b = b[..., 0]
    assert a.shape == b.shape, (a.shape, b.shape)
    if np.abs(b).max() == 0:
        return np.abs(a).max()
    if a.ndim >= 3 or (a.ndim == 1 and a.shape[0] in (1, 4, 5)):
        sc = field_scales(b)
        take = (lambda x, k: x[:, :, k]) if a.ndim >= 3 else (lambda x, k: x[k])
        worst = 0.
        for k in range(sc.size):
            d = np.abs(take(a, k) - take(b, k)).max()
            worst = max(worst, d / sc[k] if sc[k] > 0 else d)
        return worst
    return np.abs(a - b).max() / np.abs(b).max()
