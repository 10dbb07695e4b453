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


def rel_err(a, b):
    """max |a-b| / max |b| : the relative measure used for the 1e-12 parity bar (FP64, north_star)."""
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = np.abs(b).max()
    if scale == 0:
        return np.abs(a).max()
    if a.ndim >= 3:
        # per field (axis 2): a field whose own scale is not negligible is judged against its own scale
        worst = 0.
        for k in range(a.shape[2]):
            sk = np.abs(b[:, :, k]).max()
            sk = sk if sk > 1e-6 * scale else scale
            worst = max(worst, np.abs(a[:, :, k] - b[:, :, k]).max() / sk)
        return worst
    return np.abs(a - b).max() / scale
