"""The drop-in boundary against the reference's OWN objects: oracle/_ref/ref_gpu links the unmodified reference solver
objects -- input, mesh reader, GeoPreprocess, eles_*, int_inters, bdy_inters, InitSolution -- and hands the reference's arrays
(hf_array::get_ptr_cpu() pointers of the eles members, connectivity recovered from the double* tables of set_interior / set_boundary)
to the C ABI of include/hifiles_b200.h where the reference's main loop calls CalcResidual + AdvanceSolution.  Nothing of the repo's
host mirror takes part.  Compared with the reference's CPU run of the same input (ref_dump): mode 0 (staged kernels, the bit-exact
yardstick) at 1e-14, mode 1 (fused / blocked kernels) at the north star's 1e-12."""
import os
import subprocess

import pytest

import util
from test_staged_parity import make_case, check

REF_GPU = os.path.join(util.REF_DIR, "ref_gpu")
CASES = ["hex_p4_ns_hllc_rk34", "hex_p2_ns_wall_char_periodic", "quad_p3_euler_vortex_rk45", "quad_p2_ns_walls_char_out", "mixed_tri_quad_p3_ns_rusanov_walls",
         "tri_p2_euler_vortex_hllc_rk45", "pritet_p3_ns_roem_rk34", "hexpri_p2_ns_roem_overint", "tet_p2_ns_hllc_cfl_local_dt"]


def run_ref_gpu(inp, n_steps, mode):
    cwd = os.path.dirname(os.path.abspath(inp))
    out = os.path.join(cwd, os.path.basename(inp) + ".gpu%d.hfd" % mode)
    env = dict(os.environ, HIFILES_HOME=util.REF_DIR)
    r = subprocess.run([REF_GPU, os.path.basename(inp), out, str(n_steps), str(mode)], cwd=cwd, env=env, capture_output=True, text=True, timeout=900)
    if r.returncode != 0 or not os.path.exists(out):
        raise RuntimeError("ref_gpu failed:\n" + r.stdout[-3000:] + r.stderr[-3000:])
    return util.read_hfd(out), r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_reference_objects_drive_the_c_abi(tmp_path, meshgen, name):
    if not (util.have_reference() and os.path.exists(REF_GPU)):
        pytest.skip("oracle/_ref/ref_gpu not built")
    inp = make_case(tmp_path, meshgen, name)
    n_steps = 3
    ref = util.run_reference(inp, n_steps, stagewise=False)
    for mode, tol in ((0, 1e-14), (1, 1e-12)):
        got, log = run_ref_gpu(inp, n_steps, mode)
        keys = [k for k in got if k.startswith("final.") and k.endswith(".disu_upts")]
        assert keys, log
        for k in keys:
            check("mode %d %s" % (mode, k), got[k], ref[k], tol)
        check("mode %d residual norm" % mode, got["history.norm_residual"][:, -1], ref["history.norm_residual"][:, -1], 1e-12)
    if name == "hex_p4_ns_hllc_rk34":
        assert "fused hexahedron kernels: available" in log  # the named configuration runs generation 9 behind the reference's objects
    else:
        assert "blocked element kernels: available" in log
