"""CPU: ASCII restart files, reading side.  A restart file written by the UNMODIFIED reference binary is read back by the
reference itself (instrumented driver, restart_flag 1) and by the host mirror; the initial solution must be bit-identical,
also when the file holds another polynomial order than the run (opp_r interpolation, reference src/eles.cpp:3692-3710)."""
import os
import subprocess

import numpy as np
import pytest

import util

REF = os.path.join(util.REF_DIR, "HiFiLES_ref")

CASES = {
    # name: (mesh builder, order of the run that writes the file, order of the restarted run)
    "hex_same_order": ("hex", 2, 2),
    "hex_p2_to_p3": ("hex", 2, 3),
    "quadtri_p2_to_p3": ("mixed2d", 2, 3),
    "pritet_p1_to_p2": ("pritet", 1, 2),
}


def build_mesh(meshgen, kind, path):
    if kind == "hex":
        meshgen.hex_box(path, 2)
        return {}
    if kind == "mixed2d":
        meshgen.mixed_box_2d(path, 4, kind="mixed", lengths=(6.2831853071795862,) * 2, origin=(0., 0.))
        return dict(dz_cyclic=None)
    meshgen.mixed_box_3d(path, (2, 2, 2), kind=kind)
    return {}


@pytest.mark.parametrize("name", list(CASES))
def test_restart_read_matches_reference(tmp_path, hb, meshgen, name, monkeypatch):
    if not (util.have_reference() and os.path.exists(REF)):
        pytest.skip("oracle/_ref not built")
    kind, p_write, p_read = CASES[name]
    extra = build_mesh(meshgen, kind, str(tmp_path / "m.neu"))
    common = dict(adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, monitor_res_freq=1, **extra)
    meshgen.write_input(str(tmp_path / "input_write"), "m.neu", order=p_write, n_steps=2, restart_dump_freq=2, **common)
    env = dict(os.environ, HIFILES_HOME=util.REF_DIR)
    r = subprocess.run([REF, "input_write"], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and os.path.exists(tmp_path / "Rest_000000002_p0000.dat"), r.stdout[-2000:] + r.stderr[-2000:]
    inp = meshgen.write_input(str(tmp_path / "input_read"), "m.neu", order=p_read, n_steps=1, restart_flag=1, restart_iter=2, n_restart_files=1, **common)
    ref = util.run_reference(inp, 0, stagewise=False)
    monkeypatch.chdir(tmp_path)  # restart files are looked up in the working directory
    checked = 0
    with hb.Run(inp, host_only=True) as run:
        for k, v in ref.items():
            if k.endswith(".disu_upts_ic"):
                a = run.host_array(k.replace("disu_upts_ic", "disu_upts"))
                assert a.shape == v.shape
                assert np.array_equal(a, v), "%s: restart data differ from the reference (max abs %.3e)" % (k, np.abs(a - v).max())
                assert np.abs(v).max() > 0
                checked += 1
        assert run.scalar("time") == ref["meta.time0"][0] if "meta.time0" in ref else True
    assert checked >= 1


def test_hdf5_restart_is_rejected_as_in_the_reference_build_without_hdf5(tmp_path, hb, meshgen):
    meshgen.hex_box(str(tmp_path / "m.neu"), 2)
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", order=1, restart_flag=2, restart_iter=2)
    with pytest.raises(hb.HiFiLESError, match="HDF5"):
        hb.Run(inp, host_only=True)
