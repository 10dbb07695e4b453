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


@pytest.mark.parametrize("kind,patch", [("quad", dict(patch=1, patch_type=0, Mv=0.4, ra=1.5, rb=4.0, xc=0.5, yc=-0.25)),
                                        ("hex", dict(patch=1, patch_type=0, Mv=0.3, ra=1.0, rb=2.5, xc=3.0, yc=3.1)),
                                        ("hex", dict(patch=1, patch_type=1, patch_x=3.0, u_c_ic=30., v_c_ic=-5., w_c_ic=2., p_c_ic=90000.))])
def test_solution_patch_matches_reference(tmp_path, hb, meshgen, kind, patch):
    """eles::set_patch (reference src/eles.cpp:535-652): vortex ring / uniform state laid over the initial solution"""
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    if kind == "quad":
        from test_staged_parity import EULER_IC
        meshgen.quad_box(str(tmp_path / "m.neu"), 6)
        opts = dict(order=3, adv_type=3, riemann_solve_type=0, viscous=0, ic_form=0, test_case=1, dt=1e-3, dx_cyclic=20., dy_cyclic=20., dz_cyclic=None, **EULER_IC)
    else:
        meshgen.hex_box(str(tmp_path / "m.neu"), 3)
        opts = dict(order=2, adv_type=2, riemann_solve_type=0, viscous=1, dt=1e-5)
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", **dict(opts, **patch))
    ref = util.run_reference(inp, 0, stagewise=False)
    plain = util.run_reference(meshgen.write_input(str(tmp_path / "input_plain"), "m.neu", **opts), 0, stagewise=False)
    with hb.Run(inp, host_only=True) as run:
        for k, v in ref.items():
            if k.endswith(".disu_upts_ic"):
                a = run.host_array(k.replace("disu_upts_ic", "disu_upts"))
                assert np.array_equal(a, v), "%s: patched solution differs from the reference (max abs %.3e)" % (k, np.abs(a - v).max())
                assert not np.array_equal(v, plain[k]), "the patch did not touch the solution: the case checks nothing"


@pytest.mark.parametrize("kind,order", [("hex", 2), ("pritet", 1)])
def test_partitioned_restart_files_are_read_by_the_reference(tmp_path, hb, meshgen, kind, order, monkeypatch):
    """Writing side of a partitioned run (reference output::write_restart_ascii, src/output.cpp:1753-1818: one file per rank in
    Rest_<iter>/): two ranks of the host mirror write their part of the initial solution, the UNMODIFIED reference reads the two
    files back (restart_flag 1, n_restart_files 2) into its serial numbering; every value must come back as printed (15 digits)."""
    if not util.have_reference():
        pytest.skip("oracle/_ref not built")
    extra = build_mesh(meshgen, kind, str(tmp_path / "m.neu"))
    common = dict(adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, monitor_res_freq=1, order=order, **extra)
    inp0 = meshgen.write_input(str(tmp_path / "input_write"), "m.neu", n_steps=1, **common)
    monkeypatch.chdir(tmp_path)
    with hb.Run(inp0, host_only=True) as single:
        n_cells = sum(single.host_array(t + ".ele2global_ele").size for t in single.ele_types())
        ic = {t: (single.host_array(t + ".disu_upts").copy(), single.host_array(t + ".ele2global_ele").copy()) for t in single.ele_types()}
    part = (np.arange(n_cells) % 2).astype(np.int32) if kind == "hex" else (np.arange(n_cells) // ((n_cells + 1) // 2)).astype(np.int32)
    for rank in range(2):
        with hb.Run(inp0, rank=rank, nproc=2, part=part, host_only=True) as run:
            run.write_restart(7)
    assert os.path.exists(tmp_path / "Rest_000000007" / "Rest_000000007_p0000.dat") and os.path.exists(tmp_path / "Rest_000000007" / "Rest_000000007_p0001.dat")
    inp = meshgen.write_input(str(tmp_path / "input_read"), "m.neu", n_steps=1, restart_flag=1, restart_iter=7, n_restart_files=2, **common)
    ref = util.run_reference(inp, 0, stagewise=False)
    checked = 0
    for t, (u, gid) in ic.items():
        v = ref[t + ".disu_upts_ic"]
        assert v.shape == u.shape
        scale = np.abs(u).max()
        assert scale > 0 and np.abs(v - u).max() <= 2e-15 * scale * 10, "%s: %.3e" % (t, np.abs(v - u).max() / scale)
        checked += 1
    assert checked >= 1
    # and the host mirror reads its own partitioned files like the reference
    with hb.Run(inp, host_only=True) as run:
        for t, (u, gid) in ic.items():
            a = run.host_array(t + ".disu_upts")
            assert np.abs(a - u).max() <= 2e-14 * np.abs(u).max()
