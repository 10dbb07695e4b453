"""CPU: Paraview output.  The reference binary writes the initial solution as <name>_000000000.vtu before its first step
(src/HiFiLES.cpp:171-182, src/output.cpp:462-900); the host mirror's write_vtu of the same initial solution must produce
the same file: same XML, same plot-point numbering and sub-cell connectivity, numbers equal to the printed digits."""
import os
import re
import subprocess

import numpy as np
import pytest

import util

REF = os.path.join(util.REF_DIR, "HiFiLES_ref")
NUM = re.compile(r"^[-+]?(\d+\.?\d*|\.\d+)([eE][-+]?\d+)?$")


def split_vtu(path):
    """-> (structure tokens, numbers): every whitespace-separated token of the file, numeric ones replaced by '#'"""
    text, nums = [], []
    for tok in open(path).read().split():
        if NUM.match(tok):
            nums.append(float(tok))
            text.append("#")
        else:
            text.append(tok)
    return text, np.array(nums)


@pytest.mark.parametrize("kind,order,p_res,diag", [("hex", 2, 3, None), ("quadtri", 2, 4, None), ("pritet", 1, 4, None), ("pritet", 2, 2, None),
                                                   ("hex", 2, 2, "8 u V w energy Mach pressure vorticity q_criterion"), ("quadtri", 2, 3, "4 u w pressure vorticity"),
                                                   ("hex", 1, 2, "avg")])
def test_initial_vtu_matches_reference_binary(tmp_path, hb, meshgen, kind, order, p_res, diag, monkeypatch):
    if not (util.have_reference() and os.path.exists(REF)):
        pytest.skip("oracle/_ref not built")
    extra = {}
    if kind == "hex":
        meshgen.hex_box(str(tmp_path / "m.neu"), 2)
    elif kind == "quadtri":
        meshgen.mixed_box_2d(str(tmp_path / "m.neu"), 4, kind="mixed", lengths=(6.2831853071795862,) * 2, origin=(0., 0.))
        extra = dict(dz_cyclic=None)
    else:
        meshgen.mixed_box_3d(str(tmp_path / "m.neu"), (2, 2, 2), kind=kind)
    if diag == "avg":
        extra["average_fields"] = "5 rho_average u_average v_average w_average e_average"  # zero before the first step
    elif diag:
        extra["diagnostic_fields"] = diag  # optional plot fields; the gradient-based ones are zero before the first residual evaluation
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, n_steps=0,
                              p_res=p_res, data_file_name="Plot", **extra)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    for f in ("m.neu", "input"):
        os.symlink(tmp_path / f, ref_dir / f)
    r = subprocess.run([REF, "input"], cwd=str(ref_dir), env=dict(os.environ, HIFILES_HOME=util.REF_DIR), capture_output=True, text=True, timeout=600)
    assert os.path.exists(ref_dir / "Plot_000000000.vtu"), r.stdout[-2000:] + r.stderr[-2000:]
    monkeypatch.chdir(tmp_path)
    with hb.Run(inp, host_only=True) as run:
        run.write_vtu(0)
    ta, na = split_vtu(ref_dir / "Plot_000000000.vtu")
    tb, nb = split_vtu(tmp_path / "Plot_000000000.vtu")
    assert ta == tb, "XML structure / token layout differs"
    assert na.shape == nb.shape and na.size > 100
    scale = np.abs(na).max()
    assert np.abs(na - nb).max() <= 1e-13 * scale
    assert np.array_equal(na[np.abs(na - np.round(na)) == 0], nb[np.abs(na - np.round(na)) == 0])  # integers (connectivity, offsets, types) exactly


@pytest.mark.parametrize("kind,order,p_res,extra_fields", [("hex", 2, 3, {}), ("quadtri", 2, 3, dict(diagnostic_fields="3 u pressure mach")),
                                                        ("pritet", 1, 3, dict(average_fields="2 rho_average u_average"))])
def test_initial_tecplot_file_matches_reference_binary(tmp_path, hb, meshgen, kind, order, p_res, extra_fields, monkeypatch):
    """write_type 1: output::write_tec (reference src/output.cpp:165-451)"""
    if not (util.have_reference() and os.path.exists(REF)):
        pytest.skip("oracle/_ref not built")
    extra = dict(extra_fields)
    if kind == "hex":
        meshgen.hex_box(str(tmp_path / "m.neu"), 2)
    elif kind == "quadtri":
        meshgen.mixed_box_2d(str(tmp_path / "m.neu"), 4, kind="mixed", lengths=(6.2831853071795862,) * 2, origin=(0., 0.))
        extra["dz_cyclic"] = None
    else:
        meshgen.mixed_box_3d(str(tmp_path / "m.neu"), (2, 2, 2), kind=kind)
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, n_steps=0,
                              p_res=p_res, data_file_name="Plot", write_type=1, **extra)
    ref_dir = tmp_path / "ref"
    ref_dir.mkdir()
    for f in ("m.neu", "input"):
        os.symlink(tmp_path / f, ref_dir / f)
    r = subprocess.run([REF, "input"], cwd=str(ref_dir), env=dict(os.environ, HIFILES_HOME=util.REF_DIR), capture_output=True, text=True, timeout=600)
    assert os.path.exists(ref_dir / "Plot_000000000_p0000.plt"), r.stdout[-2000:] + r.stderr[-2000:]
    monkeypatch.chdir(tmp_path)
    with hb.Run(inp, host_only=True) as run:
        run.write_vtu(0)  # the plot file of the format write_type selects
    ta, na = split_vtu(ref_dir / "Plot_000000000_p0000.plt")
    tb, nb = split_vtu(tmp_path / "Plot_000000000_p0000.plt")
    assert ta == tb, "header / token layout differs"
    assert na.shape == nb.shape and na.size > 100
    assert np.abs(na - nb).max() <= 1e-13 * np.abs(na).max()
