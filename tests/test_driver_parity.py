"""GPU: the drop-in command line.  `HiFiLES <input_file>` of this repository (host mirror + device layer) and the
reference's own binary (oracle/_ref/HiFiLES_ref) are run on the same input; the residual table printed on stdout and
the log10-residual columns of history.plt (15 digits, reference src/output.cpp:2298-2378) must agree, and so must the
ASCII restart file (Rest_<iter>_p0000.dat: the solution of every element, 15 digits, src/output.cpp:1753-1818)."""
import os
import re
import subprocess

import numpy as np
import pytest

import util

OURS = os.path.join(util.ROOT, "hifiles-solver_b200", "bin", "HiFiLES")
REF = os.path.join(util.REF_DIR, "HiFiLES_ref")


def residual_rows(text):
    rows = []
    for line in text.splitlines():
        t = line.split()
        if len(t) >= 5 and re.fullmatch(r"\d+", t[0]):
            try:
                rows.append([float(x) for x in t[1:6]])
            except ValueError:
                pass
    return np.array(rows)


def history_rows(path, n_fields):
    rows = []
    for line in open(path):
        t = [x.strip() for x in line.split(",")]
        if len(t) > n_fields and re.fullmatch(r"\d+", t[0]):
            rows.append([float(x) for x in t[1:1 + n_fields]])
    return np.array(rows)


@pytest.mark.gpu
def test_command_line_matches_reference_binary(tmp_path, hb, meshgen):
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        meshgen.hex_box(str(d / "tgv.neu"), 4)
        meshgen.write_input(str(d / "input"), "tgv.neu", order=3, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1, n_steps=4, monitor_res_freq=1)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[who] = (residual_rows(r.stdout), history_rows(str(d / "history.plt"), 5))
    assert out["ref"][0].shape == out["ours"][0].shape and out["ref"][0].shape[0] == 4
    assert np.abs(out["ours"][0] - out["ref"][0]).max() <= 1e-8          # printed with 8 decimals
    assert np.abs(out["ours"][1] - out["ref"][1]).max() <= 1e-11         # log10 of the residual norms, 15 digits


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["staged", "fused"])
def test_integral_quantities_match_reference_binary(tmp_path, hb, meshgen, kernels):
    """history.plt with the Taylor-Green diagnostics of the shipped input (integral_quantities kineticenergy enstropy, plus
    the strain products): eles::CalcIntegralQuantities on the device, against the reference binary.  The reference
    evaluates them with the gradient the LAST residual evaluation left behind (the solution before the final RK stage's
    update) -- kept: the staged kernels hold that array anyway, the fused kernels store it on monitored stages."""
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        meshgen.hex_box(str(d / "tgv.neu"), 3)
        meshgen.write_input(str(d / "input"), "tgv.neu", order=3, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1, n_steps=3, monitor_res_freq=1,
                            integral_quantities="4 kineticenergy enstropy pressuredilatation devstraincolonproduct",
                            device_fused=1 if kernels == "fused" else 0)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[who] = history_rows(str(d / "history.plt"), 10)
        if who == "ours":
            head = open(d / "history.plt").read().split("ZONE")[0]
            assert 'Diagnostics[kineticenergy]' in head and 'Diagnostics[enstropy]' in head
    ref, ours = out["ref"], out["ours"]
    assert ref.shape == ours.shape == (3, 10)
    # pressure dilatation and the strain products are sums of cancelling terms: rounding of the element-wise grouping shows at 2e-12
    tol = 1e-11 if kernels == "staged" else 1e-10
    for q in range(5, 9):  # the four diagnostics; column 9 is the physical time
        assert np.abs(ours[:, q] - ref[:, q]).max() <= tol * np.abs(ref[:, q]).max(), "diagnostic %d: %s vs %s" % (q - 5, ours[:, q], ref[:, q])
    assert np.abs(ours[:, 9] - ref[:, 9]).max() <= 1e-13 * np.abs(ref[:, 9]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["quad", "hex"])
def test_vortex_error_file_matches_reference_binary(tmp_path, hb, meshgen, kind):
    """BASELINE config 1: isentropic vortex (test_case 1) on periodic quads, P=3 -- and its 3-D twin on hexahedra (fused
    kernels).  At the end of a test-case run the reference integrates the error against the analytic vortex over the
    volume cubature and appends it to error.dat (src/output.cpp:2052-2160, src/eles.cpp:5076-5290): the files must agree."""
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    from test_staged_parity import EULER_IC
    lines = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        if kind == "quad":
            meshgen.quad_box(str(d / "m.neu"), 8)
            extra = dict(dz_cyclic=None)
        else:
            meshgen.hex_box(str(d / "m.neu"), 4, lengths=(20.,) * 3, origin=(-10.,) * 3)
            extra = dict(dz_cyclic=20.)
        meshgen.write_input(str(d / "input"), "m.neu", order=3, adv_type=3, riemann_solve_type=0, viscous=0, ic_form=0, test_case=1, dt=1e-3,
                            dx_cyclic=20., dy_cyclic=20., n_steps=5, monitor_res_freq=5, **extra, **EULER_IC)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        lines[who] = open(d / "error.dat").read().strip().splitlines()
    assert len(lines["ref"]) == len(lines["ours"]) == 1
    a, b = [t.strip() for t in lines["ref"][0].split(",")], [t.strip() for t in lines["ours"][0].split(",")]
    assert a[:6] == b[:6]                                    # step, order, mesh file, adv_type, riemann_solve_type, norm type
    ea, eb = np.array([float(x) for x in a[6:]]), np.array([float(x) for x in b[6:]])
    assert ea.shape == eb.shape == ((4,) if kind == "quad" else (5,))
    assert np.all(ea[:3] > 0) and np.abs(eb - ea).max() <= 2e-6 * np.abs(ea).max()  # printed with 7 significant digits


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["quad_p2_ns_walls_char_out", "quad_p3_euler_slip_supin_supout", "hex_p2_ns_wall_char_periodic",
                                  "config5_hexpri_p2_les_wm_shockcap_hllc", "mixed_tri_quad_p3_ns_rusanov_walls", "pritet_p2_ns_walls"])
def test_surface_forces_match_reference_binary(tmp_path, hb, meshgen, name):
    """calc_force: output::CalcForces / eles::compute_wall_forces (pressure and viscous traction integrated over the wall
    faces' cubature points) -- force, CL, CD columns of history.plt and the cp / cf file against the reference binary."""
    surface_forces_case(tmp_path, meshgen, name, 0, 1e-10, 1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["hex_p2_ns_wall_char_periodic", "quad_p2_ns_walls_char_out"])
def test_surface_forces_on_the_fast_paths(tmp_path, hb, meshgen, name):
    """The same through the fast kernels: the sum-factorised hexahedron kernels with boundary faces (generation 9) and the blocked element
    kernels keep grad_disu_upts of the monitored steps' last residual evaluation for the wall traction."""
    surface_forces_case(tmp_path, meshgen, name, 1, 1e-9, 1e-8)


def surface_forces_case(tmp_path, meshgen, name, device_fused, tol_force, tol_cp):
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    from test_staged_parity import CASES, make_mesh
    if name == "pritet_p2_ns_walls":  # tetrahedra and prisms with walls on every face kind (x: prism quads / tet triangles, z: triangles)
        base = CASES["hex_p2_ns_wall_char_periodic"]
        kind, n, mkw, opts = "pritet", (2, 4, 2), dict(lengths=(1., 2., 1.), bcs={"x-": "Wall", "x+": "Wall", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Far"}), \
            dict(base[3], dx_cyclic=None, dy_cyclic=2., dz_cyclic=None)
    else:
        kind, n, mkw, opts = CASES[name]
    opts = dict(opts, calc_force=1, monitor_cp_freq=2, area_ref=1.5, n_steps=2, monitor_res_freq=1, device_fused=device_fused)
    nd = 2 if kind in ("quad", "tri", "mixed") else 3
    nf = nd + 2
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        make_mesh(meshgen, kind, str(d / "m.neu"), n, mkw)
        meshgen.write_input(str(d / "input"), "m.neu", **opts)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=900)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        hist = history_rows(str(d / "history.plt"), nf + nd + 2)
        cp = [l.split() for l in open(d / "force_files_000000002" / "cp_000000002_p0000.dat").read().splitlines()]
        out[who] = (hist, cp)
    (ha, ca), (hb_, cb) = out["ref"], out["ours"]
    assert ha.shape == hb_.shape == (2, nf + nd + 2)
    fa, fb = ha[:, nf:], hb_[:, nf:]
    assert np.abs(fa).max() > 0
    assert np.abs(fb - fa).max() <= tol_force * np.abs(fa).max(), "forces / coefficients: %s vs %s" % (fb, fa)
    assert len(ca) == len(cb) and len(ca) > 3
    for la, lb in zip(ca, cb):
        assert len(la) == len(lb)
        if len(la) == 1 or la[0] == "x":
            assert la == lb
        else:
            va, vb = np.array([float(x) for x in la]), np.array([float(x) for x in lb])
            assert np.abs(va - vb).max() <= tol_cp * max(1.0, np.abs(va).max())


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["staged", "fused"])
def test_plot_file_with_diagnostic_fields_matches_reference_binary(tmp_path, hb, meshgen, kernels):
    """Paraview file after two steps with gradient-based plot fields (vorticity, Q criterion): they use grad_disu_upts as the
    last residual evaluation of the plotted step left it, which the fused kernels store on such steps."""
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    from test_plot_cpu import split_vtu
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        meshgen.hex_box(str(d / "tgv.neu"), 3)
        meshgen.write_input(str(d / "input"), "tgv.neu", order=2, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1, n_steps=2, monitor_res_freq=100,
                            plot_freq=2, p_res=3, data_file_name="Mesh", diagnostic_fields="5 pressure mach vorticity q_criterion scaled_q_criterion",
                            device_fused=1 if kernels == "fused" else 0)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[who] = split_vtu(d / "Mesh_000000002.vtu")
    (ta, na), (tb, nb) = out["ref"], out["ours"]
    assert ta == tb and na.shape == nb.shape
    # values are printed with 15 digits; vorticity-type fields are differences of gradients: compare against the field's scale
    tol = 1e-11 if kernels == "staged" else 1e-9
    assert np.abs(na - nb).max() <= tol * np.abs(na).max()
    assert np.abs(na).max() > 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("kernels", ["staged", "fused"])
def test_time_averaged_fields_match_reference_binary(tmp_path, hb, meshgen, kernels):
    """average_fields: the running averages updated on the device after every step (eles::CalcTimeAverageQuantities),
    compared through the Paraview file after three steps"""
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    from test_plot_cpu import split_vtu
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        meshgen.hex_box(str(d / "tgv.neu"), 3)
        meshgen.write_input(str(d / "input"), "tgv.neu", order=2, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1, n_steps=3, monitor_res_freq=100,
                            plot_freq=3, p_res=2, data_file_name="Mesh", average_fields="5 rho_average u_average v_average w_average e_average",
                            device_fused=1 if kernels == "fused" else 0)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[who] = split_vtu(d / "Mesh_000000003.vtu")
    (ta, na), (tb, nb) = out["ref"], out["ours"]
    assert ta == tb and na.shape == nb.shape
    assert "Name=\"u_average\"" in ta
    assert np.abs(na - nb).max() <= 1e-11 * np.abs(na).max()


def restart_numbers(path):
    """structure (all non-numeric lines, in order) and numbers of an ASCII restart file"""
    text, nums = [], []
    for line in open(path):
        t = line.split()
        try:
            nums.extend(float(x) for x in t)
        except ValueError:
            text.append(line.strip())
    return text, np.array(nums)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["hex", "pritet"])
def test_restart_file_matches_reference_binary(tmp_path, hb, meshgen, kind):
    if not (os.path.exists(REF) and os.path.exists(OURS)):
        pytest.skip("driver binaries not built")
    out = {}
    for who, exe in (("ref", REF), ("ours", OURS)):
        d = tmp_path / who
        d.mkdir()
        if kind == "hex":
            meshgen.hex_box(str(d / "m.neu"), 3)
        else:
            meshgen.mixed_box_3d(str(d / "m.neu"), (2, 4, 2), kind=kind)
        # staged kernels (device_fused 0): they reproduce the reference's arithmetic, so the files agree to the last printed digit
        meshgen.write_input(str(d / "input"), "m.neu", order=2, adv_type=2, dt=1e-5, riemann_solve_type=2, viscous=1, n_steps=2, monitor_res_freq=1,
                            restart_dump_freq=2, device_fused=0)
        env = dict(os.environ, HIFILES_HOME=util.REF_DIR if who == "ref" else os.path.join(util.ROOT, "hifiles-solver_b200"))
        r = subprocess.run([exe, "input"], cwd=str(d), env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        out[who] = restart_numbers(str(d / "Rest_000000002_p0000.dat"))
    assert out["ours"][0] == out["ref"][0]
    assert out["ours"][1].shape == out["ref"][1].shape
    scale = np.abs(out["ref"][1]).max()
    assert np.abs(out["ours"][1] - out["ref"][1]).max() <= 1e-13 * scale
