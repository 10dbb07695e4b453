"""CPU: the numpy restatement (oracle/hifiles_oracle.py) is pinned against golden dumps of the unmodified reference
solver (tests/golden/*.npz, see make_golden.py).  Tolerance 1e-12 relative (FP64; the restatement uses BLAS products, so
sums are not in the reference's order)."""
import glob
import os
import sys

import numpy as np
import pytest

import util

sys.path.insert(0, os.path.join(util.ROOT, "oracle"))
import hifiles_oracle as ho  # noqa: E402

GOLDEN = sorted(glob.glob(os.path.join(util.ROOT, "tests", "golden", "*.npz")))


def load(path):
    z = np.load(path)
    return {k.replace("__", "."): z[k] for k in z.files}


def make_oracle(g):
    kind = ["hex", "quad", "tri", "tet", "pri"][int(g["case.kind"][0])]
    inter = {"hex": ["int_quad"], "quad": ["int_seg"], "tri": ["int_seg"], "tet": ["int_tri"], "pri": ["int_tri", "int_quad"]}[kind]
    n_dims, _, order, viscous, riemann, adv = [int(x) for x in g["meta"]]
    p = g["params"]
    P = ho.Params(gamma=p[0], prandtl=p[1], mu_inf=p[2], rt_inf=p[3], c_sth=p[4], fix_vis=p[5], ldg_beta=p[6], ldg_tau=p[7], dt=p[8],
                  viscous=viscous, riemann_solve_type=riemann, adv_type=adv, RK_a=g.get("rk_a"), RK_b=g.get("rk_b"))
    return ho.Oracle(g, kind, inter, P), kind


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_restatement_matches_reference_dump(path):
    g = load(path)
    orc, kind = make_oracle(g)
    orc.calc_residual()
    # the pointwise residual is a derivative (cancellation): judged at 1e-11, the states at 1e-12
    assert util.rel_err(orc.div, g["step0.stage0.s18_corrected_divergence.%s.div_tconf_upts" % kind]) < 1e-11
    orc.advance_solution(0)
    assert util.rel_err(orc.u[0], g["step0.stage0.advanced.%s.disu_upts" % kind]) < 1e-12
    for s in range(1, orc.n_stages()):
        orc.calc_residual()
        orc.advance_solution(s)
    steps = int(g["case.steps"][0])
    for _ in range(steps - 1):
        orc.step()
    assert util.rel_err(orc.u[0], g["final.%s.disu_upts" % kind]) < 1e-12
    assert util.rel_err(orc.norm_residual(1), g["history.norm_residual"][:, -1]) < 1e-11


def test_riemann_solvers_are_consistent():
    """F(u,u,n) = F(u).n for every solver (a property the reference's solvers share), on random admissible states."""
    rng = np.random.default_rng(7)
    u = np.empty((64, 5))
    u[:, 0] = 1 + rng.random(64)
    u[:, 1:4] = rng.standard_normal((64, 3))
    u[:, 4] = 10 + rng.random(64)
    n = rng.standard_normal((64, 3))
    n /= np.linalg.norm(n, axis=1)[:, None]
    exact = np.einsum("qkd,qd->qk", ho.calc_invf(u, 1.4), n)
    for f in (ho.rusanov_flux, ho.hllc_flux, ho.roeM_flux):
        assert np.allclose(f(u, u.copy(), n, 1.4), exact, rtol=1e-13, atol=1e-13)


def test_ldg_switch_is_antisymmetric():
    """The two sides of a partition face evaluate the switch with opposite normals and must pick opposite signs
    (reference src/inters.cpp:566-581)."""
    rng = np.random.default_rng(3)
    n = rng.standard_normal((200, 3))
    n[:50, 0] = 0.
    n[50:80, 1] = -n[50:80, 0]
    n[:20, 1] = 0.
    keep = ~((n[:, 0] == 0) & (n[:, 0] + n[:, 1] == 0) & (n[:, 0] + n[:, 2] == 0))
    a, b = ho.ldg_switched_beta(0.5, n[keep]), ho.ldg_switched_beta(0.5, -n[keep])
    assert np.all(a == -b)


@pytest.mark.parametrize("name", ["hex2_p3_ns_rusanov_rk45", "hex3_p2_ns_hllc_rk34", "hex3_p1_ns_roem_sutherland_rk24"])
def test_face_gradient_formulation_prototype(name):
    """tools/face_gradient_proto.py: the planned k_grad formulation (normal face gradient from the line in registers,
    tangential ones as in-face derivatives of the face values + edge corrections) reproduces the reference's gradient at the
    flux points (dense operators and golden dump) -- the math behind the next kernel generation, pinned on the CPU."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("face_gradient_proto", os.path.join(util.ROOT, "tools", "face_gradient_proto.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(os.path.join(util.ROOT, "tests", "golden", name + ".npz")) == 0
