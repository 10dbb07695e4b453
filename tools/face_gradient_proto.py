"""Prototype (numpy) of the round-2 formulation of k_grad: the reference-space gradient at the flux points of a hexahedron's
faces WITHOUT the volume gradient planes.

Reference chain (src/eles.cpp:1823-2052): grad_upts(d) = opp_4(d) u + opp_5(d) delta;  grad_fpts(d) = opp_6 grad_upts(d).
With tensor-product operators (1-D tables D, l(-1), l(+1), correction c5 of the two faces of a direction) the same numbers are,
for a flux point q of a face F normal to direction n on side s (L = l(s)), with tangential directions t:
  normal      G_n(q) = sum_j (L.D)[j] u(j along the line behind q) + (L.c5[n+]) delta_{n+}(q') + (L.c5[n-]) delta_{n-}(q'')
  tangential  G_t(q) = sum_j D[q_t][j] uF(j along t inside the face) + c5[t+][q_t] Delta_{t+} + c5[t-][q_t] Delta_{t-},
              uF = the element's own face values on F, Delta_{t+-} = sum_i L[i] delta_{t+-}(i along n, rest as q): the side
              face's LDG correction extrapolated to the edge it shares with F.
So a thread that holds a line in registers gets the normal component for free, and the tangential ones are 5-wide line
operators inside the face (P+1 loads -> P+1 outputs) instead of a gather of (P+1) x 15 values per flux point.
Checked here against the dense reference operators and the golden dump of the unmodified reference (tests/golden).
usage: python tools/face_gradient_proto.py [tests/golden/hex2_p3_ns_rusanov_rk45.npz]"""
import sys
import numpy as np


def fpt_of_upt(N, f, a, b, c):
    P = N - 1  # face-local flux point met by the line through solution point (a, b, c), reference src/eles_hexas.cpp:224-282
    return [(P - a) + N * b, a + N * c, b + N * c, (P - a) + N * c, (P - b) + N * c, a + N * b][f]


FACE_DIR = [2, 1, 0, 1, 0, 2]   # direction of the normal of face f
FACE_SGN = [-1, -1, 1, 1, -1, 1]
FACES_OF_DIR = {0: (4, 2), 1: (1, 3), 2: (0, 5)}  # (minus, plus)


def main(path):
    z = np.load(path)
    g = {k.replace("__", "."): z[k] for k in z.files}
    o0, o6 = g["hex.opp_0"], g["hex.opp_6"]
    o4 = [g["hex.opp_4_%d" % d] for d in range(3)]
    o5 = [g["hex.opp_5_%d" % d] for d in range(3)]
    u = g["hex.disu_upts_ic"]                                                    # (upt, ele, field)
    delta = g["step0.stage0.s09_common_invFlux.hex.delta_disu_fpts"]            # (fpt, ele, field)
    NU = u.shape[0]
    N = round(NU ** (1 / 3))
    NN = N * N
    upt = lambda a, b, c: a + N * b + NN * c
    # 1-D tables out of the dense operators (as hf_fused.cu extract_tables does)
    Lm = np.array([o0[4 * NN + fpt_of_upt(N, 4, i, 0, 0), upt(i, 0, 0)] for i in range(N)])
    Lp = np.array([o0[2 * NN + fpt_of_upt(N, 2, i, 0, 0), upt(i, 0, 0)] for i in range(N)])
    D = np.array([[o4[0][upt(i, 0, 0), upt(j, 0, 0)] for j in range(N)] for i in range(N)])
    c5 = np.zeros((6, N))
    for f in range(6):
        d = FACE_DIR[f]
        for m in range(N):
            abc = [0, 0, 0]; abc[d] = m
            c5[f, m] = o5[d][upt(*abc), f * NN + fpt_of_upt(N, f, *abc)]
    L = {-1: Lm, 1: Lp}
    # reference: reference-space gradient at the flux points, dense operators
    ne, nf = u.shape[1], u.shape[2]
    ref = np.zeros((6 * NN, ne, nf, 3))
    for d in range(3):
        gu = np.einsum("pq,qef->pef", o4[d], u) + np.einsum("pq,qef->pef", o5[d], delta)
        ref[..., d] = np.einsum("pq,qef->pef", o6, gu)
    # new formulation
    uF = np.einsum("pq,qef->pef", o0, u)                                        # own face values (fpt, ele, field)
    new = np.zeros_like(ref)
    for f in range(6):
        n, s = FACE_DIR[f], FACE_SGN[f]
        Ls = L[s]
        LD = Ls @ D                                                              # (L.D)[j]
        tdirs = [d for d in range(3) if d != n]
        for a in range(N):
            for b in range(N):
                # the face point is addressed through the solution-point coordinates of its line: (a, b) = the two tangential indices
                abc = [0, 0, 0]; abc[tdirs[0]] = a; abc[tdirs[1]] = b
                q = f * NN + fpt_of_upt(N, f, *abc)
                # normal component: line behind the point
                line = []
                for j in range(N):
                    abc[n] = j
                    line.append(upt(*abc))
                fm, fp = FACES_OF_DIR[n]
                abc[n] = 0
                qm = fm * NN + fpt_of_upt(N, fm, *abc)
                qp = fp * NN + fpt_of_upt(N, fp, *abc)
                new[q, :, :, n] = np.einsum("j,jef->ef", LD, u[line]) + (Ls @ c5[fp]) * delta[qp] + (Ls @ c5[fm]) * delta[qm]
                # tangential components: in-face derivative of the face values + edge corrections of the side faces
                for t in tdirs:
                    other = [d for d in tdirs if d != t][0]
                    qt = abc[t]
                    acc = 0.
                    for j in range(N):
                        cc = list(abc); cc[t] = j
                        acc = acc + D[qt, j] * uF[f * NN + fpt_of_upt(N, f, *cc)]
                    tm, tp = FACES_OF_DIR[t]
                    for ft in (tm, tp):
                        edge = 0.
                        for i in range(N):                                        # side face's delta extrapolated along n to the shared edge
                            cc = list(abc); cc[n] = i; cc[t] = 0
                            edge = edge + Ls[i] * delta[ft * NN + fpt_of_upt(N, ft, *cc)]
                        acc = acc + c5[ft, qt] * edge
                    new[q, :, :, t] = acc
                    _ = other
    err = np.abs(new - ref).max() / np.abs(ref).max()
    print("face gradient, new formulation vs dense reference operators: max rel err %.3e" % err)
    # and against the reference's physical gradient at the flux points (golden dump): 1/detjac * JGinv^T
    J = g["hex.JGinv_fpts"]; dj = g["hex.detjac_fpts"]                          # (l, m, fpt, ele), (fpt, ele)
    phys = np.einsum("lmpe,pefl->pefm", J, new) / dj[:, :, None, None]
    gold = g["step0.stage0.s11_correct_gradient.hex.grad_disu_fpts"]
    err2 = np.abs(phys - gold).max() / np.abs(gold).max()
    print("physical gradient at the flux points vs the golden dump of the reference: max rel err %.3e" % err2)
    return 0 if err < 1e-12 and err2 < 1e-12 else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "tests/golden/hex2_p3_ns_rusanov_rk45.npz"))
