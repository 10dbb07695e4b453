"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's per-RK-stage residual path.  Never imported by the
product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use anything under oracle/.

Pinned: tests/test_oracle_cpu.py checks this file against golden dumps of the UNMODIFIED reference solver
(tests/golden/*.npz, produced by tests/golden/make_golden.py from oracle/_ref/ref_dump, i.e. the reference's own
sources compiled where they lie under /root/reference).  The reference's own regression goldens are stale
(SURVEY.md section 4), so reference runs are the pin.

Every function cites the reference routine it restates.  Arrays keep the reference's column-major hf_array index order
(first index fastest): disu_upts(upt,ele,field), disu_fpts(fpt,ele,field), grad(pt,ele,field,dim),
tdisf_upts(upt,ele,field,dim), JGinv(l,m,pt,ele), detjac(pt,ele), tdA(fpt,ele), norm_fpts(fpt,ele,dim), opp(row,col).
Scope: Euler / Navier-Stokes on one element type with interior (incl. periodic) interfaces -- what BASELINE configs
1 and 3 exercise -- and boundary interfaces (the eleven Navier-Stokes boundary kinds without ramp and wall model: SURVEY 8a rows a8, a14).  Operator products use numpy's BLAS, so sums are not in the reference's order: agreement is to
rounding (1e-13), not bit-exact."""
import numpy as np


class Params:
    def __init__(self, gamma, prandtl, mu_inf, rt_inf, c_sth, fix_vis, ldg_beta, ldg_tau, dt, viscous, riemann_solve_type, adv_type,
                 RK_a=None, RK_b=None):
        self.__dict__.update(locals())
        del self.__dict__["self"]


# ---- pointwise physics -------------------------------------------------------------------------------------------------
def calc_invf(u, gamma):
    """calc_invf_2d / calc_invf_3d (reference src/flux.cpp:33-125).  u[..., field] -> f[..., field, dim]"""
    nd = u.shape[-1] - 2
    rho, E = u[..., 0], u[..., nd + 1]
    v = u[..., 1:nd + 1] / rho[..., None]
    p = (gamma - 1.0) * (E - 0.5 * rho * np.sum(v * v, axis=-1))
    f = np.empty(u.shape + (nd,))
    for d in range(nd):
        f[..., 0, d] = u[..., d + 1]
        for k in range(nd):
            f[..., k + 1, d] = u[..., k + 1] * v[..., d] + (p if k == d else 0.0)
        f[..., nd + 1, d] = v[..., d] * (E + p)
    return f


def calc_visf(u, g, P):
    """calc_visf_2d / calc_visf_3d (reference src/flux.cpp:129-422), no RANS.  g[..., field, dim]"""
    nd = u.shape[-1] - 2
    rho, ene = u[..., 0], u[..., nd + 1]
    v = u[..., 1:nd + 1] / rho[..., None]
    vsq = np.sum(v * v, axis=-1)
    inte = ene / rho - 0.5 * vsq
    rt_ratio = (P.gamma - 1.0) * inte / P.rt_inf
    mu = P.mu_inf * rt_ratio ** 1.5 * (1 + P.c_sth) / (rt_ratio + P.c_sth)
    mu = mu + P.fix_vis * (P.mu_inf - mu)
    drho = g[..., 0, :]                                             # (..., dim)
    dv = (g[..., 1:nd + 1, :] - drho[..., None, :] * v[..., :, None]) / rho[..., None, None]  # dv[i, d]
    dke = 0.5 * vsq[..., None] * drho + rho[..., None] * np.einsum("...i,...id->...d", v, dv)
    de = (g[..., nd + 1, :] - dke - drho * inte[..., None]) / rho[..., None]
    diag = np.einsum("...ii->...", dv) / 3.0
    tau = mu[..., None, None] * (dv + np.swapaxes(dv, -1, -2))
    for i in range(nd):
        tau[..., i, i] = 2.0 * mu * (dv[..., i, i] - diag)
    f = np.zeros(u.shape + (nd,))
    f[..., 1:nd + 1, :] = -tau
    f[..., nd + 1, :] = -(np.einsum("...i,...id->...d", v, tau) + (mu / P.prandtl)[..., None] * P.gamma * de)
    return f


def _sides(u_l, u_r, n, gamma):
    nd = n.shape[-1]
    out = []
    for u in (u_l, u_r):
        v = u[..., 1:nd + 1] / u[..., :1]
        vn = np.sum(v * n, axis=-1)
        vsq = np.sum(v * v, axis=-1)
        p = (gamma - 1.0) * (u[..., nd + 1] - 0.5 * u[..., 0] * vsq)
        h = (u[..., nd + 1] + p) / u[..., 0]
        fn = np.einsum("...kd,...d->...k", calc_invf(u, gamma), n)
        out.append((v, vn, vsq, p, h, fn))
    return out


def rusanov_flux(u_l, u_r, n, gamma):
    """inters::rusanov_flux (reference src/inters.cpp:277-324)"""
    (_, vn_l, _, p_l, _, fn_l), (_, vn_r, _, p_r, _, fn_r) = _sides(u_l, u_r, n, gamma)
    eig = np.sqrt(gamma * (p_l + p_r) / (u_l[..., 0] + u_r[..., 0])) + 0.5 * np.abs(vn_l + vn_r)
    return 0.5 * ((fn_l + fn_r) - eig[..., None] * (u_r - u_l))


def hllc_flux(u_l, u_r, n, gamma):
    """inters::hllc_flux (reference src/inters.cpp:439-532)"""
    nd = n.shape[-1]
    (_, vn_l, _, p_l, h_l, fn_l), (_, vn_r, _, p_r, h_r, fn_r) = _sides(u_l, u_r, n, gamma)
    rl, rr = u_l[..., 0], u_r[..., 0]
    sq_rho = np.sqrt(rr / rl)
    rrho = 1. / (sq_rho + 1.)
    vn_m = rrho * (vn_l + sq_rho * vn_r)
    h_m = rrho * (h_l + sq_rho * h_r)
    a_m = np.sqrt((gamma - 1.) * (h_m - 0.5 * vn_m * vn_m))
    S_R, S_L = vn_m + a_m, vn_m - a_m
    S_star = (p_r - p_l + rl * vn_l * (S_L - vn_l) - rr * vn_r * (S_R - vn_r)) / (rl * (S_L - vn_l) - rr * (S_R - vn_r))

    def star(u, fn, p, vn, S):
        rcp = S - S_star
        pst = p + u[..., 0] * (S - vn) * (S_star - vn)
        out = np.empty_like(u)
        out[..., 0] = S_star * (S * u[..., 0] - fn[..., 0]) / rcp
        for i in range(nd):
            out[..., i + 1] = (S_star * (S * u[..., i + 1] - fn[..., i + 1]) + S * pst * n[..., i]) / rcp
        out[..., nd + 1] = (S_star * (S * u[..., nd + 1] - fn[..., nd + 1]) + S * pst * S_star) / rcp
        return out
    fl_star, fr_star = star(u_l, fn_l, p_l, vn_l, S_L), star(u_r, fn_r, p_r, vn_r, S_R)
    c0, c1, c2 = (S_L >= 0)[..., None], (S_star >= 0)[..., None], (S_R >= 0)[..., None]
    return np.where(c0, fn_l, np.where(c1, fl_star, np.where(c2, fr_star, fn_r)))


def roeM_flux(u_l, u_r, n, gamma):
    """inters::roeM_flux (reference src/inters.cpp:327-437)"""
    nd = n.shape[-1]
    (v_l, vn_l, _, p_l, h_l, fn_l), (v_r, vn_r, _, p_r, h_r, fn_r) = _sides(u_l, u_r, n, gamma)
    rl, rr = u_l[..., 0], u_r[..., 0]
    drho, dp, dh, dvn = rr - rl, p_r - p_l, h_r - h_l, vn_r - vn_l
    dv = v_r - v_l
    sq_rho = np.sqrt(rr / rl)
    rrho = 1.0 / (1.0 + sq_rho)
    ratr = sq_rho * rrho
    ra = sq_rho * rl
    ha = h_l * rrho + h_r * ratr
    va = v_l * rrho[..., None] + v_r * ratr[..., None]
    qq = np.sum(va * va, axis=-1)
    va_n = np.sum(n * va, axis=-1)
    aa = np.sqrt((gamma - 1) * (ha - 0.5 * qq))
    rcp_aa = 1.0 / aa
    abs_ma = np.abs(va_n * rcp_aa)
    b1 = np.maximum(0.0, np.maximum(va_n + aa, vn_r + aa))
    b2 = np.minimum(0.0, np.minimum(va_n - aa, vn_l - aa))
    b1b2 = b1 * b2
    rcp = 1.0 / (b1 - b2)
    b1, b2, b1b2 = b1 * rcp, b2 * rcp, b1b2 * rcp
    h = 1.0 - np.where(p_l < p_r, p_l / p_r, p_r / p_l)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = np.where(abs_ma != 0, abs_ma ** h, 1.)
    g = f / (1.0 + abs_ma)
    du = u_r - u_l
    du[..., nd + 1] = rr * h_r - rl * h_l
    bdq = np.empty_like(u_l)
    bdq[..., 0] = drho - f * dp * rcp_aa * rcp_aa
    bdq[..., nd + 1] = bdq[..., 0] * ha + ra * dh
    for i in range(nd):
        bdq[..., i + 1] = bdq[..., 0] * va[..., i] + ra * (dv[..., i] - n[..., i] * dvn)
    return (b1[..., None] * fn_l - b2[..., None] * fn_r) + b1b2[..., None] * (du - g[..., None] * bdq)


def ldg_switched_beta(beta, n):
    """the 'consistent switch' of inters::ldg_flux / ldg_solution (reference src/inters.cpp:566-581, 620-634)"""
    nd = n.shape[-1]
    b = np.full(n.shape[:-1], float(beta))
    if beta == 0.:
        return b
    n0 = n[..., 0]
    s01 = n[..., 0] + n[..., 1]
    flip = n0 < 0.
    z0 = n0 == 0.
    flip |= z0 & (s01 < 0.)
    if nd == 3:
        flip |= z0 & (s01 == 0) & ((n[..., 0] + n[..., 2]) < 0.)
    return np.where(flip, -b, b)



# ---- boundary interfaces (reference src/bdy_inters.cpp) -----------------------------------------------------------------------
SUB_IN_SIMP, SUB_OUT_SIMP, SUB_IN_CHAR, SUB_OUT_CHAR, SUP_IN, SUP_OUT, SLIP_WALL, CYCLIC, ISOTHERM_WALL, ADIABAT_WALL, CHAR, SLIP_WALL_DUAL, AD_WALL = range(13)
WALL_KINDS = (SLIP_WALL, ISOTHERM_WALL, ADIABAT_WALL, AD_WALL, SLIP_WALL_DUAL)


class BC:
    """one row of run_input.bc_list as set_boundary_conditions reads it (reference include/bc.h:50-58)"""
    def __init__(self, flag, row):
        self.flag = int(flag)
        self.rho, self.velocity, self.p_static, self.T_static, self.p_total, self.T_total, self.mach = row[0], row[1:4], row[4], row[5], row[6], row[7], row[8]
        self.n_free = row[9:12]
        self.use_wm = int(row[12])


def set_boundary_conditions(sol_spec, B, u_l, n, gamma, R_ref):
    """bdy_inters::set_boundary_conditions, Navier-Stokes branch (reference src/bdy_inters.cpp:340-1008), for the points of ONE boundary:
    u_l[q, field], n[q, dim] -> u_r[q, field].  sol_spec 0: inviscid ghost state, 1: viscous boundary solution.  No pressure ramp, no wall
    model, no turbulent inlet (the inputs this restatement is pinned on use none)."""
    nd = u_l.shape[-1] - 2
    gm1 = gamma - 1.0
    rho_l, e_l = u_l[:, 0], u_l[:, nd + 1]
    v_l = u_l[:, 1:nd + 1] / rho_l[:, None]
    p_l = gm1 * (e_l - 0.5 * rho_l * (v_l * v_l).sum(1))
    vn_l = (v_l * n).sum(1)
    vb = np.asarray(B.velocity[:nd], dtype=float)
    f = B.flag
    if f == SUB_IN_SIMP:                                  # :374-394 fixed density and velocity, pressure from inside
        rho_r = np.full_like(rho_l, B.rho)
        v_r = np.broadcast_to(vb, v_l.shape).copy()
        e_r = p_l / gm1 + 0.5 * rho_r * (v_r * v_r).sum(1)
    elif f == SUB_OUT_SIMP:                               # :399-466 fixed pressure; back flow and supersonic outflow branches
        machn = np.abs(vn_l) / np.sqrt(gamma * p_l / rho_l)
        back, sup = vn_l < 0, (vn_l >= 0) & (machn >= 1)
        v_b = vn_l[:, None] * n
        vsq_b = (v_b * v_b).sum(1)
        T_b = B.T_total - 0.5 * vsq_b * gm1 / (R_ref * gamma)
        p_b = B.p_static * (1.0 + 0.5 * gm1 * (vsq_b / (gamma * R_ref * T_b))) ** (-gamma / gm1)
        rho_b = p_b / (R_ref * T_b)
        e_b = p_b / gm1 + 0.5 * rho_b * vsq_b
        e_s = B.p_static / gm1 + 0.5 * rho_l * (v_l * v_l).sum(1)
        rho_r = np.where(back, rho_b, rho_l)
        v_r = np.where(back[:, None], v_b, v_l)
        e_r = np.where(back, e_b, np.where(sup, e_l, e_s))
    elif f == SUB_IN_CHAR:                                # :471-588 total pressure / temperature, Riemann invariant from inside
        nf = np.asarray(B.n_free[:nd], dtype=float)
        c_l = np.sqrt(gamma * p_l / rho_l)
        R_plus = vn_l + 2.0 * c_l / gm1
        c_tot = gamma * R_ref * B.T_total
        alpha = (n * nf).sum(1)
        aa = 1.0 + 0.5 * gm1 * alpha * alpha
        bb = -gm1 * alpha * R_plus
        cc = 0.5 * gm1 * R_plus * R_plus - 2.0 * c_tot / gm1
        dd = np.sqrt(np.maximum(bb * bb - 4.0 * aa * cc, 0.0))
        V = np.maximum((-bb + dd) / (2.0 * aa), 0.0)
        vsq = V * V
        c_r = c_tot - 0.5 * gm1 * vsq
        vsq = np.minimum(vsq / c_r, 1.0) * c_r
        V = np.sqrt(vsq)
        c_r = c_tot - 0.5 * gm1 * vsq
        v_r = V[:, None] * nf
        T_r = c_r / (gamma * R_ref)
        p_r = B.p_total * (T_r / B.T_total) ** (gamma / gm1)
        rho_r = p_r / (R_ref * T_r)
        e_r = p_r / gm1 + 0.5 * rho_r * vsq
    elif f == SUB_OUT_CHAR:                               # :593-640 fixed pressure, entropy and Riemann invariant from inside
        c_l = np.sqrt(gamma * p_l / rho_l)
        R_plus = vn_l + 2.0 * c_l / gm1
        ent = p_l / rho_l ** gamma
        p_r = B.p_static
        rho_r = (p_r / ent) ** (1.0 / gamma)
        c_r = np.sqrt(gamma * p_r / rho_r)
        vn_r = R_plus - 2.0 * c_r / gm1
        v_r = v_l + (vn_r - vn_l)[:, None] * n
        e_r = p_r / gm1 + 0.5 * rho_r * (v_r * v_r).sum(1)
    elif f == SUP_IN:                                     # :644-660
        rho_r = np.full_like(rho_l, B.rho)
        v_r = np.broadcast_to(vb, v_l.shape).copy()
        e_r = B.p_static / gm1 + 0.5 * rho_r * (v_r * v_r).sum(1)
    elif f == SUP_OUT:                                    # :664-670
        rho_r, v_r, e_r = rho_l, v_l, e_l
    elif f == SLIP_WALL:                                  # :674-701 mirrored (inviscid) / removed (viscous) normal velocity
        rho_r = rho_l
        v_r = v_l - (2.0 if sol_spec == 0 else 1.0) * vn_l[:, None] * n
        e_r = p_l / gm1 + 0.5 * rho_r * (v_r * v_r).sum(1)
    elif f in (ISOTHERM_WALL, ADIABAT_WALL):              # :705-863 (without wall model)
        rho_r = rho_l
        v_r = (2.0 * vb - v_l) if sol_spec == 0 else np.broadcast_to(vb, v_l.shape).copy()
        vsq = (v_r * v_r).sum(1)
        e_r = rho_r * (R_ref / gm1 * B.T_static) + 0.5 * rho_r * vsq if f == ISOTHERM_WALL else p_l / gm1 + 0.5 * rho_r * vsq
    elif f == CHAR:                                       # :867-972 Riemann-invariant far field
        vn_b = (vb * n).sum(1)
        c_l = np.sqrt(gamma * p_l / rho_l)
        c_b = np.sqrt(gamma * B.p_static / B.rho)
        sup = np.abs(vn_l) / c_l >= 1
        inflow = vn_l < 0
        r_plus = np.where(inflow & sup, vn_b + 2.0 / gm1 * c_b, vn_l + 2.0 / gm1 * c_l)
        r_minus = np.where(~inflow & sup, vn_l - 2.0 / gm1 * c_l, vn_b - 2.0 / gm1 * c_b)
        c_star = 0.25 * gm1 * (r_plus - r_minus)
        vn_star = 0.5 * (r_plus + r_minus)
        one_over_s = np.where(inflow, B.rho ** gamma / B.p_static, rho_l ** gamma / p_l)
        rho_r = (1.0 / gamma * (one_over_s * c_star * c_star)) ** (1.0 / gm1)
        v_t = np.where(inflow[:, None], vb - vn_b[:, None] * n, v_l - vn_l[:, None] * n)
        v_r = vn_star[:, None] * n + v_t
        p_r = rho_r / gamma * c_star * c_star
        e_r = p_r / gm1 + 0.5 * rho_r * (v_r * v_r).sum(1)
    elif f == SLIP_WALL_DUAL:                             # :976-995
        rho_r = rho_l
        v_r = v_l - 2.0 * vn_l[:, None] * n
        e_r = e_l
    else:
        raise ValueError("boundary kind %d is not restated" % f)
    return np.concatenate([rho_r[:, None], rho_r[:, None] * v_r, e_r[:, None]], axis=1)


def set_boundary_gradients(flag, u_r, g_l, n):
    """bdy_inters::set_boundary_gradients (reference src/bdy_inters.cpp:1138-1189): g[q, field, dim]"""
    nd = g_l.shape[-1]
    if flag in (CHAR, SUP_IN, SUB_IN_SIMP, SUB_OUT_SIMP):
        return np.zeros_like(g_l)
    g = g_l.copy()
    if flag == ADIABAT_WALL:  # remove the wall-normal gradient of the internal energy
        rho = u_r[:, 0]
        vsq = (u_r[:, 1:nd + 1] ** 2).sum(1)
        inte = (u_r[:, nd + 1] - 0.5 * vsq / rho) / rho
        grad_vel = (g[:, 1:nd + 1, :] - g[:, 0:1, :] * (u_r[:, 1:nd + 1] / rho[:, None])[:, :, None]) / rho[:, None, None]
        s = inte[:, None] * g[:, 0, :] + (0.5 * vsq / (rho * rho))[:, None] * g[:, 0, :] + np.einsum("qi,qid->qd", u_r[:, 1:nd + 1], grad_vel)
        grad_inte = g[:, nd + 1, :] - s
        dn = (grad_inte * n).sum(1)
        g[:, nd + 1, :] -= dn[:, None] * n
    return g

# ---- the per-stage sequence (reference src/solver.cpp:50-223) for one element type, interior interfaces -------------------
class Oracle:
    def __init__(self, setup, prefix, inter_prefix, P):
        """setup: dict with the reference dump's names (oracle/ref_dump.cpp dump_setup): '<prefix>.opp_0', ... and
        '<inter_prefix>.idx_l' / '.idx_r' (flat fpt + n_fpts*ele of each flux-point pair, right side already permuted by
        inters::get_lut, reference src/int_inters.cpp:76-95)."""
        g = lambda k: np.asarray(setup[prefix + "." + k])
        self.P = P
        self.ne, self.nu, self.nfp, self.nf, self.nd = [int(x) for x in g("sizes")[:5]]
        nd = self.nd
        self.opp_0, self.opp_3 = g("opp_0"), g("opp_3")
        self.opp_1 = [g("opp_1_%d" % d) for d in range(nd)]
        self.opp_2 = [g("opp_2_%d" % d) for d in range(nd)]
        if P.viscous:
            self.opp_4 = [g("opp_4_%d" % d) for d in range(nd)]
            self.opp_5 = [g("opp_5_%d" % d) for d in range(nd)]
            self.opp_6 = g("opp_6")
        self.detjac_upts, self.JGinv_upts = g("detjac_upts"), g("JGinv_upts")
        self.detjac_fpts, self.JGinv_fpts = g("detjac_fpts"), g("JGinv_fpts")
        self.tdA_fpts, self.norm_fpts = g("tdA_fpts"), g("norm_fpts")
        prefixes = [inter_prefix] if isinstance(inter_prefix, str) else list(inter_prefix)  # a prism has triangular and quadrilateral faces
        self.idx_l = np.concatenate([np.asarray(setup[q + ".idx_l"]).ravel(order="F") for q in prefixes])
        self.idx_r = np.concatenate([np.asarray(setup[q + ".idx_r"]).ravel(order="F") for q in prefixes])
        self.u = [np.array(g("disu_upts_ic"), order="F"), np.zeros((self.nu, self.ne, self.nf), order="F")]
        self.div = np.zeros((self.nu, self.ne, self.nf), order="F")
        # boundary interfaces (oracle/ref_dump.cpp dump_setup: 'bdy_<face type>.idx_l', '.boundary_id', '.bc_flags', '.bc_params', '.R_ref')
        self.bdy = []  # (flat flux-point indices of one boundary's faces, BC)
        for q in [x.replace("int_", "bdy_") for x in prefixes]:
            if q + ".idx_l" not in setup:
                continue
            idx = np.asarray(setup[q + ".idx_l"])          # (fpt, inter)
            bid = np.asarray(setup[q + ".boundary_id"]).ravel()
            flags, rows = np.asarray(setup[q + ".bc_flags"]).ravel(), np.asarray(setup[q + ".bc_params"]).reshape((13, -1), order="F")
            self.R_ref = float(np.asarray(setup[q + ".R_ref"]).ravel()[0])
            for b in np.unique(bid):
                self.bdy.append((idx[:, bid == b].ravel(order="F"), BC(flags[b], rows[:, b])))

    # operator product over all (ele, field) columns: the reference's column-major dgemm (src/funcs.cpp:49-124)
    @staticmethod
    def _apply(op, x):
        return np.einsum("rc,cef->ref", op, x)

    def _flat(self, a):  # (fpt, ele, field[...]) -> (fpt*ele, field[...]) in flat (fpt + n_fpts*ele) order
        return a.reshape((self.nfp * self.ne,) + a.shape[2:], order="F")

    def calc_residual(self):
        P, nd = self.P, self.nd
        u = self.u[0]
        disu_fpts = self._apply(self.opp_0, u)                                   # eles::extrapolate_solution :1360
        if P.viscous:
            grad_upts = np.stack([self._apply(self.opp_4[d], u) for d in range(nd)], axis=-1)   # calculate_gradient :1823
        # eles::evaluate_invFlux (:1415): tdisf(j,i,k,l) = sum_m JGinv(l,m,j,i) f(k,m)
        f = calc_invf(u, P.gamma)
        tdisf = np.einsum("lmje,jekm->jekl", self.JGinv_upts, f)
        # int_inters::calculate_common_invFlux (src/int_inters.cpp:160-249)
        uf = self._flat(disu_fpts)
        nrm = self._flat(self.norm_fpts)
        tdA = self.tdA_fpts.ravel(order="F")
        u_l, u_r, n = uf[self.idx_l], uf[self.idx_r], nrm[self.idx_l]
        fn = {0: rusanov_flux, 2: roeM_flux, 3: hllc_flux}[P.riemann_solve_type](u_l, u_r, n, P.gamma)
        ntconf = np.zeros_like(uf)
        ntconf[self.idx_l] = fn * tdA[self.idx_l][:, None]
        ntconf[self.idx_r] = -fn * tdA[self.idx_r][:, None]
        riemann = {0: rusanov_flux, 2: roeM_flux, 3: hllc_flux}[P.riemann_solve_type]
        # bdy_inters::evaluate_boundaryConditions_invFlux (src/bdy_inters.cpp:213-338)
        for bi, B in self.bdy:
            ub_l, nb = uf[bi], nrm[bi]
            ub_r = set_boundary_conditions(0, B, ub_l, nb, P.gamma, self.R_ref)
            fnb = np.einsum("qkd,qd->qk", calc_invf(ub_l, P.gamma), nb) if B.flag == SLIP_WALL_DUAL else riemann(ub_l, ub_r, nb, P.gamma)
            ntconf[bi] = fnb * tdA[bi][:, None]
        if P.viscous:
            beta = ldg_switched_beta(P.ldg_beta, n)[:, None]
            u_c = 0.5 * (u_l + u_r) - beta * (u_l - u_r)                          # inters::ldg_solution :615
            delta = np.zeros_like(uf)
            delta[self.idx_l] = u_c - u_l
            delta[self.idx_r] = u_c - u_r
            for bi, B in self.bdy:                                                # boundary: u_c = u_r (ldg_solution, flux_spec 1)
                ub_l, nb = uf[bi], nrm[bi]
                delta[bi] = set_boundary_conditions(1 if B.flag in WALL_KINDS else 0, B, ub_l, nb, P.gamma, self.R_ref) - ub_l
            delta = delta.reshape((self.nfp, self.ne, self.nf), order="F")
            # eles::correct_gradient (:1890-2052)
            for d in range(nd):
                grad_upts[..., d] += self._apply(self.opp_5[d], delta)
            grad_fpts = np.stack([self._apply(self.opp_6, grad_upts[..., d]) for d in range(nd)], axis=-1)
            grad_upts = np.einsum("ldje,jekl->jekd", self.JGinv_upts, grad_upts) / self.detjac_upts[:, :, None, None]
            grad_fpts = np.einsum("ldje,jekl->jekd", self.JGinv_fpts, grad_fpts) / self.detjac_fpts[:, :, None, None]
            # eles::evaluate_viscFlux (:2285)
            tdisf = tdisf + np.einsum("lmje,jekm->jekl", self.JGinv_upts, calc_visf(u, grad_upts, P))
        # extrapolate_totalFlux (:1549), calculate_divergence (:1651)
        ntdisf = sum(self._apply(self.opp_1[d], tdisf[..., d]) for d in range(nd))
        div = sum(self._apply(self.opp_2[d], tdisf[..., d]) for d in range(nd))
        if P.viscous:
            # int_inters::calculate_common_viscFlux (src/int_inters.cpp:254-343) + inters::ldg_flux (:561-611)
            gf = self._flat(grad_fpts)
            f_l, f_r = calc_visf(u_l, gf[self.idx_l], P), calc_visf(u_r, gf[self.idx_r], P)
            b = ldg_switched_beta(P.ldg_beta, n)[:, None, None]
            f_c = (0.5 + b) * f_l + (0.5 - b) * f_r
            fnv = np.einsum("qkd,qd->qk", f_c, n) - P.ldg_tau * (u_r - u_l)
            ntconf[self.idx_l] += fnv * tdA[self.idx_l][:, None]
            ntconf[self.idx_r] += -fnv * tdA[self.idx_r][:, None]
            # bdy_inters::evaluate_boundaryConditions_viscFlux (src/bdy_inters.cpp:1024-1090): f_c = f_r (ldg_flux, flux_spec 1)
            for bi, B in self.bdy:
                if B.flag == SLIP_WALL:
                    continue
                ub_l, nb = uf[bi], nrm[bi]
                ub_r = set_boundary_conditions(1, B, ub_l, nb, P.gamma, self.R_ref)
                g_r = set_boundary_gradients(B.flag, ub_r, gf[bi], nb)
                fnb = np.einsum("qkd,qd->qk", calc_visf(ub_r, g_r, P), nb) - P.ldg_tau * (ub_r - ub_l)
                ntconf[bi] += fnb * tdA[bi][:, None]
        ntconf = ntconf.reshape((self.nfp, self.ne, self.nf), order="F")
        # calculate_corrected_divergence (:1738)
        self.div = div + self._apply(self.opp_3, ntconf - ntdisf)
        return self.div

    def advance_solution(self, stage):
        """eles::AdvanceSolution (reference src/eles.cpp:1080-1265), fixed time step, no source term"""
        P = self.P
        res = self.div / self.detjac_upts[:, :, None]
        rhs = -res
        u0, u1 = self.u
        a = P.adv_type
        if a == 0:
            u0 -= P.dt * res
        elif a == 1:
            if stage == 0:
                u1[...] = u0
            if stage < 3:
                u0 -= P.dt / 3.0 * res
            else:
                u0[...] = 3.0 / 4.0 * u0 + 1.0 / 4.0 * u1 + P.dt / 4.0 * rhs
        elif a == 2:
            if stage == 0:
                u1[...] = u0
            if stage < 2 or stage == 3:
                u0 -= P.dt / 2.0 * res
            else:
                u0[...] = 1.0 / 3.0 * u0 + 2.0 / 3.0 * u1 + P.dt / 6.0 * rhs
        else:
            u1[...] = P.RK_a[stage] * u1 + P.dt * rhs
            u0 += P.RK_b[stage] * u1

    def n_stages(self):
        return {0: 1, 1: 4, 2: 4, 3: 5, 4: 14}[self.P.adv_type]

    def step(self):
        for s in range(self.n_stages()):
            self.calc_residual()
            self.advance_solution(s)

    def norm_residual(self, norm_type=1):
        """eles::compute_res_upts + output::CalcNormResidual (reference src/eles.cpp:5045-5074, src/output.cpp:2166-2248)"""
        r = self.div / self.detjac_upts[:, :, None]
        n = self.nu * self.ne
        if norm_type == 0:
            return np.abs(r).max(axis=(0, 1))
        if norm_type == 1:
            return np.abs(r).sum(axis=(0, 1)) / n
        return np.sqrt((r * r).sum(axis=(0, 1))) / n
