// Fused residual path for affine hexahedra (BASELINE config 3: TGV, hex, P = 4).
//
// One RK stage = CalcResidual + AdvanceSolution (reference src/solver.cpp:50-223, src/eles.cpp:1080-1265) is done by
// two kernels per stage, each working element by element out of shared memory, with the operators applied in
// sum-factorised form (hex operators are tensor products with exact zeros, reference src/eles_hexas.cpp:224-282,
// 1132-1193, 1444-1537; the 1-D tables are read out of the dense matrices the host built, so they are the
// reference's numbers):
//
//   k_grad   (viscous only)   u -> own face values, LDG common solution with the neighbour's face values, corrected
//                             physical gradient at solution points, extrapolated to the faces, and there the element's
//                             own viscous flux dotted with the face's LEFT normal -> fv (4 values per flux point: the
//                             LDG common flux is linear in the two one-sided fluxes, so a side only ever needs
//                             F_vis(u,grad u).n of its neighbour, not the 15 gradient components)
//                             = eles::calculate_gradient + the LDG half of int_inters::calculate_common_invFlux +
//                               eles::correct_gradient + the one-sided half of calculate_common_viscFlux
//   k_resid                   the same gradient again (cheaper than storing it), inviscid + viscous flux at solution
//                             points, common fluxes on all six faces (Riemann + LDG, both sides of a face evaluate the
//                             same expression with the LEFT element's normal), divergence + correction, RK update,
//                             new face values -> fu[next]
//                             = evaluate_invFlux + evaluate_viscFlux + calculate_common_invFlux/viscFlux +
//                               extrapolate_totalFlux + calculate_divergence + calculate_corrected_divergence +
//                               AdvanceSolution + extrapolate_solution of the next stage
//
// Only face data crosses kernels: fu = u at flux points (5 fields), fv = one-sided viscous normal flux (4 fields),
// stored per face block [ele][face][field][fpt] so that a neighbour's face is one contiguous run and a partition
// face's receive buffer is simply an extra block behind the last element (message layout of fu = the reference's
// out_buffer_disu[inter][field][fpt], src/mpi_inters.cpp:226-229).
// Algorithmic traffic per element-stage, P = 4, RK34: see DESIGN.md (about 9 k doubles vs 56 k for the staged path).
#include "hf_device.h"
#include "hf_bc.cuh"
#include <cstring>
#include <type_traits>
#include <cstdlib>
#include <cmath>
#include <algorithm>
#include <map>
#include <array>

#define HF_FAIL(msg)   \
  do {                 \
    hf_set_error(msg); \
    return 1;          \
  } while (0)

namespace
{
constexpr int NF = 5;
constexpr int ND = 3;
constexpr int EM = 34; // doubles of per-element metrics: J[9], detjac, 6 x (tdA, nL[3])

// face -> direction of its normal, sign of the reference normal (reference src/eles_hexas.cpp:525-580)
__host__ __device__ inline int face_dir(int f) { return f == 0 || f == 5 ? 2 : (f == 1 || f == 3 ? 1 : 0); }
__host__ __device__ inline int face_sgn(int f) { return (f == 2 || f == 3 || f == 5) ? 1 : -1; }

// solution point (a,b,c) <-> face-local flux point on face f (reference src/eles_hexas.cpp:224-282)
template <int N>
__host__ __device__ inline int fpt_of_upt(int f, int a, int b, int c)
{
  constexpr int P = N - 1;
  switch (f)
  {
  case 0: return (P - a) + N * b;
  case 1: return a + N * c;
  case 2: return b + N * c;
  case 3: return (P - a) + N * c;
  case 4: return (P - b) + N * c;
  default: return a + N * b;
  }
}
// first solution point of the line behind face-local flux point j of face f, and the line's stride
template <int N>
__host__ __device__ inline void line_of_fpt(int f, int j, int &base, int &stride)
{
  constexpr int P = N - 1;
  int k = j % N, jj = j / N;
  switch (f)
  {
  case 0: base = (P - k) + N * jj; stride = N * N; break;
  case 1: base = k + N * N * jj; stride = N; break;
  case 2: base = N * k + N * N * jj; stride = 1; break;
  case 3: base = (P - k) + N * N * jj; stride = N; break;
  case 4: base = N * (P - k) + N * N * jj; stride = 1; break;
  default: base = k + N * jj; stride = N * N; break;
  }
}

struct fused_tables // small per-order tables, copied to shared memory by every block
{
  double D[36];      // D[i*N+j] = d l_j / dxi at xi_i           (from opp_2 / opp_4)
  double Lm[6], Lp[6]; // l_i(-1), l_i(+1)                        (from opp_0)
  double c3[36];     // c3[f*N+m]: opp_3 entry of face f at directional index m
  double c5[36];     // c5[f*N+m]: opp_5(dir f) entry
  unsigned short lbase[216];  // [f*N*N + j]: first solution point of the line behind flux point j of face f
  unsigned char perm[8 * 36]; // [rot + 4*is_right][j] -> neighbour's face-local flux point
  unsigned char uptab[216][8]; // per solution point: the face-local flux point (with face offset) it meets on faces 0..5, a | b<<4, c
};

struct rk_args
{
  int mode, copy_u1; // as k_rk_update in hf_device.cu
  double dt, fac, c1, c2;
  double dt_fac;     // dt / fac, formed once on the host (global time step)
};

struct fused_args
{
  int n_eles;
  const int *elist;       // element order of this launch: interior elements first, partition-adjacent ones last
  int lo, hi;             // the launch covers elist[lo .. hi)
  const double *u0;
  double *u0_out;
  double *u1;
  double *div;            // written when keep_residual
  double *grad_out;       // grad_disu_upts (upt,ele,field,dim): physical gradient at solution points, written when non-null (integral diagnostics)
  const double *fu_cur;   // face u, read (neighbours)
  double *fu_next;        // face u of the updated solution, written (own faces)
  double *fv;             // one-sided viscous normal flux at flux points (4 per point): written by k_grad, read by k_resid;
                          // generation 7: the complete common normal flux fc (5 per point) at the owned faces
  const double *em;       // [ele][EM]: JGinv[9], 1/detjac, 6 x (tdA, left normal[3])
  const int *nbr;         // [ele][6] neighbour face block
  const int *finfo;       // [ele][6] rot + 4*is_right + 8*partition face + 16*boundary face (then bits 8.. = index into the boundary table)
  const unsigned long long *bmask; // [ele][6] per face: bit j clear = own LDG weight 0.5 + beta, set = 0.5 - beta (see hf_fused_prepare)
  const double *dt_local;
  const unsigned *wait_flag; // generation 9, one launch per kernel: CTAs from position wait_from on (the partition-adjacent elements) wait until
  unsigned wait_value;       // *wait_flag >= wait_value, i.e. until the exchange they read from has landed
  int wait_from;
  int *nan_flag;          // raised (1 + element) when the residual of a point is NaN (reference src/eles.cpp:1781-1795)
  // 1-D operator tables of the run's order; kernel parameters live in the constant bank, so the unrolled line passes use
  // them as immediate constant operands (no registers, no shared memory)
  double tD[36];          // D[i*N+j] = d l_j / dxi at xi_i
  double tL[2][6];        // [0]: l_i(-1), [1]: l_i(+1)
  double tc3[36], tc5[36]; // opp_3 / opp_5 entry of face f at directional index m: [f*N+m]
  double tLD[2][6];       // generation 9: (l(s) . D)[j], the face-normal derivative of a line's face value
  double c5s[2][6];       // generation 9: opp_5 entries of a minus [0] / plus [1] face at directional index m (equal for the three directions, checked at setup)
  double lc5s[2][2];      // generation 9: [face side][s] = l(s) . c5s[face side]
  double *send_u, *send_g; // generation 9: send buffers of the partition faces ([interface][FB]); the kernels fill them directly (no pack kernel)
  double *gn;             // generation 9: [ele][face][FB] face-normal derivative of the own polynomial at the owned flux points
  const unsigned *cls9;   // generation 9: [ele] element class | face kinds << 20 (elements with equal owner masks and face info share a class)
  const uint2 *tw9;       // generation 9: [class][direction][task] packed line task of every thread (order of tools/bank_layout.py)
  const unsigned *cl9;    // generation 9: [class][CL9_WORDS] pass lists of k_face9
  const int *nidx;        // [ele][NFP] index into fu (field 0) of the neighbour's value facing each own flux point
  const double *nlf;      // RoeM only (else null): [ele][face][flux point][3] the LEFT side's normal at exactly this flux point, see hf_fused_prepare
  hf_phys P;
  rk_args rk;
  int viscous, keep_residual, do_update;
  int pf_dist;            // L2 software-prefetch distance in CTAs (0 = off)
  unsigned long long own_xor; // generation 7: bmask ^ own_xor = the flux points this element owns (LDG weight 1)
  const hf_bc *bct;       // generation 9 with boundary faces: the boundary table (finfo >> 8 of a boundary face indexes it)
  double R_ref;
  const unsigned char *bskip; // [ele] 1 = the element has a boundary face (k_face9's plain launch leaves it to the boundary launch)
  int bdy_mode;           // k_face9 variant: 0 no boundary faces, 1 the elements with a boundary face (elist), 2 all others
};

__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <int N>
__device__ __forceinline__ int face_stride(int f)
{
  return (f == 0 || f == 5) ? N * N : ((f == 1 || f == 3) ? N : 1);
}

// ---- pointwise physics of the fused path ------------------------------------------------------------------------------
// Same formulas as hf_physics.cuh (= reference src/flux.cpp, src/inters.cpp) with the divisions by rho replaced by one
// reciprocal per state and pow(x,1.5) by x*sqrt(x): results differ from the reference in the last bits only.
__device__ __forceinline__ void inv_flux_fast(const double *__restrict__ u, double *__restrict__ f, double gm1)
{
  const double ir = 1.0 / u[0];
  const double v0 = u[1] * ir, v1 = u[2] * ir, v2 = u[3] * ir;
  const double p = gm1 * (u[4] - 0.5 * u[0] * (v0 * v0 + v1 * v1 + v2 * v2));
  const double ep = u[4] + p;
  f[0] = u[1];       f[1] = p + u[1] * v0;  f[2] = u[2] * v0;      f[3] = u[3] * v0;      f[4] = v0 * ep;
  f[5] = u[2];       f[6] = u[1] * v1;      f[7] = p + u[2] * v1;  f[8] = u[3] * v1;      f[9] = v1 * ep;
  f[10] = u[3];      f[11] = u[1] * v2;     f[12] = u[2] * v2;     f[13] = p + u[3] * v2; f[14] = v2 * ep;
}

// viscous flux F(k,d) = f[k + 5 d] for k = 1..4 (the mass equation has none); g(k,d) = g[k + 5 d]
__device__ __forceinline__ void vis_flux_fast(const double *__restrict__ u, const double *__restrict__ g, double *__restrict__ f, const hf_phys &P)
{
  const double rho = u[0], ir = 1.0 / u[0];
  const double v0 = u[1] * ir, v1 = u[2] * ir, v2 = u[3] * ir;
  const double vsq = v0 * v0 + v1 * v1 + v2 * v2;
  const double inte = u[4] * ir - 0.5 * vsq;
  double mu = P.mu_inf;
  if (P.fix_vis != 1.0)
  {
    double rt = (P.gamma - 1.0) * inte / P.rt_inf;
    mu = P.mu_inf * (rt * sqrt(rt)) * (1. + P.c_sth) / (rt + P.c_sth);
    mu = mu + P.fix_vis * (P.mu_inf - mu);
  }
  double dv[3][3], de[3];
#pragma unroll
  for (int d = 0; d < 3; d++)
  {
    const double r_d = g[5 * d];
    dv[0][d] = (g[1 + 5 * d] - r_d * v0) * ir;
    dv[1][d] = (g[2 + 5 * d] - r_d * v1) * ir;
    dv[2][d] = (g[3 + 5 * d] - r_d * v2) * ir;
    const double dke = 0.5 * vsq * r_d + rho * (v0 * dv[0][d] + v1 * dv[1][d] + v2 * dv[2][d]);
    de[d] = (g[4 + 5 * d] - dke - r_d * inte) * ir;
  }
  const double diag = (dv[0][0] + dv[1][1] + dv[2][2]) * (1.0 / 3.0);
  const double txx = 2.0 * mu * (dv[0][0] - diag), tyy = 2.0 * mu * (dv[1][1] - diag), tzz = 2.0 * mu * (dv[2][2] - diag);
  const double txy = mu * (dv[0][1] + dv[1][0]), txz = mu * (dv[0][2] + dv[2][0]), tyz = mu * (dv[1][2] + dv[2][1]);
  const double kap = mu * P.gamma_over_pr; // (mu / Pr) * gamma
  f[0] = 0.;  f[1] = -txx; f[2] = -txy; f[3] = -txz; f[4] = -(v0 * txx + v1 * txy + v2 * txz + kap * de[0]);
  f[5] = 0.;  f[6] = -txy; f[7] = -tyy; f[8] = -tyz; f[9] = -(v0 * txy + v1 * tyy + v2 * tyz + kap * de[1]);
  f[10] = 0.; f[11] = -txz; f[12] = -tyz; f[13] = -tzz; f[14] = -(v0 * txz + v1 * tyz + v2 * tzz + kap * de[2]);
}

struct side_state
{
  double rho, ir, v[3], vn, vsq, p, h, fn[5];
};
__device__ __forceinline__ void make_side(const double *__restrict__ u, const double *__restrict__ n, double gm1, side_state &s)
{
  s.rho = u[0];
  s.ir = 1.0 / u[0];
  s.v[0] = u[1] * s.ir; s.v[1] = u[2] * s.ir; s.v[2] = u[3] * s.ir;
  s.vn = s.v[0] * n[0] + s.v[1] * n[1] + s.v[2] * n[2];
  s.vsq = s.v[0] * s.v[0] + s.v[1] * s.v[1] + s.v[2] * s.v[2];
  s.p = gm1 * (u[4] - 0.5 * u[0] * s.vsq);
  s.h = (u[4] + s.p) * s.ir;
  // normal inviscid flux F(u).n
  s.fn[0] = u[1] * n[0] + u[2] * n[1] + u[3] * n[2];
  s.fn[1] = u[1] * s.vn + s.p * n[0];
  s.fn[2] = u[2] * s.vn + s.p * n[1];
  s.fn[3] = u[3] * s.vn + s.p * n[2];
  s.fn[4] = s.vn * (u[4] + s.p);
}

// common inviscid normal flux (reference src/inters.cpp:277-532: rusanov_flux, roeM_flux, hllc_flux)
__device__ __forceinline__ void riemann_fast(const double *__restrict__ u_l, const double *__restrict__ u_r, const double *__restrict__ n,
                                             double *__restrict__ fn, const hf_phys &P)
{
  const double gamma = P.gamma, gm1 = P.gamma - 1.0;
  side_state L, R;
  make_side(u_l, n, gm1, L);
  make_side(u_r, n, gm1, R);
  if (P.riemann_solve_type == 3) // HLLC
  {
    const double sq_rho = sqrt(R.rho * L.ir);
    const double rrho = 1. / (sq_rho + 1.);
    const double vn_m = rrho * (L.vn + sq_rho * R.vn);
    const double h_m = rrho * (L.h + sq_rho * R.h);
    const double a_m = sqrt(gm1 * (h_m - 0.5 * vn_m * vn_m));
    const double S_R = vn_m + a_m, S_L = vn_m - a_m;
    const double ml = L.rho * (S_L - L.vn), mr = R.rho * (S_R - R.vn);
    const double S_star = (R.p - L.p + ml * L.vn - mr * R.vn) / (ml - mr);
    if (S_L >= 0)
    {
#pragma unroll
      for (int k = 0; k < 5; k++) fn[k] = L.fn[k];
    }
    else if (S_star >= 0)
    {
      const double inv = 1.0 / (S_L - S_star);
      const double pst = (L.p + ml * (S_star - L.vn));
      fn[0] = S_star * (S_L * u_l[0] - L.fn[0]) * inv;
#pragma unroll
      for (int i = 0; i < 3; i++) fn[i + 1] = (S_star * (S_L * u_l[i + 1] - L.fn[i + 1]) + S_L * pst * n[i]) * inv;
      fn[4] = (S_star * (S_L * u_l[4] - L.fn[4]) + S_L * pst * S_star) * inv;
    }
    else if (S_R >= 0)
    {
      const double inv = 1.0 / (S_R - S_star);
      const double pst = (R.p + mr * (S_star - R.vn));
      fn[0] = S_star * (S_R * u_r[0] - R.fn[0]) * inv;
#pragma unroll
      for (int i = 0; i < 3; i++) fn[i + 1] = (S_star * (S_R * u_r[i + 1] - R.fn[i + 1]) + S_R * pst * n[i]) * inv;
      fn[4] = (S_star * (S_R * u_r[4] - R.fn[4]) + S_R * pst * S_star) * inv;
    }
    else
    {
#pragma unroll
      for (int k = 0; k < 5; k++) fn[k] = R.fn[k];
    }
  }
  else if (P.riemann_solve_type == 0) // Rusanov
  {
    const double eig = sqrt(gamma * (L.p + R.p) / (L.rho + R.rho)) + 0.5 * fabs(L.vn + R.vn);
#pragma unroll
    for (int k = 0; k < 5; k++) fn[k] = 0.5 * ((L.fn[k] + R.fn[k]) - eig * (u_r[k] - u_l[k]));
  }
  else // RoeM
  {
    double va[3], dv[3], du[5], bdq[5];
    const double drho = R.rho - L.rho, dp = R.p - L.p, dh = R.h - L.h, dvn = R.vn - L.vn;
    const double sq_rho = sqrt(R.rho * L.ir);
    const double rrho = 1.0 / (1.0 + sq_rho);
    const double ratr = sq_rho * rrho;
    const double ra = sq_rho * L.rho;
    const double ha = L.h * rrho + R.h * ratr;
    double qq = 0., va_n = 0.;
#pragma unroll
    for (int i = 0; i < 3; i++)
    {
      dv[i] = R.v[i] - L.v[i];
      va[i] = L.v[i] * rrho + R.v[i] * ratr;
      qq += va[i] * va[i];
      va_n += n[i] * va[i];
    }
    const double aa = sqrt(gm1 * (ha - 0.5 * qq));
    const double rcp_aa = 1.0 / aa;
    const double abs_ma = fabs(va_n * rcp_aa);
    double b1 = fmax(0.0, fmax(va_n + aa, R.vn + aa));
    double b2 = fmin(0.0, fmin(va_n - aa, L.vn - aa));
    double b1b2 = b1 * b2;
    const double rcp_b1_b2 = 1.0 / (b1 - b2);
    b1 = b1 * rcp_b1_b2;
    b2 = b2 * rcp_b1_b2;
    b1b2 = b1b2 * rcp_b1_b2;
    const double hh = 1.0 - ((L.p < R.p) ? (L.p / R.p) : (R.p / L.p));
    const double ff = ((abs_ma != 0) ? pow(abs_ma, hh) : 1.);
    const double gg = ff / (1.0 + abs_ma);
#pragma unroll
    for (int i = 0; i < 4; i++) du[i] = u_r[i] - u_l[i];
    du[4] = R.rho * R.h - L.rho * L.h;
    bdq[0] = drho - ff * dp * rcp_aa * rcp_aa;
    bdq[4] = bdq[0] * ha + ra * dh;
#pragma unroll
    for (int i = 0; i < 3; i++) bdq[i + 1] = bdq[0] * va[i] + ra * (dv[i] - n[i] * dvn);
#pragma unroll
    for (int i = 0; i < 5; i++) fn[i] = (b1 * L.fn[i] - b2 * R.fn[i]) + b1b2 * (du[i] - gg * bdq[i]);
  }
}

// element id at a position of the launch order (identity when the rank has no partition faces)
__device__ __forceinline__ int elem_id(const fused_args &A, int pos) { return A.elist ? A.elist[pos] : pos; }

#include "hf_fused_kernels.cuh"
#include "hf_fused9.cuh"

// gather partition-face blocks into the send buffer: out[inter][block] = arr[block_of(inter)]
__global__ void k_pack_blocks(const double *__restrict__ arr, const int *__restrict__ blk, double *__restrict__ out, int n_inters, int blk_doubles)
{
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)n_inters * blk_doubles) return;
  int i = (int)(t / blk_doubles), w = (int)(t - (long long)i * blk_doubles);
  out[t] = arr[(size_t)blk[i] * blk_doubles + w];
}
} // namespace

// ---- host side -------------------------------------------------------------------------------------------------------------
struct hf_fused_state
{
  bool available = false;
  std::string why; // reason the fused path is not available
  int order = 0, n_eles = 0, n_mpi = 0;
  int n_bdy = 0; // boundary faces (generation 9 only): virtual neighbour blocks behind the receive blocks of fu
  int n_bele = 0; // elements with a boundary face
  int *belist = nullptr;          // [n_bele] those elements (device order, ascending)
  unsigned char *bskip = nullptr; // [ele] 1 = has a boundary face
  double *fu[2] = {nullptr, nullptr};
  int cur = 0;
  double *fv = nullptr;
  double *em = nullptr;
  int *nbr = nullptr;
  int *finfo = nullptr;
  unsigned long long *bmask = nullptr;
  fused_tables T;
  int *mpi_blk = nullptr; // [n_mpi] own face block of every partition interface
  int *nidx = nullptr;    // [ele][NFP] neighbour value index into fu
  double *nlf = nullptr;  // RoeM: left normal per flux point
  int *elist = nullptr;   // interior elements (ascending), then elements with a partition face (ascending)
  int n_interior = 0;
  bool elist_identity = false;
  double *out_u = nullptr, *out_g = nullptr;
  int E = 2, NT = 128;
  bool os = false; // one-sided LDG kernels (generation 7 / 9) in use
  bool gen9 = false; // generation 9 (k_face9 + k_resid9): face blocks of FB doubles
  double *gn = nullptr;
  unsigned *cls9 = nullptr, *cl9 = nullptr;
  uint2 *tw9 = nullptr;
  int n_classes9 = 0;
  double c5s[2][6] = {{0}};
  int fu_blk = 0;  // doubles per face block of fu
  int fv_blk = 0;  // doubles per face block of fv
  unsigned long long own_xor = 0;
  std::vector<unsigned long long> h_pmask; // own masks of the partition faces
  std::vector<unsigned long long> h_bmask; // host copy of bmask (partition faces may be re-decided when the communicator arrives)
  std::vector<int> h_finfo;
  std::string os_why;
};

void hf_fused_destroy(hf_ctx *c)
{
  delete c->fz;
  c->fz = nullptr;
}

// called from hf_dev_upload_eles while the host metric arrays are at hand
int hf_fused_on_upload(hf_ctx *c, hf_eles_dev &e, const hf_eles_desc *d)
{
  if (e.ele_type != 4 || e.n_dims != 3) return 0;
  const int nu = e.n_upts, nf = e.n_fpts, ne = e.n_eles, nfi = nf / 6;
  e.h_em.assign((size_t)ne * 10, 0.);
  e.h_face_geo.assign((size_t)ne * 24, 0.);
  e.h_own_sign.assign((size_t)ne * nf, 1);
  if (c->prm.riemann_solve_type == 2) e.h_norm_fpts.assign(d->norm_fpts, d->norm_fpts + (size_t)3 * nf * ne);
  double defect = 0.;
  for (int i = 0; i < ne; i++)
  {
    const double *J0 = d->JGinv_upts + (size_t)9 * nu * i;
    double scale = 0.;
    for (int q = 0; q < 9; q++) scale = std::max(scale, fabs(J0[q]));
    for (int q = 0; q < 9; q++) e.h_em[(size_t)i * 10 + q] = J0[q];
    e.h_em[(size_t)i * 10 + 9] = d->detjac_upts[(size_t)nu * i];
    for (int p = 0; p < nu; p++)
    {
      for (int q = 0; q < 9; q++) defect = std::max(defect, fabs(d->JGinv_upts[(size_t)9 * (p + (size_t)nu * i) + q] - J0[q]) / scale);
      defect = std::max(defect, fabs(d->detjac_upts[p + (size_t)nu * i] / d->detjac_upts[(size_t)nu * i] - 1.0));
    }
    for (int p = 0; p < nf; p++)
    {
      for (int q = 0; q < 9; q++) defect = std::max(defect, fabs(d->JGinv_fpts[(size_t)9 * (p + (size_t)nf * i) + q] - J0[q]) / scale);
      defect = std::max(defect, fabs(d->detjac_fpts[p + (size_t)nf * i] / d->detjac_upts[(size_t)nu * i] - 1.0));
      int f = p / nfi, p0 = f * nfi;
      size_t gp = p + (size_t)nf * i, g0 = p0 + (size_t)nf * i, S = (size_t)nf * ne;
      defect = std::max(defect, fabs(d->tdA_fpts[gp] / d->tdA_fpts[g0] - 1.0));
      double nrm[3];
      for (int k = 0; k < 3; k++)
      {
        nrm[k] = d->norm_fpts[gp + k * S];
        defect = std::max(defect, fabs(nrm[k] - d->norm_fpts[g0 + k * S]));
      }
      // the reference's "consistent switch" (src/inters.cpp:566-581, 620-634) on the exact normal of this point
      int s = 1;
      if (nrm[0] < 0.) s = -1;
      else if (nrm[0] == 0.)
      {
        if ((nrm[0] + nrm[1]) < 0.) s = -1;
        else if ((nrm[0] + nrm[1]) == 0)
        {
          if ((nrm[0] + nrm[2]) < 0.) s = -1;
        }
      }
      e.h_own_sign[(size_t)i * nf + p] = (int8_t)s;
      if (p == p0)
      {
        double *g = &e.h_face_geo[(size_t)i * 24 + 4 * f];
        g[0] = d->tdA_fpts[gp];
        for (int k = 0; k < 3; k++) g[1 + k] = nrm[k];
      }
    }
  }
  e.affine_defect = defect;
  e.affine = defect < 1e-10;
  return 0;
}

template <int N>
static bool extract_tables(hf_eles_dev &e, bool visc, fused_tables &T, std::string &why)
{
  constexpr int NU = N * N * N, NFP = 6 * N * N, NN = N * N;
  memset(&T, 0, sizeof(T));
  const std::vector<double> &o0 = e.h_op[0], &o3 = e.h_op[3];
  auto upt = [](int a, int b, int c) { return a + N * b + NN * c; };
  // 1-D tables
  for (int i = 0; i < N; i++)
  {
    T.Lm[i] = o0[(4 * NN + fpt_of_upt<N>(4, i, 0, 0)) + (size_t)NFP * upt(i, 0, 0)];
    T.Lp[i] = o0[(2 * NN + fpt_of_upt<N>(2, i, 0, 0)) + (size_t)NFP * upt(i, 0, 0)];
    for (int j = 0; j < N; j++) T.D[i * N + j] = e.h_op[4][upt(i, 0, 0) + (size_t)NU * upt(j, 0, 0)];
  }
  for (int f = 0; f < 6; f++)
    for (int m = 0; m < N; m++)
    {
      int dir = face_dir(f);
      int a = dir == 0 ? m : 0, b = dir == 1 ? m : 0, c = dir == 2 ? m : 0;
      int fp = f * NN + fpt_of_upt<N>(f, a, b, c);
      T.c3[f * N + m] = o3[upt(a, b, c) + (size_t)NU * fp];
      if (visc) T.c5[f * N + m] = e.h_op[10 + dir][upt(a, b, c) + (size_t)NU * fp];
    }
  // verify that the dense operators are exactly these tensor products (otherwise the fused kernels would not
  // compute what the reference computes)
  for (int p = 0; p < NU; p++)
  {
    int a = p % N, b = (p / N) % N, c = p / NN;
    int idx[3] = {a, b, c};
    for (int f = 0; f < 6; f++)
      for (int j = 0; j < NN; j++)
      {
        int fp = f * NN + j, dir = face_dir(f);
        bool on_line = fpt_of_upt<N>(f, a, b, c) == j;
        double L = face_sgn(f) > 0 ? T.Lp[idx[dir]] : T.Lm[idx[dir]];
        if (o0[fp + (size_t)NFP * p] != (on_line ? L : 0.0)) { why = "opp_0 is not the expected tensor product"; return false; }
        if (o3[p + (size_t)NU * fp] != (on_line ? T.c3[f * N + idx[dir]] : 0.0)) { why = "opp_3 is not the expected tensor product"; return false; }
        for (int d = 0; d < 3; d++)
        {
          double want1 = (on_line && d == dir) ? L * face_sgn(f) : 0.0;
          if (e.h_op[7 + d][fp + (size_t)NFP * p] != want1) { why = "opp_1 is not the expected tensor product"; return false; }
          if (visc)
          {
            double want5 = (on_line && d == dir) ? T.c5[f * N + idx[dir]] : 0.0;
            if (e.h_op[10 + d][p + (size_t)NU * fp] != want5) { why = "opp_5 is not the expected tensor product"; return false; }
          }
        }
      }
    for (int q = 0; q < NU; q++)
    {
      int a2 = q % N, b2 = (q / N) % N, c2 = q / NN;
      double w0 = (b == b2 && c == c2) ? T.D[a * N + a2] : 0.0;
      double w1 = (a == a2 && c == c2) ? T.D[b * N + b2] : 0.0;
      double w2 = (a == a2 && b == b2) ? T.D[c * N + c2] : 0.0;
      if (e.h_op[4][p + (size_t)NU * q] != w0 || e.h_op[5][p + (size_t)NU * q] != w1 || e.h_op[6][p + (size_t)NU * q] != w2)
      { why = "opp_2 is not the expected tensor product"; return false; }
    }
  }
  // neighbour permutations: left uses lut[j], right uses the inverse (reference src/inters.cpp:232-256)
  for (int rot = 0; rot < 4; rot++)
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++)
      {
        int v;
        if (rot == 0) v = (N - 1 - j) + N * i;
        else if (rot == 1) v = NN - (N - 1 - j) - N * i - 1;
        else if (rot == 2) v = N * j + i;
        else v = NN - N * j - i - 1;
        T.perm[rot * 36 + (i * N + j)] = (unsigned char)v;
        T.perm[(rot + 4) * 36 + v] = (unsigned char)(i * N + j);
      }
  for (int f = 0; f < 6; f++)
    for (int j = 0; j < NN; j++)
    {
      int base, stride;
      line_of_fpt<N>(f, j, base, stride);
      T.lbase[f * NN + j] = (unsigned short)base;
    }
  for (int p = 0; p < NU; p++)
  {
    int a = p % N, b = (p / N) % N, c = p / NN;
    for (int f = 0; f < 6; f++) T.uptab[p][f] = (unsigned char)(f * NN + fpt_of_upt<N>(f, a, b, c));
    T.uptab[p][6] = (unsigned char)(a | (b << 4));
    T.uptab[p][7] = (unsigned char)c;
  }
  return true;
}

static inline double nblk_total(int ne, int n_mpi) { return (double)ne * 6 + n_mpi; }

int hf_fused_available(hf_ctx *c) { return c->fz && c->fz->available; }

// line-task table of k_resid9 at P = 4: every half-warp of a line access hits 16 distinct shared-memory banks
// generated by tools/bank_layout.py 300000; word = field | c1 << 3 | c2 << 6
static const unsigned short lut9_p4[3][125] = {
  // direction 0: wavefronts per 8 half-warps (volume, face minus, face plus, flipped minus, flipped plus) = [8, 8, 8, 8, 8]; natural order: [8, 15, 8, 8, 15]
  {8, 80, 33, 73, 193, 201, 225, 195, 203, 227, 267, 283, 291, 204, 212, 220, 96, 128, 224, 25, 65, 81, 2, 18, 66, 98, 19, 131, 139, 259, 275, 260, 64, 72, 136, 9, 97, 273, 67, 75, 91, 99, 147, 155, 163, 219, 92, 132, 144, 256, 288, 17, 265, 281, 289, 10, 26, 138, 258, 266, 282, 290, 83, 140, 160, 257, 210, 218, 35, 12, 20, 28, 36, 68, 76, 148, 156, 164, 284, 292, 88, 34, 130, 146, 154, 162, 194, 202, 226, 3, 4, 100, 196, 228, 268, 276, 0, 16, 24, 32, 152, 192, 264, 272, 280, 1, 89, 209, 217, 274, 11, 27, 200, 208, 216, 129, 137, 145, 153, 161, 74, 82, 90, 211, 84},
  // direction 1: wavefronts per 8 half-warps (volume, face minus, face plus, flipped minus, flipped plus) = [8, 10, 10, 10, 10]; natural order: [16, 8, 15, 15, 8]
  {64, 96, 128, 256, 272, 288, 25, 257, 289, 3, 27, 35, 267, 283, 196, 228, 216, 65, 193, 201, 209, 217, 225, 281, 138, 154, 75, 91, 12, 28, 204, 220, 264, 153, 161, 10, 155, 4, 68, 76, 84, 92, 100, 132, 140, 148, 156, 164, 8, 73, 81, 2, 18, 34, 274, 131, 147, 163, 195, 203, 211, 219, 227, 291, 80, 1, 17, 33, 273, 74, 130, 146, 162, 194, 202, 210, 218, 226, 290, 139, 0, 16, 24, 32, 72, 88, 152, 160, 280, 9, 89, 26, 266, 282, 11, 284, 224, 66, 82, 98, 258, 19, 67, 83, 99, 259, 275, 20, 212, 260, 276, 292, 136, 144, 192, 200, 208, 97, 129, 137, 145, 265, 90, 36, 268},
  // direction 2: wavefronts per 8 half-warps (volume, face minus, face plus, flipped minus, flipped plus) = [8, 11, 8, 8, 11]; natural order: [12, 15, 8, 8, 15]
  {96, 264, 33, 273, 66, 74, 90, 210, 11, 27, 203, 267, 275, 283, 204, 220, 272, 89, 257, 10, 18, 26, 98, 130, 154, 162, 266, 274, 132, 164, 268, 276, 24, 80, 1, 17, 65, 81, 97, 137, 153, 193, 209, 225, 138, 194, 131, 259, 128, 144, 152, 160, 200, 216, 288, 73, 129, 145, 161, 217, 163, 4, 20, 156, 136, 224, 82, 67, 75, 83, 99, 139, 155, 195, 211, 219, 227, 12, 28, 140, 72, 192, 208, 9, 25, 265, 281, 2, 218, 36, 68, 84, 100, 196, 212, 228, 0, 8, 16, 32, 64, 88, 201, 146, 258, 3, 19, 35, 91, 147, 76, 292, 256, 280, 289, 34, 202, 226, 282, 290, 291, 92, 148, 260, 284},
};

// Generation 9: elements with the same owner masks and face info (rotation, side, partition flag) do the same thing at every line end.
// One table of packed line tasks (k_resid9) and one set of pass lists (k_face9) per such class; an element only stores its class
// (a structured block has a handful of classes; the tables stay in L2).
template <int N>
static void build_classes9_n(const std::vector<unsigned long long> &bmask, const std::vector<int> &finfo, unsigned long long own_xor, int ne,
                             std::vector<unsigned> &cls, std::vector<uint2> &tw, std::vector<unsigned> &cl)
{
  constexpr int NN = N * N, NTASK = NF * NN;
  const unsigned long long full = NN == 64 ? ~0ull : ((1ull << NN) - 1ull);
  std::map<std::array<unsigned long long, 12>, int> ids;
  cls.resize(ne);
  for (int i = 0; i < ne; i++)
  {
    std::array<unsigned long long, 12> key;
    unsigned long long own[6];
    int info[6];
    unsigned kinds = 0;
    for (int f = 0; f < 6; f++)
    {
      own[f] = (bmask[(size_t)i * 6 + f] ^ own_xor) & full;
      info[f] = finfo[(size_t)i * 6 + f];
      key[f] = own[f];
      key[6 + f] = (unsigned long long)info[f];
      kinds |= (own[f] == 0ull ? 0u : (own[f] == full ? 1u : 2u)) << (2 * f);
    }
    auto it = ids.find(key);
    int id;
    if (it == ids.end())
    {
      id = (int)ids.size();
      ids.emplace(key, id);
      tw.resize((size_t)(id + 1) * 3 * NTASK);
      cl.resize((size_t)(id + 1) * CL9_WORDS);
      for (int d = 0; d < 3; d++)
      {
        const int fm = d == 0 ? 4 : (d == 1 ? 1 : 0), fp = d == 0 ? 2 : (d == 1 ? 3 : 5);
        for (int t = 0; t < NTASK; t++)
        {
          // natural order: field-major, lines in ascending solution-point order; P = 4 takes the bank-conflict-free table
          const int k = t / NN, l = t % NN;
          const unsigned lut = N == 5 ? lut9_p4[d][t] : (unsigned)(k | ((l % N) << 3) | ((l / N) << 6));
          const task9 w = make_task9<N>(d, lut, own[fm], own[fp], info[fm], info[fp]);
          tw[((size_t)id * 3 + d) * NTASK + t] = make_uint2(w.w0, w.w1);
        }
      }
      make_class_lists9<N>(own, &cl[(size_t)id * CL9_WORDS]);
    }
    else
      id = it->second;
    cls[i] = (unsigned)id | (kinds << 20);
  }
}

static int build_classes9(hf_ctx *c, hf_fused_state *Z, const std::vector<unsigned long long> &bmask, const std::vector<int> &finfo)
{
  const int ne = (int)(bmask.size() / 6);
  std::vector<unsigned> cls, cl;
  std::vector<uint2> tw;
  switch (c->eles[4].order + 1)
  {
  case 2: build_classes9_n<2>(bmask, finfo, Z->own_xor, ne, cls, tw, cl); break;
  case 3: build_classes9_n<3>(bmask, finfo, Z->own_xor, ne, cls, tw, cl); break;
  case 4: build_classes9_n<4>(bmask, finfo, Z->own_xor, ne, cls, tw, cl); break;
  case 5: build_classes9_n<5>(bmask, finfo, Z->own_xor, ne, cls, tw, cl); break;
  case 6: build_classes9_n<6>(bmask, finfo, Z->own_xor, ne, cls, tw, cl); break;
  }
  Z->n_classes9 = (int)(cl.size() / CL9_WORDS);
  if (Z->n_classes9 >= (1 << 20)) { Z->gen9 = false; return 0; } // the class shares a word with the face kinds
  if (hf_alloc_copy(c, &Z->cls9, cls.data(), cls.size())) return 1;
  if (hf_alloc_copy(c, &Z->tw9, tw.data(), tw.size())) return 1;
  if (hf_alloc_copy(c, &Z->cl9, cl.data(), cl.size())) return 1;
  return 0;
}

int hf_fused_prepare(hf_ctx *c)
{
  if (c->fz) return 0;
  hf_fused_state *Z = new hf_fused_state();
  c->fz = Z;
  auto no = [&](const std::string &w) { Z->available = false; Z->why = w; return 0; };
  hf_eles_dev &e = c->eles[4];
  for (int t = 0; t < 4; t++)
    if (c->eles[t].present) return no("fused kernels exist for hexahedra only");
  if (!e.present) return no("no hexahedra");
  if (c->prm.equation != 0 || e.n_fields != NF) return no("fused kernels exist for the Euler / Navier-Stokes equations only");
  // boundary faces: generation 9 only (decided below); its face kernel evaluates the ghost state itself
  hf_bdy_inters_dev &Bd = c->bdys[2];
  if (c->bdys[0].n_inters || c->bdys[1].n_inters) return no("boundary faces of a non-quad type");
  const int n_bdy = Bd.n_inters;
  if (n_bdy)
  {
    if (getenv("HF_FUSED_BDY") && atoi(getenv("HF_FUSED_BDY")) == 0) return no("boundary interfaces present (HF_FUSED_BDY=0: blocked element kernels)");
    if (Bd.wm_upt) return no("boundary interfaces with a wall model run through the blocked element kernels");
    if ((int)Bd.h_ele_l.size() != n_bdy || (int)Bd.h_loc_l.size() != n_bdy || (int)Bd.h_bc_id.size() != n_bdy) return no("boundary interfaces: host lists not kept");
  }
  if (!e.affine) return no("elements are not affine (metric variation inside an element)");
  if (c->prm.over_int) return no("over-integration runs through the staged kernels");
  if (c->prm.LES) return no("LES runs through the staged kernels");
  if (e.order < 1 || e.order > 5) return no("order outside 1..5");
  const bool visc = c->prm.viscous != 0;
  const int N = e.order + 1, NN = N * N, NFP = 6 * NN, ne = e.n_eles;
  fused_tables T;
  bool ok = false;
  switch (N)
  {
  case 2: ok = extract_tables<2>(e, visc, T, Z->why); break;
  case 3: ok = extract_tables<3>(e, visc, T, Z->why); break;
  case 4: ok = extract_tables<4>(e, visc, T, Z->why); break;
  case 5: ok = extract_tables<5>(e, visc, T, Z->why); break;
  case 6: ok = extract_tables<6>(e, visc, T, Z->why); break;
  }
  if (!ok) return no(Z->why);
  // connectivity per (ele, face)
  std::vector<int> nbr((size_t)ne * 6, -1);
  std::vector<int> finfo((size_t)ne * 6, 0);
  std::vector<unsigned long long> bmask((size_t)ne * 6, 0ull);
  std::vector<double> em((size_t)ne * EM, 0.);
  for (int i = 0; i < ne; i++)
  {
    for (int q = 0; q < 9; q++) em[(size_t)i * EM + q] = e.h_em[(size_t)i * 10 + q];
    em[(size_t)i * EM + 9] = 1.0 / e.h_em[(size_t)i * 10 + 9]; // the kernels only ever divide by detjac
    for (int f = 0; f < 6; f++)
      for (int q = 0; q < 4; q++) em[(size_t)i * EM + 10 + 4 * f + q] = e.h_face_geo[(size_t)i * 24 + 4 * f + q];
  }
  hf_int_inters_dev &I = c->ints[2];
  for (int i = 0; i < I.n_inters; i++)
  {
    int el = I.h_ele_l[i], fl = I.h_loc_l[i], er = I.h_ele_r[i], fr = I.h_loc_r[i], rot = I.h_rot[i];
    nbr[(size_t)el * 6 + fl] = er * 6 + fr;
    nbr[(size_t)er * 6 + fr] = el * 6 + fl;
    finfo[(size_t)el * 6 + fl] = rot;
    finfo[(size_t)er * 6 + fr] = rot + 4;
    // the right element uses the left element's normal; its own tdA stays
    for (int q = 1; q < 4; q++) em[(size_t)er * EM + 10 + 4 * fr + q] = e.h_face_geo[(size_t)el * 24 + 4 * fl + q];
    for (int j = 0; j < NN; j++)
    {
      // bit = (ldg_beta switched to -beta at this flux point) XOR (this element is the right side): the weight of the
      // element's OWN value / flux in the LDG common solution / flux is 0.5 + beta when the bit is clear, 0.5 - beta when set
      signed char s = e.h_own_sign[(size_t)el * NFP + fl * NN + j];
      if (s < 0) bmask[(size_t)el * 6 + fl] |= 1ull << j;
      else bmask[(size_t)er * 6 + fr] |= 1ull << T.perm[rot * 36 + j];
    }
  }
  hf_mpi_inters_dev &M = c->mpis[2];
  if (c->mpis[0].n_inters || c->mpis[1].n_inters) return no("partition faces of a non-quad type");
  Z->n_mpi = M.n_inters;
  std::vector<int> mpi_blk(std::max(M.n_inters, 1), 0);
  for (int i = 0; i < M.n_inters; i++)
  {
    int el = M.h_ele_l[i], fl = M.h_loc_l[i];
    nbr[(size_t)el * 6 + fl] = ne * 6 + i; // receive block behind the last element
    finfo[(size_t)el * 6 + fl] = M.h_rot[i] + 8;
    mpi_blk[i] = el * 6 + fl;
    for (int j = 0; j < NN; j++)
      if (e.h_own_sign[(size_t)el * NFP + fl * NN + j] < 0) bmask[(size_t)el * 6 + fl] |= 1ull << j;
  }
  // boundary faces: every flux point owned (whatever the sign of beta), no rotation, the "neighbour" is the virtual block that k_face9 fills
  // with the boundary's common solution
  const unsigned long long full_mask = NN == 64 ? ~0ull : ((1ull << NN) - 1ull);
  const unsigned long long own_xor0 = c->prm.ldg_beta > 0. ? full_mask : 0ull;
  for (int i = 0; i < n_bdy; i++)
  {
    const int el = Bd.h_ele_l[i], fl = Bd.h_loc_l[i];
    if (Bd.h_ele_type_l[i] != 4) return no("boundary interface of a non-hexahedron");
    nbr[(size_t)el * 6 + fl] = ne * 6 + M.n_inters + i;
    finfo[(size_t)el * 6 + fl] = 16 | (Bd.h_bc_id[i] << 8);
    bmask[(size_t)el * 6 + fl] = full_mask ^ own_xor0;
  }
  Z->n_bdy = n_bdy;
  std::vector<unsigned char> bskip(ne, 0);
  std::vector<int> belist;
  for (int i = 0; i < n_bdy; i++) bskip[Bd.h_ele_l[i]] = 1;
  for (int i = 0; i < ne; i++) if (bskip[i]) belist.push_back(i);
  Z->n_bele = (int)belist.size();
  for (size_t q = 0; q < nbr.size(); q++)
    if (nbr[q] < 0) return no("an element face has no neighbour");
  // One-sided LDG (generation 7): |beta| = 0.5 makes the LDG weights exactly 1 and 0, so every flux-point pair has one
  // owner (own weight 0.5 + beta when the bmask bit is clear, 0.5 - beta when set).  On several ranks the two sides of a
  // partition face must agree on it: checked when the communicator arrives (hf_fused_after_nccl).
  Z->os = visc && fabs(c->prm.ldg_beta) == 0.5 && !getenv("HF_FUSED_GEN6");
  Z->own_xor = c->prm.ldg_beta > 0. ? (NN == 64 ? ~0ull : ((1ull << NN) - 1ull)) : 0ull;
  // generation 9 (hf_fused9.cuh): the default at P = 2, 3, 4 (measured faster than generation 7 there; HF_FUSED_GEN9=1 turns it on at any
  // order, HF_FUSED_GEN7=1 keeps generation 7).  Its tangential pass takes the correction-function derivative per side of the line, not per
  // face: the three directions of a hexahedron share one 1-D function, verified here on the reference's numbers.
  {
    bool side_uniform = visc;
    for (int m = 0; m < N && visc; m++)
    {
      Z->c5s[0][m] = T.c5[4 * N + m]; // minus faces: 4 (x), 1 (y), 0 (z)
      Z->c5s[1][m] = T.c5[2 * N + m]; // plus faces: 2 (x), 3 (y), 5 (z)
      if (T.c5[1 * N + m] != Z->c5s[0][m] || T.c5[0 * N + m] != Z->c5s[0][m] || T.c5[3 * N + m] != Z->c5s[1][m] || T.c5[5 * N + m] != Z->c5s[1][m])
        side_uniform = false;
    }
    const bool want9 = getenv("HF_FUSED_GEN9") ? atoi(getenv("HF_FUSED_GEN9")) != 0 : (N == 5 || N == 4 || N == 3 || (n_bdy > 0 && N <= 5)); // with boundary faces: every order generation 9 is tested at (P = 1..4)
    Z->gen9 = Z->os && side_uniform && want9 && !getenv("HF_FUSED_GEN7");
  }
  if (n_bdy && !Z->gen9)
    return no("boundary interfaces present (the fused kernels take them in generation 9 only: viscous, |ldg_beta| = 0.5; otherwise the blocked element kernels)");
  const int FB = (NF * NN + 1) & ~1; // generation 9: face blocks padded to an even count (16-byte aligned for the bulk copies)
  Z->fu_blk = Z->gen9 ? FB : NF * NN;
  Z->fv_blk = Z->gen9 ? FB : (Z->os ? NF * NN : 4 * NN);
  Z->h_bmask = bmask;
  Z->h_finfo = finfo;
  Z->h_pmask.resize(M.n_inters);
  for (int i = 0; i < M.n_inters; i++) Z->h_pmask[i] = bmask[(size_t)M.h_ele_l[i] * 6 + M.h_loc_l[i]] ^ Z->own_xor;
  // per own flux point: where the neighbour's value of field 0 sits in fu (block * NF*NN + permuted flux point)
  std::vector<int> nidx((size_t)ne * NFP);
  if (((double)nblk_total(ne, M.n_inters) + n_bdy) * (NF * NN + 1) > 2.0e9) return no("face arrays exceed int32 indexing");
  for (int i = 0; i < ne; i++)
    for (int f = 0; f < 6; f++)
      for (int j = 0; j < NN; j++)
        nidx[(size_t)i * NFP + f * NN + j] = nbr[(size_t)i * 6 + f] * (NF * NN) + T.perm[(finfo[(size_t)i * 6 + f] & 7) * 36 + j];
  // launch order: elements without a partition face first, so their work overlaps the halo exchange
  std::vector<char> is_halo(ne, 0);
  for (int i = 0; i < M.n_inters; i++) is_halo[M.h_ele_l[i]] = 1;
  std::vector<int> elist;
  elist.reserve(ne);
  for (int i = 0; i < ne; i++) if (!is_halo[i]) elist.push_back(i);
  Z->n_interior = (int)elist.size();
  for (int i = 0; i < ne; i++) if (is_halo[i]) elist.push_back(i);
  Z->elist_identity = true; // true when the device element order already has the interior elements first (hf_dev_set_element_order)
  for (int i = 0; i < ne; i++) if (elist[i] != i) Z->elist_identity = false;
  Z->order = e.order;
  Z->n_eles = ne;
  const size_t nblk = (size_t)ne * 6 + M.n_inters;
  if (hf_alloc_zero(c, &Z->fu[0], (nblk + n_bdy) * FB)) return 1;
  if (hf_alloc_zero(c, &Z->fu[1], (nblk + n_bdy) * FB)) return 1;
  if (visc && hf_alloc_zero(c, &Z->fv, nblk * FB)) return 1;
  if (Z->gen9)
  {
    if (hf_alloc_zero(c, &Z->gn, (size_t)ne * 6 * FB)) return 1;
    if (build_classes9(c, Z, bmask, finfo)) return 1;
  }
  if (n_bdy)
  {
    if (hf_alloc_copy(c, &Z->belist, belist.data(), belist.size())) return 1;
    if (hf_alloc_copy(c, &Z->bskip, bskip.data(), bskip.size())) return 1;
  }
  if (hf_alloc_copy(c, &Z->em, em.data(), em.size())) return 1;
  if (hf_alloc_copy(c, &Z->nbr, nbr.data(), nbr.size())) return 1;
  if (hf_alloc_copy(c, &Z->finfo, finfo.data(), finfo.size())) return 1;
  if (hf_alloc_copy(c, &Z->bmask, bmask.data(), bmask.size())) return 1;
  Z->T = T;
  if (hf_alloc_copy(c, &Z->mpi_blk, mpi_blk.data(), mpi_blk.size())) return 1;
  if (hf_alloc_copy(c, &Z->elist, elist.data(), elist.size())) return 1;
  if (hf_alloc_copy(c, &Z->nidx, nidx.data(), nidx.size())) return 1;
  if (c->prm.riemann_solve_type == 2)
  {
    // RoeM's f = |Ma_n|^h with its `Ma_n != 0 ? pow : 1` switch (reference src/inters.cpp:400-404) is discontinuous at Ma_n = 0, and the
    // reference's normals differ inside a face in the last bit: of the flux points of one face some carry an exact 0 in a component,
    // others 1e-16 (measured on the reference's dumps: a quarter of the faces).  With ONE normal per face (em) the normal Mach number
    // of a flow along the face is exactly 0 at other points than in the reference, f jumps by 40 h there, and the result leaves the
    // reference by 1e-7 .. 1e-6 (tools/roem_probe.py).  So this solver -- and only it -- gets the reference's normal of every
    // flux point: the left side's, i.e. the own one on left elements and partition faces, the neighbour's at the facing point on
    // right elements.
    if (e.h_norm_fpts.size() != (size_t)3 * NFP * ne) return no("RoeM: the flux-point normals were not kept at upload");
    std::vector<double> nlf((size_t)ne * NFP * 3);
    const size_t S = (size_t)NFP * ne;
    for (int i = 0; i < ne; i++)
      for (int f = 0; f < 6; f++)
        for (int j = 0; j < NN; j++)
        {
          const int info = finfo[(size_t)i * 6 + f];
          size_t src = (size_t)(f * NN + j) + (size_t)NFP * i;
          if ((info & 4) && !(info & 8))
          {
            const int blk = nbr[(size_t)i * 6 + f];
            const int jn = nidx[(size_t)i * NFP + f * NN + j] - blk * (NF * NN);
            src = (size_t)((blk % 6) * NN + jn) + (size_t)NFP * (blk / 6);
          }
          for (int k = 0; k < 3; k++) nlf[((size_t)(i * 6 + f) * NN + j) * 3 + k] = e.h_norm_fpts[src + k * S];
        }
    if (hf_alloc_copy(c, &Z->nlf, nlf.data(), nlf.size())) return 1;
  }
  if (M.n_inters + n_bdy)
  {
    // k_resid9 publishes every face whose neighbour block lies behind the last element into the send buffer: the boundary faces' blocks
    // ride along behind the partition faces' (never sent)
    if (hf_alloc_zero(c, &Z->out_u, (size_t)(M.n_inters + n_bdy) * FB)) return 1;
    if (visc && M.n_inters && hf_alloc_zero(c, &Z->out_g, (size_t)M.n_inters * FB)) return 1;
  }
  // host-side extracts are no longer needed
  std::vector<double>().swap(e.h_em);
  std::vector<double>().swap(e.h_face_geo);
  std::vector<int8_t>().swap(e.h_own_sign);
  Z->available = true;
  return 0;
}

namespace
{
template <int N, int E, int NT, int MINB>
int launch_all(hf_ctx *c, hf_fused_state *Z, fused_args &A, int what, int lo, int hi)
{
  if (hi <= lo) return 0;
  A.lo = lo;
  A.hi = hi;
  // what: 0 face values, 1 gradient kernel, 2 residual kernel; 3 / 4 = the one-sided (generation 7) pair
  const size_t smem_g = sizeof(smem6<N, E>), smem_r = smem_g;
  const int grid = (hi - lo + E - 1) / E;
  static bool attr_done = false;
  if (!attr_done)
  {
    HF_CUDA(cudaFuncSetAttribute(k_face_values6<N, E, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    HF_CUDA(cudaFuncSetAttribute(k_grad6<N, E, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_g));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_r));
    // ask for the largest shared-memory carve-out so that MINB blocks fit on an SM
    HF_CUDA(cudaFuncSetAttribute(k_grad6<N, E, NT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid6<N, E, NT, MINB, false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_grad7<N, E, NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(smem7<N, E>)));
    HF_CUDA(cudaFuncSetAttribute(k_resid7<N, E, NT, MINB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(smem7<N, E>)));
    HF_CUDA(cudaFuncSetAttribute(k_resid7<N, E, NT, MINB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(smem7<N, E>)));
    HF_CUDA(cudaFuncSetAttribute(k_grad7<N, E, NT, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid7<N, E, NT, MINB, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid7<N, E, NT, MINB, true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done = true;
  }
  if (what == 0)
    k_face_values6<N, E, NT><<<grid, NT, smem_g, c->stream>>>(A);
  else if (what == 1)
    k_grad6<N, E, NT, MINB><<<grid, NT, smem_g, c->stream>>>(A);
  else if (what == 3)
    k_grad7<N, E, NT, MINB><<<grid, NT, sizeof(smem7<N, E>), c->stream>>>(A);
  else if (what == 4)
  {
    hf_ktimer_begin(c);
    if (A.grad_out) k_resid7<N, E, NT, MINB, true><<<grid, NT, sizeof(smem7<N, E>), c->stream>>>(A);
    else k_resid7<N, E, NT, MINB, false><<<grid, NT, sizeof(smem7<N, E>), c->stream>>>(A);
    hf_ktimer_end(c);
  }
  else
  {
    hf_ktimer_begin(c);
    if (A.viscous && A.grad_out) k_resid6<N, E, NT, MINB, true, true><<<grid, NT, smem_r, c->stream>>>(A);
    else if (A.viscous) k_resid6<N, E, NT, MINB, true, false><<<grid, NT, smem_r, c->stream>>>(A);
    else k_resid6<N, E, NT, MINB, false, false><<<grid, NT, smem_r, c->stream>>>(A);
    hf_ktimer_end(c);
  }
  c->launches++;
  cudaError_t err = cudaGetLastError();
  static const bool debug_sync = getenv("HF_DEBUG_SYNC") != nullptr; // debugging aid: attribute an asynchronous fault to its kernel
  if (err == cudaSuccess && debug_sync) err = cudaStreamSynchronize(c->stream);
  if (err != cudaSuccess) { hf_set_error(std::string("fused kernel launch (kernel ") + std::to_string(what) + "): " + cudaGetErrorString(err)); return 1; }
  return 0;
}

// generation 9: what 5 = k_face9, 6 = k_resid9, 7 = face data of the current solution (k_resid9 in mode 2); one element per CTA
template <int N, int NT_R, int MINB_R, int NT_F, int MINB_F>
int launch9(hf_ctx *c, fused_args &A, int what, int lo, int hi)
{
  if (hi <= lo) return 0;
  A.lo = lo;
  A.hi = hi;
  const int grid = hi - lo;
  const int smem_r = (int)sizeof(smem9r<N>), smem_f = (int)sizeof(smem9f<N>);
  static bool attr_done = false;
  if (!attr_done)
  {
    HF_CUDA(cudaFuncSetAttribute(k_resid9<N, NT_R, MINB_R, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    HF_CUDA(cudaFuncSetAttribute(k_resid9<N, NT_R, MINB_R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    HF_CUDA(cudaFuncSetAttribute(k_resid9<N, NT_R, MINB_R, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_r));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_f));
    HF_CUDA(cudaFuncSetAttribute(k_resid9<N, NT_R, MINB_R, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_resid9<N, NT_R, MINB_R, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 0>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, false, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    HF_CUDA(cudaFuncSetAttribute(k_face9<N, NT_F, MINB_F, true, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_done = true;
  }
  if (what == 5)
  {
    if (A.bdy_mode == 1) // the elements with a boundary face
    {
      if (A.nlf) k_face9<N, NT_F, MINB_F, true, 1><<<grid, NT_F, smem_f, c->stream>>>(A);
      else k_face9<N, NT_F, MINB_F, false, 1><<<grid, NT_F, smem_f, c->stream>>>(A);
    }
    else if (A.bdy_mode == 2) // all others of a mesh with boundary faces
    {
      if (A.nlf) k_face9<N, NT_F, MINB_F, true, 2><<<grid, NT_F, smem_f, c->stream>>>(A);
      else k_face9<N, NT_F, MINB_F, false, 2><<<grid, NT_F, smem_f, c->stream>>>(A);
    }
    else if (A.nlf) k_face9<N, NT_F, MINB_F, true, 0><<<grid, NT_F, smem_f, c->stream>>>(A);
    else k_face9<N, NT_F, MINB_F, false, 0><<<grid, NT_F, smem_f, c->stream>>>(A);
  }
  else if (what == 7)
    k_resid9<N, NT_R, MINB_R, 2><<<grid, NT_R, smem_r, c->stream>>>(A);
  else
  {
    hf_ktimer_begin(c);
    if (A.grad_out) k_resid9<N, NT_R, MINB_R, 1><<<grid, NT_R, smem_r, c->stream>>>(A);
    else k_resid9<N, NT_R, MINB_R, 0><<<grid, NT_R, smem_r, c->stream>>>(A);
    hf_ktimer_end(c);
  }
  c->launches++;
  cudaError_t err = cudaGetLastError();
  static const bool debug_sync = getenv("HF_DEBUG_SYNC") != nullptr;
  if (err == cudaSuccess && debug_sync) err = cudaStreamSynchronize(c->stream);
  if (err != cudaSuccess) { hf_set_error(std::string("fused kernel launch (generation 9, kernel ") + std::to_string(what) + "): " + cudaGetErrorString(err)); return 1; }
  return 0;
}

int launch(hf_ctx *c, hf_fused_state *Z, fused_args &A, int what, int lo, int hi)
{
  if (what >= 5)
  {
    switch (Z->order)
    {
    case 1: return launch9<2, 32, 8, 32, 8>(c, A, what, lo, hi);
    case 2:
    {
      // measured on B200 (48^3, periodic; GDOF-stage/s): eight CTAs per SM 13.6, twelve (80 registers) 16.5 -- generation 7: 15.3
      static const int cfg9 = getenv("HF_FUSED_CFG9") ? atoi(getenv("HF_FUSED_CFG9")) : 0; // measurement aid
      if (cfg9 == 1) return launch9<3, 64, 8, 64, 8>(c, A, what, lo, hi);
      if (cfg9 == 2) return launch9<3, 64, 16, 64, 16>(c, A, what, lo, hi);
      return launch9<3, 64, 12, 64, 12>(c, A, what, lo, hi);
    }
    case 3:
    {
      // six CTAs per SM 22.5, eight (80 registers) 26.0 -- generation 7: 23.6
      static const int cfg9 = getenv("HF_FUSED_CFG9") ? atoi(getenv("HF_FUSED_CFG9")) : 0;
      if (cfg9 == 1) return launch9<4, 96, 6, 96, 6>(c, A, what, lo, hi);
      if (cfg9 == 2) return launch9<4, 96, 10, 96, 10>(c, A, what, lo, hi);
      return launch9<4, 96, 8, 96, 8>(c, A, what, lo, hi);
    }
    case 4:
    {
      // measurement aid: HF_FUSED_CFG9 bit 0 = five residual CTAs per SM (96 registers), bit 1 = 96-thread face kernel
      static const int cfg9 = getenv("HF_FUSED_CFG9") ? atoi(getenv("HF_FUSED_CFG9")) : 0;
      if (cfg9 == 1) return launch9<5, 128, 5, 128, 6>(c, A, what, lo, hi);
      if (cfg9 == 2) return launch9<5, 128, 6, 96, 6>(c, A, what, lo, hi);
      if (cfg9 == 3) return launch9<5, 128, 5, 96, 6>(c, A, what, lo, hi);
      if (cfg9 == 4) return launch9<5, 128, 6, 128, 5>(c, A, what, lo, hi); // face kernel at 96 registers, five CTAs per SM
      return launch9<5, 128, 6, 128, 6>(c, A, what, lo, hi);
    }
    case 5: return launch9<6, 192, 2, 192, 2>(c, A, what, lo, hi);
    }
    hf_set_error("fused path: unsupported order");
    return 1;
  }
  switch (Z->order)
  {
  case 1: return launch_all<2, 8, 128, 3>(c, Z, A, what, lo, hi);
  case 2: return launch_all<3, 4, 128, 3>(c, Z, A, what, lo, hi);
  case 3: return launch_all<4, 2, 128, 3>(c, Z, A, what, lo, hi);
  case 4:
  {
    // measurement aid: CTA shape per kernel family (HF_FUSED_CFG for all, HF_FUSED_CFG_G / _R for the gradient / residual kernel)
    static const int cfg_all = getenv("HF_FUSED_CFG") ? atoi(getenv("HF_FUSED_CFG")) : 0;
    static const int cfg_g = getenv("HF_FUSED_CFG_G") ? atoi(getenv("HF_FUSED_CFG_G")) : cfg_all;
    static const int cfg_r = getenv("HF_FUSED_CFG_R") ? atoi(getenv("HF_FUSED_CFG_R")) : cfg_all;
    int cfg = (what == 1 || what == 3) ? cfg_g : ((what == 2 || what == 4) ? cfg_r : cfg_all);
    // measured on B200 (64^3): both generation-7 kernels are fastest with one element per CTA and six CTAs per SM (80
    // registers; k_grad7 is within 2 % of that for every shape tried, profiles/ncu_r01_summary.md)
    if (what == 4 && !getenv("HF_FUSED_CFG_R") && !getenv("HF_FUSED_CFG")) cfg = 3;
    if (what == 3 && !getenv("HF_FUSED_CFG_G") && !getenv("HF_FUSED_CFG")) cfg = 3;
    if (cfg == 1) return launch_all<5, 1, 125, 4>(c, Z, A, what, lo, hi);
    if (cfg == 2) return launch_all<5, 1, 125, 5>(c, Z, A, what, lo, hi);
    if (cfg == 3) return launch_all<5, 1, 125, 6>(c, Z, A, what, lo, hi);
    // shapes measured and dropped (profiles/ncu_r01_summary.md): <5,1,64,8>, <5,2,250,1>, <5,1,96,5>, <5,2,160,2>, <5,1,96,6>, <5,2,256,2>, <5,3,192,2>
    return launch_all<5, 2, 125, 3>(c, Z, A, what, lo, hi);
  }
  case 5: return launch_all<6, 1, 128, 3>(c, Z, A, what, lo, hi);
  }
  hf_set_error("fused path: unsupported order");
  return 1;
}

void base_args(hf_ctx *c, hf_fused_state *Z, fused_args &A)
{
  hf_eles_dev &e = c->eles[4];
  memset(&A, 0, sizeof(A));
  A.n_eles = e.n_eles;
  static const bool force_elist = getenv("HF_FORCE_ELIST") != nullptr; // measurement aid: the indirect element order on one GPU
  A.elist = ((Z->n_mpi && !Z->elist_identity) || force_elist) ? Z->elist : nullptr;
  A.u0 = e.disu_upts[0];
  A.u0_out = e.disu_upts[0];
  A.u1 = e.disu_upts[1];
  A.div = e.div_tconf_upts;
  A.fu_cur = Z->fu[Z->cur];
  A.fu_next = Z->fu[Z->cur ^ 1];
  A.fv = Z->fv;
  A.em = Z->em;
  A.nbr = Z->nbr;
  A.finfo = Z->finfo;
  A.bmask = Z->bmask;
  A.nidx = Z->nidx;
  A.nlf = Z->nlf;
  A.dt_local = (c->prm.dt_type == 2) ? e.dt_local : nullptr;
  A.nan_flag = c->d_nan;
  for (int i = 0; i < 36; i++) { A.tD[i] = Z->T.D[i]; A.tc3[i] = Z->T.c3[i]; A.tc5[i] = Z->T.c5[i]; }
  for (int i = 0; i < 6; i++) { A.tL[0][i] = Z->T.Lm[i]; A.tL[1][i] = Z->T.Lp[i]; }
  {
    const int N = Z->order + 1;
    for (int s2 = 0; s2 < 2; s2++)
    {
      for (int j = 0; j < N; j++)
      {
        double a = 0.;
        for (int i = 0; i < N; i++) a += A.tL[s2][i] * A.tD[i * N + j];
        A.tLD[s2][j] = a;
      }
      for (int side = 0; side < 2; side++)
      {
        double a = 0.;
        for (int i = 0; i < N; i++) a += A.tL[s2][i] * Z->c5s[side][i];
        A.lc5s[side][s2] = a;
      }
    }
    for (int side = 0; side < 2; side++)
      for (int i = 0; i < N; i++) A.c5s[side][i] = Z->c5s[side][i];
  }
  A.gn = Z->gn;
  A.send_u = Z->out_u;
  A.send_g = Z->out_g;
  A.cls9 = Z->cls9;
  A.tw9 = Z->tw9;
  A.cl9 = Z->cl9;
  A.P = c->phys;
  A.viscous = c->prm.viscous;
  static const int pf = getenv("HF_FUSED_PF") ? atoi(getenv("HF_FUSED_PF")) : 0; // measured: no effect on B200 (the other resident CTAs already cover the staging latency)
  A.pf_dist = pf;
  A.own_xor = Z->own_xor;
  A.bct = Z->n_bdy ? c->bc_table : nullptr;
  A.R_ref = c->prm.R_ref;
  A.bskip = Z->bskip;
  A.bdy_mode = Z->n_bdy ? 2 : 0;
}

// exchange the partition-face blocks of arr (blk_doubles each): pack -> ncclSend/Recv into the tail of arr, on the
// communication stream; the compute stream carries on and calls exchange_wait before it touches the received blocks
int exchange_post(hf_ctx *c, hf_fused_state *Z, double *arr, double *out, int blk_doubles, int which = -1)
{
  if (Z->n_mpi == 0) return 0;
  hf_mpi_inters_dev &M = c->mpis[2];
  if (!Z->gen9) // generation 9 publishes partition faces straight into the send buffer
  {
    long long n = (long long)Z->n_mpi * blk_doubles;
    k_pack_blocks<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(arr, Z->mpi_blk, out, Z->n_mpi, blk_doubles);
    c->launches++;
  }
  return hf_halo_post(c, M, out, arr + (size_t)Z->n_eles * 6 * blk_doubles, (size_t)blk_doubles, which);
}
int exchange_wait(hf_ctx *c) { return hf_halo_wait(c); }
} // namespace

int hf_fused_extrapolate(hf_ctx *c)
{
  hf_fused_state *Z = c->fz;
  if (!Z || !Z->available) HF_FAIL("fused path not available");
  HF_CUDA(cudaSetDevice(c->device));
  fused_args A;
  base_args(c, Z, A);
  A.fu_next = Z->fu[Z->cur]; // fill the current buffer
  if (exchange_wait(c)) return 1; // an exchange into this buffer may still be in flight
  if (launch(c, Z, A, Z->gen9 ? 7 : 0, 0, Z->n_eles)) return 1;
  if (exchange_post(c, Z, Z->fu[Z->cur], Z->out_u, Z->fu_blk, 0)) return 1;
  c->ufpts_valid = true;
  return 0;
}

int hf_fused_stage(hf_ctx *c, int rk_stage, double time, int keep_residual, int do_update)
{
  (void)time;
  hf_fused_state *Z = c->fz;
  if (!Z || !Z->available) HF_FAIL("fused path not available");
  HF_CUDA(cudaSetDevice(c->device));
  if (!c->ufpts_valid && hf_fused_extrapolate(c)) return 1;
  fused_args A;
  base_args(c, Z, A);
  A.keep_residual = keep_residual;
  A.do_update = do_update;
  if (keep_residual && c->want_gradient && c->prm.viscous)
  {
    // integral diagnostics read grad_disu_upts as the last residual evaluation left it (reference src/eles.cpp:5515-5525)
    hf_eles_dev &eh = c->eles[4];
    if (!eh.grad_disu_upts && hf_alloc_zero(c, &eh.grad_disu_upts, (size_t)eh.n_upts * eh.n_eles * NF * ND)) return 1;
    A.grad_out = eh.grad_disu_upts;
  }
  const hf_params &p = c->prm;
  rk_args &R = A.rk;
  R.dt = p.dt; R.fac = 1.0; R.mode = 0; R.copy_u1 = 0; R.c1 = R.c2 = 0.;
  int stage = rk_stage & 0xff;
  if (p.adv_type == 1)
  {
    R.copy_u1 = stage == 0;
    if (stage < 3) { R.mode = 0; R.fac = 3.0; }
    else { R.mode = 1; R.fac = 4.0; R.c1 = 3.0 / 4.0; R.c2 = 1.0 / 4.0; }
  }
  else if (p.adv_type == 2)
  {
    R.copy_u1 = stage == 0;
    if (stage < 2 || stage == 3) { R.mode = 0; R.fac = 2.0; }
    else { R.mode = 1; R.fac = 6.0; R.c1 = 1.0 / 3.0; R.c2 = 2.0 / 3.0; }
  }
  else if (p.adv_type == 3 || p.adv_type == 4)
  {
    if (stage >= HF_MAX_RK) HF_FAIL("RK stage out of range");
    R.mode = 2; R.c1 = p.RK_a[stage]; R.c2 = p.RK_b[stage];
  }
  else if (p.adv_type != 0)
    HF_FAIL("ERROR: Time integration type not recognised ... ");
  R.dt_fac = R.dt / R.fac;
  // Overlap: the exchange of the face values (posted at the end of the previous stage) runs while the elements
  // without a partition face are processed; the partition-adjacent elements follow once it has arrived.  The same
  // for the viscous normal fluxes between k_grad and k_resid (reference windows: src/solver.cpp:68-73/131-140 and
  // :148-155/199-201).
  const int ni = Z->n_interior, n = Z->n_eles;
  static const bool no_overlap = getenv("HF_NO_OVERLAP") != nullptr; // measurement aid: serialise exchange and compute
  if (no_overlap && exchange_wait(c)) return 1;
  const int kg = Z->gen9 ? 5 : (Z->os ? 3 : 1), kr = Z->gen9 ? 6 : (Z->os ? 4 : 2);
  hf_tl_mark(c, 0, false);
  // Generation 9 with enough interior work: ONE launch per kernel.  The partition-adjacent elements sit at the end of the launch order and
  // wait, inside the kernel, for the completion counter of the exchange they read from (bumped on the communication stream); by the time
  // the grid reaches them the exchange posted a kernel earlier has long landed.  Saves the drain / ramp of two extra launches per stage.
  // Small interiors keep the two-range launches: a grid that starts with waiting CTAs could keep the NCCL kernel off the SMs.
  // meshes with boundary faces: the few elements that have one run the face kernel's ghost-state variant in a launch of their own (behind the
  // main launch, which skips them); partition-adjacent ones among them check the exchange counter like everybody else
  auto face_boundary_elements = [&]() -> int {
    if (!Z->gen9 || Z->n_bele == 0) return 0;
    fused_args B = A;
    B.elist = Z->belist;
    B.bdy_mode = 1;
    if (B.wait_flag) B.wait_from = 0;
    return launch(c, Z, B, 5, 0, Z->n_bele);
  };
  static const bool no_single = getenv("HF_SPLIT_LAUNCH") != nullptr;
  const bool single = Z->gen9 && Z->n_mpi > 0 && !no_overlap && !no_single && ni >= 4096 && c->d_xflag != nullptr;
  if (single)
  {
    A.wait_flag = c->d_xflag;
    A.wait_from = ni;
    A.wait_value = c->x_posted[0];
    if (launch(c, Z, A, kg, 0, n)) return 1;
    if (face_boundary_elements()) return 1;
    hf_tl_mark(c, 3, false);
    c->tl_xmark = 7;
    if (exchange_post(c, Z, Z->fv, Z->out_g, Z->fv_blk, 1)) return 1;
    A.wait_flag = c->d_xflag + 1;
    A.wait_value = c->x_posted[1];
    if (launch(c, Z, A, kr, 0, n)) return 1;
    hf_tl_mark(c, 6, false);
    c->tl_xmark = 9;
  }
  else
  {
  if (p.viscous)
  {
    if (launch(c, Z, A, kg, 0, ni)) return 1;
    hf_tl_mark(c, 1, false);
    if (exchange_wait(c)) return 1;
    hf_tl_mark(c, 2, false);
    if (launch(c, Z, A, kg, ni, n)) return 1;
    if (face_boundary_elements()) return 1;
    hf_tl_mark(c, 3, false);
    c->tl_xmark = 7;
    if (exchange_post(c, Z, Z->fv, Z->out_g, Z->fv_blk, 1)) return 1;
    if (no_overlap && exchange_wait(c)) return 1;
  }
  if (launch(c, Z, A, kr, 0, ni)) return 1;
  hf_tl_mark(c, 4, false);
  if (exchange_wait(c)) return 1;
  hf_tl_mark(c, 5, false);
  if (launch(c, Z, A, kr, ni, n)) return 1;
  hf_tl_mark(c, 6, false);
  c->tl_xmark = 9;
  }
  if (do_update)
  {
    Z->cur ^= 1;
    if (exchange_post(c, Z, Z->fu[Z->cur], Z->out_u, Z->fu_blk, 0)) return 1; // waited for by the next stage
    c->ufpts_valid = true;
  }
  c->tl_stage++;
  return 0;
}

// Called when the communicator of a multi-rank run arrives.  On a partition face each rank would apply the reference's LDG
// sign switch to its OWN normal (src/mpi_inters.cpp:400-483 with src/inters.cpp:566-581); on exact geometry the two normals
// are opposite and the choices complementary, but rounding-level normal components (1e-16 on the 2 pi box) can make both
// sides claim (or disclaim) a flux-point pair -- the reference's MPI build then differs from its own serial run by 1e-8
// (measured with a METIS partition).  In a single-domain run the element with the lower id is the face's left side and
// ITS normal decides, so the ranks exchange (own mask, global element id) once and the side with the lower id wins: the
// partitioned run then makes exactly the serial run's choice.  Without global ids (ele_global_l == NULL) the masks are
// only compared, and any disagreement anywhere sends every rank back to the two-sided generation-6 kernels.
int hf_fused_after_nccl(hf_ctx *c)
{
  hf_fused_state *Z = c->fz;
  if (c->nproc < 2) return 0;
  c->nccl_reconciled = true;
  // the fused kernels exchange other halo data than the staged ones: either every rank uses them or none does (a rank
  // whose part of the mesh has boundary faces, say, cannot)
  double all_available = (Z && Z->available) ? 1.0 : 0.0;
  if (hf_halo_allreduce_min(c, &all_available)) return 1;
  if (all_available != 1.0)
  {
    if (Z && Z->available)
    {
      Z->available = false;
      Z->why = "another rank cannot use the fused kernels on its part of the mesh";
    }
    return 0;
  }
  if (!c->prm.viscous) return 0;
  hf_mpi_inters_dev &M = c->mpis[2];
  const int N = Z->order + 1, NN = N * N, nm = Z->n_mpi;
  const unsigned long long full = NN == 64 ? ~0ull : ((1ull << NN) - 1ull);
  const bool have_gid = (int)M.h_gid.size() == nm;
  double ok = 1.0, gid_everywhere = have_gid ? 1.0 : 0.0;
  if (hf_halo_allreduce_min(c, &gid_everywhere)) return 1;
  if (nm)
  {
    static_assert(sizeof(unsigned long long) == sizeof(double), "masks travel as 8-byte words");
    std::vector<double> mine(2 * (size_t)nm), theirs(2 * (size_t)nm);
    for (int i = 0; i < nm; i++)
    {
      memcpy(&mine[2 * (size_t)i], &Z->h_pmask[i], 8);
      mine[2 * (size_t)i + 1] = have_gid ? (double)M.h_gid[i] : -1.0;
    }
    double *d_out = nullptr, *d_in = nullptr;
    if (hf_alloc_copy(c, &d_out, mine.data(), mine.size())) return 1;
    if (hf_alloc_zero(c, &d_in, theirs.size())) return 1;
    if (hf_halo_post(c, M, d_out, d_in, 2) || hf_halo_wait(c)) return 1;
    HF_CUDA(cudaMemcpyAsync(theirs.data(), d_in, theirs.size() * 8, cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    bool changed = false;
    for (int i = 0; i < nm; i++)
    {
      unsigned long long their_mask;
      memcpy(&their_mask, &theirs[2 * (size_t)i], 8);
      // the complement of the neighbour's ownership, in this side's flux-point numbering
      unsigned long long want = 0ull;
      for (int j = 0; j < NN; j++)
        if (!((their_mask >> Z->T.perm[(M.h_rot[i] & 3) * 36 + j]) & 1ull)) want |= 1ull << j;
      if (want == (Z->h_pmask[i] & full)) continue;
      if (gid_everywhere == 1.0)
      {
        if ((double)M.h_gid[i] > theirs[2 * (size_t)i + 1])
        {
          // the neighbour is the serial run's left side: take its choice
          Z->h_pmask[i] = want;
          Z->h_bmask[(size_t)M.h_ele_l[i] * 6 + M.h_loc_l[i]] = want ^ Z->own_xor;
          changed = true;
        }
      }
      else
        ok = 0.0;
    }
    if (changed)
    {
      HF_CUDA(cudaMemcpy(Z->bmask, Z->h_bmask.data(), Z->h_bmask.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
      if (Z->gen9 && build_classes9(c, Z, Z->h_bmask, Z->h_finfo)) return 1; // the element classes follow the owner masks
    }
  }
  if (hf_halo_allreduce_min(c, &ok)) return 1;
  if (ok != 1.0 && Z->os)
  {
    Z->os = false;
    Z->gen9 = false;
    Z->fu_blk = NF * NN;
    Z->fv_blk = 4 * NN;
    Z->os_why = "the ranks of a partition face disagree on the LDG owner of a flux-point pair (rounding-level normal components) and no global element ids were given";
  }
  return 0;
}

extern "C" const char *hf_dev_fused_variant(hf_ctx *c)
{
  if (!c->fz || !c->fz->available) return "none";
  if (!c->prm.viscous) return "generation 6 (inviscid: k_resid6)";
  if (c->fz->gen9) return "generation 9 (one-sided LDG: k_face9 + k_resid9)";
  if (c->fz->os) return "generation 7 (one-sided LDG: k_grad7 + k_resid7)";
  static std::string s;
  s = "generation 6 (two-sided LDG: k_grad6 + k_resid6)" + (c->fz->os_why.empty() ? std::string() : ": " + c->fz->os_why);
  return s.c_str();
}

extern "C" const char *hf_dev_fused_status(hf_ctx *c)
{
  if (!c->fz) return "fused path not prepared";
  return c->fz->available ? "available" : c->fz->why.c_str();
}
