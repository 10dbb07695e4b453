// Fused residual path for affine hexahedra (BASELINE config 3: TGV, hex, P = 4).
//
// One RK stage = CalcResidual + AdvanceSolution (reference src/solver.cpp:50-223, src/eles.cpp:1080-1265) is done by
// two kernels per stage, each working element by element out of shared memory, with the operators applied in
// sum-factorised form (hex operators are tensor products with exact zeros, reference src/eles_hexas.cpp:224-282,
// 1132-1193, 1444-1537; the 1-D tables are read out of the dense matrices the host built, so they are the
// reference's numbers):
//
//   k_grad   (viscous only)   u -> own face values, LDG common solution with the neighbour's face values, corrected
//                             physical gradient at solution points, extrapolated to the faces -> fg
//                             = eles::calculate_gradient + the LDG half of int_inters::calculate_common_invFlux +
//                               eles::correct_gradient
//   k_resid                   the same gradient again (cheaper than storing it), inviscid + viscous flux at solution
//                             points, common fluxes on all six faces (Riemann + LDG, both sides of a face evaluate the
//                             same expression with the LEFT element's normal), divergence + correction, RK update,
//                             new face values -> fu[next]
//                             = evaluate_invFlux + evaluate_viscFlux + calculate_common_invFlux/viscFlux +
//                               extrapolate_totalFlux + calculate_divergence + calculate_corrected_divergence +
//                               AdvanceSolution + extrapolate_solution of the next stage
//
// Only face data crosses kernels: fu = u at flux points, fg = grad u at flux points, stored per face block
// [ele][face][(dim)][field][fpt] so that a neighbour's face is one contiguous run and a partition face's receive
// buffer is simply an extra block behind the last element (message layout = the reference's
// out_buffer_disu[inter][field][fpt], src/mpi_inters.cpp:226-229).
// Algorithmic traffic per element-stage, P = 4, RK34: see DESIGN.md (about 9 k doubles vs 56 k for the staged path).
#include "hf_device.h"
#include <cstring>
#include <cmath>
#include <algorithm>

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
  unsigned char perm[8 * 36]; // [rot + 4*is_right][j] -> neighbour's face-local flux point
};

struct rk_args
{
  int mode, copy_u1; // as k_rk_update in hf_device.cu
  double dt, fac, c1, c2;
};

struct fused_args
{
  int n_eles;
  const double *u0;
  double *u0_out;
  double *u1;
  double *div;            // written when keep_residual
  const double *fu_cur;   // face u, read (neighbours)
  double *fu_next;        // face u of the updated solution, written (own faces)
  double *fg;             // face gradients: written by k_grad, read by k_resid
  const double *em;       // [ele][EM]
  const int *nbr;         // [ele][6] neighbour face block
  const signed char *finfo; // [ele][6] rot + 4*is_right
  const signed char *bsign; // [ele][6*N*N] sign of ldg_beta
  const double *dt_local;
  const fused_tables *tab;
  hf_phys P;
  rk_args rk;
  int viscous, keep_residual, do_update;
};

template <int N, int E>
struct smem_layout
{
  static constexpr int NU = N * N * N, NFP = 6 * N * N;
  fused_tables tab;
  double su[E][NF][NU];
  double sg[E][ND][NF][NU];
  double sx[E][NF][NFP];
  double em[E][EM];
  int nbr[E][6];
  int finfo[E][6];
};

// ---- shared phases ----------------------------------------------------------------------------------------------------
template <int N, int E, int NT>
__device__ __forceinline__ void load_block(smem_layout<N, E> &S, const fused_args &A, int e0, int ne)
{
  constexpr int NU = N * N * N;
  const int tid = threadIdx.x;
  // tables
  {
    const double *src = (const double *)A.tab;
    double *dst = (double *)&S.tab;
    constexpr int nd = sizeof(fused_tables) / sizeof(double);
    for (int i = tid; i < nd; i += NT) dst[i] = src[i];
  }
  // solution: for a field, the ne elements of this block are contiguous in (upt, ele)
  for (int k = 0; k < NF; k++)
  {
    const double *src = A.u0 + (size_t)NU * (e0 + (size_t)A.n_eles * k);
    for (int i = tid; i < ne * NU; i += NT) S.su[i / NU][k][i % NU] = src[i];
  }
  for (int i = tid; i < ne * EM; i += NT) S.em[i / EM][i % EM] = A.em[(size_t)e0 * EM + i];
  for (int i = tid; i < ne * 6; i += NT)
  {
    S.nbr[i / 6][i % 6] = A.nbr[(size_t)e0 * 6 + i];
    S.finfo[i / 6][i % 6] = A.finfo[(size_t)e0 * 6 + i];
  }
}

// own face value of field k at face-local flux point j of face f: sum_i L[i] * su[line]
template <int N>
__device__ __forceinline__ double face_value(const double *field_upts, const double *L, int base, int stride)
{
  double acc = 0.0;
#pragma unroll
  for (int i = 0; i < N; i++) acc += L[i] * field_upts[base + i * stride];
  return acc;
}

// LDG common solution minus own value at every own flux point -> S.sx   (delta_disu_fpts of the reference)
template <int N, int E, int NT>
__device__ __forceinline__ void phase_delta(smem_layout<N, E> &S, const fused_args &A, int e0, int ne)
{
  constexpr int NFP = 6 * N * N, NN = N * N;
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const double *L = face_sgn(f) > 0 ? S.tab.Lp : S.tab.Lm;
    int info = S.finfo[e][f];
    int pj = S.tab.perm[info * 36 + j];
    const double *nb = A.fu_cur + (size_t)S.nbr[e][f] * (NF * NN) + pj;
    double beta = A.P.ldg_beta * (double)A.bsign[(size_t)(e0 + e) * NFP + r];
    bool is_right = info >= 4;
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      double uo = face_value<N>(S.su[e][k], L, base, stride);
      double un = nb[k * NN];
      double ul = is_right ? un : uo, ur = is_right ? uo : un;
      double uc = __dsub_rn(__dmul_rn(0.5, __dadd_rn(ul, ur)), __dmul_rn(beta, __dsub_rn(ul, ur)));
      S.sx[e][k][r] = uc - uo;
    }
  }
}

// corrected physical gradient at the solution points -> S.sg
template <int N, int E, int NT>
__device__ __forceinline__ void phase_gradient(smem_layout<N, E> &S, int ne)
{
  constexpr int NU = N * N * N, NN = N * N;
  for (int q = threadIdx.x; q < ne * NU; q += NT)
  {
    int e = q / NU, p = q - e * NU;
    int a = p % N, b = (p / N) % N, c = p / NN;
    const double *J = S.em[e];
    double inv_detjac = 1.0 / J[9];
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      const double *u = S.su[e][k];
      const double *dl = S.sx[e][k];
      double gt[3];
      {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) acc += S.tab.D[a * N + i] * u[i + N * b + NN * c];
        acc += S.tab.c5[2 * N + a] * dl[2 * NN + fpt_of_upt<N>(2, a, b, c)];
        acc += S.tab.c5[4 * N + a] * dl[4 * NN + fpt_of_upt<N>(4, a, b, c)];
        gt[0] = acc;
      }
      {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) acc += S.tab.D[b * N + i] * u[a + N * i + NN * c];
        acc += S.tab.c5[1 * N + b] * dl[1 * NN + fpt_of_upt<N>(1, a, b, c)];
        acc += S.tab.c5[3 * N + b] * dl[3 * NN + fpt_of_upt<N>(3, a, b, c)];
        gt[1] = acc;
      }
      {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) acc += S.tab.D[c * N + i] * u[a + N * b + NN * i];
        acc += S.tab.c5[0 * N + c] * dl[0 * NN + fpt_of_upt<N>(0, a, b, c)];
        acc += S.tab.c5[5 * N + c] * dl[5 * NN + fpt_of_upt<N>(5, a, b, c)];
        gt[2] = acc;
      }
      // physical gradient: g(d) = sum_l (1/detJ * gt(l)) * JGinv(l,d)    (reference src/eles.cpp:1955-2011)
#pragma unroll
      for (int d = 0; d < 3; d++)
      {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < 3; l++) acc += (inv_detjac * gt[l]) * J[l + 3 * d];
        S.sg[e][d][k][p] = acc;
      }
    }
  }
}

// ---- kernel 1: face gradients ------------------------------------------------------------------------------------------
template <int N, int E, int NT>
__global__ void __launch_bounds__(NT) k_grad(fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  smem_layout<N, E> &S = *reinterpret_cast<smem_layout<N, E> *>(smem_raw);
  constexpr int NFP = 6 * N * N, NN = N * N;
  const int e0 = blockIdx.x * E;
  const int ne = min(E, A.n_eles - e0);
  load_block<N, E, NT>(S, A, e0, ne);
  __syncthreads();
  phase_delta<N, E, NT>(S, A, e0, ne);
  __syncthreads();
  phase_gradient<N, E, NT>(S, ne);
  __syncthreads();
  // gradient at the own flux points (opp_6 then the transform; for an affine element the two commute)
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const double *L = face_sgn(f) > 0 ? S.tab.Lp : S.tab.Lm;
    double *out = A.fg + ((size_t)(e0 + e) * 6 + f) * (ND * NF * NN) + j;
#pragma unroll
    for (int d = 0; d < ND; d++)
#pragma unroll
      for (int k = 0; k < NF; k++) out[(d * NF + k) * NN] = face_value<N>(S.sg[e][d][k], L, base, stride);
  }
}

// ---- kernel 2: residual + RK update + next face values --------------------------------------------------------------------
template <int N, int E, int NT, bool VISC>
__global__ void __launch_bounds__(NT) k_resid(fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  smem_layout<N, E> &S = *reinterpret_cast<smem_layout<N, E> *>(smem_raw);
  constexpr int NU = N * N * N, NFP = 6 * N * N, NN = N * N;
  const int e0 = blockIdx.x * E;
  const int ne = min(E, A.n_eles - e0);
  load_block<N, E, NT>(S, A, e0, ne);
  __syncthreads();
  if (VISC)
  {
    phase_delta<N, E, NT>(S, A, e0, ne);
    __syncthreads();
    phase_gradient<N, E, NT>(S, ne);
    __syncthreads();
    // own-side viscous normal flux F_vis(u_own, grad_own) . n_left at every own flux point -> S.sx
    for (int q = threadIdx.x; q < ne * NFP; q += NT)
    {
      int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
      int base, stride;
      line_of_fpt<N>(f, j, base, stride);
      const double *L = face_sgn(f) > 0 ? S.tab.Lp : S.tab.Lm;
      double u[NF], g[NF * ND], fv[NF * ND], fn[NF];
#pragma unroll
      for (int k = 0; k < NF; k++) u[k] = face_value<N>(S.su[e][k], L, base, stride);
#pragma unroll
      for (int d = 0; d < ND; d++)
#pragma unroll
        for (int k = 0; k < NF; k++) g[k + NF * d] = face_value<N>(S.sg[e][d][k], L, base, stride);
      vis_flux<ND, NF>(u, g, fv, A.P);
      normal_flux<ND, NF>(fv, &S.em[e][10 + 4 * f + 1], fn);
#pragma unroll
      for (int k = 0; k < NF; k++) S.sx[e][k][r] = fn[k];
    }
    __syncthreads();
  }
  // transformed total flux at the solution points -> S.sg (overwrites the gradient point by point)
  for (int q = threadIdx.x; q < ne * NU; q += NT)
  {
    int e = q / NU, p = q - e * NU;
    const double *J = S.em[e];
    double u[NF], f[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++) u[k] = S.su[e][k][p];
    inv_flux<ND, NF>(u, f, A.P);
    double t[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
      for (int l = 0; l < ND; l++)
      {
        double acc = 0.0;
#pragma unroll
        for (int m = 0; m < ND; m++) acc += J[l + 3 * m] * f[k + NF * m];
        t[k + NF * l] = acc;
      }
    if (VISC)
    {
      double g[NF * ND];
#pragma unroll
      for (int d = 0; d < ND; d++)
#pragma unroll
        for (int k = 0; k < NF; k++) g[k + NF * d] = S.sg[e][d][k][p];
      vis_flux<ND, NF>(u, g, f, A.P);
#pragma unroll
      for (int k = 0; k < NF; k++)
#pragma unroll
        for (int l = 0; l < ND; l++)
        {
          double acc = t[k + NF * l];
#pragma unroll
          for (int m = 0; m < ND; m++) acc += J[l + 3 * m] * f[k + NF * m];
          t[k + NF * l] = acc;
        }
    }
#pragma unroll
    for (int l = 0; l < ND; l++)
#pragma unroll
      for (int k = 0; k < NF; k++) S.sg[e][l][k][p] = t[k + NF * l];
  }
  __syncthreads();
  // common flux minus own normal flux at every own flux point -> S.sx
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const int sgn = face_sgn(f), dir = face_dir(f);
    const double *L = sgn > 0 ? S.tab.Lp : S.tab.Lm;
    int info = S.finfo[e][f];
    bool is_right = info >= 4;
    int pj = S.tab.perm[info * 36 + j];
    size_t nblk = (size_t)S.nbr[e][f];
    const double *geo = &S.em[e][10 + 4 * f];
    const double tdA = geo[0];
    const double n[3] = {geo[1], geo[2], geo[3]};
    double uo[NF], un[NF], fn[NF];
    const double *nb = A.fu_cur + nblk * (NF * NN) + pj;
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      uo[k] = face_value<N>(S.su[e][k], L, base, stride);
      un[k] = nb[k * NN];
    }
    if (is_right) riemann<ND, NF>(un, uo, n, fn, A.P);
    else riemann<ND, NF>(uo, un, n, fn, A.P);
    if (VISC)
    {
      double g[NF * ND], fv[NF * ND], fvn[NF];
      const double *ng = A.fg + nblk * (ND * NF * NN) + pj;
#pragma unroll
      for (int q2 = 0; q2 < NF * ND; q2++) g[q2] = ng[q2 * NN];
      vis_flux<ND, NF>(un, g, fv, A.P);
      normal_flux<ND, NF>(fv, n, fvn);
      double beta = A.P.ldg_beta * (double)A.bsign[(size_t)(e0 + e) * NFP + r];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        double fo = S.sx[e][k][r];
        double fl = is_right ? fvn[k] : fo, fr = is_right ? fo : fvn[k];
        double ul = is_right ? un[k] : uo[k], ur = is_right ? uo[k] : un[k];
        fn[k] += ((0.5 + beta) * fl + (0.5 - beta) * fr) - A.P.ldg_tau * (ur - ul);
      }
    }
    const double s_side = is_right ? -tdA : tdA;
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      double ntd = face_value<N>(S.sg[e][dir][k], L, base, stride);
      S.sx[e][k][r] = fn[k] * s_side - (sgn > 0 ? ntd : -ntd);
    }
  }
  __syncthreads();
  // divergence + correction, RK update
  for (int q = threadIdx.x; q < ne * NU; q += NT)
  {
    int e = q / NU, p = q - e * NU;
    int a = p % N, b = (p / N) % N, c = p / NN;
    const int ge = e0 + e;
    const double detjac = S.em[e][9];
    const double dtl = A.dt_local ? A.dt_local[ge] : A.rk.dt;
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      double acc = 0.0;
#pragma unroll
      for (int i = 0; i < N; i++) acc += S.tab.D[a * N + i] * S.sg[e][0][k][i + N * b + NN * c];
#pragma unroll
      for (int i = 0; i < N; i++) acc += S.tab.D[b * N + i] * S.sg[e][1][k][a + N * i + NN * c];
#pragma unroll
      for (int i = 0; i < N; i++) acc += S.tab.D[c * N + i] * S.sg[e][2][k][a + N * b + NN * i];
      const double *dfl = S.sx[e][k];
      acc += S.tab.c3[0 * N + c] * dfl[0 * NN + fpt_of_upt<N>(0, a, b, c)];
      acc += S.tab.c3[1 * N + b] * dfl[1 * NN + fpt_of_upt<N>(1, a, b, c)];
      acc += S.tab.c3[2 * N + a] * dfl[2 * NN + fpt_of_upt<N>(2, a, b, c)];
      acc += S.tab.c3[3 * N + b] * dfl[3 * NN + fpt_of_upt<N>(3, a, b, c)];
      acc += S.tab.c3[4 * N + a] * dfl[4 * NN + fpt_of_upt<N>(4, a, b, c)];
      acc += S.tab.c3[5 * N + c] * dfl[5 * NN + fpt_of_upt<N>(5, a, b, c)];
      size_t gi = p + (size_t)NU * (ge + (size_t)A.n_eles * k);
      if (A.keep_residual) A.div[gi] = acc;
      if (A.do_update)
      {
        double u = S.su[e][k][p];
        double rr = acc / detjac;
        if (A.rk.copy_u1) A.u1[gi] = u;
        if (A.rk.mode == 0)
          u -= dtl / A.rk.fac * rr;
        else if (A.rk.mode == 1)
          u = A.rk.c1 * u + A.rk.c2 * A.u1[gi] + dtl / A.rk.fac * (-rr);
        else
        {
          double dlt = A.rk.c1 * A.u1[gi] + dtl * (-rr);
          A.u1[gi] = dlt;
          u += A.rk.c2 * dlt;
        }
        A.u0_out[gi] = u;
        S.su[e][k][p] = u;
      }
    }
  }
  if (!A.do_update) return;
  __syncthreads();
  // face values of the updated solution for the next stage (extrapolate_solution)
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const double *L = face_sgn(f) > 0 ? S.tab.Lp : S.tab.Lm;
    double *out = A.fu_next + ((size_t)(e0 + e) * 6 + f) * (NF * NN) + j;
#pragma unroll
    for (int k = 0; k < NF; k++) out[k * NN] = face_value<N>(S.su[e][k], L, base, stride);
  }
}

// face values of the current solution (first stage, or after an upload)
template <int N, int E, int NT>
__global__ void __launch_bounds__(NT) k_face_values(fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  smem_layout<N, E> &S = *reinterpret_cast<smem_layout<N, E> *>(smem_raw);
  constexpr int NFP = 6 * N * N, NN = N * N;
  const int e0 = blockIdx.x * E;
  const int ne = min(E, A.n_eles - e0);
  load_block<N, E, NT>(S, A, e0, ne);
  __syncthreads();
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const double *L = face_sgn(f) > 0 ? S.tab.Lp : S.tab.Lm;
    double *out = A.fu_next + ((size_t)(e0 + e) * 6 + f) * (NF * NN) + j;
#pragma unroll
    for (int k = 0; k < NF; k++) out[k * NN] = face_value<N>(S.su[e][k], L, base, stride);
  }
}

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
  double *fu[2] = {nullptr, nullptr};
  int cur = 0;
  double *fg = nullptr;
  double *em = nullptr;
  int *nbr = nullptr;
  signed char *finfo = nullptr, *bsign = nullptr;
  fused_tables *tab = nullptr;
  int *mpi_blk = nullptr; // [n_mpi] own face block of every partition interface
  double *out_u = nullptr, *out_g = nullptr;
  int E = 2, NT = 128;
};

void hf_fused_destroy(hf_ctx *c)
{
  delete c->fz;
  c->fz = nullptr;
}

// called from hf_dev_upload_eles while the host metric arrays are at hand
int hf_fused_on_upload(hf_ctx *c, hf_eles_dev &e, const hf_eles_desc *d)
{
  (void)c;
  if (e.ele_type != 4 || e.n_dims != 3) return 0;
  const int nu = e.n_upts, nf = e.n_fpts, ne = e.n_eles, nfi = nf / 6;
  e.h_em.assign((size_t)ne * 10, 0.);
  e.h_face_geo.assign((size_t)ne * 24, 0.);
  e.h_own_sign.assign((size_t)ne * nf, 1);
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
  return true;
}

int hf_fused_available(hf_ctx *c) { return c->fz && c->fz->available; }

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
  for (int t = 0; t < HF_N_INTER_TYPES; t++)
    if (c->bdys[t].n_inters) return no("boundary interfaces present (fused path handles interior and partition faces)");
  if (!e.affine) return no("elements are not affine (metric variation inside an element)");
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
  std::vector<signed char> finfo((size_t)ne * 6, 0), bsign((size_t)ne * NFP, 1);
  std::vector<double> em((size_t)ne * EM, 0.);
  for (int i = 0; i < ne; i++)
  {
    for (int q = 0; q < 10; q++) em[(size_t)i * EM + q] = e.h_em[(size_t)i * 10 + q];
    for (int f = 0; f < 6; f++)
      for (int q = 0; q < 4; q++) em[(size_t)i * EM + 10 + 4 * f + q] = e.h_face_geo[(size_t)i * 24 + 4 * f + q];
  }
  hf_int_inters_dev &I = c->ints[2];
  for (int i = 0; i < I.n_inters; i++)
  {
    int el = I.h_ele_l[i], fl = I.h_loc_l[i], er = I.h_ele_r[i], fr = I.h_loc_r[i], rot = I.h_rot[i];
    nbr[(size_t)el * 6 + fl] = er * 6 + fr;
    nbr[(size_t)er * 6 + fr] = el * 6 + fl;
    finfo[(size_t)el * 6 + fl] = (signed char)rot;
    finfo[(size_t)er * 6 + fr] = (signed char)(rot + 4);
    // the right element uses the left element's normal; its own tdA stays
    for (int q = 1; q < 4; q++) em[(size_t)er * EM + 10 + 4 * fr + q] = e.h_face_geo[(size_t)el * 24 + 4 * fl + q];
    for (int j = 0; j < NN; j++)
    {
      signed char s = e.h_own_sign[(size_t)el * NFP + fl * NN + j];
      bsign[(size_t)el * NFP + fl * NN + j] = s;
      bsign[(size_t)er * NFP + fr * NN + T.perm[rot * 36 + j]] = s;
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
    finfo[(size_t)el * 6 + fl] = (signed char)M.h_rot[i];
    mpi_blk[i] = el * 6 + fl;
    for (int j = 0; j < NN; j++) bsign[(size_t)el * NFP + fl * NN + j] = e.h_own_sign[(size_t)el * NFP + fl * NN + j];
  }
  for (size_t q = 0; q < nbr.size(); q++)
    if (nbr[q] < 0) return no("an element face has no neighbour");
  Z->order = e.order;
  Z->n_eles = ne;
  const size_t nblk = (size_t)ne * 6 + M.n_inters;
  if (hf_alloc_zero(c, &Z->fu[0], nblk * NF * NN)) return 1;
  if (hf_alloc_zero(c, &Z->fu[1], nblk * NF * NN)) return 1;
  if (visc && hf_alloc_zero(c, &Z->fg, nblk * ND * NF * NN)) return 1;
  if (hf_alloc_copy(c, &Z->em, em.data(), em.size())) return 1;
  if (hf_alloc_copy(c, &Z->nbr, nbr.data(), nbr.size())) return 1;
  if (hf_alloc_copy(c, &Z->finfo, finfo.data(), finfo.size())) return 1;
  if (hf_alloc_copy(c, &Z->bsign, bsign.data(), bsign.size())) return 1;
  if (hf_alloc_copy(c, &Z->tab, &T, 1)) return 1;
  if (hf_alloc_copy(c, &Z->mpi_blk, mpi_blk.data(), mpi_blk.size())) return 1;
  if (M.n_inters)
  {
    if (hf_alloc_zero(c, &Z->out_u, (size_t)M.n_inters * NF * NN)) return 1;
    if (visc && hf_alloc_zero(c, &Z->out_g, (size_t)M.n_inters * ND * NF * NN)) return 1;
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
template <int N, int E, int NT>
int launch_all(hf_ctx *c, hf_fused_state *Z, fused_args &A, int what)
{
  // what: 0 face values, 1 gradient kernel, 2 residual kernel
  const size_t smem = sizeof(smem_layout<N, E>);
  const int grid = (Z->n_eles + E - 1) / E;
  static bool attr_done[3] = {false, false, false};
  if (what == 0)
  {
    if (!attr_done[0]) { HF_CUDA(cudaFuncSetAttribute(k_face_values<N, E, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[0] = true; }
    k_face_values<N, E, NT><<<grid, NT, smem, c->stream>>>(A);
  }
  else if (what == 1)
  {
    if (!attr_done[1]) { HF_CUDA(cudaFuncSetAttribute(k_grad<N, E, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_done[1] = true; }
    k_grad<N, E, NT><<<grid, NT, smem, c->stream>>>(A);
  }
  else
  {
    if (!attr_done[2])
    {
      HF_CUDA(cudaFuncSetAttribute(k_resid<N, E, NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      HF_CUDA(cudaFuncSetAttribute(k_resid<N, E, NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_done[2] = true;
    }
    hf_ktimer_begin(c);
    if (A.viscous) k_resid<N, E, NT, true><<<grid, NT, smem, c->stream>>>(A);
    else k_resid<N, E, NT, false><<<grid, NT, smem, c->stream>>>(A);
    hf_ktimer_end(c);
  }
  c->launches++;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { hf_set_error(std::string("fused kernel launch: ") + cudaGetErrorString(err)); return 1; }
  return 0;
}

int launch(hf_ctx *c, hf_fused_state *Z, fused_args &A, int what)
{
  switch (Z->order)
  {
  case 1: return launch_all<2, 8, 128>(c, Z, A, what);
  case 2: return launch_all<3, 4, 128>(c, Z, A, what);
  case 3: return launch_all<4, 2, 128>(c, Z, A, what);
  case 4: return launch_all<5, 2, 128>(c, Z, A, what);
  case 5: return launch_all<6, 1, 128>(c, Z, A, what);
  }
  hf_set_error("fused path: unsupported order");
  return 1;
}

void base_args(hf_ctx *c, hf_fused_state *Z, fused_args &A)
{
  hf_eles_dev &e = c->eles[4];
  memset(&A, 0, sizeof(A));
  A.n_eles = e.n_eles;
  A.u0 = e.disu_upts[0];
  A.u0_out = e.disu_upts[0];
  A.u1 = e.disu_upts[1];
  A.div = e.div_tconf_upts;
  A.fu_cur = Z->fu[Z->cur];
  A.fu_next = Z->fu[Z->cur ^ 1];
  A.fg = Z->fg;
  A.em = Z->em;
  A.nbr = Z->nbr;
  A.finfo = Z->finfo;
  A.bsign = Z->bsign;
  A.dt_local = (c->prm.dt_type == 2) ? e.dt_local : nullptr;
  A.tab = Z->tab;
  A.P = c->phys;
  A.viscous = c->prm.viscous;
}

// exchange the partition-face blocks of arr (blk_doubles each): pack -> ncclSend/Recv into the tail of arr
int exchange(hf_ctx *c, hf_fused_state *Z, double *arr, double *out, int blk_doubles)
{
  if (Z->n_mpi == 0) return 0;
  hf_mpi_inters_dev &M = c->mpis[2];
  long long n = (long long)Z->n_mpi * blk_doubles;
  k_pack_blocks<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(arr, Z->mpi_blk, out, Z->n_mpi, blk_doubles);
  c->launches++;
  if (hf_halo_post(c, M, out, arr + (size_t)Z->n_eles * 6 * blk_doubles, (size_t)blk_doubles)) return 1;
  return hf_halo_wait(c);
}
} // namespace

int hf_fused_extrapolate(hf_ctx *c)
{
  hf_fused_state *Z = c->fz;
  if (!Z || !Z->available) HF_FAIL("fused path not available");
  HF_CUDA(cudaSetDevice(c->device));
  fused_args A;
  base_args(c, Z, A);
  A.fu_next = Z->fu[Z->cur]; // fill the current buffer
  if (launch(c, Z, A, 0)) return 1;
  const int NN = (Z->order + 1) * (Z->order + 1);
  if (exchange(c, Z, Z->fu[Z->cur], Z->out_u, NF * NN)) return 1;
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
  const int NN = (Z->order + 1) * (Z->order + 1);
  fused_args A;
  base_args(c, Z, A);
  A.keep_residual = keep_residual;
  A.do_update = do_update;
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
  if (p.viscous)
  {
    if (launch(c, Z, A, 1)) return 1;
    if (exchange(c, Z, Z->fg, Z->out_g, ND * NF * NN)) return 1;
  }
  if (launch(c, Z, A, 2)) return 1;
  if (do_update)
  {
    Z->cur ^= 1;
    if (exchange(c, Z, Z->fu[Z->cur], Z->out_u, NF * NN)) return 1;
    c->ufpts_valid = true;
  }
  return 0;
}

extern "C" const char *hf_dev_fused_status(hf_ctx *c)
{
  if (!c->fz) return "fused path not prepared";
  return c->fz->available ? "available" : c->fz->why.c_str();
}
