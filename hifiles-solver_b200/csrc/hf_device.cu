// Device layer, part 1: context life cycle, uploads, the staged (one kernel per reference method) path, RK update,
// residual norms and CFL time step.  The staged kernels reproduce the reference's operation order method by
// method and work for every element type; the fused tensor-product kernels live in hf_fused.cu.
//   CalcResidual sequence           reference src/solver.cpp:50-223
//   operator products (dgemm order) reference src/funcs.cpp:49-124
//   pointwise flux + transform      reference src/eles.cpp:1415-1478, 2285-2392
//   gradient correction/transform   reference src/eles.cpp:1890-2052
//   interface loops                 reference src/int_inters.cpp:160-343, src/bdy_inters.cpp:213-338, 1024-1136,
//                                   src/mpi_inters.cpp:218-575
//   RK update                       reference src/eles.cpp:1080-1265
#include "hf_device.h"
#include "hf_bc.cuh"
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <algorithm>

static thread_local std::string g_err;
void hf_set_error(const std::string &msg) { g_err = msg; }
const std::string &hf_get_error() { return g_err; }

#define HF_FAIL(msg)            \
  do {                          \
    hf_set_error(msg);          \
    return 1;                   \
  } while (0)

#define HF_LAUNCH_CHECK(c)                                                                              \
  do {                                                                                                  \
    (c)->launches++;                                                                                    \
    cudaError_t e_ = cudaGetLastError();                                                                \
    if (e_ != cudaSuccess) { hf_set_error(std::string("kernel launch: ") + cudaGetErrorString(e_)); return 1; } \
  } while (0)

#define HF_DISPATCH(nd, nf, ...)                                                      \
  do {                                                                                \
    if ((nd) == 3 && (nf) == 5) { constexpr int ND = 3, NF = 5; __VA_ARGS__; }         \
    else if ((nd) == 2 && (nf) == 4) { constexpr int ND = 2, NF = 4; __VA_ARGS__; }    \
    else if ((nd) == 2 && (nf) == 1) { constexpr int ND = 2, NF = 1; __VA_ARGS__; }    \
    else if ((nd) == 3 && (nf) == 1) { constexpr int ND = 3, NF = 1; __VA_ARGS__; }    \
    else HF_FAIL("unsupported (n_dims, n_fields) combination");                       \
  } while (0)

static inline unsigned hf_blocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

// ---------------------------------------------------------------------------------------------------------------------
// kernels: small-operator products
// ---------------------------------------------------------------------------------------------------------------------
struct hf_ell3
{
  int n, rows, cols;
  int nnz[3];
  const double *val[3];
  const int *col[3];
  const double *dval[3]; // dense copies (null for sparse operators)
};

// out(row, c) = [out(row, c)] + sum_d sum_k E_d(row, k) * in_d(col_d(row,k), c): one running sum in ascending column
// order, i.e. the order of the reference's column-by-column dgemm with beta = 0 / 1.
template <bool ACC>
__global__ void k_op_apply(hf_ell3 E, const double *__restrict__ in, size_t in_dim_stride, double *__restrict__ out, long long n_cols)
{
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= (long long)E.rows * n_cols) return;
  int row = (int)(idx % E.rows);
  long long cg = idx / E.rows;
  double acc = ACC ? out[idx] : 0.0;
  for (int d = 0; d < E.n; d++)
  {
    const double *src = in + d * in_dim_stride + cg * E.cols;
    const double *val = E.val[d] + row;
    const int *col = E.col[d] + row;
    for (int k = 0; k < E.nnz[d]; k++) acc += val[(size_t)k * E.rows] * src[col[(size_t)k * E.rows]];
  }
  out[idx] = acc;
}


// Dense operators (triangles, tetrahedra, prisms: Dubiner-basis matrices without exact zeros) as FP64 tensor-core tiles.  The
// product out(rows, c) = sum_d Op_d(rows, cols) in_d(cols, c) over n_cols = n_fields * n_eles columns is a batched GEMM with a tiny
// left factor: 2 rows cols flops per 8 (rows + cols) bytes, 3 - 7 flop / byte at P = 3 -- near the FP64 : HBM balance of the B200,
// so the kernel must neither re-read the operator per column nor spend one shared-memory load per FMA as the thread-per-output
// kernel does.  One CTA takes 64 columns: the input tile is staged once in shared memory (coalesced), every warp owns 8-row blocks
// of the operator and walks the k dimension in steps of 4: one operator fragment (a global load that hits L1) feeds eight
// mma.sync.m8n8k4.f64 against eight column blocks -- 9 loads per 2 048 FMA.  The accumulation order inside a tile is the tensor
// core's, not the reference's ascending column order, and products are fused: results differ from k_op_apply in the last bits
// (1e-16 relative), which is why this path belongs to the fast mode (hf_dev_set_mode(ctx, 1)) and the bit-exact yardstick
// (mode 0) keeps k_op_apply.
constexpr int OPD_TC = 64;
__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <bool ACC>
__global__ void __launch_bounds__(128) k_op_dense_mma(hf_ell3 E, const double *__restrict__ in, size_t in_dim_stride, double *__restrict__ out, long long n_cols,
                                                      int cols_pad)
{
  extern __shared__ double sm_in[]; // [E.n][OPD_TC][cols_pad], zero padded
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long c0 = (long long)blockIdx.x * OPD_TC;
  const int nc = (int)min((long long)OPD_TC, n_cols - c0);
  for (int d = 0; d < E.n; d++)
  {
    const double *src = in + d * in_dim_stride + c0 * E.cols;
    double *dst = sm_in + (size_t)d * OPD_TC * cols_pad;
    for (int i = tid; i < OPD_TC * cols_pad; i += 128)
    {
      const int n = i / cols_pad, k = i - n * cols_pad;
      dst[i] = (n < nc && k < E.cols) ? src[(size_t)n * E.cols + k] : 0.0;
    }
  }
  __syncthreads();
  const int kb = (E.cols + 3) / 4, rb = (E.rows + 7) / 8;
  const int ar = lane >> 2, ak = lane & 3; // fragment coordinates: A(row ar, k ak), B(k ak, column ar), C(row ar, columns 2 ak, 2 ak + 1)
  for (int r = warp; r < rb; r += 4)
  {
    const int row = r * 8 + ar;
    const bool row_ok = row < E.rows;
    double acc[OPD_TC / 8][2];
#pragma unroll
    for (int nb = 0; nb < OPD_TC / 8; nb++)
    {
#pragma unroll
      for (int h = 0; h < 2; h++)
      {
        const int col = nb * 8 + 2 * ak + h;
        acc[nb][h] = (ACC && row_ok && col < nc) ? out[row + (size_t)E.rows * (c0 + col)] : 0.0;
      }
    }
    for (int d = 0; d < E.n; d++)
    {
      const double *val = E.dval[d];
      const double *b0 = sm_in + (size_t)d * OPD_TC * cols_pad + (size_t)ar * cols_pad + ak;
      for (int kk = 0; kk < kb; kk++)
      {
        const int k = kk * 4 + ak;
        const double a = (row_ok && k < E.cols) ? val[(size_t)k * E.rows + row] : 0.0;
#pragma unroll
        for (int nb = 0; nb < OPD_TC / 8; nb++) dmma884(acc[nb][0], acc[nb][1], a, b0[(size_t)nb * 8 * cols_pad + kk * 4]);
      }
    }
    if (row_ok)
    {
#pragma unroll
      for (int nb = 0; nb < OPD_TC / 8; nb++)
#pragma unroll
        for (int h = 0; h < 2; h++)
        {
          const int col = nb * 8 + 2 * ak + h;
          if (col < nc) out[row + (size_t)E.rows * (c0 + col)] = acc[nb][h];
        }
    }
  }
}

// Persson's modal sensor and the exponential filter, one CTA per element (reference src/eles.cpp:2918-2959 and
// <type>::shock_det_persson): uhat = V^-1 u_field, sensor = sum_top w uhat^2 / sum_all w uhat^2 (both sums in ascending
// mode order, as the reference's loops); if sensor >= s0 every field of the element is replaced by exp_filter * u.
__global__ void k_shock_capture(int n_upts, int n_eles, int n_fields, int det_field, double s0, double *__restrict__ u, const double *__restrict__ inv_vdm,
                                const double *__restrict__ w_top, const double *__restrict__ w_all, const double *__restrict__ filt,
                                double *__restrict__ sensor)
{
  extern __shared__ double sh[]; // [n_upts] squared modal values, then [n_upts * n_fields] the element's solution
  double *m2 = sh, *us = sh + n_upts;
  __shared__ int flagged;
  const int e = blockIdx.x;
  const size_t fs = (size_t)n_upts * n_eles;
  for (int q = threadIdx.x; q < n_upts * n_fields; q += blockDim.x)
  {
    const int k = q / n_upts, j = q - k * n_upts;
    us[q] = u[j + (size_t)n_upts * e + k * fs];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < n_upts; j += blockDim.x)
  {
    double acc = 0.0;
    for (int l = 0; l < n_upts; l++) acc += (1.0 * us[l + det_field * n_upts]) * inv_vdm[j + (size_t)n_upts * l];
    m2[j] = acc * acc;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    double top = 0, all = 0.;
    for (int j = 0; j < n_upts; j++)
    {
      if (w_top[j] != 0.) top += m2[j] * w_top[j];
      all = all + w_all[j] * m2[j];
    }
    const double sv = top / all;
    sensor[e] = sv;
    flagged = sv >= s0;
  }
  __syncthreads();
  if (!flagged) return;
  for (int q = threadIdx.x; q < n_upts * n_fields; q += blockDim.x)
  {
    const int k = q / n_upts, j = q - k * n_upts;
    double acc = 0.0;
    for (int l = 0; l < n_upts; l++) acc += (1.0 * us[l + k * n_upts]) * filt[j + (size_t)n_upts * l];
    u[j + (size_t)n_upts * e + k * fs] = acc;
  }
}

// y -= x  (the daxpy of calculate_corrected_divergence, reference src/eles.cpp:1746-1750)
__global__ void k_sub(double *__restrict__ y, const double *__restrict__ x, long long n)
{
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx < n) y[idx] -= x[idx];
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels: pointwise work at solution / flux points
// ---------------------------------------------------------------------------------------------------------------------
template <int ND, int NF, bool VISC>
__global__ void k_point_flux(long long n_pts, const double *__restrict__ u, const double *__restrict__ grad, const double *__restrict__ JGinv,
                             double *__restrict__ tdisf, hf_phys P)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double uu[NF], f[NF * ND], J[ND * ND];
#pragma unroll
  for (int k = 0; k < NF; k++) uu[k] = u[p + k * n_pts];
#pragma unroll
  for (int q = 0; q < ND * ND; q++) J[q] = JGinv[p * (ND * ND) + q];
  if (VISC)
  {
    double g[NF * ND];
#pragma unroll
    for (int q = 0; q < NF * ND; q++) g[q] = grad[p + q * n_pts];
    vis_flux<ND, NF>(uu, g, f, P);
  }
  else
    inv_flux<ND, NF>(uu, f, P);
#pragma unroll
  for (int k = 0; k < NF; k++)
#pragma unroll
    for (int l = 0; l < ND; l++)
    {
      double acc = VISC ? tdisf[p + (k + NF * l) * n_pts] : 0.0;
#pragma unroll
      for (int m = 0; m < ND; m++) acc += J[l + ND * m] * f[k + NF * m];
      tdisf[p + (k + NF * l) * n_pts] = acc;
    }
}


// Sub-grid-scale flux of the eddy-viscosity models at one solution point (eles::calc_sgsf_upts, reference
// src/eles.cpp:2395-2646; Smagorinsky with wall damping, WALE), same operation order.  sf(k,d) = sf[k + NF*d].
struct hf_les
{
  int sgs_model, order;
  double C_s, Kappa, prandtl_t, filter_ratio, vol_factor, gamma;
};
template <int ND, int NF>
__device__ __forceinline__ void sgs_flux_eddy(const double *u, const double *g, double detjac, const double *wd, const hf_les &Q, double *sf)
{
  const double rho = u[0];
  double v[ND], ke = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    v[i] = u[i + 1] / rho;
    ke += 0.5 * (v[i] * v[i]);
  }
  const double inte = u[NF - 1] / rho - ke;
  const double vol = detjac * Q.vol_factor;
  const double delta = Q.filter_ratio * pow(vol, 1. / ND) / (Q.order + 1.);
  double drho[ND], dene[ND], dke[ND], de[ND], du[ND][ND], S[ND][ND];
#pragma unroll
  for (int i = 0; i < ND; i++) { drho[i] = g[0 + NF * i]; dene[i] = g[(NF - 1) + NF * i]; }
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    dke[i] = ke * drho[i];
#pragma unroll
    for (int j = 0; j < ND; j++)
    {
      du[i][j] = (g[(j + 1) + NF * i] - v[j] * drho[i]) / rho; // du_j/dx_i
      dke[i] += rho * v[j] * du[i][j];
    }
    de[i] = (dene[i] - dke[i] - drho[i] * inte) / rho;
  }
#pragma unroll
  for (int i = 0; i < ND; i++)
#pragma unroll
    for (int j = 0; j < ND; j++) S[i][j] = (du[i][j] + du[j][i]) / 2.0;
  double mu_t;
  if (Q.sgs_model == 0)
  {
    double y = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++) y += wd[i] * wd[i];
    y = sqrt(y);
    double Smod = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
      for (int j = 0; j < ND; j++) Smod += 2.0 * S[i][j] * S[i][j];
    Smod = sqrt(Smod);
    mu_t = rho * fmin(y * y * Q.Kappa * Q.Kappa, Q.C_s * Q.C_s * delta * delta) * Smod;
  }
  else
  {
    // WALE: square of the velocity-gradient tensor, symmetrised, trace removed
    double gbt[ND][ND], Sq[ND][ND];
#pragma unroll
    for (int j = 0; j < ND; j++)
#pragma unroll
      for (int i = 0; i < ND; i++)
      {
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < ND; l++) acc += (1.0 * du[j][l]) * du[l][i]; // column-major product of the (i,j) = du_j/dx_i array with itself
        gbt[j][i] = acc;
      }
    // gbt[j][i] holds element (i,j) of the product; its transpose is g_bar
    double diag = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
      for (int j = 0; j < ND; j++) Sq[i][j] = (0.0 + 0.5 * gbt[i][j]) + 0.5 * gbt[j][i];
#pragma unroll
    for (int i = 0; i < ND; i++) diag += gbt[i][i] / 3.0;
#pragma unroll
    for (int i = 0; i < ND; i++) Sq[i][i] -= diag;
    double num = 0.0, denom = 0.0;
#pragma unroll
    for (int i = 0; i < ND; i++)
#pragma unroll
      for (int j = 0; j < ND; j++)
      {
        num += Sq[i][j] * Sq[i][j];
        denom += S[i][j] * S[i][j];
      }
    denom = pow(denom, 2.5) + pow(num, 1.25);
    num = pow(num, 1.5);
    mu_t = rho * Q.C_s * Q.C_s * delta * delta * num / (denom + 1.e-12);
  }
  double diag = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++) diag += S[i][i] / 3.0;
#pragma unroll
  for (int i = 0; i < ND; i++) S[i][i] -= diag;
#pragma unroll
  for (int j = 0; j < ND; j++)
  {
    sf[0 + NF * j] = 0.0;
    double e = -1.0 * Q.gamma * mu_t / Q.prandtl_t * de[j];
#pragma unroll
    for (int k = 0; k < ND; k++) e -= v[k] * 2.0 * mu_t * S[k][j];
    sf[(NF - 1) + NF * j] = e;
#pragma unroll
    for (int i = 1; i < NF - 1; i++) sf[i + NF * j] = -2.0 * mu_t * S[i - 1][j];
  }
}

template <int ND, int NF>
__device__ __forceinline__ void sgs_flux(const double *u, const double *g, double detjac, const double *wd, const hf_les &Q, const double *Lu, const double *Le,
                                         double *sf)
{
  const bool eddy = Q.sgs_model <= 2, sim = Q.sgs_model == 2 || Q.sgs_model == 4;
#pragma unroll
  for (int q = 0; q < NF * ND; q++) sf[q] = 0.;
  if (eddy) sgs_flux_eddy<ND, NF>(u, g, detjac, wd, Q, sf);
  if (sim)
  {
    // scale-similarity term from the Leonard tensors (reference src/eles.cpp:2612-2644)
    const double rho = u[0];
#pragma unroll
    for (int j = 0; j < ND; j++) sf[(NF - 1) + NF * j] += Q.gamma * rho * Le[j];
    if (ND == 2)
    {
      sf[1 + NF * 0] += rho * Lu[0];
      sf[1 + NF * 1] += rho * Lu[2];
      sf[2 + NF * 0] += sf[1 + NF * 1];
      sf[2 + NF * 1] += rho * Lu[1];
    }
    else
    {
      sf[1 + NF * 0] += rho * Lu[0];
      sf[1 + NF * 1] += rho * Lu[3];
      sf[1 + NF * 2] += rho * Lu[4];
      sf[2 + NF * 0] += sf[1 + NF * 1];
      sf[2 + NF * 1] += rho * Lu[1];
      sf[2 + NF * 2] += rho * Lu[5];
      sf[3 % NF + NF * 0] += sf[1 + NF * 2];
      sf[3 % NF + NF * 1] += sf[2 + NF * 2];
      sf[3 % NF + NF * 2] += rho * Lu[2];
    }
  }
}

// eles::calc_sgs_terms (reference src/eles.cpp:2058-2200): products of the unfiltered solution, then Leonard tensors
template <int ND, int NF>
__global__ void k_sgs_products(long long n_pts, const double *__restrict__ u, double *__restrict__ uu, double *__restrict__ ue)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double t[NF];
#pragma unroll
  for (int k = 0; k < NF; k++) t[k] = u[p + k * n_pts];
  const double rsq = t[0] * t[0];
  if (ND == 2)
  {
    uu[p] = t[1] * t[1] / rsq;
    uu[p + n_pts] = t[2] * t[2] / rsq;
    uu[p + 2 * n_pts] = t[1] * t[2] / rsq;
    t[NF - 1] -= 0.5 * (t[1] * t[1] + t[2] * t[2]) / t[0];
  }
  else
  {
    uu[p] = t[1] * t[1] / rsq;
    uu[p + n_pts] = t[2] * t[2] / rsq;
    uu[p + 2 * n_pts] = t[3 % NF] * t[3 % NF] / rsq;
    uu[p + 3 * n_pts] = t[1] * t[2] / rsq;
    uu[p + 4 * n_pts] = t[1] * t[3 % NF] / rsq;
    uu[p + 5 * n_pts] = t[2] * t[3 % NF] / rsq;
    t[NF - 1] -= 0.5 * (t[1] * t[1] + t[2] * t[2] + t[3 % NF] * t[3 % NF]) / t[0];
  }
#pragma unroll
  for (int d = 0; d < ND; d++) ue[p + d * n_pts] = t[d + 1] * t[NF - 1] / rsq;
}
template <int ND, int NF>
__global__ void k_sgs_leonard(long long n_pts, const double *__restrict__ uf, double *__restrict__ Lu, double *__restrict__ Le)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double t[NF];
#pragma unroll
  for (int k = 0; k < NF; k++) t[k] = uf[p + k * n_pts];
  const double rsq = t[0] * t[0];
  double diag;
  if (ND == 2)
  {
    Lu[p] -= (t[1] * t[1]) / rsq;
    Lu[p + n_pts] -= (t[2] * t[2]) / rsq;
    Lu[p + 2 * n_pts] -= (t[1] * t[2]) / rsq;
    diag = (Lu[p] + Lu[p + n_pts]) / 3.0;
    t[NF - 1] -= 0.5 * (t[1] * t[1] + t[2] * t[2]) / t[0];
  }
  else
  {
    Lu[p] -= (t[1] * t[1]) / rsq;
    Lu[p + n_pts] -= (t[2] * t[2]) / rsq;
    Lu[p + 2 * n_pts] -= (t[3 % NF] * t[3 % NF]) / rsq;
    Lu[p + 3 * n_pts] -= (t[1] * t[2]) / rsq;
    Lu[p + 4 * n_pts] -= (t[1] * t[3 % NF]) / rsq;
    Lu[p + 5 * n_pts] -= (t[2] * t[3 % NF]) / rsq;
    diag = (Lu[p] + Lu[p + n_pts] + Lu[p + 2 * n_pts]) / 3.0;
    t[NF - 1] -= 0.5 * (t[1] * t[1] + t[2] * t[2] + t[3 % NF] * t[3 % NF]) / t[0];
  }
#pragma unroll
  for (int d = 0; d < ND; d++) Le[p + d * n_pts] = (Le[p + d * n_pts] - t[d + 1] * t[NF - 1]) / rsq;
#pragma unroll
  for (int d = 0; d < ND; d++) Lu[p + d * n_pts] -= diag;
}

// eles::evaluate_viscFlux with LES (reference src/eles.cpp:2285-2392): viscous + SGS flux, the transformed SGS flux
// alone goes to sgsf_upts
template <int ND, int NF>
__global__ void k_point_flux_les(long long n_pts, int n_upts, const double *__restrict__ u, const double *__restrict__ grad, const double *__restrict__ JGinv,
                                 const double *__restrict__ detjac, const double *__restrict__ wall_distance, const double *__restrict__ Lu,
                                 const double *__restrict__ Le, double *__restrict__ tdisf, double *__restrict__ sgsf, hf_phys P, hf_les Q)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double uu[NF], f[NF * ND], sf[NF * ND], g[NF * ND], J[ND * ND], wd[ND];
#pragma unroll
  for (int k = 0; k < NF; k++) uu[k] = u[p + k * n_pts];
#pragma unroll
  for (int q = 0; q < ND * ND; q++) J[q] = JGinv[p * (ND * ND) + q];
#pragma unroll
  for (int q = 0; q < NF * ND; q++) g[q] = grad[p + q * n_pts];
#pragma unroll
  for (int d = 0; d < ND; d++) wd[d] = wall_distance ? wall_distance[p + d * n_pts] : 1e20;
  vis_flux<ND, NF>(uu, g, f, P);
  double lu[6] = {0, 0, 0, 0, 0, 0}, le[3] = {0, 0, 0};
  if (Lu)
  {
#pragma unroll
    for (int q = 0; q < (ND == 2 ? 3 : 6); q++) lu[q] = Lu[p + q * n_pts];
#pragma unroll
    for (int q = 0; q < ND; q++) le[q] = Le[p + q * n_pts];
  }
  sgs_flux<ND, NF>(uu, g, detjac[p], wd, Q, lu, le, sf);
#pragma unroll
  for (int q = 0; q < NF * ND; q++) f[q] = f[q] + 1.0 * sf[q];
#pragma unroll
  for (int k = 0; k < NF; k++)
#pragma unroll
    for (int l = 0; l < ND; l++)
    {
      double s = 0.0, acc = tdisf[p + (k + NF * l) * n_pts];
#pragma unroll
      for (int m = 0; m < ND; m++)
      {
        s += J[l + ND * m] * sf[k + NF * m];
        acc += J[l + ND * m] * f[k + NF * m];
      }
      sgsf[p + (k + NF * l) * n_pts] = s;
      tdisf[p + (k + NF * l) * n_pts] = acc;
    }
  (void)n_upts;
}

// transformed SGS flux at the flux points back to physical space: f = (1/detJ) J F (eles::extrapolate_sgsFlux, reference
// src/eles.cpp:2864-2893)
template <int ND, int NF>
__global__ void k_sgsf_physical(long long n_pts, double *__restrict__ sgsf, const double *__restrict__ detjac, const double *__restrict__ Jac)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  const double inv_detjac = 1.0 / detjac[p];
  double J[ND * ND], t[NF * ND];
#pragma unroll
  for (int q = 0; q < ND * ND; q++) J[q] = Jac[p * (ND * ND) + q]; // (a,b) = a + ND*b
#pragma unroll
  for (int q = 0; q < NF * ND; q++) t[q] = sgsf[p + q * n_pts];
#pragma unroll
  for (int k = 0; k < NF; k++)
#pragma unroll
    for (int d = 0; d < ND; d++)
    {
      double acc = 0.0;
#pragma unroll
      for (int m = 0; m < ND; m++) acc += (inv_detjac * t[k + NF * m]) * J[d + ND * m];
      sgsf[p + (k + NF * d) * n_pts] = acc;
    }
}

// reference-space gradient -> physical gradient: g(d,k) = sum_l (1/detJ * gt(l,k)) * JGinv(l,d)
template <int ND, int NF>
__global__ void k_transform_grad(long long n_pts, double *__restrict__ grad, const double *__restrict__ detjac, const double *__restrict__ JGinv)
{
  long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (p >= n_pts) return;
  double J[ND * ND];
#pragma unroll
  for (int q = 0; q < ND * ND; q++) J[q] = JGinv[p * (ND * ND) + q];
  double inv_detjac = 1.0 / detjac[p];
#pragma unroll
  for (int k = 0; k < NF; k++)
  {
    double gt[ND], g[ND];
#pragma unroll
    for (int l = 0; l < ND; l++) gt[l] = grad[p + (k + NF * l) * n_pts];
#pragma unroll
    for (int d = 0; d < ND; d++)
    {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < ND; l++) acc += (inv_detjac * gt[l]) * J[l + ND * d];
      g[d] = acc;
    }
#pragma unroll
    for (int d = 0; d < ND; d++) grad[p + (k + NF * d) * n_pts] = g[d];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels: interfaces (one thread per flux-point pair)
// ---------------------------------------------------------------------------------------------------------------------
template <int ND, int NF>
__device__ __forceinline__ void load_fpt(const hf_ele_view &V, int idx, double *u)
{
  size_t s = (size_t)V.n_fpts * V.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++) u[k] = V.disu_fpts[idx + k * s];
}
template <int ND, int NF>
__device__ __forceinline__ void load_grad_fpt(const hf_ele_view &V, int idx, double *g)
{
  size_t s = (size_t)V.n_fpts * V.n_eles;
#pragma unroll
  for (int q = 0; q < NF * ND; q++) g[q] = V.grad_disu_fpts[idx + q * s];
}
template <int ND, int NF>
__device__ __forceinline__ void add_sgsf_fpt(const hf_ele_view &V, int idx, double *f)
{
  size_t s = (size_t)V.n_fpts * V.n_eles;
#pragma unroll
  for (int q = 0; q < NF * ND; q++) f[q] += V.sgsf_fpts[idx + q * s];
}
template <int ND>
__device__ __forceinline__ void load_norm(const hf_ele_view &V, int idx, double *n)
{
  size_t s = (size_t)V.n_fpts * V.n_eles;
#pragma unroll
  for (int d = 0; d < ND; d++) n[d] = V.norm_fpts[idx + d * s];
}

template <int ND, int NF>
__global__ void k_int_invflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int *__restrict__ idx_r,
                              const int8_t *__restrict__ type_l, const int8_t *__restrict__ type_r, hf_phys P, int viscous)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  const hf_ele_view &R = W.v[type_r[i]];
  int il = idx_l[t], ir = idx_r[t];
  double u_l[NF], u_r[NF], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
  load_fpt<ND, NF>(R, ir, u_r);
  load_norm<ND>(L, il, n);
  riemann<ND, NF>(u_l, u_r, n, fn, P);
  double tdA_l = L.tdA_fpts[il], tdA_r = R.tdA_fpts[ir];
  size_t sl = (size_t)L.n_fpts * L.n_eles, sr = (size_t)R.n_fpts * R.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++)
  {
    L.norm_tconf_fpts[il + k * sl] = fn[k] * tdA_l;
    R.norm_tconf_fpts[ir + k * sr] = -fn[k] * tdA_r;
  }
  if (viscous)
  {
    double beta = ldg_switched_beta<ND>(P.ldg_beta, n);
    double u_c[NF];
    ldg_solution_int<NF>(u_l, u_r, u_c, beta);
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      L.delta_disu_fpts[il + k * sl] = (u_c[k] - u_l[k]);
      R.delta_disu_fpts[ir + k * sr] = (u_c[k] - u_r[k]);
    }
  }
}

template <int ND, int NF>
__global__ void k_int_viscflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int *__restrict__ idx_r,
                               const int8_t *__restrict__ type_l, const int8_t *__restrict__ type_r, hf_phys P, int les)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  const hf_ele_view &R = W.v[type_r[i]];
  int il = idx_l[t], ir = idx_r[t];
  double u_l[NF], u_r[NF], g[NF * ND], f_l[NF * ND], f_r[NF * ND], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
  load_fpt<ND, NF>(R, ir, u_r);
  load_grad_fpt<ND, NF>(L, il, g);
  vis_flux<ND, NF>(u_l, g, f_l, P);
  load_grad_fpt<ND, NF>(R, ir, g);
  vis_flux<ND, NF>(u_r, g, f_r, P);
  if (les) // physical SGS flux of both sides joins the viscous flux (reference src/int_inters.cpp:297-313)
  {
    add_sgsf_fpt<ND, NF>(L, il, f_l);
    add_sgsf_fpt<ND, NF>(R, ir, f_r);
  }
  load_norm<ND>(L, il, n);
  double beta = ldg_switched_beta<ND>(P.ldg_beta, n);
  ldg_flux<ND, NF>(0, u_l, u_r, f_l, f_r, n, fn, beta, P.ldg_tau);
  double tdA_l = L.tdA_fpts[il], tdA_r = R.tdA_fpts[ir];
  size_t sl = (size_t)L.n_fpts * L.n_eles, sr = (size_t)R.n_fpts * R.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++)
  {
    L.norm_tconf_fpts[il + k * sl] += fn[k] * tdA_l;
    R.norm_tconf_fpts[ir + k * sr] += -fn[k] * tdA_r;
  }
}

template <int ND, int NF>
__global__ void k_bdy_invflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l,
                              const int *__restrict__ bc_id, const hf_bc *__restrict__ bcs, hf_phys P, double R_ref, int viscous)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  int il = idx_l[t];
  const hf_bc B = bcs[bc_id[i]];
  double u_l[NF], u_r[NF], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
  load_norm<ND>(L, il, n);
#pragma unroll
  for (int k = 0; k < NF; k++) u_r[k] = 0.;
  set_boundary_conditions<ND, NF>(0, B, u_l, u_r, n, P.gamma, R_ref);
  if (NF > 1 && B.bc_flag == HF_SLIP_WALL_DUAL)
  {
    double f_l[NF * ND];
    inv_flux<ND, NF>(u_l, f_l, P);
    normal_flux<ND, NF>(f_l, n, fn);
  }
  else
    riemann<ND, NF>(u_l, u_r, n, fn, P);
  double tdA_l = L.tdA_fpts[il];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++) L.norm_tconf_fpts[il + k * sl] = fn[k] * tdA_l;
  if (viscous)
  {
    if (hf_is_wall(B.bc_flag)) set_boundary_conditions<ND, NF>(1, B, u_l, u_r, n, P.gamma, R_ref);
#pragma unroll
    for (int k = 0; k < NF; k++) L.delta_disu_fpts[il + k * sl] = (u_r[k] - u_l[k]);
  }
}

template <int ND, int NF>
__global__ void k_bdy_viscflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l,
                               const int *__restrict__ bc_id, const hf_bc *__restrict__ bcs, hf_phys P, double R_ref, const int *__restrict__ wm_upt,
                               const double *__restrict__ wm_dist, hf_wm Q)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  int il = idx_l[t];
  const hf_bc B = bcs[bc_id[i]];
  if (B.bc_flag == HF_SLIP_WALL) return;
  double u_l[NF], u_r[NF], g_l[NF * ND], g_r[NF * ND], f_r[NF * ND], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
  load_norm<ND>(L, il, n);
  if (NF > 1 && B.use_wm && wm_upt && wm_upt[i] >= 0)
  {
    // wall-modelled interface (reference src/bdy_inters.cpp:1095-1131): no-slip wall state, stress from the wall model
    // evaluated with the solution at the element's input point
#pragma unroll
    for (int k = 0; k < NF; k++) u_r[k] = 0.;
    set_boundary_conditions<ND, NF>(2, B, u_l, u_r, n, P.gamma, R_ref);
    double u_wm[NF];
    const size_t su = (size_t)L.n_upts * L.n_eles;
#pragma unroll
    for (int k = 0; k < NF; k++) u_wm[k] = L.disu_upts[wm_upt[i] + k * su];
    calc_wall_stress<ND, NF>(u_wm, u_r, wm_dist[i], n, fn, Q);
    const double tdA = L.tdA_fpts[il];
    const size_t sl = (size_t)L.n_fpts * L.n_eles;
#pragma unroll
    for (int k = 0; k < NF; k++) L.norm_tconf_fpts[il + k * sl] += fn[k] * tdA;
    return;
  }
  load_grad_fpt<ND, NF>(L, il, g_l);
#pragma unroll
  for (int k = 0; k < NF; k++) u_r[k] = 0.;
  set_boundary_conditions<ND, NF>(1, B, u_l, u_r, n, P.gamma, R_ref);
  set_boundary_gradients<ND, NF>(B.bc_flag, u_r, g_l, g_r, n);
  vis_flux<ND, NF>(u_r, g_r, f_r, P);
  ldg_flux<ND, NF>(1, u_l, u_r, f_r, f_r, n, fn, 0.0, P.ldg_tau);
  double tdA_l = L.tdA_fpts[il];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++) L.norm_tconf_fpts[il + k * sl] += fn[k] * tdA_l;
}

// partition ("mpi") faces: right state from the receive buffer, left side only is written
// in_disu(j_rhs, field, inter), in_grad(j_rhs, field, dim, inter)     (reference src/mpi_inters.cpp:154-215)
template <int ND, int NF>
__global__ void k_mpi_invflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l,
                              const int *__restrict__ lut, const double *__restrict__ in_disu, hf_phys P, int viscous)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  int il = idx_l[t];
  double u_l[NF], u_r[NF], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
#pragma unroll
  for (int k = 0; k < NF; k++) u_r[k] = in_disu[lut[t] + (size_t)nf * (k + NF * (size_t)i)];
  load_norm<ND>(L, il, n);
  riemann<ND, NF>(u_l, u_r, n, fn, P);
  double tdA_l = L.tdA_fpts[il];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++) L.norm_tconf_fpts[il + k * sl] = fn[k] * tdA_l;
  if (viscous)
  {
    double beta = ldg_switched_beta<ND>(P.ldg_beta, n);
    double u_c[NF];
    ldg_solution_int<NF>(u_l, u_r, u_c, beta);
#pragma unroll
    for (int k = 0; k < NF; k++) L.delta_disu_fpts[il + k * sl] = (u_c[k] - u_l[k]);
  }
}

template <int ND, int NF>
__global__ void k_mpi_viscflux(hf_views W, int n_pairs, int nf, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l,
                               const int *__restrict__ lut, const double *__restrict__ in_disu, const double *__restrict__ in_grad,
                               const double *__restrict__ in_sgsf, hf_phys P)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf;
  const hf_ele_view &L = W.v[type_l[i]];
  int il = idx_l[t];
  double u_l[NF], u_r[NF], g[NF * ND], f_l[NF * ND], f_r[NF * ND], n[ND], fn[NF];
  load_fpt<ND, NF>(L, il, u_l);
  load_grad_fpt<ND, NF>(L, il, g);
  vis_flux<ND, NF>(u_l, g, f_l, P);
#pragma unroll
  for (int k = 0; k < NF; k++) u_r[k] = in_disu[lut[t] + (size_t)nf * (k + NF * (size_t)i)];
#pragma unroll
  for (int d = 0; d < ND; d++)
#pragma unroll
    for (int k = 0; k < NF; k++) g[k + NF * d] = in_grad[lut[t] + (size_t)nf * (k + NF * (d + ND * (size_t)i))];
  vis_flux<ND, NF>(u_r, g, f_r, P);
  if (in_sgsf) // LES: physical SGS flux of both sides (reference src/mpi_inters.cpp, calculate_common_viscFlux)
  {
    add_sgsf_fpt<ND, NF>(L, il, f_l);
#pragma unroll
    for (int d = 0; d < ND; d++)
#pragma unroll
      for (int k = 0; k < NF; k++) f_r[k + NF * d] += in_sgsf[lut[t] + (size_t)nf * (k + NF * (d + ND * (size_t)i))];
  }
  load_norm<ND>(L, il, n);
  double beta = ldg_switched_beta<ND>(P.ldg_beta, n);
  ldg_flux<ND, NF>(0, u_l, u_r, f_l, f_r, n, fn, beta, P.ldg_tau);
  double tdA_l = L.tdA_fpts[il];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
#pragma unroll
  for (int k = 0; k < NF; k++) L.norm_tconf_fpts[il + k * sl] += fn[k] * tdA_l;
}

// out_disu[inter][field][fpt], out_grad[inter][dim][field][fpt]  (reference src/mpi_inters.cpp:226-229, 285-289)
__global__ void k_pack_disu(hf_views W, int n_pairs, int nf, int NF, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l, double *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf, j = t - i * nf;
  const hf_ele_view &L = W.v[type_l[i]];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
  for (int k = 0; k < NF; k++) out[j + (size_t)nf * (k + NF * (size_t)i)] = L.disu_fpts[idx_l[t] + k * sl];
}
__global__ void k_pack_grad(hf_views W, int n_pairs, int nf, int NF, int ND, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l, double *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf, j = t - i * nf;
  const hf_ele_view &L = W.v[type_l[i]];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
  for (int d = 0; d < ND; d++)
    for (int k = 0; k < NF; k++) out[j + (size_t)nf * (k + NF * (d + ND * (size_t)i))] = L.grad_disu_fpts[idx_l[t] + (k + NF * d) * sl];
}
// the same message layout for the physical SGS flux (mpi_inters::send_sgsf_fpts)
__global__ void k_pack_sgsf(hf_views W, int n_pairs, int nf, int NF, int ND, const int *__restrict__ idx_l, const int8_t *__restrict__ type_l, double *__restrict__ out)
{
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_pairs) return;
  int i = t / nf, j = t - i * nf;
  const hf_ele_view &L = W.v[type_l[i]];
  size_t sl = (size_t)L.n_fpts * L.n_eles;
  for (int d = 0; d < ND; d++)
    for (int k = 0; k < NF; k++) out[j + (size_t)nf * (k + NF * (d + ND * (size_t)i))] = L.sgsf_fpts[idx_l[t] + (k + NF * d) * sl];
}

// ---------------------------------------------------------------------------------------------------------------------
// kernels: time integration, norms, time step
// ---------------------------------------------------------------------------------------------------------------------
// mode 0: u -= c0 * (div/detjac)                      (forward Euler, RK24 stages 0-2, RK34 stages 0,1,3)
// mode 1: u = c1*u + c2*u1 + c0*(-div/detjac)         (RK24 last stage, RK34 stage 2)
// mode 2: r = a*r + dt*(-div/detjac); u += b*r        (RK45 / RK414), c1 = a, c2 = b
__global__ void k_rk_update(long long n, long long n_pts, int n_upts, double *__restrict__ u0, double *__restrict__ u1, const double *__restrict__ div,
                            const double *__restrict__ detjac, const double *__restrict__ dt_local, double dt, double fac, double c1, double c2,
                            int mode, int copy_u1, int *__restrict__ nan_flag)
{
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  long long p = idx % n_pts;
  double dtl = dt_local ? dt_local[p / n_upts] : dt;
  double u = u0[idx];
  if (copy_u1) u1[idx] = u;
  double r = div[idx] / detjac[p];
  if (r != r) *nan_flag = 1 + (int)(p / n_upts);
  if (mode == 0)
    u -= dtl / fac * (r - 0.0);
  else if (mode == 1)
  {
    double rhs = -r + 0.0;
    double uo = copy_u1 ? u : u1[idx];
    u = c1 * u + c2 * uo + dtl / fac * rhs;
  }
  else
  {
    double rhs = -r + 0.0;
    double d = c1 * u1[idx] + dtl * rhs;
    u1[idx] = d;
    u += c2 * d;
  }
  u0[idx] = u;
}

// per-block partial norms of div/detjac for one field; partial[block]
template <int NORM>
__global__ void k_res_norm(long long n_pts, const double *__restrict__ div, const double *__restrict__ detjac, double *__restrict__ partial)
{
  __shared__ double sh[256];
  double acc = 0.;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < n_pts; p += (long long)gridDim.x * blockDim.x)
  {
    double r = div[p] / detjac[p] - 0.0;
    if (NORM == 0) acc = fmax(acc, fabs(r));
    else if (NORM == 1) acc += fabs(r);
    else acc += r * r;
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1)
  {
    if (threadIdx.x < s) sh[threadIdx.x] = (NORM == 0) ? fmax(sh[threadIdx.x], sh[threadIdx.x + s]) : sh[threadIdx.x] + sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// eles::calc_dt_local (reference src/eles.cpp:1267-1356): one thread per element
template <int ND>
__global__ void k_dt_local(int n_eles, int n_upts, const double *__restrict__ u, const double *__restrict__ h_ref, double *__restrict__ dt_out, hf_phys P,
                           double CFL, int order, int viscous)
{
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_eles) return;
  size_t s = (size_t)n_upts * n_eles;
  double lam_inv = 0, lam_visc = 0;
  for (int i = 0; i < n_upts; i++)
  {
    size_t p = i + (size_t)n_upts * e;
    double rho = u[p];
    double vsq = 0.;
#pragma unroll
    for (int d = 0; d < ND; d++)
    {
      double v = u[p + (d + 1) * s] / rho;
      vsq += v * v;
    }
    double pr = (P.gamma - 1.0) * (u[p + (ND + 1) * s] - 0.5 * rho * vsq);
    double c = sqrt(P.gamma * pr / rho);
    double inte = pr / ((P.gamma - 1.0) * rho);
    double rt_ratio = (P.gamma - 1.0) * inte / (P.rt_inf);
    double mu = (P.mu_inf) * pow(rt_ratio, 1.5) * (1. + (P.c_sth)) / (rt_ratio + (P.c_sth));
    mu = mu + P.fix_vis * (P.mu_inf - mu);
    double li = sqrt(vsq) + c;
    double lv = fmax(4.0 / 3.0, P.gamma / P.prandtl) * mu / rho;
    if (lam_inv < li) lam_inv = li;
    if (lam_visc < lv) lam_visc = lv;
  }
  double h = h_ref[e];
  double dt_inv = CFL * h / lam_inv * 1.0 / (2.0 * order + 1.0);
  double dt_visc = viscous ? (CFL * 0.25 * h * h) / (lam_visc) * 1.0 / (2.0 * order + 1.0) : 1e16;
  dt_out[e] = fmin(dt_visc, dt_inv);
}

__global__ void k_min_reduce(int n, const double *__restrict__ x, double *__restrict__ partial)
{
  __shared__ double sh[256];
  double acc = 1e300;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) acc = fmin(acc, x[p]);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = blockDim.x / 2; s > 0; s >>= 1)
  {
    if (threadIdx.x < s) sh[threadIdx.x] = fmin(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
constexpr int TL_SLOTS = 16, TL_MARKS = 12;
void hf_tl_mark(hf_ctx *c, int what, bool comm)
{
  if (!c->tl_on) return;
  if (c->tl_ev.empty())
  {
    c->tl_ev.resize(TL_SLOTS * TL_MARKS);
    for (auto &e : c->tl_ev) cudaEventCreate(&e);
  }
  cudaEventRecord(c->tl_ev[(c->tl_stage % TL_SLOTS) * TL_MARKS + what], comm ? c->comm_stream : c->stream);
}
void hf_ktimer_begin(hf_ctx *c)
{
  if (!c->ktimer_on) return;
  if (c->kt_used + 2 > c->kt_ev.size())
  {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    c->kt_ev.push_back(a);
    c->kt_ev.push_back(b);
  }
  cudaEventRecord(c->kt_ev[c->kt_used], c->stream);
}
void hf_ktimer_end(hf_ctx *c)
{
  if (!c->ktimer_on) return;
  cudaEventRecord(c->kt_ev[c->kt_used + 1], c->stream);
  c->kt_used += 2;
}

hf_views hf_make_views(hf_ctx *c)
{
  hf_views W;
  memset(&W, 0, sizeof(W));
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    hf_eles_dev &e = c->eles[t];
    if (!e.present) continue;
    hf_ele_view &v = W.v[t];
    v.n_eles = e.n_eles; v.n_upts = e.n_upts; v.n_fpts = e.n_fpts; v.n_dims = e.n_dims; v.n_fields = e.n_fields;
    v.disu_fpts = e.disu_fpts;
    v.norm_tconf_fpts = e.norm_tconf_fpts;
    v.delta_disu_fpts = e.delta_disu_fpts;
    v.grad_disu_fpts = e.grad_disu_fpts;
    v.sgsf_fpts = e.sgsf_fpts;
    v.disu_upts = e.disu_upts[0];
    v.tdA_fpts = e.tdA_fpts;
    v.norm_fpts = e.norm_fpts;
  }
  return W;
}

static int build_ell(hf_ctx *c, hf_ell &E, const double *dense, int rows, int cols)
{
  int nnz = 1;
  for (int r = 0; r < rows; r++)
  {
    int cnt = 0;
    for (int k = 0; k < cols; k++) if (dense[r + (size_t)rows * k] != 0.0) cnt++;
    nnz = std::max(nnz, cnt);
  }
  std::vector<double> val((size_t)nnz * rows, 0.0);
  std::vector<int> col((size_t)nnz * rows, 0);
  for (int r = 0; r < rows; r++)
  {
    int cnt = 0, last = 0;
    for (int k = 0; k < cols; k++)
      if (dense[r + (size_t)rows * k] != 0.0)
      {
        val[(size_t)cnt * rows + r] = dense[r + (size_t)rows * k];
        col[(size_t)cnt * rows + r] = k;
        last = k;
        cnt++;
      }
    for (; cnt < nnz; cnt++) col[(size_t)cnt * rows + r] = last;
  }
  E.rows = rows; E.cols = cols; E.nnz = nnz;
  if (hf_alloc_copy(c, &E.val, val.data(), val.size())) return 1;
  if (hf_alloc_copy(c, &E.col, col.data(), col.size())) return 1;
  size_t filled = 0;
  for (size_t i = 0; i < (size_t)rows * cols; i++) filled += dense[i] != 0.0;
  if (2 * filled >= (size_t)rows * cols && hf_alloc_copy(c, &E.dval, dense, (size_t)rows * cols)) return 1;
  return 0;
}

// ---- integral diagnostics (eles::CalcIntegralQuantities, reference src/eles.cpp:5485-5628) ------------------------------------
// One CTA per element: thread j interpolates the solution and its physical gradient to volume cubature point j with
// opp_volume_cubpts (sum over solution points in ascending order, as the reference), evaluates the requested
// quantities there and weighs them with weight * detjac; thread q then adds the element's terms in cubature-point order.
// The host adds the per-element sums in element order, so the only difference to the reference's single running sum is
// the grouping by element.
struct hf_iq_kinds { int n; int kind[HF_MAX_INTEGRAL_QUANTITIES]; };
template <int ND>
__global__ void k_integral_quantities(int n_eles, int n_upts, int n_cub, const double *__restrict__ u, const double *__restrict__ grad,
                                      const double *__restrict__ opp, const double *__restrict__ w, const double *__restrict__ detjac,
                                      const int *__restrict__ pos, hf_iq_kinds K, double gamma, double *__restrict__ elem_out)
{
  constexpr int NF = ND + 2;
  extern __shared__ double terms[]; // [q][n_cub]
  const int i = blockIdx.x;
  const size_t s = pos ? pos[i] : i;
  const size_t fs = (size_t)n_upts * n_eles;
  for (int j = threadIdx.x; j < n_cub; j += blockDim.x)
  {
    double uc[NF], gc[NF][ND];
    for (int m = 0; m < NF; m++)
    {
      double a = 0.;
      for (int k = 0; k < n_upts; k++) a += opp[j + (size_t)n_cub * k] * u[k + n_upts * s + fs * m];
      uc[m] = a;
    }
    for (int m = 0; m < NF; m++)
      for (int n = 0; n < ND; n++)
      {
        double a = 0.;
        for (int k = 0; k < n_upts; k++) a += opp[j + (size_t)n_cub * k] * grad[k + n_upts * s + fs * (m + NF * n)];
        gc[m][n] = a;
      }
    const double irho = 1. / uc[0];
    double dv[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // dv[a][b] = d v_a / d x_b
    for (int a = 0; a < ND; a++)
      for (int b = 0; b < ND; b++) dv[a][b] = irho * (gc[1 + a][b] - uc[1 + a] * irho * gc[0][b]);
    const double wd = w[j] * detjac[j + (size_t)n_cub * i];
    for (int q = 0; q < K.n; q++)
    {
      double diagnostic = 0.0;
      const int kind = K.kind[q];
      if (kind == 0) // kineticenergy
      {
        double tke = 0.0;
        for (int n = 1; n < ND + 1; n++) tke += 0.5 * uc[n] * uc[n];
        diagnostic = irho * tke;
      }
      else if (kind == 1) // enstropy
      {
        const double wz = dv[1][0] - dv[0][1];
        diagnostic = wz * wz;
        if (ND == 3)
        {
          const double wx = dv[2][1] - dv[1][2], wy = dv[0][2] - dv[2][0];
          diagnostic += wx * wx + wy * wy;
        }
        diagnostic *= 0.5 / irho;
      }
      else if (kind == 2) // pressuredilatation
      {
        double tke = 0.0;
        for (int n = 1; n < ND + 1; n++) tke += 0.5 * uc[n] * uc[n];
        const double pressure = (gamma - 1.0) * (uc[ND + 1] - irho * tke);
        diagnostic = (ND == 2) ? pressure * (dv[0][0] + dv[1][1]) : pressure * (dv[0][0] + dv[1][1] + dv[2][2]);
      }
      else // straincolonproduct (3), devstraincolonproduct (4)
      {
        double S[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        S[0][0] = dv[0][0];
        S[0][1] = (dv[0][1] + dv[1][0]) / 2.0;
        S[1][0] = S[0][1];
        S[1][1] = dv[1][1];
        double diag = (S[0][0] + S[1][1]) / 3.0;
        if (ND == 3)
        {
          S[0][2] = (dv[0][2] + dv[2][0]) / 2.0;
          S[1][2] = (dv[1][2] + dv[2][1]) / 2.0;
          S[2][0] = S[0][2];
          S[2][1] = S[1][2];
          S[2][2] = dv[2][2];
          diag += S[2][2] / 3.0;
        }
        if (kind == 4)
          for (int a = 0; a < ND; a++) S[a][a] -= diag;
        for (int a = 0; a < ND; a++)
          for (int b = 0; b < ND; b++) diagnostic += S[a][b] * S[a][b];
      }
      terms[q * n_cub + j] = diagnostic * wd;
    }
  }
  __syncthreads();
  if (threadIdx.x < K.n)
  {
    double sum = 0.;
    for (int j = 0; j < n_cub; j++) sum += terms[threadIdx.x * n_cub + j];
    elem_out[(size_t)i * K.n + threadIdx.x] = sum;
  }
}

// running time averages (eles::CalcTimeAverageQuantities, reference src/eles.cpp:5630-5702): one thread per solution point
struct hf_avg_kinds { int n; int kind[HF_MAX_INTEGRAL_QUANTITIES]; };
__global__ void k_time_average(long long n_pts, int n_upts, int n_dims, const double *__restrict__ u, double *__restrict__ avg, const double *__restrict__ dt_local,
                               double dt_global, double time, double spinup_time, hf_avg_kinds K)
{
  long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= n_pts) return;
  const double rho = u[t];
  const double dt = dt_local ? dt_local[t / n_upts] : dt_global;
  double a, b;
  if (time == spinup_time) { a = 0.0; b = 1.0; }
  else { a = (time - spinup_time - dt) / (time - spinup_time); b = dt / (time - spinup_time); }
  for (int i = 0; i < K.n; i++)
  {
    double current_value;
    const int kind = K.kind[i];
    if (kind == 0) current_value = rho;
    else if (kind == 4) current_value = u[t + n_pts * (n_dims + 1)] / rho;
    else current_value = u[t + n_pts * kind] / rho; // u, v, w: fields 1, 2, 3 (in 2-D the reference's w_average reads field 3 as well)
    avg[t + n_pts * i] = a * avg[t + n_pts * i] + b * current_value;
  }
}

extern "C" {

const char *hf_dev_last_error(void) { return g_err.c_str(); }

int hf_dev_create(hf_ctx **out, int device, int rank, int nproc)
{
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    HF_FAIL(std::string("no CUDA device available: the HiFiLES B200 hot path has no CPU fallback (") + cudaGetErrorString(e) + ")");
  if (device < 0)
  {
    const char *lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) % n_dev : rank % n_dev;
  }
  HF_CUDA(cudaSetDevice(device));
  hf_ctx *c = new hf_ctx();
  c->device = device; c->rank = rank; c->nproc = nproc;
  HF_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  {
    // the halo exchange runs beside a compute kernel that fills every SM with queued CTAs: give its stream the highest
    // priority so that the NCCL kernel's CTAs are placed as soon as a slot frees up instead of behind the whole grid
    int prio_lo = 0, prio_hi = 0;
    HF_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    HF_CUDA(cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi));
  }
  c->own_stream = true;
  HF_CUDA(cudaEventCreateWithFlags(&c->ev_a, cudaEventDisableTiming));
  HF_CUDA(cudaEventCreateWithFlags(&c->ev_b, cudaEventDisableTiming));
  HF_CUDA(cudaEventCreate(&c->ev_t0));
  HF_CUDA(cudaEventCreate(&c->ev_t1));
  c->scratch_bytes = 1 << 20;
  if (hf_alloc_zero(c, &c->scratch, c->scratch_bytes / sizeof(double))) return 1;
  if (hf_alloc_zero(c, &c->d_nan, 1)) return 1;
  HF_CUDA(cudaMallocHost((void **)&c->h_nan, sizeof(int)));
  *c->h_nan = 0;
  memset(&c->prm, 0, sizeof(c->prm));
  *out = c;
  return 0;
}

int hf_dev_destroy(hf_ctx *c)
{
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  hf_halo_destroy(c);
  hf_fused_destroy(c);
  hf_elem_destroy(c);
  for (void *p : c->allocs) cudaFree(p);
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    if (c->eles[t].d_stage) cudaFree(c->eles[t].d_stage);
    if (c->eles[t].d_xfer) cudaFree(c->eles[t].d_xfer);
  }
  for (cudaEvent_t ev : c->kt_ev) cudaEventDestroy(ev);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  if (c->xfer_stream) cudaStreamDestroy(c->xfer_stream);
  if (c->ev_xfer) cudaEventDestroy(c->ev_xfer);
  if (c->ev_xfer_free) cudaEventDestroy(c->ev_xfer_free);
  if (c->h_nan) cudaFreeHost(c->h_nan);
  delete c;
  return 0;
}

int hf_dev_set_stream(hf_ctx *c, void *cuda_stream)
{
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)cuda_stream;
  c->own_stream = false;
  return 0;
}

int hf_dev_set_params(hf_ctx *c, const hf_params *p)
{
  c->prm = *p;
  hf_phys &P = c->phys;
  P.gamma = p->gamma; P.prandtl = p->prandtl; P.mu_inf = p->mu_inf; P.rt_inf = p->rt_inf; P.c_sth = p->c_sth;
  P.fix_vis = (double)p->fix_vis;
  P.ldg_beta = p->ldg_beta; P.ldg_tau = p->ldg_tau;
  for (int i = 0; i < 3; i++) P.wave_speed[i] = p->wave_speed[i];
  P.diff_coeff = p->diff_coeff; P.lambda = p->lambda;
  P.riemann_solve_type = p->riemann_solve_type;
  P.gamma_over_pr = p->gamma / p->prandtl;
  if (p->shock_cap && p->shock_cap != 1) HF_FAIL("Shock capturing method not implemented.");
  if (p->LES && (p->SGS_model < 0 || p->SGS_model > 4)) HF_FAIL("SGS model not implemented");
  if (p->equation == 0 && !(p->riemann_solve_type == 0 || p->riemann_solve_type == 2 || p->riemann_solve_type == 3))
    HF_FAIL("Riemann solver not implemented");
  if (p->viscous && p->vis_riemann_solve_type != 0) HF_FAIL("Viscous Riemann solver not implemented");
  c->have_params = true;
  return 0;
}

// Device element order (hf_dev_set_element_order): arrays are (points, element, rest); slot pos[e] holds host element e.
static std::vector<double> permute_eles(const double *src, size_t pre, int n_eles, size_t post, const std::vector<int> &pos, bool to_device)
{
  std::vector<double> out(pre * n_eles * post);
  for (size_t b = 0; b < post; b++)
    for (int e = 0; e < n_eles; e++)
    {
      const size_t host_off = pre * (e + (size_t)n_eles * b), dev_off = pre * (pos[e] + (size_t)n_eles * b);
      if (to_device) memcpy(&out[dev_off], src + host_off, pre * sizeof(double));
      else memcpy(&out[host_off], src + dev_off, pre * sizeof(double));
    }
  return out;
}

// the same permutation on the device, for uploads / downloads of whole arrays (one thread per double)
__global__ void k_permute_eles(const double *__restrict__ src, double *__restrict__ dst, long long pre, int n_eles, long long n, const int *__restrict__ pos,
                               int to_device)
{
  long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const long long a = idx % pre, r = idx / pre;
  const int e = (int)(r % n_eles);
  const long long b = r / n_eles;
  const long long other = a + pre * (pos[e] + (long long)n_eles * b); // idx is in host order, other in device order
  if (to_device) dst[other] = src[idx];
  else dst[idx] = src[other];
}
static int permute_on_device(hf_ctx *c, hf_eles_dev &e, const double *src, double *dst, size_t pre, size_t n, bool to_device)
{
  if (!e.d_pos && hf_alloc_copy(c, &e.d_pos, e.pos.data(), e.pos.size())) return 1;
  k_permute_eles<<<hf_blocks((long long)n, 256), 256, 0, c->stream>>>(src, dst, (long long)pre, e.n_eles, (long long)n, e.d_pos, to_device ? 1 : 0);
  HF_LAUNCH_CHECK(c);
  return 0;
}
static int ensure_stage(hf_ctx *c, hf_eles_dev &e, size_t n)
{
  if (e.stage_n >= n) return 0;
  if (e.d_stage) cudaFree(e.d_stage);
  e.d_stage = nullptr;
  cudaError_t err = cudaMalloc((void **)&e.d_stage, n * sizeof(double));
  if (err != cudaSuccess) { hf_set_error(std::string("cudaMalloc (transfer staging): ") + cudaGetErrorString(err)); return 1; }
  e.stage_n = n;
  return 0;
}

int hf_dev_set_element_order(hf_ctx *c, int ele_type, int n_eles, const int *pos)
{
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES) HF_FAIL("bad element type");
  hf_eles_dev &e = c->eles[ele_type];
  if (e.present) HF_FAIL("hf_dev_set_element_order must precede hf_dev_upload_eles");
  std::vector<char> seen(n_eles, 0);
  for (int i = 0; i < n_eles; i++)
  {
    if (pos[i] < 0 || pos[i] >= n_eles || seen[pos[i]]) HF_FAIL("element order is not a permutation");
    seen[pos[i]] = 1;
  }
  e.pos.assign(pos, pos + n_eles);
  return 0;
}

static int upload_eles_impl(hf_ctx *c, const hf_eles_desc *d);

int hf_dev_upload_eles(hf_ctx *c, const hf_eles_desc *d)
{
  if (d->ele_type < 0 || d->ele_type >= HF_N_ELE_TYPES) HF_FAIL("bad element type");
  hf_eles_dev &e = c->eles[d->ele_type];
  if (e.pos.empty()) return upload_eles_impl(c, d);
  if ((int)e.pos.size() != d->n_eles) HF_FAIL("element order has the wrong length");
  // permuted host copies of every per-element array, then the ordinary upload
  hf_eles_desc q = *d;
  const int ne = d->n_eles, nu = d->n_upts_per_ele, nf = d->n_fpts_per_ele, nd = d->n_dims;
  std::vector<std::vector<double>> keep;
  auto perm = [&](const double *&ptr, size_t pre, size_t post) {
    if (!ptr) return;
    keep.push_back(permute_eles(ptr, pre, ne, post, e.pos, true));
    ptr = keep.back().data();
  };
  perm(q.detjac_upts, nu, 1);
  perm(q.JGinv_upts, (size_t)nd * nd * nu, 1);
  perm(q.detjac_fpts, nf, 1);
  perm(q.JGinv_fpts, (size_t)nd * nd * nf, 1);
  perm(q.tdA_fpts, nf, 1);
  perm(q.norm_fpts, nf, nd);
  perm(q.h_ref, 1, 1);
  perm(q.disu_upts0, nu, d->n_fields);
  perm(q.JGinv_over_int_cubpts, (size_t)nd * nd * d->n_over_int_cubpts, 1);
  perm(q.wall_distance, nu, nd);
  perm(q.Jacobian_fpts, (size_t)nd * nd * nf, 1);
  return upload_eles_impl(c, &q);
}

static int upload_eles_impl(hf_ctx *c, const hf_eles_desc *d)
{
  if (!c->have_params) HF_FAIL("hf_dev_set_params must be called before hf_dev_upload_eles");
  if (d->ele_type < 0 || d->ele_type >= HF_N_ELE_TYPES) HF_FAIL("bad element type");
  HF_CUDA(cudaSetDevice(c->device));
  hf_eles_dev &e = c->eles[d->ele_type];
  if (e.present) HF_FAIL("element type uploaded twice");
  e.present = true;
  e.ele_type = d->ele_type; e.n_eles = d->n_eles; e.n_upts = d->n_upts_per_ele; e.n_fpts = d->n_fpts_per_ele;
  e.n_dims = d->n_dims; e.n_fields = d->n_fields; e.order = d->order; e.n_inters = d->n_inters_per_ele;
  if (e.n_inters > 6) HF_FAIL("too many faces per element");
  e.fpt_offset[0] = 0;
  for (int f = 0; f < e.n_inters; f++)
  {
    e.n_fpts_per_inter[f] = d->n_fpts_per_inter[f];
    e.fpt_offset[f + 1] = e.fpt_offset[f] + d->n_fpts_per_inter[f];
  }
  const int nu = e.n_upts, nf = e.n_fpts, nd = e.n_dims;
  const size_t NU = (size_t)nu * e.n_eles, NFP = (size_t)nf * e.n_eles, F = e.n_fields;
  const bool visc = c->prm.viscous != 0;
  auto keep = [&](int slot, const double *src, size_t n) { if (src) e.h_op[slot].assign(src, src + n); };
  if (build_ell(c, e.opp_0, d->opp_0, nf, nu)) return 1;
  keep(0, d->opp_0, (size_t)nf * nu);
  if (build_ell(c, e.opp_3, d->opp_3, nu, nf)) return 1;
  keep(3, d->opp_3, (size_t)nu * nf);
  for (int i = 0; i < nd; i++)
  {
    if (build_ell(c, e.opp_1[i], d->opp_1[i], nf, nu)) return 1;
    if (build_ell(c, e.opp_2[i], d->opp_2[i], nu, nu)) return 1;
    keep(4 + i, d->opp_2[i], (size_t)nu * nu);
    keep(7 + i, d->opp_1[i], (size_t)nf * nu);
    if (visc)
    {
      if (!d->opp_4[i] || !d->opp_5[i] || !d->opp_6) HF_FAIL("viscous run needs opp_4, opp_5, opp_6");
      if (build_ell(c, e.opp_4[i], d->opp_4[i], nu, nu)) return 1;
      if (build_ell(c, e.opp_5[i], d->opp_5[i], nu, nf)) return 1;
      keep(10 + i, d->opp_5[i], (size_t)nu * nf);
    }
  }
  if (visc && build_ell(c, e.opp_6, d->opp_6, nf, nu)) return 1;
  if (hf_alloc_copy(c, &e.detjac_upts, d->detjac_upts, NU)) return 1;
  if (hf_alloc_copy(c, &e.JGinv_upts, d->JGinv_upts, NU * nd * nd)) return 1;
  if (hf_alloc_copy(c, &e.detjac_fpts, d->detjac_fpts, NFP)) return 1;
  if (hf_alloc_copy(c, &e.JGinv_fpts, d->JGinv_fpts, NFP * nd * nd)) return 1;
  if (hf_alloc_copy(c, &e.tdA_fpts, d->tdA_fpts, NFP)) return 1;
  if (hf_alloc_copy(c, &e.norm_fpts, d->norm_fpts, NFP * nd)) return 1;
  if (d->h_ref && hf_alloc_copy(c, &e.h_ref, d->h_ref, (size_t)e.n_eles)) return 1;
  if (c->prm.dt_type != 0 && hf_alloc_zero(c, &e.dt_local, (size_t)e.n_eles)) return 1;
  if (d->disu_upts0) { if (hf_alloc_copy(c, &e.disu_upts[0], d->disu_upts0, NU * F)) return 1; }
  else if (hf_alloc_zero(c, &e.disu_upts[0], NU * F)) return 1;
  if (c->prm.adv_type != 0 && hf_alloc_zero(c, &e.disu_upts[1], NU * F)) return 1;
  if (hf_alloc_zero(c, &e.div_tconf_upts, NU * F)) return 1;
  if (hf_alloc_zero(c, &e.disu_fpts, NFP * F)) return 1;
  if (hf_alloc_zero(c, &e.norm_tconf_fpts, NFP * F)) return 1;
  if (visc)
  {
    if (hf_alloc_zero(c, &e.delta_disu_fpts, NFP * F)) return 1;
    if (hf_alloc_zero(c, &e.grad_disu_fpts, NFP * F * nd)) return 1;
  }
  if (c->prm.over_int)
  {
    if (!d->n_over_int_cubpts || !d->opp_over_int_cubpts || !d->over_int_filter || !d->JGinv_over_int_cubpts) HF_FAIL("over_int needs the over-integration operators");
    e.n_cub = d->n_over_int_cubpts;
    const size_t NC = (size_t)e.n_cub * e.n_eles;
    if (build_ell(c, e.opp_over_int, d->opp_over_int_cubpts, e.n_cub, nu)) return 1;
    if (build_ell(c, e.over_int_filter, d->over_int_filter, nu, e.n_cub)) return 1;
    if (hf_alloc_copy(c, &e.JGinv_over_int, d->JGinv_over_int_cubpts, NC * nd * nd)) return 1;
    if (hf_alloc_zero(c, &e.u_cub, NC * F)) return 1;
    if (hf_alloc_zero(c, &e.tdisf_cub, NC * F * nd)) return 1;
  }
  if (c->prm.LES)
  {
    if (!visc) HF_FAIL("LES not supported with inviscid flow");
    if (!d->Jacobian_fpts || d->ele_vol_factor <= 0.) HF_FAIL("LES needs Jacobian_fpts and the reference-element volume");
    if (c->prm.SGS_model == 0 && !d->wall_distance) HF_FAIL("the Smagorinsky model needs wall_distance");
    if (hf_alloc_zero(c, &e.sgsf_upts, NU * F * nd)) return 1;
    if (hf_alloc_zero(c, &e.sgsf_fpts, NFP * F * nd)) return 1;
    if (hf_alloc_copy(c, &e.Jacobian_fpts, d->Jacobian_fpts, NFP * nd * nd)) return 1;
    if (d->wall_distance && hf_alloc_copy(c, &e.wall_distance, d->wall_distance, NU * nd)) return 1;
    e.ele_vol_factor = d->ele_vol_factor;
    if (c->prm.SGS_model >= 2)
    {
      if (!d->filter_upts) HF_FAIL("the filter-based SGS models need filter_upts");
      if (build_ell(c, e.filter_upts, d->filter_upts, nu, nu)) return 1;
      if (hf_alloc_zero(c, &e.disuf_upts, NU * F)) return 1;
      if (c->prm.SGS_model != 3)
      {
        const size_t dim3 = nd == 2 ? 3 : 6;
        if (hf_alloc_zero(c, &e.uu, NU * dim3) || hf_alloc_zero(c, &e.Lu, NU * dim3)) return 1;
        if (hf_alloc_zero(c, &e.ue, NU * nd) || hf_alloc_zero(c, &e.Le, NU * nd)) return 1;
      }
    }
  }
  if (c->prm.shock_cap)
  {
    if (!d->inv_vandermonde || !d->exp_filter || !d->sensor_w_top || !d->sensor_w_all) HF_FAIL("shock_cap needs the modal transform, sensor weights and filter");
    if (hf_alloc_copy(c, &e.inv_vandermonde, d->inv_vandermonde, (size_t)nu * nu)) return 1;
    if (hf_alloc_copy(c, &e.exp_filter, d->exp_filter, (size_t)nu * nu)) return 1;
    if (hf_alloc_copy(c, &e.sensor_w_top, d->sensor_w_top, (size_t)nu)) return 1;
    if (hf_alloc_copy(c, &e.sensor_w_all, d->sensor_w_all, (size_t)nu)) return 1;
    if (hf_alloc_zero(c, &e.sensor, (size_t)e.n_eles)) return 1;
  }
  // tdisf_upts, norm_tdisf_fpts, grad_disu_upts are only needed by the staged path: allocated lazily
  if (c->fused && hf_fused_on_upload(c, e, d)) return 1;
  if (c->fused && hf_elem_on_upload(c, e, d)) return 1;
  return 0;
}

static int ensure_staged_buffers(hf_ctx *c, hf_eles_dev &e);
extern "C++" int hf_ensure_staged_buffers(hf_ctx *c, hf_eles_dev &e) { return ensure_staged_buffers(c, e); }
static int ensure_staged_buffers(hf_ctx *c, hf_eles_dev &e)
{
  const size_t NU = (size_t)e.n_upts * e.n_eles, NFP = (size_t)e.n_fpts * e.n_eles, F = e.n_fields;
  if (!e.tdisf_upts && hf_alloc_zero(c, &e.tdisf_upts, NU * F * e.n_dims)) return 1;
  if (!e.norm_tdisf_fpts && hf_alloc_zero(c, &e.norm_tdisf_fpts, NFP * F)) return 1;
  if (c->prm.viscous && !e.grad_disu_upts && hf_alloc_zero(c, &e.grad_disu_upts, NU * F * e.n_dims)) return 1;
  return 0;
}

// flat flux-point index (fpt + n_fpts*ele) of face-local point j on local face `loc` of element `ele`
static inline int dev_ele(const hf_eles_dev &e, int ele) { return e.pos.empty() ? ele : e.pos[ele]; }
static inline int flat_fpt(const hf_eles_dev &e, int ele, int loc, int j) { return e.fpt_offset[loc] + j + e.n_fpts * dev_ele(e, ele); }

// right-side permutation (reference src/inters.cpp:153-262)
static int fill_lut(int inter_type, int order, int nfp, int rot, std::vector<int> &lut)
{
  int n = order + 1;
  lut.assign(nfp, 0);
  if (inter_type == 0)
    for (int i = 0; i < nfp; i++) lut[i] = nfp - i - 1;
  else if (inter_type == 1)
  {
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n - j; i++)
      {
        int i0 = j * n - (j - 1) * j / 2 + i, i1;
        if (rot == 0) i1 = i * n - (i - 1) * i / 2 + j;
        else if (rot == 1) i1 = n * (order + 2) / 2 - 1 - (i + j) * (i + j + 1) / 2 - j;
        else if (rot == 2) i1 = j * n - (j - 1) * j / 2 + (order - j - i);
        else HF_FAIL("ERROR: Unknown rotation of triangular face...");
        lut[i0] = i1;
      }
  }
  else
  {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
      {
        int v;
        if (rot == 0) v = (n - 1 - j) + n * i;
        else if (rot == 1) v = nfp - (n - 1 - j) - n * i - 1;
        else if (rot == 2) v = n * j + i;
        else if (rot == 3) v = nfp - n * j - i - 1;
        else HF_FAIL("ERROR: Unknown rotation tag ... ");
        lut[i * n + j] = v;
      }
  }
  return 0;
}

int hf_dev_upload_int_inters(hf_ctx *c, const hf_int_inters_desc *d)
{
  HF_CUDA(cudaSetDevice(c->device));
  hf_int_inters_dev &I = c->ints[d->inter_type];
  I.n_inters = d->n_inters; I.nf = d->n_fpts_per_inter;
  const int nf = I.nf, ni = I.n_inters;
  std::vector<int> il((size_t)nf * ni), ir((size_t)nf * ni), lut;
  std::vector<int8_t> tl(ni), tr(ni);
  for (int i = 0; i < ni; i++)
  {
    const hf_eles_dev &L = c->eles[d->ele_type_l[i]], &R = c->eles[d->ele_type_r[i]];
    if (!L.present || !R.present) HF_FAIL("interface refers to an element type that was not uploaded");
    if (fill_lut(d->inter_type, c->prm.order, nf, d->rot_tag[i], lut)) return 1;
    tl[i] = (int8_t)d->ele_type_l[i]; tr[i] = (int8_t)d->ele_type_r[i];
    for (int j = 0; j < nf; j++)
    {
      il[j + (size_t)nf * i] = flat_fpt(L, d->ele_l[i], d->local_inter_l[i], j);
      ir[j + (size_t)nf * i] = flat_fpt(R, d->ele_r[i], d->local_inter_r[i], lut[j]);
    }
  }
  if (hf_alloc_copy(c, &I.idx_l, il.data(), il.size())) return 1;
  if (hf_alloc_copy(c, &I.idx_r, ir.data(), ir.size())) return 1;
  if (hf_alloc_copy(c, &I.type_l, tl.data(), tl.size())) return 1;
  if (hf_alloc_copy(c, &I.type_r, tr.data(), tr.size())) return 1;
  I.h_ele_type_l.assign(d->ele_type_l, d->ele_type_l + ni); I.h_ele_l.assign(d->ele_l, d->ele_l + ni);
  for (int i = 0; i < ni; i++) I.h_ele_l[i] = dev_ele(c->eles[I.h_ele_type_l[i]], I.h_ele_l[i]);
  I.h_loc_l.assign(d->local_inter_l, d->local_inter_l + ni);
  I.h_ele_type_r.assign(d->ele_type_r, d->ele_type_r + ni); I.h_ele_r.assign(d->ele_r, d->ele_r + ni);
  for (int i = 0; i < ni; i++) I.h_ele_r[i] = dev_ele(c->eles[I.h_ele_type_r[i]], I.h_ele_r[i]);
  I.h_loc_r.assign(d->local_inter_r, d->local_inter_r + ni);
  I.h_rot.assign(d->rot_tag, d->rot_tag + ni);
  return 0;
}

int hf_dev_set_bc_table(hf_ctx *c, int n_bc, const hf_bc *table)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (n_bc > 0 && c->bc_table && n_bc == c->n_bc)
  {
    // a changed table of the same size (ramped inlet): in place, behind the kernels already queued on the stream
    c->h_bc.assign(table, table + n_bc);
    HF_CUDA(cudaMemcpyAsync(c->bc_table, c->h_bc.data(), (size_t)n_bc * sizeof(hf_bc), cudaMemcpyHostToDevice, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
  }
  c->n_bc = n_bc;
  if (n_bc > 0 && hf_alloc_copy(c, &c->bc_table, table, (size_t)n_bc)) return 1;
  c->h_bc.assign(table, table + n_bc);
  return 0;
}

int hf_dev_upload_bdy_inters(hf_ctx *c, const hf_bdy_inters_desc *d)
{
  HF_CUDA(cudaSetDevice(c->device));
  hf_bdy_inters_dev &I = c->bdys[d->inter_type];
  I.n_inters = d->n_inters; I.nf = d->n_fpts_per_inter;
  const int nf = I.nf, ni = I.n_inters;
  std::vector<int> il((size_t)nf * ni);
  std::vector<int8_t> tl(ni);
  for (int i = 0; i < ni; i++)
  {
    const hf_eles_dev &L = c->eles[d->ele_type_l[i]];
    if (!L.present) HF_FAIL("interface refers to an element type that was not uploaded");
    if (d->bc_id[i] < 0 || d->bc_id[i] >= c->n_bc) HF_FAIL("boundary interface refers to an unknown bc id");
    tl[i] = (int8_t)d->ele_type_l[i];
    for (int j = 0; j < nf; j++) il[j + (size_t)nf * i] = flat_fpt(L, d->ele_l[i], d->local_inter_l[i], j);
  }
  if (hf_alloc_copy(c, &I.idx_l, il.data(), il.size())) return 1;
  if (hf_alloc_copy(c, &I.type_l, tl.data(), tl.size())) return 1;
  if (hf_alloc_copy(c, &I.bc_id, d->bc_id, (size_t)ni)) return 1;
  I.h_ele_type_l.assign(d->ele_type_l, d->ele_type_l + ni); I.h_ele_l.assign(d->ele_l, d->ele_l + ni);
  for (int i = 0; i < ni; i++) I.h_ele_l[i] = dev_ele(c->eles[I.h_ele_type_l[i]], I.h_ele_l[i]);
  I.h_loc_l.assign(d->local_inter_l, d->local_inter_l + ni);
  I.h_bc_id.assign(d->bc_id, d->bc_id + ni);
  if (d->wm_upt && d->wm_dist)
  {
    // input points of the wall model, translated to the device element order
    std::vector<int> w(ni, -1);
    for (int i = 0; i < ni; i++)
      if (d->wm_upt[i] >= 0)
      {
        const hf_eles_dev &L = c->eles[d->ele_type_l[i]];
        const int ele = d->wm_upt[i] / L.n_upts, upt = d->wm_upt[i] - ele * L.n_upts;
        w[i] = upt + L.n_upts * dev_ele(L, ele);
      }
    if (hf_alloc_copy(c, &I.wm_upt, w.data(), w.size())) return 1;
    if (hf_alloc_copy(c, &I.wm_dist, d->wm_dist, (size_t)ni)) return 1;
  }
  return 0;
}

int hf_dev_upload_mpi_inters(hf_ctx *c, const hf_mpi_inters_desc *d)
{
  HF_CUDA(cudaSetDevice(c->device));
  hf_mpi_inters_dev &I = c->mpis[d->inter_type];
  I.n_inters = d->n_inters; I.nf = d->n_fpts_per_inter;
  const int nf = I.nf, ni = I.n_inters;
  std::vector<int> il((size_t)nf * ni), lt((size_t)nf * ni), lut;
  std::vector<int8_t> tl(ni);
  int nfields = c->prm.n_fields, nd = c->prm.n_dims;
  for (int i = 0; i < ni; i++)
  {
    const hf_eles_dev &L = c->eles[d->ele_type_l[i]];
    if (!L.present) HF_FAIL("interface refers to an element type that was not uploaded");
    if (fill_lut(d->inter_type, c->prm.order, nf, d->rot_tag[i], lut)) return 1;
    tl[i] = (int8_t)d->ele_type_l[i];
    for (int j = 0; j < nf; j++)
    {
      il[j + (size_t)nf * i] = flat_fpt(L, d->ele_l[i], d->local_inter_l[i], j);
      lt[j + (size_t)nf * i] = lut[j];
    }
  }
  if (hf_alloc_copy(c, &I.idx_l, il.data(), il.size())) return 1;
  if (hf_alloc_copy(c, &I.lut, lt.data(), lt.size())) return 1;
  if (hf_alloc_copy(c, &I.type_l, tl.data(), tl.size())) return 1;
  I.nb_rank.assign(d->neighbour_rank, d->neighbour_rank + d->n_neighbours);
  I.nb_count.assign(d->neighbour_count, d->neighbour_count + d->n_neighbours);
  size_t nb = (size_t)ni * nf * nfields;
  if (hf_alloc_zero(c, &I.out_disu, nb)) return 1;
  if (hf_alloc_zero(c, &I.in_disu, nb)) return 1;
  if (c->prm.viscous)
  {
    if (hf_alloc_zero(c, &I.out_grad, nb * nd)) return 1;
    if (hf_alloc_zero(c, &I.in_grad, nb * nd)) return 1;
    if (c->prm.LES && (hf_alloc_zero(c, &I.out_sgsf, nb * nd) || hf_alloc_zero(c, &I.in_sgsf, nb * nd))) return 1;
  }
  if (d->ele_global_l) I.h_gid.assign(d->ele_global_l, d->ele_global_l + ni); else I.h_gid.clear();
  I.h_ele_type_l.assign(d->ele_type_l, d->ele_type_l + ni); I.h_ele_l.assign(d->ele_l, d->ele_l + ni);
  for (int i = 0; i < ni; i++) I.h_ele_l[i] = dev_ele(c->eles[I.h_ele_type_l[i]], I.h_ele_l[i]);
  I.h_loc_l.assign(d->local_inter_l, d->local_inter_l + ni);
  I.h_rot.assign(d->rot_tag, d->rot_tag + ni);
  return 0;
}

int hf_dev_finalize_setup(hf_ctx *c)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (c->fused && hf_fused_prepare(c)) return 1;
  c->finalized = true;
  HF_CUDA(cudaDeviceSynchronize());
  // the communicator may have arrived first (INTEGRATION.md allows either order): agree across the ranks now
  if (c->nproc > 1 && c->nccl_comm && !c->nccl_reconciled && hf_fused_after_nccl(c)) return 1;
  return 0;
}

int hf_dev_set_mode(hf_ctx *c, int fused)
{
  c->fused = fused;
  return 0;
}

// ---- the staged element methods ----------------------------------------------------------------------------------------
static hf_ell3 ell1(const hf_ell &E)
{
  hf_ell3 r;
  memset(&r, 0, sizeof(r));
  r.n = 1; r.rows = E.rows; r.cols = E.cols; r.nnz[0] = E.nnz; r.val[0] = E.val; r.col[0] = E.col; r.dval[0] = E.dval;
  return r;
}
static hf_ell3 elln(const hf_ell *E, int n)
{
  hf_ell3 r;
  memset(&r, 0, sizeof(r));
  r.n = n; r.rows = E[0].rows; r.cols = E[0].cols;
  for (int d = 0; d < n; d++) { r.nnz[d] = E[d].nnz; r.val[d] = E[d].val; r.col[d] = E[d].col; r.dval[d] = E[d].dval; }
  return r;
}

static int op_apply(hf_ctx *c, const hf_ell3 &E, const double *in, size_t in_dim_stride, double *out, long long n_cols, bool acc)
{
  long long n = (long long)E.rows * n_cols;
  // fast mode: dense operators (no exact zero dropped from any row: simplex / prism bases) go to the tensor-core kernel
  bool dense = c->fused != 0 && E.rows >= 8 && E.cols >= 4 && !getenv("HF_NO_DMMA");
  for (int d = 0; d < E.n && dense; d++) dense = E.dval[d] != nullptr;
  if (dense)
  {
    int cols_pad = E.cols;
    while (cols_pad % 8 != 4) cols_pad++; // column stride = 4 (mod 8) doubles: the B fragments of a half-warp hit 16 distinct banks
    const size_t smem = (size_t)E.n * OPD_TC * cols_pad * sizeof(double);
    if (smem <= 160 * 1024)
    {
      static bool attr_done = false;
      if (!attr_done)
      {
        HF_CUDA(cudaFuncSetAttribute(k_op_dense_mma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        HF_CUDA(cudaFuncSetAttribute(k_op_dense_mma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_done = true;
      }
      const unsigned grid = (unsigned)((n_cols + OPD_TC - 1) / OPD_TC);
      if (acc) k_op_dense_mma<true><<<grid, 128, smem, c->stream>>>(E, in, in_dim_stride, out, n_cols, cols_pad);
      else k_op_dense_mma<false><<<grid, 128, smem, c->stream>>>(E, in, in_dim_stride, out, n_cols, cols_pad);
      HF_LAUNCH_CHECK(c);
      return 0;
    }
  }
  if (acc) k_op_apply<true><<<hf_blocks(n, 256), 256, 0, c->stream>>>(E, in, in_dim_stride, out, n_cols);
  else k_op_apply<false><<<hf_blocks(n, 256), 256, 0, c->stream>>>(E, in, in_dim_stride, out, n_cols);
  HF_LAUNCH_CHECK(c);
  return 0;
}

int hf_dev_eles_op(hf_ctx *c, int ele_type, int op)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.present) return 0;
  HF_CUDA(cudaSetDevice(c->device));
  if (ensure_staged_buffers(c, e)) return 1;
  const int nd = e.n_dims, nfl = e.n_fields;
  const long long ncols = (long long)e.n_eles * nfl;
  const size_t NU = (size_t)e.n_upts * e.n_eles, NFP = (size_t)e.n_fpts * e.n_eles;
  const bool visc = c->prm.viscous != 0;
  switch (op)
  {
  case HF_EXTRAPOLATE_SOLUTION:
    return op_apply(c, ell1(e.opp_0), e.disu_upts[0], 0, e.disu_fpts, ncols, false);
  case HF_CALCULATE_GRADIENT:
    if (!visc) HF_FAIL("calculate_gradient called on an inviscid run");
    for (int d = 0; d < nd; d++)
      if (op_apply(c, ell1(e.opp_4[d]), e.disu_upts[0], 0, e.grad_disu_upts + d * NU * nfl, ncols, false)) return 1;
    return 0;
  case HF_EVALUATE_INVFLUX:
    HF_DISPATCH(nd, nfl, (k_point_flux<ND, NF, false><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.disu_upts[0], nullptr, e.JGinv_upts, e.tdisf_upts, c->phys)));
    HF_LAUNCH_CHECK(c);
    return 0;
  case HF_CORRECT_GRADIENT:
    if (!visc) HF_FAIL("correct_gradient called on an inviscid run");
    for (int d = 0; d < nd; d++)
      if (op_apply(c, ell1(e.opp_5[d]), e.delta_disu_fpts, 0, e.grad_disu_upts + d * NU * nfl, ncols, true)) return 1;
    for (int d = 0; d < nd; d++)
      if (op_apply(c, ell1(e.opp_6), e.grad_disu_upts + d * NU * nfl, 0, e.grad_disu_fpts + d * NFP * nfl, ncols, false)) return 1;
    HF_DISPATCH(nd, nfl, (k_transform_grad<ND, NF><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.grad_disu_upts, e.detjac_upts, e.JGinv_upts)));
    HF_LAUNCH_CHECK(c);
    HF_DISPATCH(nd, nfl, (k_transform_grad<ND, NF><<<hf_blocks(NFP, 128), 128, 0, c->stream>>>((long long)NFP, e.grad_disu_fpts, e.detjac_fpts, e.JGinv_fpts)));
    HF_LAUNCH_CHECK(c);
    return 0;
  case HF_EVALUATE_VISCFLUX:
    if (!visc) HF_FAIL("evaluate_viscFlux called on an inviscid run");
    if (c->prm.LES)
    {
      hf_les Q;
      Q.sgs_model = c->prm.SGS_model; Q.order = e.order; Q.C_s = c->prm.C_s; Q.Kappa = c->prm.Kappa; Q.prandtl_t = c->prm.prandtl_t;
      Q.filter_ratio = c->prm.filter_ratio; Q.vol_factor = e.ele_vol_factor; Q.gamma = c->prm.gamma;
      HF_DISPATCH(nd, nfl, (k_point_flux_les<ND, NF><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.n_upts, e.disu_upts[0], e.grad_disu_upts, e.JGinv_upts,
                                                                                          e.detjac_upts, e.wall_distance, e.Lu, e.Le, e.tdisf_upts, e.sgsf_upts, c->phys, Q)));
      HF_LAUNCH_CHECK(c);
      return 0;
    }
    HF_DISPATCH(nd, nfl, (k_point_flux<ND, NF, true><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.disu_upts[0], e.grad_disu_upts, e.JGinv_upts, e.tdisf_upts, c->phys)));
    HF_LAUNCH_CHECK(c);
    return 0;
  case HF_EXTRAPOLATE_TOTALFLUX:
    return op_apply(c, elln(e.opp_1, nd), e.tdisf_upts, NU * nfl, e.norm_tdisf_fpts, ncols, false);
  case HF_CALCULATE_DIVERGENCE:
    return op_apply(c, elln(e.opp_2, nd), e.tdisf_upts, NU * nfl, e.div_tconf_upts, ncols, false);
  case HF_CALCULATE_CORRECTED_DIVERGENCE:
  {
    long long n = (long long)NFP * nfl;
    k_sub<<<hf_blocks(n, 256), 256, 0, c->stream>>>(e.norm_tconf_fpts, e.norm_tdisf_fpts, n);
    HF_LAUNCH_CHECK(c);
    c->ufpts_valid = false;
    return op_apply(c, ell1(e.opp_3), e.norm_tconf_fpts, 0, e.div_tconf_upts, ncols, true);
  }
  case HF_EVALUATE_INVFLUX_OVER_INT:
  {
    if (!c->prm.over_int) HF_FAIL("evaluate_invFlux_over_int called without over_int");
    const size_t NC = (size_t)e.n_cub * e.n_eles;
    if (op_apply(c, ell1(e.opp_over_int), e.disu_upts[0], 0, e.u_cub, ncols, false)) return 1;
    HF_DISPATCH(nd, nfl, (k_point_flux<ND, NF, false><<<hf_blocks(NC, 128), 128, 0, c->stream>>>((long long)NC, e.u_cub, nullptr, e.JGinv_over_int, e.tdisf_cub, c->phys)));
    HF_LAUNCH_CHECK(c);
    for (int d = 0; d < nd; d++)
      if (op_apply(c, ell1(e.over_int_filter), e.tdisf_cub + d * NC * nfl, 0, e.tdisf_upts + d * NU * nfl, ncols, false)) return 1;
    return 0;
  }
  case HF_CALC_SGS_TERMS:
  {
    if (!c->prm.LES || c->prm.SGS_model < 2) HF_FAIL("calc_sgs_terms called without a filter-based SGS model");
    if (op_apply(c, ell1(e.filter_upts), e.disu_upts[0], 0, e.disuf_upts, ncols, false)) return 1;
    if (c->prm.SGS_model == 3)
    {
      // SVV: the filtered solution replaces the solution
      HF_CUDA(cudaMemcpyAsync(e.disu_upts[0], e.disuf_upts, NU * nfl * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      c->ufpts_valid = false;
      return 0;
    }
    const long long dim3 = nd == 2 ? 3 : 6;
    HF_DISPATCH(nd, nfl, (k_sgs_products<ND, NF><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.disu_upts[0], e.uu, e.ue)));
    HF_LAUNCH_CHECK(c);
    if (op_apply(c, ell1(e.filter_upts), e.uu, 0, e.Lu, (long long)e.n_eles * dim3, false)) return 1;
    if (op_apply(c, ell1(e.filter_upts), e.ue, 0, e.Le, (long long)e.n_eles * nd, false)) return 1;
    HF_DISPATCH(nd, nfl, (k_sgs_leonard<ND, NF><<<hf_blocks(NU, 128), 128, 0, c->stream>>>((long long)NU, e.disuf_upts, e.Lu, e.Le)));
    HF_LAUNCH_CHECK(c);
    return 0;
  }
  case HF_EXTRAPOLATE_SGSFLUX:
  {
    if (!c->prm.LES) HF_FAIL("extrapolate_sgsFlux called without LES");
    for (int d = 0; d < nd; d++)
      if (op_apply(c, ell1(e.opp_0), e.sgsf_upts + d * NU * nfl, 0, e.sgsf_fpts + d * NFP * nfl, ncols, false)) return 1;
    HF_DISPATCH(nd, nfl, (k_sgsf_physical<ND, NF><<<hf_blocks(NFP, 128), 128, 0, c->stream>>>((long long)NFP, e.sgsf_fpts, e.detjac_fpts, e.Jacobian_fpts)));
    HF_LAUNCH_CHECK(c);
    return 0;
  }
  case HF_SHOCK_CAPTURE:
  {
    if (!c->prm.shock_cap) HF_FAIL("shock_capture called without shock_cap");
    const int det_field = c->prm.shock_det_field == 0 ? 0 : nd + 1;
    const size_t smem = sizeof(double) * (size_t)e.n_upts * (1 + nfl);
    k_shock_capture<<<e.n_eles, 128, smem, c->stream>>>(e.n_upts, e.n_eles, nfl, det_field, c->prm.s0, e.disu_upts[0], e.inv_vandermonde, e.sensor_w_top,
                                                       e.sensor_w_all, e.exp_filter, e.sensor);
    HF_LAUNCH_CHECK(c);
    c->ufpts_valid = false;
    return 0;
  }
  default:
    HF_FAIL("unknown element operation");
  }
}

int hf_dev_int_inters_op(hf_ctx *c, int inter_type, int op)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  hf_int_inters_dev &I = c->ints[inter_type];
  if (I.n_inters == 0) return 0;
  HF_CUDA(cudaSetDevice(c->device));
  hf_views W = hf_make_views(c);
  int n = I.n_inters * I.nf;
  const int nd = c->prm.n_dims, nfl = c->prm.n_fields;
  if (op == HF_COMMON_INVFLUX)
    HF_DISPATCH(nd, nfl, (k_int_invflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.idx_r, I.type_l, I.type_r, c->phys, c->prm.viscous)));
  else if (op == HF_COMMON_VISCFLUX)
    HF_DISPATCH(nd, nfl, (k_int_viscflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.idx_r, I.type_l, I.type_r, c->phys, c->prm.LES)));
  else
    HF_FAIL("unknown interface operation");
  HF_LAUNCH_CHECK(c);
  return 0;
}

int hf_dev_bdy_inters_op(hf_ctx *c, int inter_type, int op, double time)
{
  (void)time; // the reference passes FlowSol->time; only the (out-of-scope) pressure ramp and turbulent inlet read it
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  hf_bdy_inters_dev &I = c->bdys[inter_type];
  if (I.n_inters == 0) return 0;
  HF_CUDA(cudaSetDevice(c->device));
  hf_views W = hf_make_views(c);
  int n = I.n_inters * I.nf;
  const int nd = c->prm.n_dims, nfl = c->prm.n_fields;
  if (op == HF_COMMON_INVFLUX)
    HF_DISPATCH(nd, nfl, (k_bdy_invflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.type_l, I.bc_id, c->bc_table, c->phys, c->prm.R_ref, c->prm.viscous)));
  else if (op == HF_COMMON_VISCFLUX)
  {
    hf_wm wmq;
    wmq.wall_model = c->prm.wall_model; wmq.gamma = c->prm.gamma; wmq.prandtl = c->prm.prandtl; wmq.prandtl_t = c->prm.prandtl_t; wmq.rt_inf = c->prm.rt_inf;
    wmq.mu_inf = c->prm.mu_inf; wmq.c_sth = c->prm.c_sth; wmq.fix_vis = (double)c->prm.fix_vis; wmq.Kappa = c->prm.Kappa;
    HF_DISPATCH(nd, nfl, (k_bdy_viscflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.type_l, I.bc_id, c->bc_table, c->phys, c->prm.R_ref, I.wm_upt, I.wm_dist, wmq)));
  }
  else
    HF_FAIL("unknown interface operation");
  HF_LAUNCH_CHECK(c);
  return 0;
}

int hf_dev_mpi_inters_op(hf_ctx *c, int inter_type, int op)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  hf_mpi_inters_dev &I = c->mpis[inter_type];
  if (I.n_inters == 0) return 0;
  HF_CUDA(cudaSetDevice(c->device));
  hf_views W = hf_make_views(c);
  int n = I.n_inters * I.nf;
  const int nd = c->prm.n_dims, nfl = c->prm.n_fields;
  switch (op)
  {
  case HF_COMMON_INVFLUX:
    HF_DISPATCH(nd, nfl, (k_mpi_invflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.type_l, I.lut, I.in_disu, c->phys, c->prm.viscous)));
    HF_LAUNCH_CHECK(c);
    return 0;
  case HF_COMMON_VISCFLUX:
    HF_DISPATCH(nd, nfl, (k_mpi_viscflux<ND, NF><<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, I.idx_l, I.type_l, I.lut, I.in_disu, I.in_grad, c->prm.LES ? I.in_sgsf : nullptr, c->phys)));
    HF_LAUNCH_CHECK(c);
    return 0;
  case 2: // send_solution: pack, then post the exchange on the comm stream
    k_pack_disu<<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, nfl, I.idx_l, I.type_l, I.out_disu);
    HF_LAUNCH_CHECK(c);
    return hf_halo_post(c, I, I.out_disu, I.in_disu, (size_t)I.nf * nfl);
  case 3: // receive_solution
    return hf_halo_wait(c);
  case 4: // send_corrected_gradient
    k_pack_grad<<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, nfl, nd, I.idx_l, I.type_l, I.out_grad);
    HF_LAUNCH_CHECK(c);
    return hf_halo_post(c, I, I.out_grad, I.in_grad, (size_t)I.nf * nfl * nd);
  case 5: // receive_corrected_gradient
    return hf_halo_wait(c);
  case 6: // send_sgsf_fpts (LES)
    if (!c->prm.LES) HF_FAIL("send_sgsf_fpts called without LES");
    k_pack_sgsf<<<hf_blocks(n, 128), 128, 0, c->stream>>>(W, n, I.nf, nfl, nd, I.idx_l, I.type_l, I.out_sgsf);
    HF_LAUNCH_CHECK(c);
    return hf_halo_post(c, I, I.out_sgsf, I.in_sgsf, (size_t)I.nf * nfl * nd);
  case 7: // receive_sgsf_fpts
    return hf_halo_wait(c);
  default:
    HF_FAIL("unknown partition-interface operation");
  }
}

// ---- CalcResidual / AdvanceSolution --------------------------------------------------------------------------------------
static int staged_residual(hf_ctx *c, double time, int rk_stage)
{
  const bool visc = c->prm.viscous != 0;
  const bool par = c->nproc > 1;
#define EACH_ELE(OP) for (int t = 0; t < HF_N_ELE_TYPES; t++) if (c->eles[t].present && hf_dev_eles_op(c, t, OP)) return 1
#define EACH_INT(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_int_inters_op(c, t, OP)) return 1
#define EACH_BDY(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_bdy_inters_op(c, t, OP, time)) return 1
#define EACH_MPI(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_mpi_inters_op(c, t, OP)) return 1
  // first RK stage of a step: filtered solution, Leonard tensors (reference src/solver.cpp:54-62)
  if (c->prm.LES && c->prm.SGS_model >= 2 && (rk_stage & 0xff) == 0) EACH_ELE(HF_CALC_SGS_TERMS);
  EACH_ELE(HF_EXTRAPOLATE_SOLUTION);
  if (par) EACH_MPI(2);
  if (visc) EACH_ELE(HF_CALCULATE_GRADIENT);
  EACH_ELE(c->prm.over_int ? HF_EVALUATE_INVFLUX_OVER_INT : HF_EVALUATE_INVFLUX);
  EACH_INT(HF_COMMON_INVFLUX);
  EACH_BDY(HF_COMMON_INVFLUX);
  if (par) { EACH_MPI(3); EACH_MPI(HF_COMMON_INVFLUX); }
  if (visc)
  {
    EACH_ELE(HF_CORRECT_GRADIENT);
    if (par) EACH_MPI(4);
    EACH_ELE(HF_EVALUATE_VISCFLUX);
    if (c->prm.LES) EACH_ELE(HF_EXTRAPOLATE_SGSFLUX);
    if (c->prm.LES && par) EACH_MPI(6);
  }
  EACH_ELE(HF_EXTRAPOLATE_TOTALFLUX);
  EACH_ELE(HF_CALCULATE_DIVERGENCE);
  if (visc)
  {
    EACH_INT(HF_COMMON_VISCFLUX);
    EACH_BDY(HF_COMMON_VISCFLUX);
    if (par) { EACH_MPI(5); if (c->prm.LES) EACH_MPI(7); EACH_MPI(HF_COMMON_VISCFLUX); }
  }
  EACH_ELE(HF_CALCULATE_CORRECTED_DIVERGENCE);
  return 0;
}

int hf_dev_calc_residual(hf_ctx *c, int rk_stage, double time)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  if (c->fused && hf_fused_available(c)) return hf_fused_stage(c, rk_stage, time, 1, 0);
  return staged_residual(c, time, rk_stage);
}

// what one stage of the time scheme does to (u0, u1), as the parameters of k_rk_update (reference src/eles.cpp:1080-1265)
extern "C++" int hf_rk_coeffs(hf_ctx *c, int stage, int *mode_out, int *copy_out, double *fac_out, double *c1_out, double *c2_out)
{
  const hf_params &p = c->prm;
  int mode = 0, copy = 0;
  double fac = 1.0, c1 = 0., c2 = 0.;
  if (p.adv_type == 0) { mode = 0; fac = 1.0; }
  else if (p.adv_type == 1)
  {
    copy = stage == 0;
    if (stage < 3) { mode = 0; fac = 3.0; }
    else { mode = 1; fac = 4.0; c1 = 3.0 / 4.0; c2 = 1.0 / 4.0; }
  }
  else if (p.adv_type == 2)
  {
    copy = stage == 0;
    if (stage < 2 || stage == 3) { mode = 0; fac = 2.0; }
    else { mode = 1; fac = 6.0; c1 = 1.0 / 3.0; c2 = 2.0 / 3.0; }
  }
  else if (p.adv_type == 3 || p.adv_type == 4)
  {
    if (stage < 0 || stage >= HF_MAX_RK) HF_FAIL("RK stage out of range");
    mode = 2; c1 = p.RK_a[stage]; c2 = p.RK_b[stage];
  }
  else
    HF_FAIL("ERROR: Time integration type not recognised ... ");
  *mode_out = mode; *copy_out = copy; *fac_out = fac; *c1_out = c1; *c2_out = c2;
  return 0;
}

static int advance_one(hf_ctx *c, hf_eles_dev &e, int stage)
{
  const hf_params &p = c->prm;
  long long n_pts = (long long)e.n_upts * e.n_eles, n = n_pts * e.n_fields;
  int mode = 0, copy = 0;
  double fac = 1.0, c1 = 0., c2 = 0.;
  if (hf_rk_coeffs(c, stage, &mode, &copy, &fac, &c1, &c2)) return 1;
  const double *dtl = (p.dt_type == 2) ? e.dt_local : nullptr;
  k_rk_update<<<hf_blocks(n, 256), 256, 0, c->stream>>>(n, n_pts, e.n_upts, e.disu_upts[0], e.disu_upts[1], e.div_tconf_upts, e.detjac_upts, dtl,
                                                        p.dt, fac, c1, c2, mode, copy, c->d_nan);
  HF_LAUNCH_CHECK(c);
  c->ufpts_valid = false;
  return 0;
}

// rk_stage: bits 0-7 the stage; bits 8.. = 1 + element type to advance one type only (0 = all types)
static int advance_all(hf_ctx *c, int rk_stage)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  HF_CUDA(cudaSetDevice(c->device));
  int stage = rk_stage & 0xff, only = (rk_stage >> 8) - 1;
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    if (!c->eles[t].present) continue;
    if (only >= 0 && t != only) continue;
    if (advance_one(c, c->eles[t], stage)) return 1;
  }
  return 0;
}
static int hf_check_nan(hf_ctx *c, bool collective = true);
// The method-by-method call sequence of the reference's main loop (CalcResidual, then AdvanceSolution per element type): the reference
// scans the residual for NaN in every calculate_corrected_divergence (src/eles.cpp:1781-1795); here the update kernel raises the
// device flag and this call reads it back -- one synchronisation per stage, which is what this path (the yardstick, not the fast
// path) can afford.  hf_dev_rk_stage / hf_dev_run_steps read the flag once per call instead.
int hf_dev_advance_solution(hf_ctx *c, int rk_stage)
{
  if (advance_all(c, rk_stage)) return 1;
  // no cross-rank agreement here: ranks without elements of a type do not make this call (eles::AdvanceSolution per element type), so a
  // collective would not match; a rank that finds NaN fails alone, as the reference's FatalError -> MPI_Abort does
  return hf_check_nan(c, false);
}

int hf_dev_rk_stage(hf_ctx *c, int rk_stage, double time, int keep_residual)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  if (c->fused && hf_fused_available(c))
  {
    if (hf_fused_stage(c, rk_stage, time, keep_residual, 1)) return 1;
  }
  else if (c->fused && hf_elem_available(c))
  {
    // every other element type / mesh in fast mode: blocked element kernels (hf_elem.cu) around the staged interface kernels
    if (hf_elem_stage(c, rk_stage, time, keep_residual)) return 1;
  }
  else
  {
    if (staged_residual(c, time, rk_stage)) return 1;
    if (advance_all(c, rk_stage)) return 1;
  }
  // shock capturing follows the update of every stage (reference src/HiFiLES.cpp:213-217)
  if (c->prm.shock_cap)
    for (int t = 0; t < HF_N_ELE_TYPES; t++)
      if (c->eles[t].present && hf_dev_eles_op(c, t, HF_SHOCK_CAPTURE)) return 1;
  return 0;
}

// reads the NaN flag back (all ranks agree) and fails, as the reference does, when it is set
static int hf_check_nan(hf_ctx *c, bool collective)
{
  HF_CUDA(cudaMemcpyAsync(c->h_nan, c->d_nan, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  double ok = *c->h_nan ? 0.0 : 1.0;
  if (collective && c->nproc > 1 && hf_halo_allreduce_min(c, &ok)) return 1; // every rank leaves together (a lone abort would leave the others in the next exchange)
  if (ok == 1.0) return 0;
  char msg[256];
  if (*c->h_nan)
    snprintf(msg, sizeof(msg), "Residual is NaN (element %d of rank %d, device order). Aborting...", *c->h_nan - 1, c->rank);
  else
    snprintf(msg, sizeof(msg), "Residual is NaN on another rank. Aborting...");
  HF_CUDA(cudaMemsetAsync(c->d_nan, 0, sizeof(int), c->stream));
  HF_FAIL(msg);
}

// ---- stage timeline ---------------------------------------------------------------------------------------------------------------
// prints, averaged over the recorded stages, the time of every mark relative to mark 0 (start of the stage on the compute stream):
//  0 stage start | 1 k_face interior done | 2 halo wait passed (face values arrived) | 3 k_face halo done | 4 k_resid interior done |
//  5 halo wait passed (common flux arrived) | 6 k_resid halo done      -- compute stream
//  7 / 8 exchange of the common flux: start / end | 9 / 10 exchange of the face values: start / end      -- communication stream
static void hf_timeline_report(hf_ctx *c)
{
  if (!c->tl_on || c->tl_ev.empty() || c->tl_stage < 2) return;
  cudaStreamSynchronize(c->stream);
  cudaStreamSynchronize(c->comm_stream);
  const int n = std::min(c->tl_stage, TL_SLOTS) - 1; // the oldest slot may be partly overwritten
  double sum[TL_MARKS] = {0};
  int cnt[TL_MARKS] = {0};
  for (int s = 0; s < n; s++)
  {
    const int slot = (c->tl_stage - 1 - s) % TL_SLOTS;
    for (int m = 1; m < TL_MARKS; m++)
    {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, c->tl_ev[slot * TL_MARKS], c->tl_ev[slot * TL_MARKS + m]) == cudaSuccess) { sum[m] += ms; cnt[m]++; }
      else cudaGetLastError();
    }
  }
  fprintf(stderr, "[stage timeline, rank %d, us after stage start, mean of %d stages]", c->rank, n);
  static const char *names[TL_MARKS] = {"start", "face_int", "wait_u", "face_halo", "resid_int", "wait_fc", "resid_halo", "xfc_begin", "xfc_end", "xu_begin", "xu_end", ""};
  for (int m = 1; m < 11; m++)
    if (cnt[m]) fprintf(stderr, " %s %.0f", names[m], 1e3 * sum[m] / cnt[m]);
  fprintf(stderr, "\n");
}

int hf_dev_run_steps(hf_ctx *c, int n_steps, double time0)
{
  c->tl_on = getenv("HF_STAGE_TIMELINE") != nullptr;
  if (c->prm.dt_type != 0) HF_FAIL("hf_dev_run_steps needs a fixed time step (dt_type 0)");
  double t = time0;
  static const bool no_guard = getenv("HF_NO_NAN_GUARD") != nullptr; // measurement aid
  for (int s = 0; s < n_steps; s++)
  {
    for (int i = 0; i < c->prm.n_rk; i++)
      if (hf_dev_rk_stage(c, i, t, (s == n_steps - 1 && i == c->prm.n_rk - 1) ? 1 : 0)) return 1;
    t += c->prm.dt;
  }
  // the reference scans the residual for NaN after every stage (src/eles.cpp:1781-1795); here the update kernels raise a sticky flag on
  // the device and the host reads it when the call's steps are enqueued -- one synchronisation per call instead of one per step (a
  // per-step read drains the launch pipeline: 2 % of a stage at 32 k elements per GPU)
  if (!no_guard && hf_check_nan(c)) return 1;
  hf_timeline_report(c);
  return 0;
}

// the NaN guard of hosts that issue hf_dev_rk_stage themselves (hf_dev_run_steps does this at its end)
int hf_dev_check_residual(hf_ctx *c) { return hf_check_nan(c); }

int hf_dev_set_dt(hf_ctx *c, double dt)
{
  c->prm.dt = dt;
  return 0;
}

int hf_dev_calc_dt(hf_ctx *c, double *dt_out)
{
  if (!c->finalized) HF_FAIL("hf_dev_finalize_setup has not been called");
  HF_CUDA(cudaSetDevice(c->device));
  if (c->prm.dt_type == 0) { *dt_out = c->prm.dt; return 0; }
  if (c->prm.equation != 0) HF_FAIL("CFL time step is only defined for the Euler / Navier-Stokes equations");
  double dt_min = 1e12;
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    hf_eles_dev &e = c->eles[t];
    if (!e.present) continue;
    if (!e.h_ref || !e.dt_local) HF_FAIL("h_ref was not uploaded: dt_type != 0 needs it");
    if (e.n_dims == 2)
      k_dt_local<2><<<hf_blocks(e.n_eles, 128), 128, 0, c->stream>>>(e.n_eles, e.n_upts, e.disu_upts[0], e.h_ref, e.dt_local, c->phys, c->prm.CFL, c->prm.order, c->prm.viscous);
    else
      k_dt_local<3><<<hf_blocks(e.n_eles, 128), 128, 0, c->stream>>>(e.n_eles, e.n_upts, e.disu_upts[0], e.h_ref, e.dt_local, c->phys, c->prm.CFL, c->prm.order, c->prm.viscous);
    HF_LAUNCH_CHECK(c);
    int nb = std::min(1024, (int)hf_blocks(e.n_eles, 256));
    k_min_reduce<<<nb, 256, 0, c->stream>>>(e.n_eles, e.dt_local, c->scratch);
    HF_LAUNCH_CHECK(c);
    std::vector<double> part(nb);
    HF_CUDA(cudaMemcpyAsync(part.data(), c->scratch, nb * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    for (double v : part) if (v < dt_min) dt_min = v;
  }
  if (c->nproc > 1 && hf_halo_allreduce_min(c, &dt_min)) return 1;
  c->prm.dt = dt_min;
  *dt_out = dt_min;
  return 0;
}

// ---- data movement -------------------------------------------------------------------------------------------------------
static int locate_array(hf_ctx *c, hf_eles_dev &e, int which, double **p, size_t *n)
{
  const size_t NU = (size_t)e.n_upts * e.n_eles, NFP = (size_t)e.n_fpts * e.n_eles, F = e.n_fields, D = e.n_dims;
  switch (which)
  {
  case HF_DISU_UPTS0: *p = e.disu_upts[0]; *n = NU * F; break;
  case HF_DISU_UPTS1: *p = e.disu_upts[1]; *n = NU * F; break;
  case HF_DIV_TCONF_UPTS: *p = e.div_tconf_upts; *n = NU * F; break;
  case HF_DISU_FPTS: *p = e.disu_fpts; *n = NFP * F; break;
  case HF_TDISF_UPTS: *p = e.tdisf_upts; *n = NU * F * D; break;
  case HF_NORM_TDISF_FPTS: *p = e.norm_tdisf_fpts; *n = NFP * F; break;
  case HF_NORM_TCONF_FPTS: *p = e.norm_tconf_fpts; *n = NFP * F; break;
  case HF_DELTA_DISU_FPTS: *p = e.delta_disu_fpts; *n = NFP * F; break;
  case HF_GRAD_DISU_UPTS: *p = e.grad_disu_upts; *n = NU * F * D; break;
  case HF_GRAD_DISU_FPTS: *p = e.grad_disu_fpts; *n = NFP * F * D; break;
  case HF_SRC_UPTS: *p = nullptr; *n = NU * F; break;
  case HF_DT_LOCAL: *p = e.dt_local; *n = e.n_eles; break;
  case HF_SENSOR: *p = e.sensor; *n = e.n_eles; break;
  case HF_SGSF_UPTS: *p = e.sgsf_upts; *n = NU * F * D; break;
  case HF_SGSF_FPTS: *p = e.sgsf_fpts; *n = NFP * F * D; break;
  case HF_DISUF_UPTS: *p = e.disuf_upts; *n = NU * F; break;
  case HF_LU: *p = e.Lu; *n = NU * (D == 2 ? 3 : 6); break;
  case HF_LE: *p = e.Le; *n = NU * D; break;
  case HF_DISU_AVERAGE_UPTS: *p = e.disu_average_upts; *n = NU * (size_t)e.n_average; break;
  default: HF_FAIL("unknown array id");
  }
  (void)c;
  return 0;
}

static size_t pts_per_ele(const hf_eles_dev &e, int which)
{
  switch (which)
  {
  case HF_DISU_FPTS: case HF_NORM_TDISF_FPTS: case HF_NORM_TCONF_FPTS: case HF_DELTA_DISU_FPTS: case HF_GRAD_DISU_FPTS: case HF_SGSF_FPTS: return e.n_fpts;
  case HF_DT_LOCAL: case HF_SENSOR: return 1;
  default: return e.n_upts;
  }
}

int hf_dev_download(hf_ctx *c, int ele_type, int which, double *host, size_t n_doubles)
{
  HF_CUDA(cudaSetDevice(c->device));
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.present) HF_FAIL("element type not present on the device");
  double *p; size_t n;
  if (locate_array(c, e, which, &p, &n)) return 1;
  if (n_doubles != n) HF_FAIL("download: size mismatch");
  if (which == HF_SRC_UPTS) { memset(host, 0, n * sizeof(double)); return 0; } // no source terms on the in-scope path
  if (!p) HF_FAIL("download: array is not materialised on the device in the current mode");
  // the fused path keeps face values in its own per-face layout: materialise the reference-layout array on request
  if (which == HF_DISU_FPTS && c->fused && hf_fused_available(c) &&
      op_apply(c, ell1(e.opp_0), e.disu_upts[0], 0, e.disu_fpts, (long long)e.n_eles * e.n_fields, false)) return 1;
  if (!e.pos.empty())
  {
    // device order -> host order in a staging buffer, then one flat copy
    if (ensure_stage(c, e, n) || permute_on_device(c, e, p, e.d_stage, pts_per_ele(e, which), n, false)) return 1;
    p = e.d_stage;
  }
  HF_CUDA(cudaMemcpyAsync(host, p, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int hf_dev_upload(hf_ctx *c, int ele_type, int which, const double *host, size_t n_doubles)
{
  HF_CUDA(cudaSetDevice(c->device));
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.present) HF_FAIL("element type not present on the device");
  double *p; size_t n;
  if (locate_array(c, e, which, &p, &n)) return 1;
  if (n_doubles != n) HF_FAIL("upload: size mismatch");
  if (!p)
  {
    if (ensure_staged_buffers(c, e) || locate_array(c, e, which, &p, &n)) return 1;
    if (!p) HF_FAIL("upload: array is not materialised on the device");
  }
  if (!e.pos.empty())
  {
    // flat copy into a staging buffer, then host order -> device order on the device
    if (ensure_stage(c, e, n)) return 1;
    HF_CUDA(cudaMemcpyAsync(e.d_stage, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (permute_on_device(c, e, e.d_stage, p, pts_per_ele(e, which), n, true)) return 1;
  }
  else
    HF_CUDA(cudaMemcpyAsync(p, host, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  c->ufpts_valid = false;
  return 0;
}

// Two-phase upload: the host -> device copy runs on a transfer stream while the compute stream is still busy with the previous
// time step; hf_dev_upload_commit then orders the compute stream behind the copy and moves the data into the array (device ->
// device, or the element permutation).  The reference's eles::cp_disu_upts_cpu_gpu (src/eles.cpp cp_* family) is synchronous; a
// host that feeds a new state every step (the end-to-end measurement of bench.py) hides the PCIe time behind the stages this way.
int hf_dev_upload_begin(hf_ctx *c, int ele_type, int which, const double *host, size_t n_doubles)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES || !c->eles[ele_type].present) HF_FAIL("element type not present on the device");
  hf_eles_dev &e = c->eles[ele_type];
  double *p; size_t n;
  if (locate_array(c, e, which, &p, &n)) return 1;
  if (n_doubles != n) HF_FAIL("upload: size mismatch");
  if (!p) HF_FAIL("upload: array is not materialised on the device");
  if (e.xfer_pending) HF_FAIL("hf_dev_upload_begin: the previous two-phase upload of this element type was not committed");
  if (!c->xfer_stream)
  {
    HF_CUDA(cudaStreamCreateWithFlags(&c->xfer_stream, cudaStreamNonBlocking));
    HF_CUDA(cudaEventCreateWithFlags(&c->ev_xfer, cudaEventDisableTiming));
    HF_CUDA(cudaEventCreateWithFlags(&c->ev_xfer_free, cudaEventDisableTiming));
  }
  if (e.xfer_n < n)
  {
    if (e.d_xfer) cudaFree(e.d_xfer);
    e.d_xfer = nullptr;
    cudaError_t err = cudaMalloc((void **)&e.d_xfer, n * sizeof(double));
    if (err != cudaSuccess) { hf_set_error(std::string("cudaMalloc (two-phase upload buffer): ") + cudaGetErrorString(err)); return 1; }
    e.xfer_n = n;
  }
  else
    HF_CUDA(cudaStreamWaitEvent(c->xfer_stream, c->ev_xfer_free, 0)); // the last commit has finished reading the buffer
  HF_CUDA(cudaMemcpyAsync(e.d_xfer, host, n * sizeof(double), cudaMemcpyHostToDevice, c->xfer_stream));
  HF_CUDA(cudaEventRecord(c->ev_xfer, c->xfer_stream));
  e.xfer_pending = true;
  e.xfer_which = which;
  return 0;
}

int hf_dev_upload_commit(hf_ctx *c, int ele_type)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES || !c->eles[ele_type].present) HF_FAIL("element type not present on the device");
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.xfer_pending) HF_FAIL("hf_dev_upload_commit without hf_dev_upload_begin");
  double *p; size_t n;
  if (locate_array(c, e, e.xfer_which, &p, &n)) return 1;
  HF_CUDA(cudaStreamWaitEvent(c->stream, c->ev_xfer, 0));
  if (!e.pos.empty())
  {
    if (permute_on_device(c, e, e.d_xfer, p, pts_per_ele(e, e.xfer_which), n, true)) return 1;
  }
  else
    HF_CUDA(cudaMemcpyAsync(p, e.d_xfer, n * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  HF_CUDA(cudaEventRecord(c->ev_xfer_free, c->stream));
  e.xfer_pending = false;
  c->ufpts_valid = false;
  return 0;
}

int hf_dev_set_volume_cubature(hf_ctx *c, int ele_type, int n_cubpts, const double *opp_volume_cubpts, const double *weights, const double *vol_detjac)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES || !c->eles[ele_type].present) HF_FAIL("element type not present on the device");
  hf_eles_dev &e = c->eles[ele_type];
  if (n_cubpts <= 0) HF_FAIL("volume cubature without points");
  e.n_vol_cub = n_cubpts;
  if (hf_alloc_copy(c, &e.opp_vol_cub, opp_volume_cubpts, (size_t)n_cubpts * e.n_upts)) return 1;
  if (hf_alloc_copy(c, &e.w_vol_cub, weights, (size_t)n_cubpts)) return 1;
  if (hf_alloc_copy(c, &e.detjac_vol_cub, vol_detjac, (size_t)n_cubpts * e.n_eles)) return 1;
  if (hf_alloc_zero(c, &e.iq_elem, (size_t)e.n_eles * HF_MAX_INTEGRAL_QUANTITIES)) return 1;
  c->want_gradient = true;
  return 0;
}

int hf_dev_time_average(hf_ctx *c, int ele_type, int n_average_fields, const int *kinds, double time, double spinup_time)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES || !c->eles[ele_type].present) HF_FAIL("element type not present on the device");
  hf_eles_dev &e = c->eles[ele_type];
  if (c->prm.equation != 0) HF_FAIL("time averages are defined for the Euler / Navier-Stokes equations");
  if (n_average_fields < 1 || n_average_fields > HF_MAX_INTEGRAL_QUANTITIES) HF_FAIL("number of average fields out of range");
  const long long n_pts = (long long)e.n_upts * e.n_eles;
  if (!e.disu_average_upts)
  {
    if (hf_alloc_zero(c, &e.disu_average_upts, (size_t)n_pts * n_average_fields)) return 1;
    e.n_average = n_average_fields;
  }
  if (e.n_average != n_average_fields) HF_FAIL("the number of average fields changed");
  hf_avg_kinds K;
  K.n = n_average_fields;
  for (int i = 0; i < n_average_fields; i++)
  {
    if (kinds[i] < 0 || kinds[i] > 4) HF_FAIL("average field not recognized");
    K.kind[i] = kinds[i];
  }
  k_time_average<<<hf_blocks(n_pts, 256), 256, 0, c->stream>>>(n_pts, e.n_upts, e.n_dims, e.disu_upts[0], e.disu_average_upts,
                                                             c->prm.dt_type == 2 ? e.dt_local : nullptr, c->prm.dt, time, spinup_time, K);
  HF_LAUNCH_CHECK(c);
  return 0;
}

int hf_dev_set_wall_distance(hf_ctx *c, int ele_type, const double *wall_distance, size_t n_doubles)
{
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES) HF_FAIL("bad element type");
  HF_CUDA(cudaSetDevice(c->device));
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.present) HF_FAIL("element type not present on the device");
  const size_t n = (size_t)e.n_upts * e.n_eles * e.n_dims;
  if (n_doubles != n) HF_FAIL("hf_dev_set_wall_distance: size mismatch");
  if (!e.wall_distance) HF_FAIL("hf_dev_set_wall_distance: the element type was uploaded without wall distances");
  std::vector<double> tmp;
  const double *src = wall_distance;
  if (!e.pos.empty())
  {
    tmp = permute_eles(wall_distance, (size_t)e.n_upts, e.n_eles, (size_t)e.n_dims, e.pos, true);
    src = tmp.data();
  }
  HF_CUDA(cudaMemcpyAsync(e.wall_distance, src, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

int hf_dev_set_keep_gradient(hf_ctx *c, int on)
{
  c->want_gradient = on != 0;
  return 0;
}

int hf_dev_integral_quantities(hf_ctx *c, int ele_type, int n_quantities, const int *kinds, double *out)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (ele_type < 0 || ele_type >= HF_N_ELE_TYPES || !c->eles[ele_type].present) HF_FAIL("element type not present on the device");
  hf_eles_dev &e = c->eles[ele_type];
  if (!e.n_vol_cub) HF_FAIL("hf_dev_set_volume_cubature has not been called for this element type");
  if (n_quantities < 1 || n_quantities > HF_MAX_INTEGRAL_QUANTITIES) HF_FAIL("number of integral quantities out of range");
  if (c->prm.equation != 0) HF_FAIL("integral quantities are defined for the Euler / Navier-Stokes equations");
  if (!e.grad_disu_upts)
    HF_FAIL("integral quantities need grad_disu_upts of the last residual evaluation (viscous run, residual kept on the monitored stage)");
  hf_iq_kinds K;
  K.n = n_quantities;
  for (int q = 0; q < n_quantities; q++)
  {
    if (kinds[q] < 0 || kinds[q] > 4) HF_FAIL("integral diagnostic quantity not recognized");
    K.kind[q] = kinds[q];
  }
  const size_t smem = (size_t)n_quantities * e.n_vol_cub * sizeof(double);
  const int *pos = e.pos.empty() ? nullptr : e.d_pos;
  if (e.n_dims == 2)
    k_integral_quantities<2><<<e.n_eles, 128, smem, c->stream>>>(e.n_eles, e.n_upts, e.n_vol_cub, e.disu_upts[0], e.grad_disu_upts, e.opp_vol_cub, e.w_vol_cub,
                                                                 e.detjac_vol_cub, pos, K, c->phys.gamma, e.iq_elem);
  else
    k_integral_quantities<3><<<e.n_eles, 128, smem, c->stream>>>(e.n_eles, e.n_upts, e.n_vol_cub, e.disu_upts[0], e.grad_disu_upts, e.opp_vol_cub, e.w_vol_cub,
                                                                 e.detjac_vol_cub, pos, K, c->phys.gamma, e.iq_elem);
  HF_LAUNCH_CHECK(c);
  std::vector<double> h((size_t)e.n_eles * n_quantities);
  HF_CUDA(cudaMemcpyAsync(h.data(), e.iq_elem, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  for (int i = 0; i < e.n_eles; i++)
    for (int q = 0; q < n_quantities; q++) out[q] += h[(size_t)i * n_quantities + q];
  return 0;
}

int hf_dev_residual_norm(hf_ctx *c, int norm_type, double *out)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (norm_type < 0 || norm_type > 2) HF_FAIL("norm_type not recognized");
  const int nb = 512;
  for (int f = 0; f < c->prm.n_fields; f++) out[f] = 0.;
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    hf_eles_dev &e = c->eles[t];
    if (!e.present) continue;
    long long n_pts = (long long)e.n_upts * e.n_eles;
    for (int f = 0; f < e.n_fields; f++)
    {
      const double *div = e.div_tconf_upts + (size_t)f * n_pts;
      double *part = c->scratch + (size_t)f * nb;
      if (norm_type == 0) k_res_norm<0><<<nb, 256, 0, c->stream>>>(n_pts, div, e.detjac_upts, part);
      else if (norm_type == 1) k_res_norm<1><<<nb, 256, 0, c->stream>>>(n_pts, div, e.detjac_upts, part);
      else k_res_norm<2><<<nb, 256, 0, c->stream>>>(n_pts, div, e.detjac_upts, part);
      HF_LAUNCH_CHECK(c);
    }
    std::vector<double> h((size_t)nb * e.n_fields);
    HF_CUDA(cudaMemcpyAsync(h.data(), c->scratch, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
    for (int f = 0; f < e.n_fields; f++)
    {
      double s = 0.;
      for (int b = 0; b < nb; b++) s = (norm_type == 0) ? std::max(s, h[(size_t)f * nb + b]) : s + h[(size_t)f * nb + b];
      out[f] = (norm_type == 0) ? std::max(out[f], s) : out[f] + s;
    }
  }
  return 0;
}

int hf_dev_sync(hf_ctx *c)
{
  if (!c) HF_FAIL("no device context");
  HF_CUDA(cudaSetDevice(c->device));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  HF_CUDA(cudaStreamSynchronize(c->comm_stream));
  HF_CUDA(cudaGetLastError());
  return 0;
}

long long hf_dev_launch_count(hf_ctx *c) { return c ? c->launches : 0; }

int hf_dev_kernel_timer(hf_ctx *c, int mode, double *ms_total, long long *n_launches)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (mode == 1) { c->ktimer_on = true; c->kt_used = 0; }
  else if (mode == 0) c->ktimer_on = false;
  double tot = 0.;
  if (mode == -1 || mode == 0)
  {
    HF_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i + 1 < c->kt_used; i += 2)
    {
      float ms = 0.f;
      HF_CUDA(cudaEventElapsedTime(&ms, c->kt_ev[i], c->kt_ev[i + 1]));
      tot += ms;
    }
  }
  if (ms_total) *ms_total = tot;
  if (n_launches) *n_launches = (long long)(c->kt_used / 2);
  return 0;
}

int hf_dev_timer_start(hf_ctx *c)
{
  HF_CUDA(cudaEventRecord(c->ev_t0, c->stream));
  return 0;
}
int hf_dev_timer_stop(hf_ctx *c, float *ms)
{
  HF_CUDA(cudaEventRecord(c->ev_t1, c->stream));
  HF_CUDA(cudaEventSynchronize(c->ev_t1));
  HF_CUDA(cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
  return 0;
}

} // extern "C"
