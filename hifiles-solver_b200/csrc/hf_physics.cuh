// Pointwise physics shared by every kernel: inviscid / viscous fluxes, Riemann solvers, LDG common values.
// Each function restates the reference's CPU routine with the same operation order (the reference's dead CUDA
// file has no HLLC, no RoeM and a 2-D-only Roe: SURVEY.md §2c).
//   calc_invf_2d/3d   reference src/flux.cpp:33-125
//   calc_visf_2d/3d   reference src/flux.cpp:129-422
//   rusanov_flux      reference src/inters.cpp:277-324
//   roeM_flux         reference src/inters.cpp:327-437
//   hllc_flux         reference src/inters.cpp:439-532
//   lax_friedrich     reference src/inters.cpp:535-557
//   ldg_flux/solution reference src/inters.cpp:561-646
#pragma once
#include <cuda_runtime.h>

struct hf_phys
{
  double gamma, prandtl, mu_inf, rt_inf, c_sth, fix_vis;
  double ldg_beta, ldg_tau;
  double wave_speed[3], diff_coeff, lambda;
  int riemann_solve_type;
  double gamma_over_pr; // (1 / Pr) * gamma, used by the fused kernels
};

// F(k,d) stored as f[k + NF*d]
template <int ND, int NF>
__device__ __forceinline__ void inv_flux(const double *__restrict__ u, double *__restrict__ f, const hf_phys &P)
{
  if (NF == 1)
  {
#pragma unroll
    for (int d = 0; d < ND; d++) f[d] = P.wave_speed[d] * u[0];
    return;
  }
  double v[ND];
  double vsq = 0.;
#pragma unroll
  for (int d = 0; d < ND; d++)
  {
    v[d] = u[d + 1] / u[0];
    vsq += v[d] * v[d];
  }
  double p = (P.gamma - 1.0) * (u[ND + 1] - (0.5 * u[0] * vsq));
#pragma unroll
  for (int d = 0; d < ND; d++)
  {
    f[0 + NF * d] = u[d + 1];
#pragma unroll
    for (int k = 0; k < ND; k++) f[(k + 1) + NF * d] = (k == d) ? p + (u[k + 1] * v[d]) : u[k + 1] * v[d];
    f[(ND + 1) + NF * d] = v[d] * (u[ND + 1] + p);
  }
}

// grad(k,d) stored as g[k + NF*d]
template <int ND, int NF>
__device__ __forceinline__ void vis_flux(const double *__restrict__ u, const double *__restrict__ g, double *__restrict__ f, const hf_phys &P)
{
  if (NF == 1)
  {
#pragma unroll
    for (int d = 0; d < ND; d++) f[d] = -P.diff_coeff * g[d];
    return;
  }
  double rho = u[0], ene = u[ND + 1];
  double v[ND];
  double vsq = 0.;
#pragma unroll
  for (int d = 0; d < ND; d++)
  {
    v[d] = u[d + 1] / rho;
    vsq += v[d] * v[d];
  }
  double inte = ene / rho - 0.5 * vsq;
  double rt_ratio = (P.gamma - 1.0) * inte / (P.rt_inf);
  double mu = (P.mu_inf) * pow(rt_ratio, 1.5) * (1. + (P.c_sth)) / (rt_ratio + (P.c_sth));
  mu = mu + P.fix_vis * (P.mu_inf - mu);
  // velocity gradients dv[i][d] = d v_i / d x_d
  double dv[ND][ND];
#pragma unroll
  for (int i = 0; i < ND; i++)
#pragma unroll
    for (int d = 0; d < ND; d++) dv[i][d] = (g[(i + 1) + NF * d] - g[0 + NF * d] * v[i]) / rho;
  double de[ND];
#pragma unroll
  for (int d = 0; d < ND; d++)
  {
    double s = 0.;
#pragma unroll
    for (int i = 0; i < ND; i++) s += v[i] * dv[i][d];
    double dke = 0.5 * vsq * g[0 + NF * d] + rho * s;
    de[d] = (g[(ND + 1) + NF * d] - dke - g[0 + NF * d] * inte) / rho;
  }
  double trace = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++) trace += dv[i][i];
  double diag = trace / 3.0;
  double tau[ND][ND];
#pragma unroll
  for (int i = 0; i < ND; i++)
#pragma unroll
    for (int d = 0; d < ND; d++)
      tau[i][d] = (i == d) ? 2.0 * mu * (dv[i][i] - diag) : mu * (dv[i < d ? i : d][i < d ? d : i] + dv[i < d ? d : i][i < d ? i : d]);
  double kap = (mu / P.prandtl) * (P.gamma);
#pragma unroll
  for (int d = 0; d < ND; d++)
  {
    f[0 + NF * d] = 0.0;
    double w = 0.;
#pragma unroll
    for (int i = 0; i < ND; i++)
    {
      f[(i + 1) + NF * d] = -tau[i][d];
      w += v[i] * tau[i][d];
    }
    f[(ND + 1) + NF * d] = -(w + kap * de[d]);
  }
}

template <int ND, int NF>
__device__ __forceinline__ void normal_flux(const double *__restrict__ f, const double *__restrict__ n, double *__restrict__ fn)
{
#pragma unroll
  for (int k = 0; k < NF; k++)
  {
    double s = 0.;
#pragma unroll
    for (int d = 0; d < ND; d++) s += f[k + NF * d] * n[d];
    fn[k] = s;
  }
}

template <int ND, int NF>
__device__ __forceinline__ void rusanov_flux(const double *u_l, const double *u_r, const double *f_l, const double *f_r, const double *n, double *fn, const hf_phys &P)
{
  double fn_l[NF], fn_r[NF];
  normal_flux<ND, NF>(f_l, n, fn_l);
  normal_flux<ND, NF>(f_r, n, fn_r);
  double vn_l = 0, vn_r = 0, vsq_l = 0, vsq_r = 0;
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    double a = u_l[i + 1] / u_l[0], b = u_r[i + 1] / u_r[0];
    vn_l += a * n[i];
    vn_r += b * n[i];
    vsq_l += a * a;
    vsq_r += b * b;
  }
  double p_l = (P.gamma - 1.0) * (u_l[ND + 1] - 0.5 * u_l[0] * vsq_l);
  double p_r = (P.gamma - 1.0) * (u_r[ND + 1] - 0.5 * u_r[0] * vsq_r);
  double eig = sqrt(P.gamma * (p_l + p_r) / (u_l[0] + u_r[0])) + 0.5 * fabs(vn_l + vn_r);
#pragma unroll
  for (int k = 0; k < NF; k++) fn[k] = 0.5 * ((fn_l[k] + fn_r[k]) - eig * (u_r[k] - u_l[k]));
}

template <int ND, int NF>
__device__ __forceinline__ void roeM_flux(const double *u_l, const double *u_r, const double *f_l, const double *f_r, const double *n, double *fn, const hf_phys &P)
{
  const double gamma = P.gamma;
  double v_l[ND], v_r[ND], va[ND], dv[ND], du[NF], bdq[NF], fn_l[NF], fn_r[NF];
  double vn_l = 0., vsq_l = 0., vn_r = 0., vsq_r = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    v_l[i] = u_l[i + 1] / u_l[0];
    v_r[i] = u_r[i + 1] / u_r[0];
    vn_l += v_l[i] * n[i];
    vn_r += v_r[i] * n[i];
    vsq_l += v_l[i] * v_l[i];
    vsq_r += v_r[i] * v_r[i];
    dv[i] = v_r[i] - v_l[i];
  }
  double p_l = (gamma - 1.0) * (u_l[ND + 1] - 0.5 * u_l[0] * vsq_l);
  double p_r = (gamma - 1.0) * (u_r[ND + 1] - 0.5 * u_r[0] * vsq_r);
  double h_l = (u_l[ND + 1] + p_l) / u_l[0];
  double h_r = (u_r[ND + 1] + p_r) / u_r[0];
  double drho = u_r[0] - u_l[0];
  double dp = p_r - p_l;
  double dh = h_r - h_l;
  double dvn = vn_r - vn_l;
  double sq_rho = sqrt(u_r[0] / u_l[0]);
  double rrho = 1.0 / (1.0 + sq_rho);
  double ratr = sq_rho * rrho;
  double ra = sq_rho * u_l[0];
  double ha = h_l * rrho + h_r * ratr;
  double qq = 0., va_n = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    va[i] = v_l[i] * rrho + v_r[i] * ratr;
    qq += va[i] * va[i];
    va_n += n[i] * va[i];
  }
  double aa = sqrt((gamma - 1) * (ha - 0.5 * qq));
  double rcp_aa = 1.0 / aa;
  double abs_ma = fabs(va_n * rcp_aa);
  double b1 = fmax(0.0, fmax(va_n + aa, vn_r + aa));
  double b2 = fmin(0.0, fmin(va_n - aa, vn_l - aa));
  double b1b2 = b1 * b2;
  double rcp_b1_b2 = 1.0 / (b1 - b2);
  b1 = b1 * rcp_b1_b2;
  b2 = b2 * rcp_b1_b2;
  b1b2 = b1b2 * rcp_b1_b2;
  double h = 1.0 - ((p_l < p_r) ? (p_l / p_r) : (p_r / p_l));
  double f = ((abs_ma != 0) ? pow(abs_ma, h) : 1.);
  double g = f / (1.0 + abs_ma);
#pragma unroll
  for (int i = 0; i < NF - 1; i++) du[i] = u_r[i] - u_l[i];
  du[ND + 1] = u_r[0] * h_r - u_l[0] * h_l;
  bdq[0] = drho - f * dp * rcp_aa * rcp_aa;
  bdq[ND + 1] = bdq[0] * ha + ra * dh;
#pragma unroll
  for (int i = 0; i < ND; i++) bdq[i + 1] = bdq[0] * va[i] + ra * (dv[i] - n[i] * dvn);
  normal_flux<ND, NF>(f_l, n, fn_l);
  normal_flux<ND, NF>(f_r, n, fn_r);
#pragma unroll
  for (int i = 0; i < NF; i++) fn[i] = (b1 * fn_l[i] - b2 * fn_r[i]) + b1b2 * (du[i] - g * bdq[i]);
}

template <int ND, int NF>
__device__ __forceinline__ void hllc_flux(const double *u_l, const double *u_r, const double *f_l, const double *f_r, const double *n, double *fn, const hf_phys &P)
{
  const double gamma = P.gamma;
  double fn_l[NF], fn_r[NF];
  double vn_l = 0., vsq_l = 0., vn_r = 0., vsq_r = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++)
  {
    double a = u_l[i + 1] / u_l[0], b = u_r[i + 1] / u_r[0];
    vn_l += a * n[i];
    vn_r += b * n[i];
    vsq_l += a * a;
    vsq_r += b * b;
  }
  double p_l = (gamma - 1.0) * (u_l[ND + 1] - 0.5 * u_l[0] * vsq_l);
  double p_r = (gamma - 1.0) * (u_r[ND + 1] - 0.5 * u_r[0] * vsq_r);
  double h_l = (u_l[ND + 1] + p_l) / u_l[0];
  double h_r = (u_r[ND + 1] + p_r) / u_r[0];
  normal_flux<ND, NF>(f_l, n, fn_l);
  normal_flux<ND, NF>(f_r, n, fn_r);
  double sq_rho = sqrt(u_r[0] / u_l[0]);
  double rrho = 1. / (sq_rho + 1.);
  double vn_m = rrho * (vn_l + sq_rho * vn_r);
  double h_m = rrho * (h_l + sq_rho * h_r);
  double a_m = sqrt((gamma - 1.) * (h_m - 0.5 * vn_m * vn_m));
  double S_R = vn_m + a_m;
  double S_L = vn_m - a_m;
  double S_star = (p_r - p_l + u_l[0] * vn_l * (S_L - vn_l) - u_r[0] * vn_r * (S_R - vn_r)) / (u_l[0] * (S_L - vn_l) - u_r[0] * (S_R - vn_r));
  if (S_L >= 0)
  {
#pragma unroll
    for (int k = 0; k < NF; k++) fn[k] = fn_l[k];
  }
  else if (S_star >= 0)
  {
    double rcp_star = S_L - S_star;
    double pst = (p_l + u_l[0] * (S_L - vn_l) * (S_star - vn_l));
    fn[0] = S_star * (S_L * u_l[0] - fn_l[0]) / rcp_star;
#pragma unroll
    for (int i = 0; i < ND; i++) fn[i + 1] = (S_star * (S_L * u_l[i + 1] - fn_l[i + 1]) + S_L * pst * n[i]) / rcp_star;
    fn[ND + 1] = (S_star * (S_L * u_l[ND + 1] - fn_l[ND + 1]) + S_L * pst * S_star) / rcp_star;
  }
  else if (S_R >= 0)
  {
    double rcp_star = S_R - S_star;
    double pst = (p_r + u_r[0] * (S_R - vn_r) * (S_star - vn_r));
    fn[0] = S_star * (S_R * u_r[0] - fn_r[0]) / rcp_star;
#pragma unroll
    for (int i = 0; i < ND; i++) fn[i + 1] = (S_star * (S_R * u_r[i + 1] - fn_r[i + 1]) + S_R * pst * n[i]) / rcp_star;
    fn[ND + 1] = (S_star * (S_R * u_r[ND + 1] - fn_r[ND + 1]) + S_R * pst * S_star) / rcp_star;
  }
  else
  {
#pragma unroll
    for (int k = 0; k < NF; k++) fn[k] = fn_r[k];
  }
}

template <int ND>
__device__ __forceinline__ void lax_friedrich(const double *u_l, const double *u_r, const double *n, double *fn, const hf_phys &P)
{
  double u_av = 0.5 * (u_l[0] + u_r[0]);
  double u_diff = (u_l[0] - u_r[0]);
  double norm_speed = 0;
#pragma unroll
  for (int i = 0; i < ND; i++) norm_speed += P.wave_speed[i] * n[i];
  fn[0] = 0.;
#pragma unroll
  for (int i = 0; i < ND; i++) fn[0] += P.wave_speed[i] * n[i] * u_av;
  fn[0] += 0.5 * P.lambda * fabs(norm_speed) * u_diff;
}

// common inviscid normal flux from both states (fluxes evaluated inside)
template <int ND, int NF>
__device__ __forceinline__ void riemann(const double *u_l, const double *u_r, const double *n, double *fn, const hf_phys &P)
{
  if (NF == 1)
  {
    lax_friedrich<ND>(u_l, u_r, n, fn, P);
    return;
  }
  double f_l[NF * ND], f_r[NF * ND];
  inv_flux<ND, NF>(u_l, f_l, P);
  inv_flux<ND, NF>(u_r, f_r, P);
  if (P.riemann_solve_type == 0) rusanov_flux<ND, NF>(u_l, u_r, f_l, f_r, n, fn, P);
  else if (P.riemann_solve_type == 2) roeM_flux<ND, NF>(u_l, u_r, f_l, f_r, n, fn, P);
  else hllc_flux<ND, NF>(u_l, u_r, f_l, f_r, n, fn, P);
}

// the "consistent switch" of the LDG flux: sign of beta from the interface normal, exact comparisons kept
template <int ND>
__device__ __forceinline__ double ldg_switched_beta(double ldg_beta, const double *n)
{
  if (ldg_beta != 0.)
  {
    if (n[0] < 0.) ldg_beta = -ldg_beta;
    else if (n[0] == 0.)
    {
      if ((n[0] + n[1]) < 0.) ldg_beta = -ldg_beta;
      else if ((n[0] + n[1]) == 0)
      {
        if (ND == 3)
        {
          if ((n[0] + n[ND - 1]) < 0.) ldg_beta = -ldg_beta;
        }
      }
    }
  }
  return ldg_beta;
}

template <int NF>
__device__ __forceinline__ void ldg_solution_int(const double *u_l, const double *u_r, double *u_c, double beta)
{
#pragma unroll
  // no FMA contraction here: delta = u_c - u_l is a difference of nearly equal numbers, and with |beta| = 1/2 the
  // reference's separately rounded product makes u_c equal one of the two states to the last bit
  for (int k = 0; k < NF; k++) u_c[k] = __dsub_rn(__dmul_rn(0.5, __dadd_rn(u_l[k], u_r[k])), __dmul_rn(beta, __dsub_rn(u_l[k], u_r[k])));
}

// flux_spec 0 (interior / partition): f_c = (1/2+beta) f_l + (1/2-beta) f_r ; flux_spec 1 (boundary): f_c = f_r
template <int ND, int NF>
__device__ __forceinline__ void ldg_flux(int flux_spec, const double *u_l, const double *u_r, const double *f_l, const double *f_r, const double *n, double *fn, double beta, double tau)
{
  double f_c[NF * ND];
  if (flux_spec == 0)
  {
#pragma unroll
    for (int q = 0; q < NF * ND; q++) f_c[q] = (0.5 + beta) * f_l[q] + (0.5 - beta) * f_r[q];
  }
  else
  {
#pragma unroll
    for (int q = 0; q < NF * ND; q++) f_c[q] = f_r[q];
  }
  normal_flux<ND, NF>(f_c, n, fn);
#pragma unroll
  for (int k = 0; k < NF; k++) fn[k] -= tau * (u_r[k] - u_l[k]);
}
