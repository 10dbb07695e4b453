// Fused residual kernels (included by hf_fused.cu inside its anonymous namespace).
//
// What bounds these kernels on B200 is the shared-memory pipe, not HBM and not the FP64 pipe: one 64-bit LDS/STS of a
// warp costs two wavefronts (16 doubles per clock and SM) against 64 FP64 FMA per clock and SM, measured in
// profiles/ (a thread-per-point formulation with one LDS per FMA ran the LSU pipe at 81 %; FP64 tensor-core tiles
// share the FP64 pipe and do not raise the FMA-per-operand ratio of a 5-wide line, profiles/microbench/dmma_b200.txt).
// So every 1-D operator is applied by a thread that owns a whole line of N solution points (N loads feed N
// derivative outputs, both face values and the correction terms), and passes that touch the same line are merged:
//
//   k_grad   L   per line: own face values (-> sf), LDG solution correction with the neighbour's values at both ends,
//                corrected reference-space derivative along the line (-> sg)             [opp_0, ldg_solution, opp_4, opp_5]
//            G   per direction d: all 15 gradient planes extrapolated along d-lines to the two faces normal to d
//                (-> gf, over the dead neighbour values), then one thread per flux point of those faces: metric
//                transform, viscous flux, dotted with the face's LEFT normal -> fv (global)          [opp_6, one-sided LDG]
//   k_resid  L   as above
//            PW  per solution point: physical gradient, inviscid + viscous flux, transformed flux, in place over sg
//            RM  per flux point: Riemann + LDG common normal flux, scaled, in place over the neighbour values (sx)
//            DX/DY/DZ  per line: own normal flux at both ends, correction, divergence along the line accumulated in
//                place; the z pass finishes with the RK update of its five points and their z-face values
//                                                                      [opp_1, opp_2, opp_3, AdvanceSolution, opp_0]
//            FO  per x- and y-line: face values of the updated solution -> fu[next]
// Bank conflicts: planes are padded to a stride = N (mod 16) doubles and y-line tasks are ordered (x, plane, z,
// element), which makes the addresses of consecutive lanes consecutive modulo 16; x-line tasks stride by N (odd),
// z-line tasks are contiguous.

template <int N, int E>
struct smem6
{
  static constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, PL = E * NU, FQ = E * NFP;
  static constexpr int PLS = PL + ((N - PL % 16) + 16) % 16; // plane stride = N (mod 16)
  double su[NF][PLS];      // solution at solution points
  double sx[NF][FQ];       // neighbour values at own flux points -> common normal flux (scaled) | gf of k_grad
  double sf[NF][FQ];       // own face values
  double sg[ND * NF][PLS]; // reference-space gradient -> transformed flux -> divergence (planes 0..4)
  double em[E][EM];
  unsigned long long bs[E][6];
  int finfo[E][6];
  int ge[E];
};

// ---- staging ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int N, int E, int NT, bool NEIGHBOURS, bool PREFETCH_FV, typename SM>
__device__ __forceinline__ void stage6(SM &S, const fused_args &A, int l0, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN;
  const int tid = threadIdx.x;
  if (tid < ne) S.ge[tid] = elem_id(A, l0 + tid);
  if (!A.elist)
  {
    // identity order: the ne elements of this CTA are contiguous in (upt, ele) for every field
    const double *src = A.u0 + (size_t)NU * l0;
    const size_t fs = (size_t)NU * A.n_eles;
    for (int i = tid; i < ne * NU; i += NT)
    {
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.su[k][i], src + i + k * fs);
    }
  }
  else
  {
    const size_t fs = (size_t)NU * A.n_eles;
    for (int i = tid; i < ne * NU; i += NT)
    {
      const int e = i / NU, p = i - e * NU;
      const double *src = A.u0 + p + (size_t)NU * A.elist[l0 + e];
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.su[k][i], src + k * fs);
    }
  }
  if constexpr (NEIGHBOURS)
  {
    for (int q = tid; q < ne * NFP; q += NT)
    {
      const int e = q / NFP, r = q - e * NFP;
      const double *nb = A.fu_cur + A.nidx[(size_t)elem_id(A, l0 + e) * NFP + r];
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.sx[k][q], nb + k * NN);
    }
  }
  for (int i = tid; i < ne * EM; i += NT)
  {
    const int e = i / EM;
    cp_async8(&S.em[0][0] + i, A.em + (size_t)elem_id(A, l0 + e) * EM + (i - e * EM));
  }
  if (tid < ne * 6)
  {
    const int e = tid / 6, f = tid - e * 6;
    const size_t gf = (size_t)elem_id(A, l0 + e) * 6 + f;
    cp_async8(&S.bs[0][0] + tid, A.bmask + gf);
    S.finfo[0][tid] = A.finfo[gf];
  }
  // Software prefetch into L2 for the CTA that will run in this slot next (A.pf_dist CTAs further on): its solution
  // rows, neighbour face blocks and (k_resid) viscous-flux blocks, so that its staging copies hit L2 instead of HBM.
  if constexpr (NEIGHBOURS)
  {
    const int p0 = l0 + A.pf_dist * E;
    if (A.pf_dist > 0 && p0 < A.hi)
    {
      const int pn = min(E, A.hi - p0);
      constexpr int ROW_LINES = (E * NU * 8 + 127) / 128 + 1;
      if (!A.elist)
      {
        for (int i = tid; i < NF * ROW_LINES; i += NT)
        {
          const int k = i / ROW_LINES, b = i - k * ROW_LINES;
          if (b * 128 < pn * NU * 8 + 120) prefetch_l2((const char *)(A.u0 + (size_t)NU * (p0 + (size_t)A.n_eles * k)) + b * 128);
        }
      }
      constexpr int FU_LINES = (NF * NN * 8 + 127) / 128 + 1, FV_LINES = (4 * NN * 8 + 127) / 128 + 1;
      for (int i = tid; i < pn * 6 * FU_LINES; i += NT)
      {
        const int ef = i / FU_LINES, b = i - ef * FU_LINES;
        const int blk = A.nbr[(size_t)elem_id(A, p0 + ef / 6) * 6 + ef % 6];
        prefetch_l2((const char *)(A.fu_cur + (size_t)blk * (NF * NN)) + b * 128);
      }
      if constexpr (PREFETCH_FV)
      {
        for (int i = tid; i < pn * 6 * FV_LINES; i += NT)
        {
          const int ef = i / FV_LINES, b = i - ef * FV_LINES;
          const size_t gf = (size_t)elem_id(A, p0 + ef / 6) * 6 + ef % 6;
          prefetch_l2((const char *)(A.fv + gf * (4 * NN)) + b * 128);
          prefetch_l2((const char *)(A.fv + (size_t)A.nbr[gf] * (4 * NN)) + b * 128);
        }
      }
    }
  }
  cp_async_commit();
}

// ---- line tasks --------------------------------------------------------------------------------------------------------
// A line task is (line, field): the thread decodes it once and then sweeps the E element slots of the CTA (and, in
// the G pass, the three gradient directions) with constant geometry.  For P = 4 and 125 threads every thread owns
// exactly one (line, field) per direction.
template <int N, int DIR>
struct line_dir
{
  static constexpr int stride = DIR == 0 ? 1 : (DIR == 1 ? N : N * N);
  static constexpr int fminus = DIR == 0 ? 4 : (DIR == 1 ? 1 : 0);
  static constexpr int fplus = DIR == 0 ? 2 : (DIR == 1 ? 3 : 5);
};
// task t in [0, NN*NF) of direction DIR -> field k, first solution point of the line inside an element, face-local
// flux points jm / jp at its minus / plus end (reference src/eles_hexas.cpp:224-282)
template <int N, int DIR>
__device__ __forceinline__ void line_task(int t, int &k, int &base, int &jm, int &jp)
{
  constexpr int P = N - 1, NN = N * N;
  int x, y;
  if (DIR == 1)
  {
    // (x, field, z): with plane strides = N (mod 16) consecutive lanes touch consecutive banks
    x = t % N;
    const int t1 = t / N;
    k = t1 % NF;
    y = t1 / NF;
  }
  else
  {
    const int l = t % NN;
    k = t / NN;
    x = l % N;
    y = l / N;
  }
  if (DIR == 0) { base = N * x + NN * y; jm = (P - x) + N * y; jp = x + N * y; }
  else if (DIR == 1) { base = x + NN * y; jm = x + N * y; jp = (P - x) + N * y; }
  else { base = x + N * y; jm = (P - x) + N * y; jp = x + N * y; }
}

// weight of the element's own value in the LDG common solution / flux at a flux point (0.5 + beta or 0.5 - beta, see the
// bmask comment in hf_fused_prepare; inters::ldg_solution / ldg_flux, reference src/inters.cpp:561-646)
__device__ __forceinline__ double ldg_own_weight(unsigned long long mask, int j, double beta)
{
  return ((mask >> j) & 1ull) ? 0.5 - beta : 0.5 + beta;
}

// L pass of one direction: own face values -> sf; with GRAD also the corrected reference-space derivative -> sg
template <int N, int E, int NT, int DIR, bool GRAD, typename SM>
__device__ __forceinline__ void pass_L(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const double *x = S.su[k] + e * NU + base;
      double v[N];
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = x[j * stride];
      double um = A.tL[0][0] * v[0], up = A.tL[1][0] * v[0];
#pragma unroll
      for (int j = 1; j < N; j++) { um += A.tL[0][j] * v[j]; up += A.tL[1][j] * v[j]; }
      const int fmq = e * NFP + FM * NN + jm, fpq = e * NFP + FP * NN + jp;
      S.sf[k][fmq] = um;
      S.sf[k][fpq] = up;
      if (GRAD)
      {
        // LDG solution correction: u_c weighs the two sides opposite to f_c, so u_c - u_own = w_own * (u_nbr - u_own)
        const double dm = ldg_own_weight(S.bs[e][FM], jm, A.P.ldg_beta) * (S.sx[k][fmq] - um);
        const double dp = ldg_own_weight(S.bs[e][FP], jp, A.P.ldg_beta) * (S.sx[k][fpq] - up);
        double *o = S.sg[DIR * NF + k] + e * NU + base;
#pragma unroll
        for (int i = 0; i < N; i++)
        {
          double acc = A.tD[i * N] * v[0];
#pragma unroll
          for (int j = 1; j < N; j++) acc += A.tD[i * N + j] * v[j];
          acc += A.tc5[FP * N + i] * dp;
          acc += A.tc5[FM * N + i] * dm;
          o[i * stride] = acc;
        }
      }
    }
  }
}

// ---- kernel 1: one-sided viscous normal flux at the faces -------------------------------------------------------------
// G pass of direction DIR: every gradient plane extrapolated along its DIR-lines to the two faces normal to DIR
template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_G(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, stride = line_dir<N, DIR>::stride;
  double *gf = &S.sx[0][0]; // [direction][field][e][side][fpt]
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int c = 0; c < ND; c++)
#pragma unroll
      for (int e = 0; e < E; e++)
      {
        if (e >= ne) break;
        const double *x = S.sg[c * NF + k] + e * NU + base;
        double gm = 0., gp = 0.;
#pragma unroll
        for (int j = 0; j < N; j++)
        {
          const double v = x[j * stride];
          gm += A.tL[0][j] * v;
          gp += A.tL[1][j] * v;
        }
        double *o = gf + (c * NF + k) * (2 * NN * E) + e * (2 * NN);
        o[jm] = gm;
        o[NN + jp] = gp;
      }
  }
}
// flux points of the two faces normal to DIR: metric transform (as eles::correct_gradient does at flux points),
// viscous flux, dotted with the face's left normal
template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_GF(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NFP = 6 * NN;
  const double *gf = &S.sx[0][0];
  for (int q = threadIdx.x; q < ne * 2 * NN; q += NT)
  {
    const int e = q / (2 * NN), r = q - e * (2 * NN), side = r / NN, j = r - side * NN;
    const int f = side ? line_dir<N, DIR>::fplus : line_dir<N, DIR>::fminus;
    const double *J = S.em[e];
    const double idj = J[9];
    double u[NF], g[NF * ND], fv[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++)
    {
      u[k] = S.sf[k][e * NFP + f * NN + j];
      const double g0 = gf[(0 * NF + k) * (2 * NN * E) + q] * idj, g1 = gf[(1 * NF + k) * (2 * NN * E) + q] * idj,
                   g2 = gf[(2 * NF + k) * (2 * NN * E) + q] * idj;
      g[k] = g0 * J[0] + g1 * J[1] + g2 * J[2];
      g[k + 5] = g0 * J[3] + g1 * J[4] + g2 * J[5];
      g[k + 10] = g0 * J[6] + g1 * J[7] + g2 * J[8];
    }
    vis_flux_fast(u, g, fv, A.P);
    const double *n = &S.em[e][10 + 4 * f + 1];
    const double n0 = n[0], n1 = n[1], n2 = n[2];
    double *out = A.fv + ((size_t)S.ge[e] * 6 + f) * (4 * NN) + j;
#pragma unroll
    for (int k = 1; k < NF; k++) out[(k - 1) * NN] = fv[k] * n0 + fv[k + 5] * n1 + fv[k + 10] * n2;
  }
}

template <int N, int E, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_grad6(const __grid_constant__ fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef smem6<N, E> SM;
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int l0 = A.lo + blockIdx.x * E;
  const int ne = min(E, A.hi - l0);
  stage6<N, E, NT, true, false>(S, A, l0, ne);
  cp_async_wait_all();
  __syncthreads();
  pass_L<N, E, NT, 0, true>(S, A, ne);
  pass_L<N, E, NT, 1, true>(S, A, ne);
  pass_L<N, E, NT, 2, true>(S, A, ne);
  __syncthreads();
  pass_G<N, E, NT, 0>(S, A, ne);
  __syncthreads();
  pass_GF<N, E, NT, 0>(S, A, ne);
  __syncthreads();
  pass_G<N, E, NT, 1>(S, A, ne);
  __syncthreads();
  pass_GF<N, E, NT, 1>(S, A, ne);
  __syncthreads();
  pass_G<N, E, NT, 2>(S, A, ne);
  __syncthreads();
  pass_GF<N, E, NT, 2>(S, A, ne);
}

// ---- kernel 2: residual + RK update + next face values ----------------------------------------------------------------
// divergence pass of one direction; DIR 2 finishes with the RK update and the z-face values of the updated solution
template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_D(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const double *x = S.sg[DIR * NF + k] + e * NU + base;
      double v[N];
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = x[j * stride];
      double nm = A.tL[0][0] * v[0], np = A.tL[1][0] * v[0];
#pragma unroll
      for (int j = 1; j < N; j++) { nm += A.tL[0][j] * v[j]; np += A.tL[1][j] * v[j]; }
      // common minus own normal flux: norm_tdisf = -(Lm . tdisf) on a minus face, +(Lp . tdisf) on a plus face
      const double dm = S.sx[k][e * NFP + FM * NN + jm] + nm;
      const double dp = S.sx[k][e * NFP + FP * NN + jp] - np;
      double *o = S.sg[k] + e * NU + base;
      double out[N];
#pragma unroll
      for (int i = 0; i < N; i++)
      {
        double acc = A.tD[i * N] * v[0];
#pragma unroll
        for (int j = 1; j < N; j++) acc += A.tD[i * N + j] * v[j];
        acc += A.tc3[FP * N + i] * dp;
        acc += A.tc3[FM * N + i] * dm;
        out[i] = acc;
      }
      if (DIR == 0)
      {
#pragma unroll
        for (int i = 0; i < N; i++) o[i * stride] = out[i];
      }
      else if (DIR == 1)
      {
#pragma unroll
        for (int i = 0; i < N; i++) o[i * stride] += out[i];
      }
      else
      {
        // div_tconf complete for the N points of this z-line: RK update (eles::AdvanceSolution)
        const int ge = S.ge[e];
        const double inv_detjac = S.em[e][9];
        const double dtl = A.dt_local ? A.dt_local[ge] : A.rk.dt;
        const double dt_fac = A.dt_local ? dtl / A.rk.fac : A.rk.dt_fac; // (dt / fac) * r, the reference's evaluation order (src/eles.cpp:1141, 1191)
        const size_t gi0 = (size_t)base + (size_t)NU * ge + (size_t)k * NU * A.n_eles;
        double *us = S.su[k] + e * NU + base;
        double unew[N];
#pragma unroll
        for (int i = 0; i < N; i++)
        {
          const double acc = o[i * stride] + out[i];
          const size_t gi = gi0 + i * stride;
          if (A.keep_residual) A.div[gi] = acc;
          if (acc != acc) *A.nan_flag = 1 + ge;
          double u = us[i * stride];
          if (A.do_update)
          {
            const double rr = acc * inv_detjac;
            if (A.rk.copy_u1) A.u1[gi] = u;
            if (A.rk.mode == 0)
              u -= dt_fac * rr;
            else if (A.rk.mode == 1)
              u = A.rk.c1 * u + A.rk.c2 * A.u1[gi] + dt_fac * (-rr);
            else
            {
              const double dlt = A.rk.c1 * A.u1[gi] + dtl * (-rr);
              A.u1[gi] = dlt;
              u += A.rk.c2 * dlt;
            }
            A.u0_out[gi] = u;
            us[i * stride] = u;
          }
          unew[i] = u;
        }
        if (A.do_update)
        {
          double um = A.tL[0][0] * unew[0], up = A.tL[1][0] * unew[0];
#pragma unroll
          for (int j = 1; j < N; j++) { um += A.tL[0][j] * unew[j]; up += A.tL[1][j] * unew[j]; }
          double *blk = A.fu_next + (size_t)ge * 6 * (NF * NN) + k * NN;
          blk[FM * (NF * NN) + jm] = um;
          blk[FP * (NF * NN) + jp] = up;
        }
      }
    }
  }
}

// face values of the solution in su along DIR-lines -> the element's own face blocks in global memory
template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_FO(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const double *x = S.su[k] + e * NU + base;
      double um = 0., up = 0.;
#pragma unroll
      for (int j = 0; j < N; j++)
      {
        const double v = x[j * stride];
        um += A.tL[0][j] * v;
        up += A.tL[1][j] * v;
      }
      double *blk = A.fu_next + (size_t)S.ge[e] * 6 * (NF * NN) + k * NN;
      blk[FM * (NF * NN) + jm] = um;
      blk[FP * (NF * NN) + jp] = up;
    }
  }
}

template <int N, int E, int NT, int MINB, bool VISC, bool GOUT>
__global__ void __launch_bounds__(NT, MINB) k_resid6(const __grid_constant__ fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef smem6<N, E> SM;
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN;
  const int tid = threadIdx.x;
  const int l0 = A.lo + blockIdx.x * E;
  const int ne = min(E, A.hi - l0);
  stage6<N, E, NT, true, VISC>(S, A, l0, ne);
  cp_async_wait_all();
  __syncthreads();
  pass_L<N, E, NT, 0, VISC>(S, A, ne);
  pass_L<N, E, NT, 1, VISC>(S, A, ne);
  pass_L<N, E, NT, 2, VISC>(S, A, ne);
  __syncthreads();
  // PW: transformed total flux at the solution points, in place over the gradient
  for (int q = tid; q < ne * NU; q += NT)
  {
    const int e = q / NU;
    const double *J = S.em[e];
    double u[NF], f[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++) u[k] = S.su[k][q];
    inv_flux_fast(u, f, A.P.gamma - 1.0);
    if constexpr (VISC)
    {
      double g[NF * ND], fv[NF * ND];
      const double idj = J[9];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        const double g0 = S.sg[k][q] * idj, g1 = S.sg[NF + k][q] * idj, g2 = S.sg[2 * NF + k][q] * idj;
        g[k] = g0 * J[0] + g1 * J[1] + g2 * J[2];
        g[k + 5] = g0 * J[3] + g1 * J[4] + g2 * J[5];
        g[k + 10] = g0 * J[6] + g1 * J[7] + g2 * J[8];
      }
      vis_flux_fast(u, g, fv, A.P);
      if constexpr (GOUT) // integral diagnostics: grad_disu_upts of this residual evaluation (a separate instantiation: the stores cost registers)
      {
        const size_t gi = (size_t)(q - e * NU) + (size_t)NU * S.ge[e], fs = (size_t)NU * A.n_eles;
#pragma unroll
        for (int d = 0; d < ND; d++)
#pragma unroll
          for (int k = 0; k < NF; k++) A.grad_out[gi + fs * (k + NF * d)] = g[k + NF * d];
      }
#pragma unroll
      for (int d = 0; d < ND; d++)
#pragma unroll
        for (int k = 1; k < NF; k++) f[k + NF * d] += fv[k + NF * d];
    }
    // tdisf(k,l) = sum_m JGinv(l,m) f(k,m)
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
      for (int l = 0; l < ND; l++) S.sg[l * NF + k][q] = J[l] * f[k] + J[l + 3] * f[k + 5] + J[l + 6] * f[k + 10];
  }
  // RM: common normal flux (Riemann + LDG) at every own flux point, in the element's own orientation and scaled by its
  // tdA, in place over the neighbour values in sx (no barrier needed against PW: disjoint arrays)
  for (int q = tid; q < ne * NFP; q += NT)
  {
    const int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    const int info = S.finfo[e][f];
    const bool is_right = (info & 4) != 0;
    const double *geo = &S.em[e][10 + 4 * f];
    const double tdA = geo[0];
    const double n[3] = {geo[1], geo[2], geo[3]};
    double fvo[4], fvn[4];
    if constexpr (VISC)
    {
      const int ge = S.ge[e];
      const int blk = A.nbr[(size_t)ge * 6 + f];
      const double *po = A.fv + ((size_t)ge * 6 + f) * (4 * NN) + j;
      const double *pn = A.fv + (size_t)blk * (4 * NN) + (A.nidx[(size_t)ge * NFP + r] - blk * (NF * NN));
#pragma unroll
      for (int k = 0; k < 4; k++) { fvo[k] = po[k * NN]; fvn[k] = pn[k * NN]; }
    }
    double uo[NF], un[NF], fn[NF];
#pragma unroll
    for (int k = 0; k < NF; k++) { un[k] = S.sx[k][q]; uo[k] = S.sf[k][q]; }
    {
      // one solver call on (left, right) selected per thread: no divergent duplicate of the solver body
      double ul[NF], ur[NF];
#pragma unroll
      for (int k = 0; k < NF; k++) { ul[k] = is_right ? un[k] : uo[k]; ur[k] = is_right ? uo[k] : un[k]; }
      double nr[3] = {n[0], n[1], n[2]};
      if (A.nlf) // RoeM: the reference's normal of exactly this flux point (hf_fused_prepare)
      {
        const double *q3 = A.nlf + ((size_t)(S.ge[e] * 6 + f) * NN + j) * 3;
        nr[0] = q3[0]; nr[1] = q3[1]; nr[2] = q3[2];
      }
      riemann_fast(ul, ur, nr, fn, A.P);
    }
    if constexpr (VISC)
    {
      // LDG: f_c = w_own f_own + (1 - w_own) f_nbr, minus tau (u_r - u_l)
      const double wo = ldg_own_weight(S.bs[e][f], j, A.P.ldg_beta), wn = 1.0 - wo, tau = A.P.ldg_tau;
      const double flip = (info & 8) ? -wn : wn; // a partition neighbour used its own (opposite) normal
      const double ts = is_right ? -tau : tau;
      fn[0] -= ts * (un[0] - uo[0]);
#pragma unroll
      for (int k = 1; k < NF; k++) fn[k] += (wo * fvo[k - 1] + flip * fvn[k - 1]) - ts * (un[k] - uo[k]);
    }
    const double s_side = is_right ? -tdA : tdA;
#pragma unroll
    for (int k = 0; k < NF; k++) S.sx[k][q] = fn[k] * s_side;
  }
  __syncthreads();
  pass_D<N, E, NT, 0>(S, A, ne);
  __syncthreads();
  pass_D<N, E, NT, 1>(S, A, ne);
  __syncthreads();
  pass_D<N, E, NT, 2>(S, A, ne);
  if (!A.do_update) return;
  __syncthreads();
  pass_FO<N, E, NT, 0>(S, A, ne);
  pass_FO<N, E, NT, 1>(S, A, ne);
}

// face values of the current solution (first stage, or after an upload)
template <int N, int E, int NT>
__global__ void __launch_bounds__(NT) k_face_values6(const __grid_constant__ fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef smem6<N, E> SM;
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  const int l0 = A.lo + blockIdx.x * E;
  const int ne = min(E, A.hi - l0);
  stage6<N, E, NT, false, false>(S, A, l0, ne);
  cp_async_wait_all();
  __syncthreads();
  pass_FO<N, E, NT, 0>(S, A, ne);
  pass_FO<N, E, NT, 1>(S, A, ne);
  pass_FO<N, E, NT, 2>(S, A, ne);
}

// =====================================================================================================================
// Generation 7: one-sided LDG (|ldg_beta| = 0.5, the reference's default and the TGV configuration).
//
// With beta = +-0.5 the LDG weights are exactly 1 and 0: at every flux-point pair ONE side (the "owner", picked by the
// reference's sign switch on the face normal, src/inters.cpp:566-581, which also fires on rounding-level normal
// components, so ownership is kept per flux point) supplies the viscous flux, and the common solution is the other
// side's value.  Consequences used here:
//   * the owner has everything the common flux needs (own u, neighbour u, own gradient): it evaluates Riemann + LDG
//     once per pair in k_grad7 and stores the complete common normal flux fc (5 values per flux point, along the LEFT
//     normal); the other side only reads it.  k_resid7 has no Riemann / viscous interface work left;
//   * the LDG solution correction of an element only involves its owned flux points, so only those neighbour values
//     are staged, only those face gradients are evaluated, and an element only publishes face values (fu) at the flux
//     points it does not own.
// Results are those of generation 6 (weights 1 and 0 are exact) with a third less traffic and fewer instructions.
template <int N, int E>
struct smem7
{
  static constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, PL = E * NU, FQ = E * NFP;
  static constexpr int PLS = PL + ((N - PL % 16) + 16) % 16;
  double su[NF][PLS];      // solution at solution points
  double sx[NF][FQ];       // neighbour values at the OWNED flux points (the others are never loaded)
  double sf[NF][FQ];       // k_grad7: own face values; k_resid7: common normal flux fc at every own flux point
  double sg[ND * NF][PLS]; // reference-space gradient -> transformed flux -> divergence (planes 0..4)
  double em[E][EM];
  double scl[E][6];        // k_resid7: +-tdA that turns fc into the element's own norm_tconf
  unsigned long long own[E][6]; // bit j: this element owns flux point j of the face
  int finfo[E][6];
  int ge[E];
  int n_owned;
  unsigned short olist[FQ]; // k_grad7: the owned flux points of the CTA, compacted
};

__device__ __forceinline__ bool own_bit(unsigned long long m, int j) { return (m >> j) & 1ull; }

template <int N, int E, int NT, bool RESID, typename SM>
__device__ __forceinline__ void stage7(SM &S, const fused_args &A, int l0, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN;
  const int tid = threadIdx.x;
  if (tid < ne) S.ge[tid] = elem_id(A, l0 + tid);
  const size_t fs = (size_t)NU * A.n_eles;
  if (!A.elist)
  {
    const double *src = A.u0 + (size_t)NU * l0;
    for (int i = tid; i < ne * NU; i += NT)
    {
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.su[k][i], src + i + k * fs);
    }
  }
  else
  {
    for (int i = tid; i < ne * NU; i += NT)
    {
      const int e = i / NU, p = i - e * NU;
      const double *src = A.u0 + p + (size_t)NU * A.elist[l0 + e];
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.su[k][i], src + k * fs);
    }
  }
  for (int q = tid; q < ne * NFP; q += NT)
  {
    const int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    const int ge = elem_id(A, l0 + e);
    const bool own = own_bit(A.bmask[(size_t)ge * 6 + f] ^ A.own_xor, j);
    int ni = 0;
    if (own || RESID) ni = A.nidx[(size_t)ge * NFP + r];
    if (own)
    {
      const double *nb = A.fu_cur + ni;
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.sx[k][q], nb + k * NN);
    }
    if constexpr (RESID)
    {
      const double *src = A.fv + (own ? (size_t)ge * (NF * NFP) + (size_t)f * (NF * NN) + j : (size_t)ni);
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async8(&S.sf[k][q], src + k * NN);
    }
  }
  for (int i = tid; i < ne * EM; i += NT)
  {
    const int e = i / EM;
    cp_async8(&S.em[0][0] + i, A.em + (size_t)elem_id(A, l0 + e) * EM + (i - e * EM));
  }
  if (tid < ne * 6)
  {
    const int e = tid / 6, f = tid - e * 6;
    const size_t gf = (size_t)elem_id(A, l0 + e) * 6 + f;
    S.own[0][tid] = A.bmask[gf] ^ A.own_xor;
    S.finfo[0][tid] = A.finfo[gf];
  }
  cp_async_commit();
}

// L pass, one-sided: own face values at both ends of the line, LDG correction with the neighbour's value where this
// element owns the flux point (weight 1, else 0), corrected reference-space derivative along the line
template <int N, int E, int NT, int DIR, bool STORE_FACE, typename SM>
__device__ __forceinline__ void pass_L7(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const double *x = S.su[k] + e * NU + base;
      double v[N];
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = x[j * stride];
      double um = A.tL[0][0] * v[0], up = A.tL[1][0] * v[0];
#pragma unroll
      for (int j = 1; j < N; j++) { um += A.tL[0][j] * v[j]; up += A.tL[1][j] * v[j]; }
      const int fmq = e * NFP + FM * NN + jm, fpq = e * NFP + FP * NN + jp;
      if constexpr (STORE_FACE)
      {
        S.sf[k][fmq] = um;
        S.sf[k][fpq] = up;
      }
      // unconditional loads (slots of flux points this element does not own hold stale data, discarded by the select)
      const double xm = S.sx[k][fmq], xp = S.sx[k][fpq];
      const double dm = own_bit(S.own[e][FM], jm) ? xm - um : 0.;
      const double dp = own_bit(S.own[e][FP], jp) ? xp - up : 0.;
      double *o = S.sg[DIR * NF + k] + e * NU + base;
#pragma unroll
      for (int i = 0; i < N; i++)
      {
        double acc = A.tD[i * N] * v[0];
#pragma unroll
        for (int j = 1; j < N; j++) acc += A.tD[i * N + j] * v[j];
        acc += A.tc5[FP * N + i] * dp;
        acc += A.tc5[FM * N + i] * dm;
        o[i * stride] = acc;
      }
    }
  }
}

// owned flux points: gradient extrapolated along the line behind the point, physical gradient, viscous flux of the own
// side, Riemann flux, complete common normal flux -> fc
template <int N, int E, int NT, typename SM>
__device__ __forceinline__ void pass_GF7(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN;
  const int n_owned = S.n_owned;
  for (int i = threadIdx.x; i < n_owned; i += NT)
  {
    const int q = S.olist[i];
    const int e = q / NFP, r = q - e * NFP, f = r / NN, j = r - f * NN;
    const int info = S.finfo[e][f];
    const bool is_right = (info & 4) != 0;
    const double *J = S.em[e];
    const double idj = J[9];
    int base, stride;
    line_of_fpt<N>(f, j, base, stride);
    const int side = face_sgn(f) > 0 ? 1 : 0;
    double w[N];
#pragma unroll
    for (int m = 0; m < N; m++) w[m] = A.tL[side][m];
    double uo[NF], un[NF], fn[NF], vn[NF];
    {
      double g[NF * ND], fv[NF * ND];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        uo[k] = S.sf[k][q];
        double gr[ND];
#pragma unroll
        for (int c = 0; c < ND; c++)
        {
          const double *x = S.sg[c * NF + k] + e * NU + base;
          double a = 0.;
#pragma unroll
          for (int m = 0; m < N; m++) a += w[m] * x[m * stride];
          gr[c] = a * idj;
        }
        g[k] = gr[0] * J[0] + gr[1] * J[1] + gr[2] * J[2];
        g[k + 5] = gr[0] * J[3] + gr[1] * J[4] + gr[2] * J[5];
        g[k + 10] = gr[0] * J[6] + gr[1] * J[7] + gr[2] * J[8];
      }
      vis_flux_fast(uo, g, fv, A.P);
      const double *n = &S.em[e][10 + 4 * f + 1];
      const double n0 = n[0], n1 = n[1], n2 = n[2];
      vn[0] = 0.;
#pragma unroll
      for (int k = 1; k < NF; k++) vn[k] = fv[k] * n0 + fv[k + 5] * n1 + fv[k + 10] * n2;
    }
    {
      const double *n = &S.em[e][10 + 4 * f + 1];
      double nl[3] = {n[0], n[1], n[2]};
      if (A.nlf) // RoeM: the reference's normal of exactly this flux point (hf_fused_prepare)
      {
        const double *q3 = A.nlf + ((size_t)(S.ge[e] * 6 + f) * NN + j) * 3;
        nl[0] = q3[0]; nl[1] = q3[1]; nl[2] = q3[2];
      }
      double ul[NF], ur[NF];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        un[k] = S.sx[k][q];
        ul[k] = is_right ? un[k] : uo[k];
        ur[k] = is_right ? uo[k] : un[k];
      }
      riemann_fast(ul, ur, nl, fn, A.P);
    }
    const double ts = is_right ? -A.P.ldg_tau : A.P.ldg_tau;
    fn[0] -= ts * (un[0] - uo[0]);
#pragma unroll
    for (int k = 1; k < NF; k++) fn[k] += vn[k] - ts * (un[k] - uo[k]);
    double *out = A.fv + ((size_t)S.ge[e] * 6 + f) * (NF * NN) + j;
#pragma unroll
    for (int k = 0; k < NF; k++) out[k * NN] = fn[k];
  }
}

template <int N, int E, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_grad7(const __grid_constant__ fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef smem7<N, E> SM;
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  constexpr int NN = N * N, NFP = 6 * NN;
  const int l0 = A.lo + blockIdx.x * E;
  const int ne = min(E, A.hi - l0);
  if (threadIdx.x == 0) S.n_owned = 0;
  stage7<N, E, NT, false>(S, A, l0, ne);
  cp_async_wait_all();
  __syncthreads();
  for (int q = threadIdx.x; q < ne * NFP; q += NT)
  {
    const int e = q / NFP, r = q - e * NFP, f = r / NN;
    if (own_bit(S.own[e][f], r - f * NN)) S.olist[atomicAdd(&S.n_owned, 1)] = (unsigned short)q;
  }
  pass_L7<N, E, NT, 0, true>(S, A, ne);
  pass_L7<N, E, NT, 1, true>(S, A, ne);
  pass_L7<N, E, NT, 2, true>(S, A, ne);
  __syncthreads();
  pass_GF7<N, E, NT>(S, A, ne);
}

// divergence pass, one-sided variant: the common flux comes from sf (scaled per face), the z pass finishes with the RK
// update and publishes the z-face values at the flux points this element does not own
template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_D7(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, NFP = 6 * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const double *x = S.sg[DIR * NF + k] + e * NU + base;
      double v[N];
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = x[j * stride];
      double nm = A.tL[0][0] * v[0], np = A.tL[1][0] * v[0];
#pragma unroll
      for (int j = 1; j < N; j++) { nm += A.tL[0][j] * v[j]; np += A.tL[1][j] * v[j]; }
      const bool own_m = own_bit(S.own[e][FM], jm), own_p = own_bit(S.own[e][FP], jp);
      // a partition neighbour evaluated fc along its own (opposite) normal
      double sm = S.scl[e][FM], sp = S.scl[e][FP];
      if ((S.finfo[e][FM] & 8) && !own_m) sm = -sm;
      if ((S.finfo[e][FP] & 8) && !own_p) sp = -sp;
      const double dm = S.sf[k][e * NFP + FM * NN + jm] * sm + nm;
      const double dp = S.sf[k][e * NFP + FP * NN + jp] * sp - np;
      double *o = S.sg[k] + e * NU + base;
      double out[N];
#pragma unroll
      for (int i = 0; i < N; i++)
      {
        double acc = A.tD[i * N] * v[0];
#pragma unroll
        for (int j = 1; j < N; j++) acc += A.tD[i * N + j] * v[j];
        acc += A.tc3[FP * N + i] * dp;
        acc += A.tc3[FM * N + i] * dm;
        out[i] = acc;
      }
      if (DIR == 0)
      {
#pragma unroll
        for (int i = 0; i < N; i++) o[i * stride] = out[i];
      }
      else if (DIR == 1)
      {
#pragma unroll
        for (int i = 0; i < N; i++) o[i * stride] += out[i];
      }
      else
      {
        const int ge = S.ge[e];
        const double inv_detjac = S.em[e][9];
        const double dtl = A.dt_local ? A.dt_local[ge] : A.rk.dt;
        const double dt_fac = A.dt_local ? dtl / A.rk.fac : A.rk.dt_fac;
        const size_t gi0 = (size_t)base + (size_t)NU * ge + (size_t)k * NU * A.n_eles;
        double *us = S.su[k] + e * NU + base;
        double unew[N];
#pragma unroll
        for (int i = 0; i < N; i++)
        {
          const double acc = o[i * stride] + out[i];
          const size_t gi = gi0 + i * stride;
          if (A.keep_residual) A.div[gi] = acc;
          if (acc != acc) *A.nan_flag = 1 + ge;
          double u = us[i * stride];
          if (A.do_update)
          {
            const double rr = acc * inv_detjac;
            if (A.rk.copy_u1) A.u1[gi] = u;
            if (A.rk.mode == 0)
              u -= dt_fac * rr;
            else if (A.rk.mode == 1)
              u = A.rk.c1 * u + A.rk.c2 * A.u1[gi] + dt_fac * (-rr);
            else
            {
              const double dlt = A.rk.c1 * A.u1[gi] + dtl * (-rr);
              A.u1[gi] = dlt;
              u += A.rk.c2 * dlt;
            }
            A.u0_out[gi] = u;
            us[i * stride] = u;
          }
          unew[i] = u;
        }
        if (A.do_update)
        {
          // the owner of a flux-point pair is the only reader of the other side's value there
          double *blk = A.fu_next + (size_t)ge * 6 * (NF * NN) + k * NN;
          if (!own_m)
          {
            double um = A.tL[0][0] * unew[0];
#pragma unroll
            for (int j = 1; j < N; j++) um += A.tL[0][j] * unew[j];
            blk[FM * (NF * NN) + jm] = um;
          }
          if (!own_p)
          {
            double up = A.tL[1][0] * unew[0];
#pragma unroll
            for (int j = 1; j < N; j++) up += A.tL[1][j] * unew[j];
            blk[FP * (NF * NN) + jp] = up;
          }
        }
      }
    }
  }
}

template <int N, int E, int NT, int DIR, typename SM>
__device__ __forceinline__ void pass_FO7(SM &S, const fused_args &A, int ne)
{
  constexpr int NN = N * N, NU = N * NN, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  for (int t = threadIdx.x; t < NF * NN; t += NT)
  {
    int k, base, jm, jp;
    line_task<N, DIR>(t, k, base, jm, jp);
#pragma unroll
    for (int e = 0; e < E; e++)
    {
      if (e >= ne) break;
      const bool pub_m = !own_bit(S.own[e][FM], jm), pub_p = !own_bit(S.own[e][FP], jp);
      const double *x = S.su[k] + e * NU + base;
      double v[N];
#pragma unroll
      for (int j = 0; j < N; j++) v[j] = x[j * stride];
      double *blk = A.fu_next + (size_t)S.ge[e] * 6 * (NF * NN) + k * NN;
      if (pub_m)
      {
        double um = 0.;
#pragma unroll
        for (int j = 0; j < N; j++) um += A.tL[0][j] * v[j];
        blk[FM * (NF * NN) + jm] = um;
      }
      if (pub_p)
      {
        double up = 0.;
#pragma unroll
        for (int j = 0; j < N; j++) up += A.tL[1][j] * v[j];
        blk[FP * (NF * NN) + jp] = up;
      }
    }
  }
}

template <int N, int E, int NT, int MINB, bool GOUT>
__global__ void __launch_bounds__(NT, MINB) k_resid7(const __grid_constant__ fused_args A)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef smem7<N, E> SM;
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  constexpr int NN = N * N, NU = N * NN;
  const int tid = threadIdx.x;
  const int l0 = A.lo + blockIdx.x * E;
  const int ne = min(E, A.hi - l0);
  stage7<N, E, NT, true>(S, A, l0, ne);
  cp_async_wait_all();
  __syncthreads();
  if (tid < ne * 6)
  {
    // norm_tconf of this element = fc * scl: minus on the right side of an interior face
    const int e = tid / 6, f = tid - e * 6;
    const double tdA = S.em[e][10 + 4 * f];
    S.scl[e][f] = (S.finfo[e][f] & 4) ? -tdA : tdA;
  }
  pass_L7<N, E, NT, 0, false>(S, A, ne);
  pass_L7<N, E, NT, 1, false>(S, A, ne);
  pass_L7<N, E, NT, 2, false>(S, A, ne);
  __syncthreads();
  for (int q = tid; q < ne * NU; q += NT)
  {
    const int e = q / NU;
    const double *J = S.em[e];
    double u[NF], f[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++) u[k] = S.su[k][q];
    inv_flux_fast(u, f, A.P.gamma - 1.0);
    {
      double g[NF * ND], fv[NF * ND];
      const double idj = J[9];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        const double g0 = S.sg[k][q] * idj, g1 = S.sg[NF + k][q] * idj, g2 = S.sg[2 * NF + k][q] * idj;
        g[k] = g0 * J[0] + g1 * J[1] + g2 * J[2];
        g[k + 5] = g0 * J[3] + g1 * J[4] + g2 * J[5];
        g[k + 10] = g0 * J[6] + g1 * J[7] + g2 * J[8];
      }
      vis_flux_fast(u, g, fv, A.P);
      if constexpr (GOUT) // integral diagnostics: grad_disu_upts of this residual evaluation (a separate instantiation: the stores cost registers)
      {
        const size_t gi = (size_t)(q - e * NU) + (size_t)NU * S.ge[e], fs = (size_t)NU * A.n_eles;
#pragma unroll
        for (int d = 0; d < ND; d++)
#pragma unroll
          for (int k = 0; k < NF; k++) A.grad_out[gi + fs * (k + NF * d)] = g[k + NF * d];
      }
#pragma unroll
      for (int d = 0; d < ND; d++)
#pragma unroll
        for (int k = 1; k < NF; k++) f[k + NF * d] += fv[k + NF * d];
    }
#pragma unroll
    for (int k = 0; k < NF; k++)
#pragma unroll
      for (int l = 0; l < ND; l++) S.sg[l * NF + k][q] = J[l] * f[k] + J[l + 3] * f[k + 5] + J[l + 6] * f[k + 10];
  }
  __syncthreads();
  pass_D7<N, E, NT, 0>(S, A, ne);
  __syncthreads();
  pass_D7<N, E, NT, 1>(S, A, ne);
  __syncthreads();
  pass_D7<N, E, NT, 2>(S, A, ne);
  if (!A.do_update) return;
  __syncthreads();
  pass_FO7<N, E, NT, 0>(S, A, ne);
  pass_FO7<N, E, NT, 1>(S, A, ne);
}

