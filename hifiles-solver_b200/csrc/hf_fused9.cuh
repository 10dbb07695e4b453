// Generation 9 of the fused residual kernels (included by hf_fused.cu inside its anonymous namespace, after hf_fused_kernels.cuh).
//
// Same contract as generation 7 (one-sided LDG, |ldg_beta| = 0.5: the owner of a flux-point pair evaluates Riemann + LDG once and
// stores the common normal flux fc; reference src/inters.cpp:566-646, src/int_inters.cpp:160-343), re-cut so that the gradient
// kernel never touches the element interior:
//
//   k_resid9  (per element)  u -> LDG-corrected gradient -> fluxes -> divergence + correction -> RK update, and with the
//                            updated solution still in shared memory / registers: own face values at ALL flux points (fu) and,
//                            at the owned ones, the face-normal derivative of the own polynomial (gn = (l.D).u of the line
//                            behind the point: 5 more FMA per line end, no extra loads)
//   k_face9   (per element)  faces only: own fu + gn blocks and the neighbour's fu block of every face with owned points ->
//                            LDG corrections delta = u_nbr - u_own, reference-space gradient at the owned flux points
//                              normal      G_n = gn + (l.c5[n+]) delta_{n+} + (l.c5[n-]) delta_{n-}
//                              tangential  G_t = D . fu (5-wide lines inside the face) + c5[t+-] x (side face's delta extrapolated
//                                          along n to the shared edge)
//                            (tools/face_gradient_proto.py pins this formulation against the reference's dumps), viscous flux of
//                            the own side, Riemann flux, complete common normal flux -> fc
//
// Why: generation 7's gradient kernel repeated the whole line pass of the residual kernel (625 solution values in, 1 875
// gradient values through shared memory) to obtain 15 numbers at 75 flux points, and was bound by the shared-memory pipe
// (profiles/ncu_r01_summary.md).  The face kernel moves a third of those words and does a quarter of the instructions.
// The residual kernel itself: neighbour / common-flux blocks arrive by bulk copies (cp.async.bulk + mbarrier, one per 1 008-byte
// face block, issued by six threads) instead of 8-byte cp.async per double, per-thread line tasks come from a lookup table that
// makes every line access of a half-warp hit 16 distinct banks with unpadded planes (tools/bank_layout.py), and everything
// a thread needs to know about its six line ends (owner bits, signs, permuted neighbour index) is decoded once into registers.
//
// Face arrays are [block][FB] with FB = NF*NN rounded up to an even count, so that every block is 16-byte aligned (the bulk
// copies' requirement); a block is [field][flux point] as in generation 7 (= the reference's out_buffer_disu[inter][field][fpt],
// src/mpi_inters.cpp:226-229, plus one pad word at even orders).

constexpr int CL9_WORDS = 40; // per element class: header (n_e | n_t << 8), 24 edge-pass entries, 12 tangential-pass entries, pad

template <int N>
struct geo9
{
  static constexpr int P = N - 1, NN = N * N, NU = N * NN, NFP = 6 * NN;
  static constexpr int FB = (NF * NN + 1) & ~1; // doubles per face block
  static constexpr int NTASK = NF * NN;         // line tasks per direction: (field, line)
};

// ---- bulk copy / mbarrier wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared, bytes a multiple of 16, both addresses 16-byte aligned; completion is counted on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}

// shared -> global, same alignment rules; completion through the bulk async-group of the issuing thread
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait_read()
{
  asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
}
// generic-proxy writes to shared memory made visible to the bulk-copy (async) proxy
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// partition-adjacent elements of a single-launch stage: wait until the halo exchange they read from has been counted complete
__device__ __forceinline__ void wait_exchange9(const fused_args &A)
{
  if (A.wait_flag != nullptr && A.lo + (int)blockIdx.x >= A.wait_from)
  {
    if (threadIdx.x == 0)
    {
      unsigned v;
      for (;;)
      {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(A.wait_flag) : "memory");
        if ((int)(v - A.wait_value) >= 0) break;
        __nanosleep(256);
      }
    }
    __syncthreads();
  }
}
// keeps a packed word packed: the compiler cannot look through it, so the fields are extracted where they are used instead of
// being hoisted into (and spilled from) two dozen registers at kernel start
__device__ __forceinline__ unsigned opaque(unsigned x)
{
  asm volatile("" : "+r"(x));
  return x;
}

// neighbour's face-local flux point facing own flux point j (reference src/inters.cpp:232-256; every one of the four rotations
// is an involution, so the right side's inverse table equals the left side's): rot 0 flips the column, 1 the row, 2 transposes,
// 3 transposes and flips both
template <int N>
__host__ __device__ __forceinline__ int perm9(int rot, int j)
{
  constexpr int P = N - 1;
  const int i = j / N, jj = j - i * N;
  const bool tr = (rot & 2) != 0;
  int r = tr ? jj : i, c = tr ? i : jj;
  if ((rot & 3) != 2 && (rot & 3) != 1) c = P - c; // 0, 3
  if ((rot & 3) == 1 || (rot & 3) == 3) r = P - r;
  return N * r + c;
}

// per direction and thread: the line task and what the thread needs to know about the two ends of its line, as offsets into the
// shared-memory arrays.  Built on the host once per element class (elements with the same owner masks and face info share a
// table, hf_fused_prepare) -- decoding it per thread cost a seventh of the kernel's instructions.
struct task9
{
  // w0: solution-point offset field * NU + first point (11) | own-face offsets field * NN + jm, jp (8 + 8) | own_m, own_p, negate_m, negate_p
  // w1: neighbour-block offsets of jm, jp (8 + 8) | common-flux offsets of jm, jp (8 + 8)
  unsigned w0, w1;
};
struct task9x // unpacked at the point of use
{
  int v, jm, jp, xm, xp, cm, cp;
  bool own_m, own_p, neg_m, neg_p;
  __device__ __forceinline__ task9x(const task9 &t)
  {
    const unsigned w0 = opaque(t.w0), w1 = opaque(t.w1);
    v = w0 & 2047; jm = (w0 >> 11) & 255; jp = (w0 >> 19) & 255;
    own_m = (w0 >> 27) & 1; own_p = (w0 >> 28) & 1; neg_m = (w0 >> 29) & 1; neg_p = (w0 >> 30) & 1;
    xm = w1 & 255; xp = (w1 >> 8) & 255; cm = (w1 >> 16) & 255; cp = w1 >> 24;
  }
};

template <int N>
struct smem9r
{
  typedef geo9<N> G;
  static constexpr int NUP = (G::NU * NF + 1) & ~1; // solution planes, padded to an even count of doubles
  double sx[6][G::FB];       // bulk copy target: the neighbour's face values, raw (faces with owned flux points); after the L pass: the second RK register
  double sc[6][G::FB];       // bulk copy target: common normal flux; own block (neighbour's values patched in on mixed faces) or the neighbour's raw block;
                             // after the z pass: staging of the x / y faces of fu
  double sc2[2][G::FB];      // bulk copy target: the neighbour's raw common-flux block of the first two mixed faces
  double em[EM];             // bulk copy target
  unsigned long long bar, pad_;
  double su[NUP];            // solution [field][NU]; unpadded planes (the task table takes care of the banks)
  double sg[(ND * NF * G::NU + 3) & ~1]; // [plane][NU]: reference-space gradient -> transformed flux -> divergence (planes 0..4); the y and z sets end as
                                  // staging of the published face data (16-byte aligned sub-blocks, see stage_ptr9)
  static constexpr bool STAGE_IN_SG = 4 * G::FB <= NF * G::NU; // P >= 4; lower orders get their own staging blocks
  double so[STAGE_IN_SG ? 2 : 8 * G::FB];
};

// where the published face data of face f is staged before its bulk store: the z faces (produced inside the z pass) in the common-flux
// blocks of the x / y faces, which the x / y passes have used up; the x / y faces (produced after the z pass) in the x set (fu) and the
// z set (gn) of sg
template <int N, typename SM>
__device__ __forceinline__ double *stage_ptr9(SM &S, int f, bool gn)
{
  typedef geo9<N> G;
  if (f == 0 || f == 5) return &S.sc[1][0] + ((f == 5 ? 1 : 0) + (gn ? 2 : 0)) * G::FB;
  double *xset = SM::STAGE_IN_SG ? S.sg : S.so, *zset = SM::STAGE_IN_SG ? S.sg + ((2 * NF * G::NU + 1) & ~1) : S.so + 4 * G::FB;
  return (gn ? zset : xset) + (f - 1) * G::FB;
}

// the line task of one thread: lut word = field | c1 << 3 | c2 << 6 with (c1, c2) the two coordinates of the line other than dir,
// ascending; own masks and face info of the two faces normal to dir (host side, once per element class)
template <int N>
__host__ inline task9 make_task9(int dir, unsigned lut, unsigned long long own_fm, unsigned long long own_fp, int info_m, int info_p)
{
  constexpr int P = N - 1, NN = N * N, NU = N * NN;
  const int k = lut & 7, c1 = (lut >> 3) & 7, c2 = (lut >> 6) & 7;
  int base, jm, jp;
  if (dir == 0) { base = N * c1 + NN * c2; jm = (P - c1) + N * c2; jp = c1 + N * c2; }
  else if (dir == 1) { base = c1 + NN * c2; jm = c1 + N * c2; jp = (P - c1) + N * c2; }
  else { base = c1 + N * c2; jm = (P - c1) + N * c2; jp = c1 + N * c2; }
  const bool om = (own_fm >> jm) & 1ull, op = (own_fp >> jp) & 1ull;
  // a partition neighbour evaluated fc along its own (opposite) normal: negate what it sent; and fc is along the LEFT normal, so
  // the right side of a face negates as well (norm_tconf of this element = +-tdA fc)
  const bool ngm = (((info_m & 8) != 0) && !om) != ((info_m & 4) != 0);
  const bool ngp = (((info_p & 8) != 0) && !op) != ((info_p & 4) != 0);
  // a boundary face's virtual neighbour block is in the own numbering
  const int xm = (info_m & 16) ? jm : perm9<N>(info_m, jm), xp = (info_p & 16) ? jp : perm9<N>(info_p, jp);
  // common flux: a face without owned points holds the neighbour's raw block, any other the own block
  const int cm = own_fm == 0ull ? xm : jm, cp = own_fp == 0ull ? xp : jp;
  task9 t;
  t.w0 = (unsigned)(k * NU + base) | ((unsigned)(k * NN + jm) << 11) | ((unsigned)(k * NN + jp) << 19) | ((unsigned)om << 27) | ((unsigned)op << 28) |
         ((unsigned)ngm << 29) | ((unsigned)ngp << 30);
  t.w1 = (unsigned)(k * NN + xm) | ((unsigned)(k * NN + xp) << 8) | ((unsigned)(k * NN + cm) << 16) | ((unsigned)(k * NN + cp) << 24);
  return t;
}

// L pass: own face values at both ends of the line, LDG correction with the neighbour's value where this element owns the flux
// point (weight 1, else 0), corrected reference-space derivative along the line.  WAIT: the neighbour blocks are still in flight; the
// part that does not need them comes first
template <int N, int DIR, bool WAIT, typename SM>
__device__ __forceinline__ void pass_L9(SM &S, const fused_args &A, const task9 &tp)
{
  constexpr int NU = N * N * N, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  const task9x t(tp);
  const double *x = S.su + t.v;
  double v[N];
#pragma unroll
  for (int j = 0; j < N; j++) v[j] = x[j * stride];
  double um = A.tL[0][0] * v[0], up = A.tL[1][0] * v[0];
#pragma unroll
  for (int j = 1; j < N; j++) { um += A.tL[0][j] * v[j]; up += A.tL[1][j] * v[j]; }
  double acc[N];
#pragma unroll
  for (int i = 0; i < N; i++)
  {
    acc[i] = A.tD[i * N] * v[0];
#pragma unroll
    for (int j = 1; j < N; j++) acc[i] += A.tD[i * N + j] * v[j];
  }
  if (WAIT) mbar_wait(&S.bar, 0);
  // unconditional loads (blocks of faces without owned points are never copied in: stale data, discarded by the select)
  const double xm = S.sx[FM][t.xm], xp = S.sx[FP][t.xp];
  const double dm = t.own_m ? xm - um : 0.;
  const double dp = t.own_p ? xp - up : 0.;
  double *o = S.sg + DIR * NF * NU + t.v;
#pragma unroll
  for (int i = 0; i < N; i++)
  {
    acc[i] += A.tc5[FP * N + i] * dp;
    acc[i] += A.tc5[FM * N + i] * dm;
    o[i * stride] = acc[i];
  }
}

// what an element publishes about the line through its updated solution: own face values at both ends (all flux points) and,
// where it owns the point, the face-normal derivative of its own polynomial -- into the staging blocks of the two faces
template <int N, int DIR, typename SM>
__device__ __forceinline__ void face_out9(SM &S, const fused_args &A, const task9x &t, const double *__restrict__ v)
{
  constexpr int FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  double um = A.tL[0][0] * v[0], up = A.tL[1][0] * v[0];
#pragma unroll
  for (int j = 1; j < N; j++) { um += A.tL[0][j] * v[j]; up += A.tL[1][j] * v[j]; }
  stage_ptr9<N>(S, FM, false)[t.jm] = um;
  stage_ptr9<N>(S, FP, false)[t.jp] = up;
  if (t.own_m)
  {
    double g = A.tLD[0][0] * v[0];
#pragma unroll
    for (int j = 1; j < N; j++) g += A.tLD[0][j] * v[j];
    stage_ptr9<N>(S, FM, true)[t.jm] = g;
  }
  if (t.own_p)
  {
    double g = A.tLD[1][0] * v[0];
#pragma unroll
    for (int j = 1; j < N; j++) g += A.tLD[1][j] * v[j];
    stage_ptr9<N>(S, FP, true)[t.jp] = g;
  }
}

// divergence pass: own normal flux at the line ends, correction with the common flux, divergence along the line accumulated in
// place; the z pass finishes with the RK update (eles::AdvanceSolution, reference src/eles.cpp:1080-1265) and the z-face data
template <int N, int DIR, typename SM>
__device__ __forceinline__ void pass_D9(SM &S, const fused_args &A, const task9 &tp, int ge)
{
  constexpr int NU = N * N * N, stride = line_dir<N, DIR>::stride, FM = line_dir<N, DIR>::fminus, FP = line_dir<N, DIR>::fplus;
  const task9x t(tp);
  const double *x = S.sg + DIR * NF * NU + t.v;
  double v[N];
#pragma unroll
  for (int j = 0; j < N; j++) v[j] = x[j * stride];
  double nm = A.tL[0][0] * v[0], np = A.tL[1][0] * v[0];
#pragma unroll
  for (int j = 1; j < N; j++) { nm += A.tL[0][j] * v[j]; np += A.tL[1][j] * v[j]; }
  const double tm = S.em[10 + 4 * FM], tq = S.em[10 + 4 * FP];
  const double sm = t.neg_m ? -tm : tm, sp = t.neg_p ? -tq : tq;
  const double dm = S.sc[FM][t.cm] * sm + nm;
  const double dp = S.sc[FP][t.cp] * sp - np;
  double *o = S.sg + (DIR == 1 ? NF * NU : 0) + t.v; // x and y passes write their own plane sets (no ordering between them), the z pass sums
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
    for (int i = 0; i < N; i++) o[i * stride] = out[i];
  }
  else
  {
    const double *oy = S.sg + NF * NU + t.v;
    const double inv_detjac = S.em[9];
    const double dtl = A.dt_local ? A.dt_local[ge] : A.rk.dt;
    const double dt_fac = A.dt_local ? dtl / A.rk.fac : A.rk.dt_fac; // (dt / fac) * r, the reference's evaluation order (src/eles.cpp:1141, 1191)
    double *us = S.su + t.v;
    double *s1 = &S.sx[0][0] + t.v; // the second RK register (staged after the L pass when the scheme reads it)
    double unew[N];
#pragma unroll
    for (int i = 0; i < N; i++)
    {
      const double acc = (o[i * stride] + oy[i * stride]) + out[i];
      if (A.keep_residual) A.div[(size_t)(t.v % NU) + i * stride + (size_t)NU * ge + (size_t)(t.v / NU) * NU * A.n_eles] = acc;
      if (acc != acc) *A.nan_flag = 1 + ge;
      double u = us[i * stride];
      if (A.do_update)
      {
        const double rr = acc * inv_detjac;
        if (A.rk.mode == 0)
          u -= dt_fac * rr;
        else if (A.rk.mode == 1)
          u = A.rk.c1 * u + A.rk.c2 * (A.rk.copy_u1 ? u : s1[i * stride]) + dt_fac * (-rr);
        else
        {
          const double dlt = A.rk.c1 * s1[i * stride] + dtl * (-rr);
          s1[i * stride] = dlt;
          u += A.rk.c2 * dlt;
        }
        us[i * stride] = u;
      }
      unew[i] = u;
    }
    if (A.do_update) face_out9<N, DIR>(S, A, t, unew);
  }
}

template <int N, int DIR, typename SM>
__device__ __forceinline__ void pass_FO9(SM &S, const fused_args &A, const task9 &tp)
{
  constexpr int stride = line_dir<N, DIR>::stride;
  const task9x t(tp);
  const double *x = S.su + t.v;
  double v[N];
#pragma unroll
  for (int j = 0; j < N; j++) v[j] = x[j * stride];
  face_out9<N, DIR>(S, A, t, v);
}

// one array [field][ele][pt] of one element <-> contiguous shared memory [field][NU]: 8-byte cp.async in (an element's 125-double
// plane is only 8-byte aligned, bulk copies need 16), coalesced stores out
template <int N, int NT>
__device__ __forceinline__ void stage_in9(double *dst, const double *arr, int n_eles, int ge)
{
  constexpr int NU = N * N * N;
  const double *src = arr + (size_t)NU * ge;
  const size_t fs = (size_t)NU * n_eles;
  for (int i = threadIdx.x; i < NU; i += NT)
  {
#pragma unroll
    for (int k = 0; k < NF; k++) cp_async8(dst + k * NU + i, src + i + k * fs);
  }
  cp_async_commit();
}
template <int N, int NT>
__device__ __forceinline__ void stage_out9(double *arr, const double *src, int n_eles, int ge)
{
  constexpr int NU = N * N * N;
  double *dst = arr + (size_t)NU * ge;
  const size_t fs = (size_t)NU * n_eles;
  for (int i = threadIdx.x; i < NU; i += NT)
  {
#pragma unroll
    for (int k = 0; k < NF; k++) dst[i + k * fs] = src[k * NU + i];
  }
}

// bulk stores of the staged face data: fu of every face, gn of the faces with owned points (issued by thread f of the first warp)
template <int N, typename SM>
__device__ __forceinline__ void publish_faces9(SM &S, const fused_args &A, int ge, int f, bool any_owned)
{
  typedef geo9<N> G;
  bulk_s2g(A.fu_next + ((size_t)ge * 6 + f) * G::FB, stage_ptr9<N>(S, f, false), G::FB * 8);
  // a partition face (its neighbour block is a receive block behind the last element): the message goes straight into the send buffer
  const int nb = __ldg(A.nbr + (size_t)ge * 6 + f) - 6 * A.n_eles;
  if (nb >= 0) bulk_s2g(A.send_u + (size_t)nb * G::FB, stage_ptr9<N>(S, f, false), G::FB * 8);
  if (any_owned) bulk_s2g(A.gn + ((size_t)ge * 6 + f) * G::FB, stage_ptr9<N>(S, f, true), G::FB * 8);
  bulk_commit_wait_read();
}

// MODE 0: residual + update + next face data; MODE 1: GOUT (also stores grad_disu_upts); MODE 2: face data of the current
// solution only (first stage, or after an upload / shock capturing changed the solution)
template <int N, int NT, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB) k_resid9(const __grid_constant__ fused_args A)
{
  typedef geo9<N> G;
  typedef smem9r<N> SM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  constexpr int NN = G::NN, NU = G::NU, FB = G::FB;
  const int tid = threadIdx.x;
  const int ge = elem_id(A, A.lo + blockIdx.x);
  if (MODE != 2) wait_exchange9(A);
  stage_in9<N, NT>(S.su, A.u0, A.n_eles, ge);
  // element class | face kinds << 20 (2 bits per face: 0 no owned flux point, 1 all owned, 2 mixed); the thread's three line tasks
  // (one per direction) come from the class's table
  const unsigned ce = __ldg(A.cls9 + ge);
  const bool has_task = tid < G::NTASK;
  task9 t0, t1, t2;
  {
    const uint2 *tw = A.tw9 + (size_t)(ce & 0xfffffu) * (3 * G::NTASK) + (has_task ? tid : 0);
    const uint2 a = __ldg(tw), b = __ldg(tw + G::NTASK), c = __ldg(tw + 2 * G::NTASK);
    t0.w0 = a.x; t0.w1 = a.y; t1.w0 = b.x; t1.w1 = b.y; t2.w0 = c.x; t2.w1 = c.y;
  }
  const unsigned kinds = ce >> 20;
  const unsigned kind_mine = (kinds >> (2 * (tid < 6 ? tid : 0))) & 3u; // of face tid
  // mixed faces (some points owned, some not: the reference's sign switch fires on rounding-level normal components,
  // src/inters.cpp:566-581) take the neighbour's common flux at the points it owns: owner mask and face info of the flux points
  // this thread will patch, fetched now
  constexpr int NITQ = (6 * NN + NT - 1) / NT;
  unsigned long long pm[NITQ];
  int pinfo[NITQ];
  if (MODE != 2 && (kinds & 0xaaau))
  {
#pragma unroll
    for (int it = 0; it < NITQ; it++)
    {
      const int q = tid + it * NT, f = q / NN;
      pm[it] = ~0ull;
      pinfo[it] = 0;
      if (q < 6 * NN && ((kinds >> (2 * f)) & 3u) == 2u)
      {
        pm[it] = __ldg(A.bmask + (size_t)ge * 6 + f) ^ A.own_xor;
        pinfo[it] = __ldg(A.finfo + (size_t)ge * 6 + f);
      }
    }
  }
  if constexpr (MODE != 2)
  {
    // the first warp posts the bulk copies: one thread per face fetches the neighbour's values where this element owns flux points,
    // and the common flux from the own block (any owned point) or the neighbour's (none)
    if (tid < 32)
    {
      int nb = 0;
      if (tid < 6) nb = __ldg(A.nbr + (size_t)ge * 6 + tid);
      if (tid == 0) mbar_init(&S.bar, 6);
      __syncwarp();
      if (tid < 6)
      {
        const int f = tid;
        const int slot = __popc(kinds & 0xaaau & ((1u << (2 * f)) - 1u)); // mixed faces below this one
        const bool both = kind_mine == 2u && slot < 2;
        mbar_expect_tx(&S.bar, (kind_mine != 0u ? (both ? 3u : 2u) : 1u) * (unsigned)(FB * 8) + (f == 0 ? (unsigned)(EM * 8) : 0u));
        if (kind_mine != 0u)
        {
          bulk_g2s(S.sx[f], A.fu_cur + (size_t)nb * FB, FB * 8, &S.bar);
          bulk_g2s(S.sc[f], A.fv + ((size_t)ge * 6 + f) * FB, FB * 8, &S.bar);
          if (both) bulk_g2s(S.sc2[slot], A.fv + (size_t)nb * FB, FB * 8, &S.bar);
        }
        else
          bulk_g2s(S.sc[f], A.fv + (size_t)nb * FB, FB * 8, &S.bar);
        if (f == 0) bulk_g2s(S.em, A.em + (size_t)ge * EM, EM * 8, &S.bar);
      }
    }
  }
  if constexpr (MODE == 2)
  {
    cp_async_wait_all();
    __syncthreads();
    if (has_task)
    {
      pass_FO9<N, 0>(S, A, t0);
      pass_FO9<N, 1>(S, A, t1);
      pass_FO9<N, 2>(S, A, t2);
    }
    fence_async_smem();
    __syncthreads();
    if (tid < 6) publish_faces9<N>(S, A, ge, tid, kind_mine != 0u);
    return;
  }
  else
  {
    cp_async_wait_all();
    __syncthreads(); // the solution is in; the mbarrier is initialised
    // SSP schemes: the first stage keeps the old solution in the second register (reference src/eles.cpp:1107-1117)
    if (A.do_update && A.rk.copy_u1) stage_out9<N, NT>(A.u1, S.su, A.n_eles, ge);
    if (has_task)
    {
      pass_L9<N, 0, true>(S, A, t0);
      pass_L9<N, 1, false>(S, A, t1);
      pass_L9<N, 2, false>(S, A, t2);
    }
    else
      mbar_wait(&S.bar, 0);
    // mixed faces: the neighbour's common flux at the points it owns, patched into the own block (from the raw copy in sc2; a
    // third and further mixed face of one element goes through global memory)
    if (kinds & 0xaaau)
    {
#pragma unroll
      for (int it = 0; it < NITQ; it++)
      {
        const int q = tid + it * NT, f = q / NN, j = q - f * NN;
        if (q < 6 * NN && !((pm[it] >> j) & 1ull))
        {
          const int slot = __popc(kinds & 0xaaau & ((1u << (2 * f)) - 1u));
          const int pj = perm9<N>(pinfo[it], j);
          if (slot < 2)
          {
#pragma unroll
            for (int k = 0; k < NF; k++) S.sc[f][k * NN + j] = S.sc2[slot][k * NN + pj];
          }
          else
          {
            const double *src = A.fv + (size_t)__ldg(A.nbr + (size_t)ge * 6 + f) * FB + pj;
#pragma unroll
            for (int k = 0; k < NF; k++) S.sc[f][k * NN + j] = src[k * NN];
          }
        }
      }
    }
    __syncthreads();
    // the neighbour values are used up: their place takes the second RK register where the scheme reads it
    const bool need_u1 = A.do_update && ((A.rk.mode == 1 && !A.rk.copy_u1) || A.rk.mode == 2);
    if (need_u1) stage_in9<N, NT>(&S.sx[0][0], A.u1, A.n_eles, ge);
    for (int q = tid; q < NU; q += NT)
    {
      const double *J = S.em;
      double u[NF], f[NF * ND];
#pragma unroll
      for (int k = 0; k < NF; k++) u[k] = S.su[k * NU + q];
      inv_flux_fast(u, f, A.P.gamma - 1.0);
      {
        double g[NF * ND], fv[NF * ND];
        const double idj = J[9];
#pragma unroll
        for (int k = 0; k < NF; k++)
        {
          const double g0 = S.sg[k * NU + q] * idj, g1 = S.sg[(NF + k) * NU + q] * idj, g2 = S.sg[(2 * NF + k) * NU + q] * idj;
          g[k] = g0 * J[0] + g1 * J[1] + g2 * J[2];
          g[k + 5] = g0 * J[3] + g1 * J[4] + g2 * J[5];
          g[k + 10] = g0 * J[6] + g1 * J[7] + g2 * J[8];
        }
        vis_flux_fast(u, g, fv, A.P);
        if constexpr (MODE == 1) // integral diagnostics: grad_disu_upts of this residual evaluation
        {
          const size_t gi = (size_t)q + (size_t)NU * ge, fs = (size_t)NU * A.n_eles;
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
        for (int l = 0; l < ND; l++) S.sg[(l * NF + k) * NU + q] = J[l] * f[k] + J[l + 3] * f[k + 5] + J[l + 6] * f[k + 10];
    }
    __syncthreads();
    if (has_task)
    {
      pass_D9<N, 0>(S, A, t0, ge);
      pass_D9<N, 1>(S, A, t1, ge);
    }
    if (need_u1) cp_async_wait_all();
    __syncthreads();
    if (has_task) pass_D9<N, 2>(S, A, t2, ge);
    if (!A.do_update) return;
    __syncthreads();
    // the updated solution (and the second register of the low-storage schemes) leave in coalesced rows; the x and y lines publish
    // their ends
    stage_out9<N, NT>(A.u0_out, S.su, A.n_eles, ge);
    if (A.rk.mode == 2) stage_out9<N, NT>(A.u1, &S.sx[0][0], A.n_eles, ge);
    if (has_task)
    {
      pass_FO9<N, 0>(S, A, t0);
      pass_FO9<N, 1>(S, A, t1);
    }
    fence_async_smem();
    __syncthreads();
    if (tid < 6) publish_faces9<N>(S, A, ge, tid, kind_mine != 0u);
  }
}

// =====================================================================================================================
// k_face9: the common normal flux at the owned flux points, from face data only
template <int N>
struct smem9f
{
  typedef geo9<N> G;
  double uf[6][G::FB];    // bulk: own face values
  double dl[6][G::FB];    // bulk: the neighbour's raw block -> LDG correction delta (own numbering, 0 where not owned), in place
  double gn[6][G::FB];    // bulk: face-normal derivative of the own polynomial
  double em[EM];          // bulk
  unsigned long long bar;
  double gt[2][6][G::FB]; // the two tangential components (directions in ascending order)
  double ed[6][2][2][NF][N]; // per face, tangential slot, side: the side face's delta extrapolated along the face normal to the shared edge
  unsigned long long own[6];
  int info[6];
  int send[6];            // partition faces: interface number (block in the send buffer), else -1
  int n_owned;
  unsigned cl[CL9_WORDS];  // the element class's pass lists
  unsigned short olist[6 * N * N];
  int vblk[6];            // boundary faces (BDY variant): the face's virtual neighbour block in fu (the boundary's common solution, for k_resid9), else -1
};

__host__ __device__ __forceinline__ int dir_minus_face9(int d) { return d == 0 ? 4 : (d == 1 ? 1 : 0); }
__host__ __device__ __forceinline__ int dir_plus_face9(int d) { return d == 0 ? 2 : (d == 1 ? 3 : 5); }
// the two directions tangential to direction n, ascending
__host__ __device__ __forceinline__ int tan_dir9(int n, int slot) { return n == 0 ? (slot ? 2 : 1) : (n == 1 ? (slot ? 2 : 0) : (slot ? 1 : 0)); }
// faces 0, 3, 4 number their flux points against the lower tangential coordinate (reference src/eles_hexas.cpp:224-282)
__host__ __device__ __forceinline__ bool face_rev9(int f) { return f == 0 || f == 3 || f == 4; }

// pass lists of k_face9 for one element class (host side)
template <int N>
__host__ inline void make_class_lists9(const unsigned long long *own, unsigned *out)
{
  for (int i = 0; i < CL9_WORDS; i++) out[i] = 0;
  int n_e = 0, n_t = 0;
  for (int F = 0; F < 6; F++)
  {
    if (own[F] == 0ull) continue;
    const int n = face_dir(F);
    for (int slot = 0; slot < 2; slot++)
    {
      out[1 + 24 + n_t++] = (unsigned)F | ((unsigned)slot << 3);
      const int t = tan_dir9(n, slot), o = tan_dir9(n, 1 - slot);
      for (int side = 0; side < 2; side++)
      {
        const int ft = side ? dir_plus_face9(t) : dir_minus_face9(t);
        const unsigned ed_off = (unsigned)(((F * 2 + slot) * 2 + side) * NF * N);
        out[1 + n_e++] = ed_off | ((unsigned)ft << 10) | ((unsigned)face_rev9(ft) << 13) | ((unsigned)(n < o) << 14) | ((unsigned)(face_sgn(F) > 0) << 15) |
                         ((unsigned)(own[ft] != 0ull) << 16);
      }
    }
  }
  out[0] = (unsigned)n_e | ((unsigned)n_t << 8);
}

// BDY: the mesh has boundary faces.  A boundary face is a face whose flux points the element all owns; its "neighbour" is the ghost state
// of bdy_inters::set_boundary_conditions (reference src/bdy_inters.cpp:213-338, 1024-1136; device restatement hf_bc.cuh) evaluated here
// from the own face values:
//   LDG common solution  u_c = u_r (ldg_solution, flux_spec 1; the viscous wall state on walls)  -> delta = u_c - u_l, and u_c goes into the
//                        face's virtual neighbour block of fu, where k_resid9 finds it like any neighbour's face values
//   inviscid             Riemann(u_l, u_r of the inviscid boundary state); slip_wall_dual: F(u_l).n
//   viscous              F_v(u_r, boundary gradient of the own corrected gradient).n - tau (u_r - u_l)  (ldg_flux, flux_spec 1); none on a slip wall
// A kernel variant of its own (the ghost-state code costs registers): meshes without boundary faces keep the other one.
// Two launches on such a mesh: BDYM = 2, the plain code over all elements, those with a boundary face leaving at once; BDYM = 1, the variant
// with the ghost-state code over the list of elements with a boundary face (a few per cent of a mesh).  BDYM = 0: no boundary faces.
template <int N, int NT, int MINB, bool ROEM, int BDYM>
// (the boundary launch is a few per cent of the mesh: its variant gets registers rather than occupancy, at most five CTAs per SM)
__global__ void __launch_bounds__(NT, BDYM == 1 ? (MINB > 6 ? 5 : (MINB > 1 ? MINB - 1 : 1)) : MINB) k_face9(const __grid_constant__ fused_args A)
{
  typedef geo9<N> G;
  typedef smem9f<N> SM;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SM &S = *reinterpret_cast<SM *>(smem_raw);
  constexpr int P = G::P, NN = G::NN, FB = G::FB, NW = NT / 32;
  static_assert(NF * N <= 32, "one lane per (field, edge point)");
  constexpr bool BDY = BDYM == 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ge = elem_id(A, A.lo + blockIdx.x);
  if constexpr (BDYM == 2)
  {
    if (A.bskip[ge]) return; // an element with a boundary face: the other launch's
  }
  wait_exchange9(A);
  // the first warp posts the bulk copies: one thread per face with owned flux points fetches the own face values, the own normal
  // derivative and the neighbour's face values
  if (warp == 0)
  {
    unsigned long long o = 0ull;
    int nb = 0;
    bool bdy = false;
    if (lane < 6)
    {
      o = __ldg(A.bmask + (size_t)ge * 6 + lane) ^ A.own_xor;
      nb = __ldg(A.nbr + (size_t)ge * 6 + lane);
      const int info = __ldg(A.finfo + (size_t)ge * 6 + lane);
      bdy = BDY && (info & 16) != 0;
      S.own[lane] = o;
      S.info[lane] = info;
      S.send[lane] = bdy ? -1 : nb - 6 * A.n_eles;
      if (BDY) S.vblk[lane] = bdy ? nb : -1;
    }
    if (lane == 0)
    {
      mbar_init(&S.bar, 6);
      S.n_owned = 0;
    }
    __syncwarp();
    if (lane < 6)
    {
      const int f = lane;
      const bool act = o != 0ull;
      mbar_expect_tx(&S.bar, (act ? (bdy ? 2u : 3u) * (unsigned)(FB * 8) : 0u) + (f == 0 ? (unsigned)(EM * 8) : 0u));
      if (act)
      {
        const size_t ob = ((size_t)ge * 6 + f) * FB;
        bulk_g2s(S.uf[f], A.fu_cur + ob, FB * 8, &S.bar);
        bulk_g2s(S.gn[f], A.gn + ob, FB * 8, &S.bar);
        if (!bdy) bulk_g2s(S.dl[f], A.fu_cur + (size_t)nb * FB, FB * 8, &S.bar); // a boundary face has no neighbour block: its delta comes from the ghost state
      }
      if (f == 0) bulk_g2s(S.em, A.em + (size_t)ge * EM, EM * 8, &S.bar);
    }
  }
  if (warp == (NW > 1 ? 1 : 0))
  {
    const unsigned *cl = A.cl9 + (size_t)(__ldg(A.cls9 + ge) & 0xfffffu) * CL9_WORDS;
    for (int i = lane; i < CL9_WORDS; i += 32) S.cl[i] = __ldg(cl + i);
  }
  __syncthreads();
  mbar_wait(&S.bar, 0);
  // ---- delta = u_nbr - u_own at the owned flux points (own numbering), 0 elsewhere on faces with owned points; the list of owned points
  {
    constexpr int NIT = (6 * NN + NT - 1) / NT;
    double d[NIT][NF];
#pragma unroll
    for (int it = 0; it < NIT; it++)
    {
      const int q = tid + it * NT;
      if (q < 6 * NN)
      {
        const int f = q / NN, j = q - f * NN;
        const unsigned long long o = S.own[f];
        if (o != 0ull)
        {
          const bool mine = (o >> j) & 1ull;
          if (BDY && (S.info[f] & 16))
          {
            // the LDG common solution of a boundary flux point is the ghost state (the viscous wall state on walls)
            double ul[NF], ur[NF];
#pragma unroll
            for (int k = 0; k < NF; k++) { ul[k] = S.uf[f][k * NN + j]; ur[k] = 0.; }
            const hf_bc &B = A.bct[S.info[f] >> 8];
            const double *nrm = &S.em[10 + 4 * f + 1];
            set_boundary_conditions<ND, NF>(0, B, ul, ur, nrm, A.P.gamma, A.R_ref);
            if (hf_is_wall(B.bc_flag)) set_boundary_conditions<ND, NF>(1, B, ul, ur, nrm, A.P.gamma, A.R_ref);
            double *vb = const_cast<double *>(A.fu_cur) + (size_t)S.vblk[f] * FB + j;
#pragma unroll
            for (int k = 0; k < NF; k++)
            {
              d[it][k] = ur[k] - ul[k];
              vb[k * NN] = ur[k];
            }
          }
          else
          {
            const int pj = perm9<N>(S.info[f], j);
#pragma unroll
            for (int k = 0; k < NF; k++) d[it][k] = mine ? S.dl[f][k * NN + pj] - S.uf[f][k * NN + j] : 0.;
          }
          if (mine) S.olist[atomicAdd(&S.n_owned, 1)] = (unsigned short)q;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < NIT; it++)
    {
      const int q = tid + it * NT;
      if (q < 6 * NN)
      {
        const int f = q / NN, j = q - f * NN;
        if (S.own[f] != 0ull)
        {
#pragma unroll
          for (int k = 0; k < NF; k++) S.dl[f][k * NN + j] = d[it][k];
        }
      }
    }
  }
  __syncthreads();
  // ---- edge pass: for every face F with owned points and each of its four side faces ft, ft's delta extrapolated along F's normal
  //      direction to the shared edge.  The class's list holds one entry per (F, tangential slot, side) with F active; one entry per
  //      warp iteration (warp-uniform geometry), a lane is one (field, edge point)
  const unsigned hdr = S.cl[0];
  {
    const int k = lane / N, m = lane - k * N;
    const int n_e = hdr & 255;
    if (lane < NF * N)
      for (int c = warp; c < n_e; c += NW)
      {
        // entry: offset into ed (10) | ft (3) | ft numbers its lower in-face coordinate backwards | F's normal is the lower of ft's two
        // in-face directions | F is a plus face | ft has owned points
        const unsigned e = S.cl[1 + c];
        double acc = 0.;
        if (e & (1u << 16))
        {
          const int ft = (e >> 10) & 7;
          const bool rev = (e >> 13) & 1, nlo = (e >> 14) & 1, plus = (e >> 15) & 1;
          const int a0 = rev ? P : 0, sgn = rev ? -1 : 1;
          const int base = nlo ? N * m + a0 : a0 + sgn * m, step = nlo ? sgn : N;
          const double *x = S.dl[ft] + k * NN + base;
#pragma unroll
          for (int i = 0; i < N; i++) acc += (plus ? A.tL[1][i] : A.tL[0][i]) * x[i * step];
        }
        (&S.ed[0][0][0][0][0])[(e & 1023) + lane] = acc;
      }
  }
  __syncthreads();
  // ---- tangential pass: in-face line derivative of the own face values + the two edge corrections; one (F, slot) with F active per
  //      warp iteration, a lane is one (field, line)
  {
    const int k = lane / N, m = lane - k * N;
    const int n_t = (hdr >> 8) & 255;
    if (lane < NF * N)
      for (int c = warp; c < n_t; c += NW)
      {
        const unsigned e = S.cl[1 + 24 + c]; // F (3) | slot (1)
        const int F = e & 7, slot = (e >> 3) & 1;
        const bool rev = face_rev9(F);
        // slot 0 runs along jj (reversed on faces 0, 3, 4), slot 1 along i; m = the line's coordinate along the other in-face index
        const int base = slot == 0 ? N * m + (rev ? P : 0) : m, step = slot == 0 ? (rev ? -1 : 1) : N, mo = (slot == 1 && rev) ? P - m : m;
        const double *x = S.uf[F] + k * NN + base;
        double v[N];
#pragma unroll
        for (int j = 0; j < N; j++) v[j] = x[j * step];
        const double edm = S.ed[F][slot][0][k][mo], edp = S.ed[F][slot][1][k][mo];
        double *o = S.gt[slot][F] + k * NN + base;
#pragma unroll
        for (int i = 0; i < N; i++)
        {
          double acc = A.tD[i * N] * v[0];
#pragma unroll
          for (int j = 1; j < N; j++) acc += A.tD[i * N + j] * v[j];
          acc += A.c5s[1][i] * edp;
          acc += A.c5s[0][i] * edm;
          o[i * step] = acc;
        }
      }
  }
  __syncthreads();
  // ---- owned flux points: physical gradient, viscous flux of the own side, Riemann flux, common normal flux
  const int n_owned = S.n_owned;
  for (int i = tid; i < n_owned; i += NT)
  {
    const int q = S.olist[i];
    const int f = q / NN, j = q - f * NN;
    const int info = S.info[f];
    const bool is_right = (info & 4) != 0;
    const double *J = S.em;
    const double idj = J[9];
    const int n = face_dir(f);
    const bool plus = face_sgn(f) > 0;
    const int fo = plus ? dir_minus_face9(n) : dir_plus_face9(n);
    const int jo = j + P - 2 * (j % N); // the other end of the line: jj -> P - jj
    const bool opp = (S.own[fo] >> jo) & 1ull;
    // l(s) . c5 of the own face and of the opposite one
    const double lc_self = plus ? A.lc5s[1][1] : A.lc5s[0][0], lc_opp = plus ? A.lc5s[0][1] : A.lc5s[1][0];
    double uo[NF], un[NF], fn[NF], vn[NF];
    if constexpr (!BDY) // meshes without boundary faces: the code the headline configuration was tuned with, untouched
    {
      {
        double g[NF * ND], fv[NF * ND];
#pragma unroll
        for (int k = 0; k < NF; k++)
        {
          uo[k] = S.uf[f][k * NN + j];
          const double ds = S.dl[f][k * NN + j];
          const double dop = S.dl[fo][k * NN + jo];
          un[k] = uo[k] + ds; // the neighbour's value back from the correction (an owned point: delta = u_nbr - u_own)
          const double an = (S.gn[f][k * NN + j] + lc_self * ds + (opp ? lc_opp * dop : 0.)) * idj;
          const double a0 = S.gt[0][f][k * NN + j] * idj, a1 = S.gt[1][f][k * NN + j] * idj;
          const double gr0 = n == 0 ? an : a0;                 // t0 = 0 unless n = 0
          const double gr1 = n == 1 ? an : (n == 0 ? a0 : a1); // direction 1 is t0 for n = 0, t1 for n = 2
          const double gr2 = n == 2 ? an : a1;                 // t1 = 2 unless n = 2
          g[k] = gr0 * J[0] + gr1 * J[1] + gr2 * J[2];
          g[k + 5] = gr0 * J[3] + gr1 * J[4] + gr2 * J[5];
          g[k + 10] = gr0 * J[6] + gr1 * J[7] + gr2 * J[8];
        }
        vis_flux_fast(uo, g, fv, A.P);
        const double *nrm = &S.em[10 + 4 * f + 1];
        const double n0 = nrm[0], n1 = nrm[1], n2 = nrm[2];
        vn[0] = 0.;
#pragma unroll
        for (int k = 1; k < NF; k++) vn[k] = fv[k] * n0 + fv[k + 5] * n1 + fv[k + 10] * n2;
      }
      {
        const double *nrm = &S.em[10 + 4 * f + 1];
        double nl[3] = {nrm[0], nrm[1], nrm[2]};
        if constexpr (ROEM) // RoeM: the reference's normal of exactly this flux point (hf_fused_prepare); a kernel variant of its own, so that the
        {                   // other solvers do not carry the pointer (the run-time test cost k_face9 2 % in registers / spills)
          const double *q3 = A.nlf + ((size_t)(ge * 6 + f) * NN + j) * 3;
          nl[0] = q3[0]; nl[1] = q3[1]; nl[2] = q3[2];
        }
        double ul[NF], ur[NF];
#pragma unroll
        for (int k = 0; k < NF; k++)
        {
          ul[k] = is_right ? un[k] : uo[k];
          ur[k] = is_right ? uo[k] : un[k];
        }
        riemann_fast(ul, ur, nl, fn, A.P);
      }
      const double ts = is_right ? -A.P.ldg_tau : A.P.ldg_tau;
      fn[0] -= ts * (un[0] - uo[0]);
#pragma unroll
      for (int k = 1; k < NF; k++) fn[k] += vn[k] - ts * (un[k] - uo[k]);
    }
    else
    {
      const double *nrm = &S.em[10 + 4 * f + 1];
      double nl[3] = {nrm[0], nrm[1], nrm[2]};
      if constexpr (ROEM) // RoeM: the reference's normal of exactly this flux point (hf_fused_prepare); a kernel variant of its own, so that the
      {                   // other solvers do not carry the pointer (the run-time test cost k_face9 2 % in registers / spills)
        const double *q3 = A.nlf + ((size_t)(ge * 6 + f) * NN + j) * 3;
        nl[0] = q3[0]; nl[1] = q3[1]; nl[2] = q3[2];
      }
      double g[NF * ND];
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        uo[k] = S.uf[f][k * NN + j];
        const double ds = S.dl[f][k * NN + j];
        const double dop = S.dl[fo][k * NN + jo];
        un[k] = uo[k] + ds; // the neighbour's value back from the correction (an owned point: delta = u_nbr - u_own)
        const double an = (S.gn[f][k * NN + j] + lc_self * ds + (opp ? lc_opp * dop : 0.)) * idj;
        const double a0 = S.gt[0][f][k * NN + j] * idj, a1 = S.gt[1][f][k * NN + j] * idj;
        const double gr0 = n == 0 ? an : a0;                 // t0 = 0 unless n = 0
        const double gr1 = n == 1 ? an : (n == 0 ? a0 : a1); // direction 1 is t0 for n = 0, t1 for n = 2
        const double gr2 = n == 2 ? an : a1;                 // t1 = 2 unless n = 2
        g[k] = gr0 * J[0] + gr1 * J[1] + gr2 * J[2];
        g[k + 5] = gr0 * J[3] + gr1 * J[4] + gr2 * J[5];
        g[k + 10] = gr0 * J[6] + gr1 * J[7] + gr2 * J[8];
      }
      if (BDY && (info & 16))
      {
        // boundary flux point (reference src/bdy_inters.cpp:213-338 inviscid, :1024-1090 viscous)
        const hf_bc &B = A.bct[info >> 8];
        double ur[NF];
#pragma unroll
        for (int k = 0; k < NF; k++) ur[k] = 0.;
        set_boundary_conditions<ND, NF>(0, B, uo, ur, nrm, A.P.gamma, A.R_ref);
        if (B.bc_flag == HF_SLIP_WALL_DUAL) // dual-consistent wall: the common flux is the left normal flux
        {
          side_state L;
          make_side(uo, nl, A.P.gamma - 1.0, L);
#pragma unroll
          for (int k = 0; k < NF; k++) fn[k] = L.fn[k];
        }
        else
          riemann_fast(uo, ur, nl, fn, A.P);
        if (B.bc_flag != HF_SLIP_WALL) // a slip wall carries no viscous flux
        {
          double gr[NF * ND], fv[NF * ND];
          set_boundary_conditions<ND, NF>(1, B, uo, ur, nrm, A.P.gamma, A.R_ref);
          set_boundary_gradients<ND, NF>(B.bc_flag, ur, g, gr, nrm);
          vis_flux_fast(ur, gr, fv, A.P);
          fn[0] -= A.P.ldg_tau * (ur[0] - uo[0]);
#pragma unroll
          for (int k = 1; k < NF; k++) fn[k] += (fv[k] * nrm[0] + fv[k + 5] * nrm[1] + fv[k + 10] * nrm[2]) - A.P.ldg_tau * (ur[k] - uo[k]);
        }
      }
      else
      {
        {
          double fv[NF * ND];
          vis_flux_fast(uo, g, fv, A.P);
          vn[0] = 0.;
#pragma unroll
          for (int k = 1; k < NF; k++) vn[k] = fv[k] * nrm[0] + fv[k + 5] * nrm[1] + fv[k + 10] * nrm[2];
        }
        {
          double ul[NF], ur[NF];
#pragma unroll
          for (int k = 0; k < NF; k++)
          {
            ul[k] = is_right ? un[k] : uo[k];
            ur[k] = is_right ? uo[k] : un[k];
          }
          riemann_fast(ul, ur, nl, fn, A.P);
        }
        const double ts = is_right ? -A.P.ldg_tau : A.P.ldg_tau;
        fn[0] -= ts * (un[0] - uo[0]);
#pragma unroll
        for (int k = 1; k < NF; k++) fn[k] += vn[k] - ts * (un[k] - uo[k]);
      }
    }
    double *out = A.fv + ((size_t)ge * 6 + f) * FB + j;
#pragma unroll
    for (int k = 0; k < NF; k++) out[k * NN] = fn[k];
    if (S.send[f] >= 0) // partition face: the same values into the message to the neighbour rank
    {
      double *msg = A.send_g + (size_t)S.send[f] * FB + j;
#pragma unroll
      for (int k = 0; k < NF; k++) msg[k * NN] = fn[k];
    }
  }
}
