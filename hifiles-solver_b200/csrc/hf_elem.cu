// Blocked element kernels: the element-local operator chain of one RK stage for EVERY element type (quadrilaterals, triangles,
// tetrahedra, prisms, and hexahedra that the sum-factorised kernels of hf_fused.cu do not take: curved ones, meshes with boundary
// faces), as two kernels per element type and stage instead of one kernel per reference method.
//
//   k_elem_grad  (viscous)   u, delta_disu_fpts -> reference-space gradient  opp_4(d) u + opp_5(d) delta   (eles::calculate_gradient +
//                            the correction of eles::correct_gradient, reference src/eles.cpp:1823-1886, 1890-1950) -> opp_6 to the
//                            flux points -> physical gradient with the flux-point metrics (:1990-2011) -> grad_disu_fpts
//   k_elem_resid             u [, delta] -> gradient again (cheaper than storing it) -> physical gradient at the solution points
//                            (:1955-1986) -> inviscid + viscous flux, transformed (evaluate_invFlux :1415-1478, evaluate_viscFlux
//                            :2285-2392) -> divergence opp_2(d) (:1651-1725) and normal flux at the flux points opp_1(d) (:1549-1620)
//                            -> opp_3 (common - own normal flux) (calculate_corrected_divergence :1738-1817) -> RK update
//                            (AdvanceSolution :1080-1265) -> opp_0 of the UPDATED solution = the next stage's extrapolate_solution
//                            (:1360-1411) -> disu_upts(0), disu_fpts
//
// The interface kernels in between (k_int_* / k_bdy_* / k_mpi_* of hf_device.cu: Riemann solvers, LDG, the twelve boundary kinds,
// wall model, halo exchange) are the staged ones, unchanged: they read disu_fpts / grad_disu_fpts and write norm_tconf_fpts /
// delta_disu_fpts.  What disappears is every element-local intermediate array of the reference (tdisf_upts, grad_disu_upts,
// norm_tdisf_fpts, div_tconf_upts: SURVEY §8a rows a4-a6, a9-a12, a15, a16 in two launches).
//
// One CTA takes a tile of E elements: NC = E * n_fields columns of every array live in shared memory as [column][point] with a
// column stride = 4 (mod 8) doubles.  Every small-operator product is a batched contraction on the FP64 tensor cores
// (mma.sync.m8n8k4.f64): the DATA tile is the A operand (8 columns x 4 points, read from shared memory without bank conflicts), the
// OPERATOR is the B operand, stored on the host in fragment order (one coalesced 256-byte read per k-step, L1-resident, zero padding
// baked in), and a lane's two results are neighbouring points of one column (one 16-byte shared-memory store).  A warp task is one
// block of 8 operator rows against CG column blocks, the tasks of all products of a phase are pooled over the CTA's warps.
// Accumulation order is the tensor core's and products are fused, so results differ from the reference's ascending dgemm sums in
// the last bits (1e-15 relative): this is the fast mode (hf_dev_set_mode(ctx, 1)); mode 0 keeps the bit-exact staged kernels.
#include "hf_device.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace
{
constexpr int EL_THREADS = 256;

// host-side description of one product: dst (mode) sum_t op_t * src_t
// an operator on the device in fragment order.  Dense: [row block][k step][lane], k steps padded to pairs.  Sparse (tensor-product
// operators of quadrilaterals / hexahedra: most 8 x 4 fragments are entirely zero): per row block only the non-zero fragments, with the
// list of their k steps.
struct el_op
{
  double *frag = nullptr;
  int *kidx = nullptr;
  bool sparse = false;
  int kb = 0;                 // dense: k steps
  std::vector<int> off, cnt;  // sparse: first fragment / number of fragments of every row block
};
struct el_term
{
  const el_op *op;
  int src, ss;      // shared-memory offset and column stride of the data (doubles)
};
struct el_prod
{
  el_term t[4];
  int n_terms, rb; // row blocks of 8
  int dst, ds;     // shared-memory offset and column stride of the result
  int mode;        // 0 dst = acc, 1 dst += acc, 2 dst -= acc
};
// what a warp executes: one block of 8 operator rows against two column blocks, every offset resolved on the host (decoding
// (product, row block, column group) per task in the kernel cost four times the instructions of the tensor-core loop itself)
struct __align__(16) el_task
{
  int n_terms, mode, dst, ds; // read as one int4
  int two, pad0, pad1, pad2;
  struct __align__(16)
  {
    int kb2, src, ss, pad; // read as one int4; kb2 < 0: -kb2 fragments of a sparse operator, their k steps listed in kidx
    const double *op;      // already at the task's row block
    const int *kidx;       // sparse operators: k step of every stored fragment of this row block
  } t[4];
};
struct el_phase
{
  const el_task *tasks;
  int n;
};

struct el_args
{
  int n_eles, nu, nf, E, mb; // mb: column blocks of 8 (E * n_fields rounded up)
  int SU, SF;                // column strides: solution-point arrays, flux-point arrays
  unsigned mg_u, mg_f;       // ceil(2^32 / nu), ceil(2^32 / nf): r / n = umulhi(r, magic) for the small r of a tile
  int o_u, o_g, o_dl, o_fc, o_gf, o_dj;
  int visc, inv_from_global, store_div, store_grad;
  int over_int;         // 1: de-aliased inviscid flux inside k_elem_resid (phases ph_oi, ph_filt; at o_cub, stride SC: the solution at the cubature points, then D flux planes)
  int n_cub, SC, o_cub;
  unsigned mg_c;        // ceil(2^32 / n_cub)
  const double *JG_cub; // JGinv_over_int_cubpts (l,m,cubpt,ele)
  int grad_from_global; // k_elem_grad has stored the physical gradient at the solution points: k_elem_resid reads it instead of forming it again
  const double *u_in;
  double *u0, *u1;
  const double *delu, *ntconf;
  double *disu_fpts, *grad_fpts, *div_out, *grad_out;
  const double *tdisf_in;
  const double *detjac_u, *JG_u, *detjac_f, *JG_f, *dt_local;
  double dt, fac, c1, c2;
  int rk_mode, rk_copy;
  int *nan_flag;
  el_phase ph_grad, ph_gf, ph_div, ph_corr, ph_face, ph_oi, ph_filt;
  size_t smem_doubles;
  hf_phys P;
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// all products of one phase, tasks pooled over the warps of the CTA.  The k steps of every operator are padded to pairs on the host
// (zero fragments) and a task always works on two column blocks (the second one a copy of the first when the tile has an odd
// block left over, its result dropped), so the inner loop carries no predicate: 2 fragment loads, 4 data loads, 4 DMMA per
// iteration, the next iteration's fragments requested before this one's tensor-core instructions.
// one term of a task with a compile-time number of k-step pairs: fully unrolled, all operator fragments requested up front
template <int KB2>
__device__ __forceinline__ void term_unrolled(const double *__restrict__ opf, const double *s0, const double *s1, double &a00, double &a01, double &a10,
                                              double &a11)
{
  double b[2 * KB2];
#pragma unroll
  for (int q = 0; q < 2 * KB2; q++) b[q] = __ldg(opf + q * 32);
#pragma unroll
  for (int q = 0; q < 2 * KB2; q++)
  {
    dmma884(a00, a01, s0[q * 4], b[q]);
    dmma884(a10, a11, s1[q * 4], b[q]);
  }
}
__device__ __forceinline__ void run_phase(const el_phase &PH, double *sm)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int ar = lane >> 2, ak = lane & 3;
  for (int task = warp; task < PH.n; task += nw)
  {
    const el_task *K = PH.tasks + task;
    const int4 hd = __ldg(reinterpret_cast<const int4 *>(K)); // n_terms, mode, dst, ds
    const int nt = hd.x, mode = hd.y, ds = hd.w;
    const bool two = __ldg(&K->two) != 0;
    double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
    for (int t = 0; t < nt; t++)
    {
      const int4 td = __ldg(reinterpret_cast<const int4 *>(&K->t[t].kb2)); // kb2, src, ss, pad
      const int kb2 = td.x, ss = td.z;
      const double *opf = reinterpret_cast<const double *>(__ldg(reinterpret_cast<const unsigned long long *>(&K->t[t].op))) + lane;
      const double *s0 = sm + td.y + ar * ss + ak;
      const double *s1 = two ? s0 + 8 * ss : s0;
      // the operator sizes of P = 3 (and below) as straight-line code: quadrilaterals / triangles 2, tetrahedra 3 and 5, prisms 5 and 9
      if (kb2 == 2) term_unrolled<2>(opf, s0, s1, a00, a01, a10, a11);
      else if (kb2 == 3) term_unrolled<3>(opf, s0, s1, a00, a01, a10, a11);
      else if (kb2 == 5) term_unrolled<5>(opf, s0, s1, a00, a01, a10, a11);
      else if (kb2 == 1) term_unrolled<1>(opf, s0, s1, a00, a01, a10, a11);
      else if (kb2 < 0)
      {
        // sparse operator: only the non-zero fragments of this row block, each with its k step
        const int *kx = reinterpret_cast<const int *>(__ldg(reinterpret_cast<const unsigned long long *>(&K->t[t].kidx)));
        const int cnt = -kb2;
        double b0 = __ldg(opf);
        int k0 = __ldg(kx);
        for (int q = 0; q < cnt; q++)
        {
          double n0 = 0.0;
          int k1 = 0;
          if (q + 1 < cnt) { n0 = __ldg(opf + (q + 1) * 32); k1 = __ldg(kx + q + 1); }
          dmma884(a00, a01, s0[k0 * 4], b0);
          dmma884(a10, a11, s1[k0 * 4], b0);
          b0 = n0; k0 = k1;
        }
      }
      else
      {
        double b0 = __ldg(opf), b1 = __ldg(opf + 32);
        for (int q = 0; q < kb2; q++)
        {
          opf += 64;
          double n0 = 0.0, n1 = 0.0;
          if (q + 1 < kb2) { n0 = __ldg(opf); n1 = __ldg(opf + 32); }
          dmma884(a00, a01, s0[0], b0);
          dmma884(a10, a11, s1[0], b0);
          dmma884(a00, a01, s0[4], b1);
          dmma884(a10, a11, s1[4], b1);
          s0 += 8; s1 += 8;
          b0 = n0; b1 = n1;
        }
      }
    }
    double2 *d0 = reinterpret_cast<double2 *>(sm + hd.z + ar * ds + 2 * ak);
    double2 *d1 = d0 + 4 * ds; // 8 columns further, in double2 units
    if (mode == 0)
    {
      *d0 = make_double2(a00, a01);
      if (two) *d1 = make_double2(a10, a11);
    }
    else
    {
      const double sg = mode == 1 ? 1.0 : -1.0;
      double2 v = *d0;
      v.x += sg * a00; v.y += sg * a01;
      *d0 = v;
      if (two)
      {
        double2 w = *d1;
        w.x += sg * a10; w.y += sg * a11;
        *d1 = w;
      }
    }
  }
}

// 8-byte asynchronous copies global -> shared (LDGSTS): a tile's loads are all in flight at once -- with one ordinary load per thread
// and loop iteration the kernels ran at a seventh of the HBM rate (too few bytes in flight per SM)
__device__ __forceinline__ void cp_async8(double *dst, const double *src)
{
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
// [field][element of the tile][point] of a global array -> shared columns (field-major inside the tile: column = field * E + element)
__device__ __forceinline__ void cp_async16(double *dst, const double *src)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int NF>
__device__ __forceinline__ void load_cols(const el_args &A, double *dst, int stride, const double *__restrict__ g, int npt, unsigned mg, int e0, int ne)
{
  const int per_field = ne * npt;
  const size_t fstride = (size_t)npt * A.n_eles;
  if ((npt & 1) == 0 && (fstride & 1) == 0)
  {
    // even point count: pairs of points are 16-byte aligned on both sides; .cg keeps the streamed tile out of L1, which then holds
    // the operator fragments
    for (int r = 2 * threadIdx.x; r < per_field; r += 2 * blockDim.x)
    {
      const int el = (int)__umulhi((unsigned)r, mg), pt = r - el * npt;
      double *d = dst + el * stride + pt;
      const double *s = g + (size_t)npt * e0 + r;
#pragma unroll
      for (int k = 0; k < NF; k++) cp_async16(d + k * A.E * stride, s + fstride * k);
    }
    return;
  }
  for (int r = threadIdx.x; r < per_field; r += blockDim.x)
  {
    const int el = (int)__umulhi((unsigned)r, mg), pt = r - el * npt;
    double *d = dst + el * stride + pt;
    const double *s = g + (size_t)npt * e0 + r;
#pragma unroll
    for (int k = 0; k < NF; k++) cp_async8(d + k * A.E * stride, s + fstride * k);
  }
}
// one value per (element, point): detjac
__device__ __forceinline__ void load_pts(double *dst, const double *__restrict__ g, int n)
{
  for (int r = threadIdx.x; r < n; r += blockDim.x) cp_async8(dst + r, g + r);
}
template <int NF>
__device__ __forceinline__ void store_cols(const el_args &A, const double *src, int stride, double *__restrict__ g, int npt, unsigned mg, int e0, int ne)
{
  const int per_field = ne * npt;
  const size_t fstride = (size_t)npt * A.n_eles;
  for (int r = threadIdx.x; r < per_field; r += blockDim.x)
  {
    const int el = (int)__umulhi((unsigned)r, mg), pt = r - el * npt;
    const double *sp = src + el * stride + pt;
    double *d = g + (size_t)npt * e0 + r;
#pragma unroll
    for (int k = 0; k < NF; k++) d[fstride * k] = sp[k * A.E * stride];
  }
}

__device__ __forceinline__ void zero_smem(double *sm, size_t n)
{
  double2 *p = reinterpret_cast<double2 *>(sm);
  for (size_t i = threadIdx.x; i < n / 2; i += blockDim.x) p[i] = make_double2(0.0, 0.0);
}

// physical gradient from the reference-space one: (1 / detjac) * JGinv^T (reference src/eles.cpp:1955-2011)
template <int ND, int NF>
__device__ __forceinline__ void to_physical(const double *gr, const double *J, double inv_detjac, double *g)
{
#pragma unroll
  for (int k = 0; k < NF; k++)
#pragma unroll
    for (int d = 0; d < ND; d++)
    {
      double acc = 0.0;
#pragma unroll
      for (int l = 0; l < ND; l++) acc += (inv_detjac * gr[k + NF * l]) * J[l + ND * d];
      g[k + NF * d] = acc;
    }
}

template <int ND, int NF, int MINB>
__global__ void __launch_bounds__(EL_THREADS, MINB) k_elem_grad(const __grid_constant__ el_args A)
{
  extern __shared__ __align__(16) double sm[];
  const int e0 = blockIdx.x * A.E;
  const int ne = min(A.E, A.n_eles - e0);
  const int ncp = A.mb * 8;
  zero_smem(sm, A.smem_doubles);
  __syncthreads();
  load_cols<NF>(A, sm + A.o_u, A.SU, A.u_in, A.nu, A.mg_u, e0, ne);
  load_cols<NF>(A, sm + A.o_dl, A.SF, A.delu, A.nf, A.mg_f, e0, ne);
  cp_async_wait_all();
  __syncthreads();
  run_phase(A.ph_grad, sm);
  __syncthreads();
  run_phase(A.ph_gf, sm);
  __syncthreads();
  if (A.grad_from_global)
  {
    // physical gradient at the solution points (reference src/eles.cpp:1955-1986) for k_elem_resid: written once here instead of the whole
    // opp_4 / opp_5 product being repeated there
    const size_t NUP = (size_t)A.nu * A.n_eles;
    for (int i = threadIdx.x; i < ne * A.nu; i += blockDim.x)
    {
      const int el = (int)__umulhi((unsigned)i, A.mg_u), pt = i - el * A.nu;
      const size_t p = (size_t)A.nu * (e0 + el) + pt;
      double J[ND * ND], gr[NF * ND], g[NF * ND];
#pragma unroll
      for (int q = 0; q < ND * ND; q++) J[q] = A.JG_u[p * (ND * ND) + q];
      const double inv_detjac = 1.0 / A.detjac_u[p];
#pragma unroll
      for (int l = 0; l < ND; l++)
#pragma unroll
        for (int k = 0; k < NF; k++) gr[k + NF * l] = sm[A.o_g + (l * ncp + k * A.E + el) * A.SU + pt];
      to_physical<ND, NF>(gr, J, inv_detjac, g);
#pragma unroll
      for (int q = 0; q < NF * ND; q++) A.grad_out[p + q * NUP] = g[q];
    }
  }
  const size_t NFP = (size_t)A.nf * A.n_eles;
  for (int i = threadIdx.x; i < ne * A.nf; i += blockDim.x)
  {
    const int el = (int)__umulhi((unsigned)i, A.mg_f), fp = i - el * A.nf;
    const size_t p = (size_t)A.nf * (e0 + el) + fp;
    double J[ND * ND], gr[NF * ND], g[NF * ND];
#pragma unroll
    for (int q = 0; q < ND * ND; q++) J[q] = A.JG_f[p * (ND * ND) + q];
    const double inv_detjac = 1.0 / A.detjac_f[p];
#pragma unroll
    for (int l = 0; l < ND; l++)
#pragma unroll
      for (int k = 0; k < NF; k++) gr[k + NF * l] = sm[A.o_gf + (l * ncp + k * A.E + el) * A.SF + fp];
    to_physical<ND, NF>(gr, J, inv_detjac, g);
#pragma unroll
    for (int q = 0; q < NF * ND; q++) A.grad_fpts[p + q * NFP] = g[q];
  }
}

template <int ND, int NF, int MINB>
__global__ void __launch_bounds__(EL_THREADS, MINB) k_elem_resid(const __grid_constant__ el_args A)
{
  extern __shared__ __align__(16) double sm[];
  const int e0 = blockIdx.x * A.E;
  const int ne = min(A.E, A.n_eles - e0);
  const int ncp = A.mb * 8;
  const size_t NUP = (size_t)A.nu * A.n_eles;
  zero_smem(sm, A.smem_doubles);
  __syncthreads();
  load_cols<NF>(A, sm + A.o_u, A.SU, A.u_in, A.nu, A.mg_u, e0, ne);
  if (A.visc && !A.over_int)
  {
    if (A.grad_from_global)
    {
#pragma unroll
      for (int d = 0; d < ND; d++) load_cols<NF>(A, sm + A.o_g + d * ncp * A.SU, A.SU, A.grad_out + (size_t)d * NF * NUP, A.nu, A.mg_u, e0, ne);
    }
    else
      load_cols<NF>(A, sm + A.o_dl, A.SF, A.delu, A.nf, A.mg_f, e0, ne);
  }
  load_cols<NF>(A, sm + A.o_fc, A.SF, A.ntconf, A.nf, A.mg_f, e0, ne);
  load_pts(sm + A.o_dj, A.detjac_u + (size_t)A.nu * e0, ne * A.nu);
  cp_async_wait_all();
  __syncthreads();
  if (A.visc && !A.grad_from_global)
  {
    run_phase(A.ph_grad, sm);
    __syncthreads();
  }
  if (A.over_int)
  {
    // polynomial de-aliasing (eles::evaluate_invFlux_over_int, reference src/eles.cpp:1480-1545): solution to the cubature points, inviscid
    // flux there with the cubature-point metrics, L2 projection back onto the solution basis -> the flux planes o_g.  (The gradient of a
    // viscous run is then read from global memory in the point loop below: the planes are taken.)
    run_phase(A.ph_oi, sm); // plane 0 of o_cub = opp_over_int_cubpts u
    __syncthreads();
    for (int i = threadIdx.x; i < ne * A.n_cub; i += blockDim.x)
    {
      const int el = (int)__umulhi((unsigned)i, A.mg_c), cp = i - el * A.n_cub;
      const size_t p = (size_t)A.n_cub * (e0 + el) + cp;
      double uu[NF], J[ND * ND], f[NF * ND];
#pragma unroll
      for (int k = 0; k < NF; k++) uu[k] = sm[A.o_cub + (k * A.E + el) * A.SC + cp];
#pragma unroll
      for (int q = 0; q < ND * ND; q++) J[q] = A.JG_cub[p * (ND * ND) + q];
      inv_flux<ND, NF>(uu, f, A.P);
#pragma unroll
      for (int k = 0; k < NF; k++)
#pragma unroll
        for (int l = 0; l < ND; l++)
        {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < ND; m++) acc += J[l + ND * m] * f[k + NF * m];
          sm[A.o_cub + ((1 + l) * ncp + k * A.E + el) * A.SC + cp] = acc;
        }
    }
    __syncthreads();
    run_phase(A.ph_filt, sm); // o_g(d) = over_int_filter flux(d)
    __syncthreads();
  }
  // fluxes at the solution points, transformed, in place over the gradient
  for (int i = threadIdx.x; i < ne * A.nu; i += blockDim.x)
  {
    const int el = (int)__umulhi((unsigned)i, A.mg_u), pt = i - el * A.nu;
    const size_t p = (size_t)A.nu * (e0 + el) + pt;
    double uu[NF], J[ND * ND], f[NF * ND], t[NF * ND];
#pragma unroll
    for (int k = 0; k < NF; k++) uu[k] = sm[A.o_u + (k * A.E + el) * A.SU + pt];
#pragma unroll
    for (int q = 0; q < ND * ND; q++) J[q] = A.JG_u[p * (ND * ND) + q];
    if (A.inv_from_global)
    {
#pragma unroll
      for (int q = 0; q < NF * ND; q++) t[q] = A.tdisf_in[p + q * NUP];
    }
    else if (A.over_int)
    {
#pragma unroll
      for (int l = 0; l < ND; l++)
#pragma unroll
        for (int k = 0; k < NF; k++) t[k + NF * l] = sm[A.o_g + (l * ncp + k * A.E + el) * A.SU + pt];
    }
    else
    {
      inv_flux<ND, NF>(uu, f, A.P);
#pragma unroll
      for (int k = 0; k < NF; k++)
#pragma unroll
        for (int l = 0; l < ND; l++)
        {
          double acc = 0.0;
#pragma unroll
          for (int m = 0; m < ND; m++) acc += J[l + ND * m] * f[k + NF * m];
          t[k + NF * l] = acc;
        }
    }
    if (A.visc)
    {
      double gr[NF * ND], g[NF * ND];
      if (A.over_int)
      {
#pragma unroll
        for (int q = 0; q < NF * ND; q++) gr[q] = A.grad_out[p + q * NUP]; // physical gradient of k_elem_grad
      }
      else
      {
#pragma unroll
        for (int l = 0; l < ND; l++)
#pragma unroll
          for (int k = 0; k < NF; k++) gr[k + NF * l] = sm[A.o_g + (l * ncp + k * A.E + el) * A.SU + pt];
      }
      if (A.grad_from_global)
      {
#pragma unroll
        for (int q = 0; q < NF * ND; q++) g[q] = gr[q];
      }
      else
        to_physical<ND, NF>(gr, J, 1.0 / sm[A.o_dj + i], g);
      if (A.store_grad && !A.grad_from_global)
      {
#pragma unroll
        for (int q = 0; q < NF * ND; q++) A.grad_out[p + q * NUP] = g[q];
      }
      vis_flux<ND, NF>(uu, g, f, A.P);
#pragma unroll
      for (int k = 0; k < NF; k++)
#pragma unroll
        for (int l = 0; l < ND; l++)
        {
          double acc = t[k + NF * l];
#pragma unroll
          for (int m = 0; m < ND; m++) acc += J[l + ND * m] * f[k + NF * m];
          t[k + NF * l] = acc;
        }
    }
#pragma unroll
    for (int l = 0; l < ND; l++)
#pragma unroll
      for (int k = 0; k < NF; k++) sm[A.o_g + (l * ncp + k * A.E + el) * A.SU + pt] = t[k + NF * l];
  }
  __syncthreads();
  run_phase(A.ph_div, sm); // divergence -> o_dl; common minus own normal flux -> o_fc
  __syncthreads();
  // the flux planes are dead: the second RK register travels into the first of them while the correction product runs
  const bool need_u1 = A.rk_mode == 2 || (A.rk_mode == 1 && !A.rk_copy);
  if (need_u1) load_cols<NF>(A, sm + A.o_g, A.SU, A.u1, A.nu, A.mg_u, e0, ne);
  run_phase(A.ph_corr, sm); // + opp_3 (common - own)
  cp_async_wait_all();
  __syncthreads();
  // RK update (the arithmetic of k_rk_update, reference src/eles.cpp:1080-1265)
  {
    const int per_field = ne * A.nu;
    for (int r = threadIdx.x; r < per_field; r += blockDim.x)
    {
      const int el = (int)__umulhi((unsigned)r, A.mg_u), pt = r - el * A.nu;
      const size_t p = (size_t)A.nu * e0 + r;
      const double inv_dj = 1.0 / sm[A.o_dj + r];
      const double dtl = A.dt_local ? A.dt_local[e0 + el] : A.dt;
#pragma unroll
      for (int k = 0; k < NF; k++)
      {
        const size_t idx = p + NUP * k;
        const int so = (k * A.E + el);
        const double div = sm[A.o_dl + so * A.SF + pt];
        if (A.store_div) A.div_out[idx] = div;
        double u = sm[A.o_u + so * A.SU + pt];
        if (A.rk_copy) A.u1[idx] = u;
        const double res = div * inv_dj;
        if (res != res) *A.nan_flag = 1 + e0 + el;
        if (A.rk_mode == 0)
          u -= dtl / A.fac * res;
        else if (A.rk_mode == 1)
        {
          const double uo = A.rk_copy ? u : sm[A.o_g + so * A.SU + pt];
          u = A.c1 * u + A.c2 * uo + dtl / A.fac * (-res);
        }
        else
        {
          const double d = A.c1 * sm[A.o_g + so * A.SU + pt] + dtl * (-res);
          A.u1[idx] = d;
          u += A.c2 * d;
        }
        A.u0[idx] = u;
        sm[A.o_u + so * A.SU + pt] = u;
      }
    }
  }
  __syncthreads();
  run_phase(A.ph_face, sm); // opp_0 of the updated solution -> o_fc
  __syncthreads();
  store_cols<NF>(A, sm + A.o_fc, A.SF, A.disu_fpts, A.nf, A.mg_f, e0, ne);
}

// only the last phase: disu_fpts = opp_0 disu_upts(0) (first stage after an upload, shock capturing)
template <int ND, int NF>
__global__ void __launch_bounds__(EL_THREADS, 2) k_elem_face(const __grid_constant__ el_args A)
{
  extern __shared__ __align__(16) double sm[];
  const int e0 = blockIdx.x * A.E;
  const int ne = min(A.E, A.n_eles - e0);
  zero_smem(sm, A.smem_doubles);
  __syncthreads();
  load_cols<NF>(A, sm + A.o_u, A.SU, A.u_in, A.nu, A.mg_u, e0, ne);
  cp_async_wait_all();
  __syncthreads();
  run_phase(A.ph_face, sm);
  __syncthreads();
  store_cols<NF>(A, sm + A.o_fc, A.SF, A.disu_fpts, A.nf, A.mg_f, e0, ne);
}

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// operator (rows x cols, column-major) in B-fragment order of mma.m8n8k4: lane = 4 * (row in block) + (k in step)
int upload_op(hf_ctx *c, el_op &O, const double *op, int rows, int cols)
{
  const int RB = (rows + 7) / 8, KBr = (cols + 3) / 4, KB = (KBr + 1) & ~1; // dense: k steps padded to pairs (run_phase)
  auto fragment = [&](int rb, int kk, double *out) {
    bool any = false;
    for (int lane = 0; lane < 32; lane++)
    {
      const int row = rb * 8 + (lane >> 2), k = kk * 4 + (lane & 3);
      out[lane] = (row < rows && k < cols) ? op[(size_t)k * rows + row] : 0.0;
      any = any || out[lane] != 0.0;
    }
    return any;
  };
  double f[32];
  long long nnz = 0;
  for (int rb = 0; rb < RB; rb++)
    for (int kk = 0; kk < KBr; kk++) nnz += fragment(rb, kk, f) ? 1 : 0;
  O.sparse = nnz * 5 <= (long long)RB * KBr * 2 && !getenv("HF_ELEM_DENSE"); // at most 40 % of the fragments carry anything (hexahedra; the 16 x 16 operators of P = 3 quadrilaterals stay dense: straight-line code)
  O.kb = KB;
  std::vector<double> frag;
  std::vector<int> kidx;
  if (O.sparse)
  {
    O.off.assign(RB, 0);
    O.cnt.assign(RB, 0);
    for (int rb = 0; rb < RB; rb++)
    {
      O.off[rb] = (int)kidx.size();
      for (int kk = 0; kk < KBr; kk++)
        if (fragment(rb, kk, f))
        {
          frag.insert(frag.end(), f, f + 32);
          kidx.push_back(kk);
        }
      if ((int)kidx.size() == O.off[rb]) // an all-zero row block still needs one fragment (zeros)
      {
        for (int q = 0; q < 32; q++) f[q] = 0.0;
        frag.insert(frag.end(), f, f + 32);
        kidx.push_back(0);
      }
      O.cnt[rb] = (int)kidx.size() - O.off[rb];
    }
    if (hf_alloc_copy(c, &O.kidx, kidx.data(), kidx.size())) return 1;
  }
  else
  {
    frag.assign((size_t)RB * KB * 32, 0.0);
    for (int rb = 0; rb < RB; rb++)
      for (int kk = 0; kk < KBr; kk++) fragment(rb, kk, &frag[((size_t)rb * KB + kk) * 32]);
  }
  return hf_alloc_copy(c, &O.frag, frag.data(), frag.size());
}
} // namespace

struct hf_elem_type
{
  bool ready = false;
  el_op op0, op1[3], op2[3], op3, op4[3], op5[3], op6, op_oi, op_filt;
  bool fused_oi = false; // over-integration inside k_elem_resid (else the staged evaluate_invFlux_over_int feeds it)
  int n_cub = 0, SC = 0;
  int E = 0, mb = 0, Eg = 0, mbg = 0, SU = 0, SF = 0; // tile of k_elem_resid / k_elem_face (E, mb) and of k_elem_grad (Eg, mbg)
  size_t smem_resid = 0, smem_grad = 0;
  // task tables on the device: [0] k_elem_resid / k_elem_face, [1] k_elem_grad; phases grad, gf, div, corr, face
  el_phase ph[2][7]; // + over-integration: interpolation to the cubature points, projection back
};
struct hf_elem_state
{
  hf_elem_type t[HF_N_ELE_TYPES];
  bool attr_done = false;
};

static int build_phases(hf_ctx *c, const hf_eles_dev &e, struct hf_elem_type &T);

int hf_elem_on_upload(hf_ctx *c, hf_eles_dev &e, const hf_eles_desc *d)
{
  if (!c->ez) c->ez = new hf_elem_state();
  hf_elem_type &T = c->ez->t[d->ele_type];
  const int nu = e.n_upts, nf = e.n_fpts, nd = e.n_dims, NF = e.n_fields;
  const bool visc = c->prm.viscous != 0;
  auto up = [&](el_op *dst, const double *op, int rows, int cols) -> int { return upload_op(c, *dst, op, rows, cols); };
  if (up(&T.op0, d->opp_0, nf, nu) || up(&T.op3, d->opp_3, nu, nf)) return 1;
  for (int i = 0; i < nd; i++)
  {
    if (up(&T.op1[i], d->opp_1[i], nf, nu) || up(&T.op2[i], d->opp_2[i], nu, nu)) return 1;
    if (visc && (up(&T.op4[i], d->opp_4[i], nu, nu) || up(&T.op5[i], d->opp_5[i], nu, nf))) return 1;
  }
  if (visc && up(&T.op6, d->opp_6, nf, nu)) return 1;
  const int nup = round_up(nu, 8), nfp = round_up(nf, 8); // multiples of 8 also cover the k steps padded to pairs (8 points)
  T.SU = nup + 4;
  T.SF = (nfp > nup ? nfp : nup) + 4; // the divergence (solution points) is accumulated in a flux-point buffer
  if (c->prm.over_int && d->n_over_int_cubpts > 0 && !getenv("HF_ELEM_NO_OI"))
  {
    T.n_cub = d->n_over_int_cubpts;
    T.SC = round_up(T.n_cub, 8) + 4;
    T.fused_oi = true;
    if (T.fused_oi && (up(&T.op_oi, d->opp_over_int_cubpts, T.n_cub, nu) || up(&T.op_filt, d->over_int_filter, nu, T.n_cub))) return 1;
  }
  // elements per CTA: whole 8-column blocks with little padding, shared memory for three CTAs per SM if the element allows it
  auto bytes_resid = [&](int mb) { return sizeof(double) * ((size_t)mb * 8 * ((size_t)T.SU * (1 + nd) + 2 * (size_t)T.SF + (T.fused_oi ? (size_t)(1 + nd) * T.SC : 0)) + (size_t)((mb * 8 / NF * nu + 1) & ~1)); };
  auto bytes_grad = [&](int mb) { return sizeof(double) * (size_t)mb * 8 * ((size_t)T.SU * (1 + nd) + (size_t)T.SF * (1 + nd)); };
  // shared memory per CTA to aim for (measured, GDOF-stage/s at 48 / 72 / 100 kB: quadrilaterals P=3 32.1 / 30.4 / 30.1, triangles +
  // quadrilaterals 10.0 / 9.4 / -, tetrahedra + prisms 2.84 / 3.29 / 3.13); HF_ELEM_KB overrides
  const size_t budget = (getenv("HF_ELEM_KB") ? (size_t)atoi(getenv("HF_ELEM_KB")) : (nd == 2 ? 48 : 72)) * 1024;
  // tile = E elements.  Padding of the E * n_fields columns to whole 8-column blocks is paid in every tensor-core instruction, so it comes
  // first (measured on P = 4 hexahedra: one element per CTA, 5 of 8 columns used, 2.76 GDOF-stage/s; three elements, 15 of 16, 3.77 although
  // only one CTA fits an SM); among the well-filled tiles the largest one inside the shared-memory budget, else the smallest one.
  auto pick = [&](auto bytes, const char *force_env) {
    const char *force = getenv(force_env);
    if (force && atoi(force) > 0) return atoi(force);
    double best_eff = 0.0;
    for (int E = 1; E <= 64; E++)
    {
      const int mb = round_up(E * NF, 8) / 8;
      if (bytes(mb) > 200 * 1024) break;
      best_eff = std::max(best_eff, (double)(E * NF) / (mb * 8));
    }
    int in_budget = 0, smallest = 0;
    for (int E = 1; E <= 64; E++)
    {
      const int mb = round_up(E * NF, 8) / 8;
      const size_t b = bytes(mb);
      if (b > 200 * 1024) break;
      if ((double)(E * NF) / (mb * 8) < std::min(0.9, best_eff)) continue;
      if (!smallest) smallest = E;
      if (b <= budget) in_budget = E;
    }
    return in_budget ? in_budget : smallest;
  };
  T.E = pick(bytes_resid, "HF_ELEM_E");
  T.Eg = visc ? pick(bytes_grad, "HF_ELEM_EG") : T.E;
  if (T.E == 0 || T.Eg == 0) return 0; // element too large for a shared-memory tile: the staged kernels run
  T.mb = round_up(T.E * NF, 8) / 8;
  T.mbg = round_up(T.Eg * NF, 8) / 8;
  T.smem_resid = bytes_resid(T.mb);
  T.smem_grad = bytes_grad(T.mbg);
  if (build_phases(c, e, T)) return 1;
  T.ready = true;
  return 0;
}

void hf_elem_destroy(hf_ctx *c)
{
  delete c->ez;
  c->ez = nullptr;
}

const char *hf_elem_status(hf_ctx *c)
{
  if (!c->ez) return "blocked element kernels: not prepared (staged-only context)";
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
    if (c->eles[t].present && !c->ez->t[t].ready) return "blocked element kernels: an element type does not fit a shared-memory tile";
  if (!((c->prm.n_dims == 2 && c->prm.n_fields == 4) || (c->prm.n_dims == 3 && c->prm.n_fields == 5) || c->prm.n_fields == 1))
    return "blocked element kernels: unsupported (n_dims, n_fields)";
  return "available";
}
int hf_elem_available(hf_ctx *c)
{
  const bool off = getenv("HF_NO_ELEM") != nullptr; // read per call: tests switch it inside one process
  return !off && strcmp(hf_elem_status(c), "available") == 0;
}

struct el_layout
{
  int o_u, o_g, o_dl, o_fc, o_gf, o_dj, o_cub;
};
static el_layout layout_of(const hf_elem_type &T, int nd, int mb, int nu = 0, int NF = 1)
{
  el_layout L;
  const int ncp = mb * 8;
  L.o_u = 0;
  L.o_g = ncp * T.SU;
  L.o_dl = L.o_g + nd * ncp * T.SU;
  L.o_fc = L.o_dl + ncp * T.SF; // k_elem_resid
  L.o_gf = L.o_dl + ncp * T.SF; // k_elem_grad
  L.o_dj = L.o_fc + ncp * T.SF; // k_elem_resid: detjac at the tile's solution points
  L.o_cub = L.o_dj + ((mb * 8 / NF * nu + 1) & ~1); // k_elem_resid with over-integration: flux planes at the cubature points
  return L;
}

// products of a phase -> warp tasks (row block x pair of column blocks), uploaded once per element type
static int compile_phase(hf_ctx *c, const std::vector<el_prod> &prods, int mb, el_phase *out)
{
  std::vector<el_task> tasks;
  const int ng = (mb + 1) / 2;
  for (const el_prod &Q : prods)
    for (int rb = 0; rb < Q.rb; rb++)
      for (int g = 0; g < ng; g++)
      {
        el_task K;
        memset(&K, 0, sizeof(K));
        K.n_terms = Q.n_terms; K.mode = Q.mode; K.ds = Q.ds;
        K.two = (2 * g + 1 < mb) ? 1 : 0;
        K.dst = Q.dst + 2 * g * 8 * Q.ds + rb * 8;
        for (int t = 0; t < Q.n_terms; t++)
        {
          const el_op &O = *Q.t[t].op;
          if (O.sparse)
          {
            K.t[t].op = O.frag + (size_t)O.off[rb] * 32;
            K.t[t].kidx = O.kidx + O.off[rb];
            K.t[t].kb2 = -O.cnt[rb];
          }
          else
          {
            K.t[t].op = O.frag + (size_t)rb * O.kb * 32;
            K.t[t].kb2 = O.kb / 2;
          }
          K.t[t].src = Q.t[t].src + 2 * g * 8 * Q.t[t].ss;
          K.t[t].ss = Q.t[t].ss;
        }
        tasks.push_back(K);
      }
  out->n = (int)tasks.size();
  out->tasks = nullptr;
  if (tasks.empty()) return 0;
  el_task *d = nullptr;
  if (hf_alloc_copy(c, &d, tasks.data(), tasks.size())) return 1;
  out->tasks = d;
  return 0;
}

static int build_phases(hf_ctx *c, const hf_eles_dev &e, hf_elem_type &T)
{
  const int nd = e.n_dims;
  const bool visc = c->prm.viscous != 0;
  const int rbu = (e.n_upts + 7) / 8, rbf = (e.n_fpts + 7) / 8;
  for (int which = 0; which < 2; which++)
  {
    const int mb = which ? T.mbg : T.mb, ncp = mb * 8;
    const el_layout L = layout_of(T, nd, mb, e.n_upts, e.n_fields);
    std::vector<el_prod> grad, gf, div, corr, face, oi, filt;
    // reference-space gradient, corrected: opp_4(d) u + opp_5(d) delta
    for (int d = 0; d < nd && visc; d++)
    {
      el_prod Q;
      memset(&Q, 0, sizeof(Q));
      Q.n_terms = 2; Q.rb = rbu; Q.dst = L.o_g + d * ncp * T.SU; Q.ds = T.SU; Q.mode = 0;
      Q.t[0] = {&T.op4[d], L.o_u, T.SU};
      Q.t[1] = {&T.op5[d], L.o_dl, T.SF};
      grad.push_back(Q);
    }
    if (which == 1)
    {
      // gradient at the flux points: opp_6 g(d)
      for (int d = 0; d < nd && visc; d++)
      {
        el_prod Q;
        memset(&Q, 0, sizeof(Q));
        Q.n_terms = 1; Q.rb = rbf; Q.dst = L.o_gf + d * ncp * T.SF; Q.ds = T.SF; Q.mode = 0;
        Q.t[0] = {&T.op6, L.o_g + d * ncp * T.SU, T.SU};
        gf.push_back(Q);
      }
    }
    else
    {
      // divergence sum_d opp_2(d) f(d) -> o_dl;  o_fc -= sum_d opp_1(d) f(d)
      el_prod Q, R;
      memset(&Q, 0, sizeof(Q));
      memset(&R, 0, sizeof(R));
      Q.n_terms = nd; Q.rb = rbu; Q.dst = L.o_dl; Q.ds = T.SF; Q.mode = 0;
      R.n_terms = nd; R.rb = rbf; R.dst = L.o_fc; R.ds = T.SF; R.mode = 2;
      for (int d = 0; d < nd; d++)
      {
        Q.t[d] = {&T.op2[d], L.o_g + d * ncp * T.SU, T.SU};
        R.t[d] = {&T.op1[d], L.o_g + d * ncp * T.SU, T.SU};
      }
      div.push_back(R); // the longer tasks first
      div.push_back(Q);
      el_prod C;
      memset(&C, 0, sizeof(C));
      C.n_terms = 1; C.rb = rbu; C.dst = L.o_dl; C.ds = T.SF; C.mode = 1;
      C.t[0] = {&T.op3, L.o_fc, T.SF};
      corr.push_back(C);
      el_prod F;
      memset(&F, 0, sizeof(F));
      F.n_terms = 1; F.rb = rbf; F.dst = L.o_fc; F.ds = T.SF; F.mode = 0;
      F.t[0] = {&T.op0, L.o_u, T.SU};
      face.push_back(F);
      if (T.fused_oi)
      {
        el_prod O;
        memset(&O, 0, sizeof(O));
        O.n_terms = 1; O.rb = (T.n_cub + 7) / 8; O.dst = L.o_cub; O.ds = T.SC; O.mode = 0;
        O.t[0] = {&T.op_oi, L.o_u, T.SU};
        oi.push_back(O);
        for (int d = 0; d < nd; d++)
        {
          el_prod P;
          memset(&P, 0, sizeof(P));
          P.n_terms = 1; P.rb = rbu; P.dst = L.o_g + d * ncp * T.SU; P.ds = T.SU; P.mode = 0;
          P.t[0] = {&T.op_filt, L.o_cub + (1 + d) * ncp * T.SC, T.SC};
          filt.push_back(P);
        }
      }
    }
    if (compile_phase(c, grad, mb, &T.ph[which][0]) || compile_phase(c, gf, mb, &T.ph[which][1]) || compile_phase(c, div, mb, &T.ph[which][2]) ||
        compile_phase(c, corr, mb, &T.ph[which][3]) || compile_phase(c, face, mb, &T.ph[which][4]) || compile_phase(c, oi, mb, &T.ph[which][5]) ||
        compile_phase(c, filt, mb, &T.ph[which][6])) return 1;
  }
  return 0;
}

static void fill_args(hf_ctx *c, hf_eles_dev &e, const hf_elem_type &T, el_args &A, bool grad_kernel = false)
{
  memset(&A, 0, sizeof(A));
  const int E = grad_kernel ? T.Eg : T.E, mb = grad_kernel ? T.mbg : T.mb;
  const bool visc = c->prm.viscous != 0;
  A.n_eles = e.n_eles; A.nu = e.n_upts; A.nf = e.n_fpts; A.E = E; A.mb = mb;
  A.mg_u = (unsigned)((0x100000000ull + e.n_upts - 1) / e.n_upts);
  A.mg_f = (unsigned)((0x100000000ull + e.n_fpts - 1) / e.n_fpts);
  A.SU = T.SU; A.SF = T.SF;
  const el_layout L = layout_of(T, e.n_dims, mb, e.n_upts, e.n_fields);
  A.o_cub = L.o_cub;
  A.o_u = L.o_u; A.o_g = L.o_g; A.o_dl = L.o_dl; A.o_fc = L.o_fc; A.o_gf = L.o_gf; A.o_dj = L.o_dj;
  A.visc = visc;
  A.u_in = e.disu_upts[0]; A.u0 = e.disu_upts[0]; A.u1 = e.disu_upts[1];
  A.delu = e.delta_disu_fpts; A.ntconf = e.norm_tconf_fpts; A.disu_fpts = e.disu_fpts; A.grad_fpts = e.grad_disu_fpts;
  A.div_out = e.div_tconf_upts; A.grad_out = e.grad_disu_upts; A.tdisf_in = e.tdisf_upts;
  A.detjac_u = e.detjac_upts; A.JG_u = e.JGinv_upts; A.detjac_f = e.detjac_fpts; A.JG_f = e.JGinv_fpts;
  A.nan_flag = c->d_nan;
  A.P = c->phys;
  const int w = grad_kernel ? 1 : 0;
  A.ph_grad = T.ph[w][0]; A.ph_gf = T.ph[w][1]; A.ph_div = T.ph[w][2]; A.ph_corr = T.ph[w][3]; A.ph_face = T.ph[w][4];
  A.ph_oi = T.ph[w][5]; A.ph_filt = T.ph[w][6];
  A.n_cub = T.n_cub; A.SC = T.SC; A.JG_cub = e.JGinv_over_int;
  A.mg_c = T.n_cub ? (unsigned)((0x100000000ull + T.n_cub - 1) / T.n_cub) : 0;
}

#define EL_LAUNCH(KERNEL, smem)                                                                                          \
  do {                                                                                                                   \
    const unsigned grid = (unsigned)((e.n_eles + A.E - 1) / A.E);                                                        \
    if (nd == 3 && nfl == 5) KERNEL<3, 5><<<grid, EL_THREADS, smem, c->stream>>>(A);                                     \
    else if (nd == 2 && nfl == 4) KERNEL<2, 4><<<grid, EL_THREADS, smem, c->stream>>>(A);                                \
    else if (nd == 2 && nfl == 1) KERNEL<2, 1><<<grid, EL_THREADS, smem, c->stream>>>(A);                                \
    else KERNEL<3, 1><<<grid, EL_THREADS, smem, c->stream>>>(A);                                                         \
    c->launches++;                                                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                                                 \
    if (e_ != cudaSuccess) { hf_set_error(std::string("kernel launch (blocked element kernel): ") + cudaGetErrorString(e_)); return 1; } \
  } while (0)

static int set_attrs(hf_ctx *c)
{
  if (c->ez->attr_done) return 0;
  const int lim = 200 * 1024;
#define EL_ATTR(K) HF_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, lim))
  EL_ATTR((k_elem_grad<3, 5, 2>)); EL_ATTR((k_elem_grad<2, 4, 4>)); EL_ATTR((k_elem_grad<2, 1, 4>)); EL_ATTR((k_elem_grad<3, 1, 2>));
  EL_ATTR((k_elem_resid<3, 5, 2>)); EL_ATTR((k_elem_resid<2, 4, 2>)); EL_ATTR((k_elem_resid<2, 1, 2>)); EL_ATTR((k_elem_resid<3, 1, 2>));
  EL_ATTR((k_elem_resid<3, 5, 3>)); EL_ATTR((k_elem_resid<2, 4, 3>)); EL_ATTR((k_elem_resid<2, 1, 3>)); EL_ATTR((k_elem_resid<3, 1, 3>));
  EL_ATTR((k_elem_resid<3, 5, 4>)); EL_ATTR((k_elem_resid<2, 4, 4>)); EL_ATTR((k_elem_resid<2, 1, 4>)); EL_ATTR((k_elem_resid<3, 1, 4>));
  EL_ATTR((k_elem_face<3, 5>)); EL_ATTR((k_elem_face<2, 4>)); EL_ATTR((k_elem_face<2, 1>)); EL_ATTR((k_elem_face<3, 1>));
#undef EL_ATTR
  c->ez->attr_done = true;
  return 0;
}

// disu_fpts = opp_0 disu_upts(0) for every element type
int hf_elem_extrapolate(hf_ctx *c)
{
  if (set_attrs(c)) return 1;
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    hf_eles_dev &e = c->eles[t];
    if (!e.present) continue;
    const hf_elem_type &T = c->ez->t[t];
    const int nd = e.n_dims, nfl = e.n_fields;
    el_args A;
    fill_args(c, e, T, A);
    A.smem_doubles = T.smem_resid / sizeof(double);
    EL_LAUNCH(k_elem_face, T.smem_resid);
  }
  c->ufpts_valid = true;
  return 0;
}

// One RK stage: CalcResidual (reference src/solver.cpp:50-223) + AdvanceSolution for every element type.
int hf_elem_stage(hf_ctx *c, int rk_stage, double time, int keep_residual)
{
  HF_CUDA(cudaSetDevice(c->device));
  if (set_attrs(c)) return 1;
  const bool visc = c->prm.viscous != 0, par = c->nproc > 1;
  const int stage = rk_stage & 0xff;
  // the physical gradient at the solution points travels from k_elem_grad to k_elem_resid through grad_disu_upts (D doubles per DOF
  // written and read) instead of the opp_4 / opp_5 product being repeated; HF_ELEM_REGRAD=1 keeps the recomputation (measurement aid)
  static const bool keep_grad = getenv("HF_ELEM_REGRAD") == nullptr;
#define EACH_INT(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_int_inters_op(c, t, OP)) return 1
#define EACH_BDY(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_bdy_inters_op(c, t, OP, time)) return 1
#define EACH_MPI(OP) for (int t = 0; t < HF_N_INTER_TYPES; t++) if (hf_dev_mpi_inters_op(c, t, OP)) return 1
  const bool les = c->prm.LES != 0;
  // LES: filtered solution and Leonard tensors at the first stage of a step (reference src/solver.cpp:54-62); the SVV model replaces the
  // solution by the filtered one, so this precedes the face values
  if (les && c->prm.SGS_model >= 2 && stage == 0)
    for (int t = 0; t < HF_N_ELE_TYPES; t++)
      if (c->eles[t].present && hf_dev_eles_op(c, t, HF_CALC_SGS_TERMS)) return 1;
  if (!c->ufpts_valid && hf_elem_extrapolate(c)) return 1;
  if (par) EACH_MPI(2);
  // over-integration: inside k_elem_resid where the element's tile has room for the cubature-point planes (and the gradient comes from
  // k_elem_grad), else the staged evaluate_invFlux_over_int writes the de-aliased inviscid flux for it
  auto oi_fused = [&](int t) { return c->prm.over_int && !les && c->ez->t[t].fused_oi && (!visc || keep_grad); };
  if (c->prm.over_int)
    for (int t = 0; t < HF_N_ELE_TYPES; t++)
      if (c->eles[t].present && !oi_fused(t) && hf_dev_eles_op(c, t, HF_EVALUATE_INVFLUX_OVER_INT)) return 1;
  EACH_INT(HF_COMMON_INVFLUX);
  EACH_BDY(HF_COMMON_INVFLUX);
  if (par) { EACH_MPI(3); EACH_MPI(HF_COMMON_INVFLUX); }
  if (visc)
  {
    for (int t = 0; t < HF_N_ELE_TYPES; t++)
    {
      hf_eles_dev &e = c->eles[t];
      if (!e.present) continue;
      const hf_elem_type &T = c->ez->t[t];
      const int nd = e.n_dims, nfl = e.n_fields;
      el_args A;
      fill_args(c, e, T, A, true);
      A.smem_doubles = T.smem_grad / sizeof(double);
      if (keep_grad || les)
      {
        if (!e.grad_disu_upts && hf_alloc_zero(c, &e.grad_disu_upts, (size_t)e.n_upts * e.n_eles * e.n_fields * e.n_dims)) return 1;
        A.grad_out = e.grad_disu_upts;
        A.grad_from_global = 1;
      }
      {
        const unsigned grid = (unsigned)((e.n_eles + A.E - 1) / A.E);
        if (nd == 3 && nfl == 5) k_elem_grad<3, 5, 2><<<grid, EL_THREADS, T.smem_grad, c->stream>>>(A);
        else if (nd == 2 && nfl == 4) k_elem_grad<2, 4, 4><<<grid, EL_THREADS, T.smem_grad, c->stream>>>(A);
        else if (nd == 2 && nfl == 1) k_elem_grad<2, 1, 4><<<grid, EL_THREADS, T.smem_grad, c->stream>>>(A);
        else k_elem_grad<3, 1, 2><<<grid, EL_THREADS, T.smem_grad, c->stream>>>(A);
        c->launches++;
        cudaError_t e_ = cudaGetLastError();
        if (e_ != cudaSuccess) { hf_set_error(std::string("kernel launch (k_elem_grad): ") + cudaGetErrorString(e_)); return 1; }
      }
    }
    if (par) EACH_MPI(4);
    if (les)
    {
      // LES is a hybrid: the point fluxes with the sub-grid-scale models (eles::evaluate_viscFlux with calc_sgsf_upts) and the SGS flux
      // at the flux points (extrapolate_sgsFlux) stay the staged kernels -- they read the physical gradient k_elem_grad has just left in
      // grad_disu_upts and write the TOTAL transformed flux to tdisf_upts, which k_elem_resid then takes as it is
      for (int t = 0; t < HF_N_ELE_TYPES; t++)
      {
        if (!c->eles[t].present) continue;
        if (!c->prm.over_int && hf_dev_eles_op(c, t, HF_EVALUATE_INVFLUX)) return 1;
        if (hf_dev_eles_op(c, t, HF_EVALUATE_VISCFLUX) || hf_dev_eles_op(c, t, HF_EXTRAPOLATE_SGSFLUX)) return 1;
      }
      if (par) EACH_MPI(6);
    }
    EACH_INT(HF_COMMON_VISCFLUX);
    EACH_BDY(HF_COMMON_VISCFLUX);
    if (par) { EACH_MPI(5); if (les) EACH_MPI(7); EACH_MPI(HF_COMMON_VISCFLUX); }
  }
  for (int t = 0; t < HF_N_ELE_TYPES; t++)
  {
    hf_eles_dev &e = c->eles[t];
    if (!e.present) continue;
    const hf_elem_type &T = c->ez->t[t];
    const int nd = e.n_dims, nfl = e.n_fields;
    el_args A;
    fill_args(c, e, T, A);
    A.smem_doubles = T.smem_resid / sizeof(double);
    A.over_int = oi_fused(t) ? 1 : 0;
    A.inv_from_global = (c->prm.over_int && !A.over_int) ? 1 : 0;
    if (les)
    {
      A.visc = 0;            // tdisf_upts already holds inviscid + viscous + sub-grid-scale flux
      A.inv_from_global = 1;
      A.tdisf_in = e.tdisf_upts;
    }
    A.store_div = keep_residual ? 1 : 0;
    if (keep_residual && visc && c->want_gradient && hf_ensure_staged_buffers(c, e)) return 1; // grad_disu_upts for the integral diagnostics
    A.grad_out = e.grad_disu_upts;
    A.store_grad = (keep_residual && visc && c->want_gradient) ? 1 : 0;
    A.grad_from_global = (visc && keep_grad) ? 1 : 0;
    A.dt = c->prm.dt;
    A.dt_local = (c->prm.dt_type == 2) ? e.dt_local : nullptr;
    if (hf_rk_coeffs(c, stage, &A.rk_mode, &A.rk_copy, &A.fac, &A.c1, &A.c2)) return 1;
    {
      // three CTAs per SM at 80 registers (a few spills in the flux phase) against two at 112 - 128: chosen by measurement
      // (measured: 2-D meshes gain from four CTAs per SM at 64 registers -- quadrilaterals 33.0 -> 36.8, mixed 2-D 10.4 -> 10.9 GDOF-stage/s --
      // 3-D ones are bounded by their shared-memory tiles and stay at three)
      const int minb = getenv("HF_ELEM_MINB") ? atoi(getenv("HF_ELEM_MINB")) : (nd == 2 ? 4 : 3);
      const unsigned grid = (unsigned)((e.n_eles + A.E - 1) / A.E);
      const size_t smem = T.smem_resid;
#define EL_RESID(ND_, NF_)                                                                                 \
      do {                                                                                                 \
        if (minb >= 4) k_elem_resid<ND_, NF_, 4><<<grid, EL_THREADS, smem, c->stream>>>(A);                \
        else if (minb >= 3) k_elem_resid<ND_, NF_, 3><<<grid, EL_THREADS, smem, c->stream>>>(A);           \
        else k_elem_resid<ND_, NF_, 2><<<grid, EL_THREADS, smem, c->stream>>>(A);                          \
      } while (0)
      if (nd == 3 && nfl == 5) EL_RESID(3, 5);
      else if (nd == 2 && nfl == 4) EL_RESID(2, 4);
      else if (nd == 2 && nfl == 1) EL_RESID(2, 1);
      else EL_RESID(3, 1);
#undef EL_RESID
      c->launches++;
      cudaError_t e_ = cudaGetLastError();
      if (e_ != cudaSuccess) { hf_set_error(std::string("kernel launch (k_elem_resid): ") + cudaGetErrorString(e_)); return 1; }
    }
  }
  c->ufpts_valid = true;
  return 0;
}

extern "C" const char *hf_dev_elem_status(hf_ctx *c) { return hf_elem_status(c); }
