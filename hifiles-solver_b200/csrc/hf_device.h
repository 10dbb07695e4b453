// Internal declarations of the device layer (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <string>
#include <cstdint>
#include "../../include/hifiles_b200.h"
#include "hf_physics.cuh"

#define HF_CUDA(call)                                                                                     \
  do {                                                                                                    \
    cudaError_t e_ = (call);                                                                              \
    if (e_ != cudaSuccess) { hf_set_error(std::string(#call) + ": " + cudaGetErrorString(e_)); return 1; } \
  } while (0)

void hf_set_error(const std::string &msg);
const std::string &hf_get_error();

// ELLPACK form of a small operator: exact zeros dropped, columns ascending, so a row sum visits the non-zero
// terms in the same order as the reference's dense dgemm (reference src/funcs.cpp:49-124).
struct hf_ell
{
  int rows = 0, cols = 0, nnz = 0;
  double *val = nullptr; // [nnz][rows]
  int *col = nullptr;    // [nnz][rows]; padding entries have val 0 and repeat the row's last column
  double *dval = nullptr; // [cols][rows] with the zeros in place, kept for operators that are at least half full (tensor-core kernel of the fast mode)
};

// Per element-type view handed to kernels by value.
struct hf_ele_view
{
  int n_eles, n_upts, n_fpts, n_dims, n_fields;
  double *disu_fpts;        // (fpt,ele,field)
  double *norm_tconf_fpts;  // (fpt,ele,field)
  double *delta_disu_fpts;  // (fpt,ele,field)
  double *grad_disu_fpts;   // (fpt,ele,field,dim)
  double *sgsf_fpts;        // (fpt,ele,field,dim), LES only
  const double *disu_upts;  // (upt,ele,field): wall-model input points
  const double *tdA_fpts;   // (fpt,ele)
  const double *norm_fpts;  // (fpt,ele,dim)
};
struct hf_views
{
  hf_ele_view v[HF_N_ELE_TYPES];
};

struct hf_eles_dev
{
  bool present = false;
  int ele_type = 0, n_eles = 0, n_upts = 0, n_fpts = 0, n_dims = 0, n_fields = 0, order = 0, n_inters = 0;
  int n_fpts_per_inter[6] = {0, 0, 0, 0, 0, 0};
  int fpt_offset[7] = {0, 0, 0, 0, 0, 0, 0};
  hf_ell opp_0, opp_1[3], opp_2[3], opp_3, opp_4[3], opp_5[3], opp_6;
  // over-integration
  int n_cub = 0;
  hf_ell opp_over_int, over_int_filter;
  double *JGinv_over_int = nullptr, *u_cub = nullptr, *tdisf_cub = nullptr;
  // LES
  double *sgsf_upts = nullptr, *sgsf_fpts = nullptr, *wall_distance = nullptr, *Jacobian_fpts = nullptr;
  double ele_vol_factor = 0.;
  hf_ell filter_upts;
  double *disuf_upts = nullptr, *uu = nullptr, *ue = nullptr, *Lu = nullptr, *Le = nullptr;
  // volume cubature of the integral diagnostics (eles::CalcIntegralQuantities)
  double *disu_average_upts = nullptr;
  int n_average = 0;
  int n_vol_cub = 0;
  double *opp_vol_cub = nullptr, *w_vol_cub = nullptr, *detjac_vol_cub = nullptr, *iq_elem = nullptr;
  // shock capturing (dense, row-major access by mode)
  double *inv_vandermonde = nullptr, *exp_filter = nullptr, *sensor_w_top = nullptr, *sensor_w_all = nullptr, *sensor = nullptr;
  double *detjac_upts = nullptr, *JGinv_upts = nullptr, *detjac_fpts = nullptr, *JGinv_fpts = nullptr;
  double *tdA_fpts = nullptr, *norm_fpts = nullptr, *h_ref = nullptr, *dt_local = nullptr;
  double *disu_upts[2] = {nullptr, nullptr};
  double *div_tconf_upts = nullptr, *disu_fpts = nullptr, *disu_fpts_alt = nullptr, *tdisf_upts = nullptr;
  double *norm_tdisf_fpts = nullptr, *norm_tconf_fpts = nullptr, *delta_disu_fpts = nullptr;
  double *grad_disu_upts = nullptr, *grad_disu_fpts = nullptr;
  // tensor-product fast path (hex / quad with per-element constant metrics)
  bool tensor = false;   // element type has tensor-product operators
  bool affine = false;   // metrics constant inside every element (to rounding): fused kernels may be used
  double *op1d = nullptr;       // 1-D tables extracted from the dense operators, see hf_fused.cuh
  double *ele_metrics = nullptr; // per element: JGinv (ND*ND), 1/detJ, then per face tdA, per face JGinv? see hf_fused.cuh
  int *face_nbr = nullptr;      // (face,ele): flat flux-point base of the neighbouring face, or -1
  int8_t *face_info = nullptr;  // (face,ele): bits 0-2 rot tag, bit 3 = this element is the right side, bit 4 = boundary/partition
  int8_t *beta_sign = nullptr;  // (fpt,ele): +1 / -1 sign applied to ldg_beta at this flux point (from the LEFT normal)
  std::vector<double> h_op[16]; // host copies of dense operators (kept for table extraction)
  // host-side extracts made at upload time for the fused path (hf_fused.cu)
  std::vector<double> h_em;        // per element: JGinv(l,m) at point 0 (ND*ND), detjac
  std::vector<double> h_face_geo;  // per (ele, face): tdA, unit normal[3] at the face's first flux point
  std::vector<int8_t> h_own_sign;  // (fpt,ele): sign of ldg_beta if this element is the left side of the face
  std::vector<double> h_norm_fpts; // (fpt,ele,dim) kept for RoeM runs: the fused kernels give that solver the exact normal of every flux point
  double affine_defect = 0.;       // max relative variation of the metrics inside an element
  int *d_pos = nullptr;            // pos on the device
  double *d_stage = nullptr;       // staging buffer of permuted uploads / downloads
  size_t stage_n = 0;
  double *d_xfer = nullptr;        // landing buffer of two-phase uploads (hf_dev_upload_begin / _commit)
  size_t xfer_n = 0;
  bool xfer_pending = false;
  int xfer_which = 0;
  std::vector<int> pos;            // device element order: slot pos[e] holds host element e (empty = identity), hf_dev_set_element_order
};

struct hf_int_inters_dev
{
  int n_inters = 0, nf = 0;
  int *idx_l = nullptr, *idx_r = nullptr; // [nf*n_inters] flat (fpt + n_fpts*ele)
  int8_t *type_l = nullptr, *type_r = nullptr;
  // host copies for neighbour-table construction
  std::vector<int> h_ele_type_l, h_ele_l, h_loc_l, h_ele_type_r, h_ele_r, h_loc_r, h_rot;
};

struct hf_bdy_inters_dev
{
  int n_inters = 0, nf = 0;
  int *idx_l = nullptr;
  int8_t *type_l = nullptr;
  int *bc_id = nullptr;
  double *pos_fpts = nullptr;
  int *wm_upt = nullptr;     // wall model: flat input solution point per interface (-1: none)
  double *wm_dist = nullptr;
  std::vector<int> h_ele_type_l, h_ele_l, h_loc_l, h_bc_id;
};

struct hf_mpi_inters_dev
{
  int n_inters = 0, nf = 0;
  int *idx_l = nullptr;
  int8_t *type_l = nullptr;
  int *lut = nullptr; // [nf*n_inters] right-side flux point inside the received interface block
  std::vector<int> nb_rank, nb_count;
  double *out_disu = nullptr, *in_disu = nullptr;   // [inter][field][fpt]
  double *out_grad = nullptr, *in_grad = nullptr;   // [inter][dim][field][fpt]
  double *out_sgsf = nullptr, *in_sgsf = nullptr;   // [inter][dim][field][fpt], LES
  std::vector<int> h_ele_type_l, h_ele_l, h_loc_l, h_rot, h_gid;
};

struct hf_ctx
{
  int device = 0, rank = 0, nproc = 1;
  cudaStream_t stream = nullptr, comm_stream = nullptr, xfer_stream = nullptr;
  cudaEvent_t ev_xfer = nullptr, ev_xfer_free = nullptr;
  bool own_stream = false;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr, ev_t0 = nullptr, ev_t1 = nullptr;
  hf_params prm;
  hf_phys phys;
  bool have_params = false;
  int fused = 1;
  bool finalized = false;
  // NaN guard of the residual (reference src/eles.cpp:1781-1795 scans div_tconf_upts after every stage): the update kernels raise
  // a device flag (1 + element in device order), read back once per time step
  int *d_nan = nullptr;
  int *h_nan = nullptr; // pinned
  // stage timeline (measurement aid, HF_STAGE_TIMELINE=1; nsys is not available on the GPU boxes): events on the compute and the
  // communication stream around every launch / exchange of the last recorded RK stages, read back by hf_timeline_report
  bool tl_on = false;
  std::vector<cudaEvent_t> tl_ev; // [slot][12]
  int tl_stage = 0, tl_xmark = 7; // tl_xmark: which pair of marks the next exchange records (7/8 common flux, 9/10 face values)
  // completion counters of the halo exchanges, bumped on the communication stream behind every exchange ([0] face values, [1] common
  // flux): the generation-9 kernels wait on them inside the kernel (one launch per kernel instead of interior + halo range)
  unsigned *d_xflag = nullptr;
  unsigned x_posted[2] = {0, 0};
  bool nccl_reconciled = false; // hf_fused_after_nccl has run (needs both the communicator and the finalized setup, in either order)
  bool want_gradient = false; // integral diagnostics requested: the fused kernels also store grad_disu_upts when they keep the residual
  bool ufpts_valid = false; // disu_fpts holds opp_0 * current disu_upts(0) (fused path bookkeeping)
  hf_eles_dev eles[HF_N_ELE_TYPES];
  hf_int_inters_dev ints[HF_N_INTER_TYPES];
  hf_bdy_inters_dev bdys[HF_N_INTER_TYPES];
  hf_mpi_inters_dev mpis[HF_N_INTER_TYPES];
  hf_bc *bc_table = nullptr;
  int n_bc = 0;
  std::vector<hf_bc> h_bc;
  double *scratch = nullptr; // reductions
  size_t scratch_bytes = 0;
  long long launches = 0;
  void *nccl_comm = nullptr;
  bool halo_pending = false;
  struct hf_fused_state *fz = nullptr; // fused-path state (hf_fused.cu)
  struct hf_elem_state *ez = nullptr;  // blocked element kernels (hf_elem.cu)
  // per-launch timing of the dominant kernel (bench roofline): event pairs recorded around its launches
  bool ktimer_on = false;
  std::vector<cudaEvent_t> kt_ev; // pool, pairs
  size_t kt_used = 0;
  std::vector<void *> allocs;
};

hf_views hf_make_views(hf_ctx *c);
int hf_fused_available(hf_ctx *c);
int hf_fused_prepare(hf_ctx *c);
int hf_fused_stage(hf_ctx *c, int rk_stage, double time, int keep_residual, int do_update);
int hf_fused_extrapolate(hf_ctx *c);
int hf_fused_on_upload(hf_ctx *c, hf_eles_dev &e, const hf_eles_desc *d);
void hf_fused_destroy(hf_ctx *c);
int hf_fused_after_nccl(hf_ctx *c);
// blocked element kernels (hf_elem.cu): the fast mode of every mesh the sum-factorised hexahedron kernels do not take
int hf_elem_on_upload(hf_ctx *c, hf_eles_dev &e, const hf_eles_desc *d);
int hf_elem_available(hf_ctx *c);
const char *hf_elem_status(hf_ctx *c);
int hf_elem_stage(hf_ctx *c, int rk_stage, double time, int keep_residual);
int hf_elem_extrapolate(hf_ctx *c);
void hf_elem_destroy(hf_ctx *c);
int hf_ensure_staged_buffers(hf_ctx *c, hf_eles_dev &e);
int hf_rk_coeffs(hf_ctx *c, int stage, int *mode, int *copy, double *fac, double *c1, double *c2);
// halo exchange over NCCL (hf_halo.cu): buffers are [inter][...] with `per_inter` doubles per interface, the message
// to neighbour p is the contiguous slice of its nb_count interfaces (reference src/mpi_inters.cpp:244-255)
int hf_halo_post(hf_ctx *c, hf_mpi_inters_dev &I, const double *out, double *in, size_t per_inter, int which = -1);
int hf_halo_wait(hf_ctx *c);
int hf_halo_allreduce_min(hf_ctx *c, double *v);
void hf_halo_destroy(hf_ctx *c);
// bracket the dominant kernel: call before / after its launch (no-ops unless the kernel timer is on)
void hf_ktimer_begin(hf_ctx *c);
void hf_ktimer_end(hf_ctx *c);
void hf_tl_mark(hf_ctx *c, int what, bool comm); // what 0..11, see hf_timeline_report

template <typename T>
int hf_alloc(hf_ctx *c, T **p, size_t n)
{
  void *q = nullptr;
  cudaError_t e = cudaMalloc(&q, (n ? n : 1) * sizeof(T));
  if (e != cudaSuccess) { hf_set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e)); return 1; }
  c->allocs.push_back(q);
  *p = (T *)q;
  return 0;
}
template <typename T>
int hf_alloc_copy(hf_ctx *c, T **p, const T *src, size_t n)
{
  if (hf_alloc(c, p, n)) return 1;
  // the context's streams are non-blocking (not ordered against the legacy default stream): copy on the compute stream and
  // wait, so that whatever is launched next on it, or on the communication stream behind an event, sees the data
  if (n)
  {
    HF_CUDA(cudaMemcpyAsync(*p, src, n * sizeof(T), cudaMemcpyHostToDevice, c->stream));
    HF_CUDA(cudaStreamSynchronize(c->stream));
  }
  return 0;
}
template <typename T>
int hf_alloc_zero(hf_ctx *c, T **p, size_t n)
{
  if (hf_alloc(c, p, n)) return 1;
  HF_CUDA(cudaMemsetAsync(*p, 0, (n ? n : 1) * sizeof(T), c->stream));
  HF_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}
