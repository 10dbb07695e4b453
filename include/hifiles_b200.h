/* hifiles_b200.h -- C ABI of the B200 device layer for the HiFiLES per-RK-stage residual hot path.
 *
 * This is the drop-in seam.  In the reference the host classes reach the device through
 *   (1) hf_array<T>::{mv_cpu_gpu,cp_cpu_gpu,cp_gpu_cpu,get_ptr_gpu}   (reference include/hf_array.h:341-355,493-594)
 *   (2) the *_gpu_kernel_wrapper free functions and bespoke_SPMV      (reference include/cuda_kernels.h:30-120)
 *   (3) legacy cublasDgemm / cublasDaxpy                              (e.g. reference src/eles.cpp:1397,1800)
 * none of which is compiled any more (SURVEY.md "Five facts" #1).  The entry points below replace all three.
 * Plain pointers and sizes only; no C++ or torch types.  Every function returns 0 on success, non-zero on
 * failure with the text available from hf_dev_last_error(); the C++ shim maps non-zero to the reference's
 * FatalError behaviour (reference include/error.h:31-43).  All device work is enqueued on the context's
 * compute stream; only download / residual_norm / calc_dt / sync block the host.
 *
 * Array layouts are the reference's: column-major hf_array, first index fastest (Appendix A of SURVEY.md):
 *   disu_upts(upt,ele,field)  disu_fpts(fpt,ele,field)  grad_disu_*(pt,ele,field,dim)  tdisf_upts(upt,ele,field,dim)
 *   detjac_*(pt,ele)  JGinv_*(l,m,pt,ele)  tdA_fpts(fpt,ele)  norm_fpts(fpt,ele,dim)  opp_k(row,col)
 */
#ifndef HIFILES_B200_H
#define HIFILES_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hf_ctx hf_ctx;

#define HF_MAX_RK 16
#define HF_N_ELE_TYPES 5   /* 0 tri, 1 quad, 2 tet, 3 prism, 4 hex (reference include/global.h:44-51) */
#define HF_N_INTER_TYPES 3 /* 0 segment, 1 triangle, 2 quad face (reference src/inters.cpp:58-98) */

/* run_input fields the hot path reads (reference include/input.h; set in src/input.cpp:75-324,527-720). */
typedef struct hf_params
{
  int equation;               /* 0 Euler/Navier-Stokes, 1 advection-diffusion */
  int viscous;
  int n_dims;
  int n_fields;
  int riemann_solve_type;     /* 0 Rusanov, 1 Lax-Friedrichs, 2 RoeM, 3 HLLC (reference src/int_inters.cpp:187-214) */
  int vis_riemann_solve_type; /* 0 LDG */
  int adv_type;               /* 0 Euler, 1 SSP-RK24, 2 SSP-RK34, 3 RK45, 4 RK414 (reference src/eles.cpp:1080-1265) */
  int dt_type;                /* 0 fixed, 1 global CFL, 2 local CFL */
  int fix_vis;
  int order;
  double gamma, prandtl, mu_inf, rt_inf, c_sth;
  double ldg_beta, ldg_tau;
  double dt, CFL;
  double R_ref, T_ref_placeholder;
  double wave_speed[3], diff_coeff, lambda;
  int n_rk;
  double RK_a[HF_MAX_RK], RK_b[HF_MAX_RK];
  /* de-aliasing and shock capturing (reference src/input.cpp:248-262) */
  int over_int;               /* 1: inviscid flux through over-integration (eles::evaluate_invFlux_over_int) */
  int shock_cap;              /* 1: Persson sensor + exponential modal filter after every RK stage (eles::shock_capture) */
  int shock_det_field;        /* 0 density, 1 total energy */
  double s0;                  /* sensor threshold */
  /* large-eddy simulation (reference src/eles.cpp:2395-2646): sub-grid-scale flux added to the viscous flux */
  int LES;                    /* 1: on */
  int SGS_model;              /* 0 Smagorinsky (with wall damping), 1 WALE, 2 WALE-similarity, 3 SVV, 4 similarity */
  double C_s, Kappa, prandtl_t, filter_ratio;
  int wall_model;             /* 0 off, 1 Werner-Wengle, 2 compressible log law (reference src/wall_model_funcs.cpp:13-118) */
} hf_params;

/* One element type (mirror of class eles, reference include/eles.h; storage src/eles.cpp:100-213). */
typedef struct hf_eles_desc
{
  int ele_type, n_eles, n_upts_per_ele, n_fpts_per_ele, n_dims, n_fields, order, n_inters_per_ele;
  const int *n_fpts_per_inter; /* [n_inters_per_ele] */
  /* dense operators, column-major (reference src/eles.cpp:3074-3596); opp_4..6 may be NULL when !viscous */
  const double *opp_0;    /* [n_fpts x n_upts] */
  const double *opp_1[3]; /* [n_fpts x n_upts] per dim */
  const double *opp_2[3]; /* [n_upts x n_upts] per dim */
  const double *opp_3;    /* [n_upts x n_fpts] */
  const double *opp_4[3]; /* [n_upts x n_upts] per dim */
  const double *opp_5[3]; /* [n_upts x n_fpts] per dim */
  const double *opp_6;    /* [n_fpts x n_upts] */
  /* metrics (reference src/eles.cpp:4035-4393) */
  const double *detjac_upts; /* (upt,ele) */
  const double *JGinv_upts;  /* (l,m,upt,ele) */
  const double *detjac_fpts; /* (fpt,ele) */
  const double *JGinv_fpts;  /* (l,m,fpt,ele) */
  const double *tdA_fpts;    /* (fpt,ele) */
  const double *norm_fpts;   /* (fpt,ele,dim) */
  const double *h_ref;       /* (ele) or NULL; only for dt_type != 0 */
  const double *disu_upts0;  /* (upt,ele,field) initial solution, or NULL (zero) */
  /* over-integration (reference src/eles.cpp:1480-1545, <type>::set_over_int, eles::set_transforms_over_int_cubtps); NULL / 0 when off */
  int n_over_int_cubpts;
  const double *opp_over_int_cubpts;   /* [n_cubpts x n_upts] interpolation to the cubature points */
  const double *over_int_filter;       /* [n_upts x n_cubpts] L2 projection back onto the solution basis */
  const double *JGinv_over_int_cubpts; /* (l,m,cubpt,ele) */
  /* shock capturing (reference src/eles.cpp:2918-2959, <type>::shock_det_persson / set_exp_filter); NULL when off */
  const double *inv_vandermonde;       /* [n_upts x n_upts] nodal -> modal */
  const double *sensor_w_top;          /* [n_upts] squared norm of the modes of the highest-degree shell, 0 elsewhere */
  const double *sensor_w_all;          /* [n_upts] squared norm of every mode */
  const double *exp_filter;            /* [n_upts x n_upts] */
  /* LES (reference src/eles.cpp:2395-2646, 2817-2914); NULL / 0 when off */
  const double *wall_distance;         /* (upt,ele,dim) vector to the nearest no-slip wall point (1e20 without walls); Smagorinsky */
  const double *Jacobian_fpts;         /* (a,b,fpt,ele) dx_a/dxi_b at the flux points: takes the SGS flux back to physical space */
  double ele_vol_factor;               /* reference-element volume: element volume = detjac * factor (<type>::calc_ele_vol) */
  const double *filter_upts;           /* [n_upts x n_upts] LES test filter (SGS models 2, 3, 4), <type>::compute_filter_upts */
} hf_eles_desc;

/* Interior interfaces of one face type (mirror of int_inters::set_interior, reference src/int_inters.cpp:67-121).
 * The reference bakes double* tables; here the same getter arithmetic (reference src/eles.cpp:4638-4871) is
 * done on the device side from these per-interface integers. */
typedef struct hf_int_inters_desc
{
  int inter_type, n_inters, n_fpts_per_inter;
  const int *ele_type_l, *ele_l, *local_inter_l;
  const int *ele_type_r, *ele_r, *local_inter_r;
  const int *rot_tag;
} hf_int_inters_desc;

/* Boundary parameter table entry (mirror of class bc after non-dimensionalisation, reference include/bc.h,
 * src/input.cpp:329-525).  bc_flag follows enum BCFLAG (reference include/global.h:55-69). */
typedef struct hf_bc
{
  int bc_flag;
  double rho, velocity[3], p_static, T_static, p_total, T_total, mach, nx, ny, nz;
  int use_wm; /* 1: this wall boundary takes its viscous flux from the wall model (bc::use_wm) */
  int T_isentropic; /* 1: ramped characteristic inlet with T_ramp_coeff < 0, T_total = T_l (p_total / p_l)^((gamma-1)/gamma) */
} hf_bc;

/* Boundary interfaces of one face type (mirror of bdy_inters::set_boundary, reference src/bdy_inters.cpp:75-178). */
typedef struct hf_bdy_inters_desc
{
  int inter_type, n_inters, n_fpts_per_inter;
  const int *ele_type_l, *ele_l, *local_inter_l;
  const int *bc_id;         /* index into the hf_bc table, per interface */
  const double *pos_fpts;   /* (fpt,inter,dim) physical flux-point coordinates, or NULL */
  /* wall model (reference src/bdy_inters.cpp:149-162): per interface the element's solution point farthest from the face
   * (flat upt + n_upts*ele, -1 = no wall model on this interface) and its distance; NULL when no boundary uses it */
  const int *wm_upt;
  const double *wm_dist;
} hf_bdy_inters_desc;

/* Partition ("MPI") interfaces of one face type (mirror of mpi_inters::set_mpi + set_nout_proc, reference
 * src/mpi_inters.cpp:154-215; ordering per neighbour as built by match_mpifaces, reference src/geometry.cpp:1132-1251). */
typedef struct hf_mpi_inters_desc
{
  int inter_type, n_inters, n_fpts_per_inter;
  const int *ele_type_l, *ele_l, *local_inter_l, *rot_tag;
  int n_neighbours;
  const int *neighbour_rank;  /* [n_neighbours] */
  const int *neighbour_count; /* [n_neighbours] interfaces exchanged with that rank, contiguous slices */
  /* optional (may be NULL): global id of the element behind every interface.  In a single-domain run the element with the
   * lower id is the interface's left side and its normal decides the LDG switch (src/inters.cpp:566-581); with the ids the fused
   * kernels make the same choice on partition faces instead of letting either rank switch on its own normal */
  const int *ele_global_l;
} hf_mpi_inters_desc;

/* which-array selectors for hf_dev_download / hf_dev_upload */
enum hf_array_id
{
  HF_DISU_UPTS0 = 0, HF_DISU_UPTS1 = 1, HF_DIV_TCONF_UPTS = 2, HF_DISU_FPTS = 3, HF_TDISF_UPTS = 4,
  HF_NORM_TDISF_FPTS = 5, HF_NORM_TCONF_FPTS = 6, HF_DELTA_DISU_FPTS = 7, HF_GRAD_DISU_UPTS = 8,
  HF_GRAD_DISU_FPTS = 9, HF_SRC_UPTS = 10, HF_DT_LOCAL = 11, HF_SENSOR = 12, /* (ele) Persson sensor of the last shock_capture */
  HF_SGSF_UPTS = 13, HF_SGSF_FPTS = 14, /* (pt,ele,field,dim) sub-grid-scale flux, LES runs */
  HF_DISUF_UPTS = 15, HF_LU = 16, HF_LE = 17, /* filtered solution (upt,ele,field), Leonard tensors (upt,ele,3|6) and (upt,ele,dim) */
  HF_DISU_AVERAGE_UPTS = 18 /* (upt,ele,average field): running time averages, hf_dev_time_average */
};

/* element operations = the eles methods CalcResidual calls (reference src/solver.cpp:65-216) */
enum hf_eles_op
{
  HF_EXTRAPOLATE_SOLUTION = 0,       /* eles::extrapolate_solution          reference src/eles.cpp:1360 */
  HF_CALCULATE_GRADIENT = 1,         /* eles::calculate_gradient            reference src/eles.cpp:1823 */
  HF_EVALUATE_INVFLUX = 2,           /* eles::evaluate_invFlux              reference src/eles.cpp:1415 */
  HF_CORRECT_GRADIENT = 3,           /* eles::correct_gradient              reference src/eles.cpp:1890 */
  HF_EVALUATE_VISCFLUX = 4,          /* eles::evaluate_viscFlux             reference src/eles.cpp:2285 */
  HF_EXTRAPOLATE_TOTALFLUX = 5,      /* eles::extrapolate_totalFlux         reference src/eles.cpp:1549 */
  HF_CALCULATE_DIVERGENCE = 6,       /* eles::calculate_divergence          reference src/eles.cpp:1651 */
  HF_CALCULATE_CORRECTED_DIVERGENCE = 7, /* eles::calculate_corrected_divergence reference src/eles.cpp:1738 */
  HF_EVALUATE_INVFLUX_OVER_INT = 8,  /* eles::evaluate_invFlux_over_int     reference src/eles.cpp:1480 */
  HF_SHOCK_CAPTURE = 9,              /* eles::shock_capture                 reference src/eles.cpp:2918 */
  HF_EXTRAPOLATE_SGSFLUX = 10,       /* eles::extrapolate_sgsFlux           reference src/eles.cpp:2817 */
  HF_CALC_SGS_TERMS = 11             /* eles::calc_sgs_terms (first RK stage) reference src/eles.cpp:2058 */
};
enum hf_inters_op
{
  HF_COMMON_INVFLUX = 0, /* int_inters::calculate_common_invFlux / bdy_inters::evaluate_boundaryConditions_invFlux */
  HF_COMMON_VISCFLUX = 1 /* int_inters::calculate_common_viscFlux / bdy_inters::evaluate_boundaryConditions_viscFlux */
};

/* ---- life cycle -------------------------------------------------------------------------------------- */
int hf_dev_create(hf_ctx **out, int device, int rank, int nproc);
int hf_dev_destroy(hf_ctx *ctx);
const char *hf_dev_last_error(void);
/* Use an externally created CUDA stream (cudaStream_t passed as void*) as the compute stream. */
int hf_dev_set_stream(hf_ctx *ctx, void *cuda_stream);
/* Join an NCCL communicator: unique_id is the 128-byte ncclUniqueId produced on rank 0 by hf_dev_nccl_unique_id.
 * Collective over the ranks.  May be called before or after hf_dev_finalize_setup: whichever of the two comes second
 * runs the cross-rank agreement of the fused path (every rank or none uses the fused kernels; LDG owner of every
 * flux-point pair on partition faces), so both calls must be made by every rank in the same order. */
int hf_dev_nccl_unique_id(void *unique_id_128_bytes);
int hf_dev_nccl_init(hf_ctx *ctx, const void *unique_id_128_bytes);

/* ---- setup (replaces eles::mv_all_cpu_gpu, int_inters::mv_all_cpu_gpu, ...) ---------------------------- */
int hf_dev_set_params(hf_ctx *ctx, const hf_params *p);
/* Optional, before hf_dev_upload_eles of that type: device storage order of the elements, pos[e] = device slot of host
 * element e (a permutation).  Every per-element array is stored in that order and hf_dev_upload / hf_dev_download
 * translate, so the host keeps the reference's numbering (reference src/mesh.cpp:188-311).  The host mirror puts the
 * elements without a partition face first, so that the fused kernels launch contiguous ranges: interior elements
 * while the halo exchange is in flight, partition-adjacent ones after it. */
int hf_dev_set_element_order(hf_ctx *ctx, int ele_type, int n_eles, const int *pos);
int hf_dev_upload_eles(hf_ctx *ctx, const hf_eles_desc *d);
int hf_dev_upload_int_inters(hf_ctx *ctx, const hf_int_inters_desc *d);
int hf_dev_set_bc_table(hf_ctx *ctx, int n_bc, const hf_bc *table);
int hf_dev_upload_bdy_inters(hf_ctx *ctx, const hf_bdy_inters_desc *d);
int hf_dev_upload_mpi_inters(hf_ctx *ctx, const hf_mpi_inters_desc *d);
/* Call once after all uploads: builds element-face neighbour tables, tensor-product operator tables, buffers. */
int hf_dev_finalize_setup(hf_ctx *ctx);

/* ---- the hot path --------------------------------------------------------------------------------------- */
/* CalcResidual(in_file_num, in_rk_stage, FlowSol)  (reference src/solver.cpp:50-223): result in div_tconf_upts. */
int hf_dev_calc_residual(hf_ctx *ctx, int rk_stage, double time);
/* eles::AdvanceSolution(in_step, adv_type) for every element type (reference src/eles.cpp:1080-1265). */
int hf_dev_advance_solution(hf_ctx *ctx, int rk_stage);
/* CalcResidual + AdvanceSolution of one stage as fused kernels (no div_tconf_upts round trip unless keep_residual). */
int hf_dev_rk_stage(hf_ctx *ctx, int rk_stage, double time, int keep_residual);
/* n_steps full time steps (all RK stages each), fused path, no host synchronisation inside. */
int hf_dev_run_steps(hf_ctx *ctx, int n_steps, double time0);
/* The per-stage NaN scan of the reference (src/eles.cpp:1781-1795: "Residual is NaN" -> abort) for hosts that call hf_dev_rk_stage
 * themselves: the update kernels raise a device flag, this reads it (one synchronisation; all ranks agree) and fails with the reference's
 * message.  hf_dev_run_steps does it once per call, hf_dev_advance_solution after every stage. */
int hf_dev_check_residual(hf_ctx *ctx);
/* Individual methods, for per-operator parity tests and for hosts that keep the reference's call sequence. */
int hf_dev_eles_op(hf_ctx *ctx, int ele_type, int op);
int hf_dev_int_inters_op(hf_ctx *ctx, int inter_type, int op);
int hf_dev_bdy_inters_op(hf_ctx *ctx, int inter_type, int op, double time);
int hf_dev_mpi_inters_op(hf_ctx *ctx, int inter_type, int op);
/* calc_time_step (reference src/solver.cpp:484-549): returns the global dt (dt_type 1) or fills dt_local (2). */
int hf_dev_calc_dt(hf_ctx *ctx, double *dt_out);
int hf_dev_set_dt(hf_ctx *ctx, double dt);

/* ---- data movement (replaces hf_array::cp_gpu_cpu / cp_cpu_gpu and eles::cp_*_gpu_cpu) ------------------- */
int hf_dev_download(hf_ctx *ctx, int ele_type, int which, double *host, size_t n_doubles);
int hf_dev_upload(hf_ctx *ctx, int ele_type, int which, const double *host, size_t n_doubles);
/* Two-phase form of hf_dev_upload for hosts that hand over a new array while the device is still computing (the reference's
 * eles::cp_disu_upts_cpu_gpu, include/eles.h:410-445, is synchronous): _begin starts the host -> device copy on a transfer stream
 * and returns at once (host must be page-locked and stay untouched until _commit has been called and the next synchronising call
 * returned); _commit orders the compute stream behind the copy and moves the data into the array.  One pending upload per
 * element type. */
int hf_dev_upload_begin(hf_ctx *ctx, int ele_type, int which, const double *host, size_t n_doubles);
int hf_dev_upload_commit(hf_ctx *ctx, int ele_type);
/* eles::compute_res_upts summed over element types as output::CalcNormResidual does (reference
 * src/eles.cpp:5045-5074, src/output.cpp:2166-2248); out[n_fields]. Local to this rank (no collective). */
int hf_dev_residual_norm(hf_ctx *ctx, int norm_type, double *out);
/* Integral diagnostics = eles::CalcIntegralQuantities (reference src/eles.cpp:5485-5628; output::CalcIntegralQuantities
 * src/output.cpp:2017-2040): interpolation of the solution and of grad_disu_upts (as left by the last residual evaluation,
 * the reference's semantics) to the volume cubature points, quantity * weight * detjac summed over points and elements.
 * hf_dev_set_volume_cubature hands over opp_volume_cubpts (cubpt, upt), weight_volume_cubpts and vol_detjac_vol_cubpts
 * (cubpt, ele) of eles::set_opp_volume_cubpts / set_transforms_vol_cubpts and makes the fused kernels store the gradient
 * whenever they keep the residual.  kinds: 0 kineticenergy, 1 enstropy, 2 pressuredilatation, 3 straincolonproduct,
 * 4 devstraincolonproduct.  out[n_quantities] is ADDED to (one call per element type); local to this rank. */
#define HF_MAX_INTEGRAL_QUANTITIES 8
/* on != 0: stages that keep the residual also leave grad_disu_upts behind in fused mode (the staged kernels always do), as
 * the reference's arrays hold it after the last CalcResidual of a step (output::CopyGPUCPU -> cp_grad_disu_upts_gpu_cpu,
 * reference src/output.cpp:2535-2556): surface forces (src/eles.cpp:5772-5785) and the vorticity-type plot fields
 * (src/eles.cpp:3781-3817) read it like the integral diagnostics.  Set by hf_dev_set_volume_cubature too. */
int hf_dev_set_keep_gradient(hf_ctx *ctx, int on);
/* Running time averages at the solution points = eles::CalcTimeAverageQuantities (reference src/eles.cpp:5630-5702), called once
 * per time step: average <- a * average + b * current with a = (time - spinup - dt) / (time - spinup), b = dt / (time - spinup)
 * (a = 0, b = 1 while time == spinup).  kinds: 0 rho_average, 1 u_average, 2 v_average, 3 w_average, 4 e_average.
 * The array (HF_DISU_AVERAGE_UPTS) is created zeroed by the first call. */
int hf_dev_time_average(hf_ctx *ctx, int ele_type, int n_average_fields, const int *kinds, double time, double spinup_time);
int hf_dev_set_volume_cubature(hf_ctx *ctx, int ele_type, int n_cubpts, const double *opp_volume_cubpts, const double *weights, const double *vol_detjac);
/* Replaces the wall-distance vectors of an element type (hf_eles_desc.wall_distance: (upt,ele,dim), host element order) after the upload:
 * a partitioned run knows the no-slip wall points of the other ranks only once the communicator exists (the reference gathers them with
 * MPI_Allgather / MPI_Bcast during its geometry setup, src/geometry.cpp:768-880). */
int hf_dev_set_wall_distance(hf_ctx *ctx, int ele_type, const double *wall_distance, size_t n_doubles);
int hf_dev_integral_quantities(hf_ctx *ctx, int ele_type, int n_quantities, const int *kinds, double *out);
/* sum of v[n] over the ranks of the context's communicator, in place on every rank (no-op on one rank) */
int hf_dev_allreduce_sum(hf_ctx *ctx, double *v, int n);
/* max of v[n] over the ranks (MPI_Reduce(MPI_MAX) of the infinity-norm residual, reference src/output.cpp:2216-2221) */
int hf_dev_allreduce_max(hf_ctx *ctx, double *v, int n);
int hf_dev_sync(hf_ctx *ctx);
/* kernels launched by this context since creation (bench.py reports the delta as gpu_launches) */
long long hf_dev_launch_count(hf_ctx *ctx);
/* select kernel family: 0 = staged reference-order kernels for every element type, 1 = fused tensor-product
 * kernels where available (default 1). */
int hf_dev_set_mode(hf_ctx *ctx, int fused);
/* "available", or the reason the fused kernels cannot be used for this mesh / input (the staged kernels then run) */
const char *hf_dev_fused_status(hf_ctx *ctx);
/* which fused kernel pair runs: "generation 7 (one-sided LDG ...)" when |ldg_beta| = 0.5 (each flux-point pair has one
 * owner that evaluates the whole common flux), else "generation 6 ..." with the reason; "none" without fused kernels */
const char *hf_dev_fused_variant(hf_ctx *ctx);
/* "available" when the blocked element kernels (hf_elem.cu: two kernels per element type and stage around the interface kernels, any
 * element type, boundary faces, curved elements) serve hf_dev_rk_stage / hf_dev_run_steps in fast mode wherever the sum-factorised
 * hexahedron kernels do not; otherwise the reason why the staged kernels run. */
const char *hf_dev_elem_status(hf_ctx *ctx);
/* CUDA-event timing on the compute stream: start/stop bracket, elapsed in milliseconds. */
int hf_dev_timer_start(hf_ctx *ctx);
int hf_dev_timer_stop(hf_ctx *ctx, float *ms);
/* Per-launch CUDA-event timing of the dominant kernel (the fused residual kernel; in staged mode the operator
 * kernel): mode 1 = start collecting, 0 = stop, -1 = read. ms_total / n_launches cover the launches since mode 1. */
int hf_dev_kernel_timer(hf_ctx *ctx, int mode, double *ms_total, long long *n_launches);

#ifdef __cplusplus
}
#endif
#endif
