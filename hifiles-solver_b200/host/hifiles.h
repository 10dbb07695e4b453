// Host-side mirror of the reference's eles / inters / solution API for the per-RK-stage residual path.
// The classes keep the reference's public names and argument meaning (reference include/eles.h:57-456,
// include/int_inters.h:47-63, include/bdy_inters.h:53-74, include/solution.h:44-89, include/solver.h:34-87) but
// everything below CalcResidual runs on the device through the C ABI in include/hifiles_b200.h.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <stdexcept>
#include <sstream>
#include <fstream>
#include <iostream>
#include <cmath>
#include <cstdint>
#include "hf_array.h"
#include "../../include/hifiles_b200.h"

// ---------------------------------------------------------------------------------------------------------
// errors: the reference prints and exit(1)s (include/error.h:31-43).  The library throws; the HiFiLES driver
// and the C ABI catch and reproduce the message, so embedding hosts are not killed.
// ---------------------------------------------------------------------------------------------------------
struct hf_fatal : public std::runtime_error
{
  explicit hf_fatal(const std::string &m) : std::runtime_error(m) {}
};
#define FatalError(msg)                                                                        \
  do {                                                                                         \
    std::ostringstream hf_oss_;                                                                \
    hf_oss_ << "Fatal error '" << (msg) << "' at " << __FILE__ << ":" << __LINE__;             \
    throw hf_fatal(hf_oss_.str());                                                             \
  } while (0)

#define MAX_V_PER_F 4
#define MAX_F_PER_C 6
#define MAX_V_PER_C 27

enum CTYPE { TRI = 0, QUAD = 1, TET = 2, PRISM = 3, HEX = 4 };
enum BCFLAG
{
  SUB_IN_SIMP = 0, SUB_OUT_SIMP = 1, SUB_IN_CHAR = 2, SUB_OUT_CHAR = 3, SUP_IN = 4, SUP_OUT = 5, SLIP_WALL = 6,
  CYCLIC = 7, ISOTHERM_WALL = 8, ADIABAT_WALL = 9, CHAR = 10, SLIP_WALL_DUAL = 11, AD_WALL = 12
};

extern const double pi;
/*! directory holding data/JacobiGQ.bin etc.: $HIFILES_HOME if set (reference src/global.cpp:33), else the
 *  package's own data directory. */
std::string hifiles_data_dir();

// ---------------------------------------------------------------------------------------------------------
// bc + input (reference include/bc.h, include/input.h, include/param_reader.h)
// ---------------------------------------------------------------------------------------------------------
class bc
{
public:
  bc();
  void setup(const std::string &in_bc_name);
  int get_bc_flag() const { return bc_flag; }
  std::string get_bc_type() const;
  std::string get_bc_name() const { return bc_name; }
  int set_bc_flag(std::string &in_type);

  double mach, rho, nx, ny, nz;
  double p_total, T_total, p_ramp_coeff, T_ramp_coeff, p_total_old, T_total_old;
  double p_static, T_static;
  hf_array<double> velocity;
  int pressure_ramp, use_wm, type, mode, n_eddy;
  double vis_y, turb_1, turb_2;

private:
  std::string bc_name;
  int bc_flag;
};

class param_reader
{
public:
  explicit param_reader(const std::string &fileName);
  // same "first word of a line == key, first hit wins" semantics as reference include/param_reader.h:91-172
  template <typename T> void getScalarValue(const std::string &optName, T &opt, T defaultVal);
  template <typename T> void getScalarValue(const std::string &optName, T &opt);
  void getVectorValueOptional(const std::string &optName, hf_array<std::string> &opt);
  void getVectorValue(const std::string &optName, hf_array<double> &opt);

private:
  bool find(const std::string &optName, std::istringstream &rest);
  std::vector<std::string> lines;
};

class input
{
public:
  input();
  void setup(const char *fileNameC, int rank);
  void read_input_file(const std::string &fileName, int rank);
  void read_boundary_param();
  void setup_params(int rank);

  std::string fileNameS;
  // basic
  int equation, order, viscous, ic_form, test_case, n_steps, restart_flag, restart_iter, n_restart_files;
  std::string mesh_file, data_file_name;
  int plot_freq, restart_dump_freq, monitor_res_freq, monitor_cp_freq, calc_force, res_norm_type, error_norm_type;
  int res_norm_field, p_res, write_type, probe;
  double area_ref;
  hf_array<std::string> integral_quantities, diagnostic_fields, average_fields;
  int n_integral_quantities, n_diagnostic_fields, n_average_fields;
  double spinup_time = 0.; // start time of the running averages (set at the first step, reference src/HiFiLES.cpp:242-243)
  // solver
  int riemann_solve_type, vis_riemann_solve_type, adv_type, dt_type;
  double dt, CFL, ldg_tau, ldg_beta, time;
  int RANS, LES, SGS_model, filter_type, wall_model;
  double C_s, filter_ratio;
  // gas
  double gamma, prandtl, prandtl_t, S_gas, T_gas, R_gas, mu_gas;
  int fix_vis;
  double Mach_free_stream, L_free_stream, T_free_stream, rho_free_stream;
  // bc
  int pressure_ramp, ramp_counter;
  double dx_cyclic, dy_cyclic, dz_cyclic;
  std::vector<bc> bc_list;
  // ic
  double Mach_c_ic, nx_c_ic, ny_c_ic, nz_c_ic, T_c_ic, u_c_ic, v_c_ic, w_c_ic, p_c_ic, rho_c_ic, mu_c_ic, uvw_c_ic;
  int patch, patch_type;
  double Mv, ra, rb, xc, yc, patch_x, x_shock_ic;
  int over_int, over_int_order, shock_cap, shock_det, expf_order, expf_cutoff, shock_det_field;
  double s0, expf_fac;
  // element options
  int upts_type_tri, fpts_type_tri, vcjh_scheme_tri, sparse_tri;
  double c_tri;
  int upts_type_quad, vcjh_scheme_quad, sparse_quad;
  double eta_quad;
  int upts_type_hexa, vcjh_scheme_hexa, sparse_hexa;
  double eta_hexa;
  int upts_type_tet, fpts_type_tet, vcjh_scheme_tet, sparse_tet;
  double c_tet, eta_tet;
  int upts_type_pri_tri, upts_type_pri_1d, vcjh_scheme_pri_1d, sparse_pri;
  double eta_pri;
  // adv-diff
  hf_array<double> wave_speed;
  double diff_coeff, lambda;
  int forcing, perturb_ic;
  hf_array<double> x_coeffs, y_coeffs, z_coeffs;
  // RK
  hf_array<double> RK_a, RK_b, RK_c;
  // reference quantities
  double T_ref, L_ref, rho_ref, uvw_ref, p_ref, mu_ref, time_ref, R_ref, c_sth, mu_inf, rt_inf, Kappa;
  // B200 extension (ignored by the reference: unknown keys are skipped, include/param_reader.h:108-128)
  int device_fused; // 1 (default): fused tensor-product kernels where available; 0: staged kernels
};
extern input run_input;

// ---------------------------------------------------------------------------------------------------------
// math utilities used by operator setup (reference src/funcs.cpp)
// ---------------------------------------------------------------------------------------------------------
double eval_lagrange(double in_r, int in_mode, hf_array<double> &in_loc_pts);
double eval_d_lagrange(double in_r, int in_mode, hf_array<double> &in_loc_pts);
double eval_legendre(double in_r, int in_mode);
double eval_d_legendre(double in_r, int in_mode);
double eval_d_vcjh_1d(double in_r, int in_mode, int in_order, double in_eta);
double compute_eta(int vcjh_scheme, int order);
bool is_perfect_square(int in_a);
bool is_perfect_cube(int in_a);
void eval_isentropic_vortex(hf_array<double> &pos, double time, double &rho, double &vx, double &vy, double &vz, double &p, int n_dims);
void eval_sine_wave_single(hf_array<double> &pos, hf_array<double> &wave_speed, double diff_coeff, double time, double &rho, hf_array<double> &grad_rho, int n_dims);
void eval_sine_wave_group(hf_array<double> &pos, hf_array<double> &wave_speed, double diff_coeff, double time, double &rho, hf_array<double> &grad_rho, int n_dims);
void eval_sphere_wave(hf_array<double> &pos, hf_array<double> &wave_speed, double time, double &rho, int n_dims);
/*! 1-D Gauss (rule 0) / Gauss-Lobatto (rule 1) points and weights from the reference's binary tables
 *  (reference src/cubature_1d.cpp:50-85). */
void cubature_1d(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights);
/*! simplex point tables of the data directory: locs(point, coordinate); rule 0 = interior (with weights), 1 = alpha-optimised */
void cubature_tri(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights);
void cubature_tet(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights);
// Dubiner bases and small dense algebra (host/simplex_basis.cpp)
double eval_jacobi(double r, int alpha, int beta, int mode);
double eval_grad_jacobi(double r, int alpha, int beta, int mode);
double eval_dubiner_basis_2d(double r, double s, int mode, int order);
double eval_dr_dubiner_basis_2d(double r, double s, int mode, int order);
double eval_ds_dubiner_basis_2d(double r, double s, int mode, int order);
double eval_dubiner_basis_3d(double r, double s, double t, int mode, int order);
double eval_grad_dubiner_basis_3d(double r, double s, double t, int mode, int order, int component);
hf_array<double> mult_arrays(hf_array<double> &A, hf_array<double> &B);
hf_array<double> transpose_array(hf_array<double> &A);
hf_array<double> inv_array(hf_array<double> &in);
void get_opp_3_tri(hf_array<double> &opp_3, hf_array<double> &loc_upts_tri, hf_array<double> &loc_1d_fpts, hf_array<double> &vandermonde_tri,
                   hf_array<double> &inv_vandermonde_tri, int n_upts_per_tri, int order, double c_tri, int vcjh_scheme_tri);

// ---------------------------------------------------------------------------------------------------------
// mesh (reference include/mesh.h, src/mesh.cpp, src/mesh_reader.cpp)
// ---------------------------------------------------------------------------------------------------------
class mesh
{
public:
  mesh();
  int get_num_cells(int in_type) const;
  int get_max_n_spts(int in_type) const;
  void create_iv2ivg();
  void set_vertex_connectivity();
  void set_face_connectivity();
  int get_corner_vlist_face(int in_ic, int in_face, int *out_vlist) const;
  static int compare_faces(const int *vlist1, const int *vlist2, int num_v_per_f, int &rtag);
  /*! keep the cells with part[global cell] == rank, in ascending global id (reference src/mesh.cpp:188-311) */
  void apply_partition(const std::vector<int> &part, int rank);
  int get_corner_vlist(int in_ic, int *v) const;

  int n_dims, n_ele_dims, n_bdy;
  int num_verts_global, num_cells_global, num_verts, num_cells, num_inters, n_unmatched_inters;
  hf_array<int> c2v, c2n_v, ctype, ic2icg, iv2ivg, bc_id;
  hf_array<double> xv;
  std::vector<std::vector<int>> v2c;
  hf_array<int> f2c, f2v, f2nv, f2loc_f, c2f, rot_tag, unmatched_inters;
};

class mesh_reader
{
public:
  mesh_reader(const std::string &in_fileName, mesh *in_mesh);
  void partial_read_connectivity(int kstart, int in_num_cells);
  void read_vertices();
  void read_boundary();

private:
  void read_header_gambit();
  void read_header_gmsh();
  void partial_read_connectivity_gambit(int kstart, int in_num_cells);
  void partial_read_connectivity_gmsh(int kstart, int in_num_cells);
  void read_vertices_gambit();
  void read_vertices_gmsh();
  void read_boundary_gambit();
  void read_boundary_gmsh();
  std::string fname;
  int mesh_format;
  mesh *mesh_ptr;
  std::vector<std::string> gmsh_bc_names; // physical names
  std::map<int, int> gmsh_phys2bc;
  int gmsh_elements_block_start;
};

// ---------------------------------------------------------------------------------------------------------
// elements
// ---------------------------------------------------------------------------------------------------------
struct solution;

class eles
{
public:
  eles();
  virtual ~eles() {}

  void setup(int in_n_eles, int in_max_n_spts_per_ele);
  void set_ics(double &time);
  void set_h_ref();
  /*! running time averages (reference src/eles.cpp:5630-5702): one device call per time step */
  void CalcTimeAverageQuantities(double &time);
  void cp_disu_average_upts_gpu_cpu();
  /*! overlay a vortex or a uniform state on the initial / restarted solution (reference eles::set_patch, src/eles.cpp:535-652) */
  void set_patch();
  /*! volume cubature of the integral diagnostics (reference src/eles.cpp:3667-3687, 4599-4632, 5485-5628) */
  void set_opp_volume_cubpts();
  void set_transforms_vol_cubpts();
  void CalcIntegralQuantities(int n_integral_quantities, hf_array<double> &integral_quantities);
  /*! error against the analytic solution of the test case, integrated over the volume cubature (reference src/eles.cpp:5076-5290) */
  hf_array<double> compute_error(int in_norm_type, double &time);
  /*! surface forces on walls (reference src/eles.cpp:5704-5990 and the interface-cubature setup behind it) */
  void set_inters_cubpts_and_transforms();
  void compute_wall_forces(hf_array<double> &inv_force, hf_array<double> &vis_force, double &temp_cl, double &temp_cd, std::ofstream &coeff_file, bool write_forces);
  void set_rank(int in_rank) { rank = in_rank; }
  void set_device(hf_ctx *in_ctx) { ctx = in_ctx; }

  /*! upload everything the device path needs (replaces the dead eles::mv_all_cpu_gpu, reference src/eles.cpp:931) */
  void mv_all_cpu_gpu();
  void cp_disu_upts_cpu_gpu();
  void cp_disu_upts_gpu_cpu();
  void cp_div_tconf_upts_gpu_cpu();
  void cp_grad_disu_upts_gpu_cpu();
  void cp_src_upts_gpu_cpu();
  void cp_array_gpu_cpu(int which, hf_array<double> &dst);

  // the hot-path methods: each is one device call (reference src/eles.cpp:1360-2052, 2285, 1080)
  void extrapolate_solution();
  void calculate_gradient();
  void evaluate_invFlux();
  void evaluate_invFlux_over_int();
  void shock_capture();
  void extrapolate_sgsFlux();
  void calc_sgs_terms();
  void compute_filter_upts();
  /*! distance vector of every solution point to the nearest no-slip wall flux point, brute force (reference src/eles.cpp:2698-2813) */
  void calc_wall_distance(std::vector<hf_array<double>> &loc_noslip_bdy);
  void cp_sensor_gpu_cpu();
  void correct_gradient();
  void evaluate_viscFlux();
  void extrapolate_totalFlux();
  void calculate_divergence();
  void calculate_corrected_divergence();
  void AdvanceSolution(int in_step, int adv_type);
  double compute_res_upts(int in_norm_type, int in_field);

  int get_ele_type() const { return ele_type; }
  int get_n_eles() const { return n_eles; }
  int get_n_dims() const { return n_dims; }
  int get_n_fields() const { return n_fields; }
  int get_n_upts_per_ele() const { return n_upts_per_ele; }
  int get_n_fpts_per_ele() const { return n_fpts_per_ele; }
  int get_n_spts_per_ele(int in_ele) { return n_spts_per_ele(in_ele); }
  int get_n_inters_per_ele() const { return n_inters_per_ele; }
  int get_n_fpts_per_inter(int f) { return n_fpts_per_inter(f); }

  void set_shape(int in_max_n_spts_per_ele);
  void set_shape_node(int in_spt, int in_ele, hf_array<double> &in_pos);
  void set_bcid(int in_ele, int in_inter, int in_bcid) { bcid(in_ele, in_inter) = in_bcid; }
  void set_n_spts(int in_ele, int in_n_spts) { n_spts_per_ele(in_ele) = in_n_spts; }
  void set_ele2global_ele(int in_ele, int in_global_ele) { ele2global_ele(in_ele) = in_global_ele; }

  void set_opp_0(int in_sparse);
  void set_opp_1(int in_sparse);
  void set_opp_2(int in_sparse);
  void set_opp_3(int in_sparse);
  void set_opp_4(int in_sparse);
  void set_opp_5(int in_sparse);
  void set_opp_6(int in_sparse);

  void set_transforms();
  // modal bases, over-integration, shock capturing (host/eles_modal.cpp)
  struct modal_mode { int i, j, k; };
  void set_modes();
  double eval_modal_basis(int m, hf_array<double> &loc);
  double modal_norm(int m);
  bool mode_is_top(int m);
  void set_modal_vandermonde();
  void set_volume_cubpts(int in_order, hf_array<double> &locs, hf_array<double> &weights);
  void set_over_int();
  void set_transforms_over_int_cubpts();
  void set_shock_capture();
  void set_transforms_upts();
  void set_transforms_fpts();
  void calc_pos(hf_array<double> &in_loc, int in_ele, hf_array<double> &out_pos);
  void calc_d_pos(hf_array<double> &in_loc, int in_ele, hf_array<double> &out_d_pos);
  virtual double calc_h_ref_specific(int in_ele);

  /*! flux-point index inside the element from (local face, face-local flux point): the arithmetic of the
   *  reference's eleven pointer getters (reference src/eles.cpp:4638-4871) */
  int get_fpt_index(int in_inter_local_fpt, int in_ele_local_inter);

  virtual void setup_ele_type_specific() = 0;
  virtual double eval_nodal_basis(int in_index, hf_array<double> &in_loc) = 0;
  virtual double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) = 0;
  virtual void fill_opp_3(hf_array<double> &opp_3) = 0;
  virtual double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) = 0;
  virtual void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) = 0;

  // ---- data (public: tests and the oracle read them through the C API) ----
  hf_ctx *ctx;
  int rank, ele_type, n_eles, n_dims, n_fields, order, viscous, n_inters_per_ele;
  int n_upts_per_ele, n_fpts_per_ele, max_n_spts_per_ele, n_adv_levels, upts_type;
  hf_array<int> n_fpts_per_inter, n_spts_per_ele, ele2global_ele, bcid;
  hf_array<int> n_cubpts_per_inter;
  std::vector<hf_array<double>> loc_inters_cubpts, weight_inters_cubpts, opp_inters_cubpts, inter_detjac_inters_cubpts, norm_inters_cubpts; // per local face
  std::vector<int> bdy_ele2ele;
  hf_array<double> loc_volume_cubpts, weight_volume_cubpts, opp_volume_cubpts, vol_detjac_vol_cubpts; // (dim,cubpt), (cubpt), (cubpt,upt), (cubpt,ele)
  hf_array<double> loc_upts, tloc_fpts, tnorm_fpts, loc_1d_upts;
  hf_array<double> shape, d_nodal_s_basis;
  hf_array<double> opp_0, opp_3, opp_6;
  hf_array<hf_array<double>> opp_1, opp_2, opp_4, opp_5;
  hf_array<double> detjac_upts, JGinv_upts, detjac_fpts, JGinv_fpts, tdA_fpts, norm_fpts, pos_upts, pos_fpts, h_ref;
  hf_array<hf_array<double>> disu_upts, div_tconf_upts;
  hf_array<double> src_upts, grad_disu_upts, dt_local, disu_average_upts;
  std::vector<modal_mode> modes;
  hf_array<double> modal_vandermonde, modal_inv_vandermonde;
  hf_array<double> loc_over_int_cubpts, weight_over_int_cubpts, opp_over_int_cubpts, over_int_filter, JGinv_over_int_cubpts;
  hf_array<double> sensor_w_top, sensor_w_all, exp_filter, sensor;
  hf_array<double> wall_distance, Jacobian_fpts, filter_upts;
};

class eles_hexas : public eles
{
public:
  void setup_ele_type_specific() override;
  double eval_nodal_basis(int in_index, hf_array<double> &in_loc) override;
  double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) override;
  void fill_opp_3(hf_array<double> &opp_3) override;
  double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) override;
  void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) override;
  double eval_div_vcjh_basis(int in_index, hf_array<double> &loc);

private:
  void set_loc_upts();
  void set_tloc_fpts();
  void set_tnorm_fpts();
};

class eles_quads : public eles
{
public:
  void setup_ele_type_specific() override;
  double eval_nodal_basis(int in_index, hf_array<double> &in_loc) override;
  double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) override;
  void fill_opp_3(hf_array<double> &opp_3) override;
  double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) override;
  void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) override;
  double eval_div_vcjh_basis(int in_index, hf_array<double> &loc);

private:
  void set_loc_upts();
  void set_tloc_fpts();
  void set_tnorm_fpts();
};

/*! Triangles, tetrahedra, prisms: Dubiner modal basis, dense operators (host/eles_simplex.cpp) */
class eles_tris : public eles
{
public:
  void setup_ele_type_specific() override;
  double eval_nodal_basis(int in_index, hf_array<double> &in_loc) override;
  double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) override;
  void fill_opp_3(hf_array<double> &opp_3) override;
  double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) override;
  void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) override;
  double calc_h_ref_specific(int in_ele) override;
  hf_array<double> vandermonde, inv_vandermonde, loc_1d_fpts;
};

class eles_tets : public eles
{
public:
  void setup_ele_type_specific() override;
  double eval_nodal_basis(int in_index, hf_array<double> &in_loc) override;
  double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) override;
  void fill_opp_3(hf_array<double> &opp_3) override;
  double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) override;
  void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) override;
  double calc_h_ref_specific(int in_ele) override;
  hf_array<double> vandermonde, inv_vandermonde;

private:
  double eval_div_dg_tet(int in_index, hf_array<double> &loc, hf_array<double> &cub, hf_array<double> &cub_w);
};

class eles_pris : public eles
{
public:
  void setup_ele_type_specific() override;
  double eval_nodal_basis(int in_index, hf_array<double> &in_loc) override;
  double eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc) override;
  void fill_opp_3(hf_array<double> &opp_3) override;
  double eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts) override;
  void eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts) override;
  double calc_h_ref_specific(int in_ele) override;
  int n_upts_tri, n_upts_1d;
  hf_array<double> loc_upts_pri_1d, loc_upts_pri_tri, vandermonde_tri, inv_vandermonde_tri, loc_1d_fpts;

private:
  double tri_part(int index_tri, int cpnt, hf_array<double> &in_loc);
  int face0_map(int index);
};

// ---------------------------------------------------------------------------------------------------------
// interfaces: int32 connectivity instead of the reference's double* tables
// ---------------------------------------------------------------------------------------------------------
class inters
{
public:
  inters();
  void setup_inters(int in_n_inters, int in_inters_type);
  /*! flux-point permutation of the right side for a rotation tag (reference src/inters.cpp:153-262) */
  void get_lut(int in_rot_tag);
  void set_device(hf_ctx *in_ctx) { ctx = in_ctx; }
  int get_n_inters() const { return n_inters; }

  hf_ctx *ctx;
  int inters_type, order, viscous, n_inters, n_fpts_per_inter, n_fields, n_dims;
  hf_array<int> lut;
  hf_array<int> ele_type_l, ele_l, local_inter_l;
};

class int_inters : public inters
{
public:
  void setup(int in_n_inters, int in_inter_type);
  void set_interior(int in_inter, int in_ele_type_l, int in_ele_type_r, int in_ele_l, int in_ele_r, int in_local_inter_l, int in_local_inter_r, int rot_tag, struct solution *FlowSol);
  void mv_all_cpu_gpu();
  void calculate_common_invFlux();
  void calculate_common_viscFlux();
  hf_array<int> ele_type_r, ele_r, local_inter_r, rot_tags;
};

class bdy_inters : public inters
{
public:
  void setup(int in_n_inters, int in_inter_type);
  void set_boundary(int in_inter, int bc_id, int in_ele_type_l, int in_ele_l, int in_local_inter_l, struct solution *FlowSol);
  void mv_all_cpu_gpu();
  void evaluate_boundaryConditions_invFlux(struct solution *FlowSol, double time_bound);
  void evaluate_boundaryConditions_viscFlux(double time_bound);
  hf_array<int> boundary_id, wm_upt;
  hf_array<double> pos_fpts, wm_dist;
  bool any_wm = false;
};

class mpi_inters : public inters
{
public:
  void setup(int in_n_inters, int in_inter_type);
  void set_nproc(int in_nproc, int in_rank) { nproc = in_nproc; rank = in_rank; }
  void set_nout_proc(int in_nout, int in_p);
  void set_mpi(int in_inter, int in_ele_type_l, int in_ele_l, int in_local_inter_l, int rot_tag, struct solution *FlowSol);
  std::vector<int> ele_global_l; // global id of the element behind every interface
  void mv_all_cpu_gpu();
  void send_solution();
  void receive_solution();
  void send_corrected_gradient();
  void receive_corrected_gradient();
  void send_sgsf_fpts();
  void receive_sgsf_fpts();
  void calculate_common_invFlux();
  void calculate_common_viscFlux();
  int nproc, rank;
  hf_array<int> rot_tags;
  std::vector<int> neighbour_rank, neighbour_count;
};

// ---------------------------------------------------------------------------------------------------------
// solution aggregate + free functions (reference include/solution.h, include/solver.h, include/geometry.h)
// ---------------------------------------------------------------------------------------------------------
struct solution
{
  solution();
  ~solution();
  int rank, nproc;
  double time;
  int n_ele_types, n_dims, num_cells_global, ini_iter;
  hf_array<eles *> mesh_eles;
  eles_quads mesh_eles_quads;
  eles_tris mesh_eles_tris;
  eles_hexas mesh_eles_hexas;
  eles_tets mesh_eles_tets;
  eles_pris mesh_eles_pris;
  int n_int_inter_types, n_bdy_inter_types, n_mpi_inter_types, n_mpi_inters;
  std::vector<int_inters> mesh_int_inters;
  std::vector<bdy_inters> mesh_bdy_inters;
  std::vector<mpi_inters> mesh_mpi_inters;
  hf_array<double> norm_residual, integral_quantities, inv_force, vis_force;
  double coeff_lift = 0., coeff_drag = 0.;
  hf_ctx *ctx; // device context shared by all objects of this solution
  int no_device; // 1: host pre-processing only (CPU-side tests of setup logic); any hot-path call then fails loudly
  /*! optional partition vector (global cell -> rank); empty = block partition of the reference's initial read */
  std::vector<int> part;
  /*! Smagorinsky on several ranks: this rank's no-slip wall flux points per face type, kept until the communicator arrives
      (FinishWallDistance gathers the other ranks' and recomputes the wall distance; reference src/geometry.cpp:768-892) */
  std::vector<hf_array<double>> loc_noslip_bdy_local;
  bool wall_distance_pending = false;
};

void SetInput(struct solution *FlowSol);
void GeoPreprocess(struct solution *FlowSol, mesh &mesh_data);
/*! partitioned Smagorinsky run: wall points of every rank over the communicator, wall distance again, device arrays replaced */
void FinishWallDistance(struct solution *FlowSol);
void ReadMesh(struct solution *FlowSol, mesh &mesh_data);
void InitSolution(struct solution *FlowSol);
void CalcResidual(int in_file_num, int in_rk_stage, struct solution *FlowSol);
void calc_time_step(struct solution *FlowSol);
/*! output::CalcIntegralQuantities (reference src/output.cpp:2017-2040): fills FlowSol->integral_quantities (summed over ranks) */
void CalcIntegralQuantities(struct solution *FlowSol);
/*! output::CalcTimeAverageQuantities (reference src/output.cpp:2042-2053) */
void CalcTimeAverageQuantities(struct solution *FlowSol);
/*! output::compute_error (reference src/output.cpp:2052-2160): appends one line to error.dat */
void compute_error(int in_file_num, struct solution *FlowSol);
/*! output::write_vtu (reference src/output.cpp:462-900): Paraview file(s) of the current solution */
void write_vtu(int in_file_num, struct solution *FlowSol);
/*! output::write_tec (reference src/output.cpp:165-451): Tecplot ASCII file of the current solution */
void write_tec(int in_file_num, struct solution *FlowSol);
/*! the plot file of the format write_type selects (0 Paraview, 1 Tecplot; CGNS is not built) */
void write_plot(int in_file_num, struct solution *FlowSol);
/*! output::CalcForces (reference src/output.cpp:1915-2012): fills inv_force, vis_force, coeff_lift, coeff_drag; optionally the cp files */
void CalcForces(int in_file_num, bool write_forces, struct solution *FlowSol);
/*! CalcResidual + AdvanceSolution (+ shock_capture) of one RK stage: one fused device call where the fused kernels are
 *  available, the reference's sequence of methods otherwise.  monitored: the residual (and, for the integral diagnostics,
 *  the gradient) of this stage will be read afterwards. */
void AdvanceStage(int in_file_num, int in_rk_stage, struct solution *FlowSol, bool monitored);
/*! k-way partition of the mesh's dual graph (METIS), the serial counterpart of the reference's ParMETIS call (src/mesh.cpp:72-183) */
void partition_mesh_kway(const mesh &m, int n_dims, int nproc, std::vector<int> &part);
/*! read ASCII restart files Rest_<iter>_p<file>.dat (reference src/solver.cpp:377-434) */
void read_restart_ascii(int in_file_num, int in_n_files, struct solution *FlowSol);
/*! write ASCII restart file(s): Rest_<iter>_p0000.dat, or Rest_<iter>/Rest_<iter>_p<rank>.dat of a partitioned run (reference src/output.cpp:1753-1818) */
void write_restart_ascii(struct solution *FlowSol, int in_file_num);
/*! output::CalcNormResidual (reference src/output.cpp:2166-2248): fills FlowSol->norm_residual */
void CalcNormResidual(struct solution *FlowSol);
int get_n_rk_steps(int adv_type);
/*! B200 extension: all RK stages of n_steps steps on the device without host round trips */
void AdvanceSteps(struct solution *FlowSol, int n_steps);
void upload_bc_table(struct solution *FlowSol); // again every time step when an inlet ramps (pressure_ramp)
void hf_check(int status);
