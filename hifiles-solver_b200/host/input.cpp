// Input-file options and non-dimensionalisation, keeping the reference's option names and semantics so the
// reference's input files run unchanged (reference src/input.cpp:62-720, include/param_reader.h:91-172,
// src/bc.cpp:35-86).  Unknown keys are ignored, the first line whose first word equals the key wins.
#include "hifiles.h"
#include <algorithm>
#include <cstdlib>

using namespace std;

input run_input;
const double pi = 3.141592653589793;

// ---- bc ----------------------------------------------------------------------------------------------------
static const char *const k_bc_types[13] = {"sub_in_simp", "sub_out_simp", "sub_in_char", "sub_out_char", "sup_in",
                                           "sup_out", "slip_wall", "cyclic", "isotherm_wall", "adiabat_wall",
                                           "char", "slip_wall_dual", "ad_wall"};

bc::bc() : mach(0), rho(0), nx(1), ny(0), nz(0), p_total(0), T_total(0), p_ramp_coeff(0), T_ramp_coeff(0),
           p_total_old(0), T_total_old(0), p_static(0), T_static(0), pressure_ramp(0), use_wm(0), type(0), mode(0),
           n_eddy(0), vis_y(0), turb_1(0), turb_2(0), bc_flag(-1) {}
void bc::setup(const string &in_bc_name) { bc_name = in_bc_name; }
string bc::get_bc_type() const { return (bc_flag >= 0 && bc_flag < 13) ? k_bc_types[bc_flag] : "unset"; }
int bc::set_bc_flag(string &in_type)
{
  transform(in_type.begin(), in_type.end(), in_type.begin(), ::tolower);
  for (int i = 0; i < 13; i++)
    if (in_type == k_bc_types[i]) { bc_flag = i; return 0; }
  return -1;
}

// ---- param_reader ------------------------------------------------------------------------------------------
param_reader::param_reader(const string &fileName)
{
  ifstream f(fileName.c_str());
  if (!f.is_open()) FatalError("Cannont open input file for reading.");
  string s;
  while (getline(f, s)) lines.push_back(s);
}

bool param_reader::find(const string &optName, istringstream &rest)
{
  // a blank line leaves the previous key in place in the reference's loop (stringstream extraction fails and
  // optKey keeps its value); reproduce that so "first match" means the same line here
  string optKey;
  for (size_t i = 0; i < lines.size(); i++)
  {
    istringstream ss(lines[i]);
    ss >> optKey;
    if (optKey == optName)
    {
      string remainder;
      getline(ss, remainder);
      rest.clear();
      rest.str(remainder);
      return true;
    }
  }
  return false;
}

template <typename T>
void param_reader::getScalarValue(const string &optName, T &opt, T defaultVal)
{
  istringstream rest;
  if (find(optName, rest))
  {
    if (!(rest >> opt))
    {
      cout << "WARNING: Unable to assign value to option " << optName << endl;
      cout << "Using default value of " << defaultVal << " instead." << endl;
      opt = defaultVal;
    }
    return;
  }
  opt = defaultVal;
}

template <typename T>
void param_reader::getScalarValue(const string &optName, T &opt)
{
  istringstream rest;
  if (find(optName, rest))
  {
    if (!(rest >> opt))
    {
      cerr << "WARNING: Unable to assign value to option " << optName << endl;
      FatalError("Required option not set: " + optName);
    }
    return;
  }
  FatalError("Required option not found: " + optName);
}

void param_reader::getVectorValueOptional(const string &optName, hf_array<string> &opt)
{
  istringstream rest;
  opt.setup(0);
  if (!find(optName, rest)) return;
  int nVals;
  if (!(rest >> nVals)) return;
  hf_array<string> tmp(nVals);
  for (int i = 0; i < nVals; i++)
    if (!(rest >> tmp(i))) return;
  opt = tmp;
}

void param_reader::getVectorValue(const string &optName, hf_array<double> &opt)
{
  istringstream rest;
  if (!find(optName, rest)) FatalError("Required option not found: " + optName);
  int nVals;
  if (!(rest >> nVals)) FatalError("Required option not set: " + optName);
  opt.setup(nVals);
  for (int i = 0; i < nVals; i++)
    if (!(rest >> opt(i))) FatalError("Required option not set: " + optName);
}

// ---- input ---------------------------------------------------------------------------------------------------
input::input()
{
  time = 0.;
  monitor_cp_freq = 0;
  area_ref = 1.;
  pressure_ramp = 0;
  ramp_counter = 0;
  CFL = 0.;
  dt = 0.;
  ldg_tau = 0.;
  ldg_beta = 0.5;
  lambda = 0.;
  diff_coeff = 0.;
  device_fused = 1;
  u_c_ic = v_c_ic = w_c_ic = p_c_ic = 0.;
  Mach_c_ic = T_c_ic = mu_c_ic = uvw_c_ic = 0.;
  nx_c_ic = 1.; ny_c_ic = nz_c_ic = 0.;
  x_shock_ic = 0.;
  c_sth = mu_inf = rt_inf = 0.;
}

void input::setup(const char *fileNameC, int rank)
{
  fileNameS.assign(fileNameC);
  read_input_file(fileNameS, rank);
  // The reference opens mesh_file relative to the working directory.  Embedding hosts (tests, bench) do not chdir:
  // when that path does not exist, fall back to the directory of the input file.
  if (!mesh_file.empty() && mesh_file[0] != '/' && !std::ifstream(mesh_file.c_str()).good())
  {
    size_t slash = fileNameS.find_last_of('/');
    if (slash != string::npos)
    {
      string alt = fileNameS.substr(0, slash + 1) + mesh_file;
      if (std::ifstream(alt.c_str()).good()) mesh_file = alt;
    }
  }
  setup_params(rank);
}

void input::read_input_file(const string &fileName, int rank)
{
  param_reader opts(fileName);

  /* ---- Basic Simulation Parameters ---- */
  opts.getScalarValue("equation", equation);
  opts.getScalarValue("order", order);
  opts.getScalarValue("viscous", viscous);
  opts.getScalarValue("mesh_file", mesh_file);
  opts.getScalarValue("ic_form", ic_form, 1);
  opts.getScalarValue("test_case", test_case, 0);
  opts.getScalarValue("n_steps", n_steps);
  opts.getScalarValue("restart_flag", restart_flag, 0);
  if (restart_flag) // 0: new case; 1: ascii restart file; 2: hdf5 restart file
  {
    opts.getScalarValue("restart_iter", restart_iter);
    if (restart_flag == 1) opts.getScalarValue("n_restart_files", n_restart_files); // ascii files need to know number of files
    else if (restart_flag == 2) FatalError("To read HDF5 resart file, HiFiLES have to be compiled with HDF5");
  }

  /* ---- Monitoring ---- */
  opts.getScalarValue("plot_freq", plot_freq, INT32_MAX);
  opts.getScalarValue("data_file_name", data_file_name, string("Mesh"));
  opts.getScalarValue("restart_dump_freq", restart_dump_freq, INT32_MAX);
  opts.getScalarValue("monitor_res_freq", monitor_res_freq, 100);
  opts.getScalarValue("calc_force", calc_force, 0);
  if (calc_force)
  {
    opts.getScalarValue("monitor_cp_freq", monitor_cp_freq);
    opts.getScalarValue("area_ref", area_ref);
  }
  opts.getScalarValue("res_norm_type", res_norm_type, 2);
  opts.getScalarValue("error_norm_type", error_norm_type, 2);
  opts.getScalarValue("res_norm_field", res_norm_field, 0);
  opts.getScalarValue("p_res", p_res, 2);
  opts.getScalarValue("write_type", write_type, 0);
  opts.getScalarValue("probe", probe, 0);
  if (probe) FatalError("probes (probe_input.cpp) are not built in this host mirror");
  opts.getVectorValueOptional("integral_quantities", integral_quantities);
  opts.getVectorValueOptional("diagnostic_fields", diagnostic_fields);
  opts.getVectorValueOptional("average_fields", average_fields);
  n_integral_quantities = integral_quantities.get_dim(0);
  for (int i = 0; i < n_integral_quantities; i++) // lower case, as the reference (src/input.cpp:120-124)
    std::transform(integral_quantities(i).begin(), integral_quantities(i).end(), integral_quantities(i).begin(), ::tolower);
  if (n_integral_quantities && !viscous) FatalError("integral quantities read the solution gradient: viscous run needed");
  n_diagnostic_fields = diagnostic_fields.get_dim(0);
  for (int i = 0; i < n_diagnostic_fields; i++)
    std::transform(diagnostic_fields(i).begin(), diagnostic_fields(i).end(), diagnostic_fields(i).begin(), ::tolower);
  n_average_fields = average_fields.get_dim(0);
  for (int i = 0; i < n_average_fields; i++)
    std::transform(average_fields(i).begin(), average_fields(i).end(), average_fields(i).begin(), ::tolower);

  /* ---- Basic Solver Parameters ---- */
  opts.getScalarValue("riemann_solve_type", riemann_solve_type);
  opts.getScalarValue("vis_riemann_solve_type", vis_riemann_solve_type, 0);
  opts.getScalarValue("adv_type", adv_type);
  opts.getScalarValue("dt_type", dt_type);
  if (dt_type == 0)
    opts.getScalarValue("dt", dt);
  else
    opts.getScalarValue("CFL", CFL);
  if (vis_riemann_solve_type == 0)
  {
    opts.getScalarValue("ldg_tau", ldg_tau, 0.);
    opts.getScalarValue("ldg_beta", ldg_beta, 0.5);
  }

  /* ---- Turbulence Modeling Parameters ---- */
  opts.getScalarValue("RANS", RANS, 0);
  opts.getScalarValue("LES", LES, 0);
  SGS_model = 0; filter_type = 0; C_s = 0.; filter_ratio = 0.;
  if (LES)
  {
    opts.getScalarValue("C_s", C_s);
    opts.getScalarValue("SGS_model", SGS_model);
    if (SGS_model == 3 || SGS_model == 2 || SGS_model == 4)
      opts.getScalarValue("filter_type", filter_type);
    opts.getScalarValue("filter_ratio", filter_ratio);
  }
  opts.getScalarValue("wall_model", wall_model, 0);

  /* ---- Gas Parameters ---- */
  opts.getScalarValue("gamma", gamma, 1.4);
  opts.getScalarValue("prandtl", prandtl, .72);
  opts.getScalarValue("prandtl_t", prandtl_t, 0.9);
  opts.getScalarValue("S_gas", S_gas, 120.);
  opts.getScalarValue("T_gas", T_gas, 291.15);
  opts.getScalarValue("R_gas", R_gas, 286.9);
  opts.getScalarValue("mu_gas", mu_gas, 1.827E-5);
  opts.getScalarValue("fix_vis", fix_vis, 1);

  opts.getScalarValue("Mach_free_stream", Mach_free_stream, 1.);
  opts.getScalarValue("L_free_stream", L_free_stream, 1.);
  opts.getScalarValue("T_free_stream", T_free_stream, 300.);
  opts.getScalarValue("rho_free_stream", rho_free_stream, 1.17723946);

  /* ---- Boundary Conditions ---- */
  pressure_ramp = 0;
  opts.getScalarValue("dx_cyclic", dx_cyclic, (double)INFINITY);
  opts.getScalarValue("dy_cyclic", dy_cyclic, (double)INFINITY);
  opts.getScalarValue("dz_cyclic", dz_cyclic, (double)INFINITY);

  /* ---- Initial Conditions ---- */
  if (equation == 0)
  {
    if (viscous)
    {
      opts.getScalarValue("Mach_c_ic", Mach_c_ic);
      opts.getScalarValue("nx_c_ic", nx_c_ic, 1.);
      opts.getScalarValue("ny_c_ic", ny_c_ic, 0.);
      opts.getScalarValue("nz_c_ic", nz_c_ic, 0.);
      opts.getScalarValue("T_c_ic", T_c_ic);
    }
    else
    {
      opts.getScalarValue("u_c_ic", u_c_ic);
      opts.getScalarValue("v_c_ic", v_c_ic);
      opts.getScalarValue("w_c_ic", w_c_ic);
      opts.getScalarValue("p_c_ic", p_c_ic);
    }
  }
  opts.getScalarValue("rho_c_ic", rho_c_ic);

  opts.getScalarValue("patch", patch, 0);
  if (patch)
  {
    opts.getScalarValue("patch_type", patch_type, 0); // 0: vortex, 1: uniform flow
    if (patch_type == 0)
    {
      opts.getScalarValue("Mv", Mv, 0.5);
      opts.getScalarValue("ra", ra, 0.075);
      opts.getScalarValue("rb", rb, 0.175);
      opts.getScalarValue("xc", xc, 0.25);
      opts.getScalarValue("yc", yc, 0.5);
    }
    else if (patch_type == 1) // uniform patch for x > patch_x with the initial-condition state
      opts.getScalarValue("patch_x", patch_x);
  }

  if (ic_form == 9 || ic_form == 10)
    opts.getScalarValue("x_shock_ic", x_shock_ic);

  /* ---- Shock Capturing / dealiasing ---- */
  opts.getScalarValue("over_int", over_int, 0);
  if (over_int)
    opts.getScalarValue("over_int_order", over_int_order);
  opts.getScalarValue("shock_cap", shock_cap, 0);
  if (shock_cap)
  {
    opts.getScalarValue("shock_det", shock_det, 0);
    opts.getScalarValue("s0", s0);
    if (shock_cap == 1)
    {
      opts.getScalarValue("expf_fac", expf_fac, 36.0);
      opts.getScalarValue("expf_order", expf_order, 4);
      opts.getScalarValue("expf_cutoff", expf_cutoff, 0);
      opts.getScalarValue("shock_det_field", shock_det_field, 0);
    }
    else
      FatalError("Shock capturing method not implemented!");
  }

  /* ---- Element parameters ---- */
  opts.getScalarValue("upts_type_tri", upts_type_tri, 0);
  // the reference looks up the key " fpts_type_tri" (leading blank, src/input.cpp:271), which can never equal
  // the first word of a line: fpts_type_tri is therefore always its default.  Keep that behaviour.
  fpts_type_tri = 0;
  opts.getScalarValue("vcjh_scheme_tri", vcjh_scheme_tri, 0);
  opts.getScalarValue("c_tri", c_tri, 0.);
  opts.getScalarValue("sparse_tri", sparse_tri, 0);
  opts.getScalarValue("upts_type_quad", upts_type_quad, 0);
  opts.getScalarValue("vcjh_scheme_quad", vcjh_scheme_quad, 0);
  opts.getScalarValue("eta_quad", eta_quad, 0.);
  opts.getScalarValue("sparse_quad", sparse_quad, 0);
  opts.getScalarValue("upts_type_hexa", upts_type_hexa, 0);
  opts.getScalarValue("vcjh_scheme_hexa", vcjh_scheme_hexa, 0);
  opts.getScalarValue("eta_hexa", eta_hexa, 0.);
  opts.getScalarValue("sparse_hexa", sparse_hexa, 0);
  opts.getScalarValue("upts_type_tet", upts_type_tet, 0);
  opts.getScalarValue("fpts_type_tet", fpts_type_tet, 0);
  opts.getScalarValue("vcjh_scheme_tet", vcjh_scheme_tet, 0);
  opts.getScalarValue("c_tet", c_tet, 0.);
  opts.getScalarValue("eta_tet", eta_tet, 0.);
  opts.getScalarValue("sparse_tet", sparse_tet, 0);
  opts.getScalarValue("upts_type_pri_tri", upts_type_pri_tri, 0);
  opts.getScalarValue("upts_type_pri_1d", upts_type_pri_1d, 0);
  opts.getScalarValue("vcjh_scheme_pri_1d", vcjh_scheme_pri_1d, 0);
  opts.getScalarValue("eta_pri", eta_pri, 0.);
  opts.getScalarValue("sparse_pri", sparse_pri); // required, as in the reference (src/input.cpp:297)

  /* ---- Advection-Diffusion Parameters ---- */
  wave_speed.setup(3);
  if (equation == 1)
  {
    opts.getScalarValue("wave_speed_x", wave_speed(0));
    opts.getScalarValue("wave_speed_y", wave_speed(1), 0.);
    opts.getScalarValue("wave_speed_z", wave_speed(2), 0.);
    opts.getScalarValue("diff_coeff", diff_coeff, 0.);
    opts.getScalarValue("lambda", lambda);
  }

  opts.getScalarValue("body_forcing", forcing, 0);
  if (forcing) FatalError("body forcing (eles::evaluate_body_force, periodic channel / hill) is not built in this host mirror");
  opts.getScalarValue("perturb_ic", perturb_ic, 0);
  if (ic_form == 6)
  {
    opts.getVectorValue("x_coeffs", x_coeffs);
    opts.getVectorValue("y_coeffs", y_coeffs);
    opts.getVectorValue("z_coeffs", z_coeffs);
  }
  opts.getScalarValue("device_fused", device_fused, 1);
  (void)rank;
}

void input::read_boundary_param()
{
  param_reader bdy_r(fileNameS);
  for (size_t i = 0; i < bc_list.size(); i++)
  {
    bc &b = bc_list[i];
    string pre = "bc_" + b.get_bc_name() + "_";
    string bc_type;
    bdy_r.getScalarValue(pre + "type", bc_type);
    if (b.set_bc_flag(bc_type) == -1)
      FatalError("Boundary condition not implemented yet");
    int flag = b.get_bc_flag();
    if (flag == SUB_IN_SIMP)
    {
      bdy_r.getScalarValue(pre + "rho", b.rho);
      b.velocity.setup(3);
      bdy_r.getScalarValue(pre + "u", b.velocity(0));
      bdy_r.getScalarValue(pre + "v", b.velocity(1));
      bdy_r.getScalarValue(pre + "w", b.velocity(2));
      bdy_r.getScalarValue(pre + "inlet_type", b.type, 0);
    }
    else if (flag == SUB_IN_CHAR)
    {
      bdy_r.getScalarValue(pre + "p_total", b.p_total);
      bdy_r.getScalarValue(pre + "T_total", b.T_total);
      bdy_r.getScalarValue(pre + "pressure_ramp", b.pressure_ramp, 0);
      bdy_r.getScalarValue(pre + "nx", b.nx, 1.);
      bdy_r.getScalarValue(pre + "ny", b.ny, 0.);
      bdy_r.getScalarValue(pre + "nz", b.nz, 0.);
      bdy_r.getScalarValue(pre + "inlet_type", b.type, 0);
      if (b.pressure_ramp)
      {
        // total pressure / temperature ramped from the *_old values, one increment per time step (reference src/input.cpp:374-382)
        pressure_ramp = 1;
        ramp_counter = 1;
        bdy_r.getScalarValue(pre + "p_ramp_coeff", b.p_ramp_coeff, 0.);
        bdy_r.getScalarValue(pre + "T_ramp_coeff", b.T_ramp_coeff, 0.);
        bdy_r.getScalarValue(pre + "p_total_old", b.p_total_old);
        bdy_r.getScalarValue(pre + "T_total_old", b.T_total_old, T_free_stream);
      }
    }
    else if (flag == SUB_OUT_SIMP || flag == SUB_OUT_CHAR)
    {
      bdy_r.getScalarValue(pre + "p_static", b.p_static);
      bdy_r.getScalarValue(pre + "T_total", b.T_total, T_free_stream);
    }
    else if (flag == SUP_IN || flag == CHAR)
    {
      bdy_r.getScalarValue(pre + "p_static", b.p_static);
      bdy_r.getScalarValue(pre + "mach", b.mach);
      bdy_r.getScalarValue(pre + "nx", b.nx, 1.);
      bdy_r.getScalarValue(pre + "ny", b.ny, 0.);
      bdy_r.getScalarValue(pre + "nz", b.nz, 0.);
      bdy_r.getScalarValue(pre + "T_static", b.T_static);
      if (flag == SUP_IN) bdy_r.getScalarValue(pre + "inlet_type", b.type, 0);
    }
    else if (flag == ISOTHERM_WALL)
    {
      if (!viscous) FatalError("Isothermal wall boundary only available to viscous simulation");
      bdy_r.getScalarValue(pre + "T_static", b.T_static);
      b.velocity.setup(3);
      bdy_r.getScalarValue(pre + "u", b.velocity(0), 0.);
      bdy_r.getScalarValue(pre + "v", b.velocity(1), 0.);
      bdy_r.getScalarValue(pre + "w", b.velocity(2), 0.);
      if (wall_model) bdy_r.getScalarValue(pre + "use_wm", b.use_wm, 0);
    }
    else if (flag == ADIABAT_WALL)
    {
      if (!viscous) FatalError("Adiabatic wall boundary only available to viscous simulation");
      b.velocity.setup(3);
      bdy_r.getScalarValue(pre + "u", b.velocity(0), 0.);
      bdy_r.getScalarValue(pre + "v", b.velocity(1), 0.);
      bdy_r.getScalarValue(pre + "w", b.velocity(2), 0.);
      if (wall_model) bdy_r.getScalarValue(pre + "use_wm", b.use_wm, 0);
    }
    if (b.type != 0)
      FatalError("turbulent (synthetic-eddy) inlets are outside the hot-path scope of this build (SURVEY.md §8f)");
  }

  // non-dimensionalise (reference src/input.cpp:441-524)
  for (size_t i = 0; i < bc_list.size(); i++)
  {
    bc &b = bc_list[i];
    int flag = b.get_bc_flag();
    if (flag == SUB_IN_SIMP)
    {
      if (viscous)
      {
        b.rho /= rho_ref;
        for (int j = 0; j < 3; j++) b.velocity(j) /= uvw_ref;
      }
    }
    else if (flag == SUB_IN_CHAR)
    {
      if (viscous)
      {
        b.T_total /= T_ref;
        b.p_total /= p_ref;
        if (b.pressure_ramp) { b.p_total_old /= p_ref; b.T_total_old /= T_ref; }
      }
    }
    else if (flag == SUB_OUT_SIMP || flag == SUB_OUT_CHAR)
    {
      if (viscous) { b.p_static /= p_ref; b.T_total /= T_ref; }
    }
    else if (flag == SUP_IN || flag == CHAR)
    {
      b.rho = b.p_static / (R_gas * b.T_static);
      b.velocity.setup(3);
      b.velocity(0) = b.mach * sqrt(gamma * R_gas * b.T_static) * b.nx;
      b.velocity(1) = b.mach * sqrt(gamma * R_gas * b.T_static) * b.ny;
      b.velocity(2) = b.mach * sqrt(gamma * R_gas * b.T_static) * b.nz;
      if (viscous)
      {
        b.rho /= rho_ref;
        b.p_static /= p_ref;
        b.T_static /= T_ref;
        for (int j = 0; j < 3; j++) b.velocity(j) /= uvw_ref;
      }
    }
    else if (flag == ISOTHERM_WALL)
    {
      if (viscous)
      {
        b.T_static /= T_ref;
        for (int j = 0; j < 3; j++) b.velocity(j) /= uvw_ref;
      }
    }
    else if (flag == ADIABAT_WALL)
    {
      if (viscous)
        for (int j = 0; j < 3; j++) b.velocity(j) /= uvw_ref;
    }
  }
}

// RK tableaux: Ketcheson 2008 low-storage SSP schemes need no a/b arrays (the update formulas are explicit in
// AdvanceSolution); adv_type 3 = Carpenter-Kennedy 1994 RK45(2N); adv_type 4 = Niegemann-Diehl-Busch 2012 RK414(2N).
// (reference data/RK_coeff.dat, included at src/input.cpp:582)
static void set_rk_coeff(input &in)
{
  int t = in.adv_type;
  if (t == 0) { in.RK_a.setup(1); in.RK_b.setup(1); in.RK_c.setup(1); }
  else if (t == 1)
  {
    in.RK_a.setup(1); in.RK_b.setup(1); in.RK_c.setup(4);
    for (int i = 0; i < 4; i++) in.RK_c(i) = i / 3.0;
  }
  else if (t == 2)
  {
    in.RK_a.setup(1); in.RK_b.setup(1); in.RK_c.setup(4);
    for (int i = 0; i < 2; i++) in.RK_c(i) = i / 2.0;
    for (int i = 2; i < 4; i++) in.RK_c(i) = (i - 2.0) / 2.0;
  }
  else if (t == 3)
  {
    in.RK_a.setup(5); in.RK_b.setup(5); in.RK_c.setup(5);
    const double a[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                         -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
    const double b[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                         1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                         2277821191437.0 / 14882151754819.0};
    const double c[5] = {0.0, 1432997174477.0 / 9575080441755.0, 2526269341429.0 / 6820363962896.0,
                         2006345519317.0 / 3224310063776.0, 2802321613138.0 / 2924317926251.0};
    for (int i = 0; i < 5; i++) { in.RK_a(i) = a[i]; in.RK_b(i) = b[i]; in.RK_c(i) = c[i]; }
  }
  else if (t == 4)
  {
    in.RK_a.setup(14); in.RK_b.setup(14); in.RK_c.setup(14);
    const double a[14] = {0.0000000000000000, -0.7188012108672410, -0.7785331173421570, -0.0053282796654044,
                          -0.8552979934029281, -3.9564138245774565, -1.5780575380587385, -2.0837094552574054,
                          -0.7483334182761610, -0.7032861106563359, 0.0013917096117681, -0.0932075369637460,
                          -0.9514200470875948, -7.1151571693922548};
    const double b[14] = {0.0367762454319673, 0.3136296607553959, 0.1531848691869027, 0.0030097086818182,
                          0.3326293790646110, 0.2440251405350864, 0.3718879239592277, 0.6204126221582444,
                          0.1524043173028741, 0.0760894927419266, 0.0077604214040978, 0.0024647284755382,
                          0.0780348340049386, 5.5059777270269628};
    const double c[14] = {0.0000000000000000, 0.0367762454319673, 0.1249685262725025, 0.2446177702277698,
                          0.2476149531070420, 0.2969311120382472, 0.3978149645802642, 0.5270854589440328,
                          0.6981269994175695, 0.8190890835352128, 0.8527059887098624, 0.8604711817462826,
                          0.8627060376969976, 0.8734213127600976};
    for (int i = 0; i < 14; i++) { in.RK_a(i) = a[i]; in.RK_b(i) = b[i]; in.RK_c(i) = c[i]; }
  }
  else
    FatalError("Time advancement scheme not implemented yet!");
}

void input::setup_params(int rank)
{
  if (p_res < 2) FatalError("Plot resolution must be at least 2");
  if (monitor_res_freq == 0) monitor_res_freq = 1000;
  if (monitor_cp_freq == 0) monitor_cp_freq = INT32_MAX;
  if (write_type == 2) FatalError("To use CGNS output, build HiFiLES with CGNS support");

  if (equation == 0)
  {
    if (riemann_solve_type == 1) FatalError("Lax-Friedrich flux not supported with NS/RANS equation");
    if (ic_form == 2 || ic_form == 3 || ic_form == 4 || ic_form == 5)
      FatalError("Initial condition not supported with NS/RANS equation");
  }
  else if (equation == 1)
  {
    if (riemann_solve_type != 1) FatalError("Riemann solver not supported with Advection-Diffusion equation");
    if (ic_form != 2 && ic_form != 3 && ic_form != 4 && ic_form != 5)
      FatalError("Initial condition not supported with Advection-Diffusion equation");
  }
  if (RANS) FatalError("RANS (Spalart-Allmaras) is outside the hot-path scope of this build (SURVEY.md §2 #12)");
  if (LES && !viscous) FatalError("LES not supported with inviscid flow");
  if (LES && (SGS_model < 0 || SGS_model > 4)) FatalError("SGS model not implemented");
  if (wall_model < 0 || wall_model > 2) FatalError("Wall model not implemented!");
  if (over_int && over_int_order < 0) FatalError("Invalid under sampling order");
  if (riemann_solve_type < 0 || riemann_solve_type > 3) FatalError("Riemann solver not implemented");

  set_rk_coeff(*this);

  if (viscous && equation == 0)
  {
    T_ref = T_free_stream;
    L_ref = L_free_stream;
    rho_ref = rho_free_stream;
    uvw_ref = Mach_free_stream * sqrt(gamma * R_gas * T_ref);
    p_ref = rho_ref * uvw_ref * uvw_ref;
    mu_ref = rho_ref * uvw_ref * L_ref;
    time_ref = L_ref / uvw_ref;
    R_ref = (R_gas * T_ref) / (uvw_ref * uvw_ref);
    c_sth = S_gas / T_gas;
    mu_inf = mu_gas / mu_ref;
    rt_inf = T_gas * R_gas / (uvw_ref * uvw_ref);
    if (dt_type == 0) dt /= time_ref;
    if (calc_force) area_ref /= (L_ref * L_ref);
    dx_cyclic /= L_ref;
    dy_cyclic /= L_ref;
    dz_cyclic /= L_ref;
    if (ic_form == 9 || ic_form == 10) x_shock_ic /= L_ref;

    uvw_c_ic = Mach_c_ic * sqrt(gamma * R_gas * T_c_ic);
    u_c_ic = (uvw_c_ic * nx_c_ic) / uvw_ref;
    v_c_ic = (uvw_c_ic * ny_c_ic) / uvw_ref;
    w_c_ic = (uvw_c_ic * nz_c_ic) / uvw_ref;
    if (fix_vis)
      mu_c_ic = mu_gas;
    else
      mu_c_ic = mu_gas * pow(T_c_ic / T_gas, 1.5) * ((T_gas + S_gas) / (T_c_ic + S_gas));
    p_c_ic = rho_c_ic * R_gas * T_c_ic / p_ref;
    mu_c_ic = mu_c_ic / mu_ref;
    rho_c_ic = rho_c_ic / rho_ref;
    T_c_ic = T_c_ic / T_ref;
    Kappa = 0.41;
  }
  else
  {
    T_ref = L_ref = rho_ref = uvw_ref = p_ref = mu_ref = time_ref = R_ref = NAN;
  }
  (void)rank;
}

// explicit instantiations used outside this file
template void param_reader::getScalarValue<int>(const string &, int &, int);
template void param_reader::getScalarValue<double>(const string &, double &, double);
template void param_reader::getScalarValue<string>(const string &, string &, string);
template void param_reader::getScalarValue<int>(const string &, int &);
template void param_reader::getScalarValue<double>(const string &, double &);
template void param_reader::getScalarValue<string>(const string &, string &);
