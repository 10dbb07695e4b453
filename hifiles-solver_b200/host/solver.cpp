// Residual orchestration: the same fixed sequence of element / interface method calls as the reference's
// CalcResidual (src/solver.cpp:50-223), InitSolution (:321-375), calc_time_step (:484-549) and
// output::CalcNormResidual (src/output.cpp:2166-2248).  Each call below is one asynchronous device call.
#include "hifiles.h"

using namespace std;

int get_n_rk_steps(int adv_type)
{
  if (adv_type == 0) return 1;
  if (adv_type == 1 || adv_type == 2) return 4;
  if (adv_type == 3) return 5;
  if (adv_type == 4) return 14;
  FatalError("ERROR: Time integration type not recognised ... ");
  return 0;
}

/*! Boundary parameter table of the device.  A characteristic subsonic inlet with pressure_ramp takes its total pressure and
 *  temperature from the ramp of this time step (reference src/bdy_inters.cpp:481-510: linear in run_input.ramp_counter from
 *  the *_old value, clipped at the target; the counter advances once per time step, src/HiFiLES.cpp:224-225). */
void upload_bc_table(struct solution *FlowSol)
{
  vector<hf_bc> table(run_input.bc_list.size());
  for (size_t i = 0; i < table.size(); i++)
  {
    bc &b = run_input.bc_list[i];
    hf_bc &t = table[i];
    memset(&t, 0, sizeof(t));
    t.bc_flag = b.get_bc_flag();
    t.rho = b.rho;
    for (int k = 0; k < 3; k++) t.velocity[k] = b.velocity.size() == 3 ? b.velocity(k) : 0.;
    t.p_static = b.p_static;
    t.T_static = b.T_static;
    t.p_total = b.p_total;
    t.T_total = b.T_total;
    if (t.bc_flag == SUB_IN_CHAR && b.pressure_ramp)
    {
      if (b.p_ramp_coeff)
      {
        double p = b.p_total_old + (b.p_total - b.p_total_old) * b.p_ramp_coeff * run_input.ramp_counter;
        if (p >= b.p_total) p = b.p_total;
        t.p_total = p;
      }
      if (b.T_ramp_coeff > 0)
      {
        double T = b.T_total_old + (b.T_total - b.T_total_old) * b.T_ramp_coeff * run_input.ramp_counter;
        if (T >= b.T_total) T = b.T_total;
        t.T_total = T;
      }
      t.T_isentropic = b.T_ramp_coeff < 0 ? 1 : 0;
    }
    t.mach = b.mach;
    t.nx = b.nx; t.ny = b.ny; t.nz = b.nz;
    t.use_wm = b.use_wm;
  }
  hf_check(hf_dev_set_bc_table(FlowSol->ctx, (int)table.size(), table.empty() ? nullptr : table.data()));
}

static void upload_params(struct solution *FlowSol)
{
  hf_params p;
  memset(&p, 0, sizeof(p));
  p.equation = run_input.equation;
  p.viscous = run_input.viscous;
  p.n_dims = FlowSol->n_dims;
  p.n_fields = (run_input.equation == 0) ? FlowSol->n_dims + 2 : 1;
  p.riemann_solve_type = run_input.riemann_solve_type;
  p.vis_riemann_solve_type = run_input.vis_riemann_solve_type;
  p.adv_type = run_input.adv_type;
  p.dt_type = run_input.dt_type;
  p.fix_vis = run_input.fix_vis;
  p.order = run_input.order;
  p.gamma = run_input.gamma;
  p.prandtl = run_input.prandtl;
  p.mu_inf = run_input.mu_inf;
  p.rt_inf = run_input.rt_inf;
  p.c_sth = run_input.c_sth;
  p.ldg_beta = run_input.ldg_beta;
  p.ldg_tau = run_input.ldg_tau;
  p.dt = run_input.dt;
  p.CFL = run_input.CFL;
  // boundary conditions use the dimensional gas constant when the run is inviscid (reference src/bdy_inters.cpp:368-369)
  p.R_ref = run_input.viscous ? run_input.R_ref : run_input.R_gas;
  for (int i = 0; i < 3; i++) p.wave_speed[i] = run_input.wave_speed.size() ? run_input.wave_speed(i) : 0.;
  p.diff_coeff = run_input.diff_coeff;
  p.lambda = run_input.lambda;
  p.n_rk = get_n_rk_steps(run_input.adv_type);
  for (int i = 0; i < run_input.RK_a.get_dim(0) && i < HF_MAX_RK; i++) p.RK_a[i] = run_input.RK_a(i);
  for (int i = 0; i < run_input.RK_b.get_dim(0) && i < HF_MAX_RK; i++) p.RK_b[i] = run_input.RK_b(i);
  p.LES = run_input.LES;
  p.SGS_model = run_input.SGS_model;
  p.C_s = run_input.C_s;
  p.Kappa = run_input.Kappa;
  p.prandtl_t = run_input.prandtl_t;
  p.filter_ratio = run_input.filter_ratio;
  p.wall_model = run_input.wall_model;
  p.over_int = run_input.over_int;
  p.shock_cap = run_input.shock_cap;
  p.shock_det_field = run_input.shock_det_field;
  p.s0 = run_input.s0;
  hf_check(hf_dev_set_params(FlowSol->ctx, &p));

  upload_bc_table(FlowSol);
}

/*! Move everything to the device (the reference did this piecemeal with mv_all_cpu_gpu calls inside
 *  GeoPreprocess, src/geometry.cpp:310-321, 553-557). */
static void upload_all(struct solution *FlowSol)
{
  upload_params(FlowSol);
  hf_check(hf_dev_set_mode(FlowSol->ctx, run_input.device_fused));
  // device element order: elements without a partition face first (ascending), the partition-adjacent ones behind them
  if (FlowSol->nproc > 1)
  {
    for (int t = 0; t < FlowSol->n_ele_types; t++)
    {
      eles *e = FlowSol->mesh_eles(t);
      const int ne = e->get_n_eles();
      if (ne == 0) continue;
      vector<char> halo(ne, 0);
      bool any = false;
      for (int i = 0; i < FlowSol->n_mpi_inter_types; i++)
      {
        mpi_inters &M = FlowSol->mesh_mpi_inters[i];
        for (int q = 0; q < M.get_n_inters(); q++)
          if (M.ele_type_l(q) == e->get_ele_type()) { halo[M.ele_l(q)] = 1; any = true; }
      }
      if (!any) continue;
      vector<int> pos(ne);
      int slot = 0;
      for (int q = 0; q < ne; q++) if (!halo[q]) pos[q] = slot++;
      for (int q = 0; q < ne; q++) if (halo[q]) pos[q] = slot++;
      hf_check(hf_dev_set_element_order(FlowSol->ctx, e->get_ele_type(), ne, pos.data()));
    }
  }
  for (int i = 0; i < FlowSol->n_ele_types; i++) FlowSol->mesh_eles(i)->mv_all_cpu_gpu();
  for (int i = 0; i < FlowSol->n_int_inter_types; i++) FlowSol->mesh_int_inters[i].mv_all_cpu_gpu();
  for (int i = 0; i < FlowSol->n_bdy_inter_types; i++) FlowSol->mesh_bdy_inters[i].mv_all_cpu_gpu();
  for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].mv_all_cpu_gpu();
  hf_check(hf_dev_finalize_setup(FlowSol->ctx));
}

void InitSolution(struct solution *FlowSol)
{
  if (run_input.restart_flag == 0)
  {
    FlowSol->ini_iter = 0;
    for (int i = 0; i < FlowSol->n_ele_types; i++)
      if (FlowSol->mesh_eles(i)->get_n_eles() != 0) FlowSol->mesh_eles(i)->set_ics(FlowSol->time);
  }
  else if (run_input.restart_flag == 1) // read ascii restart files
  {
    FlowSol->ini_iter = run_input.restart_iter;
    read_restart_ascii(run_input.restart_iter, run_input.n_restart_files, FlowSol);
  }
  else
    FatalError("HiFiLES need to be compiled with HDF5 to read hdf5 format restart file");
  // patch solution after flow field initialized (reference src/solver.cpp:352-354)
  if (run_input.patch)
    for (int i = 0; i < FlowSol->n_ele_types; i++)
      if (FlowSol->mesh_eles(i)->get_n_eles() != 0) FlowSol->mesh_eles(i)->set_patch();
  if (!FlowSol->no_device) upload_all(FlowSol);
}

void CalcResidual(int in_file_num, int in_rk_stage, struct solution *FlowSol)
{
  (void)in_file_num;
  int n = FlowSol->n_ele_types;
  if (run_input.LES == 1 && in_rk_stage == 0 && (run_input.SGS_model == 2 || run_input.SGS_model == 3 || run_input.SGS_model == 4))
    for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->calc_sgs_terms();
  for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->extrapolate_solution();
  if (FlowSol->nproc > 1)
    for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].send_solution();
  if (run_input.viscous)
    for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->calculate_gradient();
  if (run_input.over_int) for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->evaluate_invFlux_over_int();
  else for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->evaluate_invFlux();
  for (int i = 0; i < FlowSol->n_int_inter_types; i++) FlowSol->mesh_int_inters[i].calculate_common_invFlux();
  for (int i = 0; i < FlowSol->n_bdy_inter_types; i++) FlowSol->mesh_bdy_inters[i].evaluate_boundaryConditions_invFlux(FlowSol, FlowSol->time);
  if (FlowSol->nproc > 1)
  {
    for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].receive_solution();
    for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].calculate_common_invFlux();
  }
  if (run_input.viscous)
  {
    for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->correct_gradient();
    if (FlowSol->nproc > 1)
      for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].send_corrected_gradient();
    for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->evaluate_viscFlux();
    if (run_input.LES)
    {
      for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->extrapolate_sgsFlux();
      if (FlowSol->nproc > 1)
        for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].send_sgsf_fpts();
    }
  }
  for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->extrapolate_totalFlux();
  for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->calculate_divergence();
  if (run_input.viscous)
  {
    for (int i = 0; i < FlowSol->n_int_inter_types; i++) FlowSol->mesh_int_inters[i].calculate_common_viscFlux();
    for (int i = 0; i < FlowSol->n_bdy_inter_types; i++) FlowSol->mesh_bdy_inters[i].evaluate_boundaryConditions_viscFlux(FlowSol->time);
    if (FlowSol->nproc > 1)
    {
      for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].receive_corrected_gradient();
      if (run_input.LES)
        for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].receive_sgsf_fpts();
      for (int i = 0; i < FlowSol->n_mpi_inter_types; i++) FlowSol->mesh_mpi_inters[i].calculate_common_viscFlux();
    }
  }
  for (int i = 0; i < n; i++) FlowSol->mesh_eles(i)->calculate_corrected_divergence();
}

// One RK stage.  The reference's main loop calls CalcResidual, then AdvanceSolution for every element type, then
// shock_capture (src/HiFiLES.cpp:205-217); on the device the three are one fused call where the fused kernels apply.
void AdvanceStage(int in_file_num, int in_rk_stage, struct solution *FlowSol, bool monitored)
{
  if (run_input.device_fused && (hf_dev_fused_status(FlowSol->ctx) == string("available") || hf_dev_elem_status(FlowSol->ctx) == string("available")))
  {
    hf_check(hf_dev_rk_stage(FlowSol->ctx, in_rk_stage, FlowSol->time, monitored ? 1 : 0));
    if (monitored) hf_check(hf_dev_check_residual(FlowSol->ctx)); // the host synchronises on monitored stages anyway (residual norm)
    return;
  }
  CalcResidual(in_file_num, in_rk_stage, FlowSol);
  for (int j = 0; j < FlowSol->n_ele_types; j++) FlowSol->mesh_eles(j)->AdvanceSolution(in_rk_stage, run_input.adv_type);
  if (run_input.shock_cap)
    for (int j = 0; j < FlowSol->n_ele_types; j++) FlowSol->mesh_eles(j)->shock_capture();
}

void CalcTimeAverageQuantities(struct solution *FlowSol)
{
  for (int i = 0; i < FlowSol->n_ele_types; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0) FlowSol->mesh_eles(i)->CalcTimeAverageQuantities(FlowSol->time);
}

void CalcIntegralQuantities(struct solution *FlowSol)
{
  const int nintq = run_input.n_integral_quantities;
  FlowSol->integral_quantities.setup(nintq > 0 ? nintq : 1);
  for (int q = 0; q < nintq; q++) FlowSol->integral_quantities(q) = 0.;
  for (int i = 0; i < FlowSol->n_ele_types; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0) FlowSol->mesh_eles(i)->CalcIntegralQuantities(nintq, FlowSol->integral_quantities);
  if (FlowSol->nproc > 1 && nintq) hf_check(hf_dev_allreduce_sum(FlowSol->ctx, FlowSol->integral_quantities.get_ptr_cpu(), nintq));
}

void compute_error(int in_file_num, struct solution *FlowSol)
{
  const int n_fields = (run_input.equation == 0) ? FlowSol->n_dims + 2 : 1;
  hf_array<double> error(2, n_fields);
  for (int j = 0; j < n_fields; j++) error(0, j) = error(1, j) = 0.;
  for (int i = 0; i < FlowSol->n_ele_types; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0)
    {
      hf_array<double> t = FlowSol->mesh_eles(i)->compute_error(run_input.error_norm_type, FlowSol->time);
      for (int j = 0; j < n_fields; j++)
      {
        error(0, j) += t(0, j);
        if (run_input.viscous) error(1, j) += t(1, j);
      }
    }
  if (FlowSol->nproc > 1) hf_check(hf_dev_allreduce_sum(FlowSol->ctx, error.get_ptr_cpu(), 2 * n_fields));
  if (FlowSol->rank != 0) return;
  if (run_input.error_norm_type == 2)
    for (int j = 0; j < n_fields; j++)
    {
      error(0, j) = sqrt(error(0, j));
      if (run_input.viscous) error(1, j) = sqrt(error(1, j));
    }
  else if (run_input.error_norm_type != 1)
    FatalError("Error norm not supported!");
  FILE *f = fopen("error.dat", "a");
  if (!f) FatalError("cannot open error.dat");
  fprintf(f, "%d, %d, %s, %d, %d, %d, ", in_file_num, run_input.order, run_input.mesh_file.c_str(), run_input.adv_type, run_input.riemann_solve_type,
          run_input.error_norm_type);
  for (int j = 0; j < n_fields; j++) fprintf(f, (j == n_fields - 1 && run_input.viscous == 0) ? "%e\n" : "%e, ", error(0, j));
  if (run_input.viscous)
    for (int j = 0; j < n_fields; j++) fprintf(f, j == n_fields - 1 ? "%e\n" : "%e, ", error(1, j));
  fclose(f);
}

void calc_time_step(struct solution *FlowSol)
{
  if (run_input.dt_type == 0) return;
  double dt = 0.;
  hf_check(hf_dev_calc_dt(FlowSol->ctx, &dt));
  run_input.dt = dt;
}

void AdvanceSteps(struct solution *FlowSol, int n_steps)
{
  if (run_input.pressure_ramp)
  {
    // the inlet state changes every step: one step per device call
    for (int it = 0; it < n_steps; it++)
    {
      upload_bc_table(FlowSol);
      hf_check(hf_dev_run_steps(FlowSol->ctx, 1, FlowSol->time));
      FlowSol->time += run_input.dt;
      run_input.ramp_counter++;
    }
    run_input.time = FlowSol->time;
    return;
  }
  hf_check(hf_dev_run_steps(FlowSol->ctx, n_steps, FlowSol->time));
  FlowSol->time += n_steps * run_input.dt;
  run_input.time = FlowSol->time;
}

void CalcNormResidual(struct solution *FlowSol)
{
  int n_fields = (run_input.equation == 0) ? FlowSol->n_dims + 2 : 1;
  FlowSol->norm_residual.setup(6);
  double sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  hf_check(hf_dev_residual_norm(FlowSol->ctx, run_input.res_norm_type, sums));
  long long n_upts_global = 0;
  for (int i = 0; i < FlowSol->n_ele_types; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0)
      n_upts_global += (long long)FlowSol->mesh_eles(i)->get_n_eles() * FlowSol->mesh_eles(i)->get_n_upts_per_ele();
  if (FlowSol->nproc > 1)
  {
    // reference src/output.cpp:2216-2231: MPI_MAX (norm 0) / MPI_SUM of the per-rank sums and of the point count; every rank
    // gets the result, so the NaN abort below fires on all of them together
    double cnt = (double)n_upts_global;
    if (run_input.res_norm_type == 0) hf_check(hf_dev_allreduce_max(FlowSol->ctx, sums, n_fields));
    else hf_check(hf_dev_allreduce_sum(FlowSol->ctx, sums, n_fields));
    hf_check(hf_dev_allreduce_sum(FlowSol->ctx, &cnt, 1));
    n_upts_global = (long long)(cnt + 0.5);
  }
  for (int f = 0; f < n_fields; f++)
  {
    double s = sums[f];
    if (run_input.res_norm_type == 1) s = s / n_upts_global;
    else if (run_input.res_norm_type == 2) s = sqrt(s) / n_upts_global;
    else if (run_input.res_norm_type != 0) FatalError("norm_type not recognized");
    if (std::isnan(s)) FatalError("NaN residual encountered. Exiting");
    FlowSol->norm_residual(f) = s;
  }
}
