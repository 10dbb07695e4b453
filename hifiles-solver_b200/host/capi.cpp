// C entry points of the host mirror, for hosts that are not C++ (the Python tests and bench drive the library
// through these with ctypes).  One handle = one `struct solution` + its run_input, i.e. what the reference's
// main() owns (src/HiFiLES.cpp:41-130).  run_input is a process-wide singleton in the reference
// (src/global.cpp:29); it is here too, so only one handle may be live at a time.
#include <string>
#include "hifiles.h"
#include <cstring>
#include <memory>

using namespace std;

namespace
{
thread_local string g_err;
struct run_handle
{
  solution FlowSol;
  map<string, vector<int>> int_cache;
  map<string, vector<double>> dbl_cache;
};
run_handle *g_live = nullptr;

template <typename F>
int guard(F f)
{
  try
  {
    f();
    return 0;
  }
  catch (const std::exception &e)
  {
    g_err = e.what();
    return 1;
  }
}

const char *k_tname[5] = {"tri", "quad", "tet", "pri", "hex"};
const char *k_iname[3] = {"seg", "tri", "quad"};
} // namespace

extern "C"
{

const char *hifiles_last_error(void) { return g_err.c_str(); }

/* flags: bit 0 = host only (no device context, no uploads).  part may be NULL when nproc == 1; when nproc > 1 and
 * part is NULL a recursive-coordinate-bisection partition is computed. */
int hifiles_create(const char *input_file, int rank, int nproc, const int *part, long long n_part, int flags, void **out)
{
  return guard([&]() {
    if (g_live) FatalError("a HiFiLES run is already live in this process (run_input is a singleton)");
    unique_ptr<run_handle> h(new run_handle());
    run_input = input();
    run_input.setup(input_file, rank);
    h->FlowSol.rank = rank;
    h->FlowSol.nproc = nproc;
    h->FlowSol.no_device = (flags & 1) ? 1 : 0;
    if (part && n_part > 0) h->FlowSol.part.assign(part, part + n_part);
    SetInput(&h->FlowSol);
    mesh mesh_data;
    GeoPreprocess(&h->FlowSol, mesh_data);
    InitSolution(&h->FlowSol);
    g_live = h.get();
    *out = h.release();
  });
}

int hifiles_destroy(void *handle)
{
  return guard([&]() {
    run_handle *h = (run_handle *)handle;
    if (h == g_live) g_live = nullptr;
    delete h;
  });
}

int hifiles_device_ctx(void *handle, hf_ctx **ctx)
{
  return guard([&]() { *ctx = ((run_handle *)handle)->FlowSol.ctx; });
}

/* reference call sequence of one RK stage: CalcResidual then AdvanceSolution for every element type */
int hifiles_calc_residual(void *handle, int rk_stage)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    CalcResidual(S->ini_iter, rk_stage, S);
  });
}

int hifiles_advance_solution(void *handle, int rk_stage)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    for (int j = 0; j < S->n_ele_types; j++) S->mesh_eles(j)->AdvanceSolution(rk_stage, run_input.adv_type);
    if (run_input.shock_cap)
      for (int j = 0; j < S->n_ele_types; j++) S->mesh_eles(j)->shock_capture();
  });
}

/* the reference's time loop (src/HiFiLES.cpp:194-223) for n_steps steps; fused != 0 uses AdvanceSteps */
int hifiles_run(void *handle, int n_steps, int fused)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    if (fused && run_input.dt_type == 0)
    {
      AdvanceSteps(S, n_steps);
      return;
    }
    int RKSteps = get_n_rk_steps(run_input.adv_type);
    // CFL time steps (dt_type 1, 2): calc_time_step on the device before every step, then the stages -- through the blocked element
    // kernels when the fast mode was asked for and they apply, else method by method as the reference's main loop
    const bool stage_calls = fused && std::string(hf_dev_fused_status(S->ctx)) != "available" && std::string(hf_dev_elem_status(S->ctx)) == "available";
    for (int it = 0; it < n_steps; it++)
    {
      calc_time_step(S);
      if (run_input.pressure_ramp) upload_bc_table(S);
      for (int i = 0; i < RKSteps; i++)
      {
        if (stage_calls)
        {
          hf_check(hf_dev_rk_stage(S->ctx, i, S->time, (it == n_steps - 1 && i == RKSteps - 1) ? 1 : 0));
          continue;
        }
        CalcResidual(S->ini_iter + it, i, S);
        for (int j = 0; j < S->n_ele_types; j++) S->mesh_eles(j)->AdvanceSolution(i, run_input.adv_type);
        if (run_input.shock_cap)
          for (int j = 0; j < S->n_ele_types; j++) S->mesh_eles(j)->shock_capture();
      }
      S->time += run_input.dt;
      run_input.time = S->time;
      if (run_input.pressure_ramp) run_input.ramp_counter++;
    }
    if (stage_calls) hf_check(hf_dev_check_residual(S->ctx));
  });
}

int hifiles_calc_time_step(void *handle, double *dt)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    calc_time_step(S);
    *dt = run_input.dt;
  });
}

/* join the NCCL communicator of a multi-rank run: id = 128-byte ncclUniqueId made on rank 0 (hf_dev_nccl_unique_id) */
int hifiles_nccl_init(void *handle, const char *unique_id_128_bytes)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    hf_check(hf_dev_nccl_init(S->ctx, unique_id_128_bytes));
    FinishWallDistance(S); // Smagorinsky: the other ranks' wall points (reference src/geometry.cpp:768-892)
  });
}

/* output::write_vtu of the current solution into the working directory (file name from data_file_name and iter) */
int hifiles_write_vtu(void *handle, int iter)
{
  return guard([&]() { write_plot(iter, &((run_handle *)handle)->FlowSol); });
}

/* output::write_restart_ascii of the current solution into the working directory: Rest_<iter>_p0000.dat, or one file per rank in the
 * folder Rest_<iter>/ when the run is partitioned (reference src/output.cpp:1753-1818) */
int hifiles_write_restart(void *handle, int iter)
{
  return guard([&]() { write_restart_ascii(&((run_handle *)handle)->FlowSol, iter); });
}

int hifiles_norm_residual(void *handle, double *out, int n)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    CalcNormResidual(S);
    for (int i = 0; i < n && i < 6; i++) out[i] = S->norm_residual(i);
  });
}

int hifiles_copy_solution_to_host(void *handle)
{
  return guard([&]() {
    solution *S = &((run_handle *)handle)->FlowSol;
    for (int j = 0; j < S->n_ele_types; j++)
    {
      S->mesh_eles(j)->cp_disu_upts_gpu_cpu();
      S->mesh_eles(j)->cp_div_tconf_upts_gpu_cpu();
    }
  });
}

double hifiles_get_scalar(void *handle, const char *name)
{
  solution *S = &((run_handle *)handle)->FlowSol;
  string n(name);
  if (n == "dt") return run_input.dt;
  if (n == "time") return S->time;
  if (n == "gamma") return run_input.gamma;
  if (n == "prandtl") return run_input.prandtl;
  if (n == "mu_inf") return run_input.mu_inf;
  if (n == "rt_inf") return run_input.rt_inf;
  if (n == "c_sth") return run_input.c_sth;
  if (n == "fix_vis") return run_input.fix_vis;
  if (n == "ldg_beta") return run_input.ldg_beta;
  if (n == "ldg_tau") return run_input.ldg_tau;
  if (n == "R_ref") return run_input.R_ref;
  if (n == "n_dims") return S->n_dims;
  if (n == "order") return run_input.order;
  if (n == "viscous") return run_input.viscous;
  if (n == "riemann_solve_type") return run_input.riemann_solve_type;
  if (n == "adv_type") return run_input.adv_type;
  if (n == "equation") return run_input.equation;
  if (n == "n_steps") return run_input.n_steps;
  if (n == "res_norm_type") return run_input.res_norm_type;
  if (n == "n_rk") return get_n_rk_steps(run_input.adv_type);
  if (n == "n_bc") return (double)run_input.bc_list.size();
  if (n.rfind("RK_a", 0) == 0) return run_input.RK_a(atoi(n.c_str() + 4));
  if (n.rfind("RK_b", 0) == 0) return run_input.RK_b(atoi(n.c_str() + 4));
  if (n.rfind("bc_flag", 0) == 0) return run_input.bc_list[atoi(n.c_str() + 7)].get_bc_flag();
  return NAN;
}

/* Named views of host arrays ("hex.opp_0", "hex.detjac_upts", "int_quad.idx_l", ...): the same names
 * oracle/ref_dump.cpp writes, so tests can diff setup against the reference dump key by key.
 * dtype 0 = double, 1 = int32.  Pointers stay valid until the handle is destroyed. */
int hifiles_get_array(void *handle, const char *name, const void **ptr, int *dtype, int *ndim, long long *dims)
{
  return guard([&]() {
    run_handle *h = (run_handle *)handle;
    solution *S = &h->FlowSol;
    string full(name);
    size_t dot = full.find('.');
    if (dot == string::npos) FatalError("array name must be <object>.<array>");
    string obj = full.substr(0, dot), arr = full.substr(dot + 1);
    auto set_d = [&](hf_array<double> &a) {
      *ptr = a.get_ptr_cpu(); *dtype = 0;
      int nd = 4;
      while (nd > 1 && a.get_dim(nd - 1) == 1) nd--;
      *ndim = nd;
      for (int i = 0; i < nd; i++) dims[i] = a.get_dim(i);
    };
    auto set_i = [&](hf_array<int> &a) {
      *ptr = a.get_ptr_cpu(); *dtype = 1;
      int nd = 4;
      while (nd > 1 && a.get_dim(nd - 1) == 1) nd--;
      *ndim = nd;
      for (int i = 0; i < nd; i++) dims[i] = a.get_dim(i);
    };
    auto set_iv = [&](vector<int> &v, long long d0, long long d1) {
      *ptr = v.data(); *dtype = 1;
      if (d1 > 0) { *ndim = 2; dims[0] = d0; dims[1] = d1; }
      else { *ndim = 1; dims[0] = d0; }
    };
    for (int t = 0; t < 5; t++)
    {
      if (obj != k_tname[t]) continue;
      eles *e = S->mesh_eles(t);
      if (e->get_n_eles() == 0) FatalError("no elements of type " + obj);
      if (arr == "sizes")
      {
        vector<int> &v = h->int_cache[full];
        v = {e->n_eles, e->n_upts_per_ele, e->n_fpts_per_ele, e->n_fields, e->n_dims, e->n_inters_per_ele};
        set_iv(v, 6, 0);
        return;
      }
      if (arr == "n_fpts_per_inter") return set_i(e->n_fpts_per_inter);
      if (arr == "loc_upts") return set_d(e->loc_upts);
      if (arr == "tloc_fpts") return set_d(e->tloc_fpts);
      if (arr == "tnorm_fpts") return set_d(e->tnorm_fpts);
      if (arr == "opp_0") return set_d(e->opp_0);
      if (arr == "opp_3") return set_d(e->opp_3);
      if (arr == "opp_6") return set_d(e->opp_6);
      for (int d = 0; d < e->n_dims; d++)
      {
        string sd = to_string(d);
        if (arr == "opp_1_" + sd) return set_d(e->opp_1(d));
        if (arr == "opp_2_" + sd) return set_d(e->opp_2(d));
        if (arr == "opp_4_" + sd) return set_d(e->opp_4(d));
        if (arr == "opp_5_" + sd) return set_d(e->opp_5(d));
      }
      if (arr == "opp_over_int_cubpts") return set_d(e->opp_over_int_cubpts);
      if (arr == "over_int_filter") return set_d(e->over_int_filter);
      if (arr == "JGinv_over_int_cubpts") return set_d(e->JGinv_over_int_cubpts);
      if (arr == "exp_filter") return set_d(e->exp_filter);
      if (arr == "wall_distance") return set_d(e->wall_distance);
      if (arr == "filter_upts") return set_d(e->filter_upts);
      if (arr == "shape") return set_d(e->shape);
      if (arr == "n_spts_per_ele") return set_i(e->n_spts_per_ele);
      if (arr == "ele2global_ele") return set_i(e->ele2global_ele);
      if (arr == "detjac_upts") return set_d(e->detjac_upts);
      if (arr == "JGinv_upts") return set_d(e->JGinv_upts);
      if (arr == "detjac_fpts") return set_d(e->detjac_fpts);
      if (arr == "JGinv_fpts") return set_d(e->JGinv_fpts);
      if (arr == "tdA_fpts") return set_d(e->tdA_fpts);
      if (arr == "norm_fpts") return set_d(e->norm_fpts);
      if (arr == "pos_upts") return set_d(e->pos_upts);
      if (arr == "pos_fpts") return set_d(e->pos_fpts);
      if (arr == "disu_upts") return set_d(e->disu_upts(0));
      if (arr == "div_tconf_upts") return set_d(e->div_tconf_upts(0));
      if (arr == "h_ref") return set_d(e->h_ref);
      FatalError("unknown element array " + full);
    }
    for (int t = 0; t < 3; t++)
    {
      if (obj == string("int_") + k_iname[t])
      {
        int_inters &I = S->mesh_int_inters[t];
        int nf = I.n_fpts_per_inter, ni = I.n_inters;
        if (arr == "type_l") return set_i(I.ele_type_l);
        if (arr == "type_r") return set_i(I.ele_type_r);
        if (arr == "ele_l") return set_i(I.ele_l);
        if (arr == "ele_r") return set_i(I.ele_r);
        if (arr == "local_inter_l") return set_i(I.local_inter_l);
        if (arr == "local_inter_r") return set_i(I.local_inter_r);
        if (arr == "rot_tag") return set_i(I.rot_tags);
        if (arr == "idx_l" || arr == "idx_r")
        {
          // flat (fpt + n_fpts_per_ele*ele) index of each flux-point pair, left as j, right as lut(j)
          // (reference src/int_inters.cpp:76-95)
          vector<int> &v = h->int_cache[full];
          v.assign((size_t)nf * ni, 0);
          bool left = arr == "idx_l";
          for (int i = 0; i < ni; i++)
          {
            I.get_lut(I.rot_tags(i));
            eles *e = S->mesh_eles(left ? I.ele_type_l(i) : I.ele_type_r(i));
            for (int j = 0; j < nf; j++)
            {
              int fpt = left ? e->get_fpt_index(j, I.local_inter_l(i)) : e->get_fpt_index(I.lut(j), I.local_inter_r(i));
              v[j + (size_t)nf * i] = fpt + e->n_fpts_per_ele * (left ? I.ele_l(i) : I.ele_r(i));
            }
          }
          return set_iv(v, nf, ni);
        }
        FatalError("unknown interface array " + full);
      }
      if (obj == string("bdy_") + k_iname[t])
      {
        bdy_inters &I = S->mesh_bdy_inters[t];
        int nf = I.n_fpts_per_inter, ni = I.n_inters;
        if (arr == "type_l") return set_i(I.ele_type_l);
        if (arr == "ele_l") return set_i(I.ele_l);
        if (arr == "local_inter_l") return set_i(I.local_inter_l);
        if (arr == "boundary_id") return set_i(I.boundary_id);
        if (arr == "bc_params")
        {
          // the boundary table as set_boundary_conditions reads it (run_input.bc_list after read_boundary_param's non-dimensionalisation), one
          // column per boundary: rho, velocity[3], p_static, T_static, p_total, T_total, mach, nx, ny, nz, use_wm (oracle/ref_dump.cpp)
          vector<double> &v = h->dbl_cache[full];
          v.clear();
          const int nb = (int)run_input.bc_list.size();
          for (int i = 0; i < nb; i++)
          {
            bc &b = run_input.bc_list[i];
            v.push_back(b.rho);
            for (int k = 0; k < 3; k++) v.push_back(b.velocity.get_dim(0) > k ? b.velocity(k) : 0.);
            v.push_back(b.p_static); v.push_back(b.T_static); v.push_back(b.p_total); v.push_back(b.T_total); v.push_back(b.mach);
            v.push_back(b.nx); v.push_back(b.ny); v.push_back(b.nz); v.push_back((double)b.use_wm);
          }
          *ptr = v.data(); *dtype = 0; *ndim = 2; dims[0] = 13; dims[1] = nb;
          return;
        }
        if (arr == "R_ref")
        {
          vector<double> &v = h->dbl_cache[full];
          v.assign(1, run_input.viscous ? run_input.R_ref : run_input.R_gas);
          *ptr = v.data(); *dtype = 0; *ndim = 1; dims[0] = 1;
          return;
        }
        if (arr == "idx_l")
        {
          vector<int> &v = h->int_cache[full];
          v.assign((size_t)nf * ni, 0);
          for (int i = 0; i < ni; i++)
          {
            eles *e = S->mesh_eles(I.ele_type_l(i));
            for (int j = 0; j < nf; j++) v[j + (size_t)nf * i] = e->get_fpt_index(j, I.local_inter_l(i)) + e->n_fpts_per_ele * I.ele_l(i);
          }
          return set_iv(v, nf, ni);
        }
        FatalError("unknown interface array " + full);
      }
      if (obj == string("mpi_") + k_iname[t])
      {
        mpi_inters &I = S->mesh_mpi_inters[t];
        if (arr == "type_l") return set_i(I.ele_type_l);
        if (arr == "ele_l") return set_i(I.ele_l);
        if (arr == "local_inter_l") return set_i(I.local_inter_l);
        if (arr == "rot_tag") return set_i(I.rot_tags);
        if (arr == "neighbour_rank") return set_iv(I.neighbour_rank, (long long)I.neighbour_rank.size(), 0);
        if (arr == "neighbour_count") return set_iv(I.neighbour_count, (long long)I.neighbour_count.size(), 0);
        FatalError("unknown interface array " + full);
      }
    }
    FatalError("unknown object " + obj);
  });
}

int hifiles_n_inters(void *handle, int kind, int inter_type)
{
  solution *S = &((run_handle *)handle)->FlowSol;
  if (kind == 0) return S->mesh_int_inters[inter_type].n_inters;
  if (kind == 1) return S->mesh_bdy_inters[inter_type].n_inters;
  return S->mesh_mpi_inters[inter_type].n_inters;
}

int hifiles_n_eles(void *handle, int ele_type)
{
  solution *S = &((run_handle *)handle)->FlowSol;
  return S->mesh_eles(ele_type)->get_n_eles();
}

} // extern "C"
