// TEST INFRASTRUCTURE (oracle side) -- never linked into the product.
//
// Instrumented driver for the UNMODIFIED reference solver objects built by oracle/build_ref.sh.
// It follows the reference main() (src/HiFiLES.cpp:113-222) but (a) skips the VTU/history writers and
// (b) dumps every array the hot path reads or writes, so tests can pin our host setup, our numpy
// restatement (oracle/hifiles_oracle.py) and the CUDA path against the reference's own numbers.
//
//   ref_dump <input_file> <out.hfd> <n_steps> [stagewise]
//
// Output container (.hfd): repeated records  [int32 name_len][name][int32 dtype 0=f64 1=i32][int32 ndim]
// [int64 dims...][raw data, column-major exactly as hf_array stores it (include/hf_array.h:303-325)].
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <string>
#include <vector>
#include <map>
#include <iostream>
#include <fstream>
#include <sstream>
#include <algorithm>
#include <iomanip>
#include <numeric>
#include <iterator>
#include <set>
#include <list>
#include <unistd.h>

#define protected public
#define private public
#include "global.h"
#include "hf_array.h"
#include "input.h"
#include "mesh.h"
#include "eles.h"
#include "inters.h"
#include "int_inters.h"
#include "bdy_inters.h"
#include "solution.h"
#include "geometry.h"
#include "solver.h"
#undef protected
#undef private

using namespace std;

static FILE *g_out = nullptr;

static void put_rec(const string &name, int dtype, const vector<long long> &dims, const void *data, size_t bytes)
{
  int32_t nl = (int32_t)name.size();
  fwrite(&nl, 4, 1, g_out);
  fwrite(name.data(), 1, nl, g_out);
  int32_t dt = dtype, nd = (int32_t)dims.size();
  fwrite(&dt, 4, 1, g_out);
  fwrite(&nd, 4, 1, g_out);
  for (auto d : dims) { int64_t v = d; fwrite(&v, 8, 1, g_out); }
  if (bytes) fwrite(data, 1, bytes, g_out);
}

template <typename T>
static vector<long long> adims(hf_array<T> &a)
{
  vector<long long> d;
  for (int i = 0; i < 4; i++) d.push_back(a.get_dim(i));
  while (d.size() > 1 && d.back() == 1) d.pop_back();
  return d;
}
static void put(const string &name, hf_array<double> &a)
{
  auto d = adims(a);
  size_t n = 1; for (auto x : d) n *= x;
  if (a.get_ptr_cpu() == nullptr) n = 0;
  if (n == 0) return;
  put_rec(name, 0, d, a.get_ptr_cpu(), n * 8);
}
static void put(const string &name, hf_array<int> &a)
{
  auto d = adims(a);
  size_t n = 1; for (auto x : d) n *= x;
  if (a.get_ptr_cpu() == nullptr) n = 0;
  if (n == 0) return;
  put_rec(name, 1, d, a.get_ptr_cpu(), n * 4);
}
static void put_ivec(const string &name, const vector<int> &v, vector<long long> dims = {})
{
  if (dims.empty()) dims = {(long long)v.size()};
  put_rec(name, 1, dims, v.data(), v.size() * 4);
}
static void put_dvec(const string &name, const vector<double> &v, vector<long long> dims = {})
{
  if (dims.empty()) dims = {(long long)v.size()};
  put_rec(name, 0, dims, v.data(), v.size() * 8);
}

static const char *tname[5] = {"tri", "quad", "tet", "pri", "hex"};

// locate which element type's array a pointer points into; return flat index (pt + n_pts*ele) of field 0
static bool locate(struct solution *S, double *p, int which, int &etype, long long &flat)
{
  for (int t = 0; t < S->n_ele_types; t++)
  {
    eles *e = S->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    double *base; long long n;
    if (which == 0) { base = e->disu_fpts.get_ptr_cpu(); n = (long long)e->n_fpts_per_ele * e->n_eles * e->n_fields; }
    else if (which == 1) { base = e->tdA_fpts.get_ptr_cpu(); n = (long long)e->n_fpts_per_ele * e->n_eles; }
    else { base = e->norm_fpts.get_ptr_cpu(); n = (long long)e->n_fpts_per_ele * e->n_eles * e->n_dims; }
    if (p >= base && p < base + n) { etype = t; flat = p - base; return true; }
  }
  return false;
}

static void dump_setup(struct solution *S)
{
  vector<int> meta = {S->n_dims, S->n_ele_types, run_input.order, run_input.viscous, run_input.riemann_solve_type, run_input.adv_type};
  put_ivec("meta", meta);
  vector<double> par = {run_input.gamma, run_input.prandtl, run_input.mu_inf, run_input.rt_inf, run_input.c_sth,
                        (double)run_input.fix_vis, run_input.ldg_beta, run_input.ldg_tau, run_input.dt,
                        run_input.R_ref, run_input.p_c_ic, run_input.rho_c_ic, run_input.T_c_ic, run_input.uvw_c_ic, run_input.uvw_ref};
  put_dvec("params", par);
  if (run_input.adv_type == 3 || run_input.adv_type == 4)
  {
    put("rk_a", run_input.RK_a);
    put("rk_b", run_input.RK_b);
  }
  for (int t = 0; t < S->n_ele_types; t++)
  {
    eles *e = S->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    string p = string(tname[t]) + ".";
    vector<int> sz = {e->n_eles, e->n_upts_per_ele, e->n_fpts_per_ele, e->n_fields, e->n_dims, e->n_inters_per_ele};
    put_ivec(p + "sizes", sz);
    put(p + "n_fpts_per_inter", e->n_fpts_per_inter);
    put(p + "loc_upts", e->loc_upts);
    put(p + "tloc_fpts", e->tloc_fpts);
    put(p + "tnorm_fpts", e->tnorm_fpts);
    put(p + "opp_0", e->opp_0);
    for (int d = 0; d < e->n_dims; d++)
    {
      string sd = to_string(d);
      put(p + "opp_1_" + sd, e->opp_1(d));
      put(p + "opp_2_" + sd, e->opp_2(d));
      if (run_input.viscous)
      {
        put(p + "opp_4_" + sd, e->opp_4(d));
        put(p + "opp_5_" + sd, e->opp_5(d));
      }
    }
    put(p + "opp_3", e->opp_3);
    if (run_input.viscous) put(p + "opp_6", e->opp_6);
    put(p + "shape", e->shape);
    put(p + "n_spts_per_ele", e->n_spts_per_ele);
    put(p + "ele2global_ele", e->ele2global_ele);
    put(p + "detjac_upts", e->detjac_upts);
    put(p + "JGinv_upts", e->JGinv_upts);
    put(p + "detjac_fpts", e->detjac_fpts);
    put(p + "JGinv_fpts", e->JGinv_fpts);
    put(p + "tdA_fpts", e->tdA_fpts);
    put(p + "norm_fpts", e->norm_fpts);
    put(p + "pos_upts", e->pos_upts);
    put(p + "pos_fpts", e->pos_fpts);
    put(p + "disu_upts_ic", e->disu_upts(0));
    if (run_input.over_int)
    {
      put(p + "opp_over_int_cubpts", e->opp_over_int_cubpts);
      put(p + "over_int_filter", e->over_int_filter);
      put(p + "JGinv_over_int_cubpts", e->JGinv_over_int_cubpts);
    }
    if (run_input.shock_cap) put(p + "exp_filter", e->exp_filter);
    if (run_input.LES && run_input.SGS_model == 0) put(p + "wall_distance", e->wall_distance);
    if (run_input.LES && run_input.SGS_model >= 2) put(p + "filter_upts", e->filter_upts);
  }
  const char *iname[3] = {"seg", "tri", "quad"};
  for (int t = 0; t < S->n_int_inter_types; t++)
  {
    int_inters &I = S->mesh_int_inters(t);
    if (I.n_inters == 0) continue;
    string p = string("int_") + iname[t] + ".";
    int nf = I.n_fpts_per_inter, ni = I.n_inters;
    vector<int> idx_l(nf * ni), idx_r(nf * ni), ty_l(ni), ty_r(ni), tda_l(nf * ni), tda_r(nf * ni), nrm(nf * ni);
    for (int i = 0; i < ni; i++)
      for (int j = 0; j < nf; j++)
      {
        int et; long long fl;
        locate(S, I.disu_fpts_l(j, i, 0), 0, et, fl); idx_l[j + nf * i] = (int)fl; ty_l[i] = et;
        locate(S, I.disu_fpts_r(j, i, 0), 0, et, fl); idx_r[j + nf * i] = (int)fl; ty_r[i] = et;
        locate(S, I.tdA_fpts_l(j, i), 1, et, fl); tda_l[j + nf * i] = (int)fl;
        locate(S, I.tdA_fpts_r(j, i), 1, et, fl); tda_r[j + nf * i] = (int)fl;
        locate(S, I.norm_fpts(j, i, 0), 2, et, fl); nrm[j + nf * i] = (int)fl;
      }
    put_ivec(p + "idx_l", idx_l, {nf, ni});
    put_ivec(p + "idx_r", idx_r, {nf, ni});
    put_ivec(p + "type_l", ty_l);
    put_ivec(p + "type_r", ty_r);
    put_ivec(p + "tdA_idx_l", tda_l, {nf, ni});
    put_ivec(p + "tdA_idx_r", tda_r, {nf, ni});
    put_ivec(p + "norm_idx", nrm, {nf, ni});
  }
  for (int t = 0; t < S->n_bdy_inter_types; t++)
  {
    bdy_inters &I = S->mesh_bdy_inters(t);
    if (I.n_inters == 0) continue;
    string p = string("bdy_") + iname[t] + ".";
    int nf = I.n_fpts_per_inter, ni = I.n_inters;
    vector<int> idx_l(nf * ni), ty_l(ni);
    for (int i = 0; i < ni; i++)
      for (int j = 0; j < nf; j++)
      {
        int et; long long fl;
        locate(S, I.disu_fpts_l(j, i, 0), 0, et, fl); idx_l[j + nf * i] = (int)fl; ty_l[i] = et;
      }
    put_ivec(p + "idx_l", idx_l, {nf, ni});
    put_ivec(p + "type_l", ty_l);
    put(p + "boundary_id", I.boundary_id);
    vector<int> flags;
    for (int i = 0; i < run_input.bc_list.get_dim(0); i++) flags.push_back(run_input.bc_list(i).get_bc_flag());
    put_ivec(p + "bc_flags", flags);
    // the boundary table as set_boundary_conditions reads it (run_input.bc_list, reference include/bc.h:50-58), one row per boundary:
    // rho, velocity[3], p_static, T_static, p_total, T_total, mach, nx, ny, nz, use_wm
    vector<double> bp;
    const int nb = run_input.bc_list.get_dim(0);
    for (int i = 0; i < nb; i++)
    {
      bc &b = run_input.bc_list(i);
      bp.push_back(b.rho);
      for (int k = 0; k < 3; k++) bp.push_back(b.velocity.get_dim(0) > k ? b.velocity(k) : 0.);
      bp.push_back(b.p_static); bp.push_back(b.T_static); bp.push_back(b.p_total); bp.push_back(b.T_total); bp.push_back(b.mach);
      bp.push_back(b.nx); bp.push_back(b.ny); bp.push_back(b.nz); bp.push_back((double)b.use_wm);
    }
    put_dvec(p + "bc_params", bp, {13, nb});
    put_dvec(p + "R_ref", vector<double>(1, run_input.viscous ? run_input.R_ref : run_input.R_gas));
  }
}

static void dump_state(struct solution *S, const string &tag, bool all)
{
  for (int t = 0; t < S->n_ele_types; t++)
  {
    eles *e = S->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    string p = tag + "." + tname[t] + ".";
    put(p + "disu_upts", e->disu_upts(0));
    put(p + "div_tconf_upts", e->div_tconf_upts(0));
    if (run_input.shock_cap) put(p + "sensor", e->sensor);
    if (all)
    {
      put(p + "disu_fpts", e->disu_fpts);
      put(p + "tdisf_upts", e->tdisf_upts);
      put(p + "norm_tdisf_fpts", e->norm_tdisf_fpts);
      put(p + "norm_tconf_fpts", e->norm_tconf_fpts);
      if (run_input.viscous)
      {
        put(p + "delta_disu_fpts", e->delta_disu_fpts);
        put(p + "grad_disu_upts", e->grad_disu_upts);
        put(p + "grad_disu_fpts", e->grad_disu_fpts);
      }
    }
  }
}

// CalcResidual (src/solver.cpp:50-223) unrolled for the serial, non-RANS case (LES: eddy-viscosity models, no filter step), with dumps in between.
static void calc_residual_stagewise(struct solution *S, const string &tag)
{
  int n = S->n_ele_types;
  auto dump1 = [&](const string &step, const string &arr) {
    for (int t = 0; t < n; t++)
    {
      eles *e = S->mesh_eles(t);
      if (e->get_n_eles() == 0) continue;
      string p = tag + "." + step + "." + tname[t] + "." + arr;
      if (arr == "disu_fpts") put(p, e->disu_fpts);
      else if (arr == "grad_disu_upts") put(p, e->grad_disu_upts);
      else if (arr == "grad_disu_fpts") put(p, e->grad_disu_fpts);
      else if (arr == "tdisf_upts") put(p, e->tdisf_upts);
      else if (arr == "norm_tconf_fpts") put(p, e->norm_tconf_fpts);
      else if (arr == "delta_disu_fpts") put(p, e->delta_disu_fpts);
      else if (arr == "norm_tdisf_fpts") put(p, e->norm_tdisf_fpts);
      else if (arr == "div_tconf_upts") put(p, e->div_tconf_upts(0));
      else if (arr == "sgsf_upts") put(p, e->sgsf_upts);
      else if (arr == "disuf_upts") put(p, e->disuf_upts);
      else if (arr == "Lu") put(p, e->Lu);
      else if (arr == "Le") put(p, e->Le);
      else if (arr == "sgsf_fpts") put(p, e->sgsf_fpts);
    }
  };
  if (run_input.LES == 1 && (run_input.SGS_model == 2 || run_input.SGS_model == 3 || run_input.SGS_model == 4))
  {
    for (int i = 0; i < n; i++) S->mesh_eles(i)->calc_sgs_terms();
    dump1("s01_calc_sgs_terms", "disuf_upts");
    if (run_input.SGS_model != 3) { dump1("s01_calc_sgs_terms", "Lu"); dump1("s01_calc_sgs_terms", "Le"); }
  }
  for (int i = 0; i < n; i++) S->mesh_eles(i)->extrapolate_solution();
  dump1("s02_extrapolate_solution", "disu_fpts");
  if (run_input.viscous)
  {
    for (int i = 0; i < n; i++) S->mesh_eles(i)->calculate_gradient();
    dump1("s04_calculate_gradient", "grad_disu_upts");
  }
  if (run_input.over_int) for (int i = 0; i < n; i++) S->mesh_eles(i)->evaluate_invFlux_over_int();
  else for (int i = 0; i < n; i++) S->mesh_eles(i)->evaluate_invFlux();
  dump1("s05_evaluate_invFlux", "tdisf_upts");
  for (int i = 0; i < S->n_int_inter_types; i++) S->mesh_int_inters(i).calculate_common_invFlux();
  for (int i = 0; i < S->n_bdy_inter_types; i++) S->mesh_bdy_inters(i).evaluate_boundaryConditions_invFlux(S, S->time);
  dump1("s09_common_invFlux", "norm_tconf_fpts");
  if (run_input.viscous)
  {
    dump1("s09_common_invFlux", "delta_disu_fpts");
    for (int i = 0; i < n; i++) S->mesh_eles(i)->correct_gradient();
    dump1("s11_correct_gradient", "grad_disu_upts");
    dump1("s11_correct_gradient", "grad_disu_fpts");
    for (int i = 0; i < n; i++) S->mesh_eles(i)->evaluate_viscFlux();
    dump1("s13_evaluate_viscFlux", "tdisf_upts");
    if (run_input.LES)
    {
      dump1("s13_evaluate_viscFlux", "sgsf_upts");
      for (int i = 0; i < n; i++) S->mesh_eles(i)->extrapolate_sgsFlux();
      dump1("s14_extrapolate_sgsFlux", "sgsf_fpts");
    }
  }
  for (int i = 0; i < n; i++) S->mesh_eles(i)->extrapolate_totalFlux();
  dump1("s15_extrapolate_totalFlux", "norm_tdisf_fpts");
  for (int i = 0; i < n; i++) S->mesh_eles(i)->calculate_divergence();
  dump1("s16_calculate_divergence", "div_tconf_upts");
  if (run_input.viscous)
  {
    for (int i = 0; i < S->n_int_inter_types; i++) S->mesh_int_inters(i).calculate_common_viscFlux();
    for (int i = 0; i < S->n_bdy_inter_types; i++) S->mesh_bdy_inters(i).evaluate_boundaryConditions_viscFlux(S->time);
    dump1("s17_common_viscFlux", "norm_tconf_fpts");
  }
  for (int i = 0; i < n; i++) S->mesh_eles(i)->calculate_corrected_divergence();
  dump1("s18_corrected_divergence", "div_tconf_upts");
}

int main(int argc, char *argv[])
{
  if (argc < 4) { fprintf(stderr, "usage: ref_dump <input> <out.hfd> <n_steps> [stagewise]\n"); return 2; }
  int n_steps = atoi(argv[3]);
  bool stagewise = argc > 4 && atoi(argv[4]) != 0;
  struct solution FlowSol;
  mesh *mesh_data = new mesh();
  run_input.setup(argv[1], 0);
  SetInput(&FlowSol);
  GeoPreprocess(&FlowSol, *mesh_data);
  // mesh-level connectivity (before delete): faces in creation order
  g_out = fopen(argv[2], "wb");
  if (!g_out) { perror("open out"); return 1; }
  {
    mesh &m = *mesh_data;
    vector<int> f2c(2 * m.num_inters), f2l(2 * m.num_inters), rt(m.num_inters), f2nv(m.num_inters);
    for (int i = 0; i < m.num_inters; i++)
    {
      f2c[2 * i] = m.f2c(i, 0); f2c[2 * i + 1] = m.f2c(i, 1);
      f2l[2 * i] = m.f2loc_f(i, 0); f2l[2 * i + 1] = m.f2loc_f(i, 1);
      rt[i] = (m.f2c(i, 1) >= 0) ? m.rot_tag(i) : -1;
      f2nv[i] = m.f2nv(i);
    }
    put_ivec("mesh.f2c", f2c, {2, m.num_inters});
    put_ivec("mesh.f2loc_f", f2l, {2, m.num_inters});
    put_ivec("mesh.rot_tag", rt);
    put_ivec("mesh.f2nv", f2nv);
    put("mesh.c2v", m.c2v);
    put("mesh.ctype", m.ctype);
    put("mesh.bc_id", m.bc_id);
    put("mesh.xv", m.xv);
  }
  delete mesh_data;
  InitSolution(&FlowSol);
  // Accommodation for a reference defect: eles::calc_sgs_terms filters the velocity-energy products with the column
  // count of the velocity-velocity products (dim3*n_eles instead of n_dims*n_eles, src/eles.cpp:2164, 2170), writing
  // past the end of Le (heap corruption, the stock binary aborts at exit).  The first n_dims slices it computes are
  // correct; give ue / Le the larger extent so that the overrun lands in owned memory.
  if (run_input.LES && (run_input.SGS_model == 2 || run_input.SGS_model == 4))
    for (int t = 0; t < FlowSol.n_ele_types; t++)
    {
      eles *e = FlowSol.mesh_eles(t);
      if (e->get_n_eles() == 0) continue;
      int dim3 = e->n_dims == 2 ? 3 : 6;
      e->ue.setup(e->n_upts_per_ele, e->n_eles, dim3);
      e->Le.setup(e->n_upts_per_ele, e->n_eles, dim3);
      e->ue.initialize_to_zero();
      e->Le.initialize_to_zero();
    }
  dump_setup(&FlowSol);

  int RKSteps = 1;
  if (run_input.adv_type == 1 || run_input.adv_type == 2) RKSteps = 4;
  else if (run_input.adv_type == 3) RKSteps = 5;
  else if (run_input.adv_type == 4) RKSteps = 14;

  vector<double> hist;
  for (int it = 0; it < n_steps; it++)
  {
    calc_time_step(&FlowSol);
    for (int i = 0; i < RKSteps; i++)
    {
      if (stagewise && it == 0 && i == 0)
        calc_residual_stagewise(&FlowSol, "step0.stage0");
      else
        CalcResidual(FlowSol.ini_iter + it, i, &FlowSol);
      if (stagewise && it == 0)
        dump_state(&FlowSol, "step0.stage" + to_string(i) + ".residual", false);
      for (int j = 0; j < FlowSol.n_ele_types; j++)
        FlowSol.mesh_eles(j)->AdvanceSolution(i, run_input.adv_type);
      if (run_input.shock_cap)
        for (int j = 0; j < FlowSol.n_ele_types; j++)
          FlowSol.mesh_eles(j)->shock_capture();
      if (stagewise && it == 0)
        dump_state(&FlowSol, "step0.stage" + to_string(i) + ".advanced", false);
    }
    FlowSol.time += run_input.dt;
    run_input.time = FlowSol.time;
    if (run_input.pressure_ramp) run_input.ramp_counter++; // as the reference's main loop (src/HiFiLES.cpp:224-225)
    // residual norms as output::CalcNormResidual (src/output.cpp:2166-2248), serial
    int nf = 0, np = 0;
    for (int t = 0; t < FlowSol.n_ele_types; t++)
      if (FlowSol.mesh_eles(t)->get_n_eles() != 0)
      { nf = FlowSol.mesh_eles(t)->get_n_fields(); np += FlowSol.mesh_eles(t)->get_n_eles() * FlowSol.mesh_eles(t)->get_n_upts_per_ele(); }
    for (int f = 0; f < nf; f++)
    {
      double s = 0.;
      for (int t = 0; t < FlowSol.n_ele_types; t++)
        if (FlowSol.mesh_eles(t)->get_n_eles() != 0)
        {
          double v = FlowSol.mesh_eles(t)->compute_res_upts(run_input.res_norm_type, f);
          if (run_input.res_norm_type == 0) s = max(s, v); else s += v;
        }
      if (run_input.res_norm_type == 1) s = s / np;
      else if (run_input.res_norm_type == 2) s = sqrt(s) / np;
      hist.push_back(s);
    }
    vector<double> dtv = {run_input.dt, FlowSol.time};
    put_dvec("step" + to_string(it) + ".dt_time", dtv);
  }
  if (n_steps > 0)
  {
    put_dvec("history.norm_residual", hist, {(long long)(hist.size() / n_steps), n_steps});
    dump_state(&FlowSol, "final", run_input.viscous || true);
  }
  fclose(g_out);
  printf("ref_dump: wrote %s\n", argv[2]);
  return 0;
}
