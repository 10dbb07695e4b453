// TEST INFRASTRUCTURE (oracle side) -- never linked into the product.
//
// The drop-in boundary proven against the reference's OWN objects: this driver links the unmodified reference solver objects
// (oracle/build_ref.sh: input, mesh reader, GeoPreprocess, eles_*, int_inters, bdy_inters, InitSolution) and, where the reference's
// main loop calls CalcResidual + AdvanceSolution on the CPU (src/HiFiLES.cpp:199-217), hands the reference's arrays to the C ABI of
// include/hifiles_b200.h instead -- the calls a maintainer would add behind `#ifdef _GPU`:
//
//   eles::mv_all_cpu_gpu            ->  hf_dev_upload_eles        from the eles members' hf_array::get_ptr_cpu() pointers
//   int_inters::mv_all_cpu_gpu      ->  hf_dev_upload_int_inters  (element, local face, rotation tag recovered from the double* tables the
//                                                                   reference baked in set_interior, src/int_inters.cpp:67-121)
//   bdy_inters::mv_all_cpu_gpu      ->  hf_dev_upload_bdy_inters + hf_dev_set_bc_table (run_input.bc_list)
//   CalcResidual + AdvanceSolution  ->  hf_dev_rk_stage
//   eles::cp_disu_upts_gpu_cpu      ->  hf_dev_download into eles::disu_upts(0)
//
// Nothing of the repo's host mirror (hifiles-solver_b200/host) is involved: setup, operators, metrics, connectivity and the initial
// solution are the reference's.  Output: the same .hfd container as ref_dump (final.<type>.disu_upts, history.norm_residual), which
// tests/test_reference_objects_gpu.py compares with ref_dump's CPU run of the same input.
//
//   ref_gpu <input_file> <out.hfd> <n_steps> [mode: 1 fast (default) | 0 bit-exact staged kernels]
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

#include "hifiles_b200.h"

using namespace std;

static void ck(int st, const char *what)
{
  if (st != 0)
  {
    fprintf(stderr, "ref_gpu: %s failed: %s\n", what, hf_dev_last_error());
    exit(1);
  }
}
#define CK(call) ck((call), #call)

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
static void put(const string &name, hf_array<double> &a)
{
  vector<long long> d;
  for (int i = 0; i < 4; i++) d.push_back(a.get_dim(i));
  while (d.size() > 1 && d.back() == 1) d.pop_back();
  size_t n = 1;
  for (auto x : d) n *= x;
  put_rec(name, 0, d, a.get_ptr_cpu(), n * 8);
}

static const char *tname[5] = {"tri", "quad", "tet", "pri", "hex"};

// which element (type, element, local face, face-local point) a flux-point pointer of the reference's interface tables points at
struct fpt_ref
{
  int etype, ele, loc, j;
};
static fpt_ref locate(struct solution *S, double *p)
{
  for (int t = 0; t < S->n_ele_types; t++)
  {
    eles *e = S->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    double *base = e->disu_fpts.get_ptr_cpu();
    long long n = (long long)e->n_fpts_per_ele * e->n_eles; // field 0
    if (p >= base && p < base + n)
    {
      long long flat = p - base;
      fpt_ref r;
      r.etype = t;
      r.ele = (int)(flat / e->n_fpts_per_ele);
      int fp = (int)(flat % e->n_fpts_per_ele), off = 0;
      for (r.loc = 0; r.loc < e->n_inters_per_ele; r.loc++)
      {
        if (fp < off + e->n_fpts_per_inter(r.loc)) break;
        off += e->n_fpts_per_inter(r.loc);
      }
      r.j = fp - off;
      return r;
    }
  }
  fprintf(stderr, "ref_gpu: an interface pointer does not point into any disu_fpts array\n");
  exit(1);
}

int main(int argc, char *argv[])
{
  if (argc < 4) { fprintf(stderr, "usage: ref_gpu <input> <out.hfd> <n_steps> [mode]\n"); return 2; }
  const int n_steps = atoi(argv[3]);
  const int mode = argc > 4 ? atoi(argv[4]) : 1;
  struct solution FlowSol;
  mesh *mesh_data = new mesh();
  run_input.setup(argv[1], 0);
  SetInput(&FlowSol);
  GeoPreprocess(&FlowSol, *mesh_data);
  delete mesh_data;
  InitSolution(&FlowSol);
  if (run_input.LES || run_input.shock_cap || run_input.wall_model) { fprintf(stderr, "ref_gpu: LES / shock capturing / wall model are not wired in this test driver\n"); return 2; }

  hf_ctx *ctx = nullptr;
  CK(hf_dev_create(&ctx, 0, 0, 1));

  // ---- run_input -> hf_params, bc_list -> hf_bc table --------------------------------------------------------------------------
  const int n_rk = run_input.adv_type == 0 ? 1 : (run_input.adv_type <= 2 ? 4 : (run_input.adv_type == 3 ? 5 : 14));
  hf_params p;
  memset(&p, 0, sizeof(p));
  p.equation = run_input.equation;
  p.viscous = run_input.viscous;
  p.n_dims = FlowSol.n_dims;
  p.n_fields = (run_input.equation == 0) ? FlowSol.n_dims + 2 : 1;
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
  p.R_ref = run_input.viscous ? run_input.R_ref : run_input.R_gas; // src/bdy_inters.cpp:368-369
  for (int i = 0; i < 3; i++) p.wave_speed[i] = run_input.wave_speed.get_dim(0) >= 3 ? run_input.wave_speed(i) : 0.;
  p.diff_coeff = run_input.diff_coeff;
  p.lambda = run_input.lambda;
  p.n_rk = n_rk;
  if (run_input.adv_type == 3 || run_input.adv_type == 4)
    for (int i = 0; i < n_rk && i < HF_MAX_RK; i++) { p.RK_a[i] = run_input.RK_a(i); p.RK_b[i] = run_input.RK_b(i); }
  p.over_int = run_input.over_int;
  CK(hf_dev_set_params(ctx, &p));
  {
    const int nb = run_input.bc_list.get_dim(0);
    vector<hf_bc> table(nb);
    for (int i = 0; i < nb; i++)
    {
      bc &b = run_input.bc_list(i);
      hf_bc &t = table[i];
      memset(&t, 0, sizeof(t));
      t.bc_flag = b.get_bc_flag();
      t.rho = b.rho;
      for (int k = 0; k < 3; k++) t.velocity[k] = b.velocity.get_dim(0) >= 3 ? b.velocity(k) : 0.;
      t.p_static = b.p_static; t.T_static = b.T_static; t.p_total = b.p_total; t.T_total = b.T_total;
      t.mach = b.mach; t.nx = b.nx; t.ny = b.ny; t.nz = b.nz;
      t.use_wm = 0;
      // the ramp fields are only set for this kind (src/input.cpp)
      if (t.bc_flag == SUB_IN_CHAR && b.pressure_ramp) { fprintf(stderr, "ref_gpu: ramped inlets are not wired in this test driver\n"); return 2; }
    }
    CK(hf_dev_set_bc_table(ctx, nb, nb ? table.data() : nullptr));
  }
  CK(hf_dev_set_mode(ctx, mode));

  // ---- eles members -> hf_eles_desc -----------------------------------------------------------------------------------------------
  for (int t = 0; t < FlowSol.n_ele_types; t++)
  {
    eles *e = FlowSol.mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    hf_eles_desc d;
    memset(&d, 0, sizeof(d));
    d.ele_type = t;
    d.n_eles = e->n_eles;
    d.n_upts_per_ele = e->n_upts_per_ele;
    d.n_fpts_per_ele = e->n_fpts_per_ele;
    d.n_dims = e->n_dims;
    d.n_fields = e->n_fields;
    d.order = e->order;
    d.n_inters_per_ele = e->n_inters_per_ele;
    d.n_fpts_per_inter = e->n_fpts_per_inter.get_ptr_cpu();
    d.opp_0 = e->opp_0.get_ptr_cpu();
    d.opp_3 = e->opp_3.get_ptr_cpu();
    for (int i = 0; i < e->n_dims; i++)
    {
      d.opp_1[i] = e->opp_1(i).get_ptr_cpu();
      d.opp_2[i] = e->opp_2(i).get_ptr_cpu();
      if (run_input.viscous)
      {
        d.opp_4[i] = e->opp_4(i).get_ptr_cpu();
        d.opp_5[i] = e->opp_5(i).get_ptr_cpu();
      }
    }
    if (run_input.viscous) d.opp_6 = e->opp_6.get_ptr_cpu();
    d.detjac_upts = e->detjac_upts.get_ptr_cpu();
    d.JGinv_upts = e->JGinv_upts.get_ptr_cpu();
    d.detjac_fpts = e->detjac_fpts.get_ptr_cpu();
    d.JGinv_fpts = e->JGinv_fpts.get_ptr_cpu();
    d.tdA_fpts = e->tdA_fpts.get_ptr_cpu();
    d.norm_fpts = e->norm_fpts.get_ptr_cpu();
    d.h_ref = run_input.dt_type > 0 ? e->h_ref.get_ptr_cpu() : nullptr;
    d.disu_upts0 = e->disu_upts(0).get_ptr_cpu();
    if (run_input.over_int)
    {
      d.n_over_int_cubpts = e->opp_over_int_cubpts.get_dim(0);
      d.opp_over_int_cubpts = e->opp_over_int_cubpts.get_ptr_cpu();
      d.over_int_filter = e->over_int_filter.get_ptr_cpu();
      d.JGinv_over_int_cubpts = e->JGinv_over_int_cubpts.get_ptr_cpu();
    }
    CK(hf_dev_upload_eles(ctx, &d));
  }

  // ---- interior interfaces: (element, local face, rotation) back from the pointer tables -------------------------------------------
  for (int t = 0; t < FlowSol.n_int_inter_types; t++)
  {
    int_inters &I = FlowSol.mesh_int_inters(t);
    if (I.n_inters == 0) continue;
    const int ni = I.n_inters, nf = I.n_fpts_per_inter;
    vector<int> tl(ni), el(ni), ll(ni), tr(ni), er(ni), lr(ni), rot(ni);
    const int n_rot = (I.inters_type == 0) ? 1 : (I.inters_type == 1 ? 3 : 4);
    for (int i = 0; i < ni; i++)
    {
      fpt_ref L = locate(&FlowSol, I.disu_fpts_l(0, i, 0));
      tl[i] = L.etype; el[i] = L.ele; ll[i] = L.loc;
      vector<int> right(nf);
      fpt_ref R0 = locate(&FlowSol, I.disu_fpts_r(0, i, 0));
      tr[i] = R0.etype; er[i] = R0.ele; lr[i] = R0.loc;
      for (int j = 0; j < nf; j++)
      {
        fpt_ref Lj = locate(&FlowSol, I.disu_fpts_l(j, i, 0));
        if (Lj.ele != L.ele || Lj.loc != L.loc || Lj.j != j) { fprintf(stderr, "ref_gpu: unexpected left numbering\n"); return 1; }
        fpt_ref Rj = locate(&FlowSol, I.disu_fpts_r(j, i, 0));
        if (Rj.ele != R0.ele || Rj.loc != R0.loc || Rj.etype != R0.etype) { fprintf(stderr, "ref_gpu: right side spans two faces\n"); return 1; }
        right[j] = Rj.j;
      }
      int found = -1;
      for (int r = 0; r < n_rot && found < 0; r++)
      {
        I.get_lut(r); // the reference's own table for this rotation (src/inters.cpp:153-262)
        bool same = true;
        for (int j = 0; j < nf && same; j++) same = I.lut(j) == right[j];
        if (same) found = r;
      }
      if (found < 0) { fprintf(stderr, "ref_gpu: no rotation tag reproduces interface %d of type %d\n", i, t); return 1; }
      rot[i] = found;
    }
    hf_int_inters_desc d;
    d.inter_type = I.inters_type; d.n_inters = ni; d.n_fpts_per_inter = nf;
    d.ele_type_l = tl.data(); d.ele_l = el.data(); d.local_inter_l = ll.data();
    d.ele_type_r = tr.data(); d.ele_r = er.data(); d.local_inter_r = lr.data();
    d.rot_tag = rot.data();
    CK(hf_dev_upload_int_inters(ctx, &d));
  }
  // ---- boundary interfaces ------------------------------------------------------------------------------------------------------
  for (int t = 0; t < FlowSol.n_bdy_inter_types; t++)
  {
    bdy_inters &I = FlowSol.mesh_bdy_inters(t);
    if (I.n_inters == 0) continue;
    const int ni = I.n_inters, nf = I.n_fpts_per_inter, nd = FlowSol.n_dims;
    vector<int> tl(ni), el(ni), ll(ni), bid(ni);
    vector<double> pos((size_t)nf * ni * nd);
    for (int i = 0; i < ni; i++)
    {
      fpt_ref L = locate(&FlowSol, I.disu_fpts_l(0, i, 0));
      tl[i] = L.etype; el[i] = L.ele; ll[i] = L.loc;
      bid[i] = I.boundary_id(i);
      for (int j = 0; j < nf; j++)
        for (int k = 0; k < nd; k++) pos[j + (size_t)nf * (i + (size_t)ni * k)] = *I.pos_fpts(j, i, k);
    }
    hf_bdy_inters_desc d;
    memset(&d, 0, sizeof(d));
    d.inter_type = I.inters_type; d.n_inters = ni; d.n_fpts_per_inter = nf;
    d.ele_type_l = tl.data(); d.ele_l = el.data(); d.local_inter_l = ll.data();
    d.bc_id = bid.data();
    d.pos_fpts = pos.data();
    CK(hf_dev_upload_bdy_inters(ctx, &d));
  }
  CK(hf_dev_finalize_setup(ctx));
  printf("ref_gpu: fused hexahedron kernels: %s | blocked element kernels: %s | mode %d\n", hf_dev_fused_status(ctx), hf_dev_elem_status(ctx), mode);

  // ---- the reference's main loop (src/HiFiLES.cpp:199-225) with the device behind CalcResidual + AdvanceSolution --------------------
  g_out = fopen(argv[2], "wb");
  if (!g_out) { perror("open out"); return 1; }
  vector<double> hist;
  const int nfld = p.n_fields;
  for (int it = 0; it < n_steps; it++)
  {
    if (run_input.dt_type != 0)
    {
      double dt = 0.;
      CK(hf_dev_calc_dt(ctx, &dt)); // calc_time_step
      if (run_input.dt_type == 1) run_input.dt = dt;
    }
    for (int i = 0; i < n_rk; i++) CK(hf_dev_rk_stage(ctx, i, FlowSol.time, i == n_rk - 1 ? 1 : 0));
    FlowSol.time += run_input.dt;
    run_input.time = FlowSol.time;
    // output::CalcNormResidual (src/output.cpp:2166-2248), serial
    double sums[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    CK(hf_dev_residual_norm(ctx, run_input.res_norm_type, sums));
    long long np = 0;
    for (int t = 0; t < FlowSol.n_ele_types; t++)
      if (FlowSol.mesh_eles(t)->get_n_eles() != 0) np += (long long)FlowSol.mesh_eles(t)->get_n_eles() * FlowSol.mesh_eles(t)->get_n_upts_per_ele();
    for (int f = 0; f < nfld; f++)
    {
      double s = sums[f];
      if (run_input.res_norm_type == 1) s = s / np;
      else if (run_input.res_norm_type == 2) s = sqrt(s) / np;
      hist.push_back(s);
    }
  }
  // eles::cp_disu_upts_gpu_cpu: back into the reference's own arrays
  for (int t = 0; t < FlowSol.n_ele_types; t++)
  {
    eles *e = FlowSol.mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    const size_t n = (size_t)e->n_upts_per_ele * e->n_eles * e->n_fields;
    CK(hf_dev_download(ctx, t, HF_DISU_UPTS0, e->disu_upts(0).get_ptr_cpu(), n));
    put(string("final.") + tname[t] + ".disu_upts", e->disu_upts(0));
  }
  if (n_steps > 0) put_rec("history.norm_residual", 0, {(long long)nfld, (long long)n_steps}, hist.data(), hist.size() * 8);
  fclose(g_out);
  CK(hf_dev_destroy(ctx));
  printf("ref_gpu: wrote %s\n", argv[2]);
  return 0;
}
