// Element base class: storage, operator construction, metric terms, initial conditions on the host; every
// per-stage method is a single call into the device layer.
//   storage         reference src/eles.cpp:58-228
//   set_ics         reference src/eles.cpp:237-531
//   set_opp_0..6    reference src/eles.cpp:3074-3596
//   set_transforms  reference src/eles.cpp:4015-4393   (arithmetic order kept: the LDG beta switch tests the sign
//                   of rounding-level normal components, reference src/inters.cpp:570-581, so the metrics have to
//                   be bit-identical, not merely close)
#include "hifiles.h"

using namespace std;

void hf_check(int status)
{
  if (status != 0) FatalError(string("device layer: ") + hf_dev_last_error());
}

eles::eles()
{
  ctx = nullptr;
  rank = 0;
  ele_type = -1;
  n_eles = 0;
  n_dims = n_fields = order = viscous = n_inters_per_ele = 0;
  n_upts_per_ele = n_fpts_per_ele = max_n_spts_per_ele = n_adv_levels = upts_type = 0;
}

void eles::setup(int in_n_eles, int in_max_n_spts_per_ele)
{
  n_eles = in_n_eles;
  max_n_spts_per_ele = in_max_n_spts_per_ele;
  if (n_eles == 0) return;

  order = run_input.order;
  viscous = run_input.viscous;
  setup_ele_type_specific();
  if (run_input.over_int) set_over_int();
  if (run_input.shock_cap) set_shock_capture();
  if (run_input.LES && viscous && (run_input.SGS_model == 2 || run_input.SGS_model == 3 || run_input.SGS_model == 4)) compute_filter_upts();

  if (run_input.adv_type == 0) n_adv_levels = 1;
  else if (run_input.adv_type >= 1 && run_input.adv_type <= 4) n_adv_levels = 2;
  else FatalError("ERROR: Type of time integration scheme not recongized ... ");

  disu_upts.setup(n_adv_levels);
  for (int i = 0; i < n_adv_levels; i++) disu_upts(i).setup(n_upts_per_ele, n_eles, n_fields); // zero-initialised
  if (run_input.dt_type == 2) dt_local.setup(n_eles);
  src_upts.setup(n_upts_per_ele, n_eles, n_fields);
  set_shape(in_max_n_spts_per_ele);
  d_nodal_s_basis.setup(max_n_spts_per_ele, n_dims);
  ele2global_ele.setup(n_eles);
  bcid.setup(n_eles, n_inters_per_ele);
  div_tconf_upts.setup(1);
  div_tconf_upts(0).setup(n_upts_per_ele, n_eles, n_fields);
}

void eles::set_shape(int in_max_n_spts_per_ele)
{
  shape.setup(n_dims, in_max_n_spts_per_ele, n_eles);
  n_spts_per_ele.setup(n_eles);
}

void eles::set_shape_node(int in_spt, int in_ele, hf_array<double> &in_pos)
{
  for (int i = 0; i < n_dims; i++) shape(i, in_spt, in_ele) = in_pos(i);
}

int eles::get_fpt_index(int in_inter_local_fpt, int in_ele_local_inter)
{
  int fpt = in_inter_local_fpt;
  for (int i = 0; i < in_ele_local_inter; i++) fpt += n_fpts_per_inter(i);
  return fpt;
}

// ---- initial conditions --------------------------------------------------------------------------------------
void eles::set_ics(double &time)
{
  double rho, vx, vy, vz, p;
  double gamma = run_input.gamma;
  time = 0.;
  hf_array<double> pos(n_dims), ics(n_fields);

  for (int i = 0; i < n_eles; i++)
  {
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      for (int k = 0; k < n_dims; k++) pos(k) = pos_upts(j, i, k);

      if (run_input.ic_form == 0) // isentropic vortex
      {
        eval_isentropic_vortex(pos, time, rho, vx, vy, vz, p, n_dims);
        ics(0) = rho;
        ics(1) = rho * vx;
        ics(2) = rho * vy;
        if (n_dims == 2)
          ics(3) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy)));
        else
        {
          ics(3) = rho * vz;
          ics(4) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy) + (vz * vz)));
        }
      }
      else if (run_input.ic_form == 1) // uniform flow
      {
        rho = run_input.rho_c_ic;
        vx = run_input.u_c_ic;
        vy = run_input.v_c_ic;
        vz = run_input.w_c_ic;
        p = run_input.p_c_ic;
        ics(0) = rho;
        ics(1) = rho * vx;
        ics(2) = rho * vy;
        if (n_dims == 2)
          ics(3) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy)));
        else
        {
          ics(3) = rho * vz;
          ics(4) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy) + (vz * vz)));
        }
      }
      else if (run_input.ic_form == 7) // Taylor-Green vortex
      {
        double V_0 = run_input.uvw_c_ic / run_input.uvw_ref;
        if (n_dims == 2)
        {
          p = run_input.p_c_ic + run_input.rho_c_ic * pow(V_0, 2) / 4.0 * (cos(2.0 * pos(0)) + cos(2.0 * pos(1)));
          ics(0) = p / (run_input.R_ref * run_input.T_c_ic);
          ics(1) = ics(0) * V_0 * sin(pos(0)) * cos(pos(1));
          ics(2) = -ics(0) * V_0 * cos(pos(0)) * sin(pos(1));
          ics(3) = p / (gamma - 1.0) + 0.5 * (ics(1) * ics(1) + ics(2) * ics(2)) / ics(0);
        }
        else
        {
          p = run_input.p_c_ic + run_input.rho_c_ic * pow(V_0, 2) / 16.0 * (cos(2.0 * pos(0)) + cos(2.0 * pos(1))) * (cos(2.0 * pos(2)) + 2.0);
          ics(0) = p / (run_input.R_ref * run_input.T_c_ic);
          ics(1) = ics(0) * V_0 * sin(pos(0)) * cos(pos(1)) * cos(pos(2));
          ics(2) = -ics(0) * V_0 * cos(pos(0)) * sin(pos(1)) * cos(pos(2));
          ics(3) = 0.0;
          ics(4) = p / (gamma - 1.0) + 0.5 * (ics(1) * ics(1) + ics(2) * ics(2) + ics(3) * ics(3)) / ics(0);
        }
      }
      else if (run_input.ic_form == 9) // stationary shock
      {
        int found = 0;
        for (size_t k = 0; k < run_input.bc_list.size(); k++)
        {
          bc &b = run_input.bc_list[k];
          if (b.get_bc_flag() == SUP_IN || b.get_bc_flag() == CHAR)
          {
            if (pos(0) <= run_input.x_shock_ic)
            {
              rho = b.rho; vx = b.velocity(0); vy = b.velocity(1); vz = b.velocity(2); p = b.p_static;
            }
            else
            {
              rho = run_input.rho_c_ic; vx = run_input.u_c_ic; vy = run_input.v_c_ic; vz = run_input.w_c_ic; p = run_input.p_c_ic;
            }
            found = 1;
            break;
          }
        }
        if (found == 0) FatalError("Must have a Sup_In or Char boundary condition");
        ics(0) = rho;
        ics(1) = rho * vx;
        ics(2) = rho * vy;
        if (n_dims == 2)
          ics(3) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy)));
        else
        {
          ics(3) = rho * vz;
          ics(4) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy) + (vz * vz)));
        }
      }
      else if (run_input.ic_form == 10) // shock tube
      {
        vx = vy = vz = 0.;
        if (pos(0) <= run_input.x_shock_ic)
        {
          if (run_input.viscous) { p = 100000. / run_input.p_ref; rho = 1.0 / run_input.rho_ref; }
          else { p = 100000.; rho = 1.0; }
        }
        else
        {
          if (run_input.viscous) { p = 10000. / run_input.p_ref; rho = 0.125 / run_input.rho_ref; }
          else { p = 10000.; rho = 0.125; }
        }
        ics(0) = rho;
        ics(1) = rho * vx;
        ics(2) = rho * vy;
        if (n_dims == 2)
          ics(3) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy)));
        else
        {
          ics(3) = rho * vz;
          ics(4) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy) + (vz * vz)));
        }
      }
      else if (run_input.ic_form == 2 || run_input.ic_form == 3) // advection-diffusion: sine wave, single / product
      {
        hf_array<double> grad_rho(n_dims);
        if (run_input.ic_form == 2)
          eval_sine_wave_single(pos, run_input.wave_speed, run_input.diff_coeff, time, rho, grad_rho, n_dims);
        else
          eval_sine_wave_group(pos, run_input.wave_speed, run_input.diff_coeff, time, rho, grad_rho, n_dims);
        ics(0) = rho;
      }
      else if (run_input.ic_form == 4) // advection-diffusion: Gaussian pulse
      {
        eval_sphere_wave(pos, run_input.wave_speed, time, rho, n_dims);
        ics(0) = rho;
      }
      else if (run_input.ic_form == 5) // advection-diffusion: constant
        ics(0) = run_input.rho_c_ic;
      else if (run_input.ic_form == 6) // polynomial velocity profile: the reference's eval_poly_ic stops here too (src/funcs.cpp:1928)
        FatalError("Function deprecated!");
      else
        FatalError("ERROR: Invalid form of initial condition ...");

      if (run_input.perturb_ic == 1 && n_dims == 3)
      {
        double alpha = 0.1, L_x = 2. * pi, L_y = pi, L_z = 2.;
        ics(3) += alpha * exp(-pow((pos(0) - L_x / 2.) / L_x, 2)) * exp(-pow(pos(1) / L_y, 2)) * cos(4. * pi * pos(2) / L_z);
      }

      for (int k = 0; k < n_fields; k++) disu_upts(0)(j, i, k) = ics(k);
    }
  }

  set_h_ref();
}

// element reference lengths for the CFL time step (computed after the initial data are set, from the analytic field or
// from a restart file: reference src/eles.cpp:470-486, 732-749)
void eles::set_h_ref()
{
  if (run_input.dt_type > 0)
  {
    h_ref.setup(n_eles);
    for (int i = 0; i < n_eles; i++) h_ref(i) = calc_h_ref_specific(i);
  }
  else
    h_ref.setup(1);
}

double eles::calc_h_ref_specific(int e)
{
  // minimum corner-to-corner edge length (reference src/eles_quads.cpp:1287-1301, src/eles_hexas.cpp:1551-1571);
  // the shape slots of a linear element are in tensor order, so the edges are the slot pairs below.
  auto len = [&](int a, int b) {
    double s = 0.;
    for (int d = 0; d < n_dims; d++) s += pow(shape(d, a, e) - shape(d, b, e), 2.0);
    return sqrt(s);
  };
  double h = 1e300;
  if (ele_type == QUAD)
  {
    const int ed[4][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}};
    for (auto &p : ed) h = min(h, len(p[0], p[1]));
  }
  else if (ele_type == HEX)
  {
    const int ed[12][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}, {4, 5}, {5, 7}, {7, 6}, {6, 4}, {1, 5}, {3, 7}, {0, 4}, {2, 6}};
    for (auto &p : ed) h = min(h, len(p[0], p[1]));
  }
  else
    FatalError("h_ref not available for this element type");
  return h;
}

// ---- operators -------------------------------------------------------------------------------------------------
void eles::set_opp_0(int)
{
  hf_array<double> loc(n_dims);
  opp_0.setup(n_fpts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int j = 0; j < n_fpts_per_ele; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = tloc_fpts(k, j);
      opp_0(j, i) = eval_nodal_basis(i, loc);
    }
}

void eles::set_opp_1(int)
{
  hf_array<double> loc(n_dims);
  opp_1.setup(n_dims);
  for (int i = 0; i < n_dims; i++) opp_1(i).setup(n_fpts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_dims; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
      for (int k = 0; k < n_fpts_per_ele; k++)
      {
        for (int l = 0; l < n_dims; l++) loc(l) = tloc_fpts(l, k);
        opp_1(i)(k, j) = eval_nodal_basis(j, loc) * tnorm_fpts(i, k);
      }
}

void eles::set_opp_2(int)
{
  hf_array<double> loc(n_dims);
  opp_2.setup(n_dims);
  for (int i = 0; i < n_dims; i++) opp_2(i).setup(n_upts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_dims; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
      for (int k = 0; k < n_upts_per_ele; k++)
      {
        for (int l = 0; l < n_dims; l++) loc(l) = loc_upts(l, k);
        opp_2(i)(k, j) = eval_d_nodal_basis(j, i, loc);
      }
}

void eles::set_opp_3(int)
{
  opp_3.setup(n_upts_per_ele, n_fpts_per_ele);
  fill_opp_3(opp_3);
}

void eles::set_opp_4(int)
{
  hf_array<double> loc(n_dims);
  opp_4.setup(n_dims);
  for (int i = 0; i < n_dims; i++) opp_4(i).setup(n_upts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_dims; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
      for (int k = 0; k < n_upts_per_ele; k++)
      {
        for (int l = 0; l < n_dims; l++) loc(l) = loc_upts(l, k);
        opp_4(i)(k, j) = eval_d_nodal_basis(j, i, loc);
      }
}

void eles::set_opp_5(int)
{
  opp_5.setup(n_dims);
  for (int i = 0; i < n_dims; i++) opp_5(i).setup(n_upts_per_ele, n_fpts_per_ele);
  for (int i = 0; i < n_dims; i++)
    for (int j = 0; j < n_fpts_per_ele; j++)
      for (int k = 0; k < n_upts_per_ele; k++)
        opp_5(i)(k, j) = opp_3(k, j) * tnorm_fpts(i, j);
}

void eles::set_opp_6(int)
{
  hf_array<double> loc(n_dims);
  opp_6.setup(n_fpts_per_ele, n_upts_per_ele);
  for (int j = 0; j < n_upts_per_ele; j++)
    for (int l = 0; l < n_fpts_per_ele; l++)
    {
      for (int m = 0; m < n_dims; m++) loc(m) = tloc_fpts(m, l);
      opp_6(l, j) = eval_nodal_basis(j, loc);
    }
}

// ---- metrics ---------------------------------------------------------------------------------------------------
void eles::calc_pos(hf_array<double> &in_loc, int in_ele, hf_array<double> &out_pos)
{
  for (int i = 0; i < n_dims; i++)
  {
    out_pos(i) = 0.0;
    for (int j = 0; j < n_spts_per_ele(in_ele); j++)
      out_pos(i) += eval_nodal_s_basis(j, in_loc, n_spts_per_ele(in_ele)) * shape(i, j, in_ele);
  }
}

void eles::calc_d_pos(hf_array<double> &in_loc, int in_ele, hf_array<double> &out_d_pos)
{
  eval_d_nodal_s_basis(d_nodal_s_basis, in_loc, n_spts_per_ele(in_ele));
  for (int j = 0; j < n_dims; j++)
    for (int k = 0; k < n_dims; k++)
    {
      out_d_pos(j, k) = 0.0;
      for (int i = 0; i < n_spts_per_ele(in_ele); i++)
        out_d_pos(j, k) += d_nodal_s_basis(i, k) * shape(j, i, in_ele);
    }
}

void eles::set_transforms()
{
  if (n_eles == 0) return;
  set_transforms_upts();
  if (run_input.over_int) set_transforms_over_int_cubpts();
  set_transforms_fpts();
  // metrics at the volume cubature points: only needed for computing error and integral diagnostic quantities (reference src/eles.cpp:4026-4028)
  // metrics at the interface cubature points: only when surface forces are asked for (reference src/eles.cpp:4022-4024)
  if (run_input.calc_force != 0) set_inters_cubpts_and_transforms();
  if (run_input.test_case != 0 || run_input.n_integral_quantities != 0)
  {
    set_volume_cubpts(order, loc_volume_cubpts, weight_volume_cubpts);
    set_opp_volume_cubpts();
    set_transforms_vol_cubpts();
  }
}

namespace
{
// The shape-basis values depend only on the reference location and the number of shape nodes; the reference
// re-evaluates them for every element (the start-up bottleneck at 64^3).  Cache them per
// (point, n_spts): the sums below then use the very same factors in the very same order.
struct basis_cache
{
  int n_spts = -1;
  vector<double> s;  // [pt][spt]
  vector<double> ds; // [pt][spt][dim]
};
} // namespace

static void fill_basis_cache(eles *e, hf_array<double> &locs, int n_pts, int n_spts, basis_cache &c)
{
  int nd = e->n_dims;
  c.n_spts = n_spts;
  c.s.assign((size_t)n_pts * n_spts, 0.);
  c.ds.assign((size_t)n_pts * n_spts * nd, 0.);
  hf_array<double> loc(nd), d(e->max_n_spts_per_ele, nd);
  for (int p = 0; p < n_pts; p++)
  {
    for (int k = 0; k < nd; k++) loc(k) = locs(k, p);
    for (int j = 0; j < n_spts; j++) c.s[(size_t)p * n_spts + j] = e->eval_nodal_s_basis(j, loc, n_spts);
    e->eval_d_nodal_s_basis(d, loc, n_spts);
    for (int j = 0; j < n_spts; j++)
      for (int k = 0; k < nd; k++) c.ds[((size_t)p * n_spts + j) * nd + k] = d(j, k);
  }
}

void eles::set_transforms_upts()
{
  detjac_upts.setup(n_upts_per_ele, n_eles);
  JGinv_upts.setup(n_dims, n_dims, n_upts_per_ele, n_eles);
  pos_upts.setup(n_upts_per_ele, n_eles, n_dims);
  basis_cache c;
  double dp[3][3];
  for (int i = 0; i < n_eles; i++)
  {
    int ns = n_spts_per_ele(i);
    if (ns != c.n_spts) fill_basis_cache(this, loc_upts, n_upts_per_ele, ns, c);
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      const double *s = &c.s[(size_t)j * ns];
      const double *ds = &c.ds[(size_t)j * ns * n_dims];
      for (int d = 0; d < n_dims; d++)
      {
        double acc = 0.0;
        for (int q = 0; q < ns; q++) acc += s[q] * shape(d, q, i);
        pos_upts(j, i, d) = acc;
      }
      for (int a = 0; a < n_dims; a++)
        for (int b = 0; b < n_dims; b++)
        {
          double acc = 0.0;
          for (int q = 0; q < ns; q++) acc += ds[q * n_dims + b] * shape(a, q, i);
          dp[a][b] = acc;
        }
      if (n_dims == 2)
      {
        double xr = dp[0][0], xs = dp[0][1], yr = dp[1][0], ys = dp[1][1];
        detjac_upts(j, i) = xr * ys - xs * yr;
        if (detjac_upts(j, i) < 0) FatalError("Negative Jacobian at solution points");
        JGinv_upts(0, 0, j, i) = ys;
        JGinv_upts(0, 1, j, i) = -xs;
        JGinv_upts(1, 0, j, i) = -yr;
        JGinv_upts(1, 1, j, i) = xr;
      }
      else
      {
        double xr = dp[0][0], xs = dp[0][1], xt = dp[0][2];
        double yr = dp[1][0], ys = dp[1][1], yt = dp[1][2];
        double zr = dp[2][0], zs = dp[2][1], zt = dp[2][2];
        detjac_upts(j, i) = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr);
        JGinv_upts(0, 0, j, i) = ys * zt - yt * zs;
        JGinv_upts(0, 1, j, i) = xt * zs - xs * zt;
        JGinv_upts(0, 2, j, i) = xs * yt - xt * ys;
        JGinv_upts(1, 0, j, i) = yt * zr - yr * zt;
        JGinv_upts(1, 1, j, i) = xr * zt - xt * zr;
        JGinv_upts(1, 2, j, i) = xt * yr - xr * yt;
        JGinv_upts(2, 0, j, i) = yr * zs - ys * zr;
        JGinv_upts(2, 1, j, i) = xs * zr - xr * zs;
        JGinv_upts(2, 2, j, i) = xr * ys - xs * yr;
      }
    }
  }
}

void eles::set_transforms_fpts()
{
  detjac_fpts.setup(n_fpts_per_ele, n_eles);
  JGinv_fpts.setup(n_dims, n_dims, n_fpts_per_ele, n_eles);
  tdA_fpts.setup(n_fpts_per_ele, n_eles);
  norm_fpts.setup(n_fpts_per_ele, n_eles, n_dims);
  pos_fpts.setup(n_fpts_per_ele, n_eles, n_dims);
  if (run_input.LES) Jacobian_fpts.setup(n_dims, n_dims, n_fpts_per_ele, n_eles);
  basis_cache c;
  double dp[3][3], tn[3];
  for (int i = 0; i < n_eles; i++)
  {
    int ns = n_spts_per_ele(i);
    if (ns != c.n_spts) fill_basis_cache(this, tloc_fpts, n_fpts_per_ele, ns, c);
    for (int j = 0; j < n_fpts_per_ele; j++)
    {
      const double *s = &c.s[(size_t)j * ns];
      const double *ds = &c.ds[(size_t)j * ns * n_dims];
      for (int d = 0; d < n_dims; d++)
      {
        double acc = 0.0;
        for (int q = 0; q < ns; q++) acc += s[q] * shape(d, q, i);
        pos_fpts(j, i, d) = acc;
      }
      for (int a = 0; a < n_dims; a++)
        for (int b = 0; b < n_dims; b++)
        {
          double acc = 0.0;
          for (int q = 0; q < ns; q++) acc += ds[q * n_dims + b] * shape(a, q, i);
          dp[a][b] = acc;
          if (run_input.LES) Jacobian_fpts(a, b, j, i) = acc;
        }
      if (n_dims == 2)
      {
        double xr = dp[0][0], xs = dp[0][1], yr = dp[1][0], ys = dp[1][1];
        detjac_fpts(j, i) = xr * ys - xs * yr;
        if (detjac_fpts(j, i) < 0) FatalError("Negative Jacobian at flux points");
        JGinv_fpts(0, 0, j, i) = ys;
        JGinv_fpts(0, 1, j, i) = -xs;
        JGinv_fpts(1, 0, j, i) = -yr;
        JGinv_fpts(1, 1, j, i) = xr;
        tn[0] = (tnorm_fpts(0, j) * dp[1][1]) - (tnorm_fpts(1, j) * dp[1][0]);
        tn[1] = -(tnorm_fpts(0, j) * dp[0][1]) + (tnorm_fpts(1, j) * dp[0][0]);
        tdA_fpts(j, i) = sqrt(tn[0] * tn[0] + tn[1] * tn[1]);
        norm_fpts(j, i, 0) = tn[0] / tdA_fpts(j, i);
        norm_fpts(j, i, 1) = tn[1] / tdA_fpts(j, i);
      }
      else
      {
        double xr = dp[0][0], xs = dp[0][1], xt = dp[0][2];
        double yr = dp[1][0], ys = dp[1][1], yt = dp[1][2];
        double zr = dp[2][0], zs = dp[2][1], zt = dp[2][2];
        detjac_fpts(j, i) = xr * (ys * zt - yt * zs) - xs * (yr * zt - yt * zr) + xt * (yr * zs - ys * zr);
        JGinv_fpts(0, 0, j, i) = ys * zt - yt * zs;
        JGinv_fpts(0, 1, j, i) = xt * zs - xs * zt;
        JGinv_fpts(0, 2, j, i) = xs * yt - xt * ys;
        JGinv_fpts(1, 0, j, i) = yt * zr - yr * zt;
        JGinv_fpts(1, 1, j, i) = xr * zt - xt * zr;
        JGinv_fpts(1, 2, j, i) = xt * yr - xr * yt;
        JGinv_fpts(2, 0, j, i) = yr * zs - ys * zr;
        JGinv_fpts(2, 1, j, i) = xs * zr - xr * zs;
        JGinv_fpts(2, 2, j, i) = xr * ys - xs * yr;
        double t0 = tnorm_fpts(0, j), t1 = tnorm_fpts(1, j), t2 = tnorm_fpts(2, j);
        tn[0] = ((t0 * (dp[1][1] * dp[2][2] - dp[1][2] * dp[2][1])) + (t1 * (dp[1][2] * dp[2][0] - dp[1][0] * dp[2][2])) + (t2 * (dp[1][0] * dp[2][1] - dp[1][1] * dp[2][0])));
        tn[1] = ((t0 * (dp[0][2] * dp[2][1] - dp[0][1] * dp[2][2])) + (t1 * (dp[0][0] * dp[2][2] - dp[0][2] * dp[2][0])) + (t2 * (dp[0][1] * dp[2][0] - dp[0][0] * dp[2][1])));
        tn[2] = ((t0 * (dp[0][1] * dp[1][2] - dp[0][2] * dp[1][1])) + (t1 * (dp[0][2] * dp[1][0] - dp[0][0] * dp[1][2])) + (t2 * (dp[0][0] * dp[1][1] - dp[0][1] * dp[1][0])));
        tdA_fpts(j, i) = sqrt(tn[0] * tn[0] + tn[1] * tn[1] + tn[2] * tn[2]);
        norm_fpts(j, i, 0) = tn[0] / tdA_fpts(j, i);
        norm_fpts(j, i, 1) = tn[1] / tdA_fpts(j, i);
        norm_fpts(j, i, 2) = tn[2] / tdA_fpts(j, i);
      }
    }
  }
}

// ---- device mirror ---------------------------------------------------------------------------------------------
void eles::mv_all_cpu_gpu()
{
  if (n_eles == 0) return;
  hf_eles_desc d;
  memset(&d, 0, sizeof(d));
  d.ele_type = ele_type;
  d.n_eles = n_eles;
  d.n_upts_per_ele = n_upts_per_ele;
  d.n_fpts_per_ele = n_fpts_per_ele;
  d.n_dims = n_dims;
  d.n_fields = n_fields;
  d.order = order;
  d.n_inters_per_ele = n_inters_per_ele;
  d.n_fpts_per_inter = n_fpts_per_inter.get_ptr_cpu();
  d.opp_0 = opp_0.get_ptr_cpu();
  d.opp_3 = opp_3.get_ptr_cpu();
  for (int i = 0; i < n_dims; i++)
  {
    d.opp_1[i] = opp_1(i).get_ptr_cpu();
    d.opp_2[i] = opp_2(i).get_ptr_cpu();
    if (viscous)
    {
      d.opp_4[i] = opp_4(i).get_ptr_cpu();
      d.opp_5[i] = opp_5(i).get_ptr_cpu();
    }
  }
  if (viscous) d.opp_6 = opp_6.get_ptr_cpu();
  d.detjac_upts = detjac_upts.get_ptr_cpu();
  d.JGinv_upts = JGinv_upts.get_ptr_cpu();
  d.detjac_fpts = detjac_fpts.get_ptr_cpu();
  d.JGinv_fpts = JGinv_fpts.get_ptr_cpu();
  d.tdA_fpts = tdA_fpts.get_ptr_cpu();
  d.norm_fpts = norm_fpts.get_ptr_cpu();
  d.h_ref = (run_input.dt_type > 0) ? h_ref.get_ptr_cpu() : nullptr;
  d.disu_upts0 = disu_upts(0).get_ptr_cpu();
  if (run_input.over_int)
  {
    d.n_over_int_cubpts = loc_over_int_cubpts.get_dim(1);
    d.opp_over_int_cubpts = opp_over_int_cubpts.get_ptr_cpu();
    d.over_int_filter = over_int_filter.get_ptr_cpu();
    d.JGinv_over_int_cubpts = JGinv_over_int_cubpts.get_ptr_cpu();
  }
  if (run_input.LES)
  {
    static const double vol_factor[5] = {2., 4., 8. / 6., 4., 8.}; // tri, quad, tet, prism, hex: <type>::calc_ele_vol
    d.ele_vol_factor = vol_factor[ele_type];
    d.Jacobian_fpts = Jacobian_fpts.get_ptr_cpu();
    d.wall_distance = wall_distance.size() ? wall_distance.get_ptr_cpu() : nullptr;
    d.filter_upts = filter_upts.size() ? filter_upts.get_ptr_cpu() : nullptr;
  }
  if (run_input.shock_cap)
  {
    d.inv_vandermonde = modal_inv_vandermonde.get_ptr_cpu();
    d.sensor_w_top = sensor_w_top.get_ptr_cpu();
    d.sensor_w_all = sensor_w_all.get_ptr_cpu();
    d.exp_filter = exp_filter.get_ptr_cpu();
  }
  hf_check(hf_dev_upload_eles(ctx, &d));
  // surface forces and the gradient-based plot fields read grad_disu_upts of the monitored stage
  // (and so does the gradient error of the advection-diffusion test cases, eles::compute_error)
  if (viscous && (run_input.calc_force != 0 || run_input.n_diagnostic_fields != 0 || run_input.test_case == 2 || run_input.test_case == 3)) hf_check(hf_dev_set_keep_gradient(ctx, 1));
  if (run_input.n_integral_quantities != 0)
    hf_check(hf_dev_set_volume_cubature(ctx, ele_type, loc_volume_cubpts.get_dim(1), opp_volume_cubpts.get_ptr_cpu(), weight_volume_cubpts.get_ptr_cpu(),
                                        vol_detjac_vol_cubpts.get_ptr_cpu()));
}

void eles::cp_disu_upts_cpu_gpu()
{
  if (n_eles) hf_check(hf_dev_upload(ctx, ele_type, HF_DISU_UPTS0, disu_upts(0).get_ptr_cpu(), disu_upts(0).size()));
}
void eles::cp_disu_upts_gpu_cpu()
{
  if (n_eles) hf_check(hf_dev_download(ctx, ele_type, HF_DISU_UPTS0, disu_upts(0).get_ptr_cpu(), disu_upts(0).size()));
}
void eles::cp_div_tconf_upts_gpu_cpu()
{
  if (n_eles) hf_check(hf_dev_download(ctx, ele_type, HF_DIV_TCONF_UPTS, div_tconf_upts(0).get_ptr_cpu(), div_tconf_upts(0).size()));
}
void eles::cp_grad_disu_upts_gpu_cpu()
{
  if (!n_eles || !viscous) return;
  if (grad_disu_upts.size() == 0) grad_disu_upts.setup(n_upts_per_ele, n_eles, n_fields, n_dims);
  hf_check(hf_dev_download(ctx, ele_type, HF_GRAD_DISU_UPTS, grad_disu_upts.get_ptr_cpu(), grad_disu_upts.size()));
}
void eles::cp_src_upts_gpu_cpu()
{
  if (n_eles) hf_check(hf_dev_download(ctx, ele_type, HF_SRC_UPTS, src_upts.get_ptr_cpu(), src_upts.size()));
}
void eles::cp_array_gpu_cpu(int which, hf_array<double> &dst)
{
  if (n_eles) hf_check(hf_dev_download(ctx, ele_type, which, dst.get_ptr_cpu(), dst.size()));
}

void eles::extrapolate_solution() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EXTRAPOLATE_SOLUTION)); }
void eles::calculate_gradient() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_CALCULATE_GRADIENT)); }
void eles::evaluate_invFlux() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EVALUATE_INVFLUX)); }
void eles::evaluate_invFlux_over_int() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EVALUATE_INVFLUX_OVER_INT)); }
void eles::calc_sgs_terms() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_CALC_SGS_TERMS)); }
void eles::extrapolate_sgsFlux() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EXTRAPOLATE_SGSFLUX)); }
void eles::shock_capture() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_SHOCK_CAPTURE)); }
void eles::cp_sensor_gpu_cpu() { if (n_eles && run_input.shock_cap) hf_check(hf_dev_download(ctx, ele_type, HF_SENSOR, sensor.get_ptr_cpu(), sensor.size())); }
void eles::correct_gradient() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_CORRECT_GRADIENT)); }
void eles::evaluate_viscFlux() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EVALUATE_VISCFLUX)); }
void eles::extrapolate_totalFlux() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_EXTRAPOLATE_TOTALFLUX)); }
void eles::calculate_divergence() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_CALCULATE_DIVERGENCE)); }
void eles::calculate_corrected_divergence() { if (n_eles) hf_check(hf_dev_eles_op(ctx, ele_type, HF_CALCULATE_CORRECTED_DIVERGENCE)); }

void eles::AdvanceSolution(int in_step, int adv_type)
{
  // the device advances every element type in one call; issue it from the first non-empty type only so the
  // reference's "for every element type" loop (src/HiFiLES.cpp:209-210) advances exactly once
  (void)adv_type;
  if (n_eles == 0) return;
  hf_check(hf_dev_advance_solution(ctx, in_step | ((ele_type + 1) << 8)));
}

double eles::compute_res_upts(int in_norm_type, int in_field)
{
  // host-side restatement for callers that already copied div_tconf_upts back (reference src/eles.cpp:5045-5074)
  double sum = 0.;
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      double r = div_tconf_upts(0)(j, i, in_field) / detjac_upts(j, i) - src_upts(j, i, in_field);
      if (in_norm_type == 0) sum = max(sum, fabs(r));
      else if (in_norm_type == 1) sum += fabs(r);
      else if (in_norm_type == 2) sum += r * r;
    }
  return sum;
}

// ---- integral diagnostics ---------------------------------------------------------------------------------------------------
// interpolation solution points -> volume cubature points (reference eles::set_opp_volume_cubpts, src/eles.cpp:3667-3687)
void eles::set_opp_volume_cubpts()
{
  const int nc = loc_volume_cubpts.get_dim(1);
  hf_array<double> loc(n_dims);
  opp_volume_cubpts.setup(nc, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int j = 0; j < nc; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = loc_volume_cubpts(k, j);
      opp_volume_cubpts(j, i) = eval_nodal_basis(i, loc);
    }
}

// Jacobian determinant at the volume cubature points (reference eles::set_transforms_vol_cubpts, src/eles.cpp:4599-4632)
void eles::set_transforms_vol_cubpts()
{
  const int nc = loc_volume_cubpts.get_dim(1);
  hf_array<double> d_pos(n_dims, n_dims), loc(n_dims);
  vol_detjac_vol_cubpts.setup(nc, n_eles);
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < nc; j++)
    {
      for (int m = 0; m < n_dims; m++) loc(m) = loc_volume_cubpts(m, j);
      calc_d_pos(loc, i, d_pos);
      if (n_dims == 2)
        vol_detjac_vol_cubpts(j, i) = d_pos(0, 0) * d_pos(1, 1) - d_pos(0, 1) * d_pos(1, 0);
      else
        vol_detjac_vol_cubpts(j, i) = d_pos(0, 0) * (d_pos(1, 1) * d_pos(2, 2) - d_pos(1, 2) * d_pos(2, 1)) -
                                      d_pos(0, 1) * (d_pos(1, 0) * d_pos(2, 2) - d_pos(1, 2) * d_pos(2, 0)) +
                                      d_pos(0, 2) * (d_pos(1, 0) * d_pos(2, 1) - d_pos(1, 1) * d_pos(2, 0));
    }
}

// reference eles::CalcIntegralQuantities (src/eles.cpp:5485-5628): one device call, adds this type's share
void eles::CalcIntegralQuantities(int n_integral_quantities, hf_array<double> &integral_quantities)
{
  if (n_eles == 0) return;
  if (n_integral_quantities > HF_MAX_INTEGRAL_QUANTITIES) FatalError("too many integral quantities");
  int kinds[HF_MAX_INTEGRAL_QUANTITIES];
  for (int m = 0; m < n_integral_quantities; m++)
  {
    const string &q = run_input.integral_quantities(m);
    if (q == "kineticenergy") kinds[m] = 0;
    else if (q == "enstropy") kinds[m] = 1;
    else if (q == "pressuredilatation") kinds[m] = 2;
    else if (q == "straincolonproduct") kinds[m] = 3;
    else if (q == "devstraincolonproduct") kinds[m] = 4;
    else FatalError("integral diagnostic quantity not recognized");
  }
  hf_check(hf_dev_integral_quantities(ctx, ele_type, n_integral_quantities, kinds, integral_quantities.get_ptr_cpu()));
}

// reference eles::compute_error + get_pointwise_error (src/eles.cpp:5076-5276), on the host after a device -> host copy of
// the solution (and, for a viscous run, of the gradient of the last residual evaluation): it runs once, at the end of a
// test-case run.  Built: test_case 1 (isentropic vortex), 2 / 3 / 4 (advection-diffusion: plane sine wave, product of sines,
// Gaussian pulse); the Couette case fails loudly.
hf_array<double> eles::compute_error(int in_norm_type, double &time)
{
  const int tc = run_input.test_case;
  if (tc < 1 || tc > 4) FatalError("Test case not recognized in compute error, exiting");
  if (in_norm_type != 1 && in_norm_type != 2) FatalError("Error norm not supported!");
  hf_array<double> disu_cubpt(n_fields), grad_disu_cubpt(n_fields, n_dims), pos(n_dims), temp_loc(n_dims);
  hf_array<double> error_sol(n_fields), error_grad_sol(n_fields, n_dims), grad_rho(n_dims);
  hf_array<double> error_sum(2, n_fields);
  for (int m = 0; m < n_fields; m++) error_sum(0, m) = error_sum(1, m) = 0.;
  const int n_cubpts_per_ele = loc_volume_cubpts.get_dim(1);
  // the sine-wave cases compare the gradient too; without diffusion the exact solution is the undamped wave
  const double diff = viscous ? run_input.diff_coeff : 0.;
  cp_disu_upts_gpu_cpu();
  const bool with_grad = viscous && (tc == 2 || tc == 3); // the only cases whose gradient error is not identically zero
  if (with_grad) cp_grad_disu_upts_gpu_cpu();
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < n_cubpts_per_ele; j++)
    {
      const double detjac = vol_detjac_vol_cubpts(j, i);
      for (int k = 0; k < n_dims; k++) temp_loc(k) = loc_volume_cubpts(k, j);
      calc_pos(temp_loc, i, pos);
      for (int m = 0; m < n_fields; m++)
      {
        disu_cubpt(m) = 0.;
        for (int k = 0; k < n_upts_per_ele; k++) disu_cubpt(m) += opp_volume_cubpts(j, k) * disu_upts(0)(k, i, m);
        for (int n = 0; n < n_dims; n++)
        {
          grad_disu_cubpt(m, n) = 0.;
          if (with_grad)
            for (int k = 0; k < n_upts_per_ele; k++) grad_disu_cubpt(m, n) += opp_volume_cubpts(j, k) * grad_disu_upts(k, i, m, n);
          error_grad_sol(m, n) = 0.; // stays zero for the vortex and the pulse (src/eles.cpp:5149-5165, 5209-5214)
        }
      }
      if (tc == 1)
      {
        double rho, vx, vy, vz, p;
        eval_isentropic_vortex(pos, time, rho, vx, vy, vz, p, n_dims);
        error_sol(0) = disu_cubpt(0) - rho;
        error_sol(1) = disu_cubpt(1) - rho * vx;
        error_sol(2) = disu_cubpt(2) - rho * vy;
        if (n_dims == 2)
          error_sol(3) = disu_cubpt(3) - (p / (run_input.gamma - 1) + 0.5 * rho * (vx * vx + vy * vy));
        else
        {
          error_sol(3) = disu_cubpt(3) - rho * vz;
          error_sol(4) = disu_cubpt(4) - (p / (run_input.gamma - 1) + 0.5 * rho * (vx * vx + vy * vy + vz * vz));
        }
      }
      else
      {
        double rho;
        if (tc == 2) eval_sine_wave_single(pos, run_input.wave_speed, diff, time, rho, grad_rho, n_dims);
        else if (tc == 3) eval_sine_wave_group(pos, run_input.wave_speed, diff, time, rho, grad_rho, n_dims);
        else eval_sphere_wave(pos, run_input.wave_speed, time, rho, n_dims);
        error_sol(0) = disu_cubpt(0) - rho;
        if (tc != 4)
          for (int n = 0; n < n_dims; n++) error_grad_sol(0, n) = grad_disu_cubpt(0, n) - grad_rho(n);
      }
      for (int m = 0; m < n_fields; m++)
      {
        double e0 = 0., e1 = 0.;
        if (in_norm_type == 1)
        {
          e0 += fabs(error_sol(m));
          for (int n = 0; n < n_dims; n++) e1 += fabs(error_grad_sol(m, n));
        }
        else
        {
          e0 += error_sol(m) * error_sol(m);
          for (int n = 0; n < n_dims; n++) e1 += error_grad_sol(m, n) * error_grad_sol(m, n);
        }
        error_sum(0, m) += e0 * weight_volume_cubpts(j) * detjac;
        error_sum(1, m) += e1 * weight_volume_cubpts(j) * detjac;
      }
    }
  return error_sum;
}

// ---- solution patch ------------------------------------------------------------------------------------------------------------
// reference eles::set_patch (src/eles.cpp:535-652): patch_type 0 superposes an isentropic vortex ring (solid-body core of
// radius ra, decaying to zero at rb, peak Mach number Mv) on the solution inside r <= rb around (xc, yc); patch_type 1 sets
// the initial-condition state for x >= patch_x.  The expressions keep the reference's order of operations.
void eles::set_patch()
{
  const double gamma = run_input.gamma;
  const double temp_R = viscous ? run_input.R_ref : run_input.R_gas;
  hf_array<double> pos(n_dims);
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      for (int k = 0; k < n_dims; k++) pos(k) = pos_upts(j, i, k);
      double rho, vx, vy, vz = 0., p;
      if (run_input.patch_type == 0)
      {
        const double Mv = run_input.Mv, ra = run_input.ra, rb = run_input.rb, xc = run_input.xc, yc = run_input.yc;
        const double r = sqrt(pow(pos(0) - xc, 2) + pow(pos(1) - yc, 2));
        if (!(r <= rb)) continue;
        rho = disu_upts(0)(j, i, 0);
        vx = disu_upts(0)(j, i, 1) / disu_upts(0)(j, i, 0);
        vy = disu_upts(0)(j, i, 2) / disu_upts(0)(j, i, 0);
        if (n_dims == 2)
          p = (disu_upts(0)(j, i, 3) - 0.5 * rho * (vx * vx + vy * vy)) * (gamma - 1.0);
        else
        {
          vz = disu_upts(0)(j, i, 3) / disu_upts(0)(j, i, 0);
          p = (disu_upts(0)(j, i, 4) - 0.5 * rho * (vx * vx + vy * vy + vz * vz)) * (gamma - 1.0);
        }
        const double vm = Mv * sqrt(gamma * p / rho); // peak swirl velocity
        const double T0 = p / (rho * temp_R), cT = (gamma - 1) / (temp_R * gamma);
        double temper;
        if (r <= ra)
        {
          // solid-body core: (unit tangent) * vm * r / ra, associated left to right as the reference writes it
          vx -= (pos(1) - yc) / r * vm * r / ra;
          vy += (pos(0) - xc) / r * vm * r / ra;
          const double core = pow(vm, 2) / pow(ra, 2) * 0.5 * (pow(ra, 2) - pow(r, 2));
          const double ring = pow(vm, 2) * pow(ra, 2) / pow(pow(ra, 2) - pow(rb, 2), 2) *
                              (0.5 * (pow(rb, 2) - pow(ra, 2)) - 0.5 * pow(rb, 4) * (1 / pow(rb, 2) - 1 / pow(ra, 2)) - 2 * pow(rb, 2) * (log(rb / ra)));
          temper = T0 - cT * (core + ring);
        }
        else
        {
          vx -= (pos(1) - yc) / r * vm * ra / (pow(ra, 2) - pow(rb, 2)) * (r - pow(rb, 2) / r);
          vy += (pos(0) - xc) / r * vm * ra / (pow(ra, 2) - pow(rb, 2)) * (r - pow(rb, 2) / r);
          temper = T0 - cT * pow(vm, 2) * pow(ra, 2) / pow(pow(ra, 2) - pow(rb, 2), 2) *
                            (0.5 * (pow(rb, 2) - pow(r, 2)) - 0.5 * pow(rb, 4) * (1 / (pow(rb, 2)) - 1 / (pow(r, 2))) - 2 * pow(rb, 2) * (log(rb / r)));
        }
        const double rho_temp = rho;
        rho = rho * pow(temper / (p / (rho * temp_R)), 1 / (gamma - 1));
        p = p * pow(temper / (p / (rho_temp * temp_R)), gamma / (gamma - 1));
      }
      else if (run_input.patch_type == 1)
      {
        if (!(pos(0) >= run_input.patch_x)) continue;
        rho = run_input.rho_c_ic;
        vx = run_input.u_c_ic;
        vy = run_input.v_c_ic;
        vz = run_input.w_c_ic;
        p = run_input.p_c_ic;
      }
      else
      {
        FatalError("ERROR: Invalid form of patch ... ");
        return;
      }
      disu_upts(0)(j, i, 0) = rho;
      disu_upts(0)(j, i, 1) = rho * vx;
      disu_upts(0)(j, i, 2) = rho * vy;
      if (n_dims == 2)
        disu_upts(0)(j, i, 3) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy)));
      else
      {
        disu_upts(0)(j, i, 3) = rho * vz;
        disu_upts(0)(j, i, 4) = (p / (gamma - 1.0)) + (0.5 * rho * ((vx * vx) + (vy * vy) + (vz * vz)));
      }
    }
}

// ---- time averages ---------------------------------------------------------------------------------------------------------------
void eles::CalcTimeAverageQuantities(double &time)
{
  if (n_eles == 0) return;
  const int n = run_input.n_average_fields;
  if (n > HF_MAX_INTEGRAL_QUANTITIES) FatalError("too many average fields");
  int kinds[HF_MAX_INTEGRAL_QUANTITIES];
  for (int i = 0; i < n; i++)
  {
    const string &f = run_input.average_fields(i);
    if (f == "rho_average") kinds[i] = 0;
    else if (f == "u_average") kinds[i] = 1;
    else if (f == "v_average") kinds[i] = 2;
    else if (f == "w_average") kinds[i] = 3;
    else if (f == "e_average") kinds[i] = 4;
    else FatalError("average field not recognized: " + f);
  }
  hf_check(hf_dev_time_average(ctx, ele_type, n, kinds, time, run_input.spinup_time));
}

void eles::cp_disu_average_upts_gpu_cpu()
{
  if (!n_eles || !run_input.n_average_fields) return;
  if (disu_average_upts.size() == 0) disu_average_upts.setup(n_upts_per_ele, n_eles, run_input.n_average_fields);
  hf_check(hf_dev_download(ctx, ele_type, HF_DISU_AVERAGE_UPTS, disu_average_upts.get_ptr_cpu(), disu_average_upts.size()));
}
