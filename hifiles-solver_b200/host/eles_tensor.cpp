// Tensor-product element types: hexahedra and quadrilaterals.
// Point orderings, face numbering and the reversal of the in-face coordinate on some faces are the reference's
// (hexas: src/eles_hexas.cpp:198-282, 525-580, 1132-1193, 1444-1537; quads: src/eles_quads.cpp:187-262, 389-426,
// 962-1120, 1192-1275) -- they define which entry of every operator is an exact 0 or 1, which the device kernels
// rely on when they apply the operators in sum-factorised form.
#include "hifiles.h"

using namespace std;

static void set_loc_1d_spts(hf_array<double> &loc_1d_spts, int n)
{
  for (int i = 0; i < n; i++) loc_1d_spts(i) = -1.0 + ((2.0 * i) / (1.0 * (n - 1)));
}

// =============================================================================================================
// hexahedra
// =============================================================================================================
void eles_hexas::setup_ele_type_specific()
{
  ele_type = HEX;
  n_dims = 3;
  if (run_input.equation == 0) n_fields = 5;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  n_inters_per_ele = 6;
  n_upts_per_ele = (order + 1) * (order + 1) * (order + 1);
  upts_type = run_input.upts_type_hexa;
  hf_array<double> w;
  cubature_1d(upts_type, order, loc_1d_upts, w);
  set_loc_upts();
  n_fpts_per_inter.setup(6);
  for (int i = 0; i < 6; i++) n_fpts_per_inter(i) = (order + 1) * (order + 1);
  n_fpts_per_ele = n_inters_per_ele * (order + 1) * (order + 1);
  set_tloc_fpts();
  set_tnorm_fpts();
  set_opp_0(run_input.sparse_hexa);
  set_opp_1(run_input.sparse_hexa);
  set_opp_2(run_input.sparse_hexa);
  set_opp_3(run_input.sparse_hexa);
  if (viscous)
  {
    set_opp_4(run_input.sparse_hexa);
    set_opp_5(run_input.sparse_hexa);
    set_opp_6(run_input.sparse_hexa);
  }
}

void eles_hexas::set_loc_upts()
{
  int n = order + 1;
  loc_upts.setup(n_dims, n_upts_per_ele);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++)
      {
        int upt = k + n * j + n * n * i;
        loc_upts(0, upt) = loc_1d_upts(k);
        loc_upts(1, upt) = loc_1d_upts(j);
        loc_upts(2, upt) = loc_1d_upts(i);
      }
}

void eles_hexas::set_tloc_fpts()
{
  int n = order + 1;
  tloc_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < n_inters_per_ele; i++)
    for (int j = 0; j < n; j++)
      for (int k = 0; k < n; k++)
      {
        int fpt = k + n * j + n * n * i;
        double a = loc_1d_upts(k), ar = loc_1d_upts(order - k), b = loc_1d_upts(j);
        switch (i)
        {
        case 0: tloc_fpts(0, fpt) = ar;   tloc_fpts(1, fpt) = b;    tloc_fpts(2, fpt) = -1.0; break;
        case 1: tloc_fpts(0, fpt) = a;    tloc_fpts(1, fpt) = -1.0; tloc_fpts(2, fpt) = b;    break;
        case 2: tloc_fpts(0, fpt) = 1.0;  tloc_fpts(1, fpt) = a;    tloc_fpts(2, fpt) = b;    break;
        case 3: tloc_fpts(0, fpt) = ar;   tloc_fpts(1, fpt) = 1.0;  tloc_fpts(2, fpt) = b;    break;
        case 4: tloc_fpts(0, fpt) = -1.0; tloc_fpts(1, fpt) = ar;   tloc_fpts(2, fpt) = b;    break;
        case 5: tloc_fpts(0, fpt) = a;    tloc_fpts(1, fpt) = b;    tloc_fpts(2, fpt) = 1.0;  break;
        }
      }
}

void eles_hexas::set_tnorm_fpts()
{
  int n2 = (order + 1) * (order + 1);
  static const double nrm[6][3] = {{0, 0, -1}, {0, -1, 0}, {1, 0, 0}, {0, 1, 0}, {-1, 0, 0}, {0, 0, 1}};
  tnorm_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < 6; i++)
    for (int q = 0; q < n2; q++)
      for (int d = 0; d < 3; d++) tnorm_fpts(d, q + n2 * i) = nrm[i][d];
}

double eles_hexas::eval_nodal_basis(int in_index, hf_array<double> &in_loc)
{
  int n = order + 1;
  int i = in_index / (n * n);
  int j = (in_index - n * n * i) / n;
  int k = in_index - n * j - n * n * i;
  return eval_lagrange(in_loc(0), k, loc_1d_upts) * eval_lagrange(in_loc(1), j, loc_1d_upts) * eval_lagrange(in_loc(2), i, loc_1d_upts);
}

double eles_hexas::eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc)
{
  int n = order + 1;
  int i = in_index / (n * n);
  int j = (in_index - n * n * i) / n;
  int k = in_index - n * j - n * n * i;
  if (in_cpnt == 0)
    return eval_d_lagrange(in_loc(0), k, loc_1d_upts) * eval_lagrange(in_loc(1), j, loc_1d_upts) * eval_lagrange(in_loc(2), i, loc_1d_upts);
  if (in_cpnt == 1)
    return eval_lagrange(in_loc(0), k, loc_1d_upts) * eval_d_lagrange(in_loc(1), j, loc_1d_upts) * eval_lagrange(in_loc(2), i, loc_1d_upts);
  return eval_lagrange(in_loc(0), k, loc_1d_upts) * eval_lagrange(in_loc(1), j, loc_1d_upts) * eval_d_lagrange(in_loc(2), i, loc_1d_upts);
}

// 20-node serendipity hexahedron (reference src/eles_hexas.cpp:1215-1257, 1292-1356).  Node p sits at (px,py,pz): the eight
// corners counter-clockwise bottom then top, then the mid-edge nodes of the bottom ring, the vertical edges and the top ring.
//   corner:    N = px py pz / 8 (x+px)(y+py)(z+pz) (px x + py y + pz z - 2)
//   mid-edge:  N = -pu pv / 4 (u+pu)(v+pv)(w^2-1),  w the edge direction, u < v the other two
// Every product and sum below is associated as the reference writes it (overall signs are exact in floating point, so only
// the order of the magnitudes matters): the metrics of curved hexes are bit-identical to the reference's.
static const int hex20_pos[20][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1}, {1, 0, -1}, {0, 1, -1}, {-1, 0, -1},
    {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0},
    {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1}};

static inline void hex20_edge_axes(const int *p, int &u, int &v, int &w)
{
  w = p[0] == 0 ? 0 : p[1] == 0 ? 1 : 2;
  u = w == 0 ? 1 : 0;
  v = w == 2 ? 1 : 2;
}

static double hex20_basis(int m, const double *x)
{
  const int *p = hex20_pos[m];
  if (m < 8)
  {
    double f4 = ((p[0] * x[0] - 2.) + p[1] * x[1]) + p[2] * x[2];
    return (p[0] * p[1] * p[2]) * (0.125 * (x[0] + p[0])) * (x[1] + p[1]) * (x[2] + p[2]) * f4;
  }
  int u, v, w;
  hex20_edge_axes(p, u, v, w);
  return (-p[u] * p[v]) * (0.25 * (x[u] + p[u])) * (x[v] + p[v]) * (x[w] * x[w] - 1.);
}

static void hex20_d_basis(int m, const double *x, double *d)
{
  const int *p = hex20_pos[m];
  if (m < 8)
  {
    double sgn = p[0] * p[1] * p[2];
    for (int c = 0; c < 3; c++)
    {
      int a = c == 0 ? 1 : 0, b = c == 2 ? 1 : 2; // the two other directions
      double t = p[a] * x[a] + p[b] * x[b];
      // the reference adds the constant before the doubled term in three entries (src/eles_hexas.cpp:1301, 1322, 1323)
      bool const_first = (c == 0 && m == 7) || (c == 1 && (m == 6 || m == 7));
      double sum = const_first ? (t - 1.) + 2. * p[c] * x[c] : (t + 2. * p[c] * x[c]) - 1.;
      d[c] = sgn * (0.125 * (x[b] + p[b])) * (x[a] + p[a]) * sum;
    }
    return;
  }
  int u, v, w;
  hex20_edge_axes(p, u, v, w);
  double c4 = -p[u] * p[v];
  int first = w == 2 ? 0 : 2, second = w == 0 ? 1 : w == 1 ? 0 : 1; // order of the two linear factors behind (1/2) w
  d[w] = c4 * 0.5 * x[w] * (x[first] + p[first]) * (x[second] + p[second]);
  d[u] = c4 * (0.25 * (x[v] + p[v])) * (x[w] * x[w] - 1.);
  d[v] = c4 * (0.25 * (x[u] + p[u])) * (x[w] * x[w] - 1.);
}

double eles_hexas::eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts)
{
  if (in_n_spts == 20)
  {
    double x[3] = {in_loc(0), in_loc(1), in_loc(2)};
    return hex20_basis(in_index, x);
  }
  if (!is_perfect_cube(in_n_spts))
    FatalError("Shape basis not implemented yet, exiting");
  int n = (int)round(pow(in_n_spts, 1. / 3.));
  hf_array<double> l(n);
  set_loc_1d_spts(l, n);
  int i = in_index / (n * n);
  int j = (in_index - n * n * i) / n;
  int k = in_index - n * j - n * n * i;
  return eval_lagrange(in_loc(0), k, l) * eval_lagrange(in_loc(1), j, l) * eval_lagrange(in_loc(2), i, l);
}

void eles_hexas::eval_d_nodal_s_basis(hf_array<double> &d, hf_array<double> &in_loc, int in_n_spts)
{
  if (in_n_spts == 20)
  {
    double x[3] = {in_loc(0), in_loc(1), in_loc(2)}, g[3];
    for (int m = 0; m < 20; m++)
    {
      hex20_d_basis(m, x, g);
      for (int c = 0; c < 3; c++) d(m, c) = g[c];
    }
    return;
  }
  if (!is_perfect_cube(in_n_spts))
    FatalError("Shape basis not implemented yet, exiting");
  int n = (int)round(pow(in_n_spts, 1. / 3.));
  hf_array<double> l(n);
  set_loc_1d_spts(l, n);
  for (int m = 0; m < in_n_spts; m++)
  {
    int i = m / (n * n);
    int j = (m - n * n * i) / n;
    int k = m - n * j - n * n * i;
    d(m, 0) = eval_d_lagrange(in_loc(0), k, l) * eval_lagrange(in_loc(1), j, l) * eval_lagrange(in_loc(2), i, l);
    d(m, 1) = eval_lagrange(in_loc(0), k, l) * eval_d_lagrange(in_loc(1), j, l) * eval_lagrange(in_loc(2), i, l);
    d(m, 2) = eval_lagrange(in_loc(0), k, l) * eval_lagrange(in_loc(1), j, l) * eval_d_lagrange(in_loc(2), i, l);
  }
}

void eles_hexas::fill_opp_3(hf_array<double> &opp_3)
{
  hf_array<double> loc(n_dims);
  for (int i = 0; i < n_fpts_per_ele; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = loc_upts(k, j);
      opp_3(j, i) = eval_div_vcjh_basis(i, loc);
    }
}

double eles_hexas::eval_div_vcjh_basis(int in_index, hf_array<double> &loc)
{
  int scheme = run_input.vcjh_scheme_hexa;
  double eta = 0.;
  if (scheme == 0) eta = run_input.eta_hexa;
  else if (scheme < 5) eta = compute_eta(scheme, order);
  else FatalError("OFR / OESFR correction functions are not available in this build");
  int nf = n_fpts_per_inter(0);
  int i = in_index / nf;
  int j = (in_index - nf * i) / (order + 1);
  int k = in_index - nf * i - (order + 1) * j;
  switch (i)
  {
  case 0: return -eval_lagrange(loc(0), order - k, loc_1d_upts) * eval_lagrange(loc(1), j, loc_1d_upts) * eval_d_vcjh_1d(loc(2), 0, order, eta);
  case 1: return -eval_lagrange(loc(0), k, loc_1d_upts) * eval_lagrange(loc(2), j, loc_1d_upts) * eval_d_vcjh_1d(loc(1), 0, order, eta);
  case 2: return eval_lagrange(loc(1), k, loc_1d_upts) * eval_lagrange(loc(2), j, loc_1d_upts) * eval_d_vcjh_1d(loc(0), 1, order, eta);
  case 3: return eval_lagrange(loc(0), order - k, loc_1d_upts) * eval_lagrange(loc(2), j, loc_1d_upts) * eval_d_vcjh_1d(loc(1), 1, order, eta);
  case 4: return -eval_lagrange(loc(1), order - k, loc_1d_upts) * eval_lagrange(loc(2), j, loc_1d_upts) * eval_d_vcjh_1d(loc(0), 0, order, eta);
  default: return eval_lagrange(loc(0), k, loc_1d_upts) * eval_lagrange(loc(1), j, loc_1d_upts) * eval_d_vcjh_1d(loc(2), 1, order, eta);
  }
}

// =============================================================================================================
// quadrilaterals
// =============================================================================================================
void eles_quads::setup_ele_type_specific()
{
  ele_type = QUAD;
  n_dims = 2;
  if (run_input.equation == 0) n_fields = 4;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  n_inters_per_ele = 4;
  n_upts_per_ele = (order + 1) * (order + 1);
  upts_type = run_input.upts_type_quad;
  hf_array<double> w;
  cubature_1d(upts_type, order, loc_1d_upts, w);
  set_loc_upts();
  n_fpts_per_inter.setup(4);
  for (int i = 0; i < 4; i++) n_fpts_per_inter(i) = order + 1;
  n_fpts_per_ele = n_inters_per_ele * (order + 1);
  set_tloc_fpts();
  set_tnorm_fpts();
  set_opp_0(run_input.sparse_quad);
  set_opp_1(run_input.sparse_quad);
  set_opp_2(run_input.sparse_quad);
  set_opp_3(run_input.sparse_quad);
  if (viscous)
  {
    set_opp_4(run_input.sparse_quad);
    set_opp_5(run_input.sparse_quad);
    set_opp_6(run_input.sparse_quad);
  }
}

void eles_quads::set_loc_upts()
{
  int n = order + 1;
  loc_upts.setup(n_dims, n_upts_per_ele);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      int upt = j + n * i;
      loc_upts(0, upt) = loc_1d_upts(j);
      loc_upts(1, upt) = loc_1d_upts(i);
    }
}

void eles_quads::set_tloc_fpts()
{
  int n = order + 1;
  tloc_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < n; j++)
    {
      int fpt = j + n * i;
      switch (i)
      {
      case 0: tloc_fpts(0, fpt) = loc_1d_upts(j);         tloc_fpts(1, fpt) = -1.0; break;
      case 1: tloc_fpts(0, fpt) = 1.0;                    tloc_fpts(1, fpt) = loc_1d_upts(j); break;
      case 2: tloc_fpts(0, fpt) = loc_1d_upts(order - j); tloc_fpts(1, fpt) = 1.0; break;
      case 3: tloc_fpts(0, fpt) = -1.0;                   tloc_fpts(1, fpt) = loc_1d_upts(order - j); break;
      }
    }
}

void eles_quads::set_tnorm_fpts()
{
  int n = order + 1;
  static const double nrm[4][2] = {{0, -1}, {1, 0}, {0, 1}, {-1, 0}};
  tnorm_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < n; j++)
      for (int d = 0; d < 2; d++) tnorm_fpts(d, j + n * i) = nrm[i][d];
}

double eles_quads::eval_nodal_basis(int in_index, hf_array<double> &in_loc)
{
  int i = in_index / (order + 1);
  int j = in_index - (order + 1) * i;
  return eval_lagrange(in_loc(0), j, loc_1d_upts) * eval_lagrange(in_loc(1), i, loc_1d_upts);
}

double eles_quads::eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc)
{
  int i = in_index / (order + 1);
  int j = in_index - (order + 1) * i;
  if (in_cpnt == 0)
    return eval_d_lagrange(in_loc(0), j, loc_1d_upts) * eval_lagrange(in_loc(1), i, loc_1d_upts);
  return eval_lagrange(in_loc(0), j, loc_1d_upts) * eval_d_lagrange(in_loc(1), i, loc_1d_upts);
}

double eles_quads::eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts)
{
  double x = in_loc(0), y = in_loc(1);
  if (is_perfect_square(in_n_spts))
  {
    int n = (int)round(sqrt(1.0 * in_n_spts));
    hf_array<double> l(n);
    set_loc_1d_spts(l, n);
    int j = in_index / n;
    int i = in_index - n * j;
    return eval_lagrange(x, i, l) * eval_lagrange(y, j, l);
  }
  if (in_n_spts == 8)
  {
    switch (in_index)
    {
    case 0: return -0.25 * (1. - x) * (1. - y) * (1. + x + y);
    case 1: return -0.25 * (1. + x) * (1. - y) * (1. - x + y);
    case 2: return -0.25 * (1. + x) * (1. + y) * (1. - x - y);
    case 3: return -0.25 * (1. - x) * (1. + y) * (1. + x - y);
    case 4: return 0.5 * (1. - x) * (1. + x) * (1. - y);
    case 5: return 0.5 * (1. + x) * (1. + y) * (1. - y);
    case 6: return 0.5 * (1. - x) * (1. + x) * (1. + y);
    default: return 0.5 * (1. - x) * (1. + y) * (1. - y);
    }
  }
  FatalError("Shape basis not implemented yet, exiting");
  return 0.;
}

void eles_quads::eval_d_nodal_s_basis(hf_array<double> &d, hf_array<double> &in_loc, int in_n_spts)
{
  double x = in_loc(0), y = in_loc(1);
  if (is_perfect_square(in_n_spts))
  {
    int n = (int)round(sqrt(1.0 * in_n_spts));
    hf_array<double> l(n);
    set_loc_1d_spts(l, n);
    for (int k = 0; k < in_n_spts; k++)
    {
      int i = k / n;
      int j = k - n * i;
      d(k, 0) = eval_d_lagrange(x, j, l) * eval_lagrange(y, i, l);
      d(k, 1) = eval_lagrange(x, j, l) * eval_d_lagrange(y, i, l);
    }
  }
  else if (in_n_spts == 8)
  {
    d(0, 0) = -0.25 * (-1. + y) * (2. * x + y);
    d(1, 0) = 0.25 * (-1. + y) * (y - 2. * x);
    d(2, 0) = 0.25 * (1. + y) * (2. * x + y);
    d(3, 0) = -0.25 * (1. + y) * (y - 2. * x);
    d(4, 0) = x * (-1. + y);
    d(5, 0) = -0.5 * (1. + y) * (-1. + y);
    d(6, 0) = -x * (1. + y);
    d(7, 0) = 0.5 * (1. + y) * (-1. + y);
    d(0, 1) = -0.25 * (-1. + x) * (x + 2. * y);
    d(1, 1) = 0.25 * (1. + x) * (2. * y - x);
    d(2, 1) = 0.25 * (1. + x) * (x + 2. * y);
    d(3, 1) = -0.25 * (-1. + x) * (2. * y - x);
    d(4, 1) = 0.5 * (1. + x) * (-1. + x);
    d(5, 1) = -y * (1. + x);
    d(6, 1) = -0.5 * (1. + x) * (-1. + x);
    d(7, 1) = y * (-1. + x);
  }
  else
    FatalError("Shape basis not implemented yet, exiting");
}

void eles_quads::fill_opp_3(hf_array<double> &opp_3)
{
  hf_array<double> loc(n_dims);
  for (int i = 0; i < n_fpts_per_ele; i++)
    for (int j = 0; j < n_upts_per_ele; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = loc_upts(k, j);
      opp_3(j, i) = eval_div_vcjh_basis(i, loc);
    }
}

double eles_quads::eval_div_vcjh_basis(int in_index, hf_array<double> &loc)
{
  int scheme = run_input.vcjh_scheme_quad;
  double eta = 0.;
  if (scheme == 0) eta = run_input.eta_quad;
  else if (scheme < 5) eta = compute_eta(scheme, order);
  else FatalError("OFR / OESFR correction functions are not available in this build");
  int nf = n_fpts_per_inter(0);
  int i = in_index / nf;
  int j = in_index - nf * i;
  switch (i)
  {
  case 0: return -eval_lagrange(loc(0), j, loc_1d_upts) * eval_d_vcjh_1d(loc(1), 0, order, eta);
  case 1: return eval_lagrange(loc(1), j, loc_1d_upts) * eval_d_vcjh_1d(loc(0), 1, order, eta);
  case 2: return eval_lagrange(loc(0), order - j, loc_1d_upts) * eval_d_vcjh_1d(loc(1), 1, order, eta);
  default: return -eval_lagrange(loc(1), order - j, loc_1d_upts) * eval_d_vcjh_1d(loc(0), 0, order, eta);
  }
}
