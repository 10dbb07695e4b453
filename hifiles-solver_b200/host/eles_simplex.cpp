// Element types with a collapsed-coordinate (Dubiner) modal basis: triangles, tetrahedra and prisms (triangle x line).
// Point sets, face numbering, in-face orderings and the construction of the correction operator follow the
// reference, because they decide which flux point meets which across an interface and what every operator entry is:
//   triangles    reference src/eles_tris.cpp:45-131, 178-225, 402-443, 703-784, 982-997
//   tetrahedra   reference src/eles_tets.cpp:45-132, 224-284, 540-574, 705-716, 977-1303, 1305-1478, 1598-1634
//   prisms       reference src/eles_pris.cpp:47-145, 210-306, 551-607, 973-1236, 1323-1402, 1504-1557
//   VCJH filter / DG lifting on a triangle   reference src/funcs.cpp:630-667, 717-890, 962-1048
// The operators of these types are dense; on the device they go through the generic small-operator kernels.
#include "hifiles.h"

using namespace std;

static int factorial_i(int n)
{
  int f = 1;
  for (int i = 2; i <= n; i++) f *= i;
  return f;
}

// nodal basis value from modal values: l_index = sum_i invV(i, index) * P_i   (Hesthaven & Warburton eq. 3.3)
static double modal_to_nodal(hf_array<double> &inv_vdm, int index, hf_array<double> &modal, int n)
{
  double v = 0.;
  for (int i = 0; i < n; i++) v += inv_vdm(i, index) * modal(i);
  return v;
}

// ---- VCJH correction on a triangle ----------------------------------------------------------------------------------
// divergence of the DG correction function of flux point (edge, edge_fpt) at in_loc: g.n on the edge is expanded in
// Legendre polynomials, its moments against the Dubiner basis are integrated with an 11-point Gauss rule
static double eval_div_dg_tri(hf_array<double> &in_loc, int in_edge, int in_edge_fpt, int in_order, hf_array<double> &in_loc_fpts_1d,
                              hf_array<double> &cub_r, hf_array<double> &cub_w)
{
  const int n_upts_tri = (in_order + 1) * (in_order + 2) / 2, n1 = in_order + 1;
  double edge_length = 2.;
  if (in_edge == 1) edge_length = 2. * sqrt(2.);
  hf_array<double> V(n1, n1), gdotn(n1, 1);
  for (int i = 0; i < n1; i++)
  {
    gdotn(i, 0) = i == in_edge_fpt ? 1. : 0.;
    const double t = (1. + in_loc_fpts_1d(i)) / 2. * edge_length;
    for (int j = 0; j < n1; j++) V(i, j) = eval_jacobi(t, 0, 0, j);
  }
  hf_array<double> Vi = inv_array(V);
  hf_array<double> coeff_gdotn = mult_arrays(Vi, gdotn);
  hf_array<double> coeff_divg(n_upts_tri, 1);
  const int ncub = cub_r.get_dim(0);
  for (int i = 0; i < n_upts_tri; i++)
  {
    double integral = 0.;
    for (int j = 0; j < ncub; j++)
    {
      double r, s, t;
      if (in_edge == 0) { t = (cub_r(j) + 1.) / 2. * edge_length; r = -1 + t; s = -1; }
      else if (in_edge == 1) { t = (cub_r(j) + 1.) / 2. * edge_length; r = 1 - t / edge_length * 2; s = -1 + t / edge_length * 2; }
      else { t = (cub_r(j) + 1.) / 2. * edge_length; r = -1; s = 1 - t; }
      double g = 0.;
      for (int k = 0; k < n1; k++) g += coeff_gdotn(k, 0) * eval_jacobi(t, 0, 0, k);
      integral += cub_w(j) * eval_dubiner_basis_2d(r, s, i, in_order) * g;
    }
    coeff_divg(i, 0) = integral * (edge_length) / 2;
  }
  double div = 0.;
  for (int i = 0; i < n_upts_tri; i++) div += coeff_divg(i, 0) * eval_dubiner_basis_2d(in_loc(0), in_loc(1), i, in_order);
  return div;
}

// c of the one-parameter VCJH family from the scheme switch (1 = DG, 2 = SD-like, 3 = Huynh-like, 4 = c+)
static double vcjh_c_simplex(int scheme, int order, double c_user, int dims)
{
  const double ap = 1. / pow(2.0, order) * factorial_i(2 * order) / (factorial_i(order) * factorial_i(order));
  const double c_sd_1d = (2 * order) / ((2 * order + 1) * (order + 1) * (factorial_i(order) * ap) * (factorial_i(order) * ap));
  const double c_hu_1d = (2 * (order + 1)) / ((2 * order + 1) * order * (factorial_i(order) * ap) * (factorial_i(order) * ap));
  double c_plus_1d = 0., c_plus = 0.;
  if (scheme > 1)
  {
    if (order == 2) c_plus_1d = 0.206;
    else if (order == 3) c_plus_1d = 3.80e-3;
    else if (order == 4) c_plus_1d = 4.67e-5;
    else if (order == 5) c_plus_1d = 4.28e-7;
    else FatalError("C_plus scheme not implemented for this order");
    if (dims == 2)
    {
      if (order == 2) c_plus = 3.13e-2;
      else if (order == 3) c_plus = 4.67e-4;
      else if (order == 4) c_plus = 6.55e-6;
      else FatalError("C_plus scheme not implemented for this order");
    }
    else
    {
      if (order == 2) c_plus = 3.07e-2;
      else if (order == 3) c_plus = 5.44e-4;
      else if (order == 4) c_plus = 9.92e-6;
      else if (order == 5) c_plus = 1.10e-7;
      else FatalError("C_plus scheme not implemented for this order");
    }
  }
  if (scheme == 0) return c_user;
  if (scheme == 1) return 0.;
  if (scheme == 2) return (c_sd_1d / c_plus_1d) * c_plus;
  if (scheme == 3) return (c_hu_1d / c_plus_1d) * c_plus;
  if (scheme == 4) return c_plus;
  FatalError(dims == 2 ? "VCJH triangular scheme not recognized" : "VCJH tetrahedral scheme not recognized");
  return 0.;
}

// Filt = (I + M^-1 K)^-1 with M^-1 = V V^T and K = c * sum_k coeff_k (D^(k))^T D^(k) over the highest-order derivative
// operators D^(k) listed by `powers` (how often Dr, Ds, Dt are applied, in the reference's order)
static hf_array<double> vcjh_filter(hf_array<double> &vdm, vector<hf_array<double>> &D, const vector<vector<int>> &powers, const vector<double> &coeff, double c,
                                    int n)
{
  hf_array<double> K(n, n), I(n, n);
  for (int i = 0; i < n; i++) I(i, i) = 1.;
  for (size_t q = 0; q < powers.size(); q++)
  {
    hf_array<double> H(I);
    for (size_t d = 0; d < powers[q].size(); d++)
      for (int rep = 0; rep < powers[q][d]; rep++) H = mult_arrays(H, D[d]);
    hf_array<double> Ht = transpose_array(H);
    hf_array<double> HtH = mult_arrays(Ht, H);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
      {
        HtH(i, j) = c * coeff[q] * HtH(i, j);
        K(i, j) += HtH(i, j);
      }
  }
  hf_array<double> Vt = transpose_array(vdm);
  hf_array<double> Minv = mult_arrays(vdm, Vt);
  hf_array<double> T = mult_arrays(Minv, K);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) T(i, j) += I(i, j);
  return inv_array(T);
}

void get_opp_3_tri(hf_array<double> &opp_3, hf_array<double> &loc_upts_tri, hf_array<double> &loc_1d_fpts, hf_array<double> &vandermonde_tri,
                   hf_array<double> &inv_vandermonde_tri, int n_upts_per_tri, int order, double c_tri, int vcjh_scheme_tri)
{
  const int n = n_upts_per_tri;
  c_tri = vcjh_c_simplex(vcjh_scheme_tri, order, c_tri, 2);
  run_input.c_tri = c_tri;
  hf_array<double> tr(n, n), ts(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      tr(i, j) = eval_dr_dubiner_basis_2d(loc_upts_tri(0, i), loc_upts_tri(1, i), j, order);
      ts(i, j) = eval_ds_dubiner_basis_2d(loc_upts_tri(0, i), loc_upts_tri(1, i), j, order);
    }
  vector<hf_array<double>> D(2);
  D[0] = mult_arrays(ts, inv_vandermonde_tri); // applied first: Ds k times, then Dr (order - k) times
  D[1] = mult_arrays(tr, inv_vandermonde_tri);
  vector<vector<int>> powers;
  vector<double> coeff;
  for (int k = 0; k < order + 1; k++)
  {
    powers.push_back({k, order - k});
    coeff.push_back((1. / n) * (factorial_i(order) / (factorial_i(k) * factorial_i(order - k))));
  }
  hf_array<double> Filt = vcjh_filter(vandermonde_tri, D, powers, coeff, c_tri, n);

  hf_array<double> dg(n, 3 * (order + 1)), loc(2), cr, cw;
  cubature_1d(0, 10, cr, cw);
  for (int i = 0; i < 3 * (order + 1); i++)
    for (int j = 0; j < n; j++)
    {
      loc(0) = loc_upts_tri(0, j);
      loc(1) = loc_upts_tri(1, j);
      dg(j, i) = eval_div_dg_tri(loc, i / (order + 1), i % (order + 1), order, loc_1d_fpts, cr, cw);
    }
  opp_3 = mult_arrays(Filt, dg);
}

// =============================================================================================================
// triangles
// =============================================================================================================
void eles_tris::setup_ele_type_specific()
{
  ele_type = TRI;
  n_dims = 2;
  if (run_input.equation == 0) n_fields = 4;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  n_inters_per_ele = 3;
  n_upts_per_ele = (order + 2) * (order + 1) / 2;
  upts_type = run_input.upts_type_tri;

  hf_array<double> pts, w;
  cubature_tri(upts_type, order, pts, w);
  loc_upts.setup(n_dims, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++) { loc_upts(0, i) = pts(i, 0); loc_upts(1, i) = pts(i, 1); }
  vandermonde.setup(n_upts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int j = 0; j < n_upts_per_ele; j++) vandermonde(i, j) = eval_dubiner_basis_2d(loc_upts(0, i), loc_upts(1, i), j, order);
  inv_vandermonde = inv_array(vandermonde);

  n_fpts_per_inter.setup(3);
  for (int i = 0; i < 3; i++) n_fpts_per_inter(i) = order + 1;
  n_fpts_per_ele = n_inters_per_ele * (order + 1);

  // flux points: edge 0 along +r at s = -1, edge 1 the hypotenuse from (1,-1) to (-1,1), edge 2 down the r = -1 side
  cubature_1d(run_input.fpts_type_tri, order, loc_1d_fpts, w);
  tloc_fpts.setup(n_dims, n_fpts_per_ele);
  tnorm_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < order + 1; j++)
    {
      const int fpt = (order + 1) * i + j;
      if (i == 0) { tloc_fpts(0, fpt) = loc_1d_fpts(j); tloc_fpts(1, fpt) = -1.0; tnorm_fpts(0, fpt) = 0.; tnorm_fpts(1, fpt) = -1.; }
      else if (i == 1)
      {
        tloc_fpts(0, fpt) = loc_1d_fpts(order - j); tloc_fpts(1, fpt) = loc_1d_fpts(j);
        tnorm_fpts(0, fpt) = 1. / sqrt(2.); tnorm_fpts(1, fpt) = 1. / sqrt(2.);
      }
      else { tloc_fpts(0, fpt) = -1.0; tloc_fpts(1, fpt) = loc_1d_fpts(order - j); tnorm_fpts(0, fpt) = -1.; tnorm_fpts(1, fpt) = 0.; }
    }

  set_opp_0(run_input.sparse_tri);
  set_opp_1(run_input.sparse_tri);
  set_opp_2(run_input.sparse_tri);
  set_opp_3(run_input.sparse_tri);
  if (viscous)
  {
    set_opp_4(run_input.sparse_tri);
    set_opp_5(run_input.sparse_tri);
    set_opp_6(run_input.sparse_tri);
  }
}

double eles_tris::eval_nodal_basis(int in_index, hf_array<double> &in_loc)
{
  hf_array<double> modal(n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++) modal(i) = eval_dubiner_basis_2d(in_loc(0), in_loc(1), i, order);
  return modal_to_nodal(inv_vandermonde, in_index, modal, n_upts_per_ele);
}

double eles_tris::eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc)
{
  hf_array<double> modal(n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    modal(i) = in_cpnt == 0 ? eval_dr_dubiner_basis_2d(in_loc(0), in_loc(1), i, order) : eval_ds_dubiner_basis_2d(in_loc(0), in_loc(1), i, order);
  return modal_to_nodal(inv_vandermonde, in_index, modal, n_upts_per_ele);
}

void eles_tris::fill_opp_3(hf_array<double> &opp_3)
{
  get_opp_3_tri(opp_3, loc_upts, loc_1d_fpts, vandermonde, inv_vandermonde, n_upts_per_ele, order, run_input.c_tri, run_input.vcjh_scheme_tri);
}

// Shape functions of the 3- and 6-node triangle as products T_I(r) T_J(s) T_K(t), t = -1 - r - s, of 1-D polynomials
// (Hughes, The Finite Element Method, p. 166).  The reference builds them with a general polynomial algebra
// (src/funcs.cpp:1966-2458); for the two node counts its mesh readers accept the factors are the ones tabulated here
// (coefficients of descending powers, exactly representable), and the evaluation below keeps the reference's order:
// row value = sum_j c_j * pow(x, power_j) from the highest power down, value = product of the three rows, and a
// derivative is the sum of the layer differentiated in r (or s) and the layer differentiated in t times -1.
namespace
{
struct tri_shape
{
  int len;           // common (padded) number of coefficients per row
  double T[4][3];    // T[I][*], I = 1..3, right-aligned
  int node[6][3];    // (I, J, K) of every node
  int n;
};
tri_shape make_tri_shape(int n_spts)
{
  tri_shape S;
  memset(&S, 0, sizeof(S));
  S.n = n_spts;
  if (n_spts == 3)
  {
    S.len = 2;
    S.T[1][0] = 0.; S.T[1][1] = 1.;
    S.T[2][0] = 0.5; S.T[2][1] = 0.5;
    const int nd[3][3] = {{1, 1, 2}, {2, 1, 1}, {1, 2, 1}};
    memcpy(S.node, nd, sizeof(nd));
  }
  else if (n_spts == 6)
  {
    S.len = 3;
    S.T[1][0] = 0.; S.T[1][1] = 0.; S.T[1][2] = 1.;
    S.T[2][0] = 0.; S.T[2][1] = 1.; S.T[2][2] = 1.;
    S.T[3][0] = 0.5; S.T[3][1] = 0.5; S.T[3][2] = 0.;
    const int nd[6][3] = {{1, 1, 3}, {3, 1, 1}, {1, 3, 1}, {2, 1, 2}, {2, 2, 1}, {1, 2, 2}};
    memcpy(S.node, nd, sizeof(nd));
  }
  else
    FatalError("Shape order not implemented yet, exiting");
  return S;
}
double row_value(const double *c, int len, double x)
{
  double v = 0;
  for (int j = 0; j < len; j++) v += c[j] * pow(x, (double)(len - j - 1));
  return v;
}
// coefficients of sign * d/dx of a row, same length (leading zero)
void row_derivative(const double *c, int len, int sign, double *out)
{
  out[0] = 0;
  for (int j = 1; j < len; j++) out[j] = sign * c[j - 1] * (len - j);
}
double tri_shape_value(const tri_shape &S, int a, const double *coords, int deriv /* -1 value, 0 d/dr, 1 d/ds */)
{
  const double *rows[3] = {S.T[S.node[a][0]], S.T[S.node[a][1]], S.T[S.node[a][2]]};
  if (deriv < 0)
  {
    double val = 1;
    for (int i = 0; i < 3; i++) val *= row_value(rows[i], S.len, coords[i]);
    return 0 + val;
  }
  double total = 0;
  for (int layer = 0; layer < 2; layer++)
  {
    const int drow = layer == 0 ? deriv : 2;
    double d[3];
    row_derivative(rows[drow], S.len, layer == 0 ? 1 : -1, d);
    double val = 1;
    for (int i = 0; i < 3; i++) val *= row_value(i == drow ? d : rows[i], S.len, coords[i]);
    total += val;
  }
  return total;
}
} // namespace

double eles_tris::eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts)
{
  tri_shape S = make_tri_shape(in_n_spts);
  const double coords[3] = {in_loc(0), in_loc(1), -1 - in_loc(0) - in_loc(1)};
  return tri_shape_value(S, in_index, coords, -1);
}

void eles_tris::eval_d_nodal_s_basis(hf_array<double> &d_nodal_s_basis, hf_array<double> &in_loc, int in_n_spts)
{
  tri_shape S = make_tri_shape(in_n_spts);
  const double coords[3] = {in_loc(0), in_loc(1), -1 - in_loc(0) - in_loc(1)};
  for (int a = 0; a < in_n_spts; a++)
  {
    d_nodal_s_basis(a, 0) = tri_shape_value(S, a, coords, 0);
    d_nodal_s_basis(a, 1) = tri_shape_value(S, a, coords, 1);
  }
}

double eles_tris::calc_h_ref_specific(int in_ele)
{
  // diameter of the incircle of the corner triangle
  auto len = [&](int p, int q) { return sqrt(pow(shape(0, p, in_ele) - shape(0, q, in_ele), 2.0) + pow(shape(1, p, in_ele) - shape(1, q, in_ele), 2.0)); };
  const double a = len(0, 1), b = len(1, 2), c = len(2, 0);
  const double s = 0.5 * (a + b + c);
  return 2 * sqrt(((s - a) * (s - b) * (s - c)) / s);
}

// =============================================================================================================
// tetrahedra
// =============================================================================================================
void eles_tets::setup_ele_type_specific()
{
  ele_type = TET;
  n_dims = 3;
  if (run_input.equation == 0) n_fields = 5;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  n_inters_per_ele = 4;
  n_upts_per_ele = (order + 3) * (order + 2) * (order + 1) / 6;
  upts_type = run_input.upts_type_tet;

  hf_array<double> pts, w;
  cubature_tet(upts_type, order, pts, w);
  loc_upts.setup(n_dims, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int d = 0; d < 3; d++) loc_upts(d, i) = pts(i, d);
  vandermonde.setup(n_upts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int j = 0; j < n_upts_per_ele; j++) vandermonde(i, j) = eval_dubiner_basis_3d(loc_upts(0, i), loc_upts(1, i), loc_upts(2, i), j, order);
  inv_vandermonde = inv_array(vandermonde);

  const int nfi = (order + 2) * (order + 1) / 2;
  n_fpts_per_inter.setup(4);
  for (int i = 0; i < 4; i++) n_fpts_per_inter(i) = nfi;
  n_fpts_per_ele = n_inters_per_ele * nfi;

  // flux points: a triangle rule mapped onto the four faces (face 0 is the oblique one, with the first in-face
  // coordinate reversed)
  hf_array<double> tri;
  cubature_tri(run_input.fpts_type_tet, order, tri, w);
  tloc_fpts.setup(n_dims, n_fpts_per_ele);
  for (int j = 0; j < order + 1; j++)
    for (int i = 0; i < order + 1 - j; i++)
    {
      const int q = j * (order + 1) - (j - 1) * j / 2 + i;
      const int qa = j * (order + 1) - (j - 1) * j / 2 + (order - j - i);
      tloc_fpts(0, q) = tri(qa, 0);
      tloc_fpts(1, q) = tri(q, 0);
      tloc_fpts(2, q) = tri(q, 1);
      tloc_fpts(0, nfi + q) = -1;
      tloc_fpts(1, nfi + q) = tri(q, 1);
      tloc_fpts(2, nfi + q) = tri(q, 0);
      tloc_fpts(0, 2 * nfi + q) = tri(q, 0);
      tloc_fpts(1, 2 * nfi + q) = -1;
      tloc_fpts(2, 2 * nfi + q) = tri(q, 1);
      tloc_fpts(0, 3 * nfi + q) = tri(q, 1);
      tloc_fpts(1, 3 * nfi + q) = tri(q, 0);
      tloc_fpts(2, 3 * nfi + q) = -1;
    }
  tnorm_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < nfi; j++)
    {
      const int fpt = nfi * i + j;
      if (i == 0) for (int d = 0; d < 3; d++) tnorm_fpts(d, fpt) = 1. / sqrt(3.);
      else for (int d = 0; d < 3; d++) tnorm_fpts(d, fpt) = (d == i - 1) ? -1.0 : 0.;
    }

  set_opp_0(run_input.sparse_tet);
  set_opp_1(run_input.sparse_tet);
  set_opp_2(run_input.sparse_tet);
  set_opp_3(run_input.sparse_tet);
  if (viscous)
  {
    set_opp_4(run_input.sparse_tet);
    set_opp_5(run_input.sparse_tet);
    set_opp_6(run_input.sparse_tet);
  }
}

double eles_tets::eval_nodal_basis(int in_index, hf_array<double> &in_loc)
{
  hf_array<double> modal(n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++) modal(i) = eval_dubiner_basis_3d(in_loc(0), in_loc(1), in_loc(2), i, order);
  return modal_to_nodal(inv_vandermonde, in_index, modal, n_upts_per_ele);
}

double eles_tets::eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc)
{
  hf_array<double> modal(n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++) modal(i) = eval_grad_dubiner_basis_3d(in_loc(0), in_loc(1), in_loc(2), i, order, in_cpnt);
  return modal_to_nodal(inv_vandermonde, in_index, modal, n_upts_per_ele);
}

// divergence of the DG correction function of flux point in_index at loc: g.n on the face is expanded in the 2-D
// Dubiner basis, its moments against the 3-D basis are integrated with the order-7 triangle rule
double eles_tets::eval_div_dg_tet(int in_index, hf_array<double> &loc, hf_array<double> &cub, hf_array<double> &cub_w)
{
  const int nfi = n_fpts_per_inter(0);
  const int face = in_index / nfi, face_fpt = in_index - (nfi * face);
  hf_array<double> V(nfi, nfi), gdotn(nfi, 1);
  for (int i = 0; i < nfi; i++)
  {
    gdotn(i, 0) = i == face_fpt ? 1. : 0.;
    const double r = tloc_fpts(0, face * nfi + i), s = tloc_fpts(1, face * nfi + i), t = tloc_fpts(2, face * nfi + i);
    double rf, sf;
    if (face == 0) { rf = r; sf = t; }
    else if (face == 1) { rf = t; sf = s; }
    else if (face == 2) { rf = r; sf = t; }
    else { rf = s; sf = r; }
    for (int j = 0; j < nfi; j++) V(i, j) = eval_dubiner_basis_2d(rf, sf, j, order);
  }
  hf_array<double> Vi = inv_array(V);
  hf_array<double> coeff_gdotn = mult_arrays(Vi, gdotn);
  hf_array<double> coeff_divg(n_upts_per_ele, 1);
  const int ncub = cub.get_dim(0);
  for (int i = 0; i < n_upts_per_ele; i++)
  {
    double integral = 0., face_jac = 1.;
    for (int j = 0; j < ncub; j++)
    {
      const double rf = cub(j, 0), sf = cub(j, 1);
      double r, s, t;
      if (face == 0) { face_jac = sqrt(3.); r = rf; t = sf; s = -1. - t - r; }
      else if (face == 1) { face_jac = 1.; r = -1.0; s = sf; t = rf; }
      else if (face == 2) { face_jac = 1.; r = rf; s = -1.0; t = sf; }
      else { face_jac = 1.; r = sf; s = rf; t = -1.0; }
      double g = 0.;
      for (int k = 0; k < nfi; k++) g += coeff_gdotn(k, 0) * eval_dubiner_basis_2d(rf, sf, k, order);
      integral += cub_w(j) * eval_dubiner_basis_3d(r, s, t, i, order) * g;
    }
    coeff_divg(i, 0) = integral * face_jac;
  }
  double div = 0.;
  for (int i = 0; i < n_upts_per_ele; i++) div += coeff_divg(i, 0) * eval_dubiner_basis_3d(loc(0), loc(1), loc(2), i, order);
  return div;
}

void eles_tets::fill_opp_3(hf_array<double> &opp_3)
{
  const int n = n_upts_per_ele;
  const double c_tet = vcjh_c_simplex(run_input.vcjh_scheme_tet, order, run_input.c_tet, 3);
  run_input.c_tet = c_tet;
  hf_array<double> tr(n, n), ts(n, n), tt(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
    {
      tr(i, j) = eval_grad_dubiner_basis_3d(loc_upts(0, i), loc_upts(1, i), loc_upts(2, i), j, order, 0);
      ts(i, j) = eval_grad_dubiner_basis_3d(loc_upts(0, i), loc_upts(1, i), loc_upts(2, i), j, order, 1);
      tt(i, j) = eval_grad_dubiner_basis_3d(loc_upts(0, i), loc_upts(1, i), loc_upts(2, i), j, order, 2);
    }
  vector<hf_array<double>> D(3);
  D[0] = mult_arrays(tr, inv_vandermonde);
  D[1] = mult_arrays(ts, inv_vandermonde);
  D[2] = mult_arrays(tt, inv_vandermonde);
  vector<vector<int>> powers;
  vector<double> coeff;
  for (int v = 1; v <= (order + 1); v++)
    for (int w = 1; w <= v; w++)
    {
      powers.push_back({order - v + 1, v - w, w - 1});
      coeff.push_back((1. / n) * (factorial_i(order) / (factorial_i(v - 1) * factorial_i(order - (v - 1)))) *
                      (factorial_i(v - 1) / (factorial_i(w - 1) * factorial_i((v - 1) - (w - 1)))));
    }
  hf_array<double> Filt = vcjh_filter(vandermonde, D, powers, coeff, c_tet, n);

  hf_array<double> dg(n, n_fpts_per_ele), loc(n_dims), cub, cub_w;
  cubature_tri(0, 7, cub, cub_w);
  for (int i = 0; i < n_fpts_per_ele; i++)
    for (int j = 0; j < n; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = loc_upts(k, j);
      dg(j, i) = eval_div_dg_tet(i, loc, cub, cub_w);
    }
  opp_3 = mult_arrays(Filt, dg);
}

double eles_tets::eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts)
{
  const double r = in_loc(0), s = in_loc(1), t = in_loc(2);
  if (in_n_spts == 4)
  {
    switch (in_index)
    {
    case 0: return -0.5 * (r + s + t + 1.);
    case 1: return 0.5 * (r + 1.);
    case 2: return 0.5 * (s + 1.);
    default: return 0.5 * (t + 1.);
    }
  }
  if (in_n_spts == 10)
  {
    switch (in_index)
    {
    case 0: return (1. / 2. * (2. + r + s + t)) * (r + 1. + s + t);
    case 1: return (1. / 2.) * r * (r + 1.);
    case 2: return (1. / 2.) * s * (s + 1.);
    case 3: return (1. / 2.) * t * (t + 1.);
    case 4: return -(r + 1. + s + t) * (r + 1.);
    case 5: return -(r + 1. + s + t) * (s + 1.);
    case 6: return -(r + 1. + s + t) * (t + 1.);
    case 7: return (r + 1.) * (s + 1.);
    case 8: return (s + 1.) * (t + 1.);
    default: return (t + 1.) * (r + 1.);
    }
  }
  FatalError("Shape order not implemented yet, exiting");
  return 0.;
}

void eles_tets::eval_d_nodal_s_basis(hf_array<double> &d, hf_array<double> &in_loc, int in_n_spts)
{
  const double r = in_loc(0), s = in_loc(1), t = in_loc(2);
  if (in_n_spts == 4)
  {
    const double v[4][3] = {{-0.5, -0.5, -0.5}, {0.5, 0., 0.}, {0., 0.5, 0.}, {0., 0., 0.5}};
    for (int a = 0; a < 4; a++)
      for (int c = 0; c < 3; c++) d(a, c) = v[a][c];
  }
  else if (in_n_spts == 10)
  {
    const double all = 1.5 + r + s + t;
    d(0, 0) = all;                     d(0, 1) = all;                     d(0, 2) = all;
    d(1, 0) = r + 0.5;                 d(1, 1) = 0.;                      d(1, 2) = 0.;
    d(2, 0) = 0.;                      d(2, 1) = s + 0.5;                 d(2, 2) = 0.;
    d(3, 0) = 0.;                      d(3, 1) = 0.;                      d(3, 2) = t + 0.5;
    d(4, 0) = -2. * r - 2. - s - t;    d(4, 1) = -r - 1.;                 d(4, 2) = -r - 1.;
    d(5, 0) = -s - 1.;                 d(5, 1) = -2. * s - 2. - r - t;    d(5, 2) = -s - 1.;
    d(6, 0) = -t - 1.;                 d(6, 1) = -t - 1.;                 d(6, 2) = -2. * t - 2. - r - s;
    d(7, 0) = s + 1.;                  d(7, 1) = r + 1.;                  d(7, 2) = 0.;
    d(8, 0) = 0.;                      d(8, 1) = t + 1.;                  d(8, 2) = s + 1.;
    d(9, 0) = t + 1.;                  d(9, 1) = 0.;                      d(9, 2) = r + 1.;
  }
  else
    FatalError("Shape order not implemented yet, exiting");
}

double eles_tets::calc_h_ref_specific(int in_ele)
{
  // diameter of the insphere: 6 V / total face area
  double a[3], b[3], c[3], d[3], e[3];
  for (int i = 0; i < 3; i++)
  {
    a[i] = shape(i, 1, in_ele) - shape(i, 0, in_ele);
    b[i] = shape(i, 2, in_ele) - shape(i, 0, in_ele);
    c[i] = shape(i, 3, in_ele) - shape(i, 0, in_ele);
    d[i] = shape(i, 2, in_ele) - shape(i, 1, in_ele);
    e[i] = shape(i, 3, in_ele) - shape(i, 1, in_ele);
  }
  const double trip = (a[0] * b[1] * c[2] + b[0] * c[1] * a[2] + c[0] * a[1] * b[2]) - (c[0] * b[1] * a[2] + b[0] * a[1] * c[2] + a[0] * c[1] * b[2]);
  const double vol = 1. / 6. * trip;
  auto area = [](const double *p, const double *q) {
    return 0.5 * sqrt(pow(p[1] * q[2] - p[2] * q[1], 2) + pow(p[0] * q[2] - p[2] * q[0], 2) + pow(p[0] * q[1] - p[1] * q[0], 2));
  };
  const double s_a = area(a, b), s_b = area(a, c), s_c = area(b, c), s_d = area(d, e);
  return 6. * vol / (s_a + s_b + s_c + s_d);
}

// =============================================================================================================
// prisms
// =============================================================================================================
void eles_pris::setup_ele_type_specific()
{
  ele_type = PRISM;
  n_dims = 3;
  if (run_input.equation == 0) n_fields = 5;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  n_inters_per_ele = 5;
  n_upts_per_ele = (order + 2) * (order + 1) * (order + 1) / 2;
  n_upts_tri = (order + 1) * (order + 2) / 2;
  n_upts_1d = order + 1;

  hf_array<double> w, tri;
  cubature_1d(run_input.upts_type_pri_1d, order, loc_upts_pri_1d, w);
  cubature_tri(run_input.upts_type_pri_tri, order, tri, w);
  loc_upts_pri_tri.setup(2, n_upts_tri);
  for (int i = 0; i < n_upts_tri; i++) { loc_upts_pri_tri(0, i) = tri(i, 0); loc_upts_pri_tri(1, i) = tri(i, 1); }
  loc_upts.setup(n_dims, n_upts_per_ele);
  for (int i = 0; i < n_upts_1d; i++)
    for (int j = 0; j < n_upts_tri; j++)
    {
      loc_upts(0, n_upts_tri * i + j) = loc_upts_pri_tri(0, j);
      loc_upts(1, n_upts_tri * i + j) = loc_upts_pri_tri(1, j);
      loc_upts(2, n_upts_tri * i + j) = loc_upts_pri_1d(i);
    }
  vandermonde_tri.setup(n_upts_tri, n_upts_tri);
  for (int i = 0; i < n_upts_tri; i++)
    for (int j = 0; j < n_upts_tri; j++) vandermonde_tri(i, j) = eval_dubiner_basis_2d(loc_upts_pri_tri(0, i), loc_upts_pri_tri(1, i), j, order);
  inv_vandermonde_tri = inv_array(vandermonde_tri);

  const int n1 = order + 1;
  n_fpts_per_inter.setup(5);
  n_fpts_per_inter(0) = n_upts_tri;
  n_fpts_per_inter(1) = n_upts_tri;
  for (int i = 2; i < 5; i++) n_fpts_per_inter(i) = n1 * n1;
  n_fpts_per_ele = 3 * n1 * n1 + (order + 2) * (order + 1);
  if (run_input.upts_type_pri_tri != run_input.fpts_type_tet) FatalError("upts_type_pri_tri != fpts_type_tet");
  if (run_input.upts_type_pri_1d != run_input.upts_type_hexa) FatalError("upts_type_pri_1d != upts_type_hexa");

  // flux points: the two triangles (the bottom one with r and s swapped), then the three quadrilateral sides in the
  // order of the triangle's edges, in-face index = (z index, edge index)
  loc_1d_fpts = loc_upts_pri_1d;
  tloc_fpts.setup(n_dims, n_fpts_per_ele);
  for (int i = 0; i < n_upts_tri; i++)
  {
    tloc_fpts(0, i) = tri(i, 1); tloc_fpts(1, i) = tri(i, 0); tloc_fpts(2, i) = -1.;
    tloc_fpts(0, n_upts_tri + i) = tri(i, 0); tloc_fpts(1, n_upts_tri + i) = tri(i, 1); tloc_fpts(2, n_upts_tri + i) = 1.;
  }
  const int offset = 2 * n_upts_tri;
  for (int face = 0; face < 3; face++)
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n1; j++)
      {
        const int q = offset + face * n1 * n1 + i * n1 + j;
        if (face == 0) { tloc_fpts(0, q) = loc_1d_fpts(j); tloc_fpts(1, q) = -1; }
        else if (face == 1) { tloc_fpts(0, q) = loc_1d_fpts(order - j); tloc_fpts(1, q) = loc_1d_fpts(j); }
        else { tloc_fpts(0, q) = -1.; tloc_fpts(1, q) = loc_1d_fpts(order - j); }
        tloc_fpts(2, q) = loc_1d_fpts(i);
      }
  tnorm_fpts.setup(n_dims, n_fpts_per_ele);
  int fpt = -1;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < n_fpts_per_inter(i); j++)
    {
      fpt++;
      double nx = 0., ny = 0., nz = 0.;
      if (i == 0) nz = -1.;
      else if (i == 1) nz = 1.;
      else if (i == 2) ny = -1.;
      else if (i == 3) { nx = 1. / sqrt(2.); ny = 1. / sqrt(2.); }
      else nx = -1.;
      tnorm_fpts(0, fpt) = nx; tnorm_fpts(1, fpt) = ny; tnorm_fpts(2, fpt) = nz;
    }

  set_opp_0(run_input.sparse_pri);
  set_opp_1(run_input.sparse_pri);
  set_opp_2(run_input.sparse_pri);
  set_opp_3(run_input.sparse_pri);
  if (viscous)
  {
    set_opp_4(run_input.sparse_pri);
    set_opp_5(run_input.sparse_pri);
    set_opp_6(run_input.sparse_pri);
  }
}

double eles_pris::tri_part(int index_tri, int cpnt, hf_array<double> &in_loc)
{
  hf_array<double> modal(n_upts_tri);
  for (int i = 0; i < n_upts_tri; i++)
  {
    if (cpnt < 0) modal(i) = eval_dubiner_basis_2d(in_loc(0), in_loc(1), i, order);
    else if (cpnt == 0) modal(i) = eval_dr_dubiner_basis_2d(in_loc(0), in_loc(1), i, order);
    else modal(i) = eval_ds_dubiner_basis_2d(in_loc(0), in_loc(1), i, order);
  }
  return modal_to_nodal(inv_vandermonde_tri, index_tri, modal, n_upts_tri);
}

double eles_pris::eval_nodal_basis(int in_index, hf_array<double> &in_loc)
{
  const int index_tri = in_index % n_upts_tri, index_1d = in_index / n_upts_tri;
  const double tri = tri_part(index_tri, -1, in_loc);
  const double oned = eval_lagrange(in_loc(2), index_1d, loc_upts_pri_1d);
  return (tri * oned);
}

double eles_pris::eval_d_nodal_basis(int in_index, int in_cpnt, hf_array<double> &in_loc)
{
  const int index_tri = in_index % n_upts_tri, index_1d = in_index / n_upts_tri;
  if (in_cpnt == 0 || in_cpnt == 1)
  {
    const double dtri = tri_part(index_tri, in_cpnt, in_loc);
    const double oned = eval_lagrange(in_loc(2), index_1d, loc_upts_pri_1d);
    return dtri * oned;
  }
  const double tri = tri_part(index_tri, -1, in_loc);
  const double doned = eval_d_lagrange(in_loc(2), index_1d, loc_upts_pri_1d);
  return tri * doned;
}

// bottom-face flux point -> triangle solution point underneath (r and s are swapped on that face)
int eles_pris::face0_map(int index)
{
  for (int j = 0; j < (order + 1); j++)
    for (int i = 0; i < (order + 1) - j; i++)
      if (j * (order + 1) - (j - 1) * j / 2 + i == index) return (i * (order + 1) - (i - 1) * i / 2 + j);
  FatalError("Should not be here in face0_map, exiting");
  return -1;
}

void eles_pris::fill_opp_3(hf_array<double> &opp_3)
{
  hf_array<double> opp_3_tri(n_upts_tri, 3 * (order + 1));
  get_opp_3_tri(opp_3_tri, loc_upts_pri_tri, loc_1d_fpts, vandermonde_tri, inv_vandermonde_tri, n_upts_tri, order, run_input.c_tri, run_input.vcjh_scheme_tri);
  const double eta = run_input.vcjh_scheme_pri_1d == 0 ? run_input.eta_pri : compute_eta(run_input.vcjh_scheme_pri_1d, order);
  const int n1 = order + 1, nt = n_upts_tri;
  for (int upt = 0; upt < n_upts_per_ele; upt++)
  {
    const double z = loc_upts(2, upt);
    const int upt_1d = upt / nt, upt_tri = upt % nt;
    for (int q = 0; q < n_fpts_per_ele; q++)
    {
      double v = 0.;
      if (q < nt) { if (face0_map(q) == upt_tri) v = -eval_d_vcjh_1d(z, 0, order, eta); }
      else if (q < 2 * nt) { if (q - nt == upt_tri) v = eval_d_vcjh_1d(z, 1, order, eta); }
      else
      {
        const int edge = (q - 2 * nt) / (n1 * n1), face_fpt = (q - 2 * nt) - edge * n1 * n1;
        if (face_fpt / n1 == upt_1d) v = opp_3_tri(upt_tri, edge * n1 + face_fpt % n1);
      }
      opp_3(upt, q) = v;
    }
  }
}

double eles_pris::eval_nodal_s_basis(int in_index, hf_array<double> &in_loc, int in_n_spts)
{
  const double r = in_loc(0), s = in_loc(1), t = in_loc(2);
  if (in_n_spts == 6)
  {
    switch (in_index)
    {
    case 0: return 1. / 4. * (r + s) * (t - 1.);
    case 1: return -1. / 4. * (r + 1.) * (t - 1.);
    case 2: return -1. / 4. * (s + 1.) * (t - 1.);
    case 3: return -1. / 4. * (r + s) * (t + 1.);
    case 4: return 1. / 4. * (r + 1.) * (t + 1.);
    default: return 1. / 4. * (s + 1.) * (t + 1.);
    }
  }
  if (in_n_spts == 15)
  {
    switch (in_index)
    {
    case 0: return (1. / 4 * (r + s)) * (r + s + 1.) * t * (t - 1.);
    case 1: return (1. / 4) * r * (r + 1.) * t * (t - 1.);
    case 2: return (1. / 4) * s * (s + 1.) * t * (t - 1.);
    case 3: return (1. / 4 * (r + s)) * (r + s + 1.) * t * (t + 1.);
    case 4: return (1. / 4) * r * (r + 1.) * t * (t + 1.);
    case 5: return (1. / 4) * s * (s + 1.) * t * (t + 1.);
    case 6: return -(1. / 2 * (r + s)) * (r + 1.) * t * (t - 1.);
    case 7: return (1. / 2 * (r + 1.)) * (s + 1.) * t * (t - 1.);
    case 8: return -(1. / 2 * (r + s)) * (s + 1.) * t * (t - 1.);
    case 9: return (1. / 2 * (r + s)) * (t * t - 1.);
    case 10: return -(1. / 2 * (r + 1.)) * (t * t - 1.);
    case 11: return -(1. / 2 * (s + 1.)) * (t * t - 1.);
    case 12: return -(1. / 2 * (r + s)) * (r + 1.) * t * (t + 1.);
    case 13: return (1. / 2 * (r + 1.)) * (s + 1.) * t * (t + 1.);
    default: return -(1. / 2 * (r + s)) * (s + 1.) * t * (t + 1.);
    }
  }
  FatalError("Shape order not implemented yet, exiting");
  return 0.;
}

void eles_pris::eval_d_nodal_s_basis(hf_array<double> &d, hf_array<double> &in_loc, int in_n_spts)
{
  const double r = in_loc(0), s = in_loc(1), t = in_loc(2);
  if (in_n_spts == 6)
  {
    d(0, 0) = 1. / 4. * (t - 1.);   d(0, 1) = 1. / 4. * (t - 1.);   d(0, 2) = 1. / 4. * (r + s);
    d(1, 0) = -1. / 4. * (t - 1.);  d(1, 1) = 0.;                   d(1, 2) = -1. / 4. * (r + 1.);
    d(2, 0) = 0;                    d(2, 1) = -1. / 4. * (t - 1.);  d(2, 2) = -1. / 4. * (s + 1.);
    d(3, 0) = -1. / 4. * (t + 1.);  d(3, 1) = -1. / 4. * (t + 1.);  d(3, 2) = -1. / 4. * (r + s);
    d(4, 0) = 1. / 4. * (t + 1.);   d(4, 1) = 0.;                   d(4, 2) = 1. / 4. * (r + 1.);
    d(5, 0) = 0.;                   d(5, 1) = 1. / 4. * (t + 1.);   d(5, 2) = 1. / 4. * (s + 1.);
  }
  else if (in_n_spts == 15)
  {
    d(0, 0) = (1. / 4) * t * (t - 1.) * (2 * r + 2 * s + 1.);
    d(1, 0) = (1. / 4) * t * (t - 1.) * (2 * r + 1.);
    d(2, 0) = 0.;
    d(3, 0) = (1. / 4) * t * (t + 1.) * (2 * r + 2 * s + 1.);
    d(4, 0) = (1. / 4) * t * (t + 1.) * (2 * r + 1.);
    d(5, 0) = 0.;
    d(6, 0) = -(1. / 2) * t * (t - 1.) * (2 * r + 1. + s);
    d(7, 0) = (1. / 2 * (s + 1.)) * t * (t - 1.);
    d(8, 0) = -(1. / 2 * (s + 1.)) * t * (t - 1.);
    d(9, 0) = (1. / 2) * t * t - 1. / 2;
    d(10, 0) = -(1. / 2) * t * t + 1. / 2;
    d(11, 0) = 0.;
    d(12, 0) = -(1. / 2) * t * (t + 1.) * (2 * r + 1. + s);
    d(13, 0) = (1. / 2 * (s + 1.)) * t * (t + 1.);
    d(14, 0) = -(1. / 2 * (s + 1.)) * t * (t + 1.);

    d(0, 1) = (1. / 4) * t * (t - 1.) * (2 * r + 2 * s + 1.);
    d(1, 1) = 0.;
    d(2, 1) = (1. / 4) * t * (t - 1.) * (2 * s + 1.);
    d(3, 1) = (1. / 4) * t * (t + 1.) * (2 * r + 2 * s + 1.);
    d(4, 1) = 0.;
    d(5, 1) = (1. / 4) * t * (t + 1.) * (2 * s + 1.);
    d(6, 1) = -(1. / 2 * (r + 1.)) * t * (t - 1.);
    d(7, 1) = (1. / 2 * (r + 1.)) * t * (t - 1.);
    d(8, 1) = -(1. / 2) * t * (t - 1.) * (2 * s + 1. + r);
    d(9, 1) = (1. / 2) * t * t - 1. / 2;
    d(10, 1) = 0.;
    d(11, 1) = -(1. / 2) * t * t + 1. / 2;
    d(12, 1) = -(1. / 2 * (r + 1.)) * t * (t + 1.);
    d(13, 1) = (1. / 2 * (r + 1.)) * t * (t + 1.);
    d(14, 1) = -(1. / 2) * t * (t + 1.) * (2 * s + 1. + r);

    d(0, 2) = (1. / 4 * (r + s + 1.)) * (r + s) * (2 * t - 1.);
    d(1, 2) = (1. / 4) * r * (2 * t - 1.) * (r + 1.);
    d(2, 2) = (1. / 4) * s * (2 * t - 1.) * (s + 1.);
    d(3, 2) = (1. / 4 * (r + s + 1.)) * (r + s) * (2 * t + 1.);
    d(4, 2) = (1. / 4) * r * (2 * t + 1.) * (r + 1.);
    d(5, 2) = (1. / 4) * s * (2 * t + 1.) * (s + 1.);
    d(6, 2) = -(1. / 2 * (2 * t - 1.)) * (r + 1.) * (r + s);
    d(7, 2) = (1. / 2 * (2 * t - 1.)) * (s + 1.) * (r + 1.);
    d(8, 2) = -(1. / 2 * (2 * t - 1.)) * (s + 1.) * (r + s);
    d(9, 2) = t * (r + s);
    d(10, 2) = -t * (r + 1.);
    d(11, 2) = -t * (s + 1.);
    d(12, 2) = -(1. / 2 * (2 * t + 1.)) * (r + 1.) * (r + s);
    d(13, 2) = (1. / 2 * (2 * t + 1.)) * (s + 1.) * (r + 1.);
    d(14, 2) = -(1. / 2 * (2 * t + 1.)) * (s + 1.) * (r + s);
  }
  else
    FatalError("Shape order not implemented yet, exiting");
}

double eles_pris::calc_h_ref_specific(int in_ele)
{
  // smallest of the three vertical edges and the incircle diameters of the two triangles
  auto len = [&](int p, int q) {
    return sqrt(pow(shape(0, p, in_ele) - shape(0, q, in_ele), 2.0) + pow(shape(1, p, in_ele) - shape(1, q, in_ele), 2.0) +
                pow(shape(2, p, in_ele) - shape(2, q, in_ele), 2.0));
  };
  double h = 1e300;
  for (int i = 0; i < 3; i++) h = min(h, len(i, i + 3));
  for (int i = 3; i < 5; i++)
  {
    const int d = (i - 3) * 3;
    const double a = len(d, d + 1), b = len(d + 1, d + 2), c = len(d + 2, d);
    const double s = 0.5 * (a + b + c);
    h = min(h, 2 * sqrt(((s - a) * (s - b) * (s - c)) / s));
  }
  return h;
}
