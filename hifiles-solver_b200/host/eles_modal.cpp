// Modal (hierarchical orthogonal) bases of the element types and the two set-up products built on them:
//   * polynomial de-aliasing by over-integration: interpolation to a stronger volume cubature, flux there, L2 projection
//     back onto the solution basis                       reference src/eles.cpp:1480-1545, <type>::set_over_int
//   * Persson's modal shock sensor and the exponential modal filter      reference src/eles.cpp:2918-2959,
//     <type>::shock_det_persson, set_exp_filter, calc_norm_basis
// The reference repeats this per element type; here a type only describes its modes (index triple, basis value, squared
// norm, which modes form the highest-degree shell) and the constructions are written once.  Mode numbering, norms and the
// order of every product follow the reference, so the matrices are the reference's numbers:
//   hexahedra   Legendre products, modes by (sum, k, j)          src/eles_hexas.cpp:938-1059, 1096-1129, 1364-1440
//   quads       Legendre products, modes by (sum, j)             src/eles_quads.cpp:786-905, 928-959, 1116-1190
//   triangles   Dubiner modes by total degree                    src/eles_tris.cpp:432-526, 674-701
//   tetrahedra  Dubiner modes by total degree                    src/eles_tets.cpp:705-802, 946-975
//   prisms      Dubiner (r,s) x Legendre (t), modes by (sum, k, j)   src/eles_pris.cpp:609-733, 938-971, 1238-1321
#include "hifiles.h"

using namespace std;

// ---- mode tables -------------------------------------------------------------------------------------------------------
void eles::set_modes()
{
  modes.clear();
  const int o = order;
  if (ele_type == HEX)
  {
    for (int l = 0; l < 3 * o + 1; l++)
      for (int k = 0; k < l + 1; k++)
        for (int j = 0; j < l - k + 1; j++)
        {
          const int i = l - k - j;
          if (i <= o && j <= o && k <= o) modes.push_back({i, j, k});
        }
  }
  else if (ele_type == QUAD)
  {
    for (int k = 0; k < 2 * o + 1; k++)
      for (int j = 0; j < k + 1; j++)
      {
        const int i = k - j;
        if (i <= o && j <= o) modes.push_back({i, j, 0});
      }
  }
  else if (ele_type == TRI)
  {
    for (int k = 0; k < o + 1; k++)
      for (int j = 0; j < k + 1; j++) modes.push_back({k - j, j, 0});
  }
  else if (ele_type == TET)
  {
    for (int m = 0; m < o + 1; m++)
      for (int n = 0; n < m + 1; n++)
        for (int k = 0; k < n + 1; k++) modes.push_back({m - n, n - k, k});
  }
  else if (ele_type == PRISM)
  {
    for (int l = 0; l < 2 * o + 1; l++)
      for (int k = 0; k < l + 1; k++)
        for (int j = 0; j < l - k + 1; j++)
        {
          const int i = l - k - j;
          if (k <= o && i + j <= o) modes.push_back({i, j, k});
        }
  }
  if ((int)modes.size() != n_upts_per_ele) FatalError("modal basis size does not match the number of solution points");
}

double eles::eval_modal_basis(int m, hf_array<double> &loc)
{
  const modal_mode &q = modes[m];
  switch (ele_type)
  {
  case HEX: return eval_legendre(loc(0), q.i) * eval_legendre(loc(1), q.j) * eval_legendre(loc(2), q.k);
  case QUAD: return eval_legendre(loc(0), q.i) * eval_legendre(loc(1), q.j);
  case TRI: return eval_dubiner_basis_2d(loc(0), loc(1), m, order);
  case TET: return eval_dubiner_basis_3d(loc(0), loc(1), loc(2), m, order);
  default:
  {
    // sqrt(2) P_i^{0,0}(a) P_j^{2i+1,0}(b) (1-b)^i L_k(t): the triangle's Dubiner mode (i, j) times a Legendre polynomial
    const int tri_mode = (q.i + q.j) * (q.i + q.j + 1) / 2 + q.j;
    return eval_dubiner_basis_2d(loc(0), loc(1), tri_mode, order) * eval_legendre(loc(2), q.k);
  }
  }
}

// squared L2 norm of a mode (1 for the orthonormal Dubiner modes)
double eles::modal_norm(int m)
{
  const modal_mode &q = modes[m];
  const double n1 = 2.0 / (2.0 * q.i + 1.0), n2 = 2.0 / (2.0 * q.j + 1.0), n3 = 2.0 / (2.0 * q.k + 1.0);
  switch (ele_type)
  {
  case HEX: return n1 * n2 * n3;
  case QUAD: return n1 * n2;
  case PRISM: return n3;
  default: return 1.0;
  }
}

// is the mode part of the highest-degree shell (what Persson's sensor measures)
bool eles::mode_is_top(int m)
{
  const modal_mode &q = modes[m];
  switch (ele_type)
  {
  case HEX: return q.i == order || q.j == order || q.k == order;
  case QUAD: return q.i == order || q.j == order;
  case PRISM: return q.i + q.j == order || q.k == order;
  case TRI: return m >= order * (order + 1) / 2;
  default: return m >= order * (order + 1) * (order + 2) / 6;
  }
}

void eles::set_modal_vandermonde()
{
  if (modal_vandermonde.get_dim(0) == n_upts_per_ele) return;
  set_modes();
  hf_array<double> loc(n_dims);
  modal_vandermonde.setup(n_upts_per_ele, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
  {
    for (int d = 0; d < n_dims; d++) loc(d) = loc_upts(d, i);
    for (int j = 0; j < n_upts_per_ele; j++) modal_vandermonde(i, j) = eval_modal_basis(j, loc);
  }
  modal_inv_vandermonde = inv_array(modal_vandermonde);
}

// ---- volume cubature of the over-integration order -----------------------------------------------------------------------
void eles::set_volume_cubpts(int in_order, hf_array<double> &locs, hf_array<double> &weights)
{
  hf_array<double> r1, w1, tri, wt;
  const int n = in_order + 1;
  if (ele_type == HEX)
  {
    cubature_1d(0, in_order, r1, w1);
    locs.setup(3, n * n * n);
    weights.setup(n * n * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        for (int k = 0; k < n; k++)
        {
          const int q = k + n * j + n * n * i;
          locs(0, q) = r1(k); locs(1, q) = r1(j); locs(2, q) = r1(i);
          weights(q) = w1(i) * w1(j) * w1(k);
        }
  }
  else if (ele_type == QUAD)
  {
    cubature_1d(0, in_order, r1, w1);
    locs.setup(2, n * n);
    weights.setup(n * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
      {
        const int q = j + n * i;
        locs(0, q) = r1(j); locs(1, q) = r1(i);
        weights(q) = w1(j) * w1(i);
      }
  }
  else if (ele_type == TRI)
  {
    cubature_tri(0, in_order, tri, wt);
    const int np = tri.get_dim(0);
    locs.setup(2, np);
    weights.setup(np);
    for (int q = 0; q < np; q++) { locs(0, q) = tri(q, 0); locs(1, q) = tri(q, 1); weights(q) = wt(q); }
  }
  else if (ele_type == TET)
  {
    cubature_tet(0, in_order, tri, wt);
    const int np = tri.get_dim(0);
    locs.setup(3, np);
    weights.setup(np);
    for (int q = 0; q < np; q++) { for (int d = 0; d < 3; d++) locs(d, q) = tri(q, d); weights(q) = wt(q); }
  }
  else
  {
    cubature_tri(0, in_order, tri, wt);
    cubature_1d(0, in_order, r1, w1);
    const int nt = tri.get_dim(0);
    locs.setup(3, nt * n);
    weights.setup(nt * n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < nt; j++)
      {
        const int q = j + nt * i;
        locs(0, q) = tri(j, 0); locs(1, q) = tri(j, 1); locs(2, q) = r1(i);
        weights(q) = wt(j) * w1(i);
      }
  }
}

// ---- over-integration ------------------------------------------------------------------------------------------------------
void eles::set_over_int()
{
  set_modal_vandermonde();
  set_volume_cubpts(run_input.over_int_order, loc_over_int_cubpts, weight_over_int_cubpts);
  const int nc = loc_over_int_cubpts.get_dim(1);
  hf_array<double> loc(n_dims);
  // interpolation solution points -> cubature points
  opp_over_int_cubpts.setup(nc, n_upts_per_ele);
  for (int i = 0; i < n_upts_per_ele; i++)
    for (int j = 0; j < nc; j++)
    {
      for (int d = 0; d < n_dims; d++) loc(d) = loc_over_int_cubpts(d, j);
      opp_over_int_cubpts(j, i) = eval_nodal_basis(i, loc);
    }
  // L2 projection cubature points -> modal coefficients, then modal -> nodal
  hf_array<double> proj(n_upts_per_ele, nc);
  const bool normalised = ele_type == TRI || ele_type == TET;
  for (int i = 0; i < n_upts_per_ele; i++)
  {
    const modal_mode &q = modes[i];
    const double n1 = 2.0 / (2.0 * q.i + 1.0), n2 = 2.0 / (2.0 * q.j + 1.0), n3 = 2.0 / (2.0 * q.k + 1.0);
    for (int j = 0; j < nc; j++)
    {
      for (int d = 0; d < n_dims; d++) loc(d) = loc_over_int_cubpts(d, j);
      const double b = eval_modal_basis(i, loc);
      if (normalised) proj(i, j) = b * weight_over_int_cubpts(j);
      else if (ele_type == HEX) proj(i, j) = b / (n1 * n2 * n3) * weight_over_int_cubpts(j);
      else if (ele_type == QUAD) proj(i, j) = b / (n1 * n2) * weight_over_int_cubpts(j);
      else proj(i, j) = b / n3 * weight_over_int_cubpts(j);
    }
  }
  over_int_filter = mult_arrays(modal_vandermonde, proj);
}

// adj(J) at the over-integration points (reference src/eles.cpp:4150-4213)
void eles::set_transforms_over_int_cubpts()
{
  const int nc = loc_over_int_cubpts.get_dim(1);
  hf_array<double> loc(n_dims), d_pos(n_dims, n_dims);
  JGinv_over_int_cubpts.setup(n_dims, n_dims, nc, n_eles);
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < nc; j++)
    {
      for (int k = 0; k < n_dims; k++) loc(k) = loc_over_int_cubpts(k, j);
      calc_d_pos(loc, i, d_pos);
      if (n_dims == 2)
      {
        const double xr = d_pos(0, 0), xs = d_pos(0, 1), yr = d_pos(1, 0), ys = d_pos(1, 1);
        JGinv_over_int_cubpts(0, 0, j, i) = ys;
        JGinv_over_int_cubpts(0, 1, j, i) = -xs;
        JGinv_over_int_cubpts(1, 0, j, i) = -yr;
        JGinv_over_int_cubpts(1, 1, j, i) = xr;
      }
      else
      {
        const double xr = d_pos(0, 0), xs = d_pos(0, 1), xt = d_pos(0, 2);
        const double yr = d_pos(1, 0), ys = d_pos(1, 1), yt = d_pos(1, 2);
        const double zr = d_pos(2, 0), zs = d_pos(2, 1), zt = d_pos(2, 2);
        JGinv_over_int_cubpts(0, 0, j, i) = ys * zt - yt * zs;
        JGinv_over_int_cubpts(0, 1, j, i) = xt * zs - xs * zt;
        JGinv_over_int_cubpts(0, 2, j, i) = xs * yt - xt * ys;
        JGinv_over_int_cubpts(1, 0, j, i) = yt * zr - yr * zt;
        JGinv_over_int_cubpts(1, 1, j, i) = xr * zt - xt * zr;
        JGinv_over_int_cubpts(1, 2, j, i) = xt * yr - xr * yt;
        JGinv_over_int_cubpts(2, 0, j, i) = yr * zs - ys * zr;
        JGinv_over_int_cubpts(2, 1, j, i) = xs * zr - xr * zs;
        JGinv_over_int_cubpts(2, 2, j, i) = xr * ys - xs * yr;
      }
    }
}

// ---- shock capturing --------------------------------------------------------------------------------------------------------
void eles::set_shock_capture()
{
  if (run_input.shock_det != 0) FatalError("Shock detector not implemented.");
  if (run_input.shock_cap != 1) FatalError("Shock capturing method not implemented.");
  set_modal_vandermonde();
  const int n = n_upts_per_ele;
  // Persson: sensor = sum_top w_j uhat_j^2 / sum_all w_j uhat_j^2, w_j = squared norm of mode j
  sensor_w_all.setup(n);
  sensor_w_top.setup(n);
  for (int j = 0; j < n; j++)
  {
    sensor_w_all(j) = modal_norm(j);
    sensor_w_top(j) = mode_is_top(j) ? sensor_w_all(j) : 0.;
  }
  // exponential modal filter sigma(eta) = exp(-fac ((eta - eta_c) / (1 - eta_c))^order) above the cut-off, per direction
  // for the tensor-product parts, by total degree for the simplex parts
  const double eta_c = (double)run_input.expf_cutoff / (double)(order);
  auto sigma = [&](double eta) { return exp(-run_input.expf_fac * pow((eta - eta_c) / (1. - eta_c), run_input.expf_order)); };
  hf_array<double> diag(n, n);
  for (int m = 0; m < n; m++)
  {
    const modal_mode &q = modes[m];
    double f;
    if (ele_type == TRI || ele_type == TET)
    {
      const double eta = (double)(q.i + q.j + q.k) / (double)(order);
      f = eta <= eta_c ? 1 : sigma(eta);
    }
    else
    {
      f = 1.;
      double etas[3];
      int ne = 0;
      if (ele_type == PRISM) { etas[ne++] = (double)(q.i + q.j) / (double)(order); etas[ne++] = (double)(q.k) / (double)(order); }
      else
      {
        etas[ne++] = (double)(q.i) / (double)(order);
        etas[ne++] = (double)(q.j) / (double)(order);
        if (ele_type == HEX) etas[ne++] = (double)(q.k) / (double)(order);
      }
      for (int e = 0; e < ne; e++)
        if (etas[e] > eta_c) f *= sigma(etas[e]);
    }
    diag(m, m) = f;
  }
  hf_array<double> t = mult_arrays(diag, modal_inv_vandermonde);
  exp_filter = mult_arrays(modal_vandermonde, t);
  sensor.setup(n_eles);
}

// ---- LES: wall distance -----------------------------------------------------------------------------------------------------
// vector from the nearest no-slip wall flux point to every solution point; the first of equally near points wins;
// (1e20, ...) when the mesh has no such wall
void eles::calc_wall_distance(std::vector<hf_array<double>> &loc_noslip_bdy)
{
  if (n_eles == 0) return;
  wall_distance.setup(n_upts_per_ele, n_eles, n_dims);
  double vec[3], vecmin[3] = {1e20, 1e20, 1e20};
  for (int i = 0; i < n_eles; ++i)
    for (int j = 0; j < n_upts_per_ele; ++j)
    {
      double distmin = 1e20;
      for (int t = 0; t < 3; t++)
      {
        hf_array<double> &W = loc_noslip_bdy[t];
        const int nw = (int)(W.size() ? W.get_dim(2) : 0), nfp = W.get_dim(1);
        for (int k = 0; k < nw; ++k)
          for (int m = 0; m < nfp; ++m)
          {
            double dist = 0.0;
            for (int n = 0; n < n_dims; ++n)
            {
              vec[n] = pos_upts(j, i, n) - W(n, m, k);
              dist += vec[n] * vec[n];
            }
            dist = sqrt(dist);
            if (dist < distmin)
            {
              for (int n = 0; n < n_dims; ++n) vecmin[n] = vec[n];
              distmin = dist;
            }
          }
      }
      for (int n = 0; n < n_dims; ++n) wall_distance(j, i, n) = vecmin[n];
    }
}
