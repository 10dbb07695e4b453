// LES test filters (reference <type>::compute_filter_upts: src/eles_hexas.cpp:583-793, src/eles_quads.cpp:428-630,
// src/eles_tris.cpp:786-970, src/eles_tets.cpp:576-703; compute_modal_filter_1d src/funcs.cpp:669-715; gaussj
// src/funcs.cpp:2580-2650).  The filter matrix acts on the nodal values of an element: filtered = filter_upts * u.
// It is used by the filter-based sub-grid models (WALE-similarity, SVV, similarity) in eles::calc_sgs_terms.
#include "hifiles.h"

using namespace std;

namespace
{
// Gauss-Jordan elimination with full pivoting (Numerical Recipes), in place: A -> A^-1, b -> solution
void gauss_jordan(int n, hf_array<double> &A, hf_array<double> &b)
{
  vector<int> indxc(n, 0), indxr(n, 0), ipiv(n, 0);
  int icol = 0, irow = 0;
  for (int i = 0; i < n; i++)
  {
    double big = 0.0;
    for (int j = 0; j < n; j++)
      if (ipiv[j] != 1)
        for (int k = 0; k < n; k++)
          if (ipiv[k] == 0 && fabs(A(k, j)) >= big) { big = fabs(A(k, j)); irow = k; icol = j; }
    ipiv[icol] = ipiv[icol] + 1;
    if (irow != icol)
    {
      for (int l = 0; l < n; l++) swap(A(l, irow), A(l, icol));
      swap(b(irow), b(icol));
    }
    indxr[i] = irow;
    indxc[i] = icol;
    if (A(icol, icol) == 0.0) FatalError("Error: Singular matrix in gaussj");
    const double pivinv = 1.0 / A(icol, icol);
    A(icol, icol) = 1.0;
    for (int l = 0; l < n; l++) A(l, icol) = A(l, icol) * pivinv;
    b(icol) = b(icol) * pivinv;
    for (int ll = 0; ll < n; ll++)
      if (ll != icol)
      {
        const double dum = A(icol, ll);
        A(icol, ll) = 0.0;
        for (int l = 0; l < n; l++) A(l, ll) = A(l, ll) - A(l, icol) * dum;
        b(ll) = b(ll) - b(icol) * dum;
      }
  }
  for (int l = n - 1; l >= 0; l--)
    if (indxr[l] != indxc[l])
      for (int k = 0; k < n; k++) swap(A(indxr[l], k), A(indxc[l], k));
}

// 1-D filter on the solution points of a line (hexahedra and quads build theirs as tensor products of it)
hf_array<double> filter_1d(int order, hf_array<double> &X)
{
  const int N = order + 1;
  int N2 = N / 2;
  if (N % 2 != 0) N2 += 1;
  const double k_c = 1.0 / run_input.filter_ratio;
  const double dlt = 2.0 / order;
  hf_array<double> F(N, N), beta(N, N), B(N);
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) beta(j, i) = (X(j) - X(i)) / dlt;
  if (run_input.filter_type == 0 && N >= 3)
  {
    // Vasilyev's high-order commuting filter: moment conditions solved per point
    hf_array<double> A(N, N);
    for (int i = 0; i < N; ++i)
    {
      const bool centre = N % 2 == 1 && i + 1 == N2;
      B(0) = 1.0;
      B(1) = exp(-pow(pi, 2) / 24.0);
      B(2) = -B(1) * pow(pi, 2) / k_c / 12.0;
      if (centre) B(2) = 0.0;
      for (int j = 0; j < N; ++j)
      {
        A(j, 0) = 1.0;
        A(j, 1) = cos(pi * k_c * beta(j, i));
        A(j, 2) = -beta(j, i) * pi * sin(pi * k_c * beta(j, i));
        if (centre) A(j, 2) = pow(beta(j, i), 3);
      }
      for (int k = 3; k < N; ++k)
      {
        for (int j = 0; j < N; ++j) A(j, k) = pow(beta(j, i), k + 1);
        B(k) = 0.0;
      }
      gauss_jordan(N, A, B);
      for (int j = 0; j < N; ++j) F(j, i) = B(j);
    }
  }
  else if (run_input.filter_type == 1)
  {
    // discrete Gaussian, quadrature-weighted and normalised row by row
    hf_array<double> r, wf;
    cubature_1d(0, order, r, wf);
    for (int i = 0; i < N; ++i)
    {
      double norm = 0.0;
      for (int j = 0; j < N; ++j)
      {
        F(i, j) = wf(j) * exp(-6.0 * pow(k_c * beta(i, j), 2));
        norm += F(i, j);
      }
      for (int j = 0; j < N; ++j) F(i, j) /= norm;
    }
  }
  else if (run_input.filter_type == 2)
  {
    // Gaussian in Legendre-modal space
    hf_array<double> V(N, N), modal(N, N);
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) V(i, j) = eval_legendre(X(i), j);
    hf_array<double> Vi = inv_array(V);
    for (int i = 0; i < N; i++)
    {
      const double eta = i / double(N);
      modal(i, i) = exp(-pow(2.0 * eta, 2.0) / 48.0);
    }
    hf_array<double> t = mult_arrays(V, modal);
    F = mult_arrays(t, Vi);
  }
  else
  {
    for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) F(i, j) = 1.0 / N;
  }
  return F;
}

// symmetrise about the centre of the point numbering and renormalise the rows (triangles and tetrahedra)
void symmetrise(hf_array<double> &F, int N)
{
  int N2 = N / 2;
  if (N % 2 != 0) N2 += 1;
  for (int i = 0; i < N2; i++)
    for (int j = 0; j < N; j++)
    {
      F(i, j) = 0.5 * F(i, j) + F(N - i - 1, N - j - 1);
      F(N - i - 1, N - j - 1) = F(i, j);
    }
  for (int i = 0; i < N2; i++)
  {
    double norm = 0.0;
    for (int j = 0; j < N; j++) norm += F(i, j);
    for (int j = 0; j < N; j++) F(i, j) /= norm;
    for (int j = 0; j < N; j++) F(N - i - 1, N - j - 1) = F(i, j);
  }
}
} // namespace

void eles::compute_filter_upts()
{
  const int n = n_upts_per_ele;
  filter_upts.setup(n, n);
  if (ele_type == HEX || ele_type == QUAD)
  {
    const int N = order + 1;
    hf_array<double> F1 = filter_1d(order, loc_1d_upts);
    if (ele_type == HEX)
    {
      int ii = 0;
      for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j)
          for (int k = 0; k < N; ++k)
          {
            int jj = 0;
            for (int l = 0; l < N; ++l)
              for (int m = 0; m < N; ++m)
                for (int q = 0; q < N; ++q) filter_upts(ii, jj++) = F1(k, q) * F1(j, m) * F1(i, l);
            ++ii;
          }
    }
    else
    {
      int ii = 0;
      for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
        {
          int jj = 0;
          for (int k = 0; k < N; k++)
            for (int l = 0; l < N; l++) filter_upts(ii, jj++) = F1(j, l) * F1(i, k);
          ++ii;
        }
    }
    return;
  }
  if (ele_type == PRISM) FatalError("the reference builds no LES filter for prisms (src/eles_pris.cpp:134): filter-based SGS models are not available on them");
  // triangles / tetrahedra
  const double k_c = 1.0 / run_input.filter_ratio;
  const double dlt = 2.0 / order;
  if (run_input.filter_type == 0) FatalError("Vasilyev filters not implemented for tris. Exiting.");
  if (run_input.filter_type == 1)
  {
    if (ele_type == TET) FatalError("Gaussian filter not implemented for tris. Exiting.");
    hf_array<double> beta(n, n), locs, wf;
    for (int i = 0; i < n; i++)
      for (int j = i; j < n; j++) beta(i, j) = sqrt(pow(loc_upts(0, i) - loc_upts(0, j), 2) + pow(loc_upts(1, i) - loc_upts(1, j), 2)) / dlt;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < i; j++) beta(i, j) = beta(j, i);
    cubature_tri(0, order, locs, wf); // weights of the interior rule of the same size
    for (int i = 0; i < n; i++)
    {
      double norm = 0.0;
      for (int j = 0; j < n; j++)
      {
        filter_upts(i, j) = wf(j) * exp(-6.0 * pow(k_c * beta(i, j), 2));
        norm += filter_upts(i, j);
      }
      for (int j = 0; j < n; j++) filter_upts(i, j) /= norm;
    }
  }
  else if (run_input.filter_type == 2)
  {
    set_modal_vandermonde(); // Dubiner Vandermonde matrix of the solution points
    hf_array<double> diag(n, n);
    for (int i = 0; i < n; i++)
    {
      const double eta = i / double(n);
      diag(i, i) = exp(-pow(2.0 * eta, 2.0) / 48.0);
    }
    hf_array<double> t = mult_arrays(modal_vandermonde, diag);
    filter_upts = mult_arrays(t, modal_inv_vandermonde);
  }
  else
  {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) filter_upts(i, j) = 1.0 / n;
  }
  symmetrise(filter_upts, n);
}
