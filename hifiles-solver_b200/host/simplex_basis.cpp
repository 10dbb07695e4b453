// Orthonormal (Dubiner / Koornwinder) bases on the reference triangle and tetrahedron, the small dense linear algebra
// around them, and the simplex cubature tables.  They only run at set-up, but the operator matrices built from them
// are part of the numerical contract with the reference (results to 1e-12, LDG switches on rounding-level signs), so
// every routine evaluates its formula in the reference's order:
//   Jacobi polynomials / gradients   reference src/funcs.cpp:1224-1316
//   collapsed coordinates            reference src/funcs.cpp:1143-1161, 1195-1222
//   Dubiner basis 2-D / 3-D          reference src/funcs.cpp:1318-1618
//   mult / transpose / inverse       reference src/funcs.cpp:2683-2885 (full-pivot elimination, fallback dgemm order)
//   cubature tables                  reference src/cubature_tri.cpp:50-125, src/cubature_tet.cpp
#include "hifiles.h"

using namespace std;

// (n-1)! as a double
static double gamma_int(int n)
{
  double g = 1.0;
  if (n != 1)
  {
    g = n - 1;
    for (int i = 0; i < n - 2; i++) g = g * (n - 2 - i);
  }
  return g;
}

// normalised Jacobi polynomial P_mode^(alpha,beta)(r) by the three-term recurrence, lowest mode first; the values that
// the reference obtains by recursion are the same numbers because each level is the same expression of the two below
double eval_jacobi(double r, int alpha, int beta, int mode)
{
  const double w0 = pow(2.0, (-alpha - beta - 1));
  const double g_ab = gamma_int(alpha + beta + 2);
  const double g_a_b = gamma_int(alpha + 1) * gamma_int(beta + 1);
  const double p0 = sqrt(w0 * (g_ab / g_a_b));
  if (mode == 0) return p0;
  const double s1 = alpha + beta + 3;
  const double s2 = (alpha + 1) * (beta + 1);
  const double lin = (r * (alpha + beta + 2) + (alpha - beta));
  const double p1 = 0.5 * sqrt(w0 * (g_ab / g_a_b)) * sqrt(s1 / s2) * lin;
  if (mode == 1) return p1;
  double pm2 = p0, pm1 = p1, p = 0.;
  for (int n = 2; n <= mode; n++)
  {
    const double a0 = n * (n + alpha + beta) * (n + alpha) * (n + beta);
    const double a1 = ((2 * n) + alpha + beta - 1) * ((2 * n) + alpha + beta + 1);
    const double a3 = (2 * n) + alpha + beta;
    const double b0 = (n - 1) * ((n - 1) + alpha + beta) * ((n - 1) + alpha) * ((n - 1) + beta);
    const double b1 = ((2 * (n - 1)) + alpha + beta - 1) * ((2 * (n - 1)) + alpha + beta + 1);
    const double b3 = (2 * (n - 1)) + alpha + beta;
    const double c0 = -((alpha * alpha) - (beta * beta));
    const double c1 = ((2 * (n - 1)) + alpha + beta) * ((2 * (n - 1)) + alpha + beta + 2);
    const double an = (2.0 / a3) * sqrt(a0 / a1);
    const double anm1 = (2.0 / b3) * sqrt(b0 / b1);
    const double bn = c0 / c1;
    const double t0 = r * pm1;
    const double t1 = anm1 * pm2;
    const double t2 = bn * pm1;
    p = (1.0 / an) * (t0 - t1 - t2);
    pm2 = pm1;
    pm1 = p;
  }
  return p;
}

double eval_grad_jacobi(double r, int alpha, int beta, int mode)
{
  if (mode == 0) return 0.0;
  return sqrt(1.0 * mode * (mode + alpha + beta + 1)) * eval_jacobi(r, alpha + 1, beta + 1, mode - 1);
}

static void rs_to_ab(double r, double s, double &a, double &b)
{
  if (s == 1.0) a = -1.0; // collapsed vertex
  else a = (2.0 * ((1.0 + r) / (1.0 - s))) - 1.0;
  b = s;
}

// mode -> (i, j) of the 2-D basis: modes are counted by total degree k = i + j, j ascending
static void mode_ij(int mode, int order, int &i, int &j)
{
  int m = 0;
  for (int k = 0; k < order + 1; k++)
    for (int jj = 0; jj < k + 1; jj++)
    {
      if (m == mode) { i = k - jj; j = jj; return; }
      m++;
    }
  FatalError("ERROR: Invalid mode when evaluating Dubiner basis ....");
}

double eval_dubiner_basis_2d(double r, double s, int mode, int order)
{
  int i, j;
  mode_ij(mode, order, i, j);
  double a, b;
  rs_to_ab(r, s, a, b);
  const double j0 = eval_jacobi(a, 0, 0, i);
  const double j1 = eval_jacobi(b, (2 * i) + 1, 0, j);
  return sqrt(2.0) * j0 * j1 * pow(1.0 - b, i);
}

double eval_dr_dubiner_basis_2d(double r, double s, int mode, int order)
{
  int i, j;
  mode_ij(mode, order, i, j);
  double a, b;
  rs_to_ab(r, s, a, b);
  if (i == 0) return 0.;
  const double j0 = eval_grad_jacobi(a, 0, 0, i);
  const double j1 = eval_jacobi(b, (2 * i) + 1, 0, j);
  return 2.0 * sqrt(2.0) * j0 * j1 * pow(1.0 - b, i - 1);
}

double eval_ds_dubiner_basis_2d(double r, double s, int mode, int order)
{
  int i, j;
  mode_ij(mode, order, i, j);
  double a, b;
  rs_to_ab(r, s, a, b);
  const double j0 = eval_grad_jacobi(a, 0, 0, i);
  const double j1 = eval_jacobi(b, (2 * i) + 1, 0, j);
  const double j2 = eval_jacobi(a, 0, 0, i);
  const double j3 = eval_grad_jacobi(b, (2 * i) + 1, 0, j) * pow(1.0 - b, i);
  const double j4 = eval_jacobi(b, (2 * i) + 1, 0, j) * i * pow(1.0 - b, i - 1);
  if (i == 0) return sqrt(2.0) * (j2 * j3);
  return sqrt(2.0) * ((j0 * j1 * pow(1.0 - b, i - 1) * (1.0 + a)) + (j2 * (j3 - j4)));
}

// ---- 3-D ------------------------------------------------------------------------------------------------------------
static void rst_to_abc(double r, double s, double t, double &a, double &b, double &c)
{
  if (s + t == 0.0) a = -1.0;
  else a = -2.0 * (1.0 + r) / (s + t) - 1.0;
  if (t == 1.0) b = -1.0;
  else b = 2.0 * (1.0 + s) / (1.0 - t) - 1.0;
  c = t;
}

// mode -> (i, j, k): total degree m ascending, then n = j + k ascending, then k ascending
static void mode_ijk(int mode, int order, int &i, int &j, int &k)
{
  int cnt = 0;
  for (int m = 0; m < order + 1; m++)
    for (int n = 0; n < m + 1; n++)
      for (int kk = 0; kk < n + 1; kk++)
      {
        if (cnt == mode) { j = n - kk; k = kk; i = m - j - k; return; }
        cnt++;
      }
  FatalError("ERROR: Invalid mode when evaluating basis ....");
}

double eval_dubiner_basis_3d(double r, double s, double t, int mode, int order)
{
  int i, j, k;
  mode_ijk(mode, order, i, j, k);
  double a, b, c;
  rst_to_abc(r, s, t, a, b, c);
  const double j0 = eval_jacobi(a, 0, 0, i);
  const double j1 = eval_jacobi(b, (2 * i) + 1, 0, j);
  const double j2 = eval_jacobi(c, (2 * i) + (2 * j) + 2, 0, k);
  return 2.0 * sqrt(2.0) * j0 * j1 * j2 * pow(1.0 - b, i) * pow(1 - c, i + j);
}

double eval_grad_dubiner_basis_3d(double r, double s, double t, int mode, int order, int component)
{
  int i, j, k;
  mode_ijk(mode, order, i, j, k);
  double a, b, c;
  rst_to_abc(r, s, t, a, b, c);
  const double fa = eval_jacobi(a, 0, 0, i);
  const double gb = eval_jacobi(b, 2 * i + 1, 0, j);
  const double hc = eval_jacobi(c, 2 * (i + j) + 2, 0, k);
  const double dfa = eval_grad_jacobi(a, 0, 0, i);
  const double dgb = eval_grad_jacobi(b, 2 * i + 1, 0, j);
  const double dhc = eval_grad_jacobi(c, 2 * (i + j) + 2, 0, k);

  double dr = dfa * gb * hc;
  if (i > 0) dr = dr * pow(0.5 * (1. - b), i - 1);
  if (i + j > 0) dr = dr * pow(0.5 * (1. - c), i + j - 1);
  const double scale = pow(2, 2 * i + j + 1.5);
  if (component == 0) return (dr * scale);

  double ds = (0.5 * (1. + a)) * dr;
  double tmp = dgb * pow(0.5 * (1. - b), i);
  if (i > 0) tmp = tmp + (-0.5 * i) * gb * pow((0.5 * (1. - b)), i - 1);
  if (i + j > 0) tmp = tmp * pow(0.5 * (1 - c), i + j - 1);
  tmp = fa * tmp * hc;
  ds = ds + tmp;
  if (component == 1) return (ds * scale);

  double dt = 0.5 * (1. + a) * dr + 0.5 * (1. + b) * tmp;
  tmp = dhc * pow(0.5 * (1. - c), i + j);
  if (i + j > 0) tmp = tmp - 0.5 * (i + j) * (hc * pow(0.5 * (1 - c), i + j - 1));
  tmp = fa * (gb * tmp);
  tmp = tmp * pow(0.5 * (1. - b), i);
  dt = dt + tmp;
  return (dt * scale);
}

// ---- small dense linear algebra -------------------------------------------------------------------------------------------
// C = A B with the summation order of the reference's fallback dgemm: column by column, inner index ascending
hf_array<double> mult_arrays(hf_array<double> &A, hf_array<double> &B)
{
  const int m = A.get_dim(0), kk = A.get_dim(1), n = B.get_dim(1);
  if (kk != B.get_dim(0)) FatalError("ERROR: hf_array dimensions are not compatible in multiplication function");
  hf_array<double> C(m, n);
  for (int j = 0; j < n; j++)
    for (int l = 0; l < kk; l++)
    {
      const double t = 1.0 * B(l, j);
      for (int i = 0; i < m; i++) C(i, j) += t * A(i, l);
    }
  return C;
}

hf_array<double> transpose_array(hf_array<double> &A)
{
  const int n = A.get_dim(0);
  if (A.get_dim(1) != n) FatalError("ERROR: Array transpose function only accepts a 2-dimensional square hf_array");
  hf_array<double> T(n, n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) T(j, i) = A(i, j);
  return T;
}

// inverse by Gaussian elimination with full pivoting (pivot = largest squared magnitude, first found wins), right-hand
// side = identity carried along, back substitution, rows put back by the column permutation
hf_array<double> inv_array(hf_array<double> &in)
{
  const int n = in.get_dim(0);
  if (in.get_dim(1) != n) FatalError("ERROR: Can only obtain inverse of a square hf_array");
  hf_array<double> w(n, n), rhs(n, n), x(n, n), out(n, n), tmp(n);
  vector<int> colperm(n);
  for (int i = 0; i < n; i++)
  {
    colperm[i] = i;
    for (int j = 0; j < n; j++) w(i, j) = in(i, j);
    rhs(i, i) = 1.0;
  }
  for (int k = 0; k < n - 1; k++)
  {
    double best = 0;
    int pi = k, pj = k;
    for (int i = k; i < n; i++)
      for (int j = k; j < n; j++)
      {
        const double mag = w(i, j) * w(i, j);
        if (mag > best) { pi = i; pj = j; best = mag; }
      }
    swap(colperm[k], colperm[pj]);
    for (int i = 0; i < n; i++) { tmp(i) = w(i, pj); w(i, pj) = w(i, k); w(i, k) = tmp(i); }
    for (int j = 0; j < n; j++)
    {
      tmp(j) = w(pi, j); w(pi, j) = w(k, j); w(k, j) = tmp(j);
      tmp(j) = rhs(pi, j); rhs(pi, j) = rhs(k, j); rhs(k, j) = tmp(j);
    }
    for (int i = k + 1; i < n; i++)
    {
      const double first = w(i, k);
      for (int j = 0; j < n; j++)
      {
        if (j >= k) w(i, j) = w(i, j) - ((first / w(k, k)) * w(k, j));
        rhs(i, j) = rhs(i, j) - ((first / w(k, k)) * rhs(k, j));
      }
    }
    for (int j = 0; j < k + 1; j++)
      for (int i = j + 1; i < n; i++) w(i, j) = 0.0;
  }
  for (int i = n - 1; i >= 0; i--)
    for (int j = 0; j < n; j++)
    {
      double acc = 0.0;
      for (int k = i + 1; k < n; k++) acc = acc + (w(i, k) * x(k, j));
      x(i, j) = (rhs(i, j) - acc) / w(i, i);
    }
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) out(colperm[i], j) = x(i, j);
  return out;
}

// ---- cubature tables ------------------------------------------------------------------------------------------------------------
// Binary tables of the run's data directory: for each order, all r, then all s (then t), then the weights (interior
// rules only).  rule 0 = interior (Gauss-type, with weights), rule 1 = alpha-optimised points (no weights).
static void read_simplex_table(const string &file, int n_coord, int lower, int upper, bool weights, int order, hf_array<double> &locs,
                               hf_array<double> &w, int n_pts, long (*npts_of)(int))
{
  if (order > upper || order < lower) FatalError("cubature order not implemented.");
  FILE *f = fopen(file.c_str(), "rb");
  if (!f) FatalError("Unable to open cubature file " + file);
  long skip = 0;
  for (int i = lower; i < order; i++) skip += (n_coord + (weights ? 1 : 0)) * npts_of(i);
  fseek(f, (long)sizeof(double) * skip, SEEK_SET);
  locs.setup(n_pts, n_coord);
  w.setup(n_pts);
  vector<double> buf(n_pts);
  for (int c = 0; c < n_coord; c++)
  {
    if ((int)fread(buf.data(), sizeof(double), n_pts, f) != n_pts) FatalError("cubature file truncated");
    for (int i = 0; i < n_pts; i++) locs(i, c) = buf[i];
  }
  if (weights && (int)fread(w.get_ptr_cpu(), sizeof(double), n_pts, f) != n_pts) FatalError("cubature file truncated");
  fclose(f);
}
static long npts_tri(int p) { return (long)(p + 1) * (p + 2) / 2; }
static long npts_tet(int p) { return (long)(p + 1) * (p + 2) * (p + 3) / 6; }

void cubature_tri(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights)
{
  const int n = (int)npts_tri(in_order);
  if (in_rule == 0) read_simplex_table(hifiles_data_dir() + "/tri_inter.bin", 2, 0, 7, true, in_order, locs, weights, n, npts_tri);
  else if (in_rule == 1) read_simplex_table(hifiles_data_dir() + "/tri_alpha.bin", 2, 1, 15, false, in_order, locs, weights, n, npts_tri);
  else FatalError("Cubature rule not implemented");
}

void cubature_tet(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights)
{
  const int n = (int)npts_tet(in_order);
  if (in_rule == 0) read_simplex_table(hifiles_data_dir() + "/tet_inter.bin", 3, 0, 6, true, in_order, locs, weights, n, npts_tet);
  else if (in_rule == 1) read_simplex_table(hifiles_data_dir() + "/tet_alpha.bin", 3, 1, 15, false, in_order, locs, weights, n, npts_tet);
  else FatalError("Cubature rule not implemented");
}
