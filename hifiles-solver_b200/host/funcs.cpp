// 1-D basis evaluators used to build the element operators.  The operator matrices are part of the numerical
// contract with the reference (results must agree to 1e-12, and exact 0/1 entries of the tensor-product
// operators are what makes sum-factorisation bit-compatible), so each evaluator keeps the reference's
// evaluation order (reference src/funcs.cpp:316-509, 1619-1672, 1724-1739).
#include "hifiles.h"
#include <cstdlib>
#include <cstdio>

using namespace std;

std::string hifiles_data_dir()
{
  const char *h = getenv("HIFILES_HOME");
  if (h && *h) return string(h) + "/data";
  const char *d = getenv("HIFILES_B200_DATA");
  if (d && *d) return string(d);
#ifdef HIFILES_B200_DATA_DIR
  return string(HIFILES_B200_DATA_DIR);
#else
  return string("data");
#endif
}

void cubature_1d(int in_rule, int in_order, hf_array<double> &locs, hf_array<double> &weights)
{
  // table layout: for order p, (p+1) locations then (p+1) weights, stored after the p(p+1) doubles of the
  // lower orders (reference src/cubature_1d.cpp:50-85)
  string filename = hifiles_data_dir();
  if (in_rule == 0) filename += "/JacobiGQ.bin";
  else if (in_rule == 1) filename += "/JacobiGL.bin";
  else FatalError("cubature rule not implemented.");
  if (in_order < 0 || in_order > 15) FatalError("cubature order not implemented.");
  FILE *f = fopen(filename.c_str(), "rb");
  if (!f) FatalError("Unable to open cubature file " + filename);
  int n = in_order + 1;
  locs.setup(n);
  weights.setup(n);
  fseek(f, (long)sizeof(double) * (1 + in_order) * in_order, SEEK_SET);
  size_t a = fread(locs.get_ptr_cpu(), sizeof(double), n, f);
  size_t b = fread(weights.get_ptr_cpu(), sizeof(double), n, f);
  fclose(f);
  if ((int)a != n || (int)b != n) FatalError("cubature file truncated");
}

double eval_lagrange(double in_r, int in_mode, hf_array<double> &in_loc_pts)
{
  double v = 1.0;
  int n = in_loc_pts.get_dim(0);
  for (int i = 0; i < n; i++)
    if (i != in_mode)
      v = v * ((in_r - in_loc_pts(i)) / (in_loc_pts(in_mode) - in_loc_pts(i)));
  return v;
}

double eval_d_lagrange(double in_r, int in_mode, hf_array<double> &in_loc_pts)
{
  int n = in_loc_pts.get_dim(0);
  double sum = 0.0;
  for (int i = 0; i < n; i++)
  {
    if (i == in_mode) continue;
    double num = 1.0, den = 1.0;
    for (int j = 0; j < n; j++)
    {
      if (j != in_mode && j != i) num = num * (in_r - in_loc_pts(j));
      if (j != in_mode) den = den * (in_loc_pts(in_mode) - in_loc_pts(j));
    }
    sum = sum + (num / den);
  }
  return sum;
}

double eval_legendre(double in_r, int in_mode)
{
  if (in_mode == 0) return 1.0;
  if (in_mode == 1) return in_r;
  return ((2 * in_mode - 1) * in_r * eval_legendre(in_r, in_mode - 1) - (in_mode - 1) * eval_legendre(in_r, in_mode - 2)) / in_mode;
}

double eval_d_legendre(double in_r, int in_mode)
{
  double d = 0.;
  if (in_mode == 0) return 0.;
  if (in_r > -1.0 && in_r < 1.0)
    d = (in_mode * ((in_r * eval_legendre(in_r, in_mode)) - eval_legendre(in_r, in_mode - 1))) / ((in_r * in_r) - 1.0);
  else
  {
    if (in_r == -1.0) d = pow(-1.0, in_mode - 1.0) * 0.5 * in_mode * (in_mode + 1.0);
    if (in_r == 1.0) d = 0.5 * in_mode * (in_mode + 1.0);
  }
  return d;
}

double eval_d_vcjh_1d(double in_r, int in_mode, int in_order, double in_eta)
{
  double v = 0.;
  if (in_mode == 0) // left correction function
  {
    if (in_order == 0)
      v = 0.5 * pow(-1.0, in_order) * (eval_d_legendre(in_r, in_order) - ((eval_d_legendre(in_r, in_order + 1)) / (1.0 + in_eta)));
    else
      v = 0.5 * pow(-1.0, in_order) * (eval_d_legendre(in_r, in_order) - (((in_eta * eval_d_legendre(in_r, in_order - 1)) + eval_d_legendre(in_r, in_order + 1)) / (1.0 + in_eta)));
  }
  else if (in_mode == 1) // right correction function
  {
    if (in_order == 0)
      v = 0.5 * (eval_d_legendre(in_r, in_order) + ((eval_d_legendre(in_r, in_order + 1)) / (1.0 + in_eta)));
    else
      v = 0.5 * (eval_d_legendre(in_r, in_order) + (((in_eta * eval_d_legendre(in_r, in_order - 1)) + eval_d_legendre(in_r, in_order + 1)) / (1.0 + in_eta)));
  }
  return v;
}

static int factorial(int n)
{
  int r = 1;
  for (int i = 1; i <= n; i++) r *= i;
  return r;
}

double compute_eta(int vcjh_scheme, int order)
{
  double eta = 0.;
  if (order == 0 && vcjh_scheme != 1)
    FatalError("ERROR: P=0 only compatible with DG. Set VCJH scheme type to 1!");
  if (vcjh_scheme == 1) eta = 0.0;
  else if (vcjh_scheme == 2) eta = (1.0 * (order)) / (1.0 * (order + 1));
  else if (vcjh_scheme == 3) eta = (1.0 * (order + 1)) / (1.0 * order);
  else if (vcjh_scheme == 4)
  {
    double c_1d;
    if (order == 2) c_1d = 0.206;
    else if (order == 3) c_1d = 3.80e-3;
    else if (order == 4) c_1d = 4.67e-5;
    else if (order == 5) c_1d = 4.28e-7;
    else { FatalError("C_plus scheme not implemented for this order"); c_1d = 0.; }
    double ap = 1. / pow(2.0, order) * factorial(2 * order) / (factorial(order) * factorial(order));
    eta = c_1d * (2 * order + 1) / 2 * (factorial(order) * ap) * (factorial(order) * ap);
  }
  else
    FatalError("ERROR: Invalid VCJH scheme ... ");
  return eta;
}

bool is_perfect_square(int in_a)
{
  int number = (int)round(sqrt(1.0 * in_a));
  return (in_a == number * number);
}

bool is_perfect_cube(int in_a)
{
  int number = (int)round(pow(1.0 * in_a, 1. / 3.));
  return (in_a == number * number * number);
}

// Exact solutions of the scalar advection-diffusion test equation, used as initial data (ic_form 2, 3, 4) and by
// compute_error: a plane sine wave along (1,1[,1]), a product of sines, a Gaussian pulse (reference src/funcs.cpp:1742-1808).
// The decay factor is written as the reference has it, exp(-n_dims * diff_coeff * pi^2 * t), so time 0 gives the same bits.
static inline double ad_decay(double diff_coeff, double time, int n_dims) { return exp(-((double)n_dims) * diff_coeff * pi * pi * time); }

void eval_sine_wave_single(hf_array<double> &pos, hf_array<double> &wave_speed, double diff_coeff, double time, double &rho, hf_array<double> &grad_rho, int n_dims)
{
  double x = pos(0) - wave_speed(0) * time, y = pos(1) - wave_speed(1) * time;
  double angle = x + y;
  if (n_dims == 3) angle = x + y + (pos(2) - wave_speed(2) * time);
  rho = ad_decay(diff_coeff, time, n_dims) * sin(pi * angle);
  for (int d = 0; d < n_dims; d++) grad_rho(d) = pi * ad_decay(diff_coeff, time, n_dims) * cos(pi * angle);
}

void eval_sine_wave_group(hf_array<double> &pos, hf_array<double> &wave_speed, double diff_coeff, double time, double &rho, hf_array<double> &grad_rho, int n_dims)
{
  double r[3] = {0., 0., 0.};
  for (int d = 0; d < n_dims; d++) r[d] = pos(d) - wave_speed(d) * time;
  const double a = ad_decay(diff_coeff, time, n_dims);
  if (n_dims == 2)
  {
    rho = a * sin(pi * r[0]) * sin(pi * r[1]);
    grad_rho(0) = pi * a * cos(pi * r[0]) * sin(pi * r[1]);
    grad_rho(1) = pi * a * sin(pi * r[0]) * cos(pi * r[1]);
  }
  else
  {
    rho = a * sin(pi * r[0]) * sin(pi * r[1]) * sin(pi * r[2]);
    grad_rho(0) = pi * a * cos(pi * r[0]) * sin(pi * r[1]) * sin(pi * r[2]);
    grad_rho(1) = pi * a * sin(pi * r[0]) * cos(pi * r[1]) * sin(pi * r[2]);
    grad_rho(2) = pi * a * sin(pi * r[0]) * sin(pi * r[1]) * cos(pi * r[2]);
  }
}

// the reference reads three coordinates whatever n_dims is (src/funcs.cpp:1797-1808): the pulse is a 3-D initial condition
void eval_sphere_wave(hf_array<double> &pos, hf_array<double> &wave_speed, double time, double &rho, int n_dims)
{
  if (n_dims != 3) FatalError("ic_form 4 (spherical pulse) needs a three-dimensional mesh");
  double x = pos(0) - wave_speed(0) * time, y = pos(1) - wave_speed(1) * time, z = pos(2) - wave_speed(2) * time;
  rho = exp(-0.5 * (x * x + y * y + z * z));
}

void eval_isentropic_vortex(hf_array<double> &pos, double time, double &rho, double &vx, double &vy, double &vz, double &p, int n_dims)
{
  (void)n_dims;
  double ev_eps_ic = 5.0;
  double gamma = run_input.gamma;
  double x = pos(0) - time;
  double y = pos(1) - time;
  double f = 1.0 - (x * x + y * y);
  rho = pow(1.0 - ev_eps_ic * ev_eps_ic * (gamma - 1.0) / (8.0 * gamma * pi * pi) * exp(f), 1.0 / (gamma - 1.0));
  vx = 1. - ev_eps_ic * y / (2.0 * pi) * exp(f / 2.0);
  vy = 1. + ev_eps_ic * x / (2.0 * pi) * exp(f / 2.0);
  vz = 0.;
  p = pow(rho, gamma);
}
