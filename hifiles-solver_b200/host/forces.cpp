// Surface forces on solid walls (reference output::CalcForces, src/output.cpp:1915-2012; eles::compute_wall_forces,
// src/eles.cpp:5704-5990; interface cubature: <type>::set_inters_cubpts, eles::set_opp_inters_cubpts,
// eles::set_transforms_inters_cubpts, <type>::compute_inter_detjac_inters_cubpts).  Runs on the host after a
// device -> host copy of the solution and of grad_disu_upts, on monitored steps only, over the boundary elements:
// pressure and viscous traction at the cubature points of every wall face, summed into force, lift and drag
// coefficients; optionally the cp / cf distribution into force_files_<iter>/cp_<iter>_p<rank>.dat.
#include "hifiles.h"
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iomanip>
#include <vector>
#include <sys/stat.h>
#include <dirent.h>
using namespace std;

namespace
{
// geometry of one local face of a reference element: where cubature point (u, v) sits, the transformed normal, and the
// two tangents as combinations of the columns (r, s, t) of d_pos, a = first tangent, b = second (2-D: a only)
struct face_rule
{
  int kind;          // 0 line, 1 triangle, 2 quadrilateral cubature
  double c0[3], cu[3], cv[3];
  bool reversed;     // line rules: points taken in descending order
  double tn[3];
  double a[3], b[3]; // coefficients of the r, s, t columns
};

const double q2 = 0.70710678118654752440, q3 = 0.57735026918962576451; // replaced by 1/sqrt(2), 1/sqrt(3) at run time

void rules_of(int type, vector<face_rule> &R)
{
  const double i2 = 1. / sqrt(2.), i3 = 1. / sqrt(3.);
  (void)q2; (void)q3;
  auto F = [&](int kind, initializer_list<double> c0, initializer_list<double> cu, initializer_list<double> cv, bool rev, initializer_list<double> tn,
               initializer_list<double> a, initializer_list<double> b) {
    face_rule f;
    f.kind = kind; f.reversed = rev;
    copy(c0.begin(), c0.end(), f.c0); copy(cu.begin(), cu.end(), f.cu); copy(cv.begin(), cv.end(), f.cv);
    copy(tn.begin(), tn.end(), f.tn); copy(a.begin(), a.end(), f.a); copy(b.begin(), b.end(), f.b);
    R.push_back(f);
  };
  R.clear();
  if (type == QUAD)
  {
    F(0, {0, -1, 0}, {1, 0, 0}, {0, 0, 0}, false, {0, -1, 0}, {1, 0, 0}, {0, 0, 0});
    F(0, {1, 0, 0}, {0, 1, 0}, {0, 0, 0}, false, {1, 0, 0}, {0, 1, 0}, {0, 0, 0});
    F(0, {0, 1, 0}, {1, 0, 0}, {0, 0, 0}, true, {0, 1, 0}, {1, 0, 0}, {0, 0, 0});
    F(0, {-1, 0, 0}, {0, 1, 0}, {0, 0, 0}, true, {-1, 0, 0}, {0, 1, 0}, {0, 0, 0});
  }
  else if (type == TRI)
  {
    F(0, {0, -1, 0}, {1, 0, 0}, {0, 0, 0}, false, {0, -1, 0}, {1, 0, 0}, {0, 0, 0});
    F(0, {0, 0, 0}, {0, 1, 0}, {0, 0, 0}, false, {i2, i2, 0}, {1, -1, 0}, {0, 0, 0}); // x takes the mirrored point, see loc_of
    F(0, {-1, 0, 0}, {0, 1, 0}, {0, 0, 0}, true, {-1, 0, 0}, {0, 1, 0}, {0, 0, 0});
  }
  else if (type == HEX)
  {
    F(2, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, false, {0, 0, -1}, {1, 0, 0}, {0, 1, 0});
    F(2, {0, -1, 0}, {1, 0, 0}, {0, 0, 1}, false, {0, -1, 0}, {1, 0, 0}, {0, 0, 1});
    F(2, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, false, {1, 0, 0}, {0, 1, 0}, {0, 0, 1});
    F(2, {0, 1, 0}, {1, 0, 0}, {0, 0, 1}, false, {0, 1, 0}, {1, 0, 0}, {0, 0, 1});
    F(2, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1}, false, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1});
    F(2, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}, false, {0, 0, 1}, {1, 0, 0}, {0, 1, 0});
  }
  else if (type == TET)
  {
    F(1, {0, 0, -1}, {1, 0, -1}, {0, 1, -1}, false, {i3, i3, i3}, {1, 0, -1}, {0, 1, -1});
    // the reference takes the (r, t) tangents on the x = -1 face as well (src/eles_tets.cpp, compute_inter_detjac_inters_cubpts): kept
    F(1, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1}, false, {-1, 0, 0}, {1, 0, 0}, {0, 0, 1});
    F(1, {0, -1, 0}, {1, 0, 0}, {0, 0, 1}, false, {0, -1, 0}, {1, 0, 0}, {0, 0, 1});
    F(1, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, false, {0, 0, -1}, {1, 0, 0}, {0, 1, 0});
  }
  else
  {
    F(1, {0, 0, -1}, {1, 0, 0}, {0, 1, 0}, false, {0, 0, -1}, {1, 0, 0}, {0, 1, 0});
    F(1, {0, 0, 1}, {1, 0, 0}, {0, 1, 0}, false, {0, 0, 1}, {1, 0, 0}, {0, 1, 0});
    F(2, {0, -1, 0}, {1, 0, 0}, {0, 0, 1}, false, {0, -1, 0}, {1, 0, 0}, {0, 0, 1});
    F(2, {0, 0, 0}, {1, -1, 0}, {0, 0, 1}, false, {i2, i2, 0}, {1, -1, 0}, {0, 0, 1});
    F(2, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1}, false, {-1, 0, 0}, {0, 1, 0}, {0, 0, 1});
  }
}
} // namespace

// interface cubature of the element type, interpolation to it, and normals / area elements of the boundary elements
void eles::set_inters_cubpts_and_transforms()
{
  vector<face_rule> R;
  rules_of(ele_type, R);
  hf_array<double> r1, w1, tri, wt, quad, wq;
  cubature_1d(0, order, r1, w1);
  const int n1 = order + 1;
  if (ele_type == TET || ele_type == PRISM) cubature_tri(0, order, tri, wt);
  if (ele_type == HEX || ele_type == PRISM)
  {
    quad.setup(n1 * n1, 2);
    wq.setup(n1 * n1);
    for (int i = 0; i < n1; i++)
      for (int j = 0; j < n1; j++)
      {
        quad(j + n1 * i, 0) = r1(j);
        quad(j + n1 * i, 1) = r1(i);
        wq(j + n1 * i) = w1(j) * w1(i);
      }
  }
  const int nf = n_inters_per_ele;
  n_cubpts_per_inter.setup(nf);
  loc_inters_cubpts.assign(nf, hf_array<double>());
  weight_inters_cubpts.assign(nf, hf_array<double>());
  opp_inters_cubpts.assign(nf, hf_array<double>());
  for (int l = 0; l < nf; l++)
  {
    const face_rule &f = R[l];
    const int n = f.kind == 0 ? n1 : (f.kind == 1 ? tri.get_dim(0) : n1 * n1);
    n_cubpts_per_inter(l) = n;
    loc_inters_cubpts[l].setup(n_dims, n);
    weight_inters_cubpts[l].setup(n);
    for (int j = 0; j < n; j++)
    {
      double u, v = 0.;
      if (f.kind == 0) u = f.reversed ? r1(n1 - j - 1) : r1(j);
      else if (f.kind == 1) { u = tri(j, 0); v = tri(j, 1); }
      else { u = quad(j, 0); v = quad(j, 1); }
      for (int d = 0; d < n_dims; d++) loc_inters_cubpts[l](d, j) = f.c0[d] + f.cu[d] * u + f.cv[d] * v;
      // the hypotenuse of the triangle: x runs through the points in descending order while y ascends
      if (ele_type == TRI && l == 1) loc_inters_cubpts[l](0, j) = r1(n1 - j - 1);
      weight_inters_cubpts[l](j) = f.kind == 0 ? w1(j) : (f.kind == 1 ? wt(j) : wq(j));
    }
    hf_array<double> loc(n_dims);
    opp_inters_cubpts[l].setup(n, n_upts_per_ele);
    for (int i = 0; i < n_upts_per_ele; i++)
      for (int j = 0; j < n; j++)
      {
        for (int d = 0; d < n_dims; d++) loc(d) = loc_inters_cubpts[l](d, j);
        opp_inters_cubpts[l](j, i) = eval_nodal_basis(i, loc);
      }
  }
  // boundary elements: every element with a face that carries a boundary id (src/eles.cpp:4396-4433)
  bdy_ele2ele.clear();
  for (int i = 0; i < n_eles; i++)
    for (int j = 0; j < nf; j++)
      if (bcid(i, j) != -1) { bdy_ele2ele.push_back(i); break; }
  const int nb = (int)bdy_ele2ele.size();
  inter_detjac_inters_cubpts.assign(nf, hf_array<double>());
  norm_inters_cubpts.assign(nf, hf_array<double>());
  hf_array<double> loc(n_dims), d_pos(n_dims, n_dims);
  for (int l = 0; l < nf; l++)
  {
    inter_detjac_inters_cubpts[l].setup(n_cubpts_per_inter(l), nb > 0 ? nb : 1);
    norm_inters_cubpts[l].setup(n_cubpts_per_inter(l), nb > 0 ? nb : 1, n_dims);
  }
  for (int i = 0; i < nb; i++)
    for (int l = 0; l < nf; l++)
    {
      const face_rule &f = R[l];
      for (int j = 0; j < n_cubpts_per_inter(l); j++)
      {
        for (int k = 0; k < n_dims; k++) loc(k) = loc_inters_cubpts[l](k, j);
        calc_d_pos(loc, bdy_ele2ele[i], d_pos);
        double t[3] = {0, 0, 0};
        if (n_dims == 2)
        {
          // transformed normal times adj(J)
          t[0] = (f.tn[0] * d_pos(1, 1)) - (f.tn[1] * d_pos(1, 0));
          t[1] = -(f.tn[0] * d_pos(0, 1)) + (f.tn[1] * d_pos(0, 0));
          const double mag = sqrt(t[0] * t[0] + t[1] * t[1]);
          norm_inters_cubpts[l](j, i, 0) = t[0] / mag;
          norm_inters_cubpts[l](j, i, 1) = t[1] / mag;
          // length element of the edge: |a_r column_r + a_s column_s|
          double e0, e1;
          if (f.a[0] != 0. && f.a[1] != 0.) { e0 = d_pos(0, 0) - d_pos(0, 1); e1 = d_pos(1, 0) - d_pos(1, 1); }
          else if (f.a[0] != 0.) { e0 = d_pos(0, 0); e1 = d_pos(1, 0); }
          else { e0 = d_pos(0, 1); e1 = d_pos(1, 1); }
          inter_detjac_inters_cubpts[l](j, i) = sqrt(e0 * e0 + e1 * e1);
        }
        else
        {
          t[0] = ((f.tn[0] * (d_pos(1, 1) * d_pos(2, 2) - d_pos(1, 2) * d_pos(2, 1))) + (f.tn[1] * (d_pos(1, 2) * d_pos(2, 0) - d_pos(1, 0) * d_pos(2, 2))) +
                  (f.tn[2] * (d_pos(1, 0) * d_pos(2, 1) - d_pos(1, 1) * d_pos(2, 0))));
          t[1] = ((f.tn[0] * (d_pos(0, 2) * d_pos(2, 1) - d_pos(0, 1) * d_pos(2, 2))) + (f.tn[1] * (d_pos(0, 0) * d_pos(2, 2) - d_pos(0, 2) * d_pos(2, 0))) +
                  (f.tn[2] * (d_pos(0, 1) * d_pos(2, 0) - d_pos(0, 0) * d_pos(2, 1))));
          t[2] = ((f.tn[0] * (d_pos(0, 1) * d_pos(1, 2) - d_pos(0, 2) * d_pos(1, 1))) + (f.tn[1] * (d_pos(0, 2) * d_pos(1, 0) - d_pos(0, 0) * d_pos(1, 2))) +
                  (f.tn[2] * (d_pos(0, 0) * d_pos(1, 1) - d_pos(0, 1) * d_pos(1, 0))));
          const double mag = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
          for (int m = 0; m < 3; m++) norm_inters_cubpts[l](j, i, m) = t[m] / mag;
          // area element: | tangent_u x tangent_v |, the tangents being differences of columns of d_pos
          auto tangent = [&](const double *c, double *o) {
            for (int m = 0; m < 3; m++)
            {
              // at most two columns take part, with coefficients +1 / -1, the positive one first (xr - xt, xr - xs)
              double val = 0.;
              bool first = true;
              for (int col = 0; col < 3; col++)
                if (c[col] > 0.) { val = d_pos(m, col); first = false; }
              for (int col = 0; col < 3; col++)
                if (c[col] < 0.) val = first ? -d_pos(m, col) : val - d_pos(m, col);
              o[m] = val;
            }
          };
          double a[3], b[3];
          tangent(f.a, a);
          tangent(f.b, b);
          const double temp0 = a[1] * b[2] - a[2] * b[1], temp1 = a[2] * b[0] - a[0] * b[2], temp2 = a[0] * b[1] - a[1] * b[0];
          inter_detjac_inters_cubpts[l](j, i) = sqrt(temp0 * temp0 + temp1 * temp1 + temp2 * temp2);
        }
      }
    }
}

void eles::compute_wall_forces(hf_array<double> &inv_force, hf_array<double> &vis_force, double &temp_cl, double &temp_cd, std::ofstream &coeff_file,
                               bool write_forces)
{
  hf_array<double> u_l(n_fields), norm(n_dims), grad_u_l(n_fields, n_dims), dv(n_dims, n_dims), dE(n_dims), drho(n_dims), taun(n_dims), tautan(n_dims);
  hf_array<double> Finv(n_dims), Fvis(n_dims), loc(n_dims), pos(n_dims), S(n_dims, n_dims);
  const double gamma = run_input.gamma, area_ref = run_input.area_ref;
  double cl = 0., cd = 0.;
  for (int m = 0; m < n_dims; m++) Finv(m) = Fvis(m) = inv_force(m) = vis_force(m) = 0.;
  temp_cd = 0.0;
  temp_cl = 0.0;
  const double aoa = atan2(run_input.v_c_ic, run_input.u_c_ic);                     // angle of attack
  const double aos = (n_dims == 3) ? atan2(run_input.w_c_ic, run_input.u_c_ic) : 0.; // angle of side slip
  // one over the dynamic pressure
  const double factor = 1.0 / (0.5 * run_input.rho_c_ic * (run_input.u_c_ic * run_input.u_c_ic + run_input.v_c_ic * run_input.v_c_ic + run_input.w_c_ic * run_input.w_c_ic));
  cp_disu_upts_gpu_cpu();
  if (viscous) cp_grad_disu_upts_gpu_cpu();
  for (int i = 0; i < (int)bdy_ele2ele.size(); i++)
  {
    const int ele = bdy_ele2ele[i];
    for (int l = 0; l < n_inters_per_ele; l++)
    {
      if (bcid(ele, l) < 0) continue;
      const int flag = run_input.bc_list[bcid(ele, l)].get_bc_flag();
      if (flag == SLIP_WALL) coeff_file << "SLIP_WALL" << endl;
      else if (flag == ISOTHERM_WALL) coeff_file << "ISOTHERM_WALL" << endl;
      else if (flag == ADIABAT_WALL) coeff_file << "ADIABAT_WALL" << endl;
      else if (flag == SLIP_WALL_DUAL) coeff_file << "SLIP_WALL_DUAL" << endl;
      if (!(flag == SLIP_WALL || flag == ISOTHERM_WALL || flag == ADIABAT_WALL || flag == SLIP_WALL_DUAL)) continue;
      for (int j = n_cubpts_per_inter(l) - 1; j >= 0; j--)
      {
        const double detjac = inter_detjac_inters_cubpts[l](j, i), wgt = weight_inters_cubpts[l](j);
        for (int m = 0; m < n_dims; m++) loc(m) = loc_inters_cubpts[l](m, j);
        calc_pos(loc, ele, pos);
        for (int m = 0; m < n_fields; m++)
        {
          double value = 0.;
          for (int k = 0; k < n_upts_per_ele; k++) value += opp_inters_cubpts[l](j, k) * disu_upts(0)(k, ele, m);
          u_l(m) = value;
        }
        if (viscous == 1)
          for (int m = 0; m < n_fields; m++)
            for (int n = 0; n < n_dims; n++)
            {
              double value = 0.;
              for (int k = 0; k < n_upts_per_ele; k++) value += opp_inters_cubpts[l](j, k) * grad_disu_upts(k, ele, m, n);
              grad_u_l(m, n) = value;
            }
        for (int m = 0; m < n_dims; m++) norm(m) = norm_inters_cubpts[l](j, i, m);
        double v_sq = 0., p_l;
        if (flag == SLIP_WALL_DUAL)
        {
          // dual consistent: remove the normal velocity first
          double vn_l = 0.;
          for (int m = 0; m < n_dims; m++) vn_l += u_l(m + 1) * norm(m);
          vn_l /= u_l(0);
          for (int m = 0; m < n_dims; m++) u_l(m + 1) = u_l(m + 1) - (vn_l)*norm(m);
        }
        for (int m = 0; m < n_dims; m++) v_sq += (u_l(m + 1) * u_l(m + 1));
        p_l = (gamma - 1.0) * (u_l(n_dims + 1) - 0.5 * v_sq / u_l(0));
        const double cp = (p_l - run_input.p_c_ic) * factor;
        for (int m = 0; m < n_dims; m++) Finv(m) = wgt * (p_l - run_input.p_c_ic) * norm(m) * detjac * factor / area_ref;
        if (n_dims == 2)
        {
          cl = -Finv(0) * sin(aoa) + Finv(1) * cos(aoa);
          cd = Finv(0) * cos(aoa) + Finv(1) * sin(aoa);
        }
        else
        {
          cl = -Finv(0) * sin(aoa) + Finv(1) * cos(aoa);
          cd = Finv(0) * cos(aoa) * cos(aos) + Finv(1) * sin(aoa) + Finv(2) * sin(aoa) * cos(aos);
        }
        if (write_forces)
        {
          coeff_file << scientific;
          for (int m = 0; m < n_dims; m++) coeff_file << setw(18) << setprecision(12) << pos(m) << " ";
          coeff_file << setw(18) << setprecision(12) << cp;
        }
        if (viscous)
        {
          for (int m = 0; m < n_dims; m++)
          {
            drho(m) = grad_u_l(0, m);
            for (int n = 0; n < n_dims; n++) dv(n, m) = (grad_u_l(n + 1, m) - drho(m) * u_l(n + 1) / u_l(0)) / u_l(0);
            dE(m) = (grad_u_l(n_dims + 1, m) - drho(m) * u_l(n_dims + 1)) / u_l(0);
          }
          double diag = 0.;
          for (int m = 0; m < n_dims; m++) diag += dv(m, m);
          diag /= 3.0;
          double inte = u_l(n_dims + 1) / u_l(0);
          for (int m = 0; m < n_dims; m++) inte -= 0.5 * u_l(m + 1) * u_l(m + 1) / u_l(0) / u_l(0);
          const double rt_ratio = (run_input.gamma - 1.0) * inte / (run_input.rt_inf);
          double mu = (run_input.mu_inf) * pow(rt_ratio, 1.5) * (1 + (run_input.c_sth)) / (rt_ratio + (run_input.c_sth));
          mu = mu + run_input.fix_vis * (run_input.mu_inf - mu);
          for (int m = 0; m < n_dims; m++) taun(m) = 0.;
          for (int n = 0; n < n_dims; n++)
          {
            for (int m = 0; m < n_dims; m++) S(m, n) = 0.5 * (dv(m, n) + dv(n, m));
            S(n, n) -= diag;
          }
          for (int m = 0; m < n_dims; m++)
            for (int n = 0; n < n_dims; n++) taun(m) += 2. * mu * S(m, n) * norm(n);
          double taundotn = 0.;
          for (int m = 0; m < n_dims; m++) taundotn += taun(m) * norm(m);
          for (int m = 0; m < n_dims; m++) tautan(m) = taun(m) - taundotn * norm(m);
          double tauw = 0.;
          for (int m = 0; m < n_dims; m++) tauw += pow(tautan(m), 2);
          tauw = sqrt(tauw);
          const double cf = tauw * factor;
          if (write_forces) coeff_file << " " << setw(18) << setprecision(12) << cf;
          for (int m = 0; m < n_dims; m++) Fvis(m) = -wgt * taun(m) * detjac * factor / area_ref;
          if (n_dims == 2)
          {
            cl += -Fvis(0) * sin(aoa) + Fvis(1) * cos(aoa);
            cd += Fvis(0) * cos(aoa) + Fvis(1) * sin(aoa);
          }
          else
          {
            cl += -Fvis(0) * sin(aoa) + Fvis(1) * cos(aoa);
            cd += Fvis(0) * cos(aoa) * cos(aos) + Fvis(1) * sin(aoa) + Fvis(2) * sin(aoa) * cos(aos);
          }
        }
        if (write_forces) coeff_file << endl;
        for (int m = 0; m < n_dims; m++)
        {
          inv_force(m) += Finv(m);
          vis_force(m) += Fvis(m);
        }
        temp_cl += cl;
        temp_cd += cd;
      }
    }
  }
}

void CalcForces(int in_file_num, bool write_forces, struct solution *FlowSol)
{
  char file_name_s[600], forcedir_s[256];
  ofstream coeff_file;
  const int nd = FlowSol->n_dims;
  hf_array<double> temp_inv_force(nd), temp_vis_force(nd);
  if (write_forces)
  {
    snprintf(forcedir_s, sizeof(forcedir_s), "force_files_%09d", in_file_num);
    if (FlowSol->rank == 0)
    {
      struct stat st;
      if (stat(forcedir_s, &st) == -1) mkdir(forcedir_s, 0755);
      else if (DIR *dir = opendir(forcedir_s))
      {
        while (struct dirent *fn = readdir(dir))
          if (strcmp(fn->d_name, ".") != 0 && strcmp(fn->d_name, "..") != 0) remove((string(forcedir_s) + '/' + fn->d_name).c_str());
        closedir(dir);
      }
    }
    struct stat st;
    for (int spin = 0; stat(forcedir_s, &st) == -1 && spin < 10000; spin++) hf_dev_sync(FlowSol->ctx);
    snprintf(file_name_s, sizeof(file_name_s), "%s/cp_%.09d_p%.04d.dat", forcedir_s, in_file_num, FlowSol->rank);
    coeff_file.open(file_name_s);
    coeff_file << setw(18) << "x" << setw(18) << "Cp" << setw(18) << "Cf" << endl;
  }
  FlowSol->inv_force.setup(nd);
  FlowSol->vis_force.setup(nd);
  for (int m = 0; m < nd; m++) FlowSol->inv_force(m) = FlowSol->vis_force(m) = 0.;
  FlowSol->coeff_lift = 0.0;
  FlowSol->coeff_drag = 0.0;
  for (int i = 0; i < FlowSol->n_ele_types; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0)
    {
      double temp_cl, temp_cd;
      FlowSol->mesh_eles(i)->compute_wall_forces(temp_inv_force, temp_vis_force, temp_cl, temp_cd, coeff_file, write_forces);
      for (int m = 0; m < nd; m++)
      {
        FlowSol->inv_force(m) += temp_inv_force(m);
        FlowSol->vis_force(m) += temp_vis_force(m);
      }
      FlowSol->coeff_lift += temp_cl;
      FlowSol->coeff_drag += temp_cd;
    }
  if (FlowSol->nproc > 1)
  {
    double buf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int m = 0; m < nd; m++) { buf[m] = FlowSol->inv_force(m); buf[3 + m] = FlowSol->vis_force(m); }
    buf[6] = FlowSol->coeff_lift; buf[7] = FlowSol->coeff_drag;
    hf_check(hf_dev_allreduce_sum(FlowSol->ctx, buf, 8));
    for (int m = 0; m < nd; m++) { FlowSol->inv_force(m) = buf[m]; FlowSol->vis_force(m) = buf[3 + m]; }
    FlowSol->coeff_lift = buf[6]; FlowSol->coeff_drag = buf[7];
  }
  if (write_forces) coeff_file.close();
}
