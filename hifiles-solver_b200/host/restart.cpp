// ASCII restart files, reading side (reference src/solver.cpp:377-434 read_restart_ascii; per element type
// read_restart_info_ascii + set_opp_r + eval_nodal_basis_restart, e.g. src/eles_hexas.cpp:799-828, 1149-1162,
// src/eles_tris.cpp:542-575, 721-736, src/eles_tets.cpp:803-846, 996-1011, src/eles_pris.cpp:734-790, 1002-1026;
// eles::read_restart_data_ascii src/eles.cpp:655-751).  The file may hold another polynomial order than the run:
// opp_r interpolates from the file's solution points to the run's, with the nodal basis of the file's points.
// The writer (write_restart_ascii, reference src/output.cpp:1753-1818) is at the end of this file; HDF5 restart files need HDF5 (absent) and fail at input time as in the reference.
#include "hifiles.h"
#include <sys/stat.h>
#include <cerrno>
#include <fstream>
#include <cstdio>
#include <vector>
using namespace std;

namespace
{
const char *title_of(int type)
{
  static const char *t[5] = {"TRIS", "QUADS", "TETS", "PRIS", "HEXAS"};
  return t[type];
}

// geometry of the file's solution points for one element type
struct rest_info
{
  int order_rest = 0, n_upts_per_ele_rest = 0, n_upts_tri_rest = 0;
  hf_array<double> loc_1d_upts_rest;   // quads, hexas, prisms (1-D part)
  hf_array<double> loc_upts_rest;      // tris, tets (n_dims, n_upts); prisms: triangle part (2, n_upts_tri)
  hf_array<double> inv_vandermonde_rest;
};

// position the stream behind the element type's title line; false: the file holds no such elements
bool seek_title(std::ifstream &f, const char *title)
{
  string str;
  while (1)
  {
    getline(f, str);
    if (str == title) return true;
    if (f.eof()) return false;
  }
}

bool read_restart_info_ascii(std::ifstream &f, eles *e, rest_info &R)
{
  const int type = e->get_ele_type();
  string str;
  if (!seek_title(f, title_of(type))) return false;
  getline(f, str);
  f >> R.order_rest;
  getline(f, str);
  getline(f, str);
  f >> R.n_upts_per_ele_rest;
  getline(f, str);
  getline(f, str);
  if (type == QUAD || type == HEX)
  {
    R.loc_1d_upts_rest.setup(R.order_rest + 1);
    for (int i = 0; i < R.order_rest + 1; ++i) f >> R.loc_1d_upts_rest(i);
  }
  else if (type == PRISM)
  {
    f >> R.n_upts_tri_rest;
    getline(f, str);
    getline(f, str);
    R.loc_1d_upts_rest.setup(R.order_rest + 1);
    R.loc_upts_rest.setup(2, R.n_upts_tri_rest);
    for (int i = 0; i < R.order_rest + 1; i++) f >> R.loc_1d_upts_rest(i);
    getline(f, str);
    getline(f, str);
    for (int i = 0; i < R.n_upts_tri_rest; i++)
      for (int j = 0; j < 2; j++) f >> R.loc_upts_rest(j, i);
  }
  else
  {
    const int nd = e->n_dims;
    R.loc_upts_rest.setup(nd, R.n_upts_per_ele_rest);
    for (int i = 0; i < R.n_upts_per_ele_rest; i++)
      for (int j = 0; j < nd; j++) f >> R.loc_upts_rest(j, i);
  }
  // Vandermonde matrix of the file's points in the orthonormal Dubiner basis, inverted (simplex parts)
  if (type == TRI || type == TET || type == PRISM)
  {
    const int n = type == PRISM ? R.n_upts_tri_rest : R.n_upts_per_ele_rest;
    hf_array<double> V(n, n);
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
        V(i, j) = type == TET ? eval_dubiner_basis_3d(R.loc_upts_rest(0, i), R.loc_upts_rest(1, i), R.loc_upts_rest(2, i), j, R.order_rest)
                              : eval_dubiner_basis_2d(R.loc_upts_rest(0, i), R.loc_upts_rest(1, i), j, R.order_rest);
    R.inv_vandermonde_rest = inv_array(V);
  }
  return true;
}

// nodal basis function in_index of the file's points at a location of the reference element
double eval_nodal_basis_restart(int type, rest_info &R, int in_index, const double *loc)
{
  const int n1 = R.order_rest + 1;
  if (type == HEX)
  {
    int i = in_index / (n1 * n1);
    int j = (in_index - n1 * n1 * i) / n1;
    int k = in_index - n1 * j - n1 * n1 * i;
    return eval_lagrange(loc[0], k, R.loc_1d_upts_rest) * eval_lagrange(loc[1], j, R.loc_1d_upts_rest) * eval_lagrange(loc[2], i, R.loc_1d_upts_rest);
  }
  if (type == QUAD)
  {
    int i = in_index / n1;
    int j = in_index - n1 * i;
    return eval_lagrange(loc[0], j, R.loc_1d_upts_rest) * eval_lagrange(loc[1], i, R.loc_1d_upts_rest);
  }
  if (type == PRISM)
  {
    const int index_tri = in_index % R.n_upts_tri_rest, index_1d = in_index / R.n_upts_tri_rest;
    double tri = 0.;
    for (int i = 0; i < R.n_upts_tri_rest; i++) tri += R.inv_vandermonde_rest(i, index_tri) * eval_dubiner_basis_2d(loc[0], loc[1], i, R.order_rest);
    const double oned = eval_lagrange(loc[2], index_1d, R.loc_1d_upts_rest);
    return tri * oned;
  }
  // From Hesthaven, equation 3.3: V^T l = P, or l = (V^-1)^T P
  double v = 0.;
  for (int i = 0; i < R.n_upts_per_ele_rest; i++)
    v += R.inv_vandermonde_rest(i, in_index) * (type == TET ? eval_dubiner_basis_3d(loc[0], loc[1], loc[2], i, R.order_rest)
                                                             : eval_dubiner_basis_2d(loc[0], loc[1], i, R.order_rest));
  return v;
}

// position of value in a sorted array without repeated entries, -1 if absent (reference src/funcs.cpp:1677-1716)
int index_locate_int(int value, const int *a, int size)
{
  int jl = 0, ju = size - 1;
  if (a[ju] <= a[0] && ju != 0) FatalError("ERROR, hf_array not sorted, exiting");
  while (ju - jl > 1)
  {
    int jm = (ju + jl) >> 1;
    if (value >= a[jm]) jl = jm;
    else ju = jm;
  }
  if (value == a[0]) return 0;
  if (value == a[size - 1]) return size - 1;
  if (value == a[jl]) return jl;
  return -1;
}

void read_restart_data_ascii(std::ifstream &f, eles *e, rest_info &R, const hf_array<double> &opp_r)
{
  string str;
  f.clear();
  f.seekg(0, f.beg);
  while (1)
  {
    getline(f, str);
    if (str == title_of(e->get_ele_type())) break;
    if (f.eof()) return; // the file does not contain my elements
  }
  while (1)
  {
    getline(f, str);
    if (str == "n_eles") break;
    if (f.eof()) FatalError("restart file: n_eles record missing");
  }
  int num_eles_to_read;
  f >> num_eles_to_read;
  getline(f, str);
  // skip the ele2global_ele lines
  getline(f, str);
  getline(f, str);
  getline(f, str);
  const int nur = R.n_upts_per_ele_rest, nu = e->n_upts_per_ele, nf = e->n_fields;
  hf_array<double> rest(nur, nf);
  for (int i = 0; i < num_eles_to_read; i++)
  {
    int ele;
    f >> ele;
    int index = index_locate_int(ele, e->ele2global_ele.get_ptr_cpu(), e->n_eles);
    if (index != -1)
    {
      for (int j = 0; j < nur; j++)
        for (int k = 0; k < nf; k++) f >> rest(j, k);
      for (int m = 0; m < nf; m++)
        for (int j = 0; j < nu; j++)
        {
          double value = 0.;
          for (int k = 0; k < nur; k++) value += opp_r(j, k) * rest(k, m);
          e->disu_upts(0)(j, index, m) = value;
        }
    }
    else
    {
      getline(f, str);
      for (int j = 0; j < nur; j++) getline(f, str);
    }
  }
  if (!f) FatalError("restart file: truncated element data");
  e->set_h_ref();
}
} // namespace

void read_restart_ascii(int in_file_num, int in_n_files, struct solution *FlowSol)
{
  char name[256];
  auto file_name = [&](int j) {
    if (in_n_files != 1) snprintf(name, sizeof(name), "Rest_%.09d/Rest_%.09d_p%.04d.dat", in_file_num, in_file_num, j); // in folder
    else snprintf(name, sizeof(name), "Rest_%.09d_p%.04d.dat", in_file_num, j);
    return name;
  };
  const int nt = FlowSol->n_ele_types;
  std::vector<rest_info> info(nt);
  std::vector<hf_array<double>> opp_r(nt);
  std::vector<char> have(nt, 0);
  // open the restart files and read the element-type info
  for (int i = 0; i < nt; i++)
  {
    eles *e = FlowSol->mesh_eles(i);
    if (e->get_n_eles() == 0) continue;
    bool found = false;
    for (int j = 0; j < in_n_files && !found; j++)
    {
      std::ifstream f(file_name(j));
      f.precision(15);
      if (!f) FatalError("Could not open restart file ");
      f >> FlowSol->time;
      found = read_restart_info_ascii(f, e, info[i]);
    }
    if (!found) continue; // no file holds this type: its initial data stay as allocated (zero), as in the reference
    have[i] = 1;
    // set opp_r (solution at restart points to solution at solution points, reference src/eles.cpp:3692-3710)
    const int nu = e->n_upts_per_ele, nur = info[i].n_upts_per_ele_rest, nd = e->n_dims;
    opp_r[i].setup(nu, nur);
    double loc[3];
    for (int r = 0; r < nur; r++)
      for (int j = 0; j < nu; j++)
      {
        for (int k = 0; k < nd; k++) loc[k] = e->loc_upts(k, j);
        opp_r[i](j, r) = eval_nodal_basis_restart(e->get_ele_type(), info[i], r, loc);
      }
  }
  // now open all the restart files one by one and store the data belonging to this processor
  for (int j = 0; j < in_n_files; j++)
  {
    std::ifstream f(file_name(j));
    f.precision(15);
    if (f.fail()) FatalError(string("Could not open restart file ") + name);
    for (int i = 0; i < nt; i++)
    {
      eles *e = FlowSol->mesh_eles(i);
      if (e->get_n_eles() != 0 && have[i]) read_restart_data_ascii(f, e, info[i], opp_r[i]);
    }
  }
}

// ASCII restart file Rest_<iter>_p<rank>.dat: time, then per element type a header with the solution-point set and the
// solution of every element under its global id, 15 significant digits (reference src/output.cpp:1753-1818,
// src/eles.cpp:845-869, <type>::write_restart_info_ascii).  It is one of the reference's parity comparators.  A partitioned run
// writes one file per rank into the folder Rest_<iter>/ (reference src/output.cpp:1760-1792; the reference's rank 0 empties an
// existing folder first and all ranks meet at a barrier -- here every rank creates the folder if need be and overwrites its own file).
void write_restart_ascii(struct solution *FlowSol, int in_file_num)
{
  char name[256];
  if (FlowSol->nproc > 1)
  {
    char folder[64];
    snprintf(folder, sizeof(folder), "Rest_%.09d", in_file_num);
    struct stat st;
    if (stat(folder, &st) == -1 && mkdir(folder, 0755) == -1 && errno != EEXIST) FatalError("cannot create the restart folder");
    snprintf(name, sizeof(name), "Rest_%.09d/Rest_%.09d_p%.04d.dat", in_file_num, in_file_num, FlowSol->rank);
  }
  else
    snprintf(name, sizeof(name), "Rest_%.09d_p%.04d.dat", in_file_num, FlowSol->rank);
  if (FlowSol->rank == 0) cout << "Writing Restart file for step " << in_file_num << " ...." << flush;
  ofstream f(name);
  f.precision(15);
  f << FlowSol->time << endl;
  for (int t = 0; t < FlowSol->n_ele_types; t++)
  {
    eles *e = FlowSol->mesh_eles(t);
    if (e->get_n_eles() == 0) continue;
    if (FlowSol->ctx) e->cp_disu_upts_gpu_cpu();
    const int type = e->get_ele_type();
    static const char *title[5] = {"TRIS", "QUADS", "TETS", "PRIS", "HEXAS"};
    static const char *count[5] = {"Number of solution points per triangular element", "Number of solution points per quadrilateral element",
                                   "Number of solution points per element", "Number of solution points per prismatic element",
                                   "Number of solution points per hexahedral element"};
    f << title[type] << endl << "Order" << endl << e->order << endl << count[type] << endl << e->n_upts_per_ele << endl;
    if (type == QUAD || type == HEX)
    {
      f << "Location of solution points in 1D" << endl;
      for (int i = 0; i < e->order + 1; ++i) f << e->loc_1d_upts(i) << " ";
      f << endl;
    }
    else if (type == PRISM)
    {
      eles_pris *p = static_cast<eles_pris *>(e);
      f << "Number of solution points in triangle" << endl << p->n_upts_tri << endl;
      f << "Location of solution points in 1D" << endl;
      for (int i = 0; i < e->order + 1; ++i) f << p->loc_upts_pri_1d(i) << " ";
      f << endl;
      f << "Location of solution points in triangle" << endl;
      for (int i = 0; i < p->n_upts_tri; i++)
      {
        for (int j = 0; j < 2; j++) f << p->loc_upts_pri_tri(j, i) << " ";
        f << endl;
      }
    }
    else
    {
      f << (type == TRI ? "Location of solution points in triangular elements" : "Location of solution points in tetrahedral elements") << endl;
      for (int i = 0; i < e->n_upts_per_ele; i++)
      {
        for (int j = 0; j < e->n_dims; j++) f << e->loc_upts(j, i) << " ";
        f << endl;
      }
    }
    f << "n_eles" << endl << e->n_eles << endl << "ele2global_ele hf_array" << endl;
    for (int i = 0; i < e->n_eles; i++) f << e->ele2global_ele(i) << " ";
    f << endl << "data" << endl;
    for (int i = 0; i < e->n_eles; i++)
    {
      f << e->ele2global_ele(i) << endl;
      for (int j = 0; j < e->n_upts_per_ele; j++)
      {
        for (int k = 0; k < e->n_fields; k++) f << e->disu_upts(0)(j, i, k) << " ";
        f << endl;
      }
    }
    f << endl;
  }
  if (FlowSol->rank == 0) cout << "done" << endl;
}

