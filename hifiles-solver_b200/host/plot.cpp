// Paraview output (reference output::write_vtu, src/output.cpp:462-900): every element becomes one <Piece> of
// p_res-resolution plot points and linear sub-cells; density, velocity and specific total energy are interpolated from the
// solution points with opp_p.  Plot points, their numbering and the sub-cell connectivity follow the reference's
// per-type rules (set_loc_ppts / set_connectivity_plot of src/eles_{hexas,quads,tris,tets,pris}.cpp) so that the files
// agree with the reference's to the printed digits.  One .vtu per run in serial; with several ranks every rank writes
// <name>_<iter>/<name>_<iter>_<rank>.vtu and rank 0 the .pvtu index.  Optional diagnostic fields (u, v, w, energy, mach,
// pressure, vorticity, q_criterion, scaled_q_criterion, sensor: eles::calc_diagnostic_fields_ppts, src/eles.cpp:3858-4010) are
// evaluated at the plot points from the interpolated solution and gradient; time-averaged fields (average_fields) are
// interpolated from the running averages the device keeps (hf_dev_time_average).  The
// solution (and gradient) is copied device -> host first (output::CopyGPUCPU).
#include "hifiles.h"
#include <cstdio>
#include <fstream>
#include <vector>
#include <sys/stat.h>
#include <dirent.h>
#include <cstring>
using namespace std;

namespace
{
// plot-point numbering of the simplex types: rows of decreasing length
inline int tri_idx(int i, int j, int p) { return i + j * (p + 1) - (j * (j + 1)) / 2; }
inline int tet_layers(int n) { return n * (n + 1) * (n + 2) / 6; }
inline int tet_idx(int i, int j, int k, int p) { return tet_layers(p) - tet_layers(p - k) + j * (p - k) - ((j - 1) * j) / 2 + i; }

struct plot_topology
{
  int n_ppts = 0, n_peles = 0, n_verts = 0;
  hf_array<double> loc_ppts; // (dim, ppt)
  vector<int> con;           // [cell][vert]
  hf_array<double> opp_p;    // (ppt, upt)
};

void add_cell(plot_topology &T, std::initializer_list<int> v)
{
  for (int x : v) T.con.push_back(x);
  T.n_peles++;
}

void build_topology(eles *e, int p, plot_topology &T)
{
  const int type = e->get_ele_type(), nd = e->n_dims;
  const double h = 1.0 * (p - 1);
  auto put = [&](int q, int i, int j, int k) {
    T.loc_ppts(0, q) = -1.0 + ((2.0 * i) / h);
    T.loc_ppts(1, q) = -1.0 + ((2.0 * j) / h);
    if (nd == 3) T.loc_ppts(2, q) = -1.0 + ((2.0 * k) / h);
  };
  if (type == HEX)
  {
    T.n_ppts = p * p * p; T.n_verts = 8;
    T.loc_ppts.setup(3, T.n_ppts);
    for (int k = 0; k < p; k++) for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) put(i + p * j + p * p * k, i, j, k);
    for (int k = 0; k < p - 1; k++) for (int l = 0; l < p - 1; l++) for (int m = 0; m < p - 1; m++)
    {
      const int a = m + p * l + p * p * k, b = a + p * p;
      add_cell(T, {a, a + 1, a + p + 1, a + p, b, b + 1, b + p + 1, b + p});
    }
  }
  else if (type == QUAD)
  {
    T.n_ppts = p * p; T.n_verts = 4;
    T.loc_ppts.setup(2, T.n_ppts);
    for (int j = 0; j < p; j++) for (int i = 0; i < p; i++) put(i + p * j, i, j, 0);
    for (int k = 0; k < p - 1; k++) for (int l = 0; l < p - 1; l++)
    {
      const int a = l + p * k;
      add_cell(T, {a, a + 1, a + p + 1, a + p});
    }
  }
  else if (type == TRI)
  {
    T.n_ppts = (p + 1) * p / 2; T.n_verts = 3;
    T.loc_ppts.setup(2, T.n_ppts);
    for (int j = 0; j < p; j++) for (int i = 0; i < p - j; i++) put(tri_idx(i, j, p), i, j, 0);
    for (int k = 0; k < p - 1; k++) for (int l = 0; l < p - k - 1; l++) // upright sub-triangles
      add_cell(T, {tri_idx(l, k, p), tri_idx(l, k, p) + 1, tri_idx(l, k + 1, p)});
    for (int k = 0; k < p - 2; k++) for (int l = 0; l < p - k - 2; l++) // inverted ones between them
      add_cell(T, {tri_idx(l + 1, k, p), tri_idx(l + 1, k + 1, p), tri_idx(l, k + 1, p)});
  }
  else if (type == TET)
  {
    T.n_ppts = (p + 2) * (p + 1) * p / 6; T.n_verts = 4;
    T.loc_ppts.setup(3, T.n_ppts);
    for (int k = 0; k < p; k++) for (int j = 0; j < p - k; j++) for (int i = 0; i < p - k - j; i++) put(tet_idx(i, j, k, p), i, j, k);
    // corner tetrahedra of every lattice cube
    for (int k = 0; k < p - 1; k++) for (int j = 0; j < p - 1 - k; j++) for (int i = 0; i < p - 1 - k - j; i++)
      add_cell(T, {tet_idx(i, j, k, p), tet_idx(i + 1, j, k, p), tet_idx(i, j + 1, k, p), tet_idx(i, j, k + 1, p)});
    // the octahedron next to it, split into four
    for (int k = 0; k < p - 2; k++) for (int j = 0; j < p - 2 - k; j++) for (int i = 0; i < p - 2 - k - j; i++)
    {
      const int v0 = tet_idx(i + 1, j, k, p), v1 = tet_idx(i + 1, j + 1, k, p), v2 = tet_idx(i + 1, j, k + 1, p);
      const int v3 = tet_idx(i, j + 1, k + 1, p), v4 = tet_idx(i, j, k + 1, p), v5 = tet_idx(i, j + 1, k, p);
      add_cell(T, {v0, v2, v1, v4});
      add_cell(T, {v2, v3, v1, v4});
      add_cell(T, {v5, v1, v3, v4});
      add_cell(T, {v0, v4, v1, v5});
    }
    // the opposite corner
    for (int k = 0; k < p - 3; k++) for (int j = 0; j < p - 3 - k; j++) for (int i = 0; i < p - 3 - k - j; i++)
      add_cell(T, {tet_idx(i + 1, j + 1, k, p), tet_idx(i + 1, j, k + 1, p), tet_idx(i, j + 1, k + 1, p), tet_idx(i + 1, j + 1, k + 1, p)});
  }
  else
  {
    const int layer = p * (p + 1) / 2;
    T.n_ppts = layer * p; T.n_verts = 6;
    T.loc_ppts.setup(3, T.n_ppts);
    for (int k = 0; k < p; k++) for (int j = 0; j < p; j++) for (int i = 0; i < p - j; i++) put(layer * k + tri_idx(i, j, p), i, j, k);
    for (int l = 0; l < p - 1; l++) for (int j = 0; j < p - 1; j++) for (int k = 0; k < p - j - 1; k++)
    {
      const int a = tri_idx(k, j, p) + l * layer, b = a + 1, c = tri_idx(k, j + 1, p) + l * layer;
      add_cell(T, {a, b, c, a + layer, b + layer, c + layer});
    }
    for (int l = 0; l < p - 1; l++) for (int j = 0; j < p - 2; j++) for (int k = 0; k < p - j - 2; k++)
    {
      const int a = tri_idx(k + 1, j, p) + l * layer, b = tri_idx(k + 1, j + 1, p) + l * layer, c = b - 1;
      add_cell(T, {a, b, c, a + layer, b + layer, c + layer});
    }
  }
  // solution points -> plot points (reference eles::set_opp_p, src/eles.cpp:3600-3621)
  hf_array<double> loc(nd);
  T.opp_p.setup(T.n_ppts, e->n_upts_per_ele);
  for (int i = 0; i < e->n_upts_per_ele; i++)
    for (int j = 0; j < T.n_ppts; j++)
    {
      for (int k = 0; k < nd; k++) loc(k) = T.loc_ppts(k, j);
      T.opp_p(j, i) = e->eval_nodal_basis(i, loc);
    }
}

// stream output with the reference's settings: precision 15, default float format
struct num { double v; };
ostream &operator<<(ostream &o, const num &n)
{
  char b[40];
  snprintf(b, sizeof(b), "%.15g", n.v);
  return o << b;
}
} // namespace

// everything a plot file holds for one element type: topology, and per element the fields at the plot points
struct plot_fields
{
  eles *e = nullptr;
  plot_topology T;
  int n_points = 0, n_cells = 0, n_verts = 0, n_fields = 0, n_dims = 0, nu = 0, n_diag_fields = 0, n_average_fields = 0;
  bool have_grad = false, have_avg = false;
  hf_array<double> u, g, diag, avg, pos;

  // copies the arrays device -> host (output::CopyGPUCPU) and builds the plot topology
  void setup(eles *in_e, struct solution *FlowSol, int in_file_num)
  {
    e = in_e;
    n_diag_fields = run_input.n_diagnostic_fields;
    n_average_fields = run_input.n_average_fields;
    if (run_input.equation != 0) FatalError("plot files are built for the Euler / Navier-Stokes equations");
    if (!FlowSol->no_device) e->cp_disu_upts_gpu_cpu();
    // gradient of the last monitored residual evaluation; before the first step the reference's array is still zero
    have_grad = n_diag_fields > 0 && run_input.viscous && !FlowSol->no_device && in_file_num != FlowSol->ini_iter;
    if (have_grad) e->cp_grad_disu_upts_gpu_cpu();
    // running averages: zero before the first step, as the reference's freshly allocated array
    have_avg = n_average_fields > 0 && !FlowSol->no_device && in_file_num != FlowSol->ini_iter;
    if (have_avg) e->cp_disu_average_upts_gpu_cpu();
    if (n_diag_fields > 0 && run_input.shock_cap && !FlowSol->no_device) e->cp_sensor_gpu_cpu();
    build_topology(e, run_input.p_res, T);
    n_points = T.n_ppts; n_cells = T.n_peles; n_verts = T.n_verts; n_fields = e->n_fields; n_dims = e->n_dims; nu = e->n_upts_per_ele;
    u.setup(n_points, n_fields);
    g.setup(n_points, n_fields, n_dims);
    diag.setup(n_points, n_diag_fields > 0 ? n_diag_fields : 1);
    avg.setup(n_points, n_average_fields > 0 ? n_average_fields : 1);
    pos.setup(n_points, n_dims);
  }

  // fields of element j at its plot points (reference eles::calc_disu_ppts, calc_time_average_ppts, calc_grad_disu_ppts,
  // calc_sensor_ppts, calc_diagnostic_fields_ppts, calc_pos_ppts: src/eles.cpp:3714-4010)
  void eval(int j)
  {
    for (int m = 0; m < n_fields; m++)
      for (int k = 0; k < n_points; k++)
      {
        double a = 0.;
        for (int l = 0; l < nu; l++) a += e->disu_upts(0)(l, j, m) * T.opp_p(k, l);
        u(k, m) = a;
      }
    for (int m = 0; m < n_average_fields; m++)
      for (int k = 0; k < n_points; k++)
      {
        double a = 0.;
        if (have_avg)
          for (int l = 0; l < nu; l++) a += e->disu_average_upts(l, j, m) * T.opp_p(k, l);
        avg(k, m) = a;
      }
    hf_array<double> loc(n_dims), p(n_dims);
    for (int k = 0; k < n_points; k++)
    {
      for (int l = 0; l < n_dims; l++) loc(l) = T.loc_ppts(l, k);
      e->calc_pos(loc, j, p);
      for (int l = 0; l < n_dims; l++) pos(k, l) = p(l);
    }
    if (n_diag_fields == 0) return;
    for (int d = 0; d < n_dims; d++)
      for (int m = 0; m < n_fields; m++)
        for (int k = 0; k < n_points; k++)
        {
          double a = 0.;
          if (have_grad)
            for (int l = 0; l < nu; l++) a += e->grad_disu_upts(l, j, m, d) * T.opp_p(k, l);
          g(k, m, d) = a;
        }
    const double sensor = (run_input.shock_cap && e->sensor.size() > 0) ? e->sensor(j) : 0.;
    for (int k = 0; k < n_points; k++)
    {
      double v_sq = 0.;
      for (int m = 0; m < n_dims; m++) v_sq += (u(k, m + 1) * u(k, m + 1));
      v_sq /= u(k, 0) * u(k, 0);
      const double pressure = (run_input.gamma - 1.0) * (u(k, n_dims + 1) - 0.5 * u(k, 0) * v_sq);
      const double irho = 1. / u(k, 0);
      for (int q = 0; q < n_diag_fields; q++)
      {
        const string &name = run_input.diagnostic_fields(q);
        double val = 0.;
        if (name == "u") val = u(k, 1) * irho;
        else if (name == "v") val = u(k, 2) * irho;
        else if (name == "w") val = (n_dims == 2) ? 0. : u(k, 3) * irho;
        else if (name == "energy") val = u(k, n_dims + 1);
        else if (name == "mach") val = sqrt(v_sq / (run_input.gamma * pressure / u(k, 0)));
        else if (name == "pressure") val = pressure;
        else if (name == "vorticity" || name == "q_criterion" || name == "scaled_q_criterion")
        {
          if (!run_input.viscous) FatalError("Trying to calculate diagnostic field only supported by viscous simualtion");
          double dvel[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}; // d v_a / d x_b
          for (int a = 0; a < n_dims; a++)
          {
            const double va = u(k, a + 1) * irho;
            for (int b = 0; b < n_dims; b++) dvel[a][b] = irho * (g(k, a + 1, b) - va * g(k, 0, b));
          }
          if (n_dims == 2)
          {
            if (name != "vorticity") FatalError("Q criterion Not implemented in 2D");
            val = fabs(dvel[1][0] - dvel[0][1]);
          }
          else
          {
            double wx = dvel[2][1] - dvel[1][2], wy = dvel[0][2] - dvel[2][0], wz = dvel[1][0] - dvel[0][1];
            if (name == "vorticity") val = sqrt(wx * wx + wy * wy + wz * wz);
            else
            {
              wx *= 0.5; wy *= 0.5; wz *= 0.5;
              const double Sxy = 0.5 * (dvel[0][1] + dvel[1][0]), Sxz = 0.5 * (dvel[0][2] + dvel[2][0]), Syz = 0.5 * (dvel[1][2] + dvel[2][1]);
              const double SS = dvel[0][0] * dvel[0][0] + dvel[1][1] * dvel[1][1] + dvel[2][2] * dvel[2][2] + 2 * Sxy * Sxy + 2 * Sxz * Sxz + 2 * Syz * Syz;
              const double OO = 2 * wx * wx + 2 * wy * wy + 2 * wz * wz;
              val = (name == "q_criterion") ? 0.5 * (OO - SS) : 0.5 * (OO - SS) / (SS + 1.e-24);
            }
          }
        }
        else if (name == "sensor")
        {
          if (!run_input.shock_cap) FatalError("Sensor unavailable");
          val = sensor;
        }
        else
          FatalError("plot_quantity not recognized");
        if (std::isnan(val)) FatalError("NaN in the calculation of plot quantity " + name);
        diag(k, q) = val;
      }
    }
  }
};

// empty an existing output directory or create it (rank 0), then wait until it is there (all ranks, same file system)
static void prepare_directory(const char *dir_s, struct solution *FlowSol)
{
  if (FlowSol->rank == 0)
  {
    struct stat st;
    if (stat(dir_s, &st) == -1) mkdir(dir_s, 0755);
    else if (DIR *dir = opendir(dir_s))
    {
      while (struct dirent *fn = readdir(dir))
        if (strcmp(fn->d_name, ".") != 0 && strcmp(fn->d_name, "..") != 0) remove((string(dir_s) + '/' + fn->d_name).c_str());
      closedir(dir);
    }
  }
  struct stat st;
  for (int spin = 0; stat(dir_s, &st) == -1 && spin < 10000; spin++) hf_dev_sync(FlowSol->ctx);
}

void write_vtu(int in_file_num, struct solution *FlowSol)
{
  const int n_diag_fields = run_input.n_diagnostic_fields, n_average_fields = run_input.n_average_fields;
  const int my_rank = FlowSol->rank, n_proc = FlowSol->nproc;
  static const int vtktypes[5] = {5, 9, 10, 13, 12}; // tri, quad, tet, prism, hex (vtkCellType.h)
  char dumpnum_s[256], vtu_s[600], pvtu_s[300];
  const char *name = run_input.data_file_name.c_str();
  snprintf(dumpnum_s, sizeof(dumpnum_s), "%s_%.09d", name, in_file_num);
  if (n_proc > 1)
  {
    snprintf(vtu_s, sizeof(vtu_s), "%s/%s_%d.vtu", dumpnum_s, dumpnum_s, my_rank);
    snprintf(pvtu_s, sizeof(pvtu_s), "%s.pvtu", dumpnum_s);
    prepare_directory(dumpnum_s, FlowSol);
    if (my_rank == 0)
    {
      cout << "Writing Paraview file " << dumpnum_s << " ...." << flush;
      ofstream w(pvtu_s);
      w << "<?xml version=\"1.0\" ?>" << endl;
      w << "<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << endl;
      w << "	<PUnstructuredGrid GhostLevel=\"1\">" << endl;
      w << "		<PPointData Scalars=\"Density\" Vectors=\"Velocity\">" << endl;
      w << "			<PDataArray type=\"Float32\" Name=\"Density\" />" << endl;
      w << "			<PDataArray type=\"Float32\" Name=\"Velocity\" NumberOfComponents=\"3\" />" << endl;
      w << "			<PDataArray type=\"Float32\" Name=\"SpecificTotalEnergy\" />" << endl;
      for (int m = 0; m < n_average_fields; m++) w << "			<PDataArray type=\"Float32\" Name=\"" << run_input.average_fields(m) << "\" />" << endl;
      for (int m = 0; m < n_diag_fields; m++) w << "			<PDataArray type=\"Float32\" Name=\"" << run_input.diagnostic_fields(m) << "\" />" << endl;
      w << "		</PPointData>" << endl;
      w << "		<PPoints>" << endl;
      w << "			<PDataArray type=\"Float32\" Name=\"Points\" NumberOfComponents=\"3\" />" << endl;
      w << "		</PPoints>" << endl;
      for (int i = 0; i < n_proc; ++i) w << "		<Piece Source=\"" << dumpnum_s << "/" << dumpnum_s << "_" << i << ".vtu" << "\" />" << endl;
      w << "	</PUnstructuredGrid>" << endl;
      w << "</VTKFile>" << endl;
    }
  }
  else
  {
    snprintf(vtu_s, sizeof(vtu_s), "%s.vtu", dumpnum_s);
    cout << "Writing Paraview file " << dumpnum_s << " ... " << flush;
  }
  ofstream w(vtu_s);
  if (!w) FatalError(string("cannot open ") + vtu_s);
  w << "<?xml version=\"1.0\" ?>" << endl;
  w << "<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">" << endl;
  w << "	<UnstructuredGrid>" << endl;
  for (int t = 0; t < FlowSol->n_ele_types; t++)
  {
    eles *e = FlowSol->mesh_eles(t);
    const int n_eles = e->get_n_eles();
    if (n_eles == 0) continue;
    plot_fields P;
    P.setup(e, FlowSol, in_file_num);
    const int n_points = P.n_points, n_cells = P.n_cells, n_verts = P.n_verts, n_dims = P.n_dims;
    for (int j = 0; j < n_eles; j++)
    {
      P.eval(j);
      const hf_array<double> &u = P.u;
      w << "		<Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">" << endl;
      w << "			<PointData>" << endl;
      w << "				<DataArray type= \"Float32\" Name=\"Density\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_points; k++) w << num{u(k, 0)} << " ";
      w << endl << "				</DataArray>" << endl;
      w << "				<DataArray type= \"Float32\" NumberOfComponents=\"3\" Name=\"Velocity\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_points; k++)
      {
        w << num{u(k, 1) / u(k, 0)} << " " << num{u(k, 2) / u(k, 0)} << " ";
        if (n_dims == 2) w << num{0.0} << " ";
        else w << num{u(k, 3) / u(k, 0)} << " ";
      }
      w << endl << "				</DataArray>" << endl;
      w << "				<DataArray type= \"Float32\" Name=\"SpecificTotalEnergy\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_points; k++) w << num{u(k, n_dims + 1) / u(k, 0)} << " ";
      w << endl << "				</DataArray>" << endl;
      for (int m = 0; m < n_average_fields; m++)
      {
        w << "				<DataArray type= \"Float32\" Name=\"" << run_input.average_fields(m) << "\" format=\"ascii\">" << endl;
        for (int k = 0; k < n_points; k++) w << num{P.avg(k, m)} << " ";
        w << endl << "				</DataArray>" << endl;
      }
      for (int m = 0; m < n_diag_fields; m++)
      {
        w << "				<DataArray type= \"Float32\" Name=\"" << run_input.diagnostic_fields(m) << "\" format=\"ascii\">" << endl;
        for (int k = 0; k < n_points; k++) w << num{P.diag(k, m)} << " ";
        w << endl << "				</DataArray>" << endl;
      }
      w << "			</PointData>" << endl;
      w << "			<Points>" << endl;
      w << "				<DataArray type=\"Float32\" NumberOfComponents=\"3\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_points; k++)
      {
        for (int l = 0; l < n_dims; l++) w << num{P.pos(k, l)} << " ";
        if (n_dims == 2) w << "0 ";
      }
      w << endl << "				</DataArray>" << endl;
      w << "			</Points>" << endl;
      w << "			<Cells>" << endl;
      w << "				<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_cells; k++)
      {
        for (int l = 0; l < n_verts; l++) w << P.T.con[(size_t)k * n_verts + l] << " ";
        w << endl;
      }
      w << "				</DataArray>" << endl;
      w << "				<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_cells; k++) w << (k + 1) * n_verts << " ";
      w << endl << "				</DataArray>" << endl;
      w << "				<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">" << endl;
      for (int k = 0; k < n_cells; k++) w << vtktypes[t] << " ";
      w << endl << "				</DataArray>" << endl;
      w << "			</Cells>" << endl;
      w << "		</Piece>" << endl;
    }
  }
  w << "	</UnstructuredGrid>" << endl;
  w << "</VTKFile>" << endl;
  w.close();
  if (my_rank == 0) cout << "done." << endl;
}

// Tecplot output (reference output::write_tec, src/output.cpp:165-451): one finite-element zone per element type, point
// data = position, conservative variables, optional averages and diagnostic fields; connectivity one-based.
void write_tec(int in_file_num, struct solution *FlowSol)
{
  const int n_diag_fields = run_input.n_diagnostic_fields, n_average_fields = run_input.n_average_fields, n_dims = FlowSol->n_dims;
  char file_name_s[600], dumpnum_s[256];
  const char *name = run_input.data_file_name.c_str();
  if (FlowSol->nproc != 1)
  {
    snprintf(dumpnum_s, sizeof(dumpnum_s), "%s_%.09d", name, in_file_num);
    snprintf(file_name_s, sizeof(file_name_s), "%s/%s_p%.04d.plt", dumpnum_s, dumpnum_s, FlowSol->rank);
    prepare_directory(dumpnum_s, FlowSol);
  }
  else
    snprintf(file_name_s, sizeof(file_name_s), "%s_%.09d_p%.04d.plt", name, in_file_num, FlowSol->rank);
  if (FlowSol->rank == 0) cout << "Writing Tecplot file number " << in_file_num << " ...." << flush;
  ofstream w(file_name_s);
  if (!w) FatalError(string("cannot open ") + file_name_s);
  w << "Title = \"HiFiLES Solution\"" << endl;
  string fields = n_dims == 2 ? "Variables = \"x\", \"y\", \"rho\", \"mom_x\", \"mom_y\", \"rhoE\""
                              : "Variables = \"x\", \"y\", \"z\", \"rho\", \"mom_x\", \"mom_y\", \"mom_z\", \"rhoE\"";
  for (int m = 0; m < n_average_fields; m++) fields += ", \"" + run_input.average_fields(m) + "\"";
  for (int m = 0; m < n_diag_fields; m++) fields += ", \"" + run_input.diagnostic_fields(m) + "\"";
  w << fields << endl;
  static const char *zonetype[5] = {"FETRIANGLE", "FEQUADRILATERAL", "FETETRAHEDRON", "FEBRICK", "FEBRICK"};
  bool time_written = false;
  for (int t = 0; t < FlowSol->n_ele_types; t++)
  {
    eles *e = FlowSol->mesh_eles(t);
    const int n_eles = e->get_n_eles();
    if (n_eles == 0) continue;
    plot_fields P;
    P.setup(e, FlowSol, in_file_num);
    w << "ZONE N = " << n_eles * P.n_points << ", E = " << n_eles * P.n_cells << ", DATAPACKING = POINT, ZONETYPE = " << zonetype[e->get_ele_type()] << endl;
    if (!time_written)
    {
      w << "SolutionTime=" << num{FlowSol->time} << endl;
      time_written = true;
    }
    for (int j = 0; j < n_eles; j++)
    {
      P.eval(j);
      for (int k = 0; k < P.n_points; k++)
      {
        for (int l = 0; l < n_dims; l++) w << num{P.pos(k, l)} << " ";
        for (int l = 0; l < P.n_fields; l++)
        {
          if (std::isnan(P.u(k, l))) FatalError("Nan in tecplot file, exiting");
          w << num{P.u(k, l)} << " ";
        }
        for (int l = 0; l < n_average_fields; l++)
        {
          if (std::isnan(P.avg(k, l))) FatalError("Nan in tecplot file, exiting");
          w << num{P.avg(k, l)} << " ";
        }
        for (int l = 0; l < n_diag_fields; l++) w << num{P.diag(k, l)} << " ";
        w << endl;
      }
    }
    for (int j = 0; j < n_eles; j++)
      for (int k = 0; k < P.n_cells; k++)
      {
        for (int l = 0; l < P.n_verts; l++)
        {
          w << P.T.con[(size_t)k * P.n_verts + l] + j * P.n_points + 1;
          if (l != P.n_verts - 1) w << " ";
        }
        w << endl;
      }
  }
  w.close();
  if (FlowSol->rank == 0) cout << "done." << endl;
}

void write_plot(int in_file_num, struct solution *FlowSol)
{
  if (run_input.equation != 0)
  {
    // the reference's writers index momentum / energy fields the scalar test equation does not have (out-of-bounds reads of
    // disu_ppts_temp, src/output.cpp:730-765): there is no defined file to reproduce, so none is written
    if (FlowSol->rank == 0) cout << "Plot files are not written for the advection-diffusion test equation." << endl;
    return;
  }
  if (run_input.write_type == 0) write_vtu(in_file_num, FlowSol);
  else if (run_input.write_type == 1) write_tec(in_file_num, FlowSol);
  else FatalError("ERROR: Trying to write unrecognized file format ... ");
}
