// Geometry pre-processing: mesh -> element objects -> interface objects.  Ordering rules reproduced from the
// reference because they fix every index the device kernels see:
//   element-local ids  = order of appearance within each type                       (src/geometry.cpp:196-297)
//   cyclic pairing     = first later unmatched face with the same bc and centroid   (src/geometry.cpp:351-415)
//   interface ids      = face creation order within each face type                  (src/geometry.cpp:637-706)
//   partition faces    = per neighbour rank, lower rank's local order wins          (src/geometry.cpp:1132-1251)
// Difference in mechanism, not in result: the reference finds partition-face partners with MPI_Bcast of face
// centroids; here every rank owns the whole mesh file and the partition vector, so it rebuilds the (cheap)
// unmatched-face list of the other ranks locally instead of communicating.
#include "hifiles.h"
#include <algorithm>
#include <unordered_map>
#include <cstring>

using namespace std;

solution::solution()
{
  rank = 0;
  nproc = 1;
  time = 0.;
  n_ele_types = 0;
  n_dims = 0;
  num_cells_global = 0;
  ini_iter = 0;
  n_int_inter_types = n_bdy_inter_types = n_mpi_inter_types = n_mpi_inters = 0;
  ctx = nullptr;
  no_device = 0;
}

solution::~solution()
{
  if (ctx) hf_dev_destroy(ctx);
}

void SetInput(struct solution *FlowSol)
{
  // rank / nproc are set by the caller (one process per GPU) before this point; defaults are serial
  if (FlowSol->nproc < 1) FlowSol->nproc = 1;
}

// ---- global mesh (read once) and its restriction to one rank ------------------------------------------------------
namespace
{
struct rank_faces
{
  // unmatched faces of one rank that are partition faces, in unmatched-face order
  vector<int> face;            // local face index
  vector<double> centroid;     // [n][n_dims]
  vector<double> vert;         // [n][4][n_dims]
  vector<int> nv;
};

bool check_cyclic(const double *delta, const double *c0, const double *c1, double tol, int n_dims)
{
  if (n_dims == 3)
    return (fabs(fabs(c0[0] - c1[0]) - delta[0]) < tol && fabs(c0[1] - c1[1]) < tol && fabs(c0[2] - c1[2]) < tol) ||
           (fabs(c0[0] - c1[0]) < tol && fabs(fabs(c0[1] - c1[1]) - delta[1]) < tol && fabs(c0[2] - c1[2]) < tol) ||
           (fabs(c0[0] - c1[0]) < tol && fabs(c0[1] - c1[1]) < tol && fabs(fabs(c0[2] - c1[2]) - delta[2]) < tol);
  return (fabs(fabs(c0[0] - c1[0]) - delta[0]) < tol && fabs(c0[1] - c1[1]) < tol) ||
         (fabs(c0[0] - c1[0]) < tol && fabs(fabs(c0[1] - c1[1]) - delta[1]) < tol);
}

// vertex a of face 1 coincides with vertex b of face 2, directly or through one cyclic offset
bool vert_match(const double *x1, const double *x2, const double *delta, double tol, bool allow_zero)
{
  double d0 = fabs(x1[0] - x2[0]), d1 = fabs(x1[1] - x2[1]), d2 = fabs(x1[2] - x2[2]);
  if (allow_zero && d0 < tol && d1 < tol && d2 < tol) return true;
  return (fabs(d0 - delta[0]) < tol && d1 < tol && d2 < tol) || (d0 < tol && fabs(d1 - delta[1]) < tol && d2 < tol) ||
         (d0 < tol && d1 < tol && fabs(d2 - delta[2]) < tol);
}

// compare_cyclic_faces / compare_mpi_faces (reference src/geometry.cpp:1011-1104, 1342-1437): xv[k*3+m]
int compare_faces_xyz(const double *v1, const double *v2, int nv, const double *delta, double tol, int n_dims, bool allow_zero)
{
  if (n_dims == 2) return 0;
  if (nv == 4)
  {
    static const int cand[4] = {1, 3, 0, 2};
    for (int r = 0; r < 4; r++)
      if (vert_match(v1, v2 + 3 * cand[r], delta, tol, allow_zero)) return r;
  }
  else if (nv == 3)
  {
    static const int cand[3] = {0, 2, 1};
    for (int r = 0; r < 3; r++)
      if (vert_match(v1, v2 + 3 * cand[r], delta, tol, allow_zero)) return r;
  }
  else
    FatalError("ERROR: Haven't implemented this face type in compare_cyclic_face yet....");
  FatalError("Could not match vertices in compare faces");
  return -1;
}

void face_centroid(const mesh &m, int f, double *c)
{
  for (int d = 0; d < m.n_dims; d++) c[d] = 0.;
  for (int k = 0; k < m.f2nv(f); k++)
    for (int d = 0; d < m.n_dims; d++) c[d] += m.xv(m.f2v(f, k), d) / (double)m.f2nv(f);
}

// Local pairing of cyclic faces (reference src/geometry.cpp:351-415).  The reference scans all later unmatched
// faces linearly (O(n^2)); a spatial hash on the centroid finds the same first match.
int pair_cyclic_faces(mesh &m, const double *delta, double tol)
{
  int n_cyc_loc = 0;
  int nd = m.n_dims;
  int nu = m.n_unmatched_inters;
  vector<double> cen((size_t)nu * 3, 0.);
  for (int i = 0; i < nu; i++) face_centroid(m, m.unmatched_inters(i), &cen[3 * (size_t)i]);

  // hash grid with cell size >> tol; candidates are looked up at the three shifted positions
  double h = 1e-3;
  auto key = [&](const double *c) {
    long long a = llround(c[0] / h), b = llround(c[1] / h), cc = nd == 3 ? llround(c[2] / h) : 0;
    return (unsigned long long)(a * 73856093LL) ^ (unsigned long long)(b * 19349663LL) ^ (unsigned long long)(cc * 83492791LL);
  };
  unordered_map<unsigned long long, vector<int>> grid;
  bool use_grid = nu > 2000;
  if (use_grid)
    for (int i = 0; i < nu; i++) grid[key(&cen[3 * (size_t)i])].push_back(i);

  for (int i = 0; i < nu; i++)
  {
    int i1 = m.unmatched_inters(i);
    int bcid_f = m.bc_id(m.f2c(i1, 0), m.f2loc_f(i1, 0));
    if (bcid_f == -1 || bcid_f == -3) continue;
    if (run_input.bc_list[bcid_f].get_bc_flag() != CYCLIC) continue;
    const double *c0 = &cen[3 * (size_t)i];
    int found = -1;
    auto test = [&](int j) {
      if (j <= i) return false;
      int i2 = m.unmatched_inters(j);
      if (bcid_f != m.bc_id(m.f2c(i2, 0), m.f2loc_f(i2, 0)) || m.f2nv(i1) != m.f2nv(i2)) return false;
      return check_cyclic(delta, c0, &cen[3 * (size_t)j], tol, nd);
    };
    if (use_grid)
    {
      int best = -1;
      for (int d = 0; d < nd; d++)
        for (int s = -1; s <= 1; s += 2)
        {
          if (!std::isfinite(delta[d])) continue;
          double q[3] = {c0[0], c0[1], c0[2]};
          q[d] += s * delta[d];
          // probe the 3^nd neighbouring hash cells of the shifted point
          for (int a = -1; a <= 1; a++)
            for (int b = -1; b <= 1; b++)
              for (int c = (nd == 3 ? -1 : 0); c <= (nd == 3 ? 1 : 0); c++)
              {
                double qq[3] = {q[0] + a * h, q[1] + b * h, q[2] + c * h};
                auto it = grid.find(key(qq));
                if (it == grid.end()) continue;
                for (int j : it->second)
                  if (test(j) && (best < 0 || j < best)) best = j;
              }
        }
      found = best;
    }
    else
    {
      for (int j = i + 1; j < nu; j++)
        if (test(j)) { found = j; break; }
    }
    if (found >= 0)
    {
      int i2 = m.unmatched_inters(found);
      m.f2c(i1, 1) = m.f2c(i2, 0);
      m.bc_id(m.f2c(i1, 0), m.f2loc_f(i1, 0)) = -1;
      m.bc_id(m.f2c(i2, 0), m.f2loc_f(i2, 0)) = -3;
      m.f2loc_f(i1, 1) = m.f2loc_f(i2, 0);
      n_cyc_loc++;
      double v0[12], v1[12];
      memset(v0, 0, sizeof(v0));
      memset(v1, 0, sizeof(v1));
      for (int k = 0; k < m.f2nv(i1); k++)
        for (int d = 0; d < nd; d++)
        {
          v0[3 * k + d] = m.xv(m.f2v(i1, k), d);
          v1[3 * k + d] = m.xv(m.f2v(i2, k), d);
        }
      m.rot_tag(i1) = compare_faces_xyz(v0, v1, m.f2nv(i1), delta, tol, nd, false);
    }
    else
      m.bc_id(m.f2c(i1, 0), m.f2loc_f(i1, 0)) = -1; // partner lives on another rank (or does not exist)
  }
  return n_cyc_loc;
}

// list the partition faces of a rank-local mesh after cyclic pairing (reference src/geometry.cpp:431-455)
void list_partition_faces(mesh &m, rank_faces &rf, bool flag_them)
{
  int nd = m.n_dims;
  for (int i = 0; i < m.n_unmatched_inters; i++)
  {
    int i1 = m.unmatched_inters(i);
    int bcid_f = m.bc_id(m.f2c(i1, 0), m.f2loc_f(i1, 0));
    if (m.f2c(i1, 1) == -1 && bcid_f == -1)
    {
      if (flag_them) m.bc_id(m.f2c(i1, 0), m.f2loc_f(i1, 0)) = -2;
      rf.face.push_back(i1);
      double c[3] = {0, 0, 0};
      face_centroid(m, i1, c);
      for (int d = 0; d < 3; d++) rf.centroid.push_back(c[d]);
      rf.nv.push_back(m.f2nv(i1));
      for (int k = 0; k < 4; k++)
        for (int d = 0; d < 3; d++)
          rf.vert.push_back((k < m.f2nv(i1) && d < nd) ? m.xv(m.f2v(i1, k), d) : 0.);
    }
  }
}
} // namespace

// Build the mesh of one rank from the file + partition vector; stops after connectivity and boundary reading.
static void build_rank_mesh(struct solution *FlowSol, mesh &mesh_data, int rank)
{
  mesh_reader m_r(run_input.mesh_file, &mesh_data);
  m_r.partial_read_connectivity(0, mesh_data.num_cells_global);
  if (FlowSol->nproc > 1)
  {
    // no partition handed in: partition the dual graph as the reference's repartition_mesh does (METIS instead of ParMETIS)
    if (FlowSol->part.empty()) partition_mesh_kway(mesh_data, mesh_data.n_dims, FlowSol->nproc, FlowSol->part);
    if ((int)FlowSol->part.size() != mesh_data.num_cells_global)
      FatalError("partition vector of the wrong size");
    mesh_data.apply_partition(FlowSol->part, rank);
  }
  mesh_data.create_iv2ivg();
  m_r.read_vertices();
  mesh_data.set_vertex_connectivity();
  mesh_data.set_face_connectivity();
  m_r.read_boundary();
}

void ReadMesh(struct solution *FlowSol, mesh &mesh_data)
{
  {
    mesh probe;
    mesh_reader hdr(run_input.mesh_file, &probe);
    FlowSol->n_dims = probe.n_dims;
    FlowSol->num_cells_global = probe.num_cells_global;
  }
  build_rank_mesh(FlowSol, mesh_data, FlowSol->rank);
  run_input.read_boundary_param();
}

void GeoPreprocess(struct solution *FlowSol, mesh &mesh_data)
{
  ReadMesh(FlowSol, mesh_data);
  if (!FlowSol->ctx && !FlowSol->no_device) hf_check(hf_dev_create(&FlowSol->ctx, -1, FlowSol->rank, FlowSol->nproc));

  int num_tris = mesh_data.get_num_cells(TRI);
  int num_quads = mesh_data.get_num_cells(QUAD);
  int num_tets = mesh_data.get_num_cells(TET);
  int num_pris = mesh_data.get_num_cells(PRISM);
  int num_hexas = mesh_data.get_num_cells(HEX);
  if (FlowSol->n_dims == 2 && (num_tets != 0 || num_pris != 0 || num_hexas != 0))
    FatalError("Error in mesh reader, n_dims=2 and 3d elements exists");
  if (FlowSol->n_dims == 3 && (num_tris != 0 || num_quads != 0))
    FatalError("Error in mesh reader, n_dims=3 and 2d elements exists");

  FlowSol->n_ele_types = 5;
  FlowSol->mesh_eles.setup(5);
  FlowSol->mesh_eles(0) = &FlowSol->mesh_eles_tris;
  FlowSol->mesh_eles(1) = &FlowSol->mesh_eles_quads;
  FlowSol->mesh_eles(2) = &FlowSol->mesh_eles_tets;
  FlowSol->mesh_eles(3) = &FlowSol->mesh_eles_pris;
  FlowSol->mesh_eles(4) = &FlowSol->mesh_eles_hexas;
  for (int i = 0; i < 5; i++)
  {
    FlowSol->mesh_eles(i)->set_rank(FlowSol->rank);
    FlowSol->mesh_eles(i)->set_device(FlowSol->ctx);
  }
  const int counts[5] = {num_tris, num_quads, num_tets, num_pris, num_hexas};
  for (int t = 0; t < 5; t++) FlowSol->mesh_eles(t)->setup(counts[t], mesh_data.get_max_n_spts(t));

  // shape nodes, global ids, boundary ids; element-local id = order of appearance within the type
  vector<int> local_c(mesh_data.num_cells);
  int cnt[5] = {0, 0, 0, 0, 0};
  hf_array<double> pos(FlowSol->n_dims);
  static const int nfaces[5] = {3, 4, 4, 5, 6};
  for (int i = 0; i < mesh_data.num_cells; i++)
  {
    int t = mesh_data.ctype(i);
    eles *e = FlowSol->mesh_eles(t);
    int lc = cnt[t]++;
    local_c[i] = lc;
    e->set_n_spts(lc, mesh_data.c2n_v(i));
    e->set_ele2global_ele(lc, mesh_data.ic2icg(i));
    for (int j = 0; j < mesh_data.c2n_v(i); j++)
    {
      for (int d = 0; d < FlowSol->n_dims; d++) pos(d) = mesh_data.xv(mesh_data.c2v(i, j), d);
      e->set_shape_node(j, lc, pos);
    }
    for (int j = 0; j < nfaces[t]; j++) e->set_bcid(lc, j, mesh_data.bc_id(i, j));
  }

  for (int i = 0; i < 5; i++)
    if (FlowSol->mesh_eles(i)->get_n_eles() != 0) FlowSol->mesh_eles(i)->set_transforms();

  // ---- cyclic faces: interior or partition faces ----
  double delta_cyclic[3] = {run_input.dx_cyclic, run_input.dy_cyclic, FlowSol->n_dims == 3 ? run_input.dz_cyclic : 0.};
  double tol = 1.e-6;
  pair_cyclic_faces(mesh_data, delta_cyclic, tol);

  // ---- partition faces ----
  FlowSol->n_mpi_inter_types = 3;
  FlowSol->mesh_mpi_inters.assign(3, mpi_inters());
  for (int i = 0; i < 3; i++)
  {
    FlowSol->mesh_mpi_inters[i].set_nproc(FlowSol->nproc, FlowSol->rank);
    FlowSol->mesh_mpi_inters[i].set_device(FlowSol->ctx);
  }
  rank_faces mine;
  list_partition_faces(mesh_data, mine, true);
  FlowSol->n_mpi_inters = (int)mine.face.size();
  if (FlowSol->n_mpi_inters > 0 && FlowSol->nproc == 1)
    FatalError("Can't find coupled cyclic interface");
  if (FlowSol->nproc > 1)
  {
    int nd = FlowSol->n_dims;
    int n_mpi = FlowSol->n_mpi_inters;
    double delta_zero[3] = {0., 0., 0.};
    vector<int> matched(n_mpi, 0), f_mpi2f(n_mpi), rot_tag_mpi(n_mpi), partner_rank(n_mpi);
    vector<int> mpifaces_part(FlowSol->nproc, 0);
    vector<double> partner_vert((size_t)n_mpi * 12);
    int icount = 0;
    // hash of my faces by centroid for the p > rank branch (remote-major order)
    for (int p = 0; p < FlowSol->nproc; p++)
    {
      if (p == FlowSol->rank) continue;
      mesh other;
      build_rank_mesh(FlowSol, other, p);
      pair_cyclic_faces(other, delta_cyclic, tol);
      rank_faces theirs;
      list_partition_faces(other, theirs, false);
      int n_rem = (int)theirs.face.size();
      // spatial hash of the remote centroids (wrapped into the periodic box so cyclic partners collide)
      auto wrap_key = [&](const double *c) {
        long long k[3] = {0, 0, 0};
        for (int d = 0; d < nd; d++)
        {
          double x = c[d];
          if (std::isfinite(delta_cyclic[d]) && delta_cyclic[d] > 0) { x = fmod(x, delta_cyclic[d]); if (x < 0) x += delta_cyclic[d]; if (fabs(x - delta_cyclic[d]) < 1e-4) x = 0.; }
          k[d] = llround(x / 1e-3);
        }
        return (unsigned long long)(k[0] * 73856093LL) ^ (unsigned long long)(k[1] * 19349663LL) ^ (unsigned long long)(k[2] * 83492791LL);
      };
      auto is_match = [&](int irem, int iloc) {
        const double *c1 = &theirs.centroid[3 * (size_t)irem];
        const double *c2 = &mine.centroid[3 * (size_t)iloc];
        return check_cyclic(delta_cyclic, c1, c2, tol, nd) || check_cyclic(delta_zero, c1, c2, tol, nd);
      };
      auto probe = [&](unordered_map<unsigned long long, vector<int>> &g, const double *c, vector<int> &out) {
        out.clear();
        for (int a = -1; a <= 1; a++)
          for (int b = -1; b <= 1; b++)
            for (int cc = (nd == 3 ? -1 : 0); cc <= (nd == 3 ? 1 : 0); cc++)
            {
              double q[3] = {c[0] + a * 1e-3, c[1] + b * 1e-3, c[2] + cc * 1e-3};
              auto it = g.find(wrap_key(q));
              if (it != g.end()) out.insert(out.end(), it->second.begin(), it->second.end());
            }
        sort(out.begin(), out.end());
        out.erase(unique(out.begin(), out.end()), out.end());
      };
      vector<int> cand;
      if (p < FlowSol->rank)
      {
        // local-major: for each unmatched local face, first remote face that matches
        unordered_map<unsigned long long, vector<int>> g;
        for (int irem = 0; irem < n_rem; irem++) g[wrap_key(&theirs.centroid[3 * (size_t)irem])].push_back(irem);
        for (int iloc = 0; iloc < n_mpi; iloc++)
        {
          if (matched[iloc]) continue;
          probe(g, &mine.centroid[3 * (size_t)iloc], cand);
          for (int irem : cand)
            if (is_match(irem, iloc))
            {
              matched[iloc] = 1;
              mpifaces_part[p]++;
              f_mpi2f[icount] = iloc;
              partner_rank[icount] = p;
              memcpy(&partner_vert[12 * (size_t)icount], &theirs.vert[12 * (size_t)irem], 12 * sizeof(double));
              icount++;
              break;
            }
        }
      }
      else
      {
        // remote-major: for each remote face, first unmatched local face that matches
        unordered_map<unsigned long long, vector<int>> g;
        for (int iloc = 0; iloc < n_mpi; iloc++) g[wrap_key(&mine.centroid[3 * (size_t)iloc])].push_back(iloc);
        for (int irem = 0; irem < n_rem; irem++)
        {
          probe(g, &theirs.centroid[3 * (size_t)irem], cand);
          for (int iloc : cand)
            if (!matched[iloc] && is_match(irem, iloc))
            {
              matched[iloc] = 1;
              mpifaces_part[p]++;
              f_mpi2f[icount] = iloc;
              partner_rank[icount] = p;
              memcpy(&partner_vert[12 * (size_t)icount], &theirs.vert[12 * (size_t)irem], 12 * sizeof(double));
              icount++;
              break;
            }
        }
      }
    }
    for (int i = 0; i < n_mpi; i++)
      if (!matched[i]) FatalError("Some mpi_faces were not matched!!! could try changing tol, exiting!");

    // The remote rank lists the faces it shares with me in the same relative order (the ordering rule above is
    // symmetric), so the k-th face I send to p pairs with the k-th face p sends to me.  partner_vert was taken
    // from the remote face that matched by centroid, which is that same face.
    int n_seg = 0, n_tri = 0, n_quad = 0;
    for (int k = 0; k < n_mpi; k++)
    {
      int nv = mine.nv[f_mpi2f[k]];
      if (nv == 2) n_seg++; else if (nv == 3) n_tri++; else n_quad++;
    }
    FlowSol->mesh_mpi_inters[0].setup(n_seg, 0);
    FlowSol->mesh_mpi_inters[1].setup(n_tri, 1);
    FlowSol->mesh_mpi_inters[2].setup(n_quad, 2);
    int ii[3] = {0, 0, 0};
    for (int k = 0; k < n_mpi; k++)
    {
      int iloc = f_mpi2f[k];
      int f = mine.face[iloc];
      int nv = mine.nv[iloc];
      int rtag = compare_faces_xyz(&mine.vert[12 * (size_t)iloc], &partner_vert[12 * (size_t)k], nv, delta_cyclic, tol, nd, true);
      int ic_l = mesh_data.f2c(f, 0);
      int t = nv == 2 ? 0 : nv == 3 ? 1 : 2;
      FlowSol->mesh_mpi_inters[t].set_mpi(ii[t]++, mesh_data.ctype(ic_l), local_c[ic_l], mesh_data.f2loc_f(f, 0), rtag, FlowSol);
    }
    int start = 0;
    for (int p = 0; p < FlowSol->nproc; p++)
    {
      int nout[3] = {0, 0, 0};
      for (int j = 0; j < mpifaces_part[p]; j++)
      {
        int nv = mine.nv[f_mpi2f[start + j]];
        nout[nv == 2 ? 0 : nv == 3 ? 1 : 2]++;
      }
      start += mpifaces_part[p];
      for (int t = 0; t < 3; t++)
        if (nout[t]) FlowSol->mesh_mpi_inters[t].set_nout_proc(nout[t], p);
    }
  }
  else
  {
    for (int t = 0; t < 3; t++) FlowSol->mesh_mpi_inters[t].setup(0, t);
  }

  // ---- interior and boundary interfaces, in face creation order ----
  int n_int[3] = {0, 0, 0}, n_bdy[3] = {0, 0, 0};
  for (int i = 0; i < mesh_data.num_inters; i++)
  {
    int bcid_f = mesh_data.bc_id(mesh_data.f2c(i, 0), mesh_data.f2loc_f(i, 0));
    int ic_r = mesh_data.f2c(i, 1);
    int t = mesh_data.f2nv(i) - 2;
    if (bcid_f == -2) continue;
    if (bcid_f == -1)
    {
      if (ic_r == -1) FatalError("Error: Interior interface has i_cell_right=-1. Should not be here, exiting");
      n_int[t]++;
    }
    else if (bcid_f != -3)
      n_bdy[t]++;
  }
  FlowSol->n_int_inter_types = 3;
  FlowSol->n_bdy_inter_types = 3;
  FlowSol->mesh_int_inters.assign(3, int_inters());
  FlowSol->mesh_bdy_inters.assign(3, bdy_inters());
  for (int t = 0; t < 3; t++)
  {
    FlowSol->mesh_int_inters[t].set_device(FlowSol->ctx);
    FlowSol->mesh_bdy_inters[t].set_device(FlowSol->ctx);
    FlowSol->mesh_int_inters[t].setup(n_int[t], t);
    FlowSol->mesh_bdy_inters[t].setup(n_bdy[t], t);
  }
  int i_int[3] = {0, 0, 0}, i_bdy[3] = {0, 0, 0};
  for (int i = 0; i < mesh_data.num_inters; i++)
  {
    int ic_l = mesh_data.f2c(i, 0);
    int ic_r = mesh_data.f2c(i, 1);
    int bcid_f = mesh_data.bc_id(ic_l, mesh_data.f2loc_f(i, 0));
    int t = mesh_data.f2nv(i) - 2;
    if (bcid_f == -2) continue;
    if (bcid_f == -1)
      FlowSol->mesh_int_inters[t].set_interior(i_int[t]++, mesh_data.ctype(ic_l), mesh_data.ctype(ic_r), local_c[ic_l], local_c[ic_r],
                                               mesh_data.f2loc_f(i, 0), mesh_data.f2loc_f(i, 1), mesh_data.rot_tag(i), FlowSol);
    else if (bcid_f != -3)
      FlowSol->mesh_bdy_inters[t].set_boundary(i_bdy[t]++, bcid_f, mesh_data.ctype(ic_l), local_c[ic_l], mesh_data.f2loc_f(i, 0), FlowSol);
  }

  // wall distance for the Smagorinsky near-wall damping (reference src/geometry.cpp:706-892): flux points of every
  // isothermal / adiabatic wall face, per face type, in mesh-face order
  if (run_input.LES && run_input.SGS_model == 0)
  {
    vector<hf_array<double>> loc_noslip_bdy(3);
    for (int t = 0; t < 3; t++)
    {
      bdy_inters &B = FlowSol->mesh_bdy_inters[t];
      vector<int> walls;
      for (int q = 0; q < B.get_n_inters(); q++)
      {
        int flag = run_input.bc_list[B.boundary_id(q)].get_bc_flag();
        if (flag == ISOTHERM_WALL || flag == ADIABAT_WALL) walls.push_back(q);
      }
      const int nfp = t == 0 ? run_input.order + 1 : (t == 1 ? (run_input.order + 2) * (run_input.order + 1) / 2 : (run_input.order + 1) * (run_input.order + 1));
      loc_noslip_bdy[t].setup(FlowSol->n_dims, nfp, max((int)walls.size(), 1));
      loc_noslip_bdy[t].setup(FlowSol->n_dims, nfp, (int)walls.size());
      for (size_t w = 0; w < walls.size(); w++)
        for (int j = 0; j < nfp; j++)
          for (int k = 0; k < FlowSol->n_dims; k++) loc_noslip_bdy[t](k, j, (int)w) = B.pos_fpts(j, walls[w], k);
    }
    // several ranks: the distance to this rank's walls for now; the other ranks' wall points arrive with the communicator (FinishWallDistance)
    if (FlowSol->nproc > 1)
    {
      FlowSol->loc_noslip_bdy_local = loc_noslip_bdy;
      FlowSol->wall_distance_pending = true;
    }
    for (int i = 0; i < FlowSol->n_ele_types; i++) FlowSol->mesh_eles(i)->calc_wall_distance(loc_noslip_bdy);
  }
}

// The wall points of every rank, rank by rank (reference src/geometry.cpp:768-880: MPI_Allgather of the counts, one MPI_Bcast per rank into a
// global array): every rank places its block into a zeroed global array and the arrays are summed over the communicator (x + 0 + ... + 0 is
// exact), in pieces of the reduction's staging buffer.  Then eles::calc_wall_distance again (src/geometry.cpp:884-892) and the device copy.
void FinishWallDistance(struct solution *FlowSol)
{
  if (!FlowSol->wall_distance_pending) return;
  if (!FlowSol->ctx) FatalError("the Smagorinsky wall distance of a partitioned run needs the device context's communicator");
  vector<hf_array<double>> glob(3);
  for (int t = 0; t < 3; t++)
  {
    hf_array<double> &mine = FlowSol->loc_noslip_bdy_local[t];
    const int nfp = t == 0 ? run_input.order + 1 : (t == 1 ? (run_input.order + 2) * (run_input.order + 1) / 2 : (run_input.order + 1) * (run_input.order + 1));
    const size_t per = (size_t)FlowSol->n_dims * nfp;
    const size_t n_mine = mine.size() / per;
    vector<double> cnt(FlowSol->nproc, 0.);
    cnt[FlowSol->rank] = (double)n_mine;
    hf_check(hf_dev_allreduce_sum(FlowSol->ctx, cnt.data(), FlowSol->nproc));
    size_t kstart = 0, total = 0;
    for (int p = 0; p < FlowSol->nproc; p++)
    {
      if (p == FlowSol->rank) kstart = total;
      total += (size_t)llround(cnt[p]);
    }
    glob[t].setup(FlowSol->n_dims, nfp, (int)total);
    if (total == 0) continue;
    double *g = glob[t].get_ptr_cpu();
    for (size_t q = 0; q < per * total; q++) g[q] = 0.;
    if (n_mine) memcpy(g + per * kstart, mine.get_ptr_cpu(), per * n_mine * sizeof(double));
    const size_t piece = 65536;
    for (size_t off = 0; off < per * total; off += piece) hf_check(hf_dev_allreduce_sum(FlowSol->ctx, g + off, (int)min(piece, per * total - off)));
  }
  for (int i = 0; i < FlowSol->n_ele_types; i++)
  {
    eles *e = FlowSol->mesh_eles(i);
    if (e->get_n_eles() == 0) continue;
    e->calc_wall_distance(glob);
    hf_check(hf_dev_set_wall_distance(FlowSol->ctx, e->get_ele_type(), e->wall_distance.get_ptr_cpu(), e->wall_distance.size()));
  }
  FlowSol->loc_noslip_bdy_local.clear();
  FlowSol->wall_distance_pending = false;
}
