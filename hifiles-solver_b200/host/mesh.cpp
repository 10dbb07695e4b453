// Mesh reading and face connectivity.  Numbering must equal the reference's bit for bit, because it fixes the
// interface order and therefore the flux-point pair indexing the device kernels use:
//   cells   : file order within the rank's block                      (reference src/mesh_reader.cpp:132-262)
//   faces   : loop cells, loop local faces, first touch creates the face (reference src/mesh.cpp:375-485)
//   rot_tag : which corner of face 2 coincides with corner 0 of face 1  (reference src/mesh.cpp:853-952)
#include "hifiles.h"
#include <algorithm>
#include <cstring>
#include <cstdio>

using namespace std;

static const int k_num_f_per_c[5] = {3, 4, 4, 5, 6};

mesh::mesh()
{
  n_dims = n_ele_dims = n_bdy = 0;
  num_verts_global = num_cells_global = num_verts = num_cells = num_inters = n_unmatched_inters = 0;
}

int mesh::get_num_cells(int in_type) const
{
  int c = 0;
  for (int i = 0; i < num_cells; i++)
    if (ctype(i) == in_type) c++;
  return c;
}

int mesh::get_max_n_spts(int in_type) const
{
  int m = 0;
  for (int i = 0; i < num_cells; i++)
    if (ctype(i) == in_type && c2n_v(i) > m) m = c2n_v(i);
  return m;
}

void mesh::apply_partition(const vector<int> &part, int rank)
{
  // reference src/mesh.cpp:188-311 migrates cells after ParMETIS and concatenates the receives in source-rank
  // order, which for a block-distributed initial read leaves each rank's cells in ascending global id.
  vector<int> keep;
  for (int i = 0; i < num_cells; i++)
    if (part[ic2icg(i)] == rank) keep.push_back(i);
  hf_array<int> c2v_n((int)keep.size(), MAX_V_PER_C), c2n_v_n((int)keep.size()), ctype_n((int)keep.size()), ic2icg_n((int)keep.size());
  for (size_t k = 0; k < keep.size(); k++)
  {
    int i = keep[k];
    for (int j = 0; j < MAX_V_PER_C; j++) c2v_n((int)k, j) = c2v(i, j);
    c2n_v_n((int)k) = c2n_v(i);
    ctype_n((int)k) = ctype(i);
    ic2icg_n((int)k) = ic2icg(i);
  }
  c2v = c2v_n; c2n_v = c2n_v_n; ctype = ctype_n; ic2icg = ic2icg_n;
  num_cells = (int)keep.size();
}

void mesh::create_iv2ivg()
{
  vector<int> vrtlist;
  vrtlist.reserve((size_t)num_cells * 8);
  for (int i = 0; i < num_cells; i++)
    for (int j = 0; j < MAX_V_PER_C; j++)
      if (c2v(i, j) != -1) vrtlist.push_back(c2v(i, j));
  sort(vrtlist.begin(), vrtlist.end());
  vrtlist.erase(unique(vrtlist.begin(), vrtlist.end()), vrtlist.end());
  num_verts = (int)vrtlist.size();
  iv2ivg.setup(num_verts);
  copy(vrtlist.begin(), vrtlist.end(), iv2ivg.get_ptr_cpu());
  // The reference renumbers c2v to local vertex ids only in its MPI build (src/mesh.cpp:337-356); in a serial
  // read every vertex is used, so local == global there.  Renumbering always is equivalent and also covers
  // partitions.
  for (int i = 0; i < num_cells; i++)
    for (int j = 0; j < c2n_v(i); j++)
    {
      int *b = iv2ivg.get_ptr_cpu();
      int *p = lower_bound(b, b + num_verts, c2v(i, j));
      if (p == b + num_verts || *p != c2v(i, j)) FatalError("Could not find value in index_locate");
      c2v(i, j) = (int)(p - b);
    }
}

void mesh::set_vertex_connectivity()
{
  v2c.assign(num_verts, vector<int>());
  for (int ic = 0; ic < num_cells; ic++)
    for (int k = 0; k < c2n_v(ic); k++)
      v2c[c2v(ic, k)].push_back(ic);
}

int mesh::get_corner_vlist_face(int in_ic, int in_face, int *v) const
{
  int nv = 0;
  int ns = c2n_v(in_ic);
  int ct = ctype(in_ic);
  if (ct == TRI)
  {
    nv = 2;
    v[0] = in_face; v[1] = (in_face + 1) % 3;
  }
  else if (ct == QUAD)
  {
    nv = 2;
    if (is_perfect_square(ns))
    {
      int n1 = (int)lround(sqrt((double)ns));
      const int c[4] = {0, n1 - 1, ns - 1, ns - n1};
      v[0] = c[in_face]; v[1] = c[(in_face + 1) % 4];
    }
    else if (ns == 8) { v[0] = in_face; v[1] = (in_face + 1) % 4; }
    else FatalError("in_nspt not implemented");
  }
  else if (ct == TET)
  {
    nv = 3;
    static const int t[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
    for (int i = 0; i < 3; i++) v[i] = t[in_face][i];
  }
  else if (ct == PRISM)
  {
    static const int p[5][4] = {{0, 2, 1, -1}, {3, 4, 5, -1}, {0, 1, 4, 3}, {1, 2, 5, 4}, {2, 0, 3, 5}};
    nv = (in_face < 2) ? 3 : 4;
    for (int i = 0; i < nv; i++) v[i] = p[in_face][i];
  }
  else if (ct == HEX)
  {
    nv = 4;
    if (is_perfect_cube(ns))
    {
      int n1 = (int)lround(pow((double)ns, 1. / 3.));
      int shift = n1 * n1 * (n1 - 1);
      // the eight corners in the reference's corner order (src/mesh.cpp:536-574)
      const int c[8] = {0, n1 - 1, n1 * n1 - 1, n1 * (n1 - 1), shift, n1 - 1 + shift, ns - 1, ns - n1};
      static const int h[6][4] = {{1, 0, 3, 2}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
      for (int i = 0; i < 4; i++) v[i] = c[h[in_face][i]];
    }
    else if (ns == 20)
    {
      static const int h[6][4] = {{1, 0, 3, 2}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
      for (int i = 0; i < 4; i++) v[i] = h[in_face][i];
    }
    else FatalError("n_spts not implemented");
  }
  else
    FatalError("ERROR: Haven't implemented other 3D Elements yet");
  for (int i = 0; i < nv; i++) v[i] = c2v(in_ic, v[i]);
  return nv;
}

int mesh::compare_faces(const int *a, const int *b, int nv, int &rtag)
{
  if (nv == 2)
  {
    if ((a[0] == b[0] && a[1] == b[1]) || (a[0] == b[1] && a[1] == b[0])) { rtag = 0; return 1; }
    return 0;
  }
  if (nv == 3)
  {
    if (a[0] == b[0] && a[1] == b[2] && a[2] == b[1]) { rtag = 0; return 1; }
    if (a[0] == b[2] && a[1] == b[1] && a[2] == b[0]) { rtag = 1; return 1; }
    if (a[0] == b[1] && a[1] == b[0] && a[2] == b[2]) { rtag = 2; return 1; }
    return 0;
  }
  if (nv == 4)
  {
    if (a[0] == b[1] && a[1] == b[0] && a[2] == b[3] && a[3] == b[2]) { rtag = 0; return 1; }
    if (a[0] == b[3] && a[1] == b[2] && a[2] == b[1] && a[3] == b[0]) { rtag = 1; return 1; }
    if (a[0] == b[0] && a[1] == b[3] && a[2] == b[2] && a[3] == b[1]) { rtag = 2; return 1; }
    if (a[0] == b[2] && a[1] == b[1] && a[2] == b[0] && a[3] == b[3]) { rtag = 3; return 1; }
    return 0;
  }
  FatalError("ERROR: Haven't implemented this face type in compare_face yet....");
  return 0;
}

void mesh::set_face_connectivity()
{
  int max_inters = num_cells * MAX_F_PER_C;
  f2c.setup(max_inters, 2);
  f2v.setup(max_inters, MAX_V_PER_F);
  f2nv.setup(max_inters);
  f2loc_f.setup(max_inters, 2);
  c2f.setup(num_cells, MAX_F_PER_C);
  rot_tag.setup(max_inters);
  unmatched_inters.setup(max_inters);
  f2c.initialize_to_value(-1);
  f2loc_f.initialize_to_value(-1);
  c2f.initialize_to_value(-1);
  num_inters = 0;
  n_unmatched_inters = 0;

  int vlist[4], vlist2[4];
  vector<int> inter, tmp;
  for (int ic = 0; ic < num_cells; ic++)
  {
    for (int k = 0; k < k_num_f_per_c[ctype(ic)]; k++)
    {
      if (c2f(ic, k) != -1) continue;
      int nv = get_corner_vlist_face(ic, k, vlist);
      inter = v2c[vlist[0]];
      for (int i = 1; i < nv; i++)
      {
        tmp.clear();
        set_intersection(v2c[vlist[i]].begin(), v2c[vlist[i]].end(), inter.begin(), inter.end(), back_inserter(tmp));
        inter.swap(tmp);
      }
      if (inter.size() == 2)
      {
        for (size_t q = 0; q < inter.size(); q++)
        {
          int ic2 = inter[q];
          if (ic2 == ic) continue;
          for (int k2 = 0; k2 < k_num_f_per_c[ctype(ic2)]; k2++)
          {
            int nv2 = get_corner_vlist_face(ic2, k2, vlist2);
            if (nv2 != nv) continue;
            int rtag;
            if (compare_faces(vlist, vlist2, nv, rtag))
            {
              c2f(ic, k) = num_inters;
              c2f(ic2, k2) = num_inters;
              f2c(num_inters, 0) = ic;
              f2c(num_inters, 1) = ic2;
              f2loc_f(num_inters, 0) = k;
              f2loc_f(num_inters, 1) = k2;
              for (int i = 0; i < nv; i++) f2v(num_inters, i) = vlist[i];
              f2nv(num_inters) = nv;
              rot_tag(num_inters) = rtag;
              num_inters++;
              break;
            }
          }
        }
      }
      else if (inter.size() == 1)
      {
        f2c(num_inters, 0) = ic;
        f2c(num_inters, 1) = -1;
        f2loc_f(num_inters, 0) = k;
        f2loc_f(num_inters, 1) = -1;
        c2f(ic, k) = num_inters;
        for (int i = 0; i < nv; i++) f2v(num_inters, i) = vlist[i];
        f2nv(num_inters) = nv;
        unmatched_inters(n_unmatched_inters) = num_inters;
        n_unmatched_inters++;
        num_inters++;
      }
      else
        FatalError("More than two cells share one face");
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// mesh_reader
// ---------------------------------------------------------------------------------------------------------
mesh_reader::mesh_reader(const string &in_fileName, mesh *in_mesh)
{
  fname = in_fileName;
  mesh_ptr = in_mesh;
  gmsh_elements_block_start = 0;
  if (fname.size() >= 3 && !fname.compare(fname.size() - 3, 3, "neu"))
    mesh_format = 0;
  else if (fname.size() >= 3 && !fname.compare(fname.size() - 3, 3, "msh"))
    mesh_format = 1;
  else
    FatalError("Mesh format not recognized");
  if (mesh_format == 0) read_header_gambit();
  else read_header_gmsh();
}

void mesh_reader::partial_read_connectivity(int kstart, int in_num_cells)
{
  if (kstart >= mesh_ptr->num_cells_global || in_num_cells > (mesh_ptr->num_cells_global - kstart))
    FatalError("Illegal block of elements to read");
  mesh_ptr->num_cells = in_num_cells;
  if (mesh_format == 0) partial_read_connectivity_gambit(kstart, in_num_cells);
  else partial_read_connectivity_gmsh(kstart, in_num_cells);
}
void mesh_reader::read_vertices()
{
  if (mesh_format == 0) read_vertices_gambit();
  else read_vertices_gmsh();
}
void mesh_reader::read_boundary()
{
  if (mesh_format == 0) read_boundary_gambit();
  else read_boundary_gmsh();
}

// whole-file tokenizer: Gambit/Gmsh meshes at 64^3 are tens of MB of text; iostream extraction per token is
// the start-up bottleneck in the reference, so read the file once and scan with strtol/strtod.
namespace
{
struct text_file
{
  vector<char> buf;
  char *p, *end;
  explicit text_file(const string &name)
  {
    FILE *f = fopen(name.c_str(), "rb");
    if (!f) FatalError("Unable to open mesh file");
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)n + 1);
    size_t got = fread(buf.data(), 1, (size_t)n, f);
    fclose(f);
    buf[got] = 0;
    p = buf.data();
    end = p + got;
  }
  bool skip_to_line_containing(const char *key)
  {
    while (p < end)
    {
      char *e = (char *)memchr(p, '\n', end - p);
      if (!e) e = end;
      char save = *e;
      *e = 0;
      bool hit = strstr(p, key) != nullptr;
      *e = save;
      p = (e < end) ? e + 1 : end;
      if (hit) return true;
    }
    return false;
  }
  void skip_line()
  {
    char *e = (char *)memchr(p, '\n', end - p);
    p = e ? e + 1 : end;
  }
  long next_int()
  {
    char *q;
    long v = strtol(p, &q, 10);
    if (q == p) FatalError("mesh file: integer expected");
    p = q;
    return v;
  }
  double next_double()
  {
    char *q;
    double v = strtod(p, &q);
    if (q == p) FatalError("mesh file: number expected");
    p = q;
    return v;
  }
  string next_word()
  {
    while (p < end && isspace((unsigned char)*p)) p++;
    char *s = p;
    while (p < end && !isspace((unsigned char)*p)) p++;
    return string(s, p);
  }
};
} // namespace

void mesh_reader::read_header_gambit()
{
  text_file f(fname);
  for (int i = 0; i < 6; i++) f.skip_line();
  mesh_ptr->num_verts_global = (int)f.next_int();
  mesh_ptr->num_cells_global = (int)f.next_int();
  (void)f.next_int();
  mesh_ptr->n_bdy = (int)f.next_int();
  mesh_ptr->n_ele_dims = (int)f.next_int();
  mesh_ptr->n_dims = (int)f.next_int();
  if (mesh_ptr->n_dims != 2 && mesh_ptr->n_dims != 3)
    FatalError("Invalid mesh dimensionality. Expected 2D or 3D.");
}

void mesh_reader::partial_read_connectivity_gambit(int kstart, int in_num_cells)
{
  text_file f(fname);
  if (!f.skip_to_line_containing("ELEMENTS/CELLS")) FatalError("ELEMENTS/CELLS section not found");
  mesh *m = mesh_ptr;
  m->c2v.setup(in_num_cells, MAX_V_PER_C);
  m->c2n_v.setup(in_num_cells);
  m->ctype.setup(in_num_cells);
  m->ic2icg.setup(in_num_cells);
  m->c2v.initialize_to_value(-1);

  // file position -> c2v slot, per Gambit element kind (reference src/mesh_reader.cpp:192-246)
  static const int tri3[3] = {0, 1, 2}, tri6[6] = {0, 3, 1, 4, 2, 5};
  static const int quad4[4] = {0, 1, 3, 2}, quad8[8] = {0, 4, 1, 5, 2, 6, 3, 7};
  static const int tet4[4] = {0, 1, 2, 3}, tet10[10] = {0, 4, 1, 5, 7, 2, 6, 9, 8, 3};
  static const int pri6[6] = {0, 1, 2, 3, 4, 5}, pri15[15] = {0, 6, 1, 8, 7, 2, 9, 10, 11, 3, 12, 4, 14, 13, 5};
  static const int hex8[8] = {0, 2, 4, 6, 1, 3, 5, 7};
  static const int hex20[20] = {0, 11, 3, 12, 15, 4, 19, 7, 8, 10, 16, 18, 1, 9, 2, 13, 14, 5, 17, 6};

  for (int i = 0; i < kstart + in_num_cells; i++)
  {
    int id = (int)f.next_int();
    int eleType = (int)f.next_int();
    int nv = (int)f.next_int();
    if (i < kstart)
    {
      for (int k = 0; k < nv; k++) (void)f.next_int();
      continue;
    }
    int c = i - kstart;
    m->ic2icg(c) = id;
    m->c2n_v(c) = nv;
    const int *map = nullptr;
    if (eleType == 3) { m->ctype(c) = TRI; map = nv == 3 ? tri3 : nv == 6 ? tri6 : nullptr; if (!map) FatalError("triangle element type not implemented"); }
    else if (eleType == 2) { m->ctype(c) = QUAD; map = nv == 4 ? quad4 : nv == 8 ? quad8 : nullptr; if (!map) FatalError("quad element type not implemented"); }
    else if (eleType == 6) { m->ctype(c) = TET; map = nv == 4 ? tet4 : nv == 10 ? tet10 : nullptr; if (!map) FatalError("tet element type not implemented"); }
    else if (eleType == 5) { m->ctype(c) = PRISM; map = nv == 6 ? pri6 : nv == 15 ? pri15 : nullptr; if (!map) FatalError("Prism element type not implemented"); }
    else if (eleType == 4) { m->ctype(c) = HEX; map = nv == 8 ? hex8 : nv == 20 ? hex20 : nullptr; if (!map) FatalError("Hexa element type not implemented"); }
    else FatalError("Haven't implemented this element type in gambit_meshreader3, exiting ");
    for (int k = 0; k < nv; k++) m->c2v(c, map[k]) = (int)f.next_int() - 1;
    m->ic2icg(c)--;
  }
}

void mesh_reader::read_vertices_gambit()
{
  text_file f(fname);
  if (!f.skip_to_line_containing("NODAL COORDINATES")) FatalError("NODAL COORDINATES section not found");
  mesh *m = mesh_ptr;
  m->xv.setup(m->num_verts, m->n_dims);
  const int *b = m->iv2ivg.get_ptr_cpu();
  for (int i = 0; i < m->num_verts_global; i++)
  {
    int id = (int)f.next_int();
    const int *q = lower_bound(b, b + m->num_verts, id - 1);
    if (q != b + m->num_verts && *q == id - 1)
    {
      int index = (int)(q - b);
      for (int d = 0; d < m->n_dims; d++) m->xv(index, d) = f.next_double();
    }
    f.skip_line();
  }
}

void mesh_reader::read_boundary_gambit()
{
  text_file f(fname);
  mesh *m = mesh_ptr;
  m->bc_id.setup(m->num_cells, MAX_F_PER_C);
  m->bc_id.initialize_to_value(-1);
  // (re)create the bc list only when it does not exist yet: partition setup re-reads the boundary section for the
  // other ranks' meshes after the per-boundary parameters have been parsed
  if ((int)run_input.bc_list.size() != m->n_bdy) run_input.bc_list.assign(m->n_bdy, bc());
  const int *b = m->ic2icg.get_ptr_cpu();
  bool sorted = is_sorted(b, b + m->num_cells);
  map<int, int> g2l;
  if (!sorted)
    for (int i = 0; i < m->num_cells; i++) g2l[b[i]] = i;

  static const int hexf[7] = {-1, 0, 3, 5, 1, 4, 2};
  static const int tetf[5] = {-1, 3, 2, 0, 1};
  static const int prif[6] = {-1, 2, 3, 4, 0, 1};
  for (int i = 0; i < m->n_bdy; i++)
  {
    if (!f.skip_to_line_containing("BOUNDARY CONDITIONS")) FatalError("BOUNDARY CONDITIONS section not found");
    string bcname = f.next_word();
    (void)f.next_int();
    int bcNF = (int)f.next_int();
    if (run_input.bc_list[i].get_bc_name() != bcname) { run_input.bc_list[i] = bc(); run_input.bc_list[i].setup(bcname); }
    f.skip_line();
    for (int bf = 0; bf < bcNF; bf++)
    {
      int icg = (int)f.next_int() - 1;
      int eleType = (int)f.next_int();
      int k = (int)f.next_int();
      int real_face;
      if (eleType == 2 || eleType == 3) real_face = k - 1;
      else if (eleType == 4) real_face = hexf[k];
      else if (eleType == 6) real_face = tetf[k];
      else if (eleType == 5) real_face = prif[k];
      else { FatalError("Cannot handle other element type in readbnd"); real_face = -1; }
      int cellID = -1;
      if (sorted)
      {
        const int *q = lower_bound(b, b + m->num_cells, icg);
        if (q != b + m->num_cells && *q == icg) cellID = (int)(q - b);
      }
      else
      {
        auto it = g2l.find(icg);
        if (it != g2l.end()) cellID = it->second;
      }
      if (cellID != -1) m->bc_id(cellID, real_face) = i;
    }
  }
}

// ---- Gmsh 2.2 (reference src/mesh_reader.cpp:397-890) -------------------------------------------------------
static string strip_quotes(string s)
{
  size_t a = s.find_first_not_of("\" \t\r\n");
  size_t b = s.find_last_not_of("\" \t\r\n");
  if (a == string::npos) return "";
  return s.substr(a, b - a + 1);
}

void mesh_reader::read_header_gmsh()
{
  text_file f(fname);
  if (!f.skip_to_line_containing("$PhysicalNames")) FatalError("$PhysicalNames tag not found!");
  int n_groups = (int)f.next_int();
  mesh_ptr->n_bdy = n_groups - 1;
  int fluid_id = -1;
  gmsh_bc_names.clear();
  gmsh_phys2bc.clear();
  for (int i = 0; i < n_groups; i++)
  {
    int dim = (int)f.next_int();
    int id = (int)f.next_int();
    string name = strip_quotes(f.next_word());
    f.skip_line();
    if (name == "FLUID")
    {
      fluid_id = id;
      mesh_ptr->n_dims = dim;
      mesh_ptr->n_ele_dims = dim;
    }
    else
    {
      gmsh_phys2bc[id] = (int)gmsh_bc_names.size();
      gmsh_bc_names.push_back(name);
    }
  }
  if (fluid_id < 0) FatalError("Cant find fluid group in mesh file");
  if (mesh_ptr->n_dims != 2 && mesh_ptr->n_dims != 3)
    FatalError("Invalid mesh dimensionality. Expected 2D or 3D.");
  gmsh_phys2bc[-fluid_id - 1000000] = -1; // remember the fluid id under a key no physical group can take
  text_file g(fname);
  if (!g.skip_to_line_containing("$Nodes")) FatalError("$Nodes tag not found!");
  mesh_ptr->num_verts_global = (int)g.next_int();
  if (!g.skip_to_line_containing("$Elements")) FatalError("$Elements tag not found!");
  int n_entities = (int)g.next_int();
  g.skip_line();
  int icount = 0;
  for (int i = 0; i < n_entities; i++)
  {
    (void)g.next_int(); (void)g.next_int(); (void)g.next_int();
    int phys = (int)g.next_int();
    if (phys == fluid_id) icount++;
    g.skip_line();
  }
  mesh_ptr->num_cells_global = icount;
  gmsh_elements_block_start = fluid_id;
}

void mesh_reader::partial_read_connectivity_gmsh(int kstart, int in_num_cells)
{
  mesh *m = mesh_ptr;
  m->c2v.setup(in_num_cells, MAX_V_PER_C);
  m->c2n_v.setup(in_num_cells);
  m->ctype.setup(in_num_cells);
  m->ic2icg.setup(in_num_cells);
  m->c2v.initialize_to_value(-1);
  int fluid_id = gmsh_elements_block_start;
  text_file f(fname);
  if (!f.skip_to_line_containing("$Elements")) FatalError("$Elements tag not found!");
  int n_entities = (int)f.next_int();
  f.skip_line();
  // file position -> c2v slot (reference src/mesh_reader.cpp:560-640)
  static const int id3[3] = {0, 1, 2}, id6[6] = {0, 1, 2, 3, 4, 5};
  static const int q4[4] = {0, 1, 3, 2}, q8[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  static const int t4[4] = {0, 1, 2, 3}, t10[10] = {0, 1, 2, 3, 4, 7, 5, 6, 8, 9};
  static const int p6[6] = {0, 1, 2, 3, 4, 5}, p15[15] = {0, 1, 2, 3, 4, 5, 6, 8, 9, 7, 10, 11, 12, 14, 13};
  static const int h8[8] = {0, 1, 3, 2, 4, 5, 7, 6};
  int icount = 0, i = 0;
  for (int k = 0; k < n_entities; k++)
  {
    (void)f.next_int();
    int elmtype = (int)f.next_int();
    int ntags = (int)f.next_int();
    int phys = (int)f.next_int();
    for (int t = 0; t < ntags - 1; t++) (void)f.next_int();
    if (phys == fluid_id)
    {
      if (icount >= kstart && i < in_num_cells)
      {
        m->ic2icg(i) = icount;
        const int *map = nullptr;
        int nv = 0;
        switch (elmtype)
        {
        case 2: m->ctype(i) = TRI; nv = 3; map = id3; break;
        case 9: m->ctype(i) = TRI; nv = 6; map = id6; break;
        case 3: m->ctype(i) = QUAD; nv = 4; map = q4; break;
        case 16: m->ctype(i) = QUAD; nv = 8; map = q8; break;
        case 4: m->ctype(i) = TET; nv = 4; map = t4; break;
        case 11: m->ctype(i) = TET; nv = 10; map = t10; break;
        case 6: m->ctype(i) = PRISM; nv = 6; map = p6; break;
        case 18: m->ctype(i) = PRISM; nv = 15; map = p15; break;
        case 5: m->ctype(i) = HEX; nv = 8; map = h8; break;
        default: FatalError("element type not recognized");
        }
        m->c2n_v(i) = nv;
        for (int q = 0; q < nv; q++) m->c2v(i, map[q]) = (int)f.next_int() - 1;
        i++;
      }
      icount++;
    }
    f.skip_line();
  }
}

void mesh_reader::read_vertices_gmsh()
{
  text_file f(fname);
  if (!f.skip_to_line_containing("$Nodes")) FatalError("$Nodes tag not found!");
  mesh *m = mesh_ptr;
  (void)f.next_int();
  m->xv.setup(m->num_verts, m->n_dims);
  const int *b = m->iv2ivg.get_ptr_cpu();
  for (int i = 0; i < m->num_verts_global; i++)
  {
    int id = (int)f.next_int();
    const int *q = lower_bound(b, b + m->num_verts, id - 1);
    if (q != b + m->num_verts && *q == id - 1)
    {
      int index = (int)(q - b);
      for (int d = 0; d < m->n_dims; d++) m->xv(index, d) = f.next_double();
    }
    f.skip_line();
  }
}

void mesh_reader::read_boundary_gmsh()
{
  // boundary entities are lower-dimensional elements tagged with a physical group; a boundary face belongs to
  // the cell that contains all its vertices (reference src/mesh_reader.cpp:700-890)
  mesh *m = mesh_ptr;
  m->bc_id.setup(m->num_cells, MAX_F_PER_C);
  m->bc_id.initialize_to_value(-1);
  if ((int)run_input.bc_list.size() != m->n_bdy) run_input.bc_list.assign(m->n_bdy, bc());
  for (int i = 0; i < m->n_bdy; i++)
    if (run_input.bc_list[i].get_bc_name() != gmsh_bc_names[i]) { run_input.bc_list[i] = bc(); run_input.bc_list[i].setup(gmsh_bc_names[i]); }
  int fluid_id = gmsh_elements_block_start;
  text_file f(fname);
  if (!f.skip_to_line_containing("$Elements")) FatalError("$Elements tag not found!");
  int n_entities = (int)f.next_int();
  f.skip_line();
  const int *b = m->iv2ivg.get_ptr_cpu();
  int vl[4], fv[4];
  for (int k = 0; k < n_entities; k++)
  {
    (void)f.next_int();
    int elmtype = (int)f.next_int();
    int ntags = (int)f.next_int();
    int phys = (int)f.next_int();
    for (int t = 0; t < ntags - 1; t++) (void)f.next_int();
    if (phys != fluid_id)
    {
      auto it = gmsh_phys2bc.find(phys);
      if (it == gmsh_phys2bc.end()) FatalError("boundary entity with unknown physical group");
      int bcflag = it->second;
      int nv = 0;
      if (elmtype == 1 || elmtype == 8) nv = 2;      // line / 3-node line: corners first
      else if (elmtype == 2 || elmtype == 9) nv = 3;  // tri
      else if (elmtype == 3 || elmtype == 16) nv = 4; // quad
      else FatalError("Boundary elmtype not recognized");
      bool local = true;
      for (int q = 0; q < nv; q++)
      {
        int g = (int)f.next_int() - 1;
        const int *p = lower_bound(b, b + m->num_verts, g);
        if (p == b + m->num_verts || *p != g) { local = false; break; }
        vl[q] = (int)(p - b);
      }
      if (local)
      {
        // cells containing every vertex of the boundary entity
        vector<int> inter = m->v2c[vl[0]], tmp;
        for (int q = 1; q < nv; q++)
        {
          tmp.clear();
          set_intersection(m->v2c[vl[q]].begin(), m->v2c[vl[q]].end(), inter.begin(), inter.end(), back_inserter(tmp));
          inter.swap(tmp);
        }
        for (size_t c = 0; c < inter.size(); c++)
        {
          int ic = inter[c];
          for (int lf = 0; lf < k_num_f_per_c[m->ctype(ic)]; lf++)
          {
            int nfv = m->get_corner_vlist_face(ic, lf, fv);
            if (nfv != nv) continue;
            int cnt = 0;
            for (int a = 0; a < nv; a++)
              for (int q = 0; q < nv; q++)
                if (vl[a] == fv[q]) { cnt++; break; }
            if (cnt == nv) m->bc_id(ic, lf) = bcflag;
          }
        }
      }
    }
    f.skip_line();
  }
}
