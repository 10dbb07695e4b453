// Mesh partitioning: k-way partition of the dual graph of the mesh, the serial counterpart of the reference's
// ParMETIS_V3_PartMeshKway call (reference src/mesh.cpp:72-183): corner vertices of every cell as the mesh description,
// two cells are neighbours when they share ncommonnodes vertices (2 in 2-D, 3 in 3-D), unit cell weights, equal target
// weights, 5 % imbalance tolerance, seed 0.  ParMETIS itself is not in this image; METIS (the serial library ParMETIS is
// built on) ships with the CUDA toolkit as libmetis_static.a (64-bit idx_t), bound here with hand-written prototypes.
// Every rank reads the whole mesh file and METIS is deterministic for a given seed, so all ranks compute the same
// vector without communication.  No reference test pins a partition: results do not depend on it
// beyond rounding (tests/multi_gpu_check.py).
#include "hifiles.h"
#include <cstdint>
#include <cmath>

extern "C"
{
typedef int64_t metis_idx_t; // libmetis_static.a of CUDA 12.9 is built with IDXTYPEWIDTH 64
int METIS_SetDefaultOptions(metis_idx_t *options);
int METIS_PartMeshDual(metis_idx_t *ne, metis_idx_t *nn, metis_idx_t *eptr, metis_idx_t *eind, metis_idx_t *vwgt, metis_idx_t *vsize,
                       metis_idx_t *ncommon, metis_idx_t *nparts, void *tpwgts, metis_idx_t *options, metis_idx_t *objval,
                       metis_idx_t *epart, metis_idx_t *npart);
}

// corner vertices of a cell (reference mesh::get_corner_vert_in_order, src/mesh.cpp:487-580), as entries of c2v
int mesh::get_corner_vlist(int in_ic, int *v) const
{
  const int ns = c2n_v(in_ic), ct = ctype(in_ic);
  int nv = 0;
  if (ct == TRI)
  {
    if (ns != 3 && ns != 6) FatalError("in_nspt not implemented");
    nv = 3;
    for (int i = 0; i < 3; i++) v[i] = i;
  }
  else if (ct == QUAD)
  {
    nv = 4;
    if (is_perfect_square(ns))
    {
      int n1 = (int)lround(sqrt((double)ns));
      v[0] = 0; v[1] = n1 - 1; v[2] = ns - 1; v[3] = ns - n1;
    }
    else if (ns == 8) { for (int i = 0; i < 4; i++) v[i] = i; }
    else FatalError("in_nspt not implemented");
  }
  else if (ct == TET)
  {
    if (ns != 4 && ns != 10) FatalError("in_nspt not implemented");
    nv = 4;
    for (int i = 0; i < 4; i++) v[i] = i;
  }
  else if (ct == PRISM)
  {
    if (ns != 6 && ns != 15) FatalError("in_nspt not implemented");
    nv = 6;
    for (int i = 0; i < 6; i++) v[i] = i;
  }
  else if (ct == HEX)
  {
    nv = 8;
    if (is_perfect_cube(ns))
    {
      int n1 = (int)lround(pow((double)ns, 1. / 3.));
      int shift = n1 * n1 * (n1 - 1);
      const int c[8] = {0, n1 - 1, n1 * n1 - 1, n1 * (n1 - 1), shift, n1 - 1 + shift, ns - 1, ns - n1};
      for (int i = 0; i < 8; i++) v[i] = c[i];
    }
    else if (ns == 20) { for (int i = 0; i < 8; i++) v[i] = i; }
    else FatalError("n_spts not implemented");
  }
  else
    FatalError("unknown element type, in repartitioning");
  for (int i = 0; i < nv; i++) v[i] = c2v(in_ic, v[i]);
  return nv;
}

// part[global cell] in [0, nproc) for the whole mesh held by m (before apply_partition)
void partition_mesh_kway(const mesh &m, int n_dims, int nproc, std::vector<int> &part)
{
  const int nc = m.num_cells;
  part.assign(nc, 0);
  if (nproc <= 1) return;
  if (nc < nproc) FatalError("fewer cells than ranks");
  std::vector<metis_idx_t> eptr(nc + 1, 0), eind;
  eind.reserve((size_t)nc * 8);
  metis_idx_t max_v = -1;
  int v[8];
  for (int i = 0; i < nc; i++)
  {
    int nv = m.get_corner_vlist(i, v);
    for (int j = 0; j < nv; j++)
    {
      eind.push_back(v[j]);
      if (v[j] > max_v) max_v = v[j];
    }
    eptr[i + 1] = (metis_idx_t)eind.size();
  }
  // METIS wants the nodes numbered 0 .. nn-1 without gaps being required only for sizing: nn = highest id + 1
  metis_idx_t ne = nc, nn = max_v + 1, ncommon = (n_dims == 2) ? 2 : 3, nparts = nproc, objval = 0;
  metis_idx_t options[40];
  METIS_SetDefaultOptions(options);
  options[8] = 0;   // METIS_OPTION_SEED      (ParMETIS options[2] = 0 in the reference)
  options[16] = 50; // METIS_OPTION_UFACTOR   (load imbalance 1.05 = ubvec of the reference)
  options[17] = 0;  // METIS_OPTION_NUMBERING (C-style, numflag 0)
  std::vector<metis_idx_t> epart(nc, 0), npart((size_t)nn, 0);
  int status = METIS_PartMeshDual(&ne, &nn, eptr.data(), eind.data(), nullptr, nullptr, &ncommon, &nparts, nullptr, options, &objval,
                                  epart.data(), npart.data());
  if (status != 1) FatalError("METIS_PartMeshDual failed");
  for (int i = 0; i < nc; i++) part[i] = (int)epart[i];
}
