// Interface objects.  The reference wires every flux point of every interface to element storage through
// tables of double* (72 pointers per flux-point pair, reference src/inters.cpp:107-122, src/int_inters.cpp:54-63).
// Here an interface is five integers per side; the device layer expands them with the same getter arithmetic
// (reference src/eles.cpp:4638-4871) and the same right-side permutation lut (reference src/inters.cpp:153-262).
#include "hifiles.h"

using namespace std;

inters::inters()
{
  ctx = nullptr;
  inters_type = order = viscous = n_inters = n_fpts_per_inter = n_fields = n_dims = 0;
}

void inters::setup_inters(int in_n_inters, int in_inters_type)
{
  n_inters = in_n_inters;
  inters_type = in_inters_type;
  order = run_input.order;
  viscous = run_input.viscous;
  if (inters_type == 0) { n_dims = 2; n_fpts_per_inter = order + 1; }
  else if (inters_type == 1) { n_dims = 3; n_fpts_per_inter = (order + 2) * (order + 1) / 2; }
  else if (inters_type == 2) { n_dims = 3; n_fpts_per_inter = (order + 1) * (order + 1); }
  else FatalError("ERROR: Invalid interface type ... ");
  if (run_input.equation == 0) n_fields = n_dims + 2;
  else if (run_input.equation == 1) n_fields = 1;
  else FatalError("Equation not supported");
  lut.setup(n_fpts_per_inter);
  ele_type_l.setup(max(n_inters, 1));
  ele_l.setup(max(n_inters, 1));
  local_inter_l.setup(max(n_inters, 1));
}

void inters::get_lut(int in_rot_tag)
{
  int n = order + 1;
  if (inters_type == 0)
  {
    for (int i = 0; i < n_fpts_per_inter; i++) lut(i) = n_fpts_per_inter - i - 1;
  }
  else if (inters_type == 1)
  {
    for (int j = 0; j < n; j++)
      for (int i = 0; i < n - j; i++)
      {
        int index0 = j * n - (j - 1) * j / 2 + i;
        int index1;
        if (in_rot_tag == 0) index1 = i * n - (i - 1) * i / 2 + j;
        else if (in_rot_tag == 1) index1 = n * (order + 2) / 2 - 1 - (i + j) * (i + j + 1) / 2 - j;
        else if (in_rot_tag == 2) index1 = j * n - (j - 1) * j / 2 + (order - j - i);
        else { FatalError("ERROR: Unknown rotation of triangular face..."); index1 = 0; }
        lut(index0) = index1;
      }
  }
  else if (inters_type == 2)
  {
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++)
      {
        int v;
        if (in_rot_tag == 0) v = (n - 1 - j) + n * i;
        else if (in_rot_tag == 1) v = n_fpts_per_inter - (n - 1 - j) - n * i - 1;
        else if (in_rot_tag == 2) v = n * j + i;
        else if (in_rot_tag == 3) v = n_fpts_per_inter - n * j - i - 1;
        else { FatalError("ERROR: Unknown rotation tag ... "); v = 0; }
        lut(i * n + j) = v;
      }
  }
  else
    FatalError("ERROR: Invalid interface type ... ");
}

// ---- interior ----------------------------------------------------------------------------------------------------
void int_inters::setup(int in_n_inters, int in_inter_type)
{
  setup_inters(in_n_inters, in_inter_type);
  int n = max(n_inters, 1);
  ele_type_r.setup(n);
  ele_r.setup(n);
  local_inter_r.setup(n);
  rot_tags.setup(n);
}

void int_inters::set_interior(int in_inter, int in_ele_type_l, int in_ele_type_r, int in_ele_l, int in_ele_r, int in_local_inter_l, int in_local_inter_r, int rot_tag, struct solution *)
{
  ele_type_l(in_inter) = in_ele_type_l;
  ele_type_r(in_inter) = in_ele_type_r;
  ele_l(in_inter) = in_ele_l;
  ele_r(in_inter) = in_ele_r;
  local_inter_l(in_inter) = in_local_inter_l;
  local_inter_r(in_inter) = in_local_inter_r;
  rot_tags(in_inter) = rot_tag;
}

void int_inters::mv_all_cpu_gpu()
{
  if (n_inters == 0) return;
  hf_int_inters_desc d;
  d.inter_type = inters_type;
  d.n_inters = n_inters;
  d.n_fpts_per_inter = n_fpts_per_inter;
  d.ele_type_l = ele_type_l.get_ptr_cpu();
  d.ele_l = ele_l.get_ptr_cpu();
  d.local_inter_l = local_inter_l.get_ptr_cpu();
  d.ele_type_r = ele_type_r.get_ptr_cpu();
  d.ele_r = ele_r.get_ptr_cpu();
  d.local_inter_r = local_inter_r.get_ptr_cpu();
  d.rot_tag = rot_tags.get_ptr_cpu();
  hf_check(hf_dev_upload_int_inters(ctx, &d));
}

void int_inters::calculate_common_invFlux() { if (n_inters) hf_check(hf_dev_int_inters_op(ctx, inters_type, HF_COMMON_INVFLUX)); }
void int_inters::calculate_common_viscFlux() { if (n_inters) hf_check(hf_dev_int_inters_op(ctx, inters_type, HF_COMMON_VISCFLUX)); }

// ---- boundary ----------------------------------------------------------------------------------------------------
void bdy_inters::setup(int in_n_inters, int in_inter_type)
{
  setup_inters(in_n_inters, in_inter_type);
  boundary_id.setup(max(n_inters, 1));
  pos_fpts.setup(n_fpts_per_inter, max(n_inters, 1), n_dims);
}

void bdy_inters::set_boundary(int in_inter, int bc_id, int in_ele_type_l, int in_ele_l, int in_local_inter_l, struct solution *FlowSol)
{
  boundary_id(in_inter) = bc_id;
  ele_type_l(in_inter) = in_ele_type_l;
  ele_l(in_inter) = in_ele_l;
  local_inter_l(in_inter) = in_local_inter_l;
  eles *e = FlowSol->mesh_eles(in_ele_type_l);
  for (int j = 0; j < n_fpts_per_inter; j++)
  {
    int fpt = e->get_fpt_index(j, in_local_inter_l);
    for (int k = 0; k < n_dims; k++) pos_fpts(j, in_inter, k) = e->pos_fpts(fpt, in_ele_l, k);
  }
  // wall model: the solution point of the element farthest from this face feeds it (reference src/bdy_inters.cpp:149-162,
  // eles::calc_wm_upts_dist src/eles.cpp:4873-4903)
  if (wm_upt.get_dim(0) != max(n_inters, 1)) { wm_upt.setup(max(n_inters, 1)); wm_upt.initialize_to_value(-1); wm_dist.setup(max(n_inters, 1)); }
  if (run_input.bc_list[bc_id].use_wm)
  {
    double dist_min = 0., dist_max = 0.;
    int best = 0;
    for (int i = 0; i < e->n_upts_per_ele; i++)
    {
      for (int j = 0; j < n_fpts_per_inter; j++)
      {
        const int fpt = e->get_fpt_index(j, in_local_inter_l);
        double d = 0.;
        for (int k = 0; k < n_dims; k++) d += (e->pos_fpts(fpt, in_ele_l, k) - e->pos_upts(i, in_ele_l, k)) * e->norm_fpts(fpt, in_ele_l, k);
        if (j == 0 || d < dist_min) dist_min = d;
      }
      if (dist_min > dist_max) { dist_max = dist_min; best = i; }
    }
    wm_upt(in_inter) = best + e->n_upts_per_ele * in_ele_l;
    wm_dist(in_inter) = dist_max;
    any_wm = true;
  }
}

void bdy_inters::mv_all_cpu_gpu()
{
  if (n_inters == 0) return;
  hf_bdy_inters_desc d;
  d.inter_type = inters_type;
  d.n_inters = n_inters;
  d.n_fpts_per_inter = n_fpts_per_inter;
  d.ele_type_l = ele_type_l.get_ptr_cpu();
  d.ele_l = ele_l.get_ptr_cpu();
  d.local_inter_l = local_inter_l.get_ptr_cpu();
  d.wm_upt = any_wm ? wm_upt.get_ptr_cpu() : nullptr;
  d.wm_dist = any_wm ? wm_dist.get_ptr_cpu() : nullptr;
  d.bc_id = boundary_id.get_ptr_cpu();
  d.pos_fpts = pos_fpts.get_ptr_cpu();
  hf_check(hf_dev_upload_bdy_inters(ctx, &d));
}

void bdy_inters::evaluate_boundaryConditions_invFlux(struct solution *, double time_bound)
{
  if (n_inters) hf_check(hf_dev_bdy_inters_op(ctx, inters_type, HF_COMMON_INVFLUX, time_bound));
}
void bdy_inters::evaluate_boundaryConditions_viscFlux(double time_bound)
{
  if (n_inters) hf_check(hf_dev_bdy_inters_op(ctx, inters_type, HF_COMMON_VISCFLUX, time_bound));
}

// ---- partition ("mpi") interfaces ----------------------------------------------------------------------------------
void mpi_inters::setup(int in_n_inters, int in_inter_type)
{
  setup_inters(in_n_inters, in_inter_type);
  rot_tags.setup(max(n_inters, 1));
  neighbour_rank.clear();
  neighbour_count.clear();
}

void mpi_inters::set_nout_proc(int in_nout, int in_p)
{
  neighbour_rank.push_back(in_p);
  neighbour_count.push_back(in_nout);
}

void mpi_inters::set_mpi(int in_inter, int in_ele_type_l, int in_ele_l, int in_local_inter_l, int rot_tag, struct solution *FlowSol)
{
  if ((int)ele_global_l.size() < n_inters) ele_global_l.assign(n_inters, -1);
  if (FlowSol) ele_global_l[in_inter] = FlowSol->mesh_eles(in_ele_type_l)->ele2global_ele(in_ele_l);
  ele_type_l(in_inter) = in_ele_type_l;
  ele_l(in_inter) = in_ele_l;
  local_inter_l(in_inter) = in_local_inter_l;
  rot_tags(in_inter) = rot_tag;
}

void mpi_inters::mv_all_cpu_gpu()
{
  if (n_inters == 0) return;
  hf_mpi_inters_desc d;
  d.inter_type = inters_type;
  d.n_inters = n_inters;
  d.n_fpts_per_inter = n_fpts_per_inter;
  d.ele_type_l = ele_type_l.get_ptr_cpu();
  d.ele_l = ele_l.get_ptr_cpu();
  d.local_inter_l = local_inter_l.get_ptr_cpu();
  d.rot_tag = rot_tags.get_ptr_cpu();
  d.n_neighbours = (int)neighbour_rank.size();
  d.neighbour_rank = neighbour_rank.data();
  d.neighbour_count = neighbour_count.data();
  d.ele_global_l = ((int)ele_global_l.size() == n_inters) ? ele_global_l.data() : nullptr;
  hf_check(hf_dev_upload_mpi_inters(ctx, &d));
}

// split-phase halo exchange: op codes 2..5 of hf_dev_mpi_inters_op
void mpi_inters::send_solution() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 2)); }
void mpi_inters::receive_solution() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 3)); }
void mpi_inters::send_corrected_gradient() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 4)); }
void mpi_inters::receive_corrected_gradient() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 5)); }
void mpi_inters::send_sgsf_fpts() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 6)); }
void mpi_inters::receive_sgsf_fpts() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, 7)); }
void mpi_inters::calculate_common_invFlux() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, HF_COMMON_INVFLUX)); }
void mpi_inters::calculate_common_viscFlux() { if (n_inters) hf_check(hf_dev_mpi_inters_op(ctx, inters_type, HF_COMMON_VISCFLUX)); }
