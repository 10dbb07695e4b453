"""Run under torchrun (one rank per GPU): the TGV case split into bricks over the ranks must reproduce the
single-domain result.  Rank 0 also runs the whole mesh on its own GPU (staged kernels = bit-exact reference order) and
compares element by element through ele2global_ele.  Usage:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multi_gpu_check.py [n] [order] [steps] [fused|staged] [hex|tet|pri|hexpri|pritet|tri|mixed]
For the simplex / mixed kinds the mesh is cut into slabs (meshgen.slab_partition) and every element type is compared."""
import os
import pathlib
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(pathlib.Path(__file__).parent))
import conftest  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    mode = sys.argv[4] if len(sys.argv) > 4 else "fused"
    kind = sys.argv[5] if len(sys.argv) > 5 else "hex"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hb = conftest.load_package()
    import importlib
    mg = importlib.import_module("hifiles_solver_b200.meshgen")
    obj = [tempfile.mkdtemp(prefix="hf_mgpu_") if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    work = obj[0]
    mesh = os.path.join(work, "tgv.neu")
    inp = os.path.join(work, "input")
    if kind != "hex":
        return general(hb, mg, dist, rank, world, work, kind, n, order, steps, mode)
    if rank == 0:
        if os.environ.get("HF_CHECK_WALLS"):
            # channel-like mesh: periodic in x, y; isothermal wall and characteristic far field in z (boundary faces inside the fused kernels)
            mg.hex_box(mesh, n, lengths=(1., 1., 2.), bcs={"x-": "Cyclic", "x+": "Cyclic", "y-": "Cyclic", "y+": "Cyclic", "z-": "Wall", "z+": "Far"})
            mg.write_input(inp, "tgv.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1, ic_form=1, Mach_c_ic=0.2, nx_c_ic=1., ny_c_ic=0.,
                           nz_c_ic=0.05, T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.2, rho_free_stream=1.17, T_free_stream=300., L_free_stream=1.,
                           dx_cyclic=1., dy_cyclic=1., dz_cyclic=None, bc_Wall_type="isotherm_wall", bc_Wall_T_static=310., bc_Far_type="sub_out_char",
                           bc_Far_p_static=100500.)
        else:
            mg.hex_box(mesh, n)
            mg.write_input(inp, "tgv.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=3, viscous=1)
    dist.barrier()
    part = None if os.environ.get("HF_CHECK_PART") == "metis" else mg.block_partition(n, mg.blocks_for(world))  # None: METIS k-way in the host mirror
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(hb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, src=0)
    run = hb.Run(inp, rank=rank, nproc=world, part=part, nccl_id=bytes(idt.cpu().tolist()))
    if mode == "staged":
        run.set_mode(False)
    else:
        assert run.fused_status() == "available", run.fused_status()
    variant = run.fused_variant() if mode != "staged" else "staged"
    run.run(steps, fused=True)
    u = run.download("hex", "disu_upts")
    gid = run.host_array("hex.ele2global_ele")
    n_mpi = run.n_inters("mpi", 2)
    res_part = run.norm_residual()  # reduced over the ranks (reference src/output.cpp:2216-2231): every rank holds the global value
    run.close()
    # gather on rank 0
    nu = u.shape[0]
    full = torch.zeros((n ** 3, nu, 5), dtype=torch.float64, device="cuda")
    full[torch.from_numpy(gid.astype(np.int64)).cuda()] = torch.from_numpy(np.ascontiguousarray(u.transpose(1, 0, 2))).cuda()
    dist.all_reduce(full)
    ok = True
    if rank == 0:
        with hb.Run(inp) as single:
            single.set_mode(False)
            single.run(steps, fused=True)
            us = single.download("hex", "disu_upts")
            gs = single.host_array("hex.ele2global_ele")
            res_single = single.norm_residual()
        ref = np.zeros((n ** 3, nu, 5))
        ref[gs] = us.transpose(1, 0, 2)
        got = full.cpu().numpy()

        def scaled(a, b):
            sc = np.abs(b).reshape(-1, 5).max(0)
            sc[1:4] = sc[1:4].max()
            return (np.abs(a - b).reshape(-1, 5).max(0) / sc).max()

        err = scaled(got, ref)
        err_res = np.abs(np.asarray(res_part) - np.asarray(res_single)).max() / np.abs(res_single).max()
        ok = bool(err < 1e-12 and err_res < 1e-12)
        msg = ""
        # and against the unmodified reference CPU solver (serial: it numbers its elements globally)
        sys.path.insert(0, str(pathlib.Path(__file__).parent))
        import util
        if util.have_reference():
            r = util.run_reference(inp, steps, stagewise=False)
            err_ref = scaled(got, np.ascontiguousarray(r["final.hex.disu_upts"].transpose(1, 0, 2)))
            err_ref_res = np.abs(np.asarray(res_part) - r["history.norm_residual"][:, -1]).max() / np.abs(r["history.norm_residual"][:, -1]).max()
            ok = ok and bool(err_ref < 1e-12 and err_ref_res < 1e-12)
            msg = "  vs reference CPU solver %.3e (residual norm %.3e)" % (err_ref, err_ref_res)
        print("multi_gpu_check: world=%d n=%d order=%d steps=%d mode=%s [%s] partition faces on rank 0: %d  max rel err vs single domain %.3e (residual norm %.3e)%s  %s"
              % (world, n, order, steps, mode, variant, n_mpi, err, err_res, msg, "OK" if ok else "FAIL"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


def general(hb, mg, dist, rank, world, work, kind, n, order, steps, mode="staged"):
    """staged kernels (mode "staged") or the blocked element kernels of the fast mode (mode "fused": hf_elem.cu around the halo
    exchange of the staged interface kernels) on a partitioned simplex / prism / mixed mesh (triangular, quadrilateral and segment
    partition faces); the yardstick is the single-domain run of the staged kernels"""
    fast = mode != "staged"
    mesh = os.path.join(work, "m.neu")
    inp = os.path.join(work, "input")
    two_d = kind in ("tri", "mixed", "quad")
    obj = [None]
    if rank == 0:
        if two_d and os.environ.get("HF_CHECK_WALLS"):
            # channel with ONE no-slip wall (below), cut into a lower and an upper slab: the upper rank has no wall point of its own, the
            # Smagorinsky damping there needs the lower rank's (FinishWallDistance, reference src/geometry.cpp:768-892)
            info = mg.mixed_box_2d(mesh, (n, n), kind=kind, lengths=(4., 2.), origin=(0., 0.), bcs={"x-": "In", "x+": "Out", "y-": "Wall", "y+": "Top"})
            mg.write_input(inp, "m.neu", order=order, adv_type=3, riemann_solve_type=0, viscous=1, ic_form=1, dt=5e-5, fix_vis=0, Mach_c_ic=0.3, nx_c_ic=1., ny_c_ic=0.,
                           nz_c_ic=0., T_c_ic=300., rho_c_ic=1.17, Mach_free_stream=0.3, rho_free_stream=1.17, T_free_stream=300., L_free_stream=1., dx_cyclic=None,
                           dy_cyclic=None, dz_cyclic=None, bc_Cyclic_type=None, bc_In_type="char", bc_In_p_static=100747., bc_In_mach=0.3, bc_In_T_static=300.,
                           bc_In_nx=1., bc_In_ny=0., bc_Out_type="sub_out_simp", bc_Out_p_static=100000., bc_Wall_type="isotherm_wall", bc_Wall_T_static=310.,
                           bc_Top_type="slip_wall", LES=1, SGS_model=0, C_s=0.1, filter_ratio=2.0, calc_force=1, monitor_cp_freq=100000, area_ref=1.0)
        elif two_d:
            info = mg.mixed_box_2d(mesh, n, kind=kind, lengths=(6.2831853071795862,) * 2, origin=(0., 0.))
            mg.write_input(inp, "m.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, dz_cyclic=None)
        else:
            info = mg.mixed_box_3d(mesh, n, kind=kind)
            les = dict(LES=1, SGS_model=int(os.environ["HF_CHECK_LES"]), C_s=0.3, filter_ratio=2.0, filter_type=2) if os.environ.get("HF_CHECK_LES") else {}
            mg.write_input(inp, "m.neu", order=order, adv_type=2, dt=1e-5, riemann_solve_type=2, viscous=1, **les)
        obj = [mg.slab_partition(info["centroids"], world, axis=1 if (kind == "pritet" or os.environ.get("HF_CHECK_WALLS")) else 0)]
    dist.broadcast_object_list(obj, src=0)
    part = obj[0]
    idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        idt = torch.tensor(list(hb.nccl_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(idt, src=0)
    run = hb.Run(inp, rank=rank, nproc=world, part=part, nccl_id=bytes(idt.cpu().tolist()))
    if fast:
        assert run.elem_status() == "available", run.elem_status()
        run.run(steps, fused=True)
    else:
        run.set_mode(False)
        run.run(steps, fused=False)
    mine = {t: (run.download(t, "disu_upts"), run.host_array(t + ".ele2global_ele")) for t in run.ele_types()}
    n_mpi = sum(run.n_inters("mpi", i) for i in range(3))
    run.close()
    ok, worst = True, 0.
    n_cells = len(part)
    for t in ("tri", "quad", "tet", "pri", "hex"):
        shape = torch.zeros(2, dtype=torch.int64, device="cuda")
        if t in mine:
            shape[0], shape[1] = mine[t][0].shape[0], mine[t][0].shape[2]
        dist.all_reduce(shape, op=dist.ReduceOp.MAX)
        nu, nf = int(shape[0]), int(shape[1])
        if nu == 0:
            continue
        full = torch.zeros((n_cells, nu, nf), dtype=torch.float64, device="cuda")
        if t in mine:
            u, gid = mine[t]
            full[torch.from_numpy(gid.astype(np.int64)).cuda()] = torch.from_numpy(np.ascontiguousarray(u.transpose(1, 0, 2))).cuda()
        dist.all_reduce(full)
        if rank == 0:
            with hb.Run(inp) as single:
                single.set_mode(False)
                single.run(steps, fused=False)
                us, gs = single.download(t, "disu_upts"), single.host_array(t + ".ele2global_ele")
            ref = np.zeros((n_cells, nu, nf))
            ref[gs] = us.transpose(1, 0, 2)
            got = full.cpu().numpy()
            sc = np.abs(ref).reshape(-1, nf).max(0)
            sc[1:nf - 1] = sc[1:nf - 1].max()
            err = (np.abs(got - ref).reshape(-1, nf).max(0) / sc).max()
            worst = max(worst, err)
            ok = ok and bool(err < 1e-12)
    if rank == 0:
        print("multi_gpu_check: world=%d kind=%s n=%s order=%d steps=%d mode=%s partition faces on rank 0: %d  max rel err vs single domain %.3e  %s"
              % (world, kind, n, order, steps, "blocked element kernels" if fast else "staged", n_mpi, worst, "OK" if ok else "FAIL"))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
