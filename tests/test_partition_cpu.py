"""CPU, multi-process (gloo): the host-side partition logic that feeds the NCCL halo exchange.  Each rank builds its
part of a periodic hex mesh with the host mirror (host_only: no device), then the ranks exchange, over gloo, what they
would send through NCCL and check the pairing rules the reference establishes with MPI (src/geometry.cpp:424-551,
1132-1340; src/mpi_inters.cpp:154-215):
  * rank A lists as many faces for neighbour B as B lists for A, in the same order;
  * flux point j of A's k-th face towards B coincides (up to the periodic shift) with flux point lut_A(j) of B's k-th
    face towards A, where lut_A is inters::get_lut of A's rotation tag -- i.e. the receive-side index map is right;
  * every element appears on exactly one rank and the local order is ascending global id (src/mesh.cpp:188-311)."""
import os
import sys

import numpy as np
import pytest

import util


def quad_lut(rot, n):
    """inters::get_lut for quadrilateral faces (reference src/inters.cpp:232-256)"""
    nf = n * n
    lut = np.zeros(nf, dtype=int)
    for i in range(n):
        for j in range(n):
            lut[i * n + j] = {0: (n - 1 - j) + n * i, 1: nf - (n - 1 - j) - n * i - 1, 2: n * j + i, 3: nf - n * j - i - 1}[rot]
    return lut


def worker(rank, world, port, workdir, n, order, part_kind, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        sys.path.insert(0, os.path.join(util.ROOT, "tests"))
        import conftest
        hb = conftest.load_package()
        import importlib
        mg = importlib.import_module("hifiles_solver_b200.meshgen")
        dist.init_process_group("gloo", rank=rank, world_size=world)
        if part_kind == "bricks":
            part = mg.block_partition(n, mg.blocks_for(world))
        elif part_kind == "metis":
            part = None  # the host mirror partitions the dual graph itself (METIS), identically on every rank
        else:
            part = np.random.default_rng(11).integers(0, world, n ** 3).astype(np.int32)
        inp = os.path.join(workdir, "input")
        run = hb.Run(inp, rank=rank, nproc=world, part=part, host_only=True)
        nn = order + 1
        nfi = nn * nn
        gid = run.host_array("hex.ele2global_ele")
        pos = run.host_array("hex.pos_fpts")  # (fpt, ele, dim)
        n_mpi = run.n_inters("mpi", 2)
        info = dict(rank=rank, gid=gid, n_mpi=n_mpi)
        if n_mpi:
            ele = run.host_array("mpi_quad.ele_l"); loc = run.host_array("mpi_quad.local_inter_l"); rot = run.host_array("mpi_quad.rot_tag")
            info["nb_rank"] = run.host_array("mpi_quad.neighbour_rank")
            info["nb_count"] = run.host_array("mpi_quad.neighbour_count")
            info["pos"] = np.stack([pos[loc[i] * nfi:(loc[i] + 1) * nfi, ele[i], :] for i in range(n_mpi)])  # (inter, fpt, dim)
            info["rot"] = rot[:n_mpi]
        run.close()
        out = [None] * world
        dist.all_gather_object(out, info)
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0:
            q.put(out)
    except Exception as e:  # pragma: no cover
        import traceback
        q.put("rank %d: %s\n%s" % (rank, e, traceback.format_exc()))


@pytest.mark.parametrize("world,part_kind", [(2, "bricks"), (2, "random"), (4, "bricks"), (2, "metis"), (4, "metis")])
def test_halo_pairing_rules(tmp_path, hb, meshgen, world, part_kind):
    import torch.multiprocessing as mp
    n, order = 4, 2
    meshgen.hex_box(str(tmp_path / "m.neu"), n)
    meshgen.write_input(str(tmp_path / "input"), "m.neu", order=order)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 7 + {"bricks": 0, "random": 3, "metis": 5}[part_kind]
    procs = [ctx.Process(target=worker, args=(r, world, port, str(tmp_path), n, order, part_kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert not isinstance(res, str), res
    # every cell on exactly one rank, ascending global ids locally
    allg = np.concatenate([r["gid"] for r in res])
    assert sorted(allg.tolist()) == list(range(n ** 3))
    for r in res:
        assert np.all(np.diff(r["gid"]) > 0)
    if part_kind == "metis":
        # k-way partition with the reference's 5 % imbalance tolerance (src/mesh.cpp:152-155); METIS may exceed it slightly on tiny graphs
        sizes = [len(r["gid"]) for r in res]
        assert max(sizes) <= 1.15 * n ** 3 / world and min(sizes) > 0, sizes
    L = 2 * np.pi
    nn = order + 1
    total = 0
    for a in res:
        if a["n_mpi"] == 0:
            continue
        start_a = np.concatenate([[0], np.cumsum(a["nb_count"])])
        for ia, b_rank in enumerate(a["nb_rank"]):
            b = res[int(b_rank)]
            ib = list(b["nb_rank"]).index(a["rank"])
            assert a["nb_count"][ia] == b["nb_count"][ib], "ranks %d and %d disagree on the number of shared faces" % (a["rank"], b_rank)
            start_b = np.concatenate([[0], np.cumsum(b["nb_count"])])
            for k in range(int(a["nb_count"][ia])):
                pa = a["pos"][start_a[ia] + k]
                pb = b["pos"][start_b[ib] + k]
                lut = quad_lut(int(a["rot"][start_a[ia] + k]), nn)
                d = np.abs(pa - pb[lut])
                d = np.minimum(d, np.abs(d - L))  # periodic images
                assert d.max() < 1e-9, "flux points of a shared face do not coincide (ranks %d/%d, face %d)" % (a["rank"], b_rank, k)
                total += 1
    assert total > 0


@pytest.mark.parametrize("kind,n,world", [("pritet", (2, 4, 2), 2), ("hexpri", (2, 2, 4), 3), ("mixed2d", 6, 4)])
def test_metis_partition_of_mixed_meshes(tmp_path, hb, meshgen, kind, n, world):
    """The k-way partitioner on meshes of several element types (corner vertices of triangles, quadrilaterals, tetrahedra,
    prisms, hexahedra in the dual graph): every rank computes the same vector, every cell lands on exactly one rank, the
    parts are balanced.  Ranks are built one after the other in this process (host only)."""
    mesh = str(tmp_path / "m.neu")
    extra = {}
    if kind == "mixed2d":
        info = meshgen.mixed_box_2d(mesh, n, kind="mixed", lengths=(6.2831853071795862,) * 2, origin=(0., 0.))
        extra = dict(dz_cyclic=None)
    else:
        info = meshgen.mixed_box_3d(mesh, n, kind=kind)
    n_cells = len(info["centroids"])
    inp = meshgen.write_input(str(tmp_path / "input"), "m.neu", order=1, adv_type=2, dt=1e-5, riemann_solve_type=0, viscous=1, **extra)
    owned = []
    for rank in range(world):
        with hb.Run(inp, rank=rank, nproc=world, part=None, host_only=True) as run:
            mine = np.concatenate([run.host_array(t + ".ele2global_ele") for t in run.ele_types()])
            owned.append(mine)
    allc = np.concatenate(owned)
    assert sorted(allc.tolist()) == list(range(n_cells))
    sizes = [len(o) for o in owned]
    assert min(sizes) > 0 and max(sizes) <= 1.25 * n_cells / world, sizes
