"""Multi-GPU parity (runs only where >= 2 CUDA devices are visible): the partitioned run over NCCL must reproduce the
single-domain run.  The host-side partition logic is covered on CPU by test_partition_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["staged", "fused"])
def test_two_ranks_match_single_domain(mode):
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "6", "2", "2", mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("part", ["bricks", "metis"])
def test_two_ranks_generation9_match_single_domain_and_reference(part):
    """P = 4: the generation-9 kernels (k_face9 + k_resid9) with the halo exchange of their padded face blocks, brick and METIS partitions"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29539", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "4", "4", "2", "fused"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_CHECK_PART=part))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout and "generation 9" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("part", ["bricks", "metis"])
def test_two_ranks_fused_kernels_with_boundary_faces(part):
    """partition faces AND boundary faces (isothermal wall, characteristic far field) in the generation-9 kernels: the boundary faces' virtual
    neighbour blocks sit behind the halo receive blocks"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29543", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "4", "3", "2", "fused"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_CHECK_PART=part, HF_CHECK_WALLS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout and "generation 9" in r.stdout


@pytest.mark.gpu
def test_two_ranks_single_launch_stage_with_in_kernel_wait():
    """enough interior elements (>= 4096 per rank) for the one-launch-per-kernel stage: partition-adjacent elements wait inside the
    kernel for the exchange counter (hf_fused.cu, hf_fused9.cuh wait_exchange9)"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "22", "4", "1", "fused"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout and "generation 9" in r.stdout


@pytest.mark.gpu
def test_two_ranks_match_single_domain_metis_partition():
    """fused kernels on a METIS k-way partition (ragged partition boundary instead of a plane)"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "6", "2", "2", "fused"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_CHECK_PART="metis"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["staged", "fused"])
def test_two_ranks_match_single_domain_les(mode):
    """LES (WALE-similarity: eddy viscosity + Leonard tensors) on a partitioned tetrahedral mesh: the SGS-flux halo exchange; mode "fused" =
    the blocked element kernels around the staged sub-grid-scale point fluxes"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29535", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "3", "2", "2", mode, "tet"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_CHECK_LES="2"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["staged", "fused"])
@pytest.mark.parametrize("kind,n", [("pritet", "4"), ("hexpri", "4"), ("mixed", "6")])
def test_two_ranks_match_single_domain_simplex_and_mixed(kind, n, mode):
    """partition faces of every face type (segments, triangles, quadrilaterals) between element types of every kind; mode "fused" = the
    blocked element kernels of the fast mode (hf_elem.cu) around the same halo exchange"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "tests", "multi_gpu_check.py"), n, "2", "2", mode, kind]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["staged", "fused"])
def test_two_ranks_smagorinsky_wall_distance_across_ranks(mode):
    """Smagorinsky near-wall damping on a partitioned mesh whose only no-slip wall lies on the other rank: the wall points of every rank
    travel over the communicator once it exists (reference src/geometry.cpp:768-892 gathers them with MPI during the geometry setup)"""
    if n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29545", os.path.join(ROOT, "tests", "multi_gpu_check.py"), "6", "2", "2", mode, "mixed"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, HF_CHECK_WALLS="1"))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout
