"""Multi-GPU plumbing: one process per GPU, particles partitioned, full met
replica per GPU, and ONE collective -- the sum-reduce of the accumulation grids
to rank 0 at each output interval (the mpif_tm_reduce_grid slot,
src/mpi_mod.f90:2395-2579, src/timemanager_mpi.f90:468-485).  torch.distributed
(NCCL on GPUs, gloo in the CPU tests) is only the transport."""
import numpy as np


def partition_rows(n, rank, world):
    """Rows of a release owned by `rank`: round-robin in release order, the
    reference's numrel/np + addone rule (src/releaseparticles_mpi.f90:141-152)."""
    return np.arange(rank, n, world)


def take_partition(parts, rank, world, cls):
    """Compact copy of the rows of `parts` owned by `rank`."""
    idx = partition_rows(parts.numpart, rank, world)
    q = cls(max(len(idx), 1), parts.nspec)
    for name in ("xtra1", "ytra1", "cbt") + parts.F32 + parts.I32:
        getattr(q, name)[:len(idx)] = getattr(parts, name)[idx]
    q.xmass1[:len(idx)] = parts.xmass1[idx]
    q.xscav_frac1[:len(idx)] = parts.xscav_frac1[idx]
    q.numpart = len(idx)
    return q


class DeviceGridView:
    """Zero-copy torch view of an engine grid (device memory owned by libfpb)."""

    def __init__(self, ptr, nfloats):
        self.__cuda_array_interface__ = {"shape": (nfloats,), "typestr": "<f4", "data": (ptr, False),
                                         "version": 2}


def device_grid_tensor(engine, which, device):
    import torch
    ptr, n = engine.grid_device_ptr(which)
    if not ptr or n == 0:
        return None
    return torch.as_tensor(DeviceGridView(ptr, n), device=device)


def reduce_grids_to_root(tensors, root=0):
    """Sum every tensor onto `root` (in place there), like MPI_Reduce(MPI_SUM)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for t in tensors:
        if t is not None:
            dist.reduce(t, dst=root, op=dist.ReduceOp.SUM)
