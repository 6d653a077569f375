"""Multi-GPU scheme of the path: reads sharded, index replicated, one final sum-reduce of the accumulators.

This is the reference's own multi-process ("MPI small-mem") scheme: every rank skips to its share of the reads
in 2048-read slices (reference inc/SeqManager.h:329-345, READS_PER_PROC inc/const_include.h:64), holds the whole
genome, and after its last batch the per-position accumulators are summed across ranks (`Allreduce(SUM, float)`
of amount_genome, reference src/Driver.cpp:1660-1672, and `Reduce(SUM)` of the five base planes, :1719-1767).
Here one process drives one GPU and the collective is NCCL over NVLink (`torch.distributed`, backend "nccl");
the same code runs on the "gloo" backend for the CPU tests.  Nothing else on the path communicates.
"""
from __future__ import annotations

import numpy as np

from . import _abi

SLICE_READS = 2048


def shard_slices(n_reads: int, rank: int, world: int, slice_reads: int = SLICE_READS):
    """[(lo, hi)] of the read slices dealt round-robin to `rank`."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    out = []
    for k, lo in enumerate(range(0, n_reads, slice_reads)):
        if k % world == rank:
            out.append((lo, min(lo + slice_reads, n_reads)))
    return out


def shard_indices(n_reads: int, rank: int, world: int, slice_reads: int = SLICE_READS) -> np.ndarray:
    sl = shard_slices(n_reads, rank, world, slice_reads)
    return np.concatenate([np.arange(lo, hi, dtype=np.int64) for lo, hi in sl]) if sl else np.zeros(0, np.int64)


def shard_batch(batch: _abi.ReadBatch, rank: int, world: int, slice_reads: int = SLICE_READS) -> _abi.ReadBatch:
    """The rank's reads as one host batch (slices concatenated in order)."""
    parts = [batch.slice(lo, hi) for lo, hi in shard_slices(batch.n_reads, rank, world, slice_reads)]
    out = _abi.ReadBatch.__new__(_abi.ReadBatch)
    if not parts:
        out.offsets = np.zeros(1, np.int64); out.seq = np.zeros(0, np.uint8)
        out.qual = None if batch.qual is None else np.zeros(0, np.uint8)
        out.pwm = None if batch.pwm is None else np.zeros((0, 4), np.float32)
        out._mk()
        return out
    lens = np.concatenate([np.diff(p.offsets) for p in parts])
    out.offsets = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=out.offsets[1:])
    out.seq = np.concatenate([p.seq for p in parts])
    out.qual = None if batch.qual is None else np.concatenate([p.qual for p in parts])
    out.pwm = None if batch.pwm is None else np.concatenate([p.pwm for p in parts])
    out._mk()
    return out


class _CudaArray:
    """__cuda_array_interface__ view of a device pointer owned by the gmx context."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 3, "strides": None}


def device_accumulators(mapper, device):
    """The context's accumulators as torch tensors over the SAME device memory (no copy): [amount] or
    [amount, planes(5 * l_pac)] -- K3 scatters into them and the collective runs on them in place."""
    import torch
    amount_ptr, n_amount, plane_ptrs, n_plane = mapper.accumulators_device()
    out = [torch.as_tensor(_CudaArray(amount_ptr, n_amount), device=device)]
    if n_plane:
        out.append(torch.as_tensor(_CudaArray(plane_ptrs[0], 5 * n_plane), device=device))
    return out


def all_reduce_accumulators(tensors, group=None):
    """Sum the accumulators over all ranks, in place (ncclAllReduce on GPUs, gloo on the CPU)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def reduce_accumulators(tensors, dst: int = 0, group=None):
    """reference src/Driver.cpp:1719-1767: only the root needs the planes for PrintFinal."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for t in tensors:
        dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM, group=group)
