"""Test-side helpers for the multi-rank host logic (one process per GPU): everything here runs on any torch.distributed
backend (gloo in the CPU tests).  Not part of the product package."""
from __future__ import annotations

from typing import List, Tuple


def subfile_range(numfiles: int, numprocs: int, myid: int) -> Tuple[int, int]:
    """Work split of the reference (slicer-v2.cpp:162-175): contiguous sub-files per rank, remainder to the LAST rank;
    with more ranks than files only the last rank works."""
    intdiv, remaindiv = divmod(numfiles, numprocs)
    if myid != numprocs - 1:
        return myid * intdiv, (myid + 1) * intdiv
    return myid * intdiv, (myid + 1) * intdiv + remaindiv


def balanced_subfiles(numfiles: int, numprocs: int, myid: int) -> List[int]:
    """Round-robin split used by the GPU driver (every GPU gets work as soon as numfiles >= numprocs)."""
    return list(range(myid, numfiles, numprocs))


def broadcast_unique_id(make_id, rank: int, src: int = 0) -> bytes:
    """ncclUniqueId made on `src` (make_id() -> 128 bytes) and handed to every rank."""
    import torch.distributed as dist

    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    assert isinstance(box[0], (bytes, bytearray)) and len(box[0]) == 128
    return bytes(box[0])


def max_over_ranks(value: float) -> float:
    """Device times are reported as the maximum over ranks (the job ends when the slowest rank ends)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_int64_planes(planes):
    """Host-side model of slicer_reduce (ncclReduce(int64, sum) onto rank 0) for tests: torch int64 all-reduce."""
    import torch
    import torch.distributed as dist

    t = torch.as_tensor(planes).clone()
    dist.reduce(t, dst=0, op=dist.ReduceOp.SUM)
    return t
