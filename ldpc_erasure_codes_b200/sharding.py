"""Batch sharding over the GPUs of one box.

Codewords are independent (each `while(1)` iteration of the reference kernels touches only its own
`codeword[]`, OpenCL/device/ldpc_erasure_decoder.cl:27-104), so the multi-GPU form of the path is a
contiguous split of the frame range with NO collective on the data path.  What the ranks do share:
  * the global frame index (it feeds the Threefry counter, so results do not depend on the split),
  * the two cumulative counters the reference's data_out kernel reports (summed at the end),
  * the timing (max over ranks).
`torch.distributed` is only used for those scalars (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous frame range [begin, end) of `rank`: GPU g gets frames [g*B/G, (g+1)*B/G)."""
    if world <= 0 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    return (total * rank) // world, (total * (rank + 1)) // world


def weak_frame_base(per_rank: int, rank: int) -> int:
    """Weak scaling: every rank decodes `per_rank` frames; rank r owns frames [r*per_rank, (r+1)*per_rank)."""
    return per_rank * rank


def reduce_stats(stats: Dict[str, int], group=None) -> Dict[str, int]:
    """Sum of the cumulative counters (frames, ldpc_errors, rs_errors, ...) over the ranks."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return dict(stats)
    keys = sorted(stats)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([int(stats[k]) for k in keys], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def reduce_max(value: float, group=None) -> float:
    """Max over ranks (device-timed milliseconds of the slowest rank)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
