"""Multi-GPU plumbing for the frame-parallel path.

Frames are independent (instance norm is per sample), so inference shards by frame with NO data-path collective: one
process per GPU, each with its own native context and a replica of the weights.  torch.distributed is used only for
rendezvous, barriers and reducing timings / checksums (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Tuple


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment, defaulting to a single process."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def frame_shard(num_frames: int, rank: int, world: int) -> range:
    """Contiguous, balanced partition of a frame stream: the first (num_frames % world) ranks get one extra frame."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(max(num_frames, 0), world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def init_process_group(backend: str = "nccl", device=None):
    import torch.distributed as dist
    if not dist.is_initialized():
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return dist


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (the multi-GPU timing rule: report the slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def sum_over_ranks(values, device="cpu"):
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) for v in t]


def aggregate_throughput(frames_this_rank: int, seconds_this_rank: float, device="cpu") -> float:
    """Whole-job frames/s: all ranks' frames over the slowest rank's time."""
    total = sum_over_ranks([frames_this_rank], device)[0]
    slowest = max_over_ranks([seconds_this_rank], device)[0]
    return total / slowest if slowest > 0 else 0.0


def allreduce_sum_(tensor):
    """In-place SUM all-reduce of the flat gradient buffer over every rank (NCCL on GPUs, gloo in the CPU tests); a no-op
    in a single process.  SUM, not mean: Keras differentiates the batch SUM of the (B,) loss vector, so the data-parallel
    gradient of the global batch is the sum of the per-rank gradients."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
        if getattr(tensor, "is_cuda", False):
            # NCCL runs on torch's stream, the native trainer on its own: the optimizer step that follows must not start
            # before the reduced gradient has landed (and the next backward must not overwrite it while it is being read)
            import torch
            torch.cuda.synchronize(tensor.device)
    return tensor
