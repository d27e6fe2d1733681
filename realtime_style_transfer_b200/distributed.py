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


# ---- host-side placement for the copy-bound end-to-end path ------------------------------------------------------------------
def _cpus_of_node(node: int):
    cpus = set()
    for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _measure_h2d_gbs(device: int, mbytes: int = 64, repeats: int = 3) -> float:
    import torch
    buf = torch.empty(mbytes << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty_like(buf, device=torch.device("cuda", device))
    dst.copy_(buf, non_blocking=True)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(repeats):
        dst.copy_(buf, non_blocking=True)
    e1.record()
    torch.cuda.synchronize(device)
    return repeats * buf.numel() / (e0.elapsed_time(e1) * 1e6)


def bind_to_gpu_numa_node(local_rank: int, measure: bool = True) -> dict:
    """Pins this process -- and therefore the pinned host buffers it allocates afterwards (first touch) -- to the CPUs next to its
    GPU, so that the H2D / D2H copies of 8 ranks do not all cross the socket interconnect (round 1: 137 GB/s aggregate at 8
    GPUs, no gain from 2 to 4).  Sources, in order: sysfs ``numa_node`` of the GPU's PCI function; NVML's CPU affinity of the
    device; and, when the container hides both, a MEASUREMENT: a small pinned buffer is allocated under each NUMA node's CPUs
    in turn and the node with the fastest host-to-device copy wins.  Returns what it did (reported by bench.py)."""
    import os
    allowed = os.sched_getaffinity(0)
    info = {"node": None, "method": None}
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in visible.split(",") if v.strip()]
        index = int(ids[local_rank]) if len(ids) > local_rank and ids[local_rank].strip().isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:                 # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        try:
            node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        except OSError:
            node = -1
        if node >= 0:
            cpus = _cpus_of_node(node) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {"node": node, "cpus": len(cpus), "method": "sysfs"}
        try:
            words = (max(allowed) // 64) + 1
            mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1} & allowed
            if cpus and cpus != allowed:
                os.sched_setaffinity(0, cpus)
                return {"node": None, "cpus": len(cpus), "method": "nvml cpu affinity"}
        except Exception:                               # noqa: BLE001 - not supported in this container
            pass
    except Exception as e:                              # noqa: BLE001 - best effort
        info["why"] = type(e).__name__
    if not measure:
        return info
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
        groups = [(n, _cpus_of_node(n) & allowed) for n in nodes]
        groups = [(n, c) for n, c in groups if c]
        if len(groups) < 2:
            return {**info, "method": "single NUMA node", "nodes": len(groups)}
        rates = {}
        for n, cpus in groups:
            os.sched_setaffinity(0, cpus)
            rates[n] = _measure_h2d_gbs(local_rank)
        best = max(rates, key=rates.get)
        os.sched_setaffinity(0, dict(groups)[best])
        return {"node": best, "cpus": len(dict(groups)[best]), "method": "measured H2D per node",
                "h2d_gbs_by_node": {str(k): round(v, 1) for k, v in rates.items()}}
    except Exception as e:                              # noqa: BLE001
        os.sched_setaffinity(0, allowed)
        return {**info, "method": "none", "why2": type(e).__name__}
