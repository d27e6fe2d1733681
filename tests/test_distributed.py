"""CPU tests of the multi-GPU host logic with world_size-2 gloo process groups (no GPU, no NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from realtime_style_transfer_b200 import distributed as D


def test_frame_shard_partitions_exactly():
    for n in (0, 1, 7, 8, 17, 240):
        for world in (1, 2, 4, 8):
            shards = [D.frame_shard(n, r, world) for r in range(world)]
            flat = [i for s in shards for i in s]
            assert flat == list(range(n))                               # contiguous, ordered, no gaps / overlaps
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        D.frame_shard(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init_process_group("gloo")
    assert D.env_rank() == (rank, world, rank)
    frames = D.frame_shard(17, rank, world)
    # every rank "processes" its shard; rank 1 is slower
    seconds = 1.0 + rank
    mx = D.max_over_ranks([seconds, float(len(frames))])
    total = D.sum_over_ranks([len(frames)])[0]
    fps = D.aggregate_throughput(len(frames), seconds)
    dist.barrier()
    q.put((rank, mx, total, fps, list(frames)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing_rule():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, mx0, tot0, fps0, f0), (r1, mx1, tot1, fps1, f1) = results
    assert f0 + f1 == list(range(17)) and len(f0) == 9 and len(f1) == 8
    assert mx0 == mx1 == [2.0, 9.0]               # the slowest rank's time, the largest shard
    assert tot0 == tot1 == 17.0
    assert fps0 == fps1 == pytest.approx(17.0 / 2.0)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    D.init_process_group("gloo")
    # each rank holds the gradient of its own batch shard in one flat buffer (what rst_train_gradients exposes)
    g = torch.arange(10, dtype=torch.float32) * (rank + 1)
    out = D.allreduce_sum_(g)
    assert out is g                                     # in place: the native RMSprop reads the same buffer
    q.put((rank, g.tolist()))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_is_a_sum():
    """Keras differentiates the batch SUM of the loss vector, so data-parallel gradients add up (no averaging)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [3.0 * i for i in range(10)]
    assert results[0][1] == expect and results[1][1] == expect


def test_allreduce_is_a_noop_without_a_process_group():
    g = torch.ones(4)
    assert D.allreduce_sum_(g).tolist() == [1.0] * 4


def test_numa_binding_is_best_effort_without_a_gpu():
    """bind_to_gpu_numa_node never raises and never leaves the process with an empty CPU set (no GPU / no NVML here)."""
    import os
    from realtime_style_transfer_b200 import distributed as D
    before = os.sched_getaffinity(0)
    info = D.bind_to_gpu_numa_node(0, measure=False)
    assert isinstance(info, dict) and "method" in info
    assert os.sched_getaffinity(0) and os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
