#!/usr/bin/env python
"""Headline benchmark: frames/sec of the rst-960-120-128-17 stylization forward (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

A step = one transfer-network forward over a batch of 8 synthetic 17-channel 480x960 frames per GPU
(BASELINE.json configs[1]); frames are independent, so ranks shard frames with no collective (weak scaling).
`value` is timed with inputs resident in HBM; `e2e` goes through the host-buffer C-ABI entry point with the
H2D copy of the frames and the D2H copy of the stylised images inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SPEC = "rst-960-120-128-17"
BATCH = int(os.environ.get("RST_BENCH_BATCH", "8"))     # BASELINE config: 8 frames per GPU; the override is for experiments
GFLOP_PER_FRAME = 127.269           # convolution MACs x2, true channel counts
TRUNK_CONV_GFLOP = 2 * 4.247        # one 128->128 3x3 conv at 120x240, per frame
RES0_CONV0_GFLOP = 2 * 1.062


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples if len(s) >= 7 for i in range(4) if s[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def synthetic_inputs(cfg, batch, seed):
    """Channel-wise synthetic G-buffer in ShapeConfig.channels order."""
    rng = np.random.default_rng(seed)
    h, w = cfg.input_shape["content"][:2]
    parts = []
    for name, n in cfg.channels:
        if name == "FinalImage":
            t = rng.uniform(0, 4, (batch, h, w, n))
        elif name == "ViewNormal":
            t = rng.standard_normal((batch, h, w, n))
            t /= np.maximum(np.linalg.norm(t, axis=-1, keepdims=True), 1e-6)
        elif name == "SceneDepth":
            t = rng.uniform(10, 1e4, (batch, h, w, n))
        else:
            t = rng.uniform(0, 1, (batch, h, w, n))
        parts.append(t.astype(np.float32))
    return np.concatenate(parts, axis=-1)


def run_reference(args, rank, world):
    """CPU arm: the oracle port (PyTorch-CPU restatement of the reference; TensorFlow is not installable here)."""
    if rank != 0:
        return
    import torch
    from oracle import rst_oracle as O
    from realtime_style_transfer_b200.shape_config import ShapeConfig
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = ShapeConfig.from_spec(SPEC)
    spec = O.TransferSpec(cfg.input_shape["content"], cfg.output_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
    weights = O.init_transfer_weights(spec, seed=1)
    content = synthetic_inputs(cfg, 1, 0)
    params = np.random.default_rng(1).uniform(0.3, 1.0, (1, 1, spec.num_style_parameters)).astype(np.float32)
    steps = min(args.steps, 12)
    warm = min(args.warmup, 2)
    with torch.no_grad():
        for _ in range(warm):
            O.transfer_forward(spec, weights, content, params)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.transfer_forward(spec, weights, content, params)
        dt = time.perf_counter() - t0
    fps = steps / dt
    sample = f"{steps} steps of 1 frame (a 1/8 sample of the batch-8 step), oracle fp32 on {cores} host threads"
    print(json.dumps({
        "impl": "reference", "metric": "frames/sec rst-960-120-128-17", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{SPEC} single-style inference forward, 1 frame per step on CPU", "batch_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample(cfg, budget_s=20.0):
    import torch
    from oracle import rst_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = O.TransferSpec(cfg.input_shape["content"], cfg.output_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
    weights = O.init_transfer_weights(spec, seed=1)
    content = synthetic_inputs(cfg, 1, 0)
    params = np.random.default_rng(1).uniform(0.3, 1.0, (1, 1, spec.num_style_parameters)).astype(np.float32)
    with torch.no_grad():
        O.transfer_forward(spec, weights, content, params)
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < budget_s and n < 40):
            O.transfer_forward(spec, weights, content, params)
            n += 1
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{n} single-frame forwards of {SPEC} (oracle: PyTorch-CPU fp32 restatement; TF not installable)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from realtime_style_transfer_b200 import _native
    from realtime_style_transfer_b200 import distributed as rdist
    from realtime_style_transfer_b200._plan import PredictorPlan, TransferPlan
    from realtime_style_transfer_b200.shape_config import ShapeConfig

    assert args.warmup >= 3, "timing rules: at least 3 warm-up steps"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = ShapeConfig.from_spec(SPEC)
    in_shape, out_shape = cfg.input_shape["content"], cfg.output_shape
    plan = TransferPlan(in_shape, out_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
    weights = plan.initial_weights(np.random.default_rng(1))
    # randomised BatchNorm moving statistics so inference-mode BN does real work
    rng = np.random.default_rng(2)
    for k in weights:
        if k.endswith("moving_mean"):
            weights[k] = (0.05 * rng.standard_normal(weights[k].shape)).astype(np.float32)
        if k.endswith("moving_variance"):
            weights[k] = (0.5 + rng.uniform(size=weights[k].shape)).astype(np.float32)
    precision = _native.PRECISION_BF16 if args.precision == "bf16" else _native.PRECISION_FP32
    ctx = _native.NativeContext(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=cfg.bottleneck_res_y,
                                bottleneck_num_filters=cfg.bottleneck_num_filters, num_styles=1, max_batch=BATCH,
                                precision=precision, device=local_rank)
    ctx.set_weights(weights)

    # style parameters from the style predictor on one synthetic style image, replicated over the batch
    pplan = PredictorPlan(out_shape, "MOBILE_NET", plan.num_style_parameters)
    pctx = _native.NativeContext(extractor=_native.EXTRACTOR_MOBILE_NET, style_shape=out_shape[:2],
                                 predictor_num_params=plan.num_style_parameters, max_batch=1, device=local_rank)
    pctx.set_weights(pplan.initial_weights(np.random.default_rng(3)))
    style = np.random.default_rng(4).uniform(0, 1, (1,) + out_shape).astype(np.float32)
    params_h = np.repeat(pctx.predict_style_host(style)[:, None, :], BATCH, axis=0)
    pctx.close()

    content_h = torch.from_numpy(synthetic_inputs(cfg, BATCH, seed=rank)).pin_memory()
    params_pin = torch.from_numpy(np.ascontiguousarray(params_h)).pin_memory()
    out_pin = torch.empty((BATCH,) + out_shape, dtype=torch.float32).pin_memory()
    d_content = content_h.to(dev)          # 250 MB fp32: larger than the 126 MB L2, no flush needed between steps
    d_params = params_pin.to(dev)
    d_out = torch.empty((BATCH,) + out_shape, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)          # a non-default stream: the forward is captured once and replayed as a CUDA graph
    torch.cuda.set_stream(stream)

    def step():
        ctx.transfer_forward_device(d_content.data_ptr(), d_params.data_ptr(), None, d_out.data_ptr(), BATCH, stream.cuda_stream)

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # the library runs a buffer set eagerly the first time and captures its CUDA graph when it comes back: prime both, so that
    # the W warm-up steps and the timed steps all replay the graph whatever W is
    for _ in range(2):
        step()
    for _ in range(args.warmup):
        step()
    sync_all()
    launches_per_step = ctx.last_launch_count()

    # ---- timed region: device-resident inputs (value); the forward replays as one CUDA graph ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    sync_all()
    ms = e0.elapsed_time(e1)

    # ---- per-kernel pass for the roofline: the same K steps again, launched eagerly with a CUDA-event pair around
    # every kernel on the launch stream (events cannot be interleaved with a graph replay) ----
    ctx.profile(True)
    ctx.profile_reset()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(stream)
    for _ in range(args.steps):
        step()
    p1.record(stream)
    sync_all()
    eager_ms = p0.elapsed_time(p1)
    groups = ctx.profile_groups()
    ctx.profile(False)

    # ---- end-to-end: HOST buffers through the public streaming entry point (rst_transfer_submit_host / _wait):
    # every step copies its 250 MB of fp32 frames H2D and its 44 MB of stylised frames D2H inside the timed region;
    # the copies of neighbouring steps overlap the forward (two staging slots), as in a video loop with prefetch.
    e2e_steps = max(4, args.steps)
    content_np, params_np = content_h.numpy(), params_pin.numpy()
    outs = [out_pin.numpy(), torch.empty((BATCH,) + out_shape, dtype=torch.float32).pin_memory().numpy()]

    def e2e_run(n):
        prev = None
        for i in range(n):
            t = ctx.transfer_submit_host(content_np, params_np, None, outs[i % 2])
            if prev is not None:
                ctx.transfer_wait(prev)
            prev = t
        ctx.transfer_wait(prev)

    e2e_run(4)          # both staging slots: first use runs eagerly, second captures the CUDA graph
    sync_all()
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    # the synchronous single-call variant for comparison (no overlap)
    lib = ctx.lib
    import ctypes as C
    t1 = time.perf_counter()
    for _ in range(3):
        rc = lib.rst_transfer_forward_host(ctx.handle, content_np.ctypes.data_as(C.c_void_p), params_np.ctypes.data_as(C.c_void_p),
                                           None, outs[0].ctypes.data_as(C.c_void_p), BATCH)
        assert rc == 0, lib.rst_last_error(ctx.handle)
    e2e_sync_fps = world * BATCH * 3 / (time.perf_counter() - t1)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # multi-GPU timing rule: every rank did the same number of frames; report against the slowest rank
    ms_max, e2e_ms_max = rdist.max_over_ranks([ms, e2e_s * 1e3], device=dev)
    checksum = float(d_out.double().mean())

    if rank == 0:
        pk = peaks()
        fps = world * BATCH * args.steps / (ms_max / 1e3)
        e2e_fps = world * BATCH * e2e_steps / (e2e_ms_max / 1e3)
        # dominant kernel = largest share of the step among this library's kernel groups
        total_group_ms = sum(v[0] for v in groups.values()) or 1.0
        dominant = max(groups.items(), key=lambda kv: kv[1][0])
        shares = {k: round(v[0] / total_group_ms, 4) for k, v in sorted(groups.items(), key=lambda kv: -kv[1][0])}
        roof = None
        if "conv3x3_umma" in groups and groups["conv3x3_umma"][1]:
            g_ms, g_n = groups["conv3x3_umma"]
            # per step: 9 convs 128->128 and one 32->128 (executed as 64->128: padded input channels are not counted)
            gflop_per_launch = BATCH * (9 * TRUNK_CONV_GFLOP + RES0_CONV0_GFLOP) / 10.0
            achieved = gflop_per_launch / (g_ms / g_n)       # GFLOP / ms = TFLOP/s
            roof = {"kernel": "halo_gemm2_kernel (2-CTA tcgen05 halo GEMM, cta_group::2, UMMA 256x128x16; residual bottleneck "
                              "convs 128->128; the 32->128 first conv runs the 1-CTA halo_gemm_kernel<128,64,...>)",
                    "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16_tflops"],
                    # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full capture
                    # (profiles/r01_ncu_full_final_raw_subset.csv: 61.3 MB read + 19.5..21.4 MB written); algorithmic
                    # in+out is 118 MB, most of the output stays in the 126 MB L2 for the norm pass that follows
                    "traffic": 81.8e6, "peak_source": pk["source"] + ", burst figure (kernel timed alone between events)",
                    "avg_launch_ms": g_ms / g_n, "launches": g_n, "share_of_step": shares.get("conv3x3_umma")}
        whole = {"achieved_tflops": fps * GFLOP_PER_FRAME / 1e3 / world, "peak": pk["bf16_tflops_sustained"],
                 "frac_of_sustained_bf16": fps * GFLOP_PER_FRAME / 1e3 / world / pk["bf16_tflops_sustained"]}
        line = {
            "metric": "frames/sec rst-960-120-128-17", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"{SPEC} single-style inference forward, batch {BATCH} frames per GPU, device-resident "
                                   "fp32 NHWC frames (250 MB per batch > L2, no flush needed)",
                       "batch_per_gpu": BATCH, "frames_sharded_across_gpus": True, "collective": "none"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(content_np.nbytes + params_np.nbytes),
                    "d2h_bytes_per_step": int(outs[0].nbytes), "steps": e2e_steps,
                    "api": "rst_transfer_submit_host/rst_transfer_wait (pinned host buffers, 2 batches in flight)",
                    "synchronous_single_call_fps": e2e_sync_fps},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler.summary(),
            "roofline": roof,
            "whole_net": whole,
            "kernel_shares": shares,
            "sum_kernel_ms_per_step": total_group_ms / args.steps,
            "eager_profiled_ms_per_step": eager_ms / args.steps,
            "dominant_group": dominant[0],
            "checksum": checksum,
        }
        if not args.no_cpu_baseline and world == 1:       # the CPU arm is timed at N=1 only (its host cores are shared at N>1)
            line["cpu_baseline"] = cpu_baseline_sample(cfg)
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
