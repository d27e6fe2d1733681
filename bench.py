#!/usr/bin/env python
"""Headline benchmark: frames/sec of the rst-960-120-128-17 stylization forward (BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on the host cores

A step = one transfer-network forward over a batch of 8 synthetic 17-channel 480x960 frames per GPU
(BASELINE.json configs[1]); frames are independent, so ranks shard frames with no collective (weak scaling).

`value`  device-resident: the G-buffer batch is already in HBM (float16, the element type of the e2e path) and the uint8
         image stays there; timed with CUDA events over K replays of the captured forward.
`e2e`    the same frames through the public host entry points rst_transfer_submit_host_typed / rst_transfer_wait: every
         step copies its float16 G-buffer batch from pinned host memory and reads the uint8 image back, inside the timed region.
`fp32_io`   both numbers again through the float32 drop-in entry points (float32 G-buffer in, float32 image out).
`sustained` >= 3 s of back-to-back replays after the short timed region (power-capped clocks), against the sustained peak.
`configs`   BASELINE configs 1 (rst-960-120-32-3, fp32, batch 1) and 3 (rst-960-120-128-18, two styles blended).
`training`  BASELINE config 4: one data-parallel training step (forward, VGG loss, backward, NCCL all-reduce, RMSprop).
"""
from __future__ import annotations

import argparse
import glob
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
GFLOP_PER_FRAME_18 = 129.66         # rst-960-120-128-18
GFLOP_PER_FRAME_32_3 = 18.98        # rst-960-120-32-3
TRUNK_CONV_GFLOP = 2 * 4.247        # one 128->128 3x3 conv at 120x240, per frame
RES0_CONV0_GFLOP = 2 * 1.062
TRAIN_TFLOP_PER_SAMPLE = 1.56


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def traffic_record():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the tracked ncu summary."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(path):
        return None, None
    rec = json.load(open(path))
    return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"]), f"profiles/roofline_traffic.json ({rec.get('source', '')})"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed regions."""

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
                    self.samples.append((time.perf_counter(), [f.strip() for f in out.split(",")]))
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self, t0=None, t1=None):
        rows = [s for t, s in self.samples if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        sm = [float(s[0]) for s in rows if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in rows if len(s) > 1 and s[1].replace(".", "").isdigit()]
        pw = [float(s[2]) for s in rows if len(s) > 2 and s[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in rows if len(s) >= 7 for i in range(4) if s[3 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


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


def randomise_bn(weights, seed=2):
    """randomised BatchNorm moving statistics so inference-mode BN does real work"""
    rng = np.random.default_rng(seed)
    for k in weights:
        if k.endswith("moving_mean"):
            weights[k] = (0.05 * rng.standard_normal(weights[k].shape)).astype(np.float32)
        if k.endswith("moving_variance"):
            weights[k] = (0.5 + rng.uniform(size=weights[k].shape)).astype(np.float32)
    return weights


# ---------------------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------------------
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
    # a step of this arm is ONE frame of the batch-8 step (a 1/8 sample, ~0.2 s on 16 cores): --steps / --warmup are honoured
    # as given up to a bound that keeps the run within a few minutes
    steps, warm = min(args.steps, 600), min(args.warmup, 50)
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
        "steps": steps, "warmup": warm, "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{SPEC} single-style inference forward, 1 frame per step on CPU", "batch_per_step": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_sample(cfg, weights, content16, params, gpu_u8_frame0, extra_cases, budget_s=20.0):
    """The oracle port timed on single frames of the same workload (rank 0, N = 1 only).  While it is at hand it is also the
    CHECKER of this run's GPU results: frame 0 of the headline workload and one frame of BASELINE configs 1 and 3 are
    recomputed on the CPU and compared (``parity``); nothing the GPU arm measures runs through it."""
    import torch
    from oracle import rst_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    spec = O.TransferSpec(cfg.input_shape["content"], cfg.output_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
    x = content16[:1].astype(np.float32)
    parity = {}
    with torch.no_grad():
        ref = O.transfer_forward(spec, weights, x, params[:1]).numpy()
        n, t0 = 0, time.perf_counter()
        while n < 3 or (time.perf_counter() - t0 < budget_s and n < 40):
            O.transfer_forward(spec, weights, x, params[:1])
            n += 1
        dt = time.perf_counter() - t0
        e = np.abs(gpu_u8_frame0.astype(int) - (ref[0] * 255).astype(int))
        parity["headline_frame0_uint8_levels"] = {"mean": float(e.mean()), "p99": float(np.quantile(e, 0.99)), "max": int(e.max()),
                                                  "pass": bool(np.quantile(e, 0.99) <= 6)}
        for name, case in extra_cases.items():
            s2 = O.TransferSpec(case["in_shape"], case["out_shape"], case["res_y"], case["filters"], case["styles"])
            r = O.transfer_forward(s2, case["weights"], case["content"], case["params"], case.get("style_weights")).numpy()
            err = np.abs(case["gpu"] - r)
            rel = float(np.sqrt((err.astype(np.float64) ** 2).sum() / (r.astype(np.float64) ** 2).sum()))
            parity[name] = {"max_abs": float(err.max()), "rel_l2": rel, "tolerance": case["tol_desc"], "pass": bool(case["gate"](err, rel))}
    return ({"value": n / dt, "unit": "frames/s", "cores": cores, "kind": "port",
             "sample": f"{n} single-frame forwards of {SPEC} (oracle: PyTorch-CPU fp32 restatement; TF not installable)"}, parity)


# ---------------------------------------------------------------------------------------------------------------------------
# GPU arm helpers
# ---------------------------------------------------------------------------------------------------------------------------
def timed_replays(torch, stream, step, steps, sync_all):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    sync_all()
    return e0.elapsed_time(e1)


def e2e_loop(ctx, content, params, weights, outs, n):
    prev = None
    for i in range(n):
        t = ctx.transfer_submit_host(content, params, weights, outs[i % 2])
        if prev is not None:
            ctx.transfer_wait(prev)
        prev = t
    ctx.transfer_wait(prev)


def bench_training(torch, dist, _native, _plan, rdist, dev, local_rank, rank, world, steps=3, warmup=2, tf32_math=False):
    """BASELINE config 4: predictor + transfer net forward/backward in training mode, VGG16 Gram/content loss (x3 forward,
    x1 backward), gradient all-reduce (SUM) over NCCL, RMSprop; 8 samples per GPU at 480x960, 17 channels, 128 filters."""
    b, h, w, f = 8, 480, 960, 128
    in_shape, out_shape = (h, w, 17), (h, w, 3)
    tplan = _plan.TransferPlan(in_shape, out_shape, h // 4, f, 1)
    pplan = _plan.PredictorPlan(out_shape, "MOBILE_NET", tplan.num_style_parameters, 100)
    rng = np.random.default_rng(1234)
    weights = dict(tplan.initial_weights(rng))
    weights.update(pplan.initial_weights(rng))
    tr = _native.NativeTrainer(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=h // 4, bottleneck_num_filters=f,
                               max_batch=b, extractor=_native.EXTRACTOR_MOBILE_NET, style_shape=out_shape[:2], device=local_rank)
    if tf32_math:
        tr.set_math(_native.PRECISION_TF32)       # plain tf32 operands: TensorFlow's own default for float32 models on Ampere and later
    tr.model.set_weights(weights)
    vgg, cin = {}, 3
    for blk, n, co in (("block1", 2, 64), ("block2", 2, 128), ("block3", 3, 256), ("block4", 3, 512), ("block5", 3, 512)):
        for i in range(1, n + 1):
            vgg[f"{blk}_conv{i}/kernel"] = rng.normal(0, np.sqrt(2.0 / (9 * cin)), (3, 3, cin, co)).astype(np.float32)
            vgg[f"{blk}_conv{i}/bias"] = np.zeros(co, np.float32)
            cin = co
    tr.loss.set_weights(vgg)
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    content = torch.rand((b,) + in_shape, device=dev, generator=g)
    style = torch.rand((b,) + out_shape, device=dev, generator=g)
    gt = torch.rand((b,) + out_shape, device=dev, generator=g)
    losses = torch.empty((b, 4), device=dev)
    from realtime_style_transfer_b200.models.styleTransferTrainingModel import _DeviceArray
    grads = torch.as_tensor(_DeviceArray(tr.gradients_ptr(), tr.num_gradient_elements), device=dev)
    tstream = torch.cuda.ExternalStream(tr.stream_ptr(), device=dev)
    cur = torch.cuda.current_stream(dev)
    ar_ms = []

    def step(timed):
        tr.forward_backward(content.data_ptr(), style.data_ptr(), gt.data_ptr(), style.data_ptr(), losses.data_ptr(), b)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(cur)
        rdist.allreduce_sum_(grads)                  # NCCL all-reduce (SUM) of the flat gradient buffer; no-op at N = 1
        a1.record(cur)
        torch.cuda.synchronize(dev)
        if timed:
            ar_ms.append(a0.elapsed_time(a1))
        tr.apply_gradients()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    for _ in range(warmup):
        step(False)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(tstream)
    for _ in range(steps):
        step(True)
    e1.record(tstream)
    sync_all()
    ms = e0.elapsed_time(e1)
    ms_max, ar_max = rdist.max_over_ranks([ms, float(np.mean(ar_ms))], device=dev)
    loss_rows = losses.cpu().numpy()
    launches = int(tr.lib.rst_last_launch_count(tr.model.handle))
    tr.close()
    del grads
    torch.cuda.empty_cache()
    per_step = ms_max / steps
    return {
        "workload": "training step: MobileNetV3 style predictor + rst-960-120-128-17 transfer net forward/backward (BatchNorm on "
                    "batch statistics), VGG16 Gram/content/TV loss x3 forward + x1 backward, RMSprop; 8 samples per GPU",
        "samples_per_s": b * world / (per_step / 1e3), "ms_per_step": per_step, "steps": steps, "warmup": warmup,
        "batch_per_gpu": b, "n_gpus": world,
        "allreduce_ms": ar_max, "allreduce_elements": int(tr.num_gradient_elements),
        "collective": "NCCL all-reduce (SUM) of the flat fp32 gradient buffer" if world > 1 else "none (single process)",
        "achieved_tflops_per_gpu": TRAIN_TFLOP_PER_SAMPLE * b / (per_step / 1e3),
        "math": "plain tf32 operands on tcgen05 for the trunk and VGG convolutions (rst_train_set_math(TF32)), fp32 elsewhere" if tf32_math else
                "fp32-accurate (error-compensated split tf32 on tcgen05 for the trunk and VGG convolutions, fp32 elsewhere)",
        "timing": "CUDA events on the trainer's stream around the timed steps, max over ranks",
        "loss_mean": float(loss_rows[:, 0].mean()), "style_loss_mean": float(loss_rows[:, 2].mean()),
        "gpu_launches_per_step": launches,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-training", action="store_true", help="skip the config-4 training record")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / fp32_io / configs records (kernel experiments)")
    ap.add_argument("--sustained-seconds", type=float, default=3.2)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from realtime_style_transfer_b200 import _native, _plan
    from realtime_style_transfer_b200 import distributed as rdist
    torch.cuda.set_device(local_rank)
    numa = rdist.bind_to_gpu_numa_node(local_rank)      # before any pinned allocation
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
    weights = randomise_bn(plan.initial_weights(np.random.default_rng(1)))
    ctx = _native.NativeContext(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=cfg.bottleneck_res_y,
                                bottleneck_num_filters=cfg.bottleneck_num_filters, num_styles=1, max_batch=BATCH,
                                precision=_native.PRECISION_BF16, device=local_rank)
    ctx.set_weights(weights)

    # style parameters from the style predictor on one synthetic style image, replicated over the batch
    pplan = PredictorPlan(out_shape, "MOBILE_NET", plan.num_style_parameters)
    pctx = _native.NativeContext(extractor=_native.EXTRACTOR_MOBILE_NET, style_shape=out_shape[:2],
                                 predictor_num_params=plan.num_style_parameters, max_batch=1, device=local_rank)
    pctx.set_weights(pplan.initial_weights(np.random.default_rng(3)))
    style = np.random.default_rng(4).uniform(0, 1, (1,) + out_shape).astype(np.float32)
    params_h = np.repeat(pctx.predict_style_host(style)[:, None, :], BATCH, axis=0)
    pctx.close()

    content32 = synthetic_inputs(cfg, BATCH, seed=rank)
    content16_pin = torch.from_numpy(content32.astype(np.float16)).pin_memory()       # what the EXR planes hold
    content32_pin = torch.from_numpy(content32).pin_memory()
    params_pin = torch.from_numpy(np.ascontiguousarray(params_h)).pin_memory()
    d_content16 = content16_pin.to(dev)
    d_content32 = content32_pin.to(dev)
    d_params = params_pin.to(dev)
    d_out8 = torch.empty((BATCH,) + out_shape, dtype=torch.uint8, device=dev)
    d_out32 = torch.empty((BATCH,) + out_shape, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(dev)          # a non-default stream: the forward is captured once and replayed as a CUDA graph
    torch.cuda.set_stream(stream)

    def step():
        ctx.transfer_forward_device(d_content16.data_ptr(), d_params.data_ptr(), None, d_out8.data_ptr(), BATCH, stream.cuda_stream,
                                    content_dtype=_native.DTYPE_F16, out_dtype=_native.DTYPE_U8)

    def step32():
        ctx.transfer_forward_device(d_content32.data_ptr(), d_params.data_ptr(), None, d_out32.data_ptr(), BATCH, stream.cuda_stream)

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
    t_value0 = time.perf_counter()
    ms = timed_replays(torch, stream, step, args.steps, sync_all)

    # ---- per-kernel pass for the roofline: the same K steps again, launched eagerly with a CUDA-event pair around
    # every kernel on the launch stream (events cannot be interleaved with a graph replay) ----
    ctx.profile(True)
    ctx.profile_reset()
    eager_ms = timed_replays(torch, stream, step, args.steps, sync_all)
    groups = ctx.profile_groups()
    ctx.profile(False)

    # ---- sustained regime: seconds of back-to-back replays of the same step under the power cap ----
    sustained = None
    if not args.no_extras:
        n_sus = max(args.steps, int(args.sustained_seconds * 1e3 / max(ms / args.steps, 1e-3)) + 1)
        t_s0 = time.perf_counter()
        sus_ms = timed_replays(torch, stream, step, n_sus, sync_all)
        t_s1 = time.perf_counter()
        sustained = {"steps": n_sus, "seconds": sus_ms / 1e3, "ms_per_step": sus_ms / n_sus,
                     "clocks": sampler.summary(t_s0 + 0.5, t_s1)}

    # ---- frame latency at smaller batches (the reference is a per-frame renderer: predict_video_using_checkpoint.py runs
    # batch_size=1): same context, same device buffers, first b frames ----
    batch_sweep = {}
    if not args.no_extras:
        for b in (1, 2, 4):
            if b >= BATCH:
                continue

            def step_b(b=b):
                ctx.transfer_forward_device(d_content16.data_ptr(), d_params.data_ptr(), None, d_out8.data_ptr(), b, stream.cuda_stream,
                                            content_dtype=_native.DTYPE_F16, out_dtype=_native.DTYPE_U8)
            for _ in range(2 + 3):
                step_b()
            ms_b = timed_replays(torch, stream, step_b, args.steps, sync_all)
            ms_b = rdist.max_over_ranks([ms_b], device=dev)[0]
            batch_sweep[f"batch_{b}"] = {"ms_per_forward": ms_b / args.steps, "frames_per_s": world * b * args.steps / (ms_b / 1e3)}

    # ---- end-to-end: HOST buffers through the public streaming entry point (rst_transfer_submit_host_typed / _wait): every
    # step copies its float16 G-buffer batch H2D and its uint8 stylised frames D2H inside the timed region; the copies of
    # neighbouring steps overlap the forward (two staging slots), as in a video loop with prefetch.
    e2e_steps = max(4, args.steps)
    content16_np, content32_np, params_np = content16_pin.numpy(), content32_pin.numpy(), params_pin.numpy()
    outs8 = [torch.empty((BATCH,) + out_shape, dtype=torch.uint8).pin_memory().numpy() for _ in range(2)]
    e2e_loop(ctx, content16_np, params_np, None, outs8, 4)      # both staging slots: first use runs eagerly, second captures
    sync_all()
    t0 = time.perf_counter()
    e2e_loop(ctx, content16_np, params_np, None, outs8, e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t_e2e_end = time.perf_counter()
    frame0_u8 = outs8[(e2e_steps - 1) % 2][0].copy()
    # the synchronous single-call variant for comparison (no overlap)
    t1 = time.perf_counter()
    for _ in range(3):
        ctx.transfer_forward_host(content16_np, params_np, None, out_dtype=np.uint8)
    e2e_sync_s = (time.perf_counter() - t1) / 3

    # ---- the float32 drop-in entry points on the same frames ----
    fp32_io = None
    if not args.no_extras:
        for _ in range(2 + 3):
            step32()
        ms32 = timed_replays(torch, stream, step32, args.steps, sync_all)
        outs32 = [torch.empty((BATCH,) + out_shape, dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        n32 = max(4, args.steps // 2)
        e2e_loop(ctx, content32_np, params_np, None, outs32, 4)
        sync_all()
        t0 = time.perf_counter()
        e2e_loop(ctx, content32_np, params_np, None, outs32, n32)
        torch.cuda.synchronize(dev)
        e2e32_s = time.perf_counter() - t0
        d = np.abs((outs32[(n32 - 1) % 2][0] * 255).astype(int) - frame0_u8.astype(int))
        fp32_io = {"ms": ms32, "e2e_s": e2e32_s, "e2e_steps": n32, "h2d": int(content32_np.nbytes + params_np.nbytes),
                   "d2h": int(outs32[0].nbytes), "uint8_vs_float_levels_max": int(d.max())}
        del outs32
    sampler_main = sampler.summary(t_value0, t_e2e_end)      # the device-timed region, its eager repeat, the sustained run and e2e
    checksum = float(d_out8.double().mean()) / 255.0

    # ---- BASELINE configs 1 and 3 ----
    extra_cases, config_recs = {}, {}
    if not args.no_extras:
        # config 3: rst-960-120-128-18, two styles blended per pixel by the weight map
        cfg3 = ShapeConfig(num_styles=2, num_channels=18)
        plan3 = TransferPlan(cfg3.input_shape["content"], cfg3.output_shape, 120, 128, 2)
        w3 = randomise_bn(plan3.initial_weights(np.random.default_rng(5)))
        ctx3 = _native.NativeContext(in_shape=cfg3.input_shape["content"], out_shape=cfg3.output_shape, bottleneck_res_y=120,
                                     bottleneck_num_filters=128, num_styles=2, max_batch=BATCH, precision=_native.PRECISION_BF16,
                                     device=local_rank)
        ctx3.set_weights(w3)
        c3 = synthetic_inputs(cfg3, BATCH, seed=7 + rank)
        p3 = np.random.default_rng(8).uniform(0.3, 1.2, (BATCH, 2, plan3.num_style_parameters)).astype(np.float32)
        yy, xx = np.meshgrid(np.linspace(0, 1, 480, dtype=np.float32), np.linspace(0, 1, 960, dtype=np.float32), indexing="ij")
        sw3 = np.repeat((0.5 + 0.5 * np.sin(6.28 * xx) * np.cos(3.14 * yy))[None, :, :, None], BATCH, axis=0).astype(np.float32)
        d_c3, d_p3, d_w3 = (torch.from_numpy(a).to(dev) for a in (c3.astype(np.float16), p3, sw3))
        d_o3 = torch.empty((BATCH, 480, 960, 3), dtype=torch.float32, device=dev)

        def step3():
            ctx3.transfer_forward_device(d_c3.data_ptr(), d_p3.data_ptr(), d_w3.data_ptr(), d_o3.data_ptr(), BATCH, stream.cuda_stream,
                                         content_dtype=_native.DTYPE_F16)
        for _ in range(2 + 3):
            step3()
        ms3 = timed_replays(torch, stream, step3, args.steps, sync_all)
        ms3 = rdist.max_over_ranks([ms3], device=dev)[0]
        config_recs["config3_dual_style"] = {
            "workload": "rst-960-120-128-18, two predicted style-parameter sets blended per pixel by the weight map, batch 8, bf16, "
                        "device-resident float16 G-buffer, float32 image",
            "frames_per_s": world * BATCH * args.steps / (ms3 / 1e3), "ms_per_step": ms3 / args.steps,
            "achieved_tflops_per_gpu": BATCH * args.steps / (ms3 / 1e3) * GFLOP_PER_FRAME_18 / 1e3}
        extra_cases["config3_dual_style"] = dict(
            in_shape=cfg3.input_shape["content"], out_shape=cfg3.output_shape, res_y=120, filters=128, styles=2, weights=w3,
            content=c3[:1].astype(np.float16).astype(np.float32), params=p3[:1], style_weights=sw3[:1], gpu=d_o3[:1].cpu().numpy(),
            tol_desc="bf16 path: relative L2 <= 2e-2", gate=lambda err, rel: rel <= 2e-2)
        ctx3.close()
        del d_c3, d_o3, d_w3
        # config 1: rst-960-120-32-3, one style, batch 1, fp32 path
        cfg1 = ShapeConfig.from_spec("rst-960-120-32-3", hdr=False)
        plan1 = TransferPlan(cfg1.input_shape["content"], cfg1.output_shape, 120, 32, 1)
        w1 = randomise_bn(plan1.initial_weights(np.random.default_rng(9)))
        ctx1 = _native.NativeContext(in_shape=cfg1.input_shape["content"], out_shape=cfg1.output_shape, bottleneck_res_y=120,
                                     bottleneck_num_filters=32, num_styles=1, max_batch=1, precision=_native.PRECISION_FP32,
                                     device=local_rank)
        ctx1.set_weights(w1)
        c1 = np.random.default_rng(10).uniform(0, 1, (1, 480, 960, 3)).astype(np.float32)
        p1 = np.random.default_rng(11).uniform(0.3, 1.2, (1, 1, plan1.num_style_parameters)).astype(np.float32)
        d_c1, d_p1 = torch.from_numpy(c1).to(dev), torch.from_numpy(p1).to(dev)
        d_o1 = torch.empty((1, 480, 960, 3), dtype=torch.float32, device=dev)

        def step1():
            ctx1.transfer_forward_device(d_c1.data_ptr(), d_p1.data_ptr(), None, d_o1.data_ptr(), 1, stream.cuda_stream)
        for _ in range(2 + 3):
            step1()
        ms1 = timed_replays(torch, stream, step1, args.steps, sync_all)
        ms1 = rdist.max_over_ranks([ms1], device=dev)[0]
        c1_pin, p1_pin = torch.from_numpy(c1).pin_memory().numpy(), torch.from_numpy(p1).pin_memory().numpy()
        o1 = [torch.empty((1, 480, 960, 3), dtype=torch.float32).pin_memory().numpy() for _ in range(2)]
        e2e_loop(ctx1, c1_pin, p1_pin, None, o1, 4)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        e2e_loop(ctx1, c1_pin, p1_pin, None, o1, args.steps)
        torch.cuda.synchronize(dev)
        e1s = time.perf_counter() - t0
        config_recs["config1_fp32_b1"] = {
            "workload": "rst-960-120-32-3, one style, batch 1, fp32 path (max abs error <= 1e-4; residual 3x3 convolutions as split-tf32 "
                        "tcgen05 GEMMs, the 9x9 / strided / transposed layers on CUDA cores), float32 in / out",
            "frames_per_s": world * args.steps / (ms1 / 1e3), "ms_per_frame": ms1 / args.steps,
            "e2e_frames_per_s": world * args.steps / e1s,
            "achieved_tflops_per_gpu": args.steps / (ms1 / 1e3) * GFLOP_PER_FRAME_32_3 / 1e3,
            "gpu_launches_per_frame": int(ctx1.last_launch_count())}
        extra_cases["config1_fp32_b1"] = dict(
            in_shape=cfg1.input_shape["content"], out_shape=cfg1.output_shape, res_y=120, filters=32, styles=1, weights=w1,
            content=c1, params=p1, gpu=d_o1.cpu().numpy(), tol_desc="fp32 path: max abs error <= 1e-4",
            gate=lambda err, rel: float(err.max()) <= 1e-4)
        ctx1.close()

    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- BASELINE config 4: the data-parallel training step (every rank; NCCL all-reduce of the gradients for N > 1) ----
    ctx.close()
    del d_content32, d_content16, d_out32
    torch.cuda.empty_cache()
    training = None
    if not args.no_training:
        torch.cuda.set_stream(torch.cuda.default_stream(dev))
        training = bench_training(torch, dist, _native, _plan, rdist, dev, local_rank, rank, world)
        fast = bench_training(torch, dist, _native, _plan, rdist, dev, local_rank, rank, world, tf32_math=True)
        training["tf32_math"] = {k: fast[k] for k in ("samples_per_s", "ms_per_step", "allreduce_ms", "achieved_tflops_per_gpu", "math", "loss_mean")}

    # multi-GPU timing rule: every rank did the same number of frames; report against the slowest rank
    vals = [ms, e2e_s * 1e3, e2e_sync_s * 1e3, sustained["seconds"] * 1e3 if sustained else 0.0,
            fp32_io["ms"] if fp32_io else 0.0, fp32_io["e2e_s"] * 1e3 if fp32_io else 0.0]
    ms_max, e2e_ms_max, e2e_sync_ms_max, sus_ms_max, ms32_max, e2e32_ms_max = rdist.max_over_ranks(vals, device=dev)

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
            traffic, traffic_src = traffic_record()
            roof = {"kernel": "halo_gemm2_kernel (2-CTA tcgen05 halo GEMM, cta_group::2, UMMA 256x128x16; residual bottleneck "
                              "convs 128->128; the 32->128 first conv runs the 1-CTA halo_gemm_kernel<128,64,...>)",
                    "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16_tflops"],
                    "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": BATCH * 2 * 120 * 240 * 128 * 2,
                    "peak_source": pk["source"] + ", burst figure (kernel timed alone between events)",
                    "avg_launch_ms": g_ms / g_n, "launches": g_n, "share_of_step": shares.get("conv3x3_umma")}
        # HBM-bound passes: algorithmic bytes (each tensor read / written once) over the event-timed group
        px = 120 * 240 * 128 * 2 * BATCH                          # one bottleneck tensor, bytes
        norm_bytes = 5 * 2 * px + 2 * px + 4 * 3 * px + 2 * px + 2 * (2 * px) + BATCH * 480 * 960 * 3 * (4 + 1)
        pack_bytes = BATCH * 480 * 960 * (17 * 2 + 32 * 2)
        hbm = {}
        for grp, nbytes, what in (("cin_apply_bf16", norm_bytes, "13 conditional-instance-norm passes (style affine, ReLU, skip add, sigmoid / uint8 head)"),
                                 ("pack_input", pack_bytes, "float16 NHWC-17 -> packed bf16 stem rows")):
            if grp in groups and groups[grp][0] > 0:
                gbs = nbytes * args.steps / (groups[grp][0] / 1e3) / 1e9
                hbm[grp] = {"what": what, "algorithmic_bytes_per_step": int(nbytes), "ms_per_step": groups[grp][0] / args.steps,
                            "achieved_gbs": gbs, "peak_gbs": pk["hbm_gbs"], "frac": gbs / pk["hbm_gbs"]}
        tflops = fps * GFLOP_PER_FRAME / 1e3 / world
        whole = {"achieved_tflops": tflops, "regime": f"burst ({ms_max / 1e3:.3f} s timed region)", "peak_burst": pk["bf16_tflops"],
                 "frac_of_burst_bf16": tflops / pk["bf16_tflops"]}
        # north_star's "conv path": the convolution kernels alone (every *_umma group of the eager, event-timed pass), i.e. the
        # step without the norm / blend and pack passes
        conv_ms = sum(v[0] for k, v in groups.items() if k.endswith("_umma")) / args.steps
        conv_tflops = BATCH * GFLOP_PER_FRAME / conv_ms if conv_ms > 0 else 0.0          # GFLOP / ms = TFLOP/s
        whole["conv_path"] = {"ms_per_step": conv_ms, "achieved_tflops": conv_tflops, "frac_of_burst_bf16": conv_tflops / pk["bf16_tflops"],
                              "frac_of_sustained_bf16": conv_tflops / pk["bf16_tflops_sustained"]}
        line = {
            "metric": "frames/sec rst-960-120-128-17", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{SPEC} single-style inference forward, batch {BATCH} frames per GPU, device-resident float16 NHWC "
                                   "G-buffer in, uint8 image out (the element types of the e2e path); no L2 flush needed: a step "
                                   "rewrites 1.3 GB of activations, 10x the 126 MB L2",
                       "batch_per_gpu": BATCH, "frames_sharded_across_gpus": True, "collective": "none"},
            "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": int(content16_np.nbytes + params_np.nbytes),
                    "d2h_bytes_per_step": int(outs8[0].nbytes), "steps": e2e_steps,
                    "api": "rst_transfer_submit_host_typed/rst_transfer_wait (pinned host buffers, float16 G-buffer in, uint8 image "
                           "out = trunc(255*y) as the reference's video loop computes it, 2 batches in flight)",
                    "synchronous_single_call_fps": world * BATCH / (e2e_sync_ms_max / 1e3), "numa": numa},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": sampler_main if sampler_main["samples"] else sampler.summary(),
            "roofline": roof,
            "roofline_hbm_passes": hbm,
            "whole_net": whole,
            "kernel_shares": shares,
            "sum_kernel_ms_per_step": total_group_ms / args.steps,
            "eager_profiled_ms_per_step": eager_ms / args.steps,
            "dominant_group": dominant[0],
            "checksum": checksum,
        }
        if sustained:
            sus_fps = world * BATCH * sustained["steps"] / (sus_ms_max / 1e3)
            sus_tflops = sus_fps * GFLOP_PER_FRAME / 1e3 / world
            line["sustained"] = {"value": sus_fps, "unit": "frames/s", "seconds": sus_ms_max / 1e3, "steps": sustained["steps"],
                                 "ms_per_step": sus_ms_max / sustained["steps"], "achieved_tflops": sus_tflops,
                                 "peak_sustained": pk["bf16_tflops_sustained"],
                                 "frac_of_sustained": sus_tflops / pk["bf16_tflops_sustained"], "clocks": sustained["clocks"]}
        if fp32_io:
            line["fp32_io"] = {"value": world * BATCH * args.steps / (ms32_max / 1e3), "unit": "frames/s",
                               "e2e": world * BATCH * fp32_io["e2e_steps"] / (e2e32_ms_max / 1e3),
                               "h2d_bytes_per_step": fp32_io["h2d"], "d2h_bytes_per_step": fp32_io["d2h"],
                               "api": "rst_transfer_forward / rst_transfer_submit_host (float32 in, float32 out: the drop-in default)",
                               "uint8_path_vs_trunc255_of_float_path_max_levels": fp32_io["uint8_vs_float_levels_max"]}
        if batch_sweep:
            line["batch_sweep"] = batch_sweep
        if config_recs:
            line["configs"] = config_recs
        if training:
            line["training"] = training
        if not args.no_cpu_baseline and world == 1:       # the CPU arm is timed at N=1 only (its host cores are shared at N>1)
            line["cpu_baseline"], parity = cpu_baseline_sample(cfg, weights, content16_np, params_np, frame0_u8, extra_cases)
            line["parity"] = parity
            for name, rec in config_recs.items():
                if name in parity:
                    rec["parity"] = parity[name]
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
