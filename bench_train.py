"""Training-step benchmark (BASELINE.json configs[3]: predictor + transfer fwd/bwd in training mode, VGG loss x3 fwd + x1
dgrad, RMSprop, gradient all-reduce; B samples per GPU at 480x960, F=128, MobileNetV3 predictor).  Secondary to bench.py
(the headline metric is inference frames/s); prints ONE JSON line with samples/s.

  python bench_train.py --batch 8 --steps 5 --warmup 2
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench_train.py --gpus N
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from realtime_style_transfer_b200 import _native, _plan, distributed as rdist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=8, help="samples per GPU")
    ap.add_argument("--extractor", default="MOBILE_NET", choices=["MOBILE_NET", "DUMMY"])
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=960)
    ap.add_argument("--filters", type=int, default=128)
    ap.add_argument("--math", default="default", choices=["default", "fp32", "tf32"],
                    help="default = fp32: split-tf32 tensor-core convs (fp32-level accuracy); tf32: plain tf32 operands")
    args = ap.parse_args()
    rank, world, local = rdist.env_rank()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        rdist.init_process_group("nccl", device=dev)
    h, w, f, b = args.height, args.width, args.filters, args.batch
    in_shape, out_shape = (h, w, 17), (h, w, 3)
    tplan = _plan.TransferPlan(in_shape, out_shape, h // 4, f, 1)
    pplan = _plan.PredictorPlan(out_shape, args.extractor, tplan.num_style_parameters, 100)
    rng = np.random.default_rng(1234)
    weights = dict(tplan.initial_weights(rng))
    weights.update(pplan.initial_weights(rng))
    tr = _native.NativeTrainer(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=h // 4, bottleneck_num_filters=f,
                               max_batch=b, extractor=getattr(_native, "EXTRACTOR_" + args.extractor), style_shape=out_shape[:2],
                               device=local)
    if args.math == "tf32":
        tr.set_math(_native.PRECISION_TF32)
    tr.model.set_weights(weights)
    vgg = {}
    cin = 3
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

    def step():
        tr.forward_backward(content.data_ptr(), style.data_ptr(), gt.data_ptr(), style.data_ptr(), losses.data_ptr(), b)
        rdist.allreduce_sum_(grads)
        torch.cuda.synchronize(dev)
        tr.apply_gradients()

    def sync_all():
        torch.cuda.synchronize(dev)
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sync_all()
    dt = time.perf_counter() - t0
    dt = rdist.max_over_ranks([dt], device=dev)[0]
    if rank == 0:
        ms = dt / args.steps * 1e3
        # ~1.56 TFLOP per sample at 480x960 / F=128 (72 % of it the frozen VGG)
        flop_per_sample = 1.56e12 * (h * w) / (480 * 960)
        print(json.dumps({
            "metric": "training_samples_per_second", "value": b * world / (dt / args.steps), "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "dtype": {"default": "f32 (split tf32 on tensor cores)", "fp32": "f32 (split tf32 on tensor cores)", "tf32": "tf32"}[args.math], "data": "synthetic",
            "config": {"workload": f"train-{w}-{h // 4}-{f}-17 B={b}/GPU, {args.extractor} predictor, VGG16 loss, RMSprop"},
            "approx_tflops_per_gpu": flop_per_sample * b / (dt / args.steps) / 1e12,
            "loss": float(losses[:, 0].mean().item()),
            "gpu_launches": int(tr.lib.rst_last_launch_count(tr.model.handle))}))
    tr.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
