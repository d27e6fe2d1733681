"""Experiment: one batch of 8 frames against two half batches on two streams (do the norm passes of one half hide under the
convolutions of the other?)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import bench
from realtime_style_transfer_b200 import _native
from realtime_style_transfer_b200._plan import TransferPlan
from realtime_style_transfer_b200.shape_config import ShapeConfig

cfg = ShapeConfig.from_spec(bench.SPEC)
in_shape, out_shape = cfg.input_shape["content"], cfg.output_shape
plan = TransferPlan(in_shape, out_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
weights = bench.randomise_bn(plan.initial_weights(np.random.default_rng(1)))
dev = torch.device("cuda", 0)
split = [int(a) for a in (sys.argv[1] if len(sys.argv) > 1 else "4,4").split(",")]

def make(b):
    c = _native.NativeContext(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=cfg.bottleneck_res_y,
                              bottleneck_num_filters=cfg.bottleneck_num_filters, num_styles=1, max_batch=b,
                              precision=_native.PRECISION_BF16, device=0)
    c.set_weights(weights)
    return c

content = torch.from_numpy(bench.synthetic_inputs(cfg, 8, seed=0).astype(np.float16)).to(dev)
params = torch.from_numpy(np.random.default_rng(5).normal(0, 1, (8, 1, plan.num_style_parameters)).astype(np.float32)).to(dev)
out = torch.empty((8,) + out_shape, dtype=torch.uint8, device=dev)
out2 = torch.empty_like(out)

def fwd(c, lo, n, stream, o):
    c.transfer_forward_device(content[lo:lo + n].data_ptr(), params[lo:lo + n].data_ptr(), None, o[lo:lo + n].data_ptr(), n,
                              stream.cuda_stream, content_dtype=_native.DTYPE_F16, out_dtype=_native.DTYPE_U8)

s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
c8 = make(8)
def time_it(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s0)
    for _ in range(iters): fn()
    j = torch.cuda.Event(); j.record(s1); s0.wait_event(j)
    e1.record(s0)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

def whole(): fwd(c8, 0, 8, s0, out)
t8 = time_it(whole)
print(f"batch 8, one stream: {t8:.3f} ms  {8 / t8 * 1e3:.0f} frames/s")
ca, cb = make(split[0]), make(split[1])
def serial():
    fwd(ca, 0, split[0], s0, out2); fwd(cb, split[0], split[1], s0, out2)
ts = time_it(serial)
print(f"{split} on one stream: {ts:.3f} ms  {8 / ts * 1e3:.0f} frames/s")
def dual():
    fwd(ca, 0, split[0], s0, out2); fwd(cb, split[0], split[1], s1, out2)
# start s1 after s0's start event
td = time_it(dual)
print(f"{split} on two streams: {td:.3f} ms  {8 / td * 1e3:.0f} frames/s   same bytes: {bool((out == out2).all())}")
