"""Per-group device time of the fp32 inference path (rst_profile_*): `python tools/experiments/fp32_path_profile.py [spec] [batch]`.
Used for the config-1 numbers of profiles/r02_00_summary.md."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from realtime_style_transfer_b200 import _native
from realtime_style_transfer_b200._plan import TransferPlan
from realtime_style_transfer_b200.shape_config import ShapeConfig
spec = sys.argv[1] if len(sys.argv) > 1 else "rst-960-120-32-3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = ShapeConfig.from_spec(spec)
in_shape, out_shape = cfg.input_shape["content"], cfg.output_shape
plan = TransferPlan(in_shape, out_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters, 1)
w = plan.initial_weights(np.random.default_rng(1))
ctx = _native.NativeContext(in_shape=in_shape, out_shape=out_shape, bottleneck_res_y=cfg.bottleneck_res_y,
                            bottleneck_num_filters=cfg.bottleneck_num_filters, num_styles=1, max_batch=B)
ctx.set_weights(w)
x = np.random.default_rng(0).uniform(0, 1, (B,) + in_shape).astype(np.float32)
p = np.random.default_rng(1).uniform(0.3, 1.0, (B, 1, plan.num_style_parameters)).astype(np.float32)
for _ in range(3): ctx.transfer_forward_host(x, p)
ctx.profile(True); ctx.profile_reset()
for _ in range(5): ctx.transfer_forward_host(x, p)
g = ctx.profile_groups()
tot = sum(v[0] for v in g.values())
for k, v in sorted(g.items(), key=lambda kv: -kv[1][0]): print(f"{k:20s} {v[0]/5:8.3f} ms/fwd  {v[1]//5:4d} launches  {100*v[0]/tot:5.1f} %")
print("total", tot / 5)
