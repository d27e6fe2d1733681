"""CPU oracle for the stylization hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product path (``realtime_style_transfer_b200``) never does.

What it is: a PyTorch-CPU fp32 (optionally fp64) restatement of the arithmetic that the
reference expresses as TensorFlow 2.9 / Keras graphs.  TensorFlow is not installable in this
environment so the reference itself cannot be run; the TF/Keras op semantics
are restated from their published definitions and each function cites the reference call
site it follows (paths relative to /root/reference).

PARITY PINNING STATUS
  * ``apply_style_weights``  -- PINNED by the reference's own known-answer test
    (realtime_style_transfer/models/styleTransferTest.py:28-49), see tests/golden/.
  * everything else (conv SAME padding, Conv2DTranspose, BatchNorm, CIN, MobileNetV3Small,
    VGG16, Gram, losses) -- **parity unpinned**: the reference ships no golden vectors,
    fixtures or checkpoints for them.  They are cross-checked against an
    independent pure-numpy loop restatement (oracle/naive_np.py) on small cases.

All tensors at this interface are NHWC numpy/torch arrays like the reference's.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

BN_EPS = 1e-3          # Keras BatchNormalization default epsilon (styleTransfer.py:201)
CIN_EPS = 1e-5         # ConditionalInstanceNormalization epsilon (styleTransfer.py:51)


# --------------------------------------------------------------------------------------
# TF / Keras primitive semantics
# --------------------------------------------------------------------------------------
def tf_same_pad(size: int, k: int, s: int) -> Tuple[int, int]:
    """TF 'SAME' padding (before, after) for one spatial dim.  out = ceil(size/s)."""
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _nchw(x: Tensor) -> Tensor:
    return x.permute(0, 3, 1, 2)


def _nhwc(x: Tensor) -> Tensor:
    return x.permute(0, 2, 3, 1)


def conv2d_same(x: Tensor, kernel: Tensor, bias: Optional[Tensor], stride: int = 1,
                groups: int = 1) -> Tensor:
    """tf.keras.layers.Conv2D(padding='same').  x NHWC, kernel HWIO (kh,kw,in/groups,out)."""
    kh, kw = int(kernel.shape[0]), int(kernel.shape[1])
    pt, pb = tf_same_pad(x.shape[1], kh, stride)
    pl, pr = tf_same_pad(x.shape[2], kw, stride)
    xin = F.pad(_nchw(x), (pl, pr, pt, pb))
    w = kernel.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xin, w, bias, stride=stride, groups=groups)
    return _nhwc(y)


def depthwise_conv2d(x: Tensor, kernel: Tensor, stride: int, pads: Tuple[int, int, int, int]) -> Tensor:
    """Keras DepthwiseConv2D, kernel (kh,kw,C,1); explicit pads (top,bottom,left,right)."""
    c = x.shape[3]
    xin = F.pad(_nchw(x), (pads[2], pads[3], pads[0], pads[1]))
    w = kernel.permute(2, 3, 0, 1).contiguous()  # (C,1,kh,kw)
    return _nhwc(F.conv2d(xin, w, None, stride=stride, groups=c))


def conv2d_transpose_same(x: Tensor, kernel: Tensor, bias: Optional[Tensor], stride: int) -> Tensor:
    """tf.keras.layers.Conv2DTranspose(padding='same') (styleTransfer.py:115-119).

    kernel is Keras layout (kh, kw, out, in).  TF defines the op as the input-gradient of the
    forward SAME conv: out[y] = sum_k in[(y + pad_before - k)/s] * K[k] over exact divisions,
    output size s*H, pad_before taken from the forward conv on the (s*H)-sized tensor.
    """
    kh, kw = int(kernel.shape[0]), int(kernel.shape[1])
    ho, wo = x.shape[1] * stride, x.shape[2] * stride
    pt, _ = tf_same_pad(ho, kh, stride)
    pl, _ = tf_same_pad(wo, kw, stride)
    w = kernel.permute(3, 2, 0, 1).contiguous()  # (in, out, kh, kw) -- torch conv_transpose2d layout
    full = F.conv_transpose2d(_nchw(x), w, bias, stride=stride, padding=0)
    return _nhwc(full[:, :, pt:pt + ho, pl:pl + wo])


def batchnorm(x: Tensor, gamma, beta, mean, var, eps: float = BN_EPS, training: bool = False, moving_out=None,
              momentum: float = 0.99):
    """Keras BatchNormalization(axis=-1).  training=True uses batch statistics over (N,H,W).
    moving_out (a list) receives the updated (moving_mean, moving_variance): Keras' fused path (4-D NHWC input) feeds the
    Bessel-corrected batch variance into the moving average (normalization/batch_normalization.py,
    `_bessels_correction_test_only` left True) -- parity unpinned, TF is un-vendored."""
    if training:
        bm = x.mean(dim=(0, 1, 2))
        bv = x.var(dim=(0, 1, 2), unbiased=False)
        if moving_out is not None:
            n = x.shape[0] * x.shape[1] * x.shape[2]
            ub = bv * (n / max(n - 1, 1))
            moving_out.append(((momentum * mean + (1 - momentum) * bm).detach(), (momentum * var + (1 - momentum) * ub).detach()))
        mean, var = bm, bv
    return (x - mean) * torch.rsqrt(var + eps) * gamma + beta


def hard_sigmoid(x: Tensor) -> Tensor:
    return F.relu6(x + 3.0) * (1.0 / 6.0)


def hard_swish(x: Tensor) -> Tensor:
    return x * hard_sigmoid(x)


# --------------------------------------------------------------------------------------
# Transfer network (realtime_style_transfer/models/styleTransfer.py)
# --------------------------------------------------------------------------------------
def apply_style_weights(style_weights: Optional[Tensor], style_params: Tensor) -> Tensor:
    """styleTransfer.py:36-44.  params (B,1,S,F); weights (B,H,W,S).  Exactly 2 styles blend:
    a plain weighted sum, no clamping or normalisation; otherwise pass-through."""
    if style_params.shape[-2] == 2:
        return (style_params.unsqueeze(1) * style_weights.unsqueeze(-1)).sum(dim=-2)
    return style_params


def cin(x: Tensor, scale: Tensor, bias: Tensor, eps: float = CIN_EPS) -> Tensor:
    """ConditionalInstanceNormalization.call, styleTransfer.py:57-71 (moments over H,W,
    population variance, x*inv - mean*inv, then bias + x*scale)."""
    mean = x.mean(dim=(1, 2), keepdim=True)
    var = x.var(dim=(1, 2), unbiased=False, keepdim=True)
    inv = torch.rsqrt(var + eps)
    xh = x * inv + (-mean * inv)
    return bias + xh * scale


def block_counts(in_h: int, out_h: int, bottleneck_res_y: int) -> Tuple[int, int]:
    """styleTransfer.py:217 and :258 -- must keep the reference's exact float expression."""
    n_contract = math.ceil(math.log2(in_h) - math.log2(bottleneck_res_y))
    bott_h = int(in_h * 2 ** -n_contract)
    n_expand = math.ceil(math.log2(out_h) - math.log2(bott_h))
    return n_contract, n_expand


CONTRACT_FILTERS = [16, 32, 32, 32]               # styleTransfer.py:218-223
EXPAND_FILTERS = [32, 16, 8, 4, 3, 3, 3, 3]       # styleTransfer.py:247-256


class TransferSpec:
    """Layer plan of create_style_transfer_model (styleTransfer.py:213-332)."""

    def __init__(self, input_shape, output_shape, bottleneck_res_y, bottleneck_num_filters, num_styles):
        self.input_shape = tuple(input_shape)
        self.output_shape = tuple(output_shape)
        self.filters = int(bottleneck_num_filters)
        self.num_styles = int(num_styles)
        self.n_contract, self.n_expand = block_counts(input_shape[0], output_shape[0], bottleneck_res_y)
        self.contract = [("start", input_shape[2], 32, 9, 1)]
        cin_ = 32
        for i in range(self.n_contract):
            self.contract.append((str(i), cin_, CONTRACT_FILTERS[i], 3, 2))
            cin_ = CONTRACT_FILTERS[i]
        self.res_in = cin_
        self.expand = []
        cin_ = self.filters
        for i in range(self.n_expand):
            self.expand.append((str(i), cin_, EXPAND_FILTERS[i], 3, 2))
            cin_ = EXPAND_FILTERS[i]
        self.expand.append(("last", cin_, 3, 9, 1))
        self.num_style_parameters = 5 * 4 * self.filters + sum(2 * e[2] for e in self.expand)

    def weight_shapes(self) -> Dict[str, Tuple[int, ...]]:
        shapes = {}
        for name, ci, co, k, _ in self.contract:
            p = f"contract_{name}"
            shapes[f"{p}/conv/kernel"] = (k, k, ci, co)
            shapes[f"{p}/conv/bias"] = (co,)
            for v in ("gamma", "beta", "moving_mean", "moving_variance"):
                shapes[f"{p}/bn/{v}"] = (co,)
        for b in range(5):
            ci = self.res_in if b == 0 else self.filters
            shapes[f"residual_block_{b}/conv0/kernel"] = (3, 3, ci, self.filters)
            shapes[f"residual_block_{b}/conv0/bias"] = (self.filters,)
            shapes[f"residual_block_{b}/conv1/kernel"] = (3, 3, self.filters, self.filters)
            shapes[f"residual_block_{b}/conv1/bias"] = (self.filters,)
        for name, ci, co, k, _ in self.expand:
            shapes[f"expand_{name}/conv/kernel"] = (k, k, co, ci)   # Conv2DTranspose: (kh,kw,out,in)
            shapes[f"expand_{name}/conv/bias"] = (co,)
        return shapes


def init_transfer_weights(spec: TransferSpec, seed: int = 1, trained_like: bool = False) -> Dict[str, np.ndarray]:
    """Seeded synthetic weights from the reference initialisers: contract/expand kernels
    N(0,0.02) (styleTransfer.py:97,190), residual kernels U(0,0.05) (:146), biases 0, BN gamma 1 /
    beta 0; moving statistics are randomised so inference-mode BN is exercised.
    trained_like=True gives a zero-mean He-scaled residual trunk and non-zero biases."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in spec.weight_shapes().items():
        if name.endswith("/kernel"):
            if name.startswith("residual_block"):
                if trained_like:
                    fan_in = shape[0] * shape[1] * shape[2]
                    w = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
                else:
                    w = torch.rand(shape, generator=g) * 0.05
            else:
                w = torch.randn(shape, generator=g) * (0.05 if trained_like else 0.02)
        elif name.endswith("/bias"):
            w = torch.randn(shape, generator=g) * 0.1 if trained_like else torch.zeros(shape)
        elif name.endswith("gamma"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g) if trained_like else torch.ones(shape)
        elif name.endswith("beta"):
            w = 0.1 * torch.randn(shape, generator=g) if trained_like else torch.zeros(shape)
        elif name.endswith("moving_mean"):
            w = 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("moving_variance"):
            w = 0.5 + torch.rand(shape, generator=g)
        else:
            raise KeyError(name)
        out[name] = w.numpy().astype(np.float32)
    return out


def style_weight_mips(style_weights: Tensor, num_mips: int) -> Dict[int, Tensor]:
    """styleTransfer.py:335-345: pyramid keyed by WIDTH, AvgPool2D(2) each level."""
    mips = {int(style_weights.shape[2]): style_weights}
    last = style_weights
    for _ in range(num_mips):
        last = _nhwc(F.avg_pool2d(_nchw(last), 2))
        mips[int(last.shape[2])] = last
    return mips


def transfer_forward(spec: TransferSpec, weights: Dict[str, np.ndarray], content, style_params,
                     style_weights=None, dtype=torch.float32, training: bool = False,
                     taps: Optional[Dict[str, Tensor]] = None, emulate_bf16: bool = False) -> Tensor:
    """create_style_transfer_model forward, styleTransfer.py:213-332.

    content (B,H,W,C); style_params (B,S,P); style_weights (B,Ho,Wo,S-1) when S>1.
    ``taps`` (optional dict) receives every intermediate activation for layer-level parity.

    emulate_bf16: an IDEAL bf16 execution of the same graph -- the content, every convolution kernel and every activation that
    a mixed-bfloat16 implementation stores between layers are rounded to bf16 (round to nearest even), everything else
    (accumulation, biases, BatchNorm affine, instance-norm statistics and affine, sigmoid) is exact in ``dtype`` (use
    float64).  The rounding points are those of the CUDA path: conv output after bias/ReLU(/BN/ReLU),
    the normalised tensor after its activation / skip add, the transposed-conv outputs; the 3-channel head stays fp32.
    It bounds what ANY bf16 implementation can achieve against the fp32 reference, which is how the tests separate inherent
    bf16 error (instance norm amplifies it on near-constant channels) from kernel defects."""
    rb = (lambda t: t.to(torch.bfloat16).to(dtype)) if emulate_bf16 else (lambda t: t)
    W = {k: (rb(torch.as_tensor(v).to(dtype)) if k.endswith("/kernel") else torch.as_tensor(v).to(dtype)) for k, v in weights.items()}
    x = rb(torch.as_tensor(content).to(dtype))
    sp = torch.as_tensor(style_params).to(dtype).unsqueeze(1)          # (B,1,S,P)  :305
    mips = None
    if spec.num_styles > 1:
        sw = torch.as_tensor(style_weights).to(dtype)
        assert sw.shape[1] == spec.output_shape[0] and sw.shape[2] == spec.output_shape[1]
        sw = torch.cat([1 - sw.sum(dim=-1, keepdim=True), sw], dim=-1)  # :297-302
        mips = style_weight_mips(sw, spec.n_expand + 1)
    cursor = 0

    def take(n):
        nonlocal cursor
        p = sp[..., cursor:cursor + n]
        cursor += n
        return p

    def tap(name, t):
        if taps is not None:
            taps[name] = t

    for name, _, co, k, s in spec.contract:                              # contract, :188-205
        p = f"contract_{name}"
        x = F.relu(conv2d_same(x, W[f"{p}/conv/kernel"], W[f"{p}/conv/bias"], s))
        x = F.relu(batchnorm(x, W[f"{p}/bn/gamma"], W[f"{p}/bn/beta"], W[f"{p}/bn/moving_mean"],
                             W[f"{p}/bn/moving_variance"], training=training))
        x = rb(x)
        tap(p, x)
    f = spec.filters
    for b in range(5):                                                   # residual_block, :144-185
        params = take(4 * f)
        w_mip = mips[int(x.shape[2])] if mips is not None else None      # :317
        fx = x
        for i in range(2):
            p = f"residual_block_{b}/conv{i}"
            fx = rb(F.relu(conv2d_same(fx, W[f"{p}/kernel"], W[f"{p}/bias"], 1)))
            tap(p + "/relu", fx)
            sl = params[..., 2 * f * i: 2 * f * (i + 1)]
            scale = apply_style_weights(w_mip, sl[..., :f])
            bias = apply_style_weights(w_mip, sl[..., f:])
            fx = cin(fx, scale, bias)
            if i == 0:
                fx = rb(F.relu(fx))
            tap(p + "/cin", fx)
        x = rb(fx if b == 0 else x + fx)
        tap(f"residual_block_{b}", x)
    for name, _, co, k, s in spec.expand:                                # expand, :95-141
        p = f"expand_{name}"
        params = take(2 * co)
        w_mip = mips[int(x.shape[2]) * s] if mips is not None else None  # :325-326
        x = conv2d_transpose_same(x, W[f"{p}/conv/kernel"], W[f"{p}/conv/bias"], s)
        if name != "last":
            x = rb(x)
        tap(p + "/conv", x)
        x = cin(x, apply_style_weights(w_mip, params[..., :co]), apply_style_weights(w_mip, params[..., co:]))
        x = torch.sigmoid(x) if name == "last" else rb(F.relu(x))
        tap(p, x)
    assert cursor == spec.num_style_parameters
    return x.to(torch.float32) if dtype == torch.float32 else x


# --------------------------------------------------------------------------------------
# Style predictor (realtime_style_transfer/models/stylePrediction.py) -- Keras MobileNetV3Small
# restated from keras/applications/mobilenet_v3.py (Keras 2.9, un-vendored dependency).
# --------------------------------------------------------------------------------------
def _depth(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# (expansion, out, kernel, stride, use_se, activation) -- MobileNetV3Small stack_fn, alpha=1
MBV3_SMALL_BLOCKS = [
    (1, 16, 3, 2, True, "relu"),
    (72. / 16, 24, 3, 2, False, "relu"),
    (88. / 24, 24, 3, 1, False, "relu"),
    (4, 40, 5, 2, True, "hswish"),
    (6, 40, 5, 1, True, "hswish"),
    (6, 40, 5, 1, True, "hswish"),
    (3, 48, 5, 1, True, "hswish"),
    (3, 48, 5, 1, True, "hswish"),
    (6, 96, 5, 2, True, "hswish"),
    (6, 96, 5, 1, True, "hswish"),
    (6, 96, 5, 1, True, "hswish"),
]


def mobilenet_plan():
    """Expanded per-block channel plan: list of dicts with prefix, cin, cexp, cout, k, s, se, act."""
    plan = []
    cin_ = 16
    for bid, (e, co, k, s, se, act) in enumerate(MBV3_SMALL_BLOCKS):
        cexp = _depth(cin_ * e)
        plan.append(dict(prefix="expanded_conv" if bid == 0 else f"expanded_conv_{bid}", block_id=bid,
                         cin=cin_, cexp=cexp, cout=co, k=k, s=s,
                         se=_depth(cexp * 0.25) if se else 0, act=act))
        cin_ = co
    return plan, _depth(cin_ * 6)


def predictor_weight_shapes(extractor: str, num_top_parameters: int, num_style_parameters: int = 100):
    shapes = {}

    def bn(prefix, c):
        for v in ("gamma", "beta", "moving_mean", "moving_variance"):
            shapes[f"{prefix}/{v}"] = (c,)

    if extractor == "DUMMY":
        shapes["dummy_conv/kernel"] = (9, 9, 3, 1)
        shapes["dummy_conv/bias"] = (1,)
        feat = 1
    elif extractor == "MOBILE_NET":
        shapes["mobilenet/Conv/kernel"] = (3, 3, 3, 16)
        bn("mobilenet/Conv/BatchNorm", 16)
        plan, last = mobilenet_plan()
        for b in plan:
            p = "mobilenet/" + b["prefix"]
            if b["block_id"]:
                shapes[f"{p}/expand/kernel"] = (1, 1, b["cin"], b["cexp"])
                bn(f"{p}/expand/BatchNorm", b["cexp"])
            shapes[f"{p}/depthwise/depthwise_kernel"] = (b["k"], b["k"], b["cexp"], 1)
            bn(f"{p}/depthwise/BatchNorm", b["cexp"])
            if b["se"]:
                shapes[f"{p}/squeeze_excite/Conv/kernel"] = (1, 1, b["cexp"], b["se"])
                shapes[f"{p}/squeeze_excite/Conv/bias"] = (b["se"],)
                shapes[f"{p}/squeeze_excite/Conv_1/kernel"] = (1, 1, b["se"], b["cexp"])
                shapes[f"{p}/squeeze_excite/Conv_1/bias"] = (b["cexp"],)
            shapes[f"{p}/project/kernel"] = (1, 1, b["cexp"], b["cout"])
            bn(f"{p}/project/BatchNorm", b["cout"])
        shapes["mobilenet/Conv_1/kernel"] = (1, 1, plan[-1]["cout"], last)
        bn("mobilenet/Conv_1/BatchNorm", last)
        feat = last
    else:
        raise ValueError(f"{extractor} is not a valid value for feature_extractor")
    shapes["StylePredictor/kernel"] = (1, 1, feat, num_style_parameters)
    shapes["StylePredictor/bias"] = (num_style_parameters,)
    shapes["StyleNormPredictor/kernel"] = (1, 1, num_style_parameters, num_top_parameters)
    shapes["StyleNormPredictor/bias"] = (num_top_parameters,)
    return shapes


def init_predictor_weights(extractor: str, num_top_parameters: int, seed: int = 2,
                           num_style_parameters: int = 100) -> Dict[str, np.ndarray]:
    """Heads: VarianceScaling(1/3, fan_out, uniform), bias 0.5 (stylePrediction.py:9-16,:62,:69).
    MobileNet body: imagenet weights are unreachable offline -> random He init, BN stats randomised."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in predictor_weight_shapes(extractor, num_top_parameters, num_style_parameters).items():
        if name.startswith("Style") and name.endswith("kernel"):
            fan_out = shape[0] * shape[1] * shape[3]
            lim = math.sqrt(3.0 * (1.0 / 3.0) / fan_out)
            w = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif name.startswith("Style") and name.endswith("bias"):
            w = torch.full(shape, 0.5)
        elif name.endswith("kernel"):
            fan_in = shape[0] * shape[1] * (shape[2] if not name.endswith("depthwise_kernel") else 1)
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_in)
        elif name.endswith("bias"):
            w = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("gamma"):
            w = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("beta"):
            w = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("moving_mean"):
            w = 0.1 * torch.randn(shape, generator=g)
        elif name.endswith("moving_variance"):
            w = 0.5 + torch.rand(shape, generator=g)
        else:
            raise KeyError(name)
        out[name] = w.numpy().astype(np.float32)
    return out


def _correct_pad(size: int, k: int) -> Tuple[int, int]:
    """keras imagenet_utils.correct_pad for one dim."""
    adjust = 1 - size % 2
    return k // 2 - adjust, k // 2


def mobilenet_v3_small(W: Dict[str, Tensor], x: Tensor, training: bool = False) -> Tensor:
    def bn(prefix, t):
        return batchnorm(t, W[f"{prefix}/gamma"], W[f"{prefix}/beta"], W[f"{prefix}/moving_mean"],
                         W[f"{prefix}/moving_variance"], eps=1e-3, training=training)

    def act(name, t):
        return F.relu(t) if name == "relu" else hard_swish(t)

    x = act("hswish", bn("mobilenet/Conv/BatchNorm", conv2d_same(x, W["mobilenet/Conv/kernel"], None, 2)))
    plan, _ = mobilenet_plan()
    for b in plan:
        p = "mobilenet/" + b["prefix"]
        shortcut = x
        if b["block_id"]:
            x = act(b["act"], bn(f"{p}/expand/BatchNorm", conv2d_same(x, W[f"{p}/expand/kernel"], None, 1)))
        k, s = b["k"], b["s"]
        if s == 2:
            pt, pb = _correct_pad(x.shape[1], k)
            pl, pr = _correct_pad(x.shape[2], k)
        else:
            pt, pb = tf_same_pad(x.shape[1], k, 1)
            pl, pr = tf_same_pad(x.shape[2], k, 1)
        x = depthwise_conv2d(x, W[f"{p}/depthwise/depthwise_kernel"], s, (pt, pb, pl, pr))
        x = act(b["act"], bn(f"{p}/depthwise/BatchNorm", x))
        if b["se"]:
            z = x.mean(dim=(1, 2), keepdim=True)
            z = F.relu(conv2d_same(z, W[f"{p}/squeeze_excite/Conv/kernel"], W[f"{p}/squeeze_excite/Conv/bias"]))
            z = hard_sigmoid(conv2d_same(z, W[f"{p}/squeeze_excite/Conv_1/kernel"],
                                         W[f"{p}/squeeze_excite/Conv_1/bias"]))
            x = x * z
        x = bn(f"{p}/project/BatchNorm", conv2d_same(x, W[f"{p}/project/kernel"], None, 1))
        if s == 1 and b["cin"] == b["cout"]:
            x = shortcut + x
    x = act("hswish", bn("mobilenet/Conv_1/BatchNorm", conv2d_same(x, W["mobilenet/Conv_1/kernel"], None, 1)))
    return x


def predictor_forward(extractor: str, weights: Dict[str, np.ndarray], style_image, dtype=torch.float32,
                      training: bool = False) -> Tensor:
    """create_style_prediction_model forward, stylePrediction.py:25-75.  (B,H,W,3) -> (B,P)."""
    W = {k: torch.as_tensor(v).to(dtype) for k, v in weights.items()}
    x = torch.as_tensor(style_image).to(dtype)
    if extractor == "DUMMY":
        x = conv2d_same(x, W["dummy_conv/kernel"], W["dummy_conv/bias"], 5)       # :30-31
    elif extractor == "MOBILE_NET":
        x = mobilenet_v3_small(W, x * 2.0 - 1.0, training)                        # :33-37
    else:
        raise ValueError(f"{extractor} is not a valid value for feature_extractor")
    x = x.mean(dim=(1, 2), keepdim=True)                                          # :54
    x = conv2d_same(x, W["StylePredictor/kernel"], W["StylePredictor/bias"])      # :59-63
    x = conv2d_same(x, W["StyleNormPredictor/kernel"], W["StyleNormPredictor/bias"])  # :66-70
    return x[:, 0, 0, :]


def inference_forward(spec: TransferSpec, transfer_w, extractor: str, predictor_w, content, style,
                      style_weights=None, dtype=torch.float32) -> Tensor:
    """make_style_transfer_inference_model, styleTransferInferenceModel.py:9-48."""
    style = torch.as_tensor(style).to(dtype)
    params = torch.stack([predictor_forward(extractor, predictor_w, style[:, s], dtype)
                          for s in range(style.shape[1])], dim=1)
    return transfer_forward(spec, transfer_w, content, params, style_weights, dtype)


# --------------------------------------------------------------------------------------
# Loss (realtime_style_transfer/models/styleLoss.py)
# --------------------------------------------------------------------------------------
VGG16_CFG = [("block1", 2, 64), ("block2", 2, 128), ("block3", 3, 256), ("block4", 3, 512), ("block5", 3, 512)]
VGG_STYLE_LAYERS = ["block1_conv2", "block2_conv2", "block3_conv3", "block4_conv3"]   # styleLoss.py:80
VGG_CONTENT_LAYERS = ["block5_conv3"]                                               # styleLoss.py:81
VGG_CAFFE_MEAN_BGR = (103.939, 116.779, 123.68)


def vgg16_weight_shapes():
    shapes, cin_ = {}, 3
    for blk, n, co in VGG16_CFG:
        for i in range(1, n + 1):
            shapes[f"{blk}_conv{i}/kernel"] = (3, 3, cin_, co)
            shapes[f"{blk}_conv{i}/bias"] = (co,)
            cin_ = co
    return shapes


def init_vgg16_weights(seed: int = 3) -> Dict[str, np.ndarray]:
    """imagenet weights are unreachable offline: random He-init."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in vgg16_weight_shapes().items():
        if name.endswith("kernel"):
            w = torch.randn(shape, generator=g) * math.sqrt(2.0 / (shape[0] * shape[1] * shape[2]))
        else:
            w = 0.05 * torch.randn(shape, generator=g)
        out[name] = w.numpy().astype(np.float32)
    return out


class _RoundTF32(torch.autograd.Function):
    """Round to the tf32 grid (8-bit exponent, 10 explicit mantissa bits), straight-through gradient."""

    @staticmethod
    def forward(ctx, x):
        m, e = torch.frexp(x)                        # x = m * 2**e, 0.5 <= |m| < 1: 11 significant bits -> scale by 2**11
        return torch.ldexp(torch.round(m * 2048.0) / 2048.0, e)

    @staticmethod
    def backward(ctx, g):
        return g


def tf32_round(x: Tensor) -> Tensor:
    return _RoundTF32.apply(x)


def vgg16_features(W: Dict[str, Tensor], image01: Tensor, tf32: bool = False) -> Dict[str, Tensor]:
    """StyleLossModelVGG.call, styleLoss.py:69-109: x*255 -> caffe preprocess -> VGG16 taps.
    tf32=True restates TensorFloat-32 execution -- TensorFlow's default for float32 convolutions on Ampere-and-later GPUs
    (tf.config.experimental.enable_tensor_float_32_execution): the operands of every convolution with >= 64 input channels
    are rounded to tf32, products and sums stay exact (here: the dtype of the call).  The 3-channel first layer stays fp32."""
    x = image01 * 255.0
    x = x.flip(-1) - torch.tensor(VGG_CAFFE_MEAN_BGR, dtype=x.dtype)
    feats = {}
    first = True
    for bi, (blk, n, _) in enumerate(VGG16_CFG):
        for i in range(1, n + 1):
            k = W[f"{blk}_conv{i}/kernel"]
            if tf32 and not first:
                k = tf32_round(k)
            x = F.relu(conv2d_same(x, k, W[f"{blk}_conv{i}/bias"]))
            if tf32:
                x = tf32_round(x)                    # activations are stored on the tf32 grid: the next conv's operand
            feats[f"{blk}_conv{i}"] = x
            first = False
        if bi < 4:
            x = _nhwc(F.max_pool2d(_nchw(x), 2))
    return feats


def gram_matrix(x: Tensor) -> Tensor:
    """get_gram_matrix_model, styleLoss.py:11-18: einsum('bijc,bijd->bcd') / (H*W)."""
    return torch.einsum("bijc,bijd->bcd", x, x) / float(x.shape[1] * x.shape[2])


def mean_l2_loss_on_batch(t: Tensor) -> Tensor:
    """styleLoss.py:290-292: mean over all but the batch axis of 0.5*t^2 -> (B,)."""
    return (0.5 * t * t).flatten(1).mean(dim=1)


def total_variation(img: Tensor) -> Tensor:
    """tf.image.total_variation: sum|dy| + sum|dx| over (H,W,C) -> (B,)."""
    dy = (img[:, 1:] - img[:, :-1]).abs().flatten(1).sum(dim=1)
    dx = (img[:, :, 1:] - img[:, :, :-1]).abs().flatten(1).sum(dim=1)
    return dy + dx


def style_loss_vgg(vgg_weights, prediction, gt_content, gt_style, dtype=torch.float32,
                   content_factor=1e4, style_factor=1e-3, tv_factor=1e-1, tf32: bool = False) -> Dict[str, Tensor]:
    """make_style_loss_function with StyleLossModelVGG and with_depth_loss=False,
    styleLoss.py:295-369 (factors :101-104).  All outputs are (B,) vectors."""
    W = {k: torch.as_tensor(v).to(dtype) for k, v in vgg_weights.items()}
    pred = torch.as_tensor(prediction).to(dtype) if not isinstance(prediction, Tensor) else prediction
    gt_style = torch.as_tensor(gt_style).to(dtype)
    if gt_style.dim() == 5:
        assert gt_style.shape[1] == 1, "Loss model does not support multiple styles."
        gt_style = gt_style[:, 0]
    fc = vgg16_features(W, torch.as_tensor(gt_content).to(dtype), tf32)
    fs = vgg16_features(W, gt_style, tf32)
    fp = vgg16_features(W, pred, tf32)
    feature = torch.stack([mean_l2_loss_on_batch(fp[l] - fc[l]) for l in VGG_CONTENT_LAYERS]).mean(0) * content_factor
    style = torch.stack([mean_l2_loss_on_batch(gram_matrix(fp[l]) - gram_matrix(fs[l]))
                         for l in VGG_STYLE_LAYERS]).mean(0) * style_factor
    tv = total_variation(pred) * tv_factor
    return {"loss": feature + style + tv, "feature_loss": feature, "style_loss": style,
            "total_variation_loss": tv}


# --------------------------------------------------------------------------------------
# Synthetic inputs
# --------------------------------------------------------------------------------------
def synthetic_content(batch: int, h: int, w: int, channels: Sequence[Tuple[str, int]], seed: int = 0,
                      unit_depth: bool = False) -> np.ndarray:
    """Channel-wise synthetic G-buffer in ShapeConfig.channels order."""
    g = torch.Generator().manual_seed(seed)
    parts = []
    for name, n in channels:
        if name == "FinalImage":
            t = torch.rand((batch, h, w, n), generator=g) * 4.0
        elif name == "ViewNormal":
            t = torch.randn((batch, h, w, n), generator=g)
            t = t / t.norm(dim=-1, keepdim=True).clamp_min(1e-6)
        elif name == "SceneDepth":
            t = torch.rand((batch, h, w, n), generator=g)
            if not unit_depth:
                t = 10.0 + t * (1e4 - 10.0)
        else:
            t = torch.rand((batch, h, w, n), generator=g)
        parts.append(t)
    return torch.cat(parts, dim=-1).numpy().astype(np.float32)


def synthetic_style_weights(batch: int, h: int, w: int, seed: int = 5) -> np.ndarray:
    """Smooth low-frequency weight field in [0,1]: bilinear upsample of an 8x16 U[0,1) grid."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand((batch, 1, 8, 16), generator=g)
    fine = F.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True)
    return fine.permute(0, 2, 3, 1).contiguous().numpy().astype(np.float32)


# --------------------------------------------------------------------------------------
# Training step (train_network.py:102-138 + Keras train_step; styleTransferTrainingModel.py:26-33)
# --------------------------------------------------------------------------------------
RMSPROP = dict(lr=1e-3, rho=0.9, eps=1e-7)      # tf.keras.optimizers.RMSprop() defaults in TF 2.9 (momentum 0, not centred)


def training_forward_backward(spec: TransferSpec, transfer_w, extractor: str, predictor_w, vgg_w, content, style, gt_content,
                              dtype=torch.float64, moving=None, tap_grads=None, pred_grad=None, vgg_tf32=False):
    """One forward/backward of the reference's training model with num_styles == 1:
    y_pred = transfer(content, predictor(style)) in training mode (BatchNorm uses batch statistics), loss vector from the
    VGG loss model, gradient of the batch SUM of the loss w.r.t. every trainable variable (VGG is frozen).
    pred_grad: back-propagate this d(loss)/d(y_pred) instead of the loss model's own gradient (isolates the network's
    backward pass from the loss, whose gradient is very sensitive to rounding noise in y_pred).
    Returns (losses dict of (B,) tensors, grads {name: tensor}, y_pred)."""
    tw = {k: torch.as_tensor(v).to(dtype).requires_grad_(not k.endswith(("moving_mean", "moving_variance")))
          for k, v in transfer_w.items()}
    pw = {k: torch.as_tensor(v).to(dtype).requires_grad_(not k.endswith(("moving_mean", "moving_variance")))
          for k, v in predictor_w.items()}
    style_t = torch.as_tensor(style).to(dtype)
    if style_t.dim() == 5:
        style_t = style_t[:, 0]
    params = _predictor_forward_t(extractor, pw, style_t, training=True)[:, None, :]
    taps = {} if tap_grads is not None else None
    y = _transfer_forward_t(spec, tw, torch.as_tensor(content).to(dtype), params, training=True, moving=moving, taps=taps)
    if taps is not None:
        taps["style_params"] = params
    losses = style_loss_vgg(vgg_w, y, gt_content, style_t, dtype=dtype, tf32=vgg_tf32)
    total = losses["loss"].sum()
    names = [("t", k) for k, v in tw.items() if v.requires_grad] + [("p", k) for k, v in pw.items() if v.requires_grad]
    tensors = [tw[k] if w == "t" else pw[k] for w, k in names]
    tap_names = list(taps) if taps is not None else []
    if pred_grad is not None:
        gs = torch.autograd.grad(y, tensors + [taps[k] for k in tap_names], grad_outputs=torch.as_tensor(pred_grad).to(dtype),
                                 allow_unused=True)
    else:
        gs = torch.autograd.grad(total, tensors + [taps[k] for k in tap_names], allow_unused=True)
    if tap_grads is not None:      # name -> (activation, d loss-sum / d activation) of the "<layer>/out" tensors
        for k, g in zip(tap_names, gs[len(tensors):]):
            tap_grads[k] = (taps[k].detach(), g)
    gs = gs[:len(tensors)]
    grads = {k: (g if g is not None else torch.zeros_like(t)) for (w, k), g, t in zip(names, gs, tensors)}
    return {k: v.detach() for k, v in losses.items()}, grads, y.detach()


def _transfer_forward_t(spec, W, x, style_params, training, moving=None, taps=None):
    """transfer_forward on already-converted torch tensors (keeps the autograd graph), single style.
    moving (a dict) receives the updated BatchNorm moving statistics of a training-mode call."""
    sp = style_params.unsqueeze(1)
    cursor = 0
    for name, _, co, k, s in spec.contract:
        p = f"contract_{name}"
        x = F.relu(conv2d_same(x, W[f"{p}/conv/kernel"], W[f"{p}/conv/bias"], s))
        mo = [] if moving is not None else None
        x = F.relu(batchnorm(x, W[f"{p}/bn/gamma"], W[f"{p}/bn/beta"], W[f"{p}/bn/moving_mean"],
                             W[f"{p}/bn/moving_variance"], training=training, moving_out=mo))
        if mo:
            moving[f"{p}/bn/moving_mean"], moving[f"{p}/bn/moving_variance"] = mo[0]
        if taps is not None:
            taps[f"{p}/out"] = x
    f = spec.filters
    for b in range(5):
        params = sp[..., cursor:cursor + 4 * f]
        cursor += 4 * f
        fx = x
        for i in range(2):
            p = f"residual_block_{b}/conv{i}"
            fx = F.relu(conv2d_same(fx, W[f"{p}/kernel"], W[f"{p}/bias"], 1))
            sl = params[..., 2 * f * i: 2 * f * (i + 1)]
            fx = cin(fx, sl[..., :f], sl[..., f:])
            if i == 0:
                fx = F.relu(fx)
                if taps is not None:
                    taps[f"{p}/out"] = fx
        x = fx if b == 0 else x + fx
        if taps is not None:
            taps[f"residual_block_{b}/conv1/out"] = x
    for name, _, co, k, s in spec.expand:
        p = f"expand_{name}"
        params = sp[..., cursor:cursor + 2 * co]
        cursor += 2 * co
        x = conv2d_transpose_same(x, W[f"{p}/conv/kernel"], W[f"{p}/conv/bias"], s)
        x = cin(x, params[..., :co], params[..., co:])
        x = torch.sigmoid(x) if name == "last" else F.relu(x)
        if taps is not None:
            taps[f"{p}/out"] = x
    return x


def _predictor_forward_t(extractor, W, x, training):
    if extractor == "DUMMY":
        x = conv2d_same(x, W["dummy_conv/kernel"], W["dummy_conv/bias"], 5)
    elif extractor == "MOBILE_NET":
        x = mobilenet_v3_small(W, x * 2.0 - 1.0, training)
    else:
        raise ValueError(f"{extractor} is not a valid value for feature_extractor")
    x = x.mean(dim=(1, 2), keepdim=True)
    x = conv2d_same(x, W["StylePredictor/kernel"], W["StylePredictor/bias"])
    x = conv2d_same(x, W["StyleNormPredictor/kernel"], W["StyleNormPredictor/bias"])
    return x[:, 0, 0, :]


def rmsprop_update(weights: Dict[str, np.ndarray], grads: Dict[str, Tensor], slots: Dict[str, np.ndarray],
                   lr=RMSPROP["lr"], rho=RMSPROP["rho"], eps=RMSPROP["eps"]):
    """Keras RMSprop (TF 2.9 optimizer_v2/rmsprop.py, momentum 0, centered False):
    rms = rho*rms + (1-rho)*g^2 ;  var -= lr * g / (sqrt(rms) + eps).   parity unpinned (TF is un-vendored)."""
    new_w, new_s = dict(weights), dict(slots)
    for k, g in grads.items():
        g = g.detach().double().numpy()
        rms = rho * slots.get(k, np.zeros_like(g)) + (1 - rho) * g * g
        new_s[k] = rms
        new_w[k] = (weights[k].astype(np.float64) - lr * g / (np.sqrt(rms) + eps)).astype(np.float32)
    return new_w, new_s
