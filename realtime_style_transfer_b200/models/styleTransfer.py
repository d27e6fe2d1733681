"""Style-conditioned transfer network -- host mirror of the reference's
realtime_style_transfer/models/styleTransfer.py (same public names and argument meaning).

``create_style_transfer_model`` returns ``(model, num_style_parameters)`` exactly like the reference
(styleTransfer.py:213-214, :331-332).  The model's forward runs in librst_sm100.so:
encoder convs + BatchNorm, five residual blocks with conditional instance normalisation, transposed
conv decoder, sigmoid head, and the dual-style weight-map blend.
"""
from __future__ import annotations

import logging
import typing

import numpy as np

from .. import _native
from .._plan import NUM_PARAMS_PER_FEATURE, TransferPlan
from ._base import InputSpec, NativeModel, _is_torch_cuda, as_content, as_numpy

log = logging.getLogger(__name__)


class StyleParamStack:
    """Running cursor over the last axis of the style-parameter tensor (styleTransfer.py:12-33)."""

    def __init__(self, style_params, style_weights):
        self.style_params = style_params
        self.style_weights = style_weights
        self.lower_bound = 0

    def get_params(self, num_params):
        lo = self.lower_bound
        self.lower_bound = lo + num_params
        return self.style_params[..., lo:lo + num_params]

    def make_content_and_style_input(self, content, num_params):
        inputs = {"content": content, "style_params": self.get_params(num_params)}
        if self.style_weights is not None:
            inputs["style_weights"] = self.style_weights
        return inputs


def _apply_style_weights(style_weights, style_params):
    """Per-pixel blend of two styles' parameters (styleTransfer.py:36-44), computed on the GPU.

    style_weights (B,H,W,2), style_params (B,1,2,F) -> (B,H,W,F); with any other style count the
    parameters pass through untouched, as in the reference."""
    params = as_numpy(style_params)
    if params.shape[-2] != 2:
        return params
    import torch
    weights = as_numpy(style_weights)
    b, h, w, _ = weights.shape
    f = params.shape[-1]
    dev = torch.device("cuda", torch.cuda.current_device())
    d_w = torch.from_numpy(weights).to(dev)
    d_p = torch.from_numpy(params.reshape(b, 2, f)).to(dev)
    d_o = torch.empty((b, h, w, f), dtype=torch.float32, device=dev)
    _native.op_apply_style_weights(d_w.data_ptr(), d_p.data_ptr(), d_o.data_ptr(), b, h, w, f,
                                   torch.cuda.current_stream().cuda_stream)
    return d_o.cpu().numpy()


class ConditionalInstanceNormalization:
    """Shape helpers of the reference layer (styleTransfer.py:47-92).  The layer owns no variables:
    scale and bias arrive as style parameters; normalisation runs fused in the native forward."""
    NumParamsPerFeature: int = NUM_PARAMS_PER_FEATURE

    def __init__(self, num_feature_maps, num_styles, name, epsilon=1e-5):
        self.name = f"ConditionalInstanceNormalization_{name}"
        self.epsilon = epsilon
        self.num_feature_maps = num_feature_maps
        self.num_styles = num_styles

    def __call__(self, inputs, activation=_native.ACT_NONE):
        import torch
        x = as_numpy(inputs["content"])
        params = as_numpy(inputs["style_params"])            # (B,1,S,2F)
        b, h, w, f = x.shape
        s = params.shape[-2]
        dev = torch.device("cuda", torch.cuda.current_device())
        d_x = torch.from_numpy(x).to(dev)
        d_p = torch.from_numpy(params.reshape(b, s, 2 * f)).to(dev)
        d_w = torch.from_numpy(as_numpy(inputs["style_weights"])).to(dev) if self.num_styles > 1 else None
        d_y = torch.empty_like(d_x)
        _native.op_cin(d_x.data_ptr(), d_p.data_ptr(), d_w.data_ptr() if d_w is not None else 0, d_y.data_ptr(),
                       b, h, w, f, s, activation, torch.cuda.current_stream().cuda_stream)
        return d_y.cpu().numpy()

    def get_config(self):
        return {"epsilon": self.epsilon, "num_feature_maps": self.num_feature_maps,
                "num_styles": self.num_feature_maps}

    @classmethod
    def get_style_params_shape_and_num(cls, num_feature_maps, num_styles):
        num_style_params = cls.NumParamsPerFeature * num_feature_maps
        return (1, num_styles, num_style_params), num_style_params

    @classmethod
    def get_style_weights_shape(cls, content_shape: typing.Tuple, num_styles, multiplier=1) -> typing.Tuple:
        shape = list(content_shape)
        shape[-1] = num_styles
        shape[-2] *= multiplier
        shape[-3] *= multiplier
        return tuple(shape)


def calc_next_conv_dims(initial, filters, mult) -> typing.Tuple:
    if len(initial) == 3:
        return (int(initial[0] * mult), int(initial[1] * mult), filters)
    return (initial[0], initial[1] * mult, initial[2] * mult, filters)


# ---- the reference's sub-model builders (styleTransfer.py:95-205, :335-345) -----------------------------------------------------
# create_style_transfer_model runs the whole network natively; these stand-alone builders exist for callers that assemble or
# inspect single blocks as the reference module allows.  Each returns a callable with the reference's name, variables drawn from
# the reference's initialisers and input dictionary; the arithmetic runs through the operator-level C ABI (rst_op_conv2d,
# rst_op_cin), BatchNorm in inference mode on the host (one affine per channel).
_ACT_CODES = {None: _native.ACT_NONE, "linear": _native.ACT_NONE, "relu": _native.ACT_RELU, "sigmoid": _native.ACT_SIGMOID}


def _act_code(activation):
    if activation is None or isinstance(activation, str):
        if activation not in _ACT_CODES:
            raise ValueError(f"unsupported activation {activation!r}: relu, sigmoid or None")
        return _ACT_CODES[activation]
    name = getattr(activation, "__name__", str(activation)).lower()          # tf.nn.relu / tf.nn.sigmoid style callables
    for key in ("relu", "sigmoid"):
        if key in name:
            return _ACT_CODES[key]
    raise ValueError(f"unsupported activation {activation!r}: relu, sigmoid or None")


def _conv_op(x: np.ndarray, kernel: np.ndarray, bias: np.ndarray, stride: int, transposed: bool, act: int) -> np.ndarray:
    """Conv2D / Conv2DTranspose with padding='same' on the GPU (rst_op_conv2d)."""
    import torch
    b, h, w, ci = x.shape
    kh, kw = kernel.shape[:2]
    co = kernel.shape[2] if transposed else kernel.shape[3]
    ho, wo = (h * stride, w * stride) if transposed else (-(-h // stride), -(-w // stride))
    dev = torch.device("cuda", torch.cuda.current_device())
    d_x, d_k, d_b = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev) for a in (x, kernel, bias))
    d_y = torch.empty((b, ho, wo, co), dtype=torch.float32, device=dev)
    _native.op_conv2d(d_x.data_ptr(), d_k.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), b, h, w, ci, co, kh, kw, stride,
                      transposed, act, _native.PRECISION_FP32, torch.cuda.current_stream().cuda_stream)
    return d_y.cpu().numpy()


class _Block:
    """A sub-model of the transfer network: name, ``variables`` (name -> float32 array, reference layouts), input / output shape."""

    def __init__(self, name, input_shape, output_shape, variables):
        self.name, self.input_shape, self.output_shape, self.variables = name, tuple(input_shape), tuple(output_shape), variables

    @property
    def weights(self):
        return list(self.variables.values())

    def get_weights(self):
        return [v.copy() for v in self.variables.values()]

    def set_weights(self, values):
        values = list(values)
        if len(values) != len(self.variables):
            raise ValueError(f"{self.name}: expected {len(self.variables)} arrays, got {len(values)}")
        for (k, old), new in zip(list(self.variables.items()), values):
            new = np.asarray(new, dtype=np.float32)
            if new.shape != old.shape:
                raise ValueError(f"{self.name}/{k}: shape {new.shape} != {old.shape}")
            self.variables[k] = new


class _ContractBlock(_Block):
    def __init__(self, input_shape, filters, size, strides, name, rng):
        h, w, ci = input_shape
        variables = {"conv/kernel": rng.normal(0.0, 0.02, (size, size, ci, filters)).astype(np.float32),
                     "conv/bias": np.zeros(filters, np.float32), "bn/gamma": np.ones(filters, np.float32),
                     "bn/beta": np.zeros(filters, np.float32), "bn/moving_mean": np.zeros(filters, np.float32),
                     "bn/moving_variance": np.ones(filters, np.float32)}
        super().__init__(f"contract_{name}", input_shape, (-(-h // strides), -(-w // strides), filters), variables)
        self.strides = strides

    def __call__(self, x):
        """ReLU(BN(ReLU(conv(x) + b))) with the moving statistics (styleTransfer.py:194-203; Keras BN eps 1e-3)."""
        v = self.variables
        y = _conv_op(as_numpy(x), v["conv/kernel"], v["conv/bias"], self.strides, False, _native.ACT_RELU)
        scale = v["bn/gamma"] / np.sqrt(v["bn/moving_variance"] + np.float32(1e-3))
        return np.maximum(y * scale + (v["bn/beta"] - v["bn/moving_mean"] * scale), 0.0).astype(np.float32)


class _StyledBlock(_Block):
    def _cin(self, x, params, style_weights, act, name):
        layer = ConditionalInstanceNormalization(x.shape[-1], self.num_styles, name)
        inputs = {"content": x, "style_params": params}
        if self.num_styles > 1:
            inputs["style_weights"] = style_weights
        return layer(inputs, activation=act)


class _ResidualBlock(_StyledBlock):
    def __init__(self, input_shape, num_styles, filters, size, strides, name, is_first, rng):
        h, w, ci = input_shape
        variables = {"conv0/kernel": rng.uniform(0.0, 0.05, (size, size, ci, filters)).astype(np.float32),
                     "conv0/bias": np.zeros(filters, np.float32),
                     "conv1/kernel": rng.uniform(0.0, 0.05, (size, size, filters, filters)).astype(np.float32),
                     "conv1/bias": np.zeros(filters, np.float32)}
        super().__init__(f"residual_block_{name}", input_shape, (h, w, filters), variables)
        self.num_styles, self.filters, self.strides, self.is_first = num_styles, filters, strides, is_first
        self.num_style_parameters = 2 * NUM_PARAMS_PER_FEATURE * filters          # [scale0 | bias0 | scale1 | bias1]

    def __call__(self, inputs):
        """fx = CIN_i(ReLU(conv_i(fx) + b)) for i in 0, 1 with a ReLU after the first CIN only; + x unless is_first
        (styleTransfer.py:167-184).  inputs: {'content', 'style_params' (B,1,S,4F)[, 'style_weights' (B,H,W,S)]}."""
        x = as_numpy(inputs["content"])
        params = as_numpy(inputs["style_params"])
        sw = inputs.get("style_weights")
        f, v = self.filters, self.variables
        fx = x
        for i in range(2):
            fx = _conv_op(fx, v[f"conv{i}/kernel"], v[f"conv{i}/bias"], self.strides, False, _native.ACT_RELU)
            fx = self._cin(fx, params[..., 2 * f * i:2 * f * (i + 1)], sw, _native.ACT_RELU if i == 0 else _native.ACT_NONE,
                           f"{self.name}_{i}")
        return fx if self.is_first else (x + fx).astype(np.float32)


class _ExpandBlock(_StyledBlock):
    def __init__(self, input_shape, num_styles, filters, size, strides, name, activation, rng):
        h, w, ci = input_shape
        variables = {"conv/kernel": rng.normal(0.0, 0.02, (size, size, filters, ci)).astype(np.float32),      # (kh,kw,out,in)
                     "conv/bias": np.zeros(filters, np.float32)}
        super().__init__(f"expand_{name}", input_shape, (h * strides, w * strides, filters), variables)
        self.num_styles, self.filters, self.strides, self.act = num_styles, filters, strides, _act_code(activation)
        self.num_style_parameters = NUM_PARAMS_PER_FEATURE * filters

    def __call__(self, inputs):
        """act(CIN(convT(x) + b)) (styleTransfer.py:115-139).  inputs as for residual_block, 'style_params' (B,1,S,2F); the
        style weights are those of the OUTPUT resolution."""
        v = self.variables
        y = _conv_op(as_numpy(inputs["content"]), v["conv/kernel"], v["conv/bias"], self.strides, True, _native.ACT_NONE)
        return self._cin(y, as_numpy(inputs["style_params"]), inputs.get("style_weights"), self.act, self.name)


def expand(input_shape: typing.Tuple, num_styles, filters, size, strides, name, activation="relu", seed=None):
    """styleTransfer.py:95-141: Conv2DTranspose('same') -> ConditionalInstanceNormalization -> activation."""
    return _ExpandBlock(input_shape, num_styles, filters, size, strides, name, activation, np.random.default_rng(seed))


def residual_block(input_shape: typing.Tuple, num_styles, filters, size, strides, name, is_first=False, seed=None):
    """styleTransfer.py:144-185."""
    return _ResidualBlock(input_shape, num_styles, filters, size, strides, name, is_first, np.random.default_rng(seed))


def contract(input_shape, filters, size, strides, name, seed=None):
    """styleTransfer.py:188-205: Conv2D('same') -> ReLU -> BatchNormalization -> ReLU."""
    return _ContractBlock(input_shape, filters, size, strides, name, np.random.default_rng(seed))


def _get_style_weight_mips(style_weights, num_mips):
    """styleTransfer.py:335-345: {width: weights}, each level AvgPool2D(2) ('valid': odd trailing rows / columns dropped) of the
    one above; host arithmetic (the native forward builds the same pyramid on the device, rst_api.cu::build_mips)."""
    level = np.asarray(as_numpy(style_weights), dtype=np.float32)
    mips = {level.shape[-2]: level}
    for _ in range(num_mips):
        b, h, w, c = level.shape
        level = level[:, :h // 2 * 2, :w // 2 * 2].reshape(b, h // 2, 2, w // 2, 2, c).mean(axis=(2, 4), dtype=np.float32)
        mips[level.shape[-2]] = level
    return mips


class StyleTransferModel(NativeModel):
    """Object returned by create_style_transfer_model; inputs dict {'content','style_params'[,'style_weights']}."""

    def __init__(self, plan: TransferPlan, name: str, seed=None):
        super().__init__(name)
        self.plan = plan
        self.num_style_parameters = plan.num_style_parameters
        self._variables = plan.initial_weights(np.random.default_rng(seed))
        self.input = {
            "content": InputSpec(plan.input_shape, "content"),
            "style_params": InputSpec((plan.num_styles, plan.num_style_parameters), "style_params"),
        }
        if plan.num_styles > 1:
            self.input["style_weights"] = InputSpec(
                ConditionalInstanceNormalization.get_style_weights_shape(plan.output_shape, plan.num_styles - 1),
                "style_weights")
        self.inputs = self.input
        self.output_shape = (None,) + plan.output_shape

    def _context_kwargs(self, max_batch):
        p = self.plan
        return dict(in_shape=p.input_shape, out_shape=p.output_shape, bottleneck_res_y=p.bottleneck_res_y,
                    bottleneck_num_filters=p.filters, num_styles=p.num_styles)

    def _check_inputs(self, inputs):
        content = inputs["content"]
        params = inputs["style_params"]
        if tuple(content.shape[1:]) != self.plan.input_shape:
            raise ValueError(f"content shape {tuple(content.shape)} incompatible with {self.input['content'].shape}")
        if tuple(params.shape[1:]) != (self.plan.num_styles, self.num_style_parameters):
            raise ValueError(f"style_params shape {tuple(params.shape)} incompatible with "
                             f"{self.input['style_params'].shape}")
        weights = inputs.get("style_weights") if self.plan.num_styles > 1 else None
        if self.plan.num_styles > 1:
            if weights is None:
                raise ValueError("style_weights input required when num_styles > 1")
            expect = self.plan.output_shape[:2] + (self.plan.num_styles - 1,)
            assert tuple(weights.shape[1:]) == expect, \
                f"Style weights must be the same dimensions as the output shape. {tuple(weights.shape)} vs. {expect}"
        return content, params, weights

    def __call__(self, inputs, training=False):
        """numpy in -> numpy out (host path); torch CUDA tensors in -> torch CUDA tensor out (device path)."""
        content, params, weights = self._check_inputs(inputs)
        if _is_torch_cuda(content):
            import torch
            b = content.shape[0]
            ctx = self._get_ctx(b)
            content = content.contiguous().float()
            params = params.contiguous().float().to(content.device)
            weights = weights.contiguous().float().to(content.device) if weights is not None else None
            out = torch.empty((b,) + self.plan.output_shape, dtype=torch.float32, device=content.device)
            ctx.transfer_forward_device(content.data_ptr(), params.data_ptr(),
                                        weights.data_ptr() if weights is not None else None, out.data_ptr(), b,
                                        torch.cuda.current_stream(content.device).cuda_stream)
            return out
        return self.predict(inputs)

    def predict(self, x, batch_size=None, verbose=0, output_dtype=np.float32, **kwargs):
        """Keras' predict.  Extensions: a float16 'content' array is uploaded as float16, and output_dtype=np.uint8 returns
        trunc(255 * y) computed on the device -- the quantisation the reference's callers apply to the result
        (predict_using_checkpoint.py:99, predict_video_using_checkpoint.py:98)."""
        content, params, weights = self._check_inputs(x)
        content, params = as_content(content), as_numpy(params)
        weights = as_numpy(weights) if weights is not None else None
        n = content.shape[0]
        step = n if not batch_size else int(batch_size)
        ctx = self._get_ctx(min(step, n) if n else 1)
        outs = []
        for i in range(0, n, max(step, 1)):
            outs.append(ctx.transfer_forward_host(content[i:i + step], params[i:i + step],
                                                  weights[i:i + step] if weights is not None else None, out_dtype=output_dtype))
        if not outs:
            return np.zeros((0,) + self.plan.output_shape, np.dtype(output_dtype))
        return np.concatenate(outs, axis=0) if len(outs) > 1 else outs[0]


def _pinned_view(x):
    """A pinned torch CPU tensor (float32 / float16, contiguous) -> its numpy view, else None."""
    if hasattr(x, "is_pinned") and x.is_pinned() and x.is_contiguous() and str(x.dtype) in ("torch.float32", "torch.float16"):
        return x.numpy()
    return None


def _predict_frames(self, batches, pinned: bool = True, output_dtype=np.float32):
    """Streaming variant of the reference's video loop (predict_video_using_checkpoint.py:90-98): yields one
    stylised numpy batch per input dict, with the host<->device copies of neighbouring batches overlapping the forward.
    numpy inputs are staged through two pinned buffers (one host copy per batch); a 'content' given as a PINNED torch CPU
    tensor is handed to the DMA engine as it is (no host copy -- the caller must not modify it until its result has been
    yielded).  float16 'content' crosses PCIe as float16 and output_dtype=np.uint8 returns trunc(255 * y) (what the reference's
    loop computes from the float result at :98): together 136 MB instead of 295 MB per batch of 8 frames of rst-960-120-128-17."""
    import torch
    slots, pending = [], []
    ctx = None
    for k, element in enumerate(batches):
        content, params, weights = self._check_inputs(element)
        direct = _pinned_view(content)
        content = direct if direct is not None else as_content(content)
        params = as_numpy(params)
        weights = as_numpy(weights) if weights is not None else None
        b = content.shape[0]
        if ctx is None:
            ctx = self._get_ctx(b)
            in_dtype = content.dtype             # the first batch decides the element type of the staging buffers
            for _ in range(2):
                def mk(shape, dtype=np.float32):
                    if pinned:
                        return torch.empty(shape, dtype=getattr(torch, np.dtype(dtype).name)).pin_memory().numpy()
                    return np.empty(shape, dtype)
                slots.append({"content": mk((ctx.cfg.max_batch,) + self.plan.input_shape, in_dtype),
                              "params": mk((ctx.cfg.max_batch, self.plan.num_styles, self.num_style_parameters)),
                              "weights": mk((ctx.cfg.max_batch,) + self.plan.output_shape[:2] + (max(self.plan.num_styles - 1, 1),)),
                              "out": mk((ctx.cfg.max_batch,) + self.plan.output_shape, output_dtype)})
        if len(pending) == 2:
            t, sl, n, _keep = pending.pop(0)
            ctx.transfer_wait(t)
            yield sl["out"][:n].copy()
        sl = slots[k % 2]
        if direct is None or content.dtype != sl["content"].dtype:
            sl["content"][:b] = content
            src = sl["content"][:b]
        else:
            src = content
        sl["params"][:b] = params
        if weights is not None:
            sl["weights"][:b] = weights
        ticket = ctx.transfer_submit_host(src, sl["params"][:b], sl["weights"][:b] if weights is not None else None, sl["out"][:b])
        pending.append((ticket, sl, b, element))          # the element keeps a directly-used pinned tensor alive
    for t, sl, n, _keep in pending:
        ctx.transfer_wait(t)
        yield sl["out"][:n].copy()


StyleTransferModel.predict_frames = _predict_frames


def create_style_transfer_model(input_shape, output_shape, bottleneck_res_y, bottleneck_num_filters, num_styles,
                                name="StyleTransferModel"):
    log.info(f"Using {num_styles} styles")
    plan = TransferPlan(input_shape, output_shape, bottleneck_res_y, bottleneck_num_filters, num_styles)
    log.info(f"Contracting with {plan.num_contract_blocks} blocks to "
             f"{plan.bottleneck_hw[1]}x{plan.bottleneck_hw[0]}x{bottleneck_num_filters}")
    log.info(f"expanding with {plan.num_expand_blocks} blocks to "
             f"{output_shape[1]}x{output_shape[0]}x{output_shape[2]}")
    model = StyleTransferModel(plan, name)
    return model, plan.num_style_parameters
