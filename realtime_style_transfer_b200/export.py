"""ONNX export of the transfer network and the style predictor -- the counterpart of the reference's
``save_using_checkpoint.py:76-103`` (``tf2onnx.convert.from_keras`` on ``.transfer`` and ``.style_predictor``).

Neither TensorFlow, tf2onnx nor the ``onnx`` package exist in this environment, so the ModelProto is written directly in the
protobuf wire format (onnx/onnx.proto3, opset 13): the file is what ``onnx.save`` would produce for the same graph and loads in
onnxruntime / Unreal's NNI like the reference's exports.  The graphs keep the Keras models' interface: NHWC float32 inputs named
``content`` / ``style_params`` / ``style_weights`` (transfer) and ``style`` (predictor), a symbolic batch dimension ``N``.

  transfer:   Transpose -> [Conv, Relu, BatchNormalization, Relu] x (1 + contract blocks)
              -> 5 x residual block [Conv, Relu, InstanceNormalization(eps 1e-5), Mul/Add with the sliced style parameters]
              -> [ConvTranspose, InstanceNormalization, Mul/Add, Relu | Sigmoid] x (expand blocks + 1) -> Transpose
              (two styles: the per-pixel blend ``w0*p0 + w1*p1`` of styleTransfer.py:36-44 over AveragePool mips of the weight map)
  predictor:  x*2-1 -> MobileNetV3Small (Conv / BatchNormalization / HardSigmoid / Mul / GlobalAveragePool) -> two 1x1 Conv.

``tests/test_export.py`` executes the written files with a small ONNX interpreter (oracle/onnx_ref.py, test infrastructure) and
compares them with the oracle network.  This module is host-side serialisation only: no arithmetic of the hot path lives here.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Sequence

import numpy as np

from ._plan import PredictorPlan, TransferPlan

OPSET = 13
IR_VERSION = 8
_FLOAT, _INT64 = 1, 7          # TensorProto.DataType


# ---- protobuf wire format ------------------------------------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def _key(field: int, wire: int) -> bytes:
    return _varint(field << 3 | wire)


def _f_varint(field: int, v: int) -> bytes:
    return _key(field, 0) + _varint(int(v))


def _f_bytes(field: int, b: bytes) -> bytes:
    return _key(field, 2) + _varint(len(b)) + b


def _f_str(field: int, s: str) -> bytes:
    return _f_bytes(field, s.encode())


def _f_float(field: int, v: float) -> bytes:
    return _key(field, 5) + struct.pack("<f", v)


def _tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    dtype = _INT64 if arr.dtype == np.int64 else _FLOAT
    if dtype == _FLOAT:
        arr = arr.astype("<f4")
    body = b"".join(_f_varint(1, d) for d in arr.shape) + _f_varint(2, dtype) + _f_str(8, name) + _f_bytes(9, arr.tobytes())
    return body


def _attribute(name: str, value) -> bytes:
    body = _f_str(1, name)
    if isinstance(value, float):
        body += _f_float(2, value) + _f_varint(20, 1)
    elif isinstance(value, (int, np.integer)):
        body += _f_varint(3, int(value)) + _f_varint(20, 2)
    elif isinstance(value, str):
        body += _f_bytes(4, value.encode()) + _f_varint(20, 3)
    elif isinstance(value, (list, tuple)) and all(isinstance(v, (int, np.integer)) for v in value):
        body += b"".join(_f_varint(8, int(v)) for v in value) + _f_varint(20, 7)
    elif isinstance(value, (list, tuple)):
        body += b"".join(_f_float(7, float(v)) for v in value) + _f_varint(20, 6)
    else:
        raise TypeError(f"attribute {name}: unsupported value {value!r}")
    return body


def _value_info(name: str, shape: Sequence) -> bytes:
    dims = b""
    for d in shape:
        dims += _f_bytes(1, _f_str(2, d) if isinstance(d, str) else _f_varint(1, int(d)))
    tensor_type = _f_varint(1, _FLOAT) + _f_bytes(2, dims)
    return _f_str(1, name) + _f_bytes(2, _f_bytes(1, tensor_type))


class _Graph:
    def __init__(self, name: str):
        self.name = name
        self.nodes: List[bytes] = []
        self.initializers: List[bytes] = []
        self.inputs: List[bytes] = []
        self.outputs: List[bytes] = []
        self._n = 0

    def _fresh(self, hint: str) -> str:
        self._n += 1
        return f"{hint}_{self._n}"

    def const(self, arr, hint="const") -> str:
        name = self._fresh(hint)
        self.initializers.append(_tensor(name, np.asarray(arr)))
        return name

    def node(self, op: str, inputs: Sequence[str], name: str = "", **attrs) -> str:
        out = self._fresh(name or op.lower())
        body = b"".join(_f_str(1, i) for i in inputs) + _f_str(2, out) + _f_str(3, out) + _f_str(4, op)
        body += b"".join(_f_bytes(5, _attribute(k, v)) for k, v in attrs.items())
        self.nodes.append(body)
        return out

    def rename_last_output(self, new: str):
        """Gives the most recent node's output the public name of a graph output."""
        body = self.nodes[-1]
        fields, pos, out = [], 0, None
        while pos < len(body):
            start = pos
            tag = body[pos]; pos += 1                               # all tags here are < 16: one byte
            n, shift = 0, 0
            while True:
                b = body[pos]; pos += 1
                n |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            payload = body[pos:pos + n]
            pos += n
            fields.append((tag >> 3, payload))
            del start
        rebuilt = b""
        for fn, payload in fields:
            if fn in (2, 3):
                payload = new.encode()
            rebuilt += _f_bytes(fn, payload)
        self.nodes[-1] = rebuilt
        return new

    def serialize(self, producer="realtime_style_transfer_b200") -> bytes:
        g = b"".join(_f_bytes(1, n) for n in self.nodes) + _f_str(2, self.name)
        g += b"".join(_f_bytes(5, t) for t in self.initializers)
        g += b"".join(_f_bytes(11, i) for i in self.inputs) + b"".join(_f_bytes(12, o) for o in self.outputs)
        opset = _f_str(1, "") + _f_varint(2, OPSET)
        return _f_varint(1, IR_VERSION) + _f_str(2, producer) + _f_str(3, "0.2") + _f_bytes(7, g) + _f_bytes(8, opset)


def _same_pads(size: int, k: int, s: int):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _conv(g: _Graph, x: str, kernel: np.ndarray, bias, hw, stride: int, name: str, pads=None, depthwise=False) -> str:
    """Keras Conv2D(padding='same') (kernel (kh,kw,I,O)) or DepthwiseConv2D (kernel (kh,kw,C,1)) as an ONNX Conv on NCHW."""
    kh, kw = kernel.shape[:2]
    if pads is None:
        (pt, pb), (pl, pr) = _same_pads(hw[0], kh, stride), _same_pads(hw[1], kw, stride)
        pads = [pt, pl, pb, pr]
    w = kernel.transpose(2, 3, 0, 1) if depthwise else kernel.transpose(3, 2, 0, 1)
    ins = [x, g.const(w, name + "_W")]
    if bias is not None:
        ins.append(g.const(bias, name + "_B"))
    return g.node("Conv", ins, name, kernel_shape=[kh, kw], strides=[stride, stride], pads=list(pads),
                  group=int(kernel.shape[2]) if depthwise else 1)


def _bn(g: _Graph, x: str, w: Dict[str, np.ndarray], prefix: str, name: str, eps=1e-3) -> str:
    return g.node("BatchNormalization", [x, g.const(w[f"{prefix}/gamma"], name + "_g"), g.const(w[f"{prefix}/beta"], name + "_b"),
                                         g.const(w[f"{prefix}/moving_mean"], name + "_m"),
                                         g.const(w[f"{prefix}/moving_variance"], name + "_v")], name, epsilon=float(eps))


# ---- transfer network (styleTransfer.py:213-332) ----------------------------------------------------------------------------
def transfer_onnx_bytes(plan: TransferPlan, weights: Dict[str, np.ndarray]) -> bytes:
    h, w_, c = plan.input_shape
    S, P, F = plan.num_styles, plan.num_style_parameters, plan.filters
    g = _Graph("StyleTransferModel")
    g.inputs.append(_value_info("content", ["N", h, w_, c]))
    g.inputs.append(_value_info("style_params", ["N", S, P]))
    mips = {}
    if S > 1:
        assert S == 2, "only exactly two styles blend (styleTransfer.py:38)"
        oh, ow = plan.output_shape[:2]
        g.inputs.append(_value_info("style_weights", ["N", oh, ow, S - 1]))
        w1 = g.node("Transpose", ["style_weights"], "weights_nchw", perm=[0, 3, 1, 2])
        w0 = g.node("Sub", [g.const(np.float32(1.0).reshape(1, 1, 1, 1), "one"), w1], "weights_rest")      # 1 - sum(w), :297-302
        lvl = (w0, w1)
        width = ow
        mips[width] = lvl
        for _ in range(plan.num_expand_blocks + 1):                                                          # :335-345
            lvl = tuple(g.node("AveragePool", [t], "mip", kernel_shape=[2, 2], strides=[2, 2]) for t in lvl)
            width //= 2
            mips[width] = lvl
    cursor = [0]
    shape4 = g.const(np.asarray([0, -1, 1, 1], np.int64), "shape_nc11")

    def take(n: int) -> List[str]:
        """-> per style: the (N, n, 1, 1) slice [cursor, cursor+n) of that style's parameters."""
        a, b = cursor[0], cursor[0] + n
        cursor[0] = b
        out = []
        for s in range(S):
            sl = g.node("Slice", ["style_params", g.const(np.asarray([s, a], np.int64), "starts"),
                                  g.const(np.asarray([s + 1, b], np.int64), "ends"), g.const(np.asarray([1, 2], np.int64), "axes")],
                        "params")
            out.append(g.node("Reshape", [sl, shape4], "params_nc11"))
        return out

    def blend(per_style: List[str], width: int) -> str:
        if S == 1:
            return per_style[0]
        w0, w1 = mips[width]
        return g.node("Add", [g.node("Mul", [per_style[0], w0]), g.node("Mul", [per_style[1], w1])], "blend")   # :36-44

    def cin(x: str, ch: int, width: int, name: str) -> str:
        scale, bias = take(ch), take(ch)
        xn = g.node("InstanceNormalization", [x, g.const(np.ones(ch, np.float32), "ones"), g.const(np.zeros(ch, np.float32), "zeros")],
                    name + "_norm", epsilon=1e-5)                                                              # :57-71
        return g.node("Add", [g.node("Mul", [xn, blend(scale, width)]), blend(bias, width)], name)

    x = g.node("Transpose", ["content"], "content_nchw", perm=[0, 3, 1, 2])
    hw = (h, w_)
    names = ["start"] + [str(i) for i in range(plan.num_contract_blocks)]
    for i, n in enumerate(names):                                                                              # contract, :188-205
        p = f"contract_{n}"
        stride = 1 if i == 0 else 2
        x = g.node("Relu", [_conv(g, x, weights[f"{p}/conv/kernel"], weights[f"{p}/conv/bias"], hw, stride, p)])
        x = g.node("Relu", [_bn(g, x, weights, f"{p}/bn", p + "_bn")])
        hw = (-(-hw[0] // stride), -(-hw[1] // stride))
    for b in range(5):                                                                                         # residual_block, :144-185
        fx = x
        params_cursor = cursor[0]
        for i in range(2):
            p = f"residual_block_{b}/conv{i}"
            fx = g.node("Relu", [_conv(g, fx, weights[f"{p}/kernel"], weights[f"{p}/bias"], hw, 1, f"res{b}_conv{i}")])
            # parameter layout of a block: [scale0 F | bias0 F | scale1 F | bias1 F]
            cursor[0] = params_cursor + 2 * F * i
            fx = cin(fx, F, hw[1], f"res{b}_cin{i}")
            if i == 0:
                fx = g.node("Relu", [fx])
        cursor[0] = params_cursor + 4 * F
        x = fx if b == 0 else g.node("Add", [x, fx], f"res{b}_out")
    enames = [str(i) for i in range(plan.num_expand_blocks)] + ["last"]
    for n in enames:                                                                                           # expand, :95-141
        p = f"expand_{n}"
        k = weights[f"{p}/conv/kernel"]                          # Conv2DTranspose kernel (kh, kw, out, in)
        stride = 1 if n == "last" else 2
        kh, kw = k.shape[:2]
        (pt, pb), (pl, pr) = _same_pads(hw[0] * stride, kh, stride), _same_pads(hw[1] * stride, kw, stride)
        x = g.node("ConvTranspose", [x, g.const(k.transpose(3, 2, 0, 1), p + "_W"), g.const(weights[f"{p}/conv/bias"], p + "_B")],
                   p, kernel_shape=[kh, kw], strides=[stride, stride], pads=[pt, pl, pb, pr])
        hw = (hw[0] * stride, hw[1] * stride)
        x = cin(x, int(k.shape[2]), hw[1], p + "_cin")
        x = g.node("Sigmoid" if n == "last" else "Relu", [x])
    assert cursor[0] == P, (cursor[0], P)
    g.node("Transpose", [x], "stylised", perm=[0, 2, 3, 1])
    g.rename_last_output("stylised")
    g.outputs.append(_value_info("stylised", ["N"] + list(plan.output_shape)))
    return g.serialize()


# ---- style predictor (stylePrediction.py:25-75) --------------------------------------------------------------------------------
def predictor_onnx_bytes(plan: PredictorPlan, weights: Dict[str, np.ndarray]) -> bytes:
    from ._plan import _MBV3_SMALL, _depth
    h, w_, c = plan.input_shape
    g = _Graph("StylePredictionModel")
    g.inputs.append(_value_info("style", ["N", h, w_, c]))
    x = g.node("Transpose", ["style"], "style_nchw", perm=[0, 3, 1, 2])
    hw = (h, w_)

    def hswish(t):
        return g.node("Mul", [t, g.node("HardSigmoid", [t], alpha=1.0 / 6.0, beta=0.5)], "hswish")

    if plan.feature_extractor == "DUMMY":
        x = _conv(g, x, weights["dummy_conv/kernel"], weights["dummy_conv/bias"], hw, 5, "dummy_conv")
    else:
        x = g.node("Add", [g.node("Mul", [x, g.const(np.float32(2.0).reshape(1, 1, 1, 1), "two")]),
                           g.const(np.float32(-1.0).reshape(1, 1, 1, 1), "minus_one")], "rescale")           # Rescaling(2, -1)
        x = hswish(_bn(g, _conv(g, x, weights["mobilenet/Conv/kernel"], None, hw, 2, "stem"), weights, "mobilenet/Conv/BatchNorm", "stem_bn"))
        hw = (-(-hw[0] // 2), -(-hw[1] // 2))
        ci = 16
        for bid, (e, co, k, s, se) in enumerate(_MBV3_SMALL):
            p = "mobilenet/expanded_conv" + (f"_{bid}" if bid else "")
            act = (lambda t: g.node("Relu", [t])) if bid < 3 else hswish
            shortcut = x
            cexp = _depth(ci * e)
            if bid:
                x = act(_bn(g, _conv(g, x, weights[f"{p}/expand/kernel"], None, hw, 1, f"b{bid}_expand"), weights, f"{p}/expand/BatchNorm", f"b{bid}_expand_bn"))
            dk = weights[f"{p}/depthwise/depthwise_kernel"]
            if s == 2:                                             # ZeroPadding2D(correct_pad) + 'valid' (mobilenet_v3.py)
                pt, pb = k // 2 - (1 - hw[0] % 2), k // 2
                pl, pr = k // 2 - (1 - hw[1] % 2), k // 2
                x = _conv(g, x, dk, None, hw, 2, f"b{bid}_dw", pads=[pt, pl, pb, pr], depthwise=True)
                hw = ((hw[0] + pt + pb - k) // 2 + 1, (hw[1] + pl + pr - k) // 2 + 1)
            else:
                x = _conv(g, x, dk, None, hw, 1, f"b{bid}_dw", depthwise=True)
            x = act(_bn(g, x, weights, f"{p}/depthwise/BatchNorm", f"b{bid}_dw_bn"))
            if se:
                z = g.node("GlobalAveragePool", [x], f"b{bid}_se_pool")
                z = g.node("Relu", [_conv(g, z, weights[f"{p}/squeeze_excite/Conv/kernel"], weights[f"{p}/squeeze_excite/Conv/bias"], (1, 1), 1, f"b{bid}_se1")])
                z = _conv(g, z, weights[f"{p}/squeeze_excite/Conv_1/kernel"], weights[f"{p}/squeeze_excite/Conv_1/bias"], (1, 1), 1, f"b{bid}_se2")
                x = g.node("Mul", [x, g.node("HardSigmoid", [z], alpha=1.0 / 6.0, beta=0.5)], f"b{bid}_se")
            x = _bn(g, _conv(g, x, weights[f"{p}/project/kernel"], None, hw, 1, f"b{bid}_project"), weights, f"{p}/project/BatchNorm", f"b{bid}_project_bn")
            if s == 1 and ci == co:
                x = g.node("Add", [shortcut, x], f"b{bid}_add")
            ci = co
        x = hswish(_bn(g, _conv(g, x, weights["mobilenet/Conv_1/kernel"], None, hw, 1, "conv_1"), weights, "mobilenet/Conv_1/BatchNorm", "conv_1_bn"))
    x = g.node("GlobalAveragePool", [x], "gap")                                                               # :54
    x = _conv(g, x, weights["StylePredictor/kernel"], weights["StylePredictor/bias"], (1, 1), 1, "StylePredictor")        # :59-63
    x = _conv(g, x, weights["StyleNormPredictor/kernel"], weights["StyleNormPredictor/bias"], (1, 1), 1, "StyleNormPredictor")   # :66-70
    g.node("Flatten", [x], "style_params", axis=1)
    g.rename_last_output("style_params")
    g.outputs.append(_value_info("style_params", ["N", plan.num_top_parameters]))
    return g.serialize()


def export_onnx(model, path) -> str:
    """``model``: the object returned by create_style_transfer_model / create_style_prediction_model (or their ``.transfer`` /
    ``.style_predictor`` attributes of the combined models).  Writes ``path`` and returns it."""
    path = str(path)
    plan = model.plan
    if isinstance(plan, TransferPlan):
        data = transfer_onnx_bytes(plan, model.weights)
    elif isinstance(plan, PredictorPlan):
        data = predictor_onnx_bytes(plan, model.weights)
    else:
        raise TypeError(f"cannot export {type(model).__name__}")
    with open(path, "wb") as f:
        f.write(data)
    return path
