"""A small ONNX reader + CPU interpreter -- TEST INFRASTRUCTURE (see oracle/rst_oracle.py header).

Neither ``onnx`` nor ``onnxruntime`` exist in this environment, so the files written by
``realtime_style_transfer_b200/export.py`` are checked by decoding the protobuf wire format here (independently of the
writer: nothing is imported from the package) and executing the graph with the operator semantics of the ONNX specification
(opset 13) on torch-CPU float64 tensors.  Covers exactly the operators the two exported graphs use.
"""
from __future__ import annotations

import struct
from typing import Dict, List

import numpy as np
import torch
import torch.nn.functional as F


def _fields(buf: bytes):
    pos, n = 0, len(buf)
    while pos < n:
        tag, shift = 0, 0
        while True:
            b = buf[pos]; pos += 1
            tag |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        fn, wt = tag >> 3, tag & 7
        if wt == 0:
            v, shift = 0, 0
            while True:
                b = buf[pos]; pos += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            if v >= 1 << 63:
                v -= 1 << 64
        elif wt == 2:
            ln, shift = 0, 0
            while True:
                b = buf[pos]; pos += 1
                ln |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = struct.unpack_from("<f", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<d", buf, pos)[0]
            pos += 8
        else:
            raise ValueError(f"wire type {wt}")
        yield fn, v


def _tensor(buf: bytes):
    dims, dtype, name, raw = [], 1, "", b""
    for fn, v in _fields(buf):
        if fn == 1:
            dims.append(v)
        elif fn == 2:
            dtype = v
        elif fn == 8:
            name = v.decode()
        elif fn == 9:
            raw = v
    np_dtype = {1: "<f4", 7: "<i8"}[dtype]
    return name, np.frombuffer(raw, dtype=np_dtype).reshape(dims).copy()


def _attribute(buf: bytes):
    name, out, floats, ints, typ = "", None, [], [], 0
    for fn, v in _fields(buf):
        if fn == 1:
            name = v.decode()
        elif fn == 2:
            out = float(v)
        elif fn == 3:
            out = int(v)
        elif fn == 4:
            out = v.decode()
        elif fn == 7:
            floats.append(float(v))
        elif fn == 8:
            ints.append(int(v))
        elif fn == 20:
            typ = v
    if typ == 7:
        out = ints
    elif typ == 6:
        out = floats
    return name, out


def _value_info(buf: bytes):
    name, shape = "", []
    for fn, v in _fields(buf):
        if fn == 1:
            name = v.decode()
        elif fn == 2:
            for f2, t in _fields(v):
                if f2 == 1:                                   # tensor_type
                    for f3, s in _fields(t):
                        if f3 == 2:                           # shape
                            for f4, d in _fields(s):
                                if f4 == 1:
                                    for f5, dv in _fields(d):
                                        shape.append(dv if f5 == 1 else dv.decode())
    return name, shape


class Model:
    def __init__(self, data: bytes):
        self.ir_version, self.opset, self.producer = None, None, ""
        self.nodes: List[dict] = []
        self.initializers: Dict[str, np.ndarray] = {}
        self.inputs, self.outputs = [], []
        for fn, v in _fields(data):
            if fn == 1:
                self.ir_version = v
            elif fn == 2:
                self.producer = v.decode()
            elif fn == 8:
                for f2, x in _fields(v):
                    if f2 == 2:
                        self.opset = x
            elif fn == 7:
                for f2, x in _fields(v):
                    if f2 == 1:
                        node = {"inputs": [], "outputs": [], "op": "", "attrs": {}}
                        for f3, y in _fields(x):
                            if f3 == 1:
                                node["inputs"].append(y.decode())
                            elif f3 == 2:
                                node["outputs"].append(y.decode())
                            elif f3 == 4:
                                node["op"] = y.decode()
                            elif f3 == 5:
                                k, val = _attribute(y)
                                node["attrs"][k] = val
                        self.nodes.append(node)
                    elif f2 == 5:
                        name, arr = _tensor(x)
                        self.initializers[name] = arr
                    elif f2 == 11:
                        self.inputs.append(_value_info(x))
                    elif f2 == 12:
                        self.outputs.append(_value_info(x))

    # ---- operator semantics (ONNX opset 13) --------------------------------------------------------------------------------
    def run(self, feeds: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
        env = {k: torch.as_tensor(v) for k, v in self.initializers.items()}
        env = {k: (v.double() if v.dtype == torch.float32 else v) for k, v in env.items()}
        for name, _shape in self.inputs:
            env[name] = torch.as_tensor(np.asarray(feeds[name])).double()
        for n in self.nodes:
            x = [env[i] for i in n["inputs"]]
            a = n["attrs"]
            op = n["op"]
            if op == "Transpose":
                y = x[0].permute(*a["perm"])
            elif op == "Conv":
                pads = a.get("pads", [0, 0, 0, 0])                       # [top, left, bottom, right]
                xin = F.pad(x[0], (pads[1], pads[3], pads[0], pads[2]))
                y = F.conv2d(xin, x[1], x[2] if len(x) > 2 else None, stride=tuple(a.get("strides", [1, 1])), groups=a.get("group", 1))
            elif op == "ConvTranspose":
                pads, st = a.get("pads", [0, 0, 0, 0]), a.get("strides", [1, 1])
                full = F.conv_transpose2d(x[0], x[1], x[2] if len(x) > 2 else None, stride=tuple(st))
                y = full[:, :, pads[0]:full.shape[2] - pads[2], pads[1]:full.shape[3] - pads[3]]
            elif op == "Relu":
                y = torch.relu(x[0])
            elif op == "Sigmoid":
                y = torch.sigmoid(x[0])
            elif op == "HardSigmoid":
                y = torch.clamp(a.get("alpha", 0.2) * x[0] + a.get("beta", 0.5), 0.0, 1.0)
            elif op == "BatchNormalization":
                sh = (1, -1, 1, 1)
                y = (x[0] - x[3].view(sh)) / torch.sqrt(x[4].view(sh) + a.get("epsilon", 1e-5)) * x[1].view(sh) + x[2].view(sh)
            elif op == "InstanceNormalization":
                mean = x[0].mean(dim=(2, 3), keepdim=True)
                var = x[0].var(dim=(2, 3), unbiased=False, keepdim=True)
                y = x[1].view(1, -1, 1, 1) * (x[0] - mean) / torch.sqrt(var + a.get("epsilon", 1e-5)) + x[2].view(1, -1, 1, 1)
            elif op == "Mul":
                y = x[0] * x[1]
            elif op == "Add":
                y = x[0] + x[1]
            elif op == "Sub":
                y = x[0] - x[1]
            elif op == "Slice":
                y = x[0]
                starts, ends, axes = x[1].tolist(), x[2].tolist(), x[3].tolist()
                for s0, e0, ax in zip(starts, ends, axes):
                    y = y.narrow(ax, s0, min(e0, y.shape[ax]) - s0)
            elif op == "Reshape":
                shape = [int(x[0].shape[i]) if s == 0 else int(s) for i, s in enumerate(x[1].tolist())]
                y = x[0].reshape(shape)
            elif op == "AveragePool":
                y = F.avg_pool2d(x[0], tuple(a["kernel_shape"]), tuple(a.get("strides", a["kernel_shape"])))
            elif op == "GlobalAveragePool":
                y = x[0].mean(dim=(2, 3), keepdim=True)
            elif op == "Flatten":
                ax = a.get("axis", 1)
                y = x[0].reshape(int(np.prod(x[0].shape[:ax])) if ax else 1, -1)
            else:
                raise NotImplementedError(f"ONNX operator {op}")
            env[n["outputs"][0]] = y
        return {name: env[name].numpy() for name, _ in self.outputs}


def load(path) -> Model:
    with open(path, "rb") as f:
        return Model(f.read())
