"""Independent pure-numpy loop restatement of the TF op semantics the oracle relies on.

TEST INFRASTRUCTURE (see oracle/rst_oracle.py header).  These are written straight from the
TensorFlow op definitions (explicit index arithmetic, float64 accumulation, no library conv)
so that oracle/rst_oracle.py -- which leans on torch.nn.functional -- is checked by a second,
structurally different implementation on small cases.  parity unpinned vs. TensorFlow itself
(TF is not installable here); pinned only where the reference has its own known-answer test.
"""
from __future__ import annotations

import numpy as np


def same_pad(size, k, s):
    out = -(-size // s)
    total = max((out - 1) * s + k - size, 0)
    return out, total // 2


def conv2d_same(x, kernel, bias, stride):
    """x (B,H,W,Ci), kernel (kh,kw,Ci,Co): out[b,oy,ox,co] = sum x[b, oy*s-pt+ky, ox*s-pl+kx, ci]*K."""
    b, h, w, ci = x.shape
    kh, kw, _, co = kernel.shape
    ho, pt = same_pad(h, kh, stride)
    wo, pl = same_pad(w, kw, stride)
    out = np.zeros((b, ho, wo, co), np.float64)
    for oy in range(ho):
        for ox in range(wo):
            for ky in range(kh):
                iy = oy * stride - pt + ky
                if iy < 0 or iy >= h:
                    continue
                for kx in range(kw):
                    ix = ox * stride - pl + kx
                    if ix < 0 or ix >= w:
                        continue
                    out[:, oy, ox, :] += x[:, iy, ix, :].astype(np.float64) @ kernel[ky, kx].astype(np.float64)
    if bias is not None:
        out += bias
    return out


def conv2d_transpose_same(x, kernel, bias, stride):
    """Scatter form of the input-gradient of the SAME forward conv.  kernel (kh,kw,Co,Ci).
    Every input pixel (i,j) adds x[i,j,:] @ K[ky,kx].T to full-output pixel (s*i+ky, s*j+kx);
    the result is cropped by the forward conv's pad_before to size (s*H, s*W)."""
    b, h, w, ci = x.shape
    kh, kw, co, _ = kernel.shape
    ho, wo = h * stride, w * stride
    _, pt = same_pad(ho, kh, stride)
    _, pl = same_pad(wo, kw, stride)
    fh, fw = (h - 1) * stride + kh, (w - 1) * stride + kw
    full = np.zeros((b, fh, fw, co), np.float64)
    for i in range(h):
        for j in range(w):
            for ky in range(kh):
                for kx in range(kw):
                    full[:, i * stride + ky, j * stride + kx, :] += \
                        x[:, i, j, :].astype(np.float64) @ kernel[ky, kx].astype(np.float64).T
    out = np.zeros((b, ho, wo, co), np.float64)
    yy = min(ho, fh - pt)
    xx = min(wo, fw - pl)
    out[:, :yy, :xx] = full[:, pt:pt + yy, pl:pl + xx]
    if bias is not None:
        out += bias
    return out


def cin(x, scale, bias, eps=1e-5):
    x = x.astype(np.float64)
    mean = x.mean(axis=(1, 2), keepdims=True)
    var = ((x - mean) ** 2).mean(axis=(1, 2), keepdims=True)
    inv = 1.0 / np.sqrt(var + eps)
    return bias + (x * inv - mean * inv) * scale


def apply_style_weights_loop(style_weights, style_params):
    """The explicit 4-deep loop of the reference's known-answer test
    (realtime_style_transfer/models/styleTransferTest.py:41-47)."""
    b, h, w, _ = style_weights.shape
    f = style_params.shape[-1]
    out = np.zeros((b, h, w, f))
    for bi in range(b):
        for x in range(h):
            for y in range(w):
                for c in range(f):
                    out[bi, x, y, c] = style_weights[bi, x, y, 0] * style_params[bi, 0, 0, c] + \
                                       style_weights[bi, x, y, 1] * style_params[bi, 0, 1, c]
    return out


def vertical_gradient(min_max, shape):
    """_generate_vertical_gradient_tensor of the reference test (styleTransferTest.py:12-24):
    note it divides the row index by shape[0] (the batch), so values leave [0,1]."""
    rows = []
    for _b in range(shape[0]):
        for i in range(shape[1]):
            rows.append([min_max[0] + (i / shape[0]) * (min_max[1] - min_max[0]) for _j in range(shape[2])])
    return np.asarray(rows, np.float32).reshape(shape)


def gram(x):
    b, h, w, c = x.shape
    f = x.reshape(b, h * w, c).astype(np.float64)
    return np.einsum("bpc,bpd->bcd", f, f) / (h * w)
