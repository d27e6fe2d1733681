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


# ------------------------------------------------------------------------------------------------------------------
# Second restatement of the pieces that had only ONE (oracle/rst_oracle.py): Keras MobileNetV3Small's building blocks,
# training-mode BatchNormalization and RMSprop.  Written from the published Keras 2.9 sources
# (keras/applications/mobilenet_v3.py, keras/applications/imagenet_utils.py::correct_pad,
# keras/layers/normalization/batch_normalization.py, keras/optimizers/optimizer_v2/rmsprop.py) with explicit loops and
# float64 arithmetic; no torch, no shared helper with rst_oracle.  parity unpinned vs. TensorFlow itself.
# ------------------------------------------------------------------------------------------------------------------
def correct_pad(size, kernel_size):
    """imagenet_utils.correct_pad for one spatial dim: ((k//2) - adjust, k//2) with adjust = 1 - size % 2."""
    adjust = 1 - size % 2
    correct = kernel_size // 2
    return correct - adjust, correct


def zero_pad(x, top, bottom, left, right):
    b, h, w, c = x.shape
    out = np.zeros((b, h + top + bottom, w + left + right, c), np.float64)
    out[:, top:top + h, left:left + w] = x
    return out


def depthwise_valid(x, kernel, stride):
    """DepthwiseConv2D(padding='valid'), kernel (kh,kw,C,1): out[b,oy,ox,c] = sum_k x[b,oy*s+ky,ox*s+kx,c] * K[ky,kx,c]."""
    b, h, w, c = x.shape
    kh, kw = kernel.shape[:2]
    ho, wo = (h - kh) // stride + 1, (w - kw) // stride + 1
    out = np.zeros((b, ho, wo, c), np.float64)
    for oy in range(ho):
        for ox in range(wo):
            for ky in range(kh):
                for kx in range(kw):
                    out[:, oy, ox, :] += x[:, oy * stride + ky, ox * stride + kx, :] * kernel[ky, kx, :, 0]
    return out


def depthwise_same(x, kernel, stride):
    """DepthwiseConv2D(padding='same'): TF SAME padding, then the valid loop."""
    kh, kw = kernel.shape[:2]
    _, pt = same_pad(x.shape[1], kh, stride)
    _, pl = same_pad(x.shape[2], kw, stride)
    ho, wo = -(-x.shape[1] // stride), -(-x.shape[2] // stride)
    total_h = max((ho - 1) * stride + kh - x.shape[1], 0)
    total_w = max((wo - 1) * stride + kw - x.shape[2], 0)
    return depthwise_valid(zero_pad(x, pt, total_h - pt, pl, total_w - pl), kernel, stride)


def batchnorm_inference(x, gamma, beta, moving_mean, moving_var, eps=1e-3):
    return (x - moving_mean) / np.sqrt(moving_var + eps) * gamma + beta


def batchnorm_training(x, gamma, beta, moving_mean, moving_var, eps=1e-3, momentum=0.99):
    """Keras BatchNormalization(training=True) on NHWC input: normalise with the batch mean and the POPULATION variance over
    (N,H,W); the moving variance is updated with the Bessel-corrected (n/(n-1)) batch variance, as the fused kernel does.
    Returns (y, new_moving_mean, new_moving_var)."""
    n = x.shape[0] * x.shape[1] * x.shape[2]
    flat = x.reshape(n, x.shape[3]).astype(np.float64)
    mean = flat.sum(axis=0) / n
    var = ((flat - mean) ** 2).sum(axis=0) / n
    y = (x - mean) / np.sqrt(var + eps) * gamma + beta
    unbiased = var * n / max(n - 1, 1)
    return y, moving_mean * momentum + mean * (1 - momentum), moving_var * momentum + unbiased * (1 - momentum)


def relu(x):
    return np.maximum(x, 0.0)


def hard_sigmoid(x):
    """mobilenet_v3.hard_sigmoid: ReLU(6)(x + 3) / 6."""
    return np.minimum(np.maximum(x + 3.0, 0.0), 6.0) / 6.0


def hard_swish(x):
    return x * hard_sigmoid(x)


def make_divisible(v, divisor=8):
    """mobilenet_v3._depth."""
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def inverted_res_block(x, W, prefix, expansion, filters, kernel_size, stride, se_ratio, activation, block_id):
    """mobilenet_v3._inverted_res_block (inference mode).  W: name -> array with this module's variable names
    (``<prefix>/expand/kernel`` ...).  Returns the block output."""
    act = relu if activation == "relu" else hard_swish
    shortcut = x
    infilters = x.shape[3]
    cexp = make_divisible(infilters * expansion)

    def bn(name, t):
        return batchnorm_inference(t, W[f"{prefix}/{name}/BatchNorm/gamma"], W[f"{prefix}/{name}/BatchNorm/beta"],
                                   W[f"{prefix}/{name}/BatchNorm/moving_mean"], W[f"{prefix}/{name}/BatchNorm/moving_variance"])

    if block_id:
        x = act(bn("expand", conv2d_same(x, W[f"{prefix}/expand/kernel"], None, 1)))
    if stride == 2:
        pt, pb = correct_pad(x.shape[1], kernel_size)
        pl, pr = correct_pad(x.shape[2], kernel_size)
        x = depthwise_valid(zero_pad(x, pt, pb, pl, pr), W[f"{prefix}/depthwise/depthwise_kernel"], 2)
    else:
        x = depthwise_same(x, W[f"{prefix}/depthwise/depthwise_kernel"], 1)
    x = act(bn("depthwise", x))
    if se_ratio:
        z = x.mean(axis=(1, 2), keepdims=True)
        z = relu(conv2d_same(z, W[f"{prefix}/squeeze_excite/Conv/kernel"], W[f"{prefix}/squeeze_excite/Conv/bias"], 1))
        z = hard_sigmoid(conv2d_same(z, W[f"{prefix}/squeeze_excite/Conv_1/kernel"], W[f"{prefix}/squeeze_excite/Conv_1/bias"], 1))
        x = x * z
    x = bn("project", conv2d_same(x, W[f"{prefix}/project/kernel"], None, 1))
    if stride == 1 and infilters == filters:
        x = shortcut + x
    return x


# MobileNetV3Small(alpha=1.0) stack_fn: (expansion, filters, kernel, stride, se_ratio, activation)
MOBILENET_V3_SMALL = [
    (1, 16, 3, 2, 0.25, "relu"), (72. / 16, 24, 3, 2, None, "relu"), (88. / 24, 24, 3, 1, None, "relu"),
    (4, 40, 5, 2, 0.25, "hard_swish"), (6, 40, 5, 1, 0.25, "hard_swish"), (6, 40, 5, 1, 0.25, "hard_swish"),
    (3, 48, 5, 1, 0.25, "hard_swish"), (3, 48, 5, 1, 0.25, "hard_swish"), (6, 96, 5, 2, 0.25, "hard_swish"),
    (6, 96, 5, 1, 0.25, "hard_swish"), (6, 96, 5, 1, 0.25, "hard_swish"),
]


def mobilenet_v3_small(W, x):
    """MobileNetV3Small(include_top=False, include_preprocessing=False), inference mode; x already in [-1, 1].
    Variable names carry the ``mobilenet/`` prefix of the oracle's registry."""
    p = "mobilenet"
    x = conv2d_same(x, W[f"{p}/Conv/kernel"], None, 2)
    x = hard_swish(batchnorm_inference(x, W[f"{p}/Conv/BatchNorm/gamma"], W[f"{p}/Conv/BatchNorm/beta"],
                                       W[f"{p}/Conv/BatchNorm/moving_mean"], W[f"{p}/Conv/BatchNorm/moving_variance"]))
    for bid, (e, f, k, s, se, act) in enumerate(MOBILENET_V3_SMALL):
        x = inverted_res_block(x, W, f"{p}/expanded_conv" + (f"_{bid}" if bid else ""), e, f, k, s, se, act, bid)
    last = make_divisible(x.shape[3] * 6)
    x = conv2d_same(x, W[f"{p}/Conv_1/kernel"], None, 1)
    assert x.shape[3] == last
    return hard_swish(batchnorm_inference(x, W[f"{p}/Conv_1/BatchNorm/gamma"], W[f"{p}/Conv_1/BatchNorm/beta"],
                                          W[f"{p}/Conv_1/BatchNorm/moving_mean"], W[f"{p}/Conv_1/BatchNorm/moving_variance"]))


def style_predictor(W, style01):
    """create_style_prediction_model(MOBILE_NET) forward (stylePrediction.py:33-37, :54-72): Rescaling(2, -1), backbone,
    global average pool, two linear 1x1 convolutions."""
    x = mobilenet_v3_small(W, style01.astype(np.float64) * 2.0 - 1.0)
    x = x.mean(axis=(1, 2), keepdims=True)
    x = conv2d_same(x, W["StylePredictor/kernel"], W["StylePredictor/bias"], 1)
    x = conv2d_same(x, W["StyleNormPredictor/kernel"], W["StyleNormPredictor/bias"], 1)
    return x[:, 0, 0, :]


def rmsprop_step(var, grad, rms, learning_rate=1e-3, rho=0.9, epsilon=1e-7):
    """One dense update of keras RMSprop with momentum 0, centered False, element by element:
    rms_t = rho * rms + (1 - rho) * g^2;  var_t = var - lr * g / (sqrt(rms_t) + epsilon)."""
    var, grad, rms = (np.asarray(a, np.float64).copy() for a in (var, grad, rms))
    flat_v, flat_g, flat_r = var.reshape(-1), grad.reshape(-1), rms.reshape(-1)
    for i in range(flat_v.size):
        flat_r[i] = rho * flat_r[i] + (1.0 - rho) * flat_g[i] * flat_g[i]
        flat_v[i] = flat_v[i] - learning_rate * flat_g[i] / (np.sqrt(flat_r[i]) + epsilon)
    return var, rms


def max_pool2(x):
    b, h, w, c = x.shape
    out = np.full((b, h // 2, w // 2, c), -np.inf)
    for dy in range(2):
        for dx in range(2):
            out = np.maximum(out, x[:, dy:2 * (h // 2):2, dx:2 * (w // 2):2, :])
    return out


def total_variation(img):
    """tf.image.total_variation: sum |x[:,1:]-x[:,:-1]| + sum |x[:,:,1:]-x[:,:,:-1]| per image."""
    img = img.astype(np.float64)
    return np.abs(img[:, 1:] - img[:, :-1]).sum(axis=(1, 2, 3)) + np.abs(img[:, :, 1:] - img[:, :, :-1]).sum(axis=(1, 2, 3))
