"""Training loss -- host mirror of realtime_style_transfer/models/styleLoss.py.

``StyleLossModelVGG(input_shape)`` + ``make_style_loss_function(loss_model, output_shape, num_styles, with_depth_loss)``
keep the reference's names and return ``(compute_loss, model)`` where ``compute_loss(y_pred, y_true)`` yields the dict
``{'loss','feature_loss','style_loss','total_variation_loss'}`` of per-sample ``(B,)`` vectors (styleLoss.py:344-352).
The arithmetic (VGG16 features, Gram matrices, reductions, and the backward w.r.t. the prediction) runs in
librst_sm100.so (csrc/loss.cu).  The MiDaS depth loss needs a network download from tfhub.dev and is out of scope.
"""
from __future__ import annotations

import logging
import math

import numpy as np

from .. import _native
from ._base import as_numpy

log = logging.getLogger(__name__)

_VGG16 = [("block1", 2, 64), ("block2", 2, 128), ("block3", 3, 256), ("block4", 3, 512), ("block5", 3, 512)]


def gram_matrix(input_tensor):
    """einsum('bijc,bijd->bcd') / (H*W)  (styleLoss.py:21-37), on the GPU."""
    import torch
    x = as_numpy(input_tensor)
    b, h, w, c = x.shape
    dev = torch.device("cuda", torch.cuda.current_device())
    d_x = torch.from_numpy(x).to(dev)
    d_g = torch.empty((b, c, c), dtype=torch.float32, device=dev)
    _native.op_gram(d_x.data_ptr(), d_g.data_ptr(), b, h, w, c, torch.cuda.current_stream().cuda_stream)
    return d_g.cpu().numpy()


def mean_l2_loss_on_batch(tensor):
    """styleLoss.py:290-292 (host helper for small tensors)."""
    t = as_numpy(tensor)
    return (0.5 * t * t).reshape(t.shape[0], -1).mean(axis=1)


class StyleLossModelBase:
    def __init__(self, name):
        self.name = name
        self.trainable = False
        self.content_layers = None
        self.style_layers = None
        self.content_loss_factor = 1
        self.style_loss_factor = 1
        self.total_variation_loss_factor = 1
        self.depth_loss_factor = 1


class StyleLossModelVGG(StyleLossModelBase):
    """VGG16 feature extractor configuration and (frozen) weights (styleLoss.py:69-109)."""

    def __init__(self, input_shape, seed=None):
        super().__init__('StyleLossModelVGG')
        self.input_shape = tuple(input_shape)
        self.style_layers = ['block1_conv2', 'block2_conv2', 'block3_conv3', 'block4_conv3']
        self.content_layers = ['block5_conv3']
        self.num_style_layers = len(self.style_layers)
        self.content_loss_factor = 1e4
        self.style_loss_factor = 1e-3
        self.total_variation_loss_factor = 1e-1
        self.depth_loss_factor = 1e-2
        log.warning("VGG16 is randomly initialised: the ImageNet weights Keras would download are not reachable here; "
                    "assign real weights with set_weights().")
        rng = np.random.default_rng(seed)
        self._variables = {}
        cin = 3
        for blk, n, co in _VGG16:
            for i in range(1, n + 1):
                self._variables[f"{blk}_conv{i}/kernel"] = rng.normal(0, math.sqrt(2.0 / (9 * cin)), (3, 3, cin, co)).astype(np.float32)
                self._variables[f"{blk}_conv{i}/bias"] = np.zeros(co, np.float32)
                cin = co
        self._native = None
        self._dirty = True
        # arithmetic of the VGG convolutions on the tensor cores: PRECISION_FP32 = split tf32 (fp32-level accuracy, default),
        # PRECISION_TF32 = plain tf32 operands (TensorFlow's own default on Ampere and later; faster, losses within 1e-3)
        self.math = _native.PRECISION_FP32

    @property
    def weights(self):
        return dict(self._variables)

    def set_weights(self, weights):
        for k, v in weights.items():
            if k not in self._variables or tuple(np.shape(v)) != self._variables[k].shape:
                raise ValueError(f"bad VGG16 variable {k!r} with shape {np.shape(v)}")
            self._variables[k] = np.ascontiguousarray(v, np.float32)
        self._dirty = True
        self._version = getattr(self, "_version", 0) + 1

    def native(self, batch: int) -> "_native.NativeLoss":
        if self._native is None or self._native.max_batch < batch:
            if self._native is not None:
                self._native.close()
            self._native = _native.NativeLoss(self.input_shape[0], self.input_shape[1], batch)
            self._dirty = True
        if self._dirty or getattr(self, "_native_math", None) != self.math:
            self._native.set_math(self.math)
            self._native_math = self.math
            self._native.set_weights(self._variables)
            self._native.set_factors(self.content_loss_factor, self.style_loss_factor, self.total_variation_loss_factor)
            self._dirty = False
        return self._native


def make_style_loss_function(loss_feature_extractor_model, output_shape, num_styles, with_depth_loss=True):
    assert num_styles == 1, f"Loss model does not support multiple styles. Found {num_styles}"
    if with_depth_loss:
        raise NotImplementedError("depth loss needs the MiDaS network from tfhub.dev (styleLoss.py:254); out of scope")
    if not isinstance(loss_feature_extractor_model, StyleLossModelVGG):
        raise NotImplementedError("only StyleLossModelVGG (the extractor train_network.py:85 uses) is built")
    model = loss_feature_extractor_model
    model.trainable = False

    def compute_loss(y_pred, y_true):
        """y_pred (B,H,W,3); y_true {'content': (B,H,W,3), 'style': (B,1,H,W,3)} -> dict of (B,) float32 arrays."""
        import torch
        pred, content, style = as_numpy(y_pred), as_numpy(y_true['content']), as_numpy(y_true['style'])
        if style.ndim == 5:
            assert style.shape[1] == 1, f"Loss model does not support multiple styles. Found {style.shape[1]}"
            style = style[:, 0]
        b = pred.shape[0]
        dev = torch.device("cuda", torch.cuda.current_device())
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (pred, content, style)]
        out = torch.empty((b, 4), dtype=torch.float32, device=dev)
        model.native(b).forward(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), out.data_ptr(), b,
                                torch.cuda.current_stream().cuda_stream)
        o = out.cpu().numpy()
        return {"loss": o[:, 0], "feature_loss": o[:, 1], "style_loss": o[:, 2], "total_variation_loss": o[:, 3]}

    return compute_loss, model
