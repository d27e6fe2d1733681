"""Loss surface -- mirror of realtime_style_transfer/models/styleLoss.py (names only for now).

SURVEY.md section 8 rows a12-a14 (VGG16 features, Gram matrices, content/style/TV loss).  The Gram
matrix operator is available natively (rst_op_gram); the full loss model is the next row to build.
"""
from __future__ import annotations

import numpy as np

from .. import _native
from ._base import as_numpy


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


class StyleLossModelVGG:
    """Configuration holder for the VGG16 loss model (styleLoss.py:69-109)."""
    style_layers = ['block1_conv2', 'block2_conv2', 'block3_conv3', 'block4_conv3']
    content_layers = ['block5_conv3']

    def __init__(self, input_shape):
        self.name = 'StyleLossModelVGG'
        self.input_shape = tuple(input_shape)
        self.trainable = False
        self.num_style_layers = len(self.style_layers)
        self.content_loss_factor = 1e4
        self.style_loss_factor = 1e-3
        self.total_variation_loss_factor = 1e-1
        self.depth_loss_factor = 1e-2


def make_style_loss_function(loss_feature_extractor_model, output_shape, num_styles, with_depth_loss=True):
    assert num_styles == 1, f"Loss model does not support multiple styles. Found {num_styles}"
    if with_depth_loss:
        raise NotImplementedError("depth loss needs the MiDaS network from tfhub.dev (styleLoss.py:254); out of scope")

    def compute_loss(y_pred, y_true):
        raise NotImplementedError("native VGG/Gram loss forward is not built yet (SURVEY.md section 8 a12-a14)")

    return compute_loss, loss_feature_extractor_model
