"""Style predictor -- host mirror of realtime_style_transfer/models/stylePrediction.py.

``create_style_prediction_model(input_shape, feature_extractor, num_top_parameters)`` returns a model
mapping a style image (B,H,W,3) in [0,1] to ``num_top_parameters`` style parameters
(stylePrediction.py:25-75): Rescaling(2,-1) -> MobileNetV3Small (no top) -> global average pool ->
1x1 conv to 100 -> 1x1 conv to P.  DUMMY replaces the backbone by one Conv2D(1, 9, 5).
"""
from __future__ import annotations

import logging

import numpy as np

from .. import _native
from .._plan import PredictorPlan
from ._base import InputSpec, NativeModel, _is_torch_cuda, as_numpy

log = logging.getLogger(__name__)

DENSE_KERNEL_INITIALIZER = {
    "class_name": "VarianceScaling",
    "config": {"scale": 1. / 3., "mode": "fan_out", "distribution": "uniform"},
}


class StyleFeatureExtractor:
    DUMMY = 'DUMMY'
    EFFICIENT_NET = 'EFFICIENT_NET'
    MOBILE_NET = 'MOBILE_NET'


_EXTRACTOR_CODE = {StyleFeatureExtractor.DUMMY: _native.EXTRACTOR_DUMMY,
                   StyleFeatureExtractor.MOBILE_NET: _native.EXTRACTOR_MOBILE_NET}


class StylePredictionModel(NativeModel):
    def __init__(self, plan: PredictorPlan, name: str, seed=None):
        super().__init__(name)
        self.plan = plan
        self._variables = plan.initial_weights(np.random.default_rng(seed))
        self.input = InputSpec(plan.input_shape, "style")
        self.inputs = self.input
        self.output_shape = (None, plan.num_top_parameters)

    def _context_kwargs(self, max_batch):
        return dict(extractor=_EXTRACTOR_CODE[self.plan.feature_extractor], style_shape=self.plan.input_shape,
                    predictor_num_params=self.plan.num_top_parameters)

    def _precision(self):
        return _native.PRECISION_FP32      # the predictor always runs the fp32 kernels

    def __call__(self, style_image, training=False):
        if tuple(style_image.shape[1:]) != self.plan.input_shape:
            raise ValueError(f"style image shape {tuple(style_image.shape)} incompatible with {self.input.shape}")
        if _is_torch_cuda(style_image):
            import torch
            b = style_image.shape[0]
            ctx = self._get_ctx(b)
            x = style_image.contiguous().float()
            out = torch.empty((b, self.plan.num_top_parameters), dtype=torch.float32, device=x.device)
            ctx.predict_style_device(x.data_ptr(), out.data_ptr(), b, torch.cuda.current_stream(x.device).cuda_stream)
            return out
        return self.predict(style_image)

    def predict(self, x, batch_size=None, verbose=0, **kwargs):
        x = as_numpy(x)
        if tuple(x.shape[1:]) != self.plan.input_shape:
            raise ValueError(f"style image shape {tuple(x.shape)} incompatible with {self.input.shape}")
        n = x.shape[0]
        step = n if not batch_size else int(batch_size)
        ctx = self._get_ctx(min(step, n) if n else 1)
        outs = [ctx.predict_style_host(x[i:i + step]) for i in range(0, n, max(step, 1))]
        if not outs:
            return np.zeros((0, self.plan.num_top_parameters), np.float32)
        return np.concatenate(outs, axis=0) if len(outs) > 1 else outs[0]


def create_style_prediction_model(input_shape, feature_extractor: StyleFeatureExtractor, num_top_parameters,
                                  num_style_parameters=100, name="StylePredictionModel"):
    if feature_extractor == StyleFeatureExtractor.EFFICIENT_NET:
        raise NotImplementedError("EFFICIENT_NET style extractor is outside the accelerated path "
                                  "(ShapeConfig never selects it, shape_config.py:29)")
    plan = PredictorPlan(input_shape, feature_extractor, num_top_parameters, num_style_parameters)
    if feature_extractor == StyleFeatureExtractor.MOBILE_NET:
        log.warning("MobileNetV3Small is randomly initialised: the ImageNet weights Keras would download are "
                    "not reachable here; load a checkpoint for meaningful style parameters.")
    log.info(f"Bottlenecking to {num_style_parameters} parameters for {num_top_parameters} norm parameters")
    return StylePredictionModel(plan, name)
