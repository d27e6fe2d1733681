"""Predictor + transfer network glue -- mirror of
realtime_style_transfer/models/styleTransferInferenceModel.py:9-48.

The factory keeps the reference's signature and returns an object with ``.inputs .inference .transfer
.style_predictor``.  ``inference.predict`` runs predictor-per-style then the transfer net inside ONE
native context (one H2D of the inputs, one D2H of the stylised frames).
"""
from __future__ import annotations

import logging
import typing

import numpy as np

from .. import _native
from ._base import InputSpec, NativeModel, as_numpy
from .stylePrediction import _EXTRACTOR_CODE

log = logging.getLogger(__name__)


class StyleTransferInferenceModel(NativeModel):
    def __init__(self, num_styles, style_predictor, style_transfer_model, name):
        super().__init__(name)
        self.num_styles = num_styles
        self.style_predictor = style_predictor
        self.transfer = style_transfer_model
        self.input = {
            "content": style_transfer_model.input["content"],
            "style": InputSpec((num_styles,) + tuple(style_predictor.input.shape[1:]), "style"),
        }
        if "style_weights" in style_transfer_model.input:
            self.input["style_weights"] = style_transfer_model.input["style_weights"]
        self.inputs = self.input
        self.output_shape = style_transfer_model.output_shape

    # variables live in the two sub-models; this model only aggregates them
    def _all_variables(self):
        merged = dict(self.transfer._variables)
        merged.update(self.style_predictor._variables)
        return merged

    @property
    def weights(self):
        return self._all_variables()

    def count_params(self):
        return self.transfer.count_params() + self.style_predictor.count_params()

    def get_weights(self):
        return [v.copy() for v in self._all_variables().values()]

    def set_weights(self, weights):
        if not isinstance(weights, dict):
            weights = dict(zip(self._all_variables().keys(), list(weights)))
        t = {k: v for k, v in weights.items() if k in self.transfer._variables}
        p = {k: v for k, v in weights.items() if k in self.style_predictor._variables}
        unknown = [k for k in weights if k not in t and k not in p]
        if unknown:
            raise ValueError(f"unknown variables {unknown[:5]}")
        if t:
            self.transfer.set_weights(t)
        if p:
            self.style_predictor.set_weights(p)

    def _context_kwargs(self, max_batch):
        kw = self.transfer._context_kwargs(max_batch)
        kw.update(extractor=_EXTRACTOR_CODE[self.style_predictor.plan.feature_extractor],
                  style_shape=self.style_predictor.plan.input_shape)
        return kw

    def _weights_version(self):
        # the variables live in the sub-models and may be assigned there directly; each context (this model's and the
        # sub-models' own) remembers the versions IT uploaded, so none of them consumes a flag another one needs
        return (self.transfer._version, self.style_predictor._version)

    def __call__(self, inputs, training=False):
        return self.predict(inputs)

    def predict(self, x, batch_size=None, verbose=0, **kwargs):
        content = as_numpy(x["content"])
        style = as_numpy(x["style"])
        weights = as_numpy(x["style_weights"]) if "style_weights" in self.input else None
        if tuple(content.shape[1:]) != tuple(self.input["content"].shape[1:]):
            raise ValueError(f"content shape {content.shape} incompatible with {self.input['content'].shape}")
        if tuple(style.shape[1:]) != tuple(self.input["style"].shape[1:]):
            raise ValueError(f"style shape {style.shape} incompatible with {self.input['style'].shape}")
        n = content.shape[0]
        step = n if not batch_size else int(batch_size)
        ctx = self._get_ctx(min(step, n) if n else 1)
        outs = []
        for i in range(0, n, max(step, 1)):
            outs.append(ctx.inference_forward_host(content[i:i + step], style[i:i + step],
                                                   weights[i:i + step] if weights is not None else None))
        if not outs:
            return np.zeros((0,) + tuple(self.output_shape[1:]), np.float32)
        return np.concatenate(outs, axis=0) if len(outs) > 1 else outs[0]


def make_style_transfer_inference_model(num_styles,
                                        style_predictor_factory_func: typing.Callable[[int], typing.Any],
                                        style_transfer_factory_func: typing.Callable[[], typing.Any],
                                        name="StyleTransferInferenceModel"):
    style_transfer_model, num_style_parameters = style_transfer_factory_func()
    style_predictor = style_predictor_factory_func(num_style_parameters)
    model = StyleTransferInferenceModel(num_styles, style_predictor, style_transfer_model, name)
    inputs = model.input

    class StyleTransferModels:
        def __init__(self):
            self.inputs = inputs
            self.inference = model
            self.transfer = style_transfer_model
            self.style_predictor = style_predictor

    return StyleTransferModels()
