"""Training-model factory -- mirror of realtime_style_transfer/models/styleTransferTrainingModel.py:39-70.

Builds the inference model with ``num_styles=1`` and attaches the loss model; exposes
``.loss_model .training .inference .transfer .style_predictor`` like the reference.  The training
step itself (forward + VGG/Gram loss + backward + RMSprop) is SURVEY.md section 8 row a12-a15.
"""
from __future__ import annotations

import typing

from .styleTransferInferenceModel import make_style_transfer_inference_model  # noqa: F401  (re-exported, as in the reference)


def make_style_transfer_training_model(style_predictor_factory_func: typing.Callable[[int], typing.Any],
                                       style_transfer_factory_func: typing.Callable[[], typing.Any],
                                       style_loss_func_factory_func: typing.Callable[[], typing.Any],
                                       name="StyleTransferTrainingModel"):
    inference_model = make_style_transfer_inference_model(
        num_styles=1,
        style_predictor_factory_func=style_predictor_factory_func,
        style_transfer_factory_func=style_transfer_factory_func,
        name=name)
    style_loss_func, loss_model = style_loss_func_factory_func()

    class StyleTransferModels:
        def __init__(self):
            self.loss_model = loss_model
            self.style_loss_func = style_loss_func
            self.training = inference_model.inference     # same variables; fit() is not built yet
            self.inference = inference_model.inference
            self.transfer = inference_model.transfer
            self.style_predictor = inference_model.style_predictor

    return StyleTransferModels()
