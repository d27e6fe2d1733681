"""Training model -- mirror of realtime_style_transfer/models/styleTransferTrainingModel.py.

``make_style_transfer_training_model`` keeps the reference's signature and returns an object with
``.loss_model .training .inference .transfer .style_predictor`` (:62-70).  ``.training`` is the slice of ``tf.keras.Model``
train_network.py uses: ``compile(optimizer=RMSprop())``, ``build``, ``fit(x=, validation_data=, epochs=, initial_epoch=,
callbacks=)``, plus ``train_step / test_step / compute_loss / compute_metrics / reset_metrics``.

One train step (Keras ``Model.train_step`` in TF 2.9) runs entirely in librst_sm100.so:
forward in training mode, the VGG loss model, back-propagation of the batch SUM of the (B,) loss vector, RMSprop.
Data parallel: one process per GPU; the flat gradient buffer is all-reduced (SUM) over NCCL between backward and update.
"""
from __future__ import annotations

import logging
import typing

import numpy as np

from .. import _native, distributed, optimizers
from ._base import NativeModel, as_numpy
from .stylePrediction import _EXTRACTOR_CODE
from .styleTransferInferenceModel import make_style_transfer_inference_model

log = logging.getLogger(__name__)

LOSS_KEYS = ("loss", "feature_loss", "style_loss", "total_variation_loss")


class History:
    def __init__(self):
        self.history: typing.Dict[str, list] = {}
        self.epoch: typing.List[int] = []


class _DeviceArray:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can wrap it without a copy."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


# noinspection PyAbstractClass
class StyleTransferTrainingModel(NativeModel):
    def __init__(self, style_loss_func, loss_model, inference_model, *args, **kwargs):
        super().__init__(kwargs.get("name", "StyleTransferTrainingModel"))
        self.style_loss_func = style_loss_func
        self.loss_model = loss_model
        self.inference_model = inference_model
        self.input = inference_model.input
        self.output_shape = inference_model.output_shape
        self.style_losses = {}
        self.optimizer = None
        self.stop_training = False
        self._trainer: typing.Optional[_native.NativeTrainer] = None
        self._host_stale = False
        self._grad_tensor = None
        self._mirrored = None       # variable versions of the sub-models the trainer's device copy corresponds to
        # the tensor-core convolutions (residual trunk, loss model) use error-compensated split tf32 = fp32-level accuracy;
        # _native.PRECISION_TF32 switches them to plain tf32 operands (TensorFlow's behaviour on Ampere and later, faster)
        self.math = _native.PRECISION_FP32

    def _versions(self):
        inf = self.inference_model
        return (inf.transfer._version, inf.style_predictor._version, id(self.loss_model), getattr(self.loss_model, "_version", 0),
                getattr(self.loss_model, "math", None), self.math)

    # -- variables live in the inference model's two sub-models ---------------------------------------------------
    def _all_variables(self):
        # after train_step the trained values live in the native trainer: weights / get_weights / trainable_variables /
        # save_weights (checkpoint callbacks in a custom loop) must never see the stale host copies
        self.sync_to_host()
        return self.inference_model._all_variables()

    @property
    def weights(self):
        return self._all_variables()

    @property
    def trainable_variables(self):
        return {k: v for k, v in self._all_variables().items() if not k.endswith(("moving_mean", "moving_variance"))}

    def set_weights(self, weights):
        self.inference_model.set_weights(weights)

    def get_weights(self):
        return [v.copy() for v in self._all_variables().values()]

    def count_params(self):
        return self.inference_model.count_params()

    # -- Keras surface ---------------------------------------------------------------------------------------------
    def compile(self, optimizer=None, run_eagerly=False, **kwargs):
        if optimizer is None or optimizer == "rmsprop":
            optimizer = optimizers.RMSprop()
        if not isinstance(optimizer, optimizers.RMSprop):
            raise NotImplementedError("only RMSprop (train_network.py:102) is built")
        super().compile(optimizer=optimizer, run_eagerly=run_eagerly, **kwargs)

    def compute_loss(self, x=None, y=None, y_pred=None, sample_weight=None):
        losses = self.style_loss_func(y_pred, y)
        self.style_losses = losses
        return losses['loss']

    def compute_metrics(self, x, y, y_pred, sample_weight):
        return {n: float(np.mean(l)) for n, l in self.style_losses.items()}

    def reset_metrics(self):
        self.style_losses = {}

    def __call__(self, inputs, training=False):
        self.sync_to_host()
        return self.inference_model(inputs)

    def predict(self, x, batch_size=None, verbose=0, **kwargs):
        self.sync_to_host()
        return self.inference_model.predict(x, batch_size=batch_size, verbose=verbose, **kwargs)

    # -- native trainer --------------------------------------------------------------------------------------------
    def _get_trainer(self, batch: int) -> _native.NativeTrainer:
        inf = self.inference_model
        if self._trainer is not None and self._trainer.cfg.max_batch < batch:
            raise NotImplementedError(f"batch {batch} exceeds the batch size the trainer was created for "
                                      f"({self._trainer.cfg.max_batch}); keep the training batch size constant")
        if self._trainer is None:
            plan = inf.transfer.plan
            if tuple(self.loss_model.input_shape[:2]) != tuple(plan.output_shape[:2]):
                raise ValueError(f"loss model works at {self.loss_model.input_shape}, transfer net produces {plan.output_shape}")
            self._trainer = _native.NativeTrainer(
                in_shape=plan.input_shape, out_shape=plan.output_shape, bottleneck_res_y=plan.bottleneck_res_y,
                bottleneck_num_filters=plan.filters, max_batch=batch,
                extractor=_EXTRACTOR_CODE[inf.style_predictor.plan.feature_extractor],
                style_shape=inf.style_predictor.plan.input_shape, device=self.device)
            self._mirrored = None
        if self._mirrored != self._versions():
            # variables were assigned on the Python side (initialisation, load_weights, set_weights): upload them.
            # The RMSprop accumulators are kept, as tf.keras keeps its slots when variables are assigned.
            self._trainer.model.set_weights(self.inference_model._all_variables(), commit=True)
            lm = self.loss_model
            if self.math == _native.PRECISION_TF32:
                self._trainer.set_math(_native.PRECISION_TF32)
            else:
                self._trainer.loss.set_math(getattr(lm, "math", _native.PRECISION_FP32))
            self._trainer.loss.set_weights(lm.weights)
            self._trainer.loss.set_factors(lm.content_loss_factor, lm.style_loss_factor, lm.total_variation_loss_factor)
            self._mirrored = self._versions()
            self._host_stale = False
        return self._trainer

    def gradient_tensor(self):
        """The trainer's flat gradient buffer as a torch CUDA tensor (no copy): what the data-parallel all-reduce sums."""
        import torch
        if self._grad_tensor is None:
            tr = self._trainer
            self._grad_tensor = torch.as_tensor(_DeviceArray(tr.gradients_ptr(), tr.num_gradient_elements),
                                                device=torch.device("cuda", self.device))
        return self._grad_tensor

    def sync_to_host(self):
        """Copies the trained variables back into the Python-side models (save_weights, checkpoints, inference)."""
        if self._trainer is None or not self._host_stale:
            return
        tr = self._trainer
        self._host_stale = False
        tr.sync_weights()
        fresh = {k: tr.model.get_weight(k, v.shape) for k, v in self.inference_model._all_variables().items()}
        self.inference_model.set_weights(fresh)
        self._mirrored = self._versions()      # the trainer already holds these values: no re-upload on the next step

    def train_step(self, data):
        """data = (x, y): x {'content': (B,H,W,C), 'style': (B,1,h,w,3)}, y {'content': (B,H,W,3), 'style': (B,1,H,W,3)}.
        Returns compute_metrics' dict (batch means of the loss terms)."""
        import torch
        if self.optimizer is None:
            raise RuntimeError("compile() the model with an optimizer before training")
        x, y = data
        content, style = as_numpy(x["content"]), as_numpy(x["style"])
        gt_content, gt_style = as_numpy(y["content"]), as_numpy(y["style"])
        if style.ndim == 5:
            assert style.shape[1] == 1, f"Training supports exactly one style. Found {style.shape[1]}"
            style = style[:, 0]
        if gt_style.ndim == 5:
            gt_style = gt_style[:, 0]
        b = content.shape[0]
        tr = self._get_trainer(b)
        dev = torch.device("cuda", self.device)
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True) for a in (content, style, gt_content, gt_style)]
        d_losses = torch.empty((b, 4), dtype=torch.float32, device=dev)
        torch.cuda.synchronize(dev)
        tr.forward_backward(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d_losses.data_ptr(), b)
        distributed.allreduce_sum_(self.gradient_tensor())      # no-op in a single process
        opt = self.optimizer
        tr.apply_gradients(opt.learning_rate, opt.rho, opt.epsilon)
        opt.iterations += 1
        self._host_stale = True
        o = d_losses.cpu().numpy()
        self.style_losses = {k: o[:, i] for i, k in enumerate(LOSS_KEYS)}
        return self.compute_metrics(x, y, None, None)

    def test_step(self, data):
        """Validation: inference-mode forward (moving statistics) and the loss model, no update."""
        x, y = data
        y_pred = self.predict(x)
        self.compute_loss(x, y, y_pred)
        return self.compute_metrics(x, y, y_pred, None)

    def fit(self, x=None, validation_data=None, epochs=1, initial_epoch=0, callbacks=None, steps_per_epoch=None,
            validation_steps=None, verbose=1, **kwargs):
        """x: an iterable of (inputs, targets) batches, re-iterated every epoch (a tf.data-like dataset or a list)."""
        callbacks = list(callbacks or [])
        history = History()
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
            else:
                cb.model = self
        _call(callbacks, "on_train_begin", None)
        self.stop_training = False
        for epoch in range(initial_epoch, epochs):
            _call(callbacks, "on_epoch_begin", epoch, None)
            sums, n = {}, 0
            for step, batch in enumerate(x):
                if steps_per_epoch is not None and step >= steps_per_epoch:
                    break
                _call(callbacks, "on_train_batch_begin", step, None)
                logs = self.train_step(batch)
                for k, v in logs.items():
                    sums[k] = sums.get(k, 0.0) + v
                n += 1
                _call(callbacks, "on_train_batch_end", step, logs)
            logs = {k: v / max(n, 1) for k, v in sums.items()}
            if validation_data is not None:
                vs, vn = {}, 0
                for step, batch in enumerate(validation_data):
                    if validation_steps is not None and step >= validation_steps:
                        break
                    for k, v in self.test_step(batch).items():
                        vs[k] = vs.get(k, 0.0) + v
                    vn += 1
                logs.update({f"val_{k}": v / max(vn, 1) for k, v in vs.items()})
            self.sync_to_host()
            history.epoch.append(epoch)
            for k, v in logs.items():
                history.history.setdefault(k, []).append(v)
            if verbose:
                log.info("epoch %d: %s", epoch, ", ".join(f"{k}={v:.6g}" for k, v in logs.items()))
            _call(callbacks, "on_epoch_end", epoch, logs)
            if self.stop_training:
                break
        _call(callbacks, "on_train_end", None)
        self.history = history
        return history

    def close(self):
        if self._trainer is not None:
            self._grad_tensor = None
            self._trainer.close()
            self._trainer = None


def _call(callbacks, method, *args):
    for cb in callbacks:
        fn = getattr(cb, method, None)
        if fn is not None:
            fn(*args)


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

    training_model = StyleTransferTrainingModel(style_loss_func, loss_model, inference_model.inference)

    class StyleTransferModels:
        def __init__(self):
            self.loss_model = loss_model
            self.style_loss_func = style_loss_func
            self.training = training_model
            self.inference = inference_model.inference
            self.transfer = inference_model.transfer
            self.style_predictor = inference_model.style_predictor

    return StyleTransferModels()
