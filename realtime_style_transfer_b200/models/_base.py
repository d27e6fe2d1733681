"""Keras-flavoured model facade over the native context.

The reference's callers use a small slice of ``tf.keras.Model``:
``model(dict)``, ``.predict(dict, batch_size=, verbose=)``, ``.trainable``, ``.compile(...)``,
``.build(...)``, ``.load_weights(path)`` -> status with ``.assert_nontrivial_match()``, ``.input``,
``.output_shape``.  This class provides exactly that slice; all arithmetic happens in
librst_sm100.so (no CPU fallback).
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import numpy as np

from .. import mixed_precision
from .._native import NativeContext, RstError


class InputSpec:
    """Stand-in for a Keras symbolic input: carries ``.shape`` with a leading None batch axis."""

    def __init__(self, shape, name=None):
        self.shape = (None,) + tuple(shape)
        self.name = name

    def __repr__(self):
        return f"InputSpec(shape={self.shape}, name={self.name!r})"


class LoadStatus:
    """Mirror of the TF checkpoint load status object (predict_using_checkpoint.py:84-85)."""

    def __init__(self, matched, missing, unused):
        self.matched, self.missing, self.unused = list(matched), list(missing), list(unused)

    def assert_nontrivial_match(self):
        if not self.matched:
            raise AssertionError("Nothing except the root object matched a checkpointed value.")
        return self

    def assert_consumed(self):
        if self.missing or self.unused:
            raise AssertionError(f"Unresolved variables: missing={self.missing[:5]} unused={self.unused[:5]}")
        return self

    def assert_existing_objects_matched(self):
        if self.missing:
            raise AssertionError(f"Model variables without a checkpoint value: {self.missing[:5]}")
        return self

    def expect_partial(self):
        return self


def _is_torch_cuda(x) -> bool:
    return hasattr(x, "data_ptr") and hasattr(x, "is_cuda") and x.is_cuda


class NativeModel:
    """Common machinery: named variables on the host, lazily built native context on the GPU."""

    def __init__(self, name: str):
        self.name = name
        self.trainable = True
        self._variables: Dict[str, np.ndarray] = {}
        self._ctx: Optional[NativeContext] = None
        self._ctx_key = None
        self._version = 0           # bumped whenever a variable is assigned
        self._uploaded = None       # _weights_version() at the last upload into the native context
        self._compiled = False
        self.device = int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("RST_DEVICE") is None \
            else int(os.environ["RST_DEVICE"])

    # -- Keras surface ------------------------------------------------------------------------
    def compile(self, optimizer=None, run_eagerly=False, **kwargs):
        self.optimizer = optimizer
        self._compiled = True

    def build(self, input_shape=None):
        return None

    def summary(self, print_fn=print):
        total = 0
        for k, v in self._variables.items():
            print_fn(f"{k:70s} {tuple(v.shape)}")
            total += v.size
        print_fn(f"Total params: {total}")

    @property
    def weights(self):
        return dict(self._variables)

    def count_params(self) -> int:
        return int(sum(v.size for v in self._variables.values()))

    def get_weights(self):
        return [v.copy() for v in self._variables.values()]

    def set_weights(self, weights):
        """Accepts a name->array dict, or a list in variable order (Keras style)."""
        if isinstance(weights, dict):
            items = weights.items()
        else:
            weights = list(weights)
            if len(weights) != len(self._variables):
                raise ValueError(f"expected {len(self._variables)} arrays, got {len(weights)}")
            items = zip(self._variables.keys(), weights)
        for name, value in items:
            if name not in self._variables:
                raise ValueError(f"unknown variable {name!r} for model {self.name}")
            value = np.asarray(value, np.float32)
            if value.shape != self._variables[name].shape:
                raise ValueError(f"shape mismatch for {name}: {value.shape} vs {self._variables[name].shape}")
            self._variables[name] = np.ascontiguousarray(value)
        self._mark_dirty()

    def _mark_dirty(self):
        self._version += 1

    def _weights_version(self):
        """Identifies the variable values this model would upload; a context re-uploads only when it changes."""
        return self._version

    def save_weights(self, filepath, save_format=None):
        """``*.npz`` -> numpy archive; anything else -> a TF2 object-based checkpoint prefix (Keras' default 'tf' format:
        ``<prefix>.index`` + ``<prefix>.data-00000-of-00001``), readable by load_weights."""
        path = str(filepath)
        if path.endswith(".npz") or save_format == "npz":
            if not path.endswith(".npz"):
                path += ".npz"
            np.savez(path, **self._all_variables())
            return path
        from ..checkpoint import write_checkpoint
        return write_checkpoint(path, self._all_variables())

    def _all_variables(self) -> Dict[str, np.ndarray]:
        return self._variables

    def save(self, filepath, include_optimizer=False, save_format=None, **kwargs):
        """Keras ``Model.save`` as the reference's export script uses it (save_using_checkpoint.py:76-103).  ``*.onnx`` writes
        the graph + weights as ONNX (what that script produces through tf2onnx); a TensorFlow SavedModel cannot be written
        without TensorFlow -- use ``save_weights`` for a TF2 checkpoint of the variables."""
        path = str(filepath)
        if path.endswith(".onnx") or save_format == "onnx":
            from ..export import export_onnx
            return export_onnx(self, path if path.endswith(".onnx") else path + ".onnx")
        raise NotImplementedError("SavedModel export needs TensorFlow; supported: model.save('<name>.onnx') and save_weights()")

    def load_weights(self, filepath):
        """``.npz`` written by save_weights, or a TF2 object-based checkpoint prefix
        (tracing/checkpoint.py:37 writes those); returns a status object like TF's."""
        from ..checkpoint import read_checkpoint_variables, match_checkpoint_to_model
        path = str(filepath)
        if path.endswith(".npz") and os.path.exists(path):
            data = np.load(path)
            source = {k: data[k] for k in data.files}
            mine = self._all_variables()
            matched = [k for k in mine if k in source and source[k].shape == mine[k].shape]
            self._assign({k: source[k] for k in matched})
            return LoadStatus(matched, [k for k in mine if k not in matched], [k for k in source if k not in mine])
        ckpt_vars = read_checkpoint_variables(path)
        assignment, missing, unused = match_checkpoint_to_model(ckpt_vars, self)
        self._assign(assignment)
        return LoadStatus(list(assignment), missing, unused)

    def _assign(self, named: Dict[str, np.ndarray]):
        self.set_weights(named)

    # -- native context -----------------------------------------------------------------------
    def _precision(self) -> int:
        return mixed_precision.native_precision()

    def _context_kwargs(self, max_batch: int) -> dict:
        raise NotImplementedError

    def _native_weights(self) -> Dict[str, np.ndarray]:
        return self._all_variables()

    def _get_ctx(self, batch: int) -> NativeContext:
        key = (self._precision(), self.device)
        if self._ctx is not None and (self._ctx_key != key or self._ctx.cfg.max_batch < batch):
            self._ctx.close()
            self._ctx = None
        if self._ctx is None:
            self._ctx = NativeContext(max_batch=max(batch, 1), precision=key[0], device=self.device,
                                      **self._context_kwargs(batch))
            self._ctx_key = key
            self._uploaded = None
        version = self._weights_version()
        if self._uploaded != version:
            self._ctx.set_weights(self._native_weights(), commit=True)
            self._uploaded = version
        return self._ctx

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None


def as_content(x) -> np.ndarray:
    """G-buffer input: float16 arrays stay float16 (reduced-byte ingest, rst_transfer_*_typed), anything else -> float32."""
    if isinstance(x, np.ndarray) and x.dtype == np.float16:
        return np.ascontiguousarray(x)
    return as_numpy(x)


def as_numpy(x) -> np.ndarray:
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    elif hasattr(x, "numpy") and not isinstance(x, np.ndarray):
        x = x.numpy()
    return np.ascontiguousarray(x, dtype=np.float32)
