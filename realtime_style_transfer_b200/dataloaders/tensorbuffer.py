"""G-buffer ingest: the raw float32 tensor-buffer wire format of the Unreal side.

Mirror of realtime_style_transfer/dataloaders/tensorbuffer.py:8-16: a file of ``prod(shape)`` little-endian float32 values,
read back in C order.  The result is a numpy array (optionally in page-locked memory so that it can be handed to
``model.predict_frames`` / ``rst_transfer_submit_host`` without another copy); planes follow ``ShapeConfig.channels`` order
(FinalImage, BaseColor, [ShadowMask], AO, Metallic, Specular, Roughness, ViewNormal, SceneDepth, LightingModel), unnormalised.
"""
from __future__ import annotations

import math
from pathlib import Path

import numpy as np


def load_tensor_from_buffer(buffer_filepath, shape, pinned: bool = False) -> np.ndarray:
    num_elements = math.prod(shape)
    path = Path(buffer_filepath)
    size = path.stat().st_size
    if size < num_elements * 4:
        raise ValueError(f"{path}: {size} bytes, need {num_elements * 4} for shape {tuple(shape)}")
    if pinned:
        import torch
        out = torch.empty(tuple(shape), dtype=torch.float32).pin_memory().numpy()
        out[...] = np.fromfile(path, dtype="<f4", count=num_elements).reshape(shape)
        return out
    return np.fromfile(path, dtype="<f4", count=num_elements).reshape(shape).astype(np.float32, copy=False)


def save_tensor_to_buffer(buffer_filepath, tensor) -> None:
    np.ascontiguousarray(tensor, dtype="<f4").tofile(str(buffer_filepath))


def iter_tensor_buffers(filepaths, shape, batch: int = 1, pinned: bool = False):
    """Yields (batch, *shape) arrays from a sequence of per-frame buffer files (the last batch may be short)."""
    chunk = []
    for p in filepaths:
        chunk.append(load_tensor_from_buffer(p, shape))
        if len(chunk) == batch:
            yield np.stack(chunk)
            chunk = []
    if chunk:
        yield np.stack(chunk)
