"""Host-side frame preparation -- mirror of the pieces of realtime_style_transfer/dataloaders/common.py that the G-buffer ingest
uses: ``preprocess_numpy_image`` (common.py:45-58) and a minimal stand-in for the ``tf.data.Dataset`` objects the reference's
dataset factories return (iterable, ``.num_samples``, ``.batch`` / ``.prefetch`` / ``.take``).

``tf.image.resize`` (bilinear, antialias off, half-pixel centres) and ``tf.image.resize_with_crop_or_pad`` are restated in numpy
from their TensorFlow 2.9 definitions; TensorFlow is not installable here, so they are checked against torch's
``interpolate(mode='bilinear', align_corners=False)``, which uses the same sampling rule (tests/test_host.py).
"""
from __future__ import annotations

import math

import numpy as np


def resize_bilinear(image: np.ndarray, size) -> np.ndarray:
    """tf.image.resize(image, size) with the defaults: (H,W,C) -> (size[0], size[1], C), float32."""
    image = np.asarray(image, np.float32)
    out_h, out_w = int(size[0]), int(size[1])
    in_h, in_w = image.shape[:2]

    def taps(n_in, n_out):
        src = (np.arange(n_out, dtype=np.float64) + 0.5) * (n_in / n_out) - 0.5
        lo = np.floor(src)
        frac = (src - lo).astype(np.float32)
        lower = np.clip(lo, 0, n_in - 1).astype(np.int64)
        upper = np.clip(np.ceil(src), 0, n_in - 1).astype(np.int64)
        return lower, upper, frac

    y0, y1, fy = taps(in_h, out_h)
    x0, x1, fx = taps(in_w, out_w)
    top = image[y0][:, x0] + (image[y0][:, x1] - image[y0][:, x0]) * fx[None, :, None]
    bot = image[y1][:, x0] + (image[y1][:, x1] - image[y1][:, x0]) * fx[None, :, None]
    return (top + (bot - top) * fy[:, None, None]).astype(np.float32)


def resize_with_crop_or_pad(image: np.ndarray, target_h: int, target_w: int) -> np.ndarray:
    """tf.image.resize_with_crop_or_pad: centred crop and / or zero pad to (target_h, target_w)."""
    h, w = image.shape[:2]
    if h > target_h:
        off = (h - target_h) // 2
        image = image[off:off + target_h]
    if w > target_w:
        off = (w - target_w) // 2
        image = image[:, off:off + target_w]
    h, w = image.shape[:2]
    if h < target_h or w < target_w:
        out = np.zeros((target_h, target_w) + image.shape[2:], image.dtype)
        top, left = (target_h - h) // 2, (target_w - w) // 2
        out[top:top + h, left:left + w] = image
        image = out
    return image


def preprocess_numpy_image(image: np.ndarray, shape) -> np.ndarray:
    """common.py:45-58, including its aspect-ratio expression (rows / columns of the image against shape[0] / shape[1])."""
    aspect_ratio_image = image.shape[0] / image.shape[1]
    aspect_ratio_target = shape[0] / shape[1]
    should_scale_to_target_y = aspect_ratio_image > aspect_ratio_target
    new_size = (math.ceil(shape[1] * aspect_ratio_image), shape[1]) if should_scale_to_target_y else (
        shape[0], math.ceil(shape[0] / aspect_ratio_image))
    if tuple(image.shape[:2]) != tuple(new_size):
        image = resize_bilinear(image, new_size)
    return resize_with_crop_or_pad(np.asarray(image), shape[0], shape[1])


class FrameDataset:
    """The slice of tf.data.Dataset the reference's loops use: re-iterable, ``num_samples``, ``batch``, ``prefetch``, ``take``."""

    def __init__(self, generator_fn, num_samples: int):
        self._gen, self.num_samples = generator_fn, int(num_samples)

    def __iter__(self):
        return iter(self._gen())

    def prefetch(self, _n):
        return self

    def take(self, n):
        def gen():
            for i, el in enumerate(self._gen()):
                if i >= n:
                    return
                yield el
        return FrameDataset(gen, min(self.num_samples, n))

    def batch(self, n):
        def gen():
            chunk = []
            for el in self._gen():
                chunk.append(el)
                if len(chunk) == n:
                    yield _stack(chunk)
                    chunk = []
            if chunk:
                yield _stack(chunk)
        return FrameDataset(gen, -(-self.num_samples // n))


def _stack(chunk):
    if isinstance(chunk[0], tuple):
        return tuple(np.stack([c[i] for c in chunk]) for i in range(len(chunk[0])))
    return np.stack(chunk)
