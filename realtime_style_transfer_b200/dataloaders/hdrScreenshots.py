"""Unreal HDR screenshot ingest -- mirror of realtime_style_transfer/dataloaders/hdrScreenshots.py:14-29.

A screenshot is a ``<stem>.png`` plus one ``<stem>_<ChannelName>.exr`` per G-buffer plane; ``expected_channels`` is the
``ShapeConfig.channels`` list of (name, num_channels) pairs and fixes the plane order of the returned (H, W, C) array.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import exr


_RGB = ("R", "G", "B")


def _planes_of(image: "exr.ExrImage", count: int) -> np.ndarray:
    """(H, W, count) from one EXR: RGB order for colour planes, the red channel for scalar planes (Unreal writes scalars
    into R), file order otherwise."""
    if count == 3:
        names = _RGB
    elif count == 1:
        names = _RGB[:1]
    else:
        names = tuple(image.channels())
    return np.stack([image.channel(n) for n in names], axis=-1)


def load_unreal_hdr_screenshot(base_png_filepath, expected_channels, dtype=np.float32):
    """Returns ((H, W, sum of plane widths), the png path), planes concatenated in ``expected_channels`` order.
    dtype float32 is the reference's result; dtype float16 keeps HALF planes as they are stored (same values, half the bytes:
    the array can be passed as 'content' to predict / predict_frames, which then upload float16)."""
    png = Path(base_png_filepath)
    keep_half = np.dtype(dtype) == np.float16
    stack = [_planes_of(exr.load(png.with_name(f"{png.stem}_{name}.exr"), keep_half=keep_half), width)
             for name, width in expected_channels]
    return np.concatenate(stack, axis=-1).astype(dtype, copy=False), png


def iter_unreal_hdr_screenshots(content_image_dir, expected_channels, batch: int = 1, dtype=np.float32):
    """Yields (batch, H, W, C) arrays for every ``*.png`` stem in the directory (the last batch may be short)."""
    chunk = []
    for png in sorted(Path(content_image_dir).glob('*.png')):
        chunk.append(load_unreal_hdr_screenshot(png, expected_channels, dtype)[0])
        if len(chunk) == batch:
            yield np.stack(chunk)
            chunk = []
    if chunk:
        yield np.stack(chunk)


def get_unreal_hdr_screenshot_dataset(content_image_dir, expected_channels, shape, **kwargs):
    """hdrScreenshots.py:32-34: every ``*.png`` stem of the directory."""
    screenshot_pngs = list(Path(content_image_dir).glob('*.png'))
    return get_unreal_hdr_screenshot_dataset_from_filepaths(screenshot_pngs, expected_channels, shape, **kwargs)


def get_unreal_hdr_screenshot_dataset_from_filepaths(screenshot_png_paths, expected_channels, shape, **kwargs):
    """hdrScreenshots.py:37-70 without tf.data: a re-iterable of G-buffer frames resized / centre-cropped to ``shape``
    (``common.preprocess_numpy_image``), shuffled once with ``seed`` if given, unreadable screenshots skipped with a warning.
    ``dtype=np.float16`` keeps HALF planes narrow when no resize is needed.  (The reference's ``output_shape`` variant also
    yields the 8-bit screenshot as ground truth; decoding PNG needs an image library that is not part of this package.)"""
    import logging
    from .common import FrameDataset, preprocess_numpy_image
    log = logging.getLogger(__name__)
    paths = list(screenshot_png_paths)
    if "seed" in kwargs:
        import random
        random.Random(kwargs["seed"]).shuffle(paths)
    if "output_shape" in kwargs:
        raise NotImplementedError("ground-truth PNG decoding is outside the accelerated path")
    dtype = kwargs.get("dtype", np.float32)

    def frames():
        for png in paths:
            try:
                channels, _ = load_unreal_hdr_screenshot(png, expected_channels, dtype)
                out = preprocess_numpy_image(channels, shape)
                yield out if out.dtype == np.dtype(dtype) else out.astype(np.float32)
            except Exception as e:                      # noqa: BLE001 - the reference skips unreadable screenshots too
                log.warning(f"Skipping f{png} due to an error: {e}")

    return FrameDataset(frames, len(paths))
