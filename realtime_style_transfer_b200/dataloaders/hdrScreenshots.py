"""Unreal HDR screenshot ingest -- mirror of realtime_style_transfer/dataloaders/hdrScreenshots.py:14-29.

A screenshot is a ``<stem>.png`` plus one ``<stem>_<ChannelName>.exr`` per G-buffer plane; ``expected_channels`` is the
``ShapeConfig.channels`` list of (name, num_channels) pairs and fixes the plane order of the returned (H, W, C) array.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

from . import exr


def load_unreal_hdr_screenshot(base_png_filepath, expected_channels):
    base_png_filepath = Path(base_png_filepath)
    channel_list = []
    for channel_name, num_channels in expected_channels:
        channel_path = base_png_filepath.parent / f"{base_png_filepath.stem}_{channel_name}.exr"
        exr_data = exr.load(channel_path)
        if num_channels == 3:
            image_tensor = np.stack([exr_data.channel('R'), exr_data.channel('G'), exr_data.channel('B')], axis=-1)
        elif num_channels == 1:
            image_tensor = np.expand_dims(exr_data.channel('R'), axis=-1)
        else:
            image_tensor = np.stack([channel for _, channel in exr_data.channels().items()], axis=-1)
        channel_list.append(image_tensor)
    all_channels = np.concatenate(channel_list, axis=-1).astype(np.float32, copy=False)
    return all_channels, base_png_filepath


def iter_unreal_hdr_screenshots(content_image_dir, expected_channels, batch: int = 1):
    """Yields (batch, H, W, C) float32 arrays for every ``*.png`` stem in the directory (the last batch may be short)."""
    chunk = []
    for png in sorted(Path(content_image_dir).glob('*.png')):
        chunk.append(load_unreal_hdr_screenshot(png, expected_channels)[0])
        if len(chunk) == batch:
            yield np.stack(chunk)
            chunk = []
    if chunk:
        yield np.stack(chunk)
