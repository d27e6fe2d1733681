"""Shape bookkeeping for the stylization path.

Mirror of the reference's ``ShapeConfig`` (realtime_style_transfer/shape_config.py:4-84): same
constructor arguments, attribute names, ``from_spec('rst-960-120-128-17')`` grammar and G-buffer
channel table, so scripts written against the reference keep working.  Host-only, integer logic.
"""
from __future__ import annotations

import json

import numpy as np

BASE_RESOLUTION = (960, 1920)          # (rows, cols) at resolution_divider == 1, shape_config.py:14-15

# G-buffer planes in the order the loader concatenates them (shape_config.py:54-73).
_FINAL_IMAGE = [("FinalImage", 3)]
_BASE_COLOR = [("BaseColor", 3)]
_SHADOW_MASK = [("ShadowMask", 1)]
_MATERIAL_AND_GEOMETRY = [
    ("AmbientOcclusion", 1),
    ("Metallic", 1),
    ("Specular", 1),
    ("Roughness", 1),
    ("ViewNormal", 3),
    ("SceneDepth", 1),
    ("LightingModel", 3),
]


def gbuffer_channels(num_channels: int):
    """Requested channel count -> ordered (name, width) planes (shape_config.py:54-73)."""
    planes = list(_FINAL_IMAGE)
    if num_channels > 3:
        planes += _BASE_COLOR
    if num_channels >= 18:
        planes += _SHADOW_MASK
    if num_channels >= 17:
        planes += _MATERIAL_AND_GEOMETRY
    return planes


class ShapeConfig:
    def __init__(self, num_styles=1, hdr=True, bottleneck_res_y=120, bottleneck_num_filters=128,
                 resolution_divider=2, num_channels=18):
        from .models.stylePrediction import StyleFeatureExtractor

        self.bottleneck_res_y = bottleneck_res_y
        self.bottleneck_num_filters = bottleneck_num_filters
        self.num_styles = num_styles
        self.channels = gbuffer_channels(num_channels)
        self.num_channels = sum(width for _, width in self.channels)

        rows, cols = (BASE_RESOLUTION[0] // resolution_divider, BASE_RESOLUTION[1] // resolution_divider)
        self.output_shape = (rows, cols, 3)
        self.image_shape = (rows, cols, 3)
        content_shape = (rows, cols, self.num_channels) if hdr else self.image_shape
        self.input_shape = {"content": content_shape, "style": (num_styles,) + self.output_shape}
        if num_styles > 1:
            self.input_shape["style_weights"] = (rows, cols, num_styles - 1)

        self.style_feature_extractor_type = StyleFeatureExtractor.MOBILE_NET
        self.with_depth_loss = True

    @staticmethod
    def from_spec(spec: str, num_styles=1, hdr=True):
        """``rst-<res_x>-<bottleneck_res_y>-<bottleneck_filters>-<channels>`` (shape_config.py:32-48)."""
        _prefix, res_x, res_y, filters, channels = spec.split("-")[:5]
        return ShapeConfig(num_styles, hdr, int(res_y), int(filters), BASE_RESOLUTION[1] // int(res_x),
                           int(channels))

    def spec(self) -> str:
        return (f"rst-{self.output_shape[1]}-{self.bottleneck_res_y}-"
                f"{self.bottleneck_num_filters}-{self.num_channels}")

    def __str__(self):
        return json.dumps(self.__dict__, indent=4)

    def get_dummy_input_element(self):
        """All-zero (inputs, ground_truth) pair with batch 1 (shape_config.py:75-84); numpy here."""
        element = {name: np.zeros((1,) + tuple(shape), np.float32) for name, shape in self.input_shape.items()}
        ground_truth = {
            "content": np.zeros((1,) + self.output_shape, np.float32),
            "style": np.zeros((1, self.num_styles) + self.output_shape, np.float32),
        }
        return element, ground_truth
