"""Host-side layer plan and variable registry (pure Python, no GPU).

Restates the block arithmetic of create_style_transfer_model (reference
realtime_style_transfer/models/styleTransfer.py:213-279) and the variable inventory of SURVEY.md
appendix B so that models can be constructed, inspected and have weights assigned on a machine
without a GPU.  The native library builds the same plan; tests check the two agree.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np

CONTRACT_FILTER_SIZES = [(16, 3, 2), (32, 3, 2), (32, 3, 2), (32, 3, 2)]
EXPAND_FILTER_SIZES = [(32, 3, 2), (16, 3, 2), (8, 3, 2), (4, 3, 2), (3, 3, 2), (3, 3, 2), (3, 3, 2), (3, 3, 2)]
NUM_RESIDUAL_BLOCKS = 5
NUM_PARAMS_PER_FEATURE = 2        # ConditionalInstanceNormalization.NumParamsPerFeature


class TransferPlan:
    def __init__(self, input_shape, output_shape, bottleneck_res_y, bottleneck_num_filters, num_styles):
        self.input_shape = tuple(int(v) for v in input_shape)
        self.output_shape = tuple(int(v) for v in output_shape)
        self.bottleneck_res_y = int(bottleneck_res_y)
        self.filters = int(bottleneck_num_filters)
        self.num_styles = int(num_styles)
        # Same float expression as the reference (styleTransfer.py:217); log2(a)-log2(b) and
        # log2(a/b) round differently on non power-of-two ratios.
        self.num_contract_blocks = math.ceil(math.log2(self.input_shape[0]) - math.log2(self.bottleneck_res_y))
        if not 1 <= self.num_contract_blocks <= len(CONTRACT_FILTER_SIZES):
            raise ValueError(f"unsupported contract depth {self.num_contract_blocks}")
        scale = 2 ** -self.num_contract_blocks
        self.bottleneck_hw = (int(self.input_shape[0] * scale), int(self.input_shape[1] * scale))
        self.num_expand_blocks = math.ceil(math.log2(self.output_shape[0]) - math.log2(self.bottleneck_hw[0]))
        if not 1 <= self.num_expand_blocks <= len(EXPAND_FILTER_SIZES):
            raise ValueError(f"unsupported expand depth {self.num_expand_blocks}")
        self.residual_in_filters = CONTRACT_FILTER_SIZES[self.num_contract_blocks - 1][0]
        self.expand_filters = [EXPAND_FILTER_SIZES[i][0] for i in range(self.num_expand_blocks)] + [3]
        self.num_style_parameters = (NUM_RESIDUAL_BLOCKS * 2 * NUM_PARAMS_PER_FEATURE * self.filters
                                     + sum(NUM_PARAMS_PER_FEATURE * f for f in self.expand_filters))

    def variables(self) -> "OrderedDict[str, Tuple[int, ...]]":
        v = OrderedDict()

        def conv_bn(name, k, ci, co):
            v[f"{name}/conv/kernel"] = (k, k, ci, co)
            v[f"{name}/conv/bias"] = (co,)
            for s in ("gamma", "beta", "moving_mean", "moving_variance"):
                v[f"{name}/bn/{s}"] = (co,)

        ci = self.input_shape[2]
        conv_bn("contract_start", 9, ci, 32)
        ci = 32
        for i in range(self.num_contract_blocks):
            co, k, _ = CONTRACT_FILTER_SIZES[i]
            conv_bn(f"contract_{i}", k, ci, co)
            ci = co
        for b in range(NUM_RESIDUAL_BLOCKS):
            for i in range(2):
                v[f"residual_block_{b}/conv{i}/kernel"] = (3, 3, ci, self.filters)
                v[f"residual_block_{b}/conv{i}/bias"] = (self.filters,)
                ci = self.filters
        for i in range(self.num_expand_blocks):
            co, k, _ = EXPAND_FILTER_SIZES[i]
            v[f"expand_{i}/conv/kernel"] = (k, k, co, ci)
            v[f"expand_{i}/conv/bias"] = (co,)
            ci = co
        v["expand_last/conv/kernel"] = (9, 9, 3, ci)
        v["expand_last/conv/bias"] = (3,)
        return v

    def initial_weights(self, rng: np.random.Generator) -> Dict[str, np.ndarray]:
        """Reference initialisers: N(0,0.02) contract/expand kernels, U(0,0.05) residual kernels,
        zero biases, BatchNorm gamma=1 beta=0 mean=0 var=1 (styleTransfer.py:97, :146, :190)."""
        out = {}
        for name, shape in self.variables().items():
            if name.endswith("kernel"):
                if name.startswith("residual_block"):
                    w = rng.uniform(0.0, 0.05, shape)
                else:
                    w = rng.normal(0.0, 0.02, shape)
            elif name.endswith(("gamma", "moving_variance")):
                w = np.ones(shape)
            else:
                w = np.zeros(shape)
            out[name] = w.astype(np.float32)
        return out


def _depth(v, divisor=8):
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


# Keras MobileNetV3Small stack (alpha 1.0): expansion, out, kernel, stride, squeeze-excite, activation
_MBV3_SMALL = [(1, 16, 3, 2, True), (72 / 16, 24, 3, 2, False), (88 / 24, 24, 3, 1, False), (4, 40, 5, 2, True),
               (6, 40, 5, 1, True), (6, 40, 5, 1, True), (3, 48, 5, 1, True), (3, 48, 5, 1, True),
               (6, 96, 5, 2, True), (6, 96, 5, 1, True), (6, 96, 5, 1, True)]


class PredictorPlan:
    def __init__(self, input_shape, feature_extractor: str, num_top_parameters: int, num_style_parameters: int = 100):
        self.input_shape = tuple(int(v) for v in input_shape)
        self.feature_extractor = feature_extractor
        self.num_top_parameters = int(num_top_parameters)
        self.num_style_parameters = int(num_style_parameters)
        if self.num_style_parameters != 100:
            # the native predictor head (StylePredictor 576 -> 100 -> P) has the reference's default bottleneck width built in
            # (stylePrediction.py:26); no reference script passes another value
            raise NotImplementedError(f"num_style_parameters={num_style_parameters}: only the reference default of 100 is built")
        if feature_extractor not in ("DUMMY", "MOBILE_NET"):
            # the reference also has EFFICIENT_NET (EfficientNetV2S); it is never selected by ShapeConfig
            raise ValueError(f"{feature_extractor} is not a valid value for feature_extractor. "
                             "Must be a StyleFeatureExtractor")

    def variables(self) -> "OrderedDict[str, Tuple[int, ...]]":
        v = OrderedDict()

        def bn(prefix, c):
            for s in ("gamma", "beta", "moving_mean", "moving_variance"):
                v[f"{prefix}/{s}"] = (c,)

        if self.feature_extractor == "DUMMY":
            v["dummy_conv/kernel"] = (9, 9, 3, 1)
            v["dummy_conv/bias"] = (1,)
            feat = 1
        else:
            v["mobilenet/Conv/kernel"] = (3, 3, 3, 16)
            bn("mobilenet/Conv/BatchNorm", 16)
            ci = 16
            for bid, (e, co, k, _s, se) in enumerate(_MBV3_SMALL):
                p = "mobilenet/expanded_conv" + (f"_{bid}" if bid else "")
                cexp = _depth(ci * e)
                if bid:
                    v[f"{p}/expand/kernel"] = (1, 1, ci, cexp)
                    bn(f"{p}/expand/BatchNorm", cexp)
                v[f"{p}/depthwise/depthwise_kernel"] = (k, k, cexp, 1)
                bn(f"{p}/depthwise/BatchNorm", cexp)
                if se:
                    cse = _depth(cexp * 0.25)
                    v[f"{p}/squeeze_excite/Conv/kernel"] = (1, 1, cexp, cse)
                    v[f"{p}/squeeze_excite/Conv/bias"] = (cse,)
                    v[f"{p}/squeeze_excite/Conv_1/kernel"] = (1, 1, cse, cexp)
                    v[f"{p}/squeeze_excite/Conv_1/bias"] = (cexp,)
                v[f"{p}/project/kernel"] = (1, 1, cexp, co)
                bn(f"{p}/project/BatchNorm", co)
                ci = co
            feat = _depth(ci * 6)
            v["mobilenet/Conv_1/kernel"] = (1, 1, ci, feat)
            bn("mobilenet/Conv_1/BatchNorm", feat)
        v["StylePredictor/kernel"] = (1, 1, feat, self.num_style_parameters)
        v["StylePredictor/bias"] = (self.num_style_parameters,)
        v["StyleNormPredictor/kernel"] = (1, 1, self.num_style_parameters, self.num_top_parameters)
        v["StyleNormPredictor/bias"] = (self.num_top_parameters,)
        return v

    def initial_weights(self, rng: np.random.Generator) -> Dict[str, np.ndarray]:
        """Heads: VarianceScaling(scale=1/3, fan_out, uniform) kernels, bias 0.5
        (stylePrediction.py:9-16, :62, :69).  The reference initialises the MobileNet body from
        ImageNet weights, which cannot be fetched here: He-normal kernels, identity BatchNorm."""
        out = {}
        for name, shape in self.variables().items():
            if name.startswith("Style") and name.endswith("kernel"):
                fan_out = shape[0] * shape[1] * shape[3]
                limit = math.sqrt(3.0 * (1.0 / 3.0) / fan_out)
                w = rng.uniform(-limit, limit, shape)
            elif name.startswith("Style"):
                w = np.full(shape, 0.5)
            elif name.endswith("depthwise_kernel"):
                w = rng.normal(0.0, math.sqrt(2.0 / (shape[0] * shape[1])), shape)
            elif name.endswith("kernel"):
                # Keras Conv2D default is glorot_uniform
                fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
                limit = math.sqrt(6.0 / (fan_in + fan_out))
                w = rng.uniform(-limit, limit, shape)
            elif name.endswith(("gamma", "moving_variance")):
                w = np.ones(shape)
            else:
                w = np.zeros(shape)
            out[name] = w.astype(np.float32)
        return out
