"""CPU tests of the ONNX export (reference: save_using_checkpoint.py:76-103).

The written files are decoded and executed by oracle/onnx_ref.py (an independent ONNX reader + interpreter; neither onnx nor
onnxruntime is installable here) and compared with the oracle network on the same weights and inputs."""
import numpy as np
import pytest
import torch

from oracle import onnx_ref
from oracle import rst_oracle as O
from realtime_style_transfer_b200 import export
from realtime_style_transfer_b200.models import stylePrediction, styleTransfer
from realtime_style_transfer_b200.shape_config import ShapeConfig


@pytest.mark.parametrize("channels,filters,styles,out_mult", [(17, 32, 1, 1), (18, 16, 2, 1), (3, 8, 1, 2)])
def test_transfer_onnx_matches_the_oracle(tmp_path, channels, filters, styles, out_mult):
    shape_in, shape_out = (32, 64, channels), (32 * out_mult, 64 * out_mult, 3)     # out_mult 2: one more expand block than contract
    model, p = styleTransfer.create_style_transfer_model(shape_in, shape_out, 8, filters, styles)
    spec = O.TransferSpec(shape_in, shape_out, 8, filters, styles)
    weights = O.init_transfer_weights(spec, seed=3, trained_like=True)
    model.set_weights(weights)
    path = export.export_onnx(model, tmp_path / "transfer.onnx")
    m = onnx_ref.load(path)
    assert m.ir_version == 8 and m.opset == 13
    names = [n for n, _ in m.inputs]
    assert names == ["content", "style_params"] + (["style_weights"] if styles == 2 else [])
    assert dict(m.inputs)["content"] == ["N", 32, 64, channels] and dict(m.outputs)["stylised"] == ["N"] + list(shape_out)
    ops = {n["op"] for n in m.nodes}
    assert {"Conv", "ConvTranspose", "InstanceNormalization", "BatchNormalization", "Sigmoid"} <= ops
    content = O.synthetic_content(2, 32, 64, ShapeConfig(num_channels=channels).channels, seed=1, unit_depth=True)
    params = np.random.default_rng(2).uniform(0.3, 1.2, (2, styles, p)).astype(np.float32)
    feeds = {"content": content, "style_params": params}
    sw = None
    if styles == 2:
        sw = O.synthetic_style_weights(2, shape_out[0], shape_out[1])
        feeds["style_weights"] = sw
    got = m.run(feeds)["stylised"]
    ref = O.transfer_forward(spec, weights, content, params, sw, dtype=torch.float64).numpy()
    assert got.shape == ref.shape
    assert np.abs(got - ref).max() < 1e-5           # the initialisers are float32 copies of the variables


@pytest.mark.parametrize("extractor,hw", [("MOBILE_NET", (64, 96)), ("MOBILE_NET", (40, 72)), ("DUMMY", (32, 48))])
def test_predictor_onnx_matches_the_oracle(tmp_path, extractor, hw):
    shape = hw + (3,)
    model = stylePrediction.create_style_prediction_model(shape, extractor, 70)
    weights = O.init_predictor_weights(extractor, 70, seed=5)
    model.set_weights(weights)
    path = export.export_onnx(model, tmp_path / "predictor.onnx")
    m = onnx_ref.load(path)
    assert [n for n, _ in m.inputs] == ["style"] and dict(m.outputs)["style_params"] == ["N", 70]
    style = np.random.default_rng(6).uniform(0, 1, (2,) + shape).astype(np.float32)
    got = m.run({"style": style})["style_params"]
    ref = O.predictor_forward(extractor, weights, style, dtype=torch.float64).numpy()
    assert got.shape == ref.shape == (2, 70)
    assert np.abs(got - ref).max() < 1e-5 * max(1.0, np.abs(ref).max())


def test_model_save_writes_onnx(tmp_path):
    model, _ = styleTransfer.create_style_transfer_model((32, 64, 3), (32, 64, 3), 8, 8, 1)
    path = model.save(str(tmp_path / "m.transfer.onnx"))
    assert onnx_ref.load(path).nodes
    with pytest.raises(NotImplementedError):
        model.save(str(tmp_path / "m.transfer.tf"), save_format="tf")
