"""CPU tests of the TensorFlow-free TF2 checkpoint reader / writer."""
import os
import struct

import numpy as np
import pytest

from realtime_style_transfer_b200 import checkpoint as ck
from realtime_style_transfer_b200.models import stylePrediction, styleTransfer, styleTransferInferenceModel


def test_crc32c_known_answers():
    # RFC 3720 test vectors
    assert ck.crc32c(b"123456789") == 0xE3069283
    assert ck.crc32c(bytes(32)) == 0x8A9136AA
    assert ck.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43


def test_bundle_roundtrip_and_format(tmp_path):
    rng = np.random.default_rng(0)
    variables = {"contract_start/conv/kernel": rng.standard_normal((9, 9, 3, 32)).astype(np.float32),
                 "contract_start/conv/bias": rng.standard_normal(32).astype(np.float32),
                 "contract_start/bn/gamma": rng.standard_normal(32).astype(np.float32),
                 "residual_block_2/conv1/kernel": rng.standard_normal((3, 3, 8, 8)).astype(np.float32),
                 "expand_last/conv/kernel": rng.standard_normal((9, 9, 3, 16)).astype(np.float32),
                 "mobilenet/expanded_conv_3/depthwise/depthwise_kernel": rng.standard_normal((5, 5, 96, 1)).astype(np.float32),
                 "StylePredictor/bias": np.full(100, 0.5, np.float32)}
    extra = {"optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE": np.zeros((), np.float32)}
    prefix = ck.write_checkpoint(str(tmp_path / "ckpt-3"), variables, extra)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57              # LevelDB table magic
    index = ck.read_index(prefix)
    assert "_CHECKPOINTABLE_OBJECT_GRAPH" in index and index["_CHECKPOINTABLE_OBJECT_GRAPH"]["dtype"] == 7
    got = ck.read_checkpoint_variables(prefix)
    by_full = {e["full_name"]: e["value"] for e in got.values() if e["full_name"]}
    np.testing.assert_array_equal(by_full["contract_start_conv/kernel"], variables["contract_start/conv/kernel"])
    np.testing.assert_array_equal(by_full["batch_normalization/gamma"], variables["contract_start/bn/gamma"])
    np.testing.assert_array_equal(by_full["expanded_conv_3/depthwise/depthwise_kernel"],
                                  variables["mobilenet/expanded_conv_3/depthwise/depthwise_kernel"])
    # directory / state-file forms resolve through model_checkpoint_path like tf.train.latest_checkpoint
    assert set(ck.read_checkpoint_variables(str(tmp_path))) == set(got)
    assert set(ck.read_checkpoint_variables(str(tmp_path / "checkpoint"))) == set(got)
    # a flipped byte in the index is caught by the block CRC
    bad = bytearray(raw)
    bad[10] ^= 0xFF
    open(prefix + ".index", "wb").write(bad)
    with pytest.raises(ValueError):
        ck.read_index(prefix)
    with pytest.raises(FileNotFoundError):
        ck.read_checkpoint_variables(str(tmp_path / "nope"))


def test_load_weights_from_tf_checkpoint_matches_by_keras_names(tmp_path):
    def build():
        return styleTransferInferenceModel.make_style_transfer_inference_model(
            num_styles=1,
            style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model((64, 128, 3), "MOBILE_NET", n),
            style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model((64, 128, 17), (64, 128, 3), 16, 32, 1))
    src, dst = build(), build()
    rng = np.random.default_rng(1)
    new = {k: rng.standard_normal(v.shape).astype(np.float32) for k, v in src.inference.weights.items()}
    src.inference.set_weights(new)
    prefix = src.inference.save_weights(str(tmp_path / "weights" / "latest_epoch_weights"))     # tracing/checkpoint.py:37
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    status = dst.inference.load_weights(prefix)
    status.assert_nontrivial_match()
    status.assert_existing_objects_matched()
    for k, v in new.items():
        np.testing.assert_array_equal(dst.inference.weights[k], v)
    # the three contract BatchNorm layers are unnamed in Keras: matched in creation order
    assert np.array_equal(dst.transfer.weights["contract_1/bn/moving_variance"], new["contract_1/bn/moving_variance"])
    # a transfer-only model loading the full checkpoint: optimizer slots / predictor variables stay unused
    t2, _ = styleTransfer.create_style_transfer_model((64, 128, 17), (64, 128, 3), 16, 32, 1)
    st = t2.load_weights(prefix)
    st.assert_nontrivial_match()
    assert st.unused and not st.missing
    with pytest.raises(AssertionError):
        st.assert_consumed()
