"""CPU tests of the host-side mirror of the reference surface and of the C-ABI library's exports."""
import ctypes
import os
import re

import numpy as np
import pytest

import realtime_style_transfer_b200 as rst
from realtime_style_transfer_b200 import _native, mixed_precision, optimizers
from realtime_style_transfer_b200._plan import PredictorPlan, TransferPlan
from realtime_style_transfer_b200.models import (stylePrediction, styleTransfer, styleTransferInferenceModel,
                                                 styleTransferTrainingModel, styleLoss)
from realtime_style_transfer_b200.shape_config import ShapeConfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- ShapeConfig (shape_config.py:4-84) -------------------------------------------------------
def test_shape_config_defaults_match_reference():
    c = ShapeConfig()
    assert c.num_styles == 1 and c.bottleneck_res_y == 120 and c.bottleneck_num_filters == 128
    assert c.num_channels == 18 and c.output_shape == (480, 960, 3) and c.image_shape == (480, 960, 3)
    assert c.input_shape == {"content": (480, 960, 18), "style": (1, 480, 960, 3)}
    assert c.style_feature_extractor_type == stylePrediction.StyleFeatureExtractor.MOBILE_NET
    assert c.with_depth_loss is True


@pytest.mark.parametrize("n,names", [
    (3, ["FinalImage"]),
    (6, ["FinalImage", "BaseColor"]),
    (17, ["FinalImage", "BaseColor", "AmbientOcclusion", "Metallic", "Specular", "Roughness", "ViewNormal",
          "SceneDepth", "LightingModel"]),
    (18, ["FinalImage", "BaseColor", "ShadowMask", "AmbientOcclusion", "Metallic", "Specular", "Roughness",
          "ViewNormal", "SceneDepth", "LightingModel"]),
])
def test_channel_table(n, names):
    c = ShapeConfig(num_channels=n)
    assert [p[0] for p in c.channels] == names
    assert c.num_channels == n == sum(p[1] for p in c.channels)


def test_from_spec_and_sdr_and_two_styles():
    c = ShapeConfig.from_spec("rst-960-120-128-17")
    assert c.input_shape["content"] == (480, 960, 17) and c.spec() == "rst-960-120-128-17"
    c = ShapeConfig.from_spec("rst-960-120-32-3", num_styles=2, hdr=False)
    assert c.input_shape["content"] == (480, 960, 3)
    assert c.input_shape["style"] == (2, 480, 960, 3) and c.input_shape["style_weights"] == (480, 960, 1)
    c = ShapeConfig.from_spec("rst-1920-120-128-18")
    assert c.output_shape == (960, 1920, 3)
    el, gt = ShapeConfig(num_styles=2).get_dummy_input_element()
    assert el["content"].shape == (1, 480, 960, 18) and el["style_weights"].shape == (1, 480, 960, 1)
    assert gt["style"].shape == (1, 2, 480, 960, 3)


# ---- plan / factories -------------------------------------------------------------------------
def test_transfer_plan_geometries_of_the_reference_tests():
    p = TransferPlan((480, 960, 17), (480, 960, 3), 120, 128, 1)
    assert (p.num_contract_blocks, p.num_expand_blocks, p.num_style_parameters) == (2, 2, 2662)
    p = TransferPlan((480, 960, 3), (1920, 3840, 3), 120, 128, 2)       # styleTransferInferenceModelTest.py:18-44
    assert (p.num_contract_blocks, p.num_expand_blocks) == (2, 4)
    p = TransferPlan((240, 480, 3), (480, 960, 3), 30, 4, 1)            # styleTransferTrainingModelTest.py:15-44
    assert (p.num_contract_blocks, p.num_expand_blocks) == (3, 4)
    v = TransferPlan((480, 960, 17), (480, 960, 3), 120, 128, 1).variables()
    assert v["contract_start/conv/kernel"] == (9, 9, 17, 32)
    assert v["residual_block_0/conv0/kernel"] == (3, 3, 32, 128)
    assert v["expand_0/conv/kernel"] == (3, 3, 32, 128) and v["expand_last/conv/kernel"] == (9, 9, 3, 16)
    assert sum(int(np.prod(s)) for s in v.values()) == 1464019 + 320


def test_plans_agree_with_oracle_registry():
    from oracle import rst_oracle as O
    for args in [((480, 960, 17), (480, 960, 3), 120, 128, 1), ((240, 480, 3), (480, 960, 3), 30, 4, 1)]:
        assert dict(TransferPlan(*args).variables()) == O.TransferSpec(*args).weight_shapes()
    for ext in ("DUMMY", "MOBILE_NET"):
        assert dict(PredictorPlan((480, 960, 3), ext, 742).variables()) == O.predictor_weight_shapes(ext, 742)


def test_factories_keep_reference_signatures():
    cfg = ShapeConfig.from_spec("rst-960-120-32-3", num_styles=2)
    models = styleTransferInferenceModel.make_style_transfer_inference_model(
        num_styles=cfg.num_styles,
        style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model(
            cfg.input_shape["style"][1:], stylePrediction.StyleFeatureExtractor.DUMMY, n),
        style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model(
            cfg.input_shape["content"], cfg.output_shape, cfg.bottleneck_res_y, cfg.bottleneck_num_filters,
            cfg.num_styles))
    assert set(models.inputs) == {"content", "style", "style_weights"}
    assert models.inference.output_shape == (None, 480, 960, 3)
    assert models.transfer.input["style_params"].shape == (None, 2, 742)
    assert models.inference.input["style"].shape == (None, 2, 480, 960, 3)
    assert models.style_predictor.output_shape == (None, 742)
    models.inference.trainable = False
    models.inference.compile(run_eagerly=False)
    with pytest.raises(ValueError):
        stylePrediction.create_style_prediction_model((480, 960, 3), "NOT_AN_EXTRACTOR", 10)


def test_initialisers_follow_reference():
    m, _ = styleTransfer.create_style_transfer_model((480, 960, 3), (480, 960, 3), 120, 32, 1)
    w = m.weights
    res = w["residual_block_2/conv1/kernel"]
    assert res.min() >= 0 and res.max() <= 0.05 and abs(res.mean() - 0.025) < 2e-3     # U(0, 0.05)
    assert abs(w["contract_0/conv/kernel"].std() - 0.02) < 2e-3                          # N(0, 0.02)
    assert not w["expand_last/conv/bias"].any() and (w["contract_1/bn/gamma"] == 1).all()
    sp = stylePrediction.create_style_prediction_model((480, 960, 3), "DUMMY", 742)
    assert (sp.weights["StylePredictor/bias"] == 0.5).all()
    lim = np.sqrt(1.0 / 742)
    assert np.abs(sp.weights["StyleNormPredictor/kernel"]).max() <= lim + 1e-7


def test_set_weights_and_npz_roundtrip(tmp_path):
    m, _ = styleTransfer.create_style_transfer_model((32, 64, 3), (32, 64, 3), 8, 4, 1)
    new = {k: np.full_like(v, 0.25) for k, v in m.weights.items()}
    m.set_weights(new)
    path = m.save_weights(str(tmp_path / "w"))
    m2, _ = styleTransfer.create_style_transfer_model((32, 64, 3), (32, 64, 3), 8, 4, 1)
    status = m2.load_weights(path)
    status.assert_nontrivial_match().assert_consumed()
    assert all((m2.weights[k] == 0.25).all() for k in new)
    with pytest.raises(ValueError):
        m.set_weights({"contract_start/conv/kernel": np.zeros((3, 3, 3, 3), np.float32)})
    with pytest.raises(ValueError):
        m.set_weights({"nope": np.zeros(1, np.float32)})


def test_style_param_stack_and_helpers():
    stack = styleTransfer.StyleParamStack(np.arange(20).reshape(1, 1, 1, 20), None)
    assert stack.get_params(4).tolist() == [[[[0, 1, 2, 3]]]]
    assert stack.make_content_and_style_input("c", 6)["style_params"].shape == (1, 1, 1, 6)
    assert stack.lower_bound == 10
    cin = styleTransfer.ConditionalInstanceNormalization
    assert cin.get_style_params_shape_and_num(128, 2) == ((1, 2, 256), 256)
    assert cin.get_style_weights_shape((120, 240, 128), 2, multiplier=2) == (240, 480, 2)
    assert styleTransfer.calc_next_conv_dims((480, 960, 17), 32, 0.25) == (120, 240, 32)


def test_training_factory_and_loss_guards():
    loss_model = styleLoss.StyleLossModelVGG((480, 960, 3))
    assert loss_model.content_loss_factor == 1e4 and loss_model.style_loss_factor == 1e-3
    assert loss_model.style_layers == ['block1_conv2', 'block2_conv2', 'block3_conv3', 'block4_conv3']
    with pytest.raises(AssertionError):
        styleLoss.make_style_loss_function(loss_model, (480, 960, 3), 2, with_depth_loss=False)
    tm = styleTransferTrainingModel.make_style_transfer_training_model(
        style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model((480, 960, 3), "DUMMY", n),
        style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model((240, 480, 3), (480, 960, 3), 30,
                                                                                      4, num_styles=1),
        style_loss_func_factory_func=lambda: styleLoss.make_style_loss_function(loss_model, (480, 960, 3), 1,
                                                                                with_depth_loss=False))
    assert tm.inference.output_shape == (None, 480, 960, 3)
    # the Keras slice train_network.py:102-138 touches
    assert tm.training.inference_model is tm.inference and tm.loss_model is loss_model
    with pytest.raises(RuntimeError):
        tm.training.train_step(({"content": np.zeros((1, 240, 480, 3)), "style": np.zeros((1, 1, 480, 960, 3))},
                                {"content": np.zeros((1, 480, 960, 3)), "style": np.zeros((1, 1, 480, 960, 3))}))
    with pytest.raises(NotImplementedError):
        tm.training.compile(optimizer="adam")
    opt = optimizers.RMSprop()
    assert (opt.learning_rate, opt.rho, opt.epsilon, opt.momentum, opt.centered) == (1e-3, 0.9, 1e-7, 0.0, False)
    with pytest.raises(NotImplementedError):
        optimizers.RMSprop(momentum=0.9)
    tm.training.compile(run_eagerly=False, optimizer=opt)
    tm.training.build(input_shape={"content": (None, 240, 480, 3), "style": (None, 1, 480, 960, 3)})
    assert set(tm.training.trainable_variables) == {k for k in tm.inference.weights if "moving_" not in k}


def test_mixed_precision_policy():
    mixed_precision.set_global_policy("mixed_bfloat16")
    assert mixed_precision.native_precision() == _native.PRECISION_BF16
    mixed_precision.set_global_policy("float32")
    assert mixed_precision.native_precision() == _native.PRECISION_FP32
    with pytest.raises(ValueError):
        mixed_precision.set_global_policy("float16")


# ---- the C-ABI library ------------------------------------------------------------------------
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "rst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rst_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_native.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(_native.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rst_b200.h but not exported"
    assert sorted(_native.SIGNATURES) == declared, "ctypes signature table out of sync with the header"
    lib.rst_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.rst_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m, p = styleTransfer.create_style_transfer_model((32, 64, 3), (32, 64, 3), 8, 4, 1)
    with pytest.raises(_native.RstError):
        m.predict({"content": np.zeros((1, 32, 64, 3), np.float32), "style_params": np.zeros((1, 1, p), np.float32)})


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import, include or dlopen it."""
    pkg = os.path.join(ROOT, "realtime_style_transfer_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<][^\n]*oracle|oracle[/\\.]_ref|import_module\([^)]*oracle",
                     re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), f"{f} reaches into oracle/"


def test_tensorbuffer_roundtrip(tmp_path):
    from realtime_style_transfer_b200.dataloaders import tensorbuffer
    x = np.random.default_rng(0).standard_normal((6, 8, 17)).astype(np.float32)
    tensorbuffer.save_tensor_to_buffer(tmp_path / "f0.bin", x)
    tensorbuffer.save_tensor_to_buffer(tmp_path / "f1.bin", x * 2)
    np.testing.assert_array_equal(tensorbuffer.load_tensor_from_buffer(tmp_path / "f0.bin", (6, 8, 17)), x)
    batches = list(tensorbuffer.iter_tensor_buffers([tmp_path / "f0.bin", tmp_path / "f1.bin", tmp_path / "f0.bin"], (6, 8, 17), 2))
    assert [b.shape for b in batches] == [(2, 6, 8, 17), (1, 6, 8, 17)]
    np.testing.assert_array_equal(batches[0][1], x * 2)
    with pytest.raises(ValueError):
        tensorbuffer.load_tensor_from_buffer(tmp_path / "f0.bin", (6, 8, 18))


def test_exr_reader_roundtrip_and_screenshot_layout(tmp_path):
    """dataloaders/hdrScreenshots.py:14-29 without pyroexr: scan-line EXR (NONE / ZIPS / ZIP, HALF / FLOAT)."""
    from realtime_style_transfer_b200.dataloaders import exr, hdrScreenshots
    rng = np.random.default_rng(0)
    h, w = 37, 50                                          # not a multiple of the 16-line ZIP block
    planes = {c: rng.uniform(0, 4, (h, w)).astype(np.float32) for c in "RGB"}
    for comp in ("NONE", "ZIPS", "ZIP"):
        for ptype, tol in (("FLOAT", 0.0), ("HALF", 2e-3)):
            path = tmp_path / f"t_{comp}_{ptype}.exr"
            exr.save(path, planes, compression=comp, pixel_type=ptype)
            img = exr.load(path)
            assert img.header["compression"] == comp and list(img.channels()) == ["B", "G", "R"] and img.shape == (h, w)
            for c in "RGB":
                assert np.abs(img.channel(c) - planes[c]).max() <= tol * 4
    # smooth data really goes through the zlib path (compressed block smaller than raw)
    smooth = {"R": np.tile(np.linspace(0, 1, w, dtype=np.float32), (h, 1))}
    exr.save(tmp_path / "s.exr", smooth, compression="ZIP", pixel_type="FLOAT")
    assert (tmp_path / "s.exr").stat().st_size < h * w * 4 // 2
    assert np.array_equal(exr.load(tmp_path / "s.exr").channel("R"), smooth["R"])
    with pytest.raises(ValueError):
        (tmp_path / "bad.exr").write_bytes(b"\0" * 64)
        exr.load(tmp_path / "bad.exr")
    # one screenshot = <stem>.png + <stem>_<Channel>.exr per plane, concatenated in ShapeConfig.channels order
    cfg = ShapeConfig(hdr=True, num_styles=1)
    truth = []
    for name, n in cfg.channels:
        data = rng.uniform(0, 1, (h, w, n)).astype(np.float32)
        truth.append(data)
        exr.save(tmp_path / f"shot_{name}.exr", {c: data[..., i] for i, c in enumerate("RGB"[:n])}, "ZIP", "FLOAT")
    (tmp_path / "shot.png").write_bytes(b"")
    arr, path = hdrScreenshots.load_unreal_hdr_screenshot(tmp_path / "shot.png", cfg.channels)
    assert arr.shape == (h, w, sum(n for _, n in cfg.channels)) and np.array_equal(arr, np.concatenate(truth, axis=-1))
    batches = list(hdrScreenshots.iter_unreal_hdr_screenshots(tmp_path, cfg.channels, batch=2))
    assert len(batches) == 1 and batches[0].shape == (1,) + arr.shape


# ---- weight-upload tracking of the combined inference model (shared dirty flag) ---------------------------
class _StubCtx:
    """Stands in for NativeContext: counts uploads, echoes one bias so stale weights are visible."""
    uploads = 0

    def __init__(self, **kw):
        import types
        self.cfg = types.SimpleNamespace(max_batch=kw.get("max_batch", 1), out_h=8, out_w=8)
        self.w = {}

    def set_weights(self, weights, commit=True):
        type(self).uploads += 1
        self.w = {k: np.array(v) for k, v in weights.items()}

    def _bias(self):
        return float(self.w["expand_last/conv/bias"][0])

    def inference_forward_host(self, content, style, weights=None):
        return np.full((content.shape[0], 1), self._bias(), np.float32)

    def transfer_forward_host(self, content, params, weights=None, out_dtype=np.float32):
        return np.full((content.shape[0], 1), self._bias(), np.float32)

    def close(self):
        pass


def test_inference_model_uploads_once_and_never_runs_stale(monkeypatch):
    from realtime_style_transfer_b200.models import _base
    monkeypatch.setattr(_base, "NativeContext", _StubCtx)
    _StubCtx.uploads = 0
    shape_in, shape_out = (32, 64, 3), (32, 64, 3)
    models = styleTransferInferenceModel.make_style_transfer_inference_model(
        1, lambda n: stylePrediction.create_style_prediction_model(shape_out, "DUMMY", n),
        lambda: styleTransfer.create_style_transfer_model(shape_in, shape_out, 8, 32, 1))
    x = {"content": np.zeros((1,) + shape_in, np.float32), "style": np.zeros((1, 1) + shape_out, np.float32)}
    for _ in range(3):
        models.inference.predict(x)
    assert _StubCtx.uploads == 1, "every predict() re-uploaded (and re-committed) all weights"
    # assign through the sub-model, consume the change in the sub-model's OWN context, then use the combined model
    models.transfer.set_weights({"expand_last/conv/bias": np.full(3, 7.0, np.float32)})
    params = np.zeros((1, 1, models.transfer.plan.num_style_parameters), np.float32)
    assert models.transfer.predict({"content": x["content"], "style_params": params})[0, 0] == 7.0
    assert models.inference.predict(x)[0, 0] == 7.0, "combined model ran with stale transfer weights"
    n = _StubCtx.uploads
    models.inference.predict(x)
    assert _StubCtx.uploads == n


def test_half_exr_planes_stay_float16_for_the_reduced_byte_ingest(tmp_path):
    """load_unreal_hdr_screenshot(dtype=float16): HALF planes are returned as stored (same values as the float32 path, half the
    bytes) -- the array the float16 entry points upload; FLOAT planes are narrowed only on request."""
    from realtime_style_transfer_b200.dataloaders import exr, hdrScreenshots
    rng = np.random.default_rng(3)
    h, w = 20, 36
    cfg = ShapeConfig(hdr=True, num_styles=1, num_channels=17)
    for name, n in cfg.channels:
        data = rng.uniform(0, 4, (h, w, n)).astype(np.float32)
        exr.save(tmp_path / f"shot_{name}.exr", {c: data[..., i] for i, c in enumerate("RGB"[:n])}, "ZIP", "HALF")
    (tmp_path / "shot.png").write_bytes(b"")
    f32, _ = hdrScreenshots.load_unreal_hdr_screenshot(tmp_path / "shot.png", cfg.channels)
    f16, _ = hdrScreenshots.load_unreal_hdr_screenshot(tmp_path / "shot.png", cfg.channels, dtype=np.float16)
    assert f32.dtype == np.float32 and f16.dtype == np.float16 and f16.shape == f32.shape == (h, w, 17)
    assert np.array_equal(f16.astype(np.float32), f32)             # HALF -> float32 is exact: identical values
    assert exr.load(tmp_path / "shot_FinalImage.exr", keep_half=True).channel("R").dtype == np.float16
    batches = list(hdrScreenshots.iter_unreal_hdr_screenshots(tmp_path, cfg.channels, batch=2, dtype=np.float16))
    assert len(batches) == 1 and batches[0].shape == (1, h, w, 17) and batches[0].dtype == np.float16


def test_preprocess_numpy_image_and_screenshot_dataset(tmp_path):
    """common.preprocess_numpy_image (common.py:45-58): bilinear resize with half-pixel centres + centred crop, against torch's
    interpolate (same sampling rule); and the dataset factory names of hdrScreenshots.py:32-70."""
    import torch
    from realtime_style_transfer_b200.dataloaders import common, exr, hdrScreenshots
    rng = np.random.default_rng(4)
    img = rng.uniform(0, 4, (30, 50, 5)).astype(np.float32)
    for size in ((60, 100), (45, 80), (17, 23), (30, 50)):
        got = common.resize_bilinear(img, size)
        ref = torch.nn.functional.interpolate(torch.as_tensor(img).permute(2, 0, 1)[None], size=size, mode="bilinear",
                                              align_corners=False, antialias=False)[0].permute(1, 2, 0).numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-5
    assert np.array_equal(common.preprocess_numpy_image(img, (30, 50, 5)), img)              # already the target shape
    out = common.preprocess_numpy_image(img, (24, 48, 5))                                     # wider target: scale rows, crop columns
    assert out.shape == (24, 48, 5)
    padded = common.resize_with_crop_or_pad(img, 34, 40)
    assert padded.shape == (34, 40, 5) and np.array_equal(padded[2:32], img[:, 5:45]) and not padded[:2].any()
    cfg = ShapeConfig(hdr=True, num_styles=1, num_channels=17)
    for stem in ("a", "b", "c"):
        for name, n in cfg.channels:
            data = rng.uniform(0, 1, (20, 40, n)).astype(np.float32)
            exr.save(tmp_path / f"{stem}_{name}.exr", {c: data[..., i] for i, c in enumerate("RGB"[:n])}, "ZIP", "HALF")
        (tmp_path / f"{stem}.png").write_bytes(b"")
    ds = hdrScreenshots.get_unreal_hdr_screenshot_dataset(tmp_path, cfg.channels, (20, 40, 17), seed=1)
    assert ds.num_samples == 3
    frames = list(ds.prefetch(5))
    assert len(frames) == 3 and frames[0].shape == (20, 40, 17) and frames[0].dtype == np.float32
    batches = list(ds.batch(2))
    assert [b.shape[0] for b in batches] == [2, 1] and len(list(ds)) == 3                      # re-iterable
    small = list(hdrScreenshots.get_unreal_hdr_screenshot_dataset(tmp_path, cfg.channels, (10, 20, 17)))
    assert small[0].shape == (10, 20, 17)


def test_header_compiles_as_c_and_links_from_plain_c(tmp_path):
    """include/rst_b200.h is a C header (no C++ / CUDA / torch types) and the shared library links from a plain C program: the
    drop-in boundary any host language binds.  Without a GPU rst_create must fail cleanly with a message, not crash."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    header_dir = os.path.join(ROOT, "include")
    lib_dir = os.path.join(ROOT, "realtime_style_transfer_b200", "csrc")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c",
                           os.path.join(header_dir, "rst_b200.h")])
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "rst_b200.h"
int main(void) {
    rst_config cfg;
    rst_ctx* ctx = NULL;
    const unsigned char digits[9] = {'1','2','3','4','5','6','7','8','9'};
    memset(&cfg, 0, sizeof cfg);
    cfg.in_h = 64; cfg.in_w = 128; cfg.in_c = 17; cfg.out_h = 64; cfg.out_w = 128; cfg.bottleneck_res_y = 16;
    cfg.bottleneck_num_filters = 128; cfg.num_styles = 1; cfg.max_batch = 1; cfg.precision = RST_PRECISION_BF16;
    printf("%s\n", rst_version());
    printf("crc %08x\n", (unsigned)rst_host_crc32c(digits, 9));
    printf("dtype %d %d %d\n", RST_DTYPE_F32, RST_DTYPE_F16, RST_DTYPE_U8);
    int rc = rst_create(&cfg, 0, &ctx);
    printf("create rc=%d msg=%s\n", rc, rc ? rst_last_error(NULL) : "ok");
    if (!rc) rst_destroy(ctx);
    return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-I", header_dir, str(src), "-o", str(exe), "-L", lib_dir, "-lrst_sm100",
                           f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    assert lines[0].startswith("rst_b200") and lines[1] == "crc e3069283" and lines[2] == "dtype 0 1 2"
    assert lines[3].startswith("create rc=")
    import torch
    if not torch.cuda.is_available():
        assert "rc=0" not in lines[3] and "no CUDA device" in lines[3]


def test_reads_hand_assembled_exr_files(tmp_path):
    """EXR files assembled byte by byte in this test from the OpenEXR file-layout description (scan-line, single part) by an
    encoder that shares no code with dataloaders/exr.py: an uncompressed file with a HALF and a FLOAT channel and a data
    window that does not start at (0, 0), and a ZIP file (16-line blocks, byte interleave + delta predictor written as explicit
    Python loops, the last block short).  The Unreal capture writes such files; none is available offline."""
    import struct
    import zlib
    from realtime_style_transfer_b200.dataloaders import exr

    def attr(name, typ, payload):
        return name.encode() + b"\0" + typ.encode() + b"\0" + struct.pack("<i", len(payload)) + payload

    def build(path, planes, types, compression, window):
        x0, y0, x1, y1 = window
        w, h = x1 - x0 + 1, y1 - y0 + 1
        names = sorted(planes)                                            # channels are stored in alphabetical order
        chlist = b""
        for n in names:
            chlist += n.encode() + b"\0" + struct.pack("<iB3xii", {"HALF": 1, "FLOAT": 2}[types[n]], 0, 1, 1)
        chlist += b"\0"
        header = struct.pack("<iI", 20000630, 2)
        header += attr("channels", "chlist", chlist)
        header += attr("compression", "compression", bytes([compression]))
        header += attr("dataWindow", "box2i", struct.pack("<iiii", x0, y0, x1, y1))
        header += attr("displayWindow", "box2i", struct.pack("<iiii", 0, 0, x1, y1))
        header += attr("lineOrder", "lineOrder", b"\0")
        header += attr("pixelAspectRatio", "float", struct.pack("<f", 1.0))
        header += attr("screenWindowCenter", "v2f", struct.pack("<ff", 0.0, 0.0))
        header += attr("screenWindowWidth", "float", struct.pack("<f", 1.0))
        header += b"\0"
        lines = 16 if compression == 3 else 1
        blocks = []
        for by in range(0, h, lines):
            raw = b""
            for yy in range(by, min(by + lines, h)):
                for n in names:                                           # per scan line: every channel's row, one after the other
                    dt = "<f2" if types[n] == "HALF" else "<f4"
                    raw += planes[n][yy].astype(dt).tobytes()
            if compression == 3:
                half = (len(raw) + 1) // 2
                re = bytearray(len(raw))
                for i, b in enumerate(raw):                               # even bytes first, odd bytes second
                    re[(i // 2) if i % 2 == 0 else half + i // 2] = b
                enc = bytearray(len(re))
                prev = 0
                for i, b in enumerate(re):                                # delta predictor: d[i] = t[i] - t[i-1] + 128 (mod 256)
                    enc[i] = b if i == 0 else (b - prev + 128 + 256) & 0xFF
                    prev = b
                packed = zlib.compress(bytes(enc))
                payload = packed if len(packed) < len(raw) else raw      # a block is stored raw when compression does not help
            else:
                payload = raw
            blocks.append(struct.pack("<ii", y0 + by, len(payload)) + payload)
        table_pos = len(header)
        offsets, pos = [], table_pos + 8 * len(blocks)
        for b in blocks:
            offsets.append(pos)
            pos += len(b)
        path.write_bytes(header + struct.pack(f"<{len(blocks)}Q", *offsets) + b"".join(blocks))

    rng = np.random.default_rng(5)
    h, w = 21, 30                                                          # 21 rows: the second ZIP block has 5 lines
    half_vals = rng.uniform(0, 4, (h, w)).astype(np.float16).astype(np.float32)
    float_vals = rng.uniform(10, 1e4, (h, w)).astype(np.float32)
    build(tmp_path / "plain.exr", {"R": half_vals, "Z": float_vals}, {"R": "HALF", "Z": "FLOAT"}, 0, (3, 2, 3 + w - 1, 2 + h - 1))
    img = exr.load(tmp_path / "plain.exr")
    assert img.shape == (h, w) and list(img.channels()) == ["R", "Z"]
    assert np.array_equal(img.channel("R"), half_vals) and np.array_equal(img.channel("Z"), float_vals)
    smooth = np.tile(np.linspace(0, 1, w, dtype=np.float32), (h, 1)).astype(np.float16).astype(np.float32)
    build(tmp_path / "zip.exr", {"B": smooth * 0.5, "G": smooth, "R": half_vals}, {"B": "HALF", "G": "HALF", "R": "HALF"}, 3,
          (0, 0, w - 1, h - 1))
    img = exr.load(tmp_path / "zip.exr", keep_half=True)
    assert img.header["compression"] == "ZIP" and img.channel("G").dtype == np.float16
    assert np.array_equal(img.channel("R").astype(np.float32), half_vals)
    assert np.array_equal(img.channel("G").astype(np.float32), smooth) and np.array_equal(img.channel("B").astype(np.float32), smooth * 0.5)


def test_predictor_bottleneck_width_other_than_the_reference_default_is_rejected_early():
    """create_style_prediction_model(..., num_style_parameters=N): the native head has the reference default (100,
    stylePrediction.py:26) built in; another width must fail at construction, not with a shape error at the first forward."""
    stylePrediction.create_style_prediction_model((64, 128, 3), "DUMMY", 742, num_style_parameters=100)
    with pytest.raises(NotImplementedError):
        stylePrediction.create_style_prediction_model((64, 128, 3), "DUMMY", 742, num_style_parameters=64)
    with pytest.raises(NotImplementedError):
        PredictorPlan((64, 128, 3), "MOBILE_NET", 742, 128)


def test_checkpoint_duplicate_keras_names_prefer_the_inference_model(tmp_path):
    """A training checkpoint of the reference can hold two variables with the same Keras name (the frozen loss model's MobileNet
    next to the predictor's): entries below `loss_model` must never be loaded into the inference variables."""
    from realtime_style_transfer_b200 import checkpoint as ck
    model = stylePrediction.create_style_prediction_model((64, 128, 3), "MOBILE_NET", 50)
    good = np.full((3, 3, 3, 16), 2.0, np.float32)
    bad = np.full((3, 3, 3, 16), -7.0, np.float32)
    for order in (("loss_model/x", "inference_model/y"), ("a_loss_model/x", "z_inference/y")):
        ckpt = {
            f"{order[0] if 'loss' in order[0] else order[1]}/kernel/.ATTRIBUTES/VARIABLE_VALUE": {"value": bad, "full_name": "Conv/kernel"},
            f"{order[1] if 'loss' in order[0] else order[0]}/kernel/.ATTRIBUTES/VARIABLE_VALUE": {"value": good, "full_name": "Conv/kernel"},
        }
        assignment, _missing, _unused = ck.match_checkpoint_to_model(ckpt, model)
        assert np.array_equal(assignment["mobilenet/Conv/kernel"], good)


def test_reference_builder_helpers_exist_with_the_reference_shapes():
    """styleTransfer.py:95, :144, :188, :335: expand / residual_block / contract / _get_style_weight_mips are part of the module the
    reference exposes; the mirror provides them with the reference's names, variable layouts, initialiser ranges and shapes."""
    import torch
    from oracle import rst_oracle as O
    from realtime_style_transfer_b200.models import styleTransfer as ST
    c = ST.contract((480, 960, 17), 32, 9, 1, "start", seed=0)
    assert c.name == "contract_start" and c.output_shape == (480, 960, 32)
    assert [v.shape for v in c.weights] == [(9, 9, 17, 32), (32,), (32,), (32,), (32,), (32,)]
    assert abs(float(c.variables["conv/kernel"].std()) - 0.02) < 2e-3                     # N(0, 0.02)
    c2 = ST.contract((480, 960, 32), 16, 3, 2, "0")
    assert c2.output_shape == (240, 480, 16)
    r = ST.residual_block((120, 240, 32), 1, 128, 3, 1, "0", is_first=True, seed=0)
    assert r.name == "residual_block_0" and r.output_shape == (120, 240, 128) and r.num_style_parameters == 4 * 128
    k = r.variables["conv1/kernel"]
    assert k.shape == (3, 3, 128, 128) and k.min() >= 0.0 and k.max() <= 0.05               # U(0, 0.05)
    e = ST.expand((120, 240, 128), 2, 32, 3, 2, "0")
    assert e.name == "expand_0" and e.output_shape == (240, 480, 32) and e.num_style_parameters == 2 * 32
    assert e.variables["conv/kernel"].shape == (3, 3, 32, 128)                              # Conv2DTranspose: (kh, kw, out, in)
    with pytest.raises(ValueError):
        ST.expand((8, 8, 4), 1, 3, 3, 2, "x", activation="tanh")
    with pytest.raises(ValueError):
        r.set_weights(r.get_weights()[:2])
    w = np.random.default_rng(1).uniform(0, 1, (2, 20, 36, 2)).astype(np.float32)
    mips = ST._get_style_weight_mips(w, 3)
    ref = O.style_weight_mips(torch.as_tensor(w), 3)
    assert sorted(mips) == sorted(ref) == [4, 9, 18, 36]
    for width, level in mips.items():
        np.testing.assert_allclose(level, ref[width].numpy(), atol=1e-6)
