"""CPU tests of the oracle itself: golden vectors, an independent numpy restatement, shapes."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import naive_np, rst_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_known_answer_apply_style_weights():
    """The reference's only numeric test (styleTransferTest.py:28-49): 7-decimal agreement."""
    g = np.load(os.path.join(GOLDEN, "apply_style_weights_known_answer.npz"))
    got = O.apply_style_weights(torch.as_tensor(g["style_weights"]), torch.as_tensor(g["style_params"])).numpy()
    assert got.shape == (2, 10, 20, 6)
    np.testing.assert_almost_equal(got, g["expected"], decimal=7)
    # weights in the reference test deliberately leave [0,1]: no clamping / normalisation may happen
    assert g["style_weights"].max() > 1.0 and g["style_weights"].min() < 0.0


def test_apply_style_weights_passthrough_single_style():
    p = torch.rand(2, 1, 1, 6)
    assert O.apply_style_weights(None, p) is p


def test_golden_tiny_transfer_is_stable():
    g = np.load(os.path.join(GOLDEN, "tiny_transfer_fp64.npz"))
    spec = O.TransferSpec((16, 32, 5), (16, 32, 3), 4, 8, 2)
    weights = {k[3:]: g[k] for k in g.files if k.startswith("w::")}
    out = O.transfer_forward(spec, weights, g["content"], g["style_params"], g["style_weights"],
                             dtype=torch.float64).numpy()
    np.testing.assert_allclose(out, g["output"], rtol=0, atol=1e-12)
    out32 = O.transfer_forward(spec, weights, g["content"], g["style_params"], g["style_weights"]).numpy()
    assert np.abs(out32 - g["output"]).max() < 1e-5


@pytest.mark.parametrize("h,w,k,s", [(6, 8, 3, 1), (6, 8, 3, 2), (7, 9, 3, 2), (12, 10, 9, 1), (10, 10, 9, 5),
                                     (8, 6, 5, 2)])
def test_conv_same_matches_naive(h, w, k, s):
    rng = np.random.default_rng(h * 100 + k)
    x = rng.standard_normal((2, h, w, 3)).astype(np.float32)
    kern = rng.standard_normal((k, k, 3, 4)).astype(np.float32)
    b = rng.standard_normal(4).astype(np.float32)
    ref = naive_np.conv2d_same(x, kern, b, s)
    got = O.conv2d_same(torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(kern, dtype=torch.float64),
                        torch.as_tensor(b, dtype=torch.float64), s).numpy()
    assert got.shape == ref.shape == (2, -(-h // s), -(-w // s), 4)
    np.testing.assert_allclose(got, ref, atol=1e-10)


@pytest.mark.parametrize("h,w,k,s", [(5, 6, 3, 2), (4, 4, 9, 1), (6, 5, 3, 1), (3, 4, 5, 2)])
def test_conv_transpose_same_matches_naive(h, w, k, s):
    rng = np.random.default_rng(h * 10 + k)
    x = rng.standard_normal((2, h, w, 4)).astype(np.float32)
    kern = rng.standard_normal((k, k, 3, 4)).astype(np.float32)     # (kh,kw,out,in)
    b = rng.standard_normal(3).astype(np.float32)
    ref = naive_np.conv2d_transpose_same(x, kern, b, s)
    got = O.conv2d_transpose_same(torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(kern, dtype=torch.float64),
                                  torch.as_tensor(b, dtype=torch.float64), s).numpy()
    assert got.shape == ref.shape == (2, h * s, w * s, 3)
    np.testing.assert_allclose(got, ref, atol=1e-10)


def test_conv_transpose_is_gradient_of_conv():
    """TF defines Conv2DTranspose as the input-gradient of Conv2D with the same padding."""
    torch.manual_seed(0)
    x = torch.randn(1, 8, 10, 3, dtype=torch.float64, requires_grad=True)     # forward-conv input (2H,2W,out)
    kern = torch.randn(3, 3, 3, 5, dtype=torch.float64)                       # fwd conv: in=3, out=5
    y = O.conv2d_same(x, kern, None, 2)                                       # (1,4,5,5)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    # Conv2DTranspose kernel layout (kh,kw,out,in) == forward kernel (kh,kw,in_fwd,out_fwd)
    got = O.conv2d_transpose_same(gy, kern, None, 2)
    np.testing.assert_allclose(got.numpy(), gx.numpy(), atol=1e-10)


def test_cin_matches_naive_and_normalises():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((2, 6, 7, 5)) * 3 + 10).astype(np.float32)
    scale = rng.standard_normal((2, 1, 1, 5)).astype(np.float32)
    bias = rng.standard_normal((2, 1, 1, 5)).astype(np.float32)
    ref = naive_np.cin(x, scale, bias)
    got = O.cin(torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(scale, dtype=torch.float64),
                torch.as_tensor(bias, dtype=torch.float64)).numpy()
    np.testing.assert_allclose(got, ref, atol=1e-9)
    unit = O.cin(torch.as_tensor(x, dtype=torch.float64), torch.ones(1, dtype=torch.float64),
                 torch.zeros(1, dtype=torch.float64))
    assert abs(float(unit.mean())) < 1e-9 and abs(float(unit.var(dim=(1, 2), unbiased=False).mean()) - 1) < 1e-4


def test_gram_matches_naive():
    x = np.random.default_rng(4).standard_normal((2, 5, 6, 7)).astype(np.float32)
    np.testing.assert_allclose(O.gram_matrix(torch.as_tensor(x, dtype=torch.float64)).numpy(), naive_np.gram(x),
                               atol=1e-10)


def test_block_counts_keep_reference_float_expression():
    # standard geometry and the reference tests' two other geometries
    assert O.block_counts(480, 480, 120) == (2, 2)
    assert O.block_counts(480, 1920, 120) == (2, 4)
    assert O.block_counts(240, 480, 30) == (3, 4)
    assert math.log2(480) - math.log2(120) < 2.0     # lands below the integer: ceil() is what makes it 2


@pytest.mark.parametrize("spec_args,expected_p", [
    (((480, 960, 17), (480, 960, 3), 120, 128, 1), 2662),
    (((480, 960, 3), (480, 960, 3), 120, 32, 1), 742),
    (((480, 960, 3), (1920, 3840, 3), 120, 128, 2), 20 * 128 + 2 * (32 + 16 + 8 + 4 + 3)),
    (((240, 480, 3), (480, 960, 3), 30, 4, 1), 20 * 4 + 2 * (32 + 16 + 8 + 4 + 3)),
])
def test_num_style_parameters(spec_args, expected_p):
    assert O.TransferSpec(*spec_args).num_style_parameters == expected_p


def test_parameter_counts_match_survey():
    spec = O.TransferSpec((480, 960, 17), (480, 960, 3), 120, 128, 1)
    n = sum(int(np.prod(s)) for s in spec.weight_shapes().values())
    assert n == 1464019 + 320
    shapes = O.predictor_weight_shapes("MOBILE_NET", 2662)
    body = sum(int(np.prod(s)) for k, s in shapes.items() if k.startswith("mobilenet/"))
    assert body == 939120          # what Keras reports for MobileNetV3Small(include_top=False)
    assert sum(int(np.prod(s)) for s in shapes.values()) - body == 326562


def test_transfer_forward_shapes_and_range():
    spec = O.TransferSpec((32, 64, 3), (64, 128, 3), 8, 4, 2)
    assert (spec.n_contract, spec.n_expand) == (2, 3)
    w = O.init_transfer_weights(spec, seed=1)
    content = O.synthetic_content(2, 32, 64, [("FinalImage", 3)])
    params = np.random.default_rng(0).uniform(0.2, 1.0, (2, 2, spec.num_style_parameters)).astype(np.float32)
    sw = O.synthetic_style_weights(2, 64, 128)
    taps = {}
    y = O.transfer_forward(spec, w, content, params, sw, taps=taps)
    assert tuple(y.shape) == (2, 64, 128, 3) and y.dtype == torch.float32
    assert float(y.min()) > 0 and float(y.max()) < 1
    assert tuple(taps["contract_1"].shape) == (2, 8, 16, 32)
    assert tuple(taps["expand_0"].shape) == (2, 16, 32, 32)


def test_predictor_shapes():
    w = O.init_predictor_weights("MOBILE_NET", 742)
    x = np.random.default_rng(0).uniform(0, 1, (1, 64, 96, 3)).astype(np.float32)
    W = {k: torch.as_tensor(v) for k, v in w.items()}
    feat = O.mobilenet_v3_small(W, torch.as_tensor(x) * 2 - 1)
    assert tuple(feat.shape) == (1, 2, 3, 576)
    assert tuple(O.predictor_forward("MOBILE_NET", w, x).shape) == (1, 742)
    wd = O.init_predictor_weights("DUMMY", 50)
    assert tuple(O.predictor_forward("DUMMY", wd, x).shape) == (1, 50)
    with pytest.raises(ValueError):
        O.predictor_forward("NOPE", wd, x)


def test_loss_is_per_sample_vector_and_zero_at_identity():
    vgg = O.init_vgg16_weights()
    img = np.random.default_rng(1).uniform(0, 1, (2, 32, 32, 3)).astype(np.float32)
    out = O.style_loss_vgg(vgg, img, img, img[:, None])
    assert tuple(out["loss"].shape) == (2,)
    assert float(out["feature_loss"].abs().max()) == 0 and float(out["style_loss"].abs().max()) == 0
    tv = O.total_variation(torch.as_tensor(img))
    np.testing.assert_allclose(out["loss"].numpy(), 0.1 * tv.numpy(), rtol=1e-6)
    with pytest.raises(AssertionError):
        O.style_loss_vgg(vgg, img, img, np.stack([img, img], axis=1))


def test_bf16_emulation_mode_rounds_activations_and_bounds_the_error():
    """emulate_bf16 is the fp32 graph with bf16 storage: identical when off, small but non-zero deviation when on, every
    stored activation exactly representable in bf16."""
    import torch
    spec = O.TransferSpec((32, 64, 17), (32, 64, 3), 8, 32, 1)
    w = O.init_transfer_weights(spec, seed=3)
    from realtime_style_transfer_b200.shape_config import ShapeConfig
    c = O.synthetic_content(1, 32, 64, ShapeConfig(num_channels=17).channels, seed=1, unit_depth=True)
    p = np.random.default_rng(0).uniform(0.3, 1.2, (1, 1, spec.num_style_parameters)).astype(np.float32)
    ref = O.transfer_forward(spec, w, c, p).numpy()
    same = O.transfer_forward(spec, w, c, p, emulate_bf16=False).numpy()
    assert np.array_equal(ref, same)
    taps = {}
    emu = O.transfer_forward(spec, w, c, p, dtype=torch.float64, emulate_bf16=True, taps=taps).numpy()
    err = np.abs(emu - ref)
    assert 0 < err.max() < 5e-2 and np.sqrt((err ** 2).sum() / (ref ** 2).sum()) < 1e-2
    for name in ("contract_start", "residual_block_0/conv0/relu", "residual_block_2", "expand_0/conv", "expand_1"):
        t = taps[name]
        assert torch.equal(t, t.to(torch.bfloat16).to(t.dtype)), name
    assert not torch.equal(taps["expand_last/conv"], taps["expand_last/conv"].to(torch.bfloat16).to(torch.float64))


# ---- second restatements (oracle/naive_np.py) of the pieces that had only one: MobileNetV3Small, training BN, RMSprop ----
def test_mobilenet_v3_small_against_the_loop_restatement():
    """Whole style predictor (stride-2 blocks with correct_pad on even AND odd sizes, squeeze-excite, hard-swish, residual
    adds, the two linear heads) on a 40x72 style image: torch oracle vs the numpy loop restatement."""
    from oracle import naive_np as NP
    w = O.init_predictor_weights("MOBILE_NET", 50, seed=4)
    style = np.random.default_rng(5).uniform(0, 1, (1, 40, 72, 3)).astype(np.float32)     # 40 -> 20 -> 10 -> 5 (odd) -> 3
    ref = NP.style_predictor({k: v.astype(np.float64) for k, v in w.items()}, style)
    import torch
    got = O.predictor_forward("MOBILE_NET", w, style, dtype=torch.float64).numpy()
    assert got.shape == ref.shape == (1, 50)
    assert np.abs(got - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())
    got32 = O.predictor_forward("MOBILE_NET", w, style).numpy()
    assert np.abs(got32 - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_inverted_res_block_stride2_padding_cases():
    """correct_pad: even sizes pad (k//2 - 1, k//2), odd sizes (k//2, k//2), for the 3x3 and 5x5 depthwise kernels."""
    from oracle import naive_np as NP
    import torch
    assert NP.correct_pad(8, 3) == (0, 1) and NP.correct_pad(9, 3) == (1, 1)
    assert NP.correct_pad(8, 5) == (1, 2) and NP.correct_pad(9, 5) == (2, 2)
    for size, k in ((8, 3), (9, 3), (8, 5), (9, 5)):
        assert O._correct_pad(size, k) == NP.correct_pad(size, k)
        rng = np.random.default_rng(size * k)
        x = rng.standard_normal((1, size, size + 1, 4))
        kern = rng.standard_normal((k, k, 4, 1))
        pt, pb = NP.correct_pad(size, k)
        pl, pr = NP.correct_pad(size + 1, k)
        ref = NP.depthwise_valid(NP.zero_pad(x, pt, pb, pl, pr), kern, 2)
        got = O.depthwise_conv2d(torch.as_tensor(x), torch.as_tensor(kern), 2, (pt, pb, pl, pr)).numpy()
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-12


def test_training_batchnorm_and_moving_statistics_against_the_loop_restatement():
    from oracle import naive_np as NP
    import torch
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 5, 7, 6)) * 2 + 1
    gamma, beta = rng.uniform(0.5, 1.5, 6), rng.standard_normal(6)
    mm, mv = rng.standard_normal(6), rng.uniform(0.5, 2, 6)
    for momentum in (0.99, 0.999):
        y, nm, nv = NP.batchnorm_training(x, gamma, beta, mm, mv, eps=1e-3, momentum=momentum)
        out = []
        got = O.batchnorm(torch.as_tensor(x), torch.as_tensor(gamma), torch.as_tensor(beta), torch.as_tensor(mm),
                          torch.as_tensor(mv), training=True, moving_out=out, momentum=momentum).numpy()
        assert np.abs(got - y).max() < 1e-12
        assert np.abs(out[0][0].numpy() - nm).max() < 1e-12 and np.abs(out[0][1].numpy() - nv).max() < 1e-12
    inf = NP.batchnorm_inference(x, gamma, beta, mm, mv)
    got = O.batchnorm(torch.as_tensor(x), torch.as_tensor(gamma), torch.as_tensor(beta), torch.as_tensor(mm),
                      torch.as_tensor(mv)).numpy()
    assert np.abs(got - inf).max() < 1e-12


def test_rmsprop_against_the_loop_restatement():
    from oracle import naive_np as NP
    import torch
    rng = np.random.default_rng(1)
    w = {"a": rng.standard_normal((3, 4)).astype(np.float32), "b": rng.standard_normal(5).astype(np.float32)}
    slots = {k: np.zeros_like(v, np.float64) for k, v in w.items()}
    ref_w = {k: v.astype(np.float64) for k, v in w.items()}
    ref_s = {k: np.zeros_like(v, np.float64) for k, v in w.items()}
    for step in range(3):
        g = {k: rng.standard_normal(v.shape) * 10.0 ** (step - 2) for k, v in w.items()}
        w, slots = O.rmsprop_update(w, {k: torch.as_tensor(v) for k, v in g.items()}, slots)
        for k in ref_w:
            ref_w[k], ref_s[k] = NP.rmsprop_step(ref_w[k], g[k], ref_s[k])
            assert np.abs(slots[k] - ref_s[k]).max() < 1e-15
            assert np.abs(w[k] - ref_w[k]).max() < 1e-6           # the oracle stores float32 variables, as Keras does
            ref_w[k] = w[k].astype(np.float64)


def test_total_variation_and_maxpool_against_the_loop_restatement():
    from oracle import naive_np as NP
    import torch
    rng = np.random.default_rng(2)
    img = rng.uniform(0, 1, (2, 6, 9, 3))
    assert np.abs(O.total_variation(torch.as_tensor(img)).numpy() - NP.total_variation(img)).max() < 1e-12
    x = rng.standard_normal((1, 6, 8, 5))
    got = torch.nn.functional.max_pool2d(torch.as_tensor(x).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).numpy()
    assert np.array_equal(got, NP.max_pool2(x))
