"""GPU parity tests of the fp32 path, all through the C ABI, against the CPU oracle.

Tolerance (BASELINE.json north_star): max abs error <= 1e-4 on the FP32 path at the network output.
"""
import os

import numpy as np
import pytest
import torch

from oracle import rst_oracle as O
from realtime_style_transfer_b200 import _native, mixed_precision
from realtime_style_transfer_b200._plan import PredictorPlan, TransferPlan
from realtime_style_transfer_b200.models import stylePrediction, styleTransfer, styleTransferInferenceModel, styleLoss
from realtime_style_transfer_b200.shape_config import ShapeConfig

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
FP32_TOL = 1e-4


def dev(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def stream():
    return torch.cuda.current_stream().cuda_stream


# ---- registry -----------------------------------------------------------------------------------
def test_native_registry_matches_host_plan(cuda_device):
    plan = TransferPlan((64, 128, 17), (64, 128, 3), 16, 32, 2)
    pplan = PredictorPlan((64, 128, 3), "MOBILE_NET", plan.num_style_parameters)
    ctx = _native.NativeContext(in_shape=(64, 128, 17), out_shape=(64, 128, 3), bottleneck_res_y=16,
                                bottleneck_num_filters=32, num_styles=2, max_batch=1,
                                extractor=_native.EXTRACTOR_MOBILE_NET, style_shape=(64, 128))
    expect = dict(plan.variables())
    expect.update(pplan.variables())
    assert ctx.weight_specs() == expect
    assert ctx.num_style_params == plan.num_style_parameters
    assert ctx.lib.rst_num_contract_blocks(ctx.handle) == plan.num_contract_blocks
    assert ctx.lib.rst_num_expand_blocks(ctx.handle) == plan.num_expand_blocks
    ctx.close()


def test_create_rejects_bad_configs(cuda_device):
    with pytest.raises(_native.RstError):
        _native.NativeContext(in_shape=(0 + 7, 8, 3), out_shape=(7, 8, 3), bottleneck_res_y=120,
                              bottleneck_num_filters=8)          # cannot contract 7 rows to 120
    with pytest.raises(_native.RstError):
        _native.NativeContext(in_shape=(64, 128, 3), out_shape=(64, 128, 3), bottleneck_res_y=16,
                              bottleneck_num_filters=8, num_styles=3)
    ctx = _native.NativeContext(in_shape=(64, 128, 3), out_shape=(64, 128, 3), bottleneck_res_y=16,
                                bottleneck_num_filters=8)
    with pytest.raises(_native.RstError):   # forward before weights are committed
        ctx.transfer_forward_host(np.zeros((1, 64, 128, 3), np.float32), np.zeros((1, 1, ctx.num_style_params), np.float32))
    with pytest.raises(_native.RstError):
        ctx.set_weights({"contract_start/conv/kernel": np.zeros((3, 3, 3, 32), np.float32)})
    ctx.close()


# ---- single operators ---------------------------------------------------------------------------
@pytest.mark.parametrize("b,h,w,ci,co,k,s,transposed", [
    (2, 12, 20, 17, 32, 9, 1, False),     # stem
    (2, 12, 20, 32, 16, 3, 2, False),     # strided encoder, even size: pads (0,1)
    (1, 13, 21, 16, 32, 3, 2, False),     # odd size: pads (1,1)
    (2, 10, 14, 32, 128, 3, 1, False),    # residual
    (1, 10, 14, 128, 128, 3, 1, False),
    (2, 6, 9, 128, 32, 3, 2, True),       # decoder upsample
    (1, 12, 18, 32, 16, 3, 2, True),
    (2, 12, 20, 16, 3, 9, 1, True),       # head
    (1, 5, 7, 3, 1, 9, 5, False),         # DUMMY predictor conv
    (0, 8, 8, 4, 4, 3, 1, False),         # empty batch
    # images of 64x64 and more: thin stride-1 layers take the direct (shared-memory halo) kernel
    (2, 70, 100, 17, 32, 9, 1, False),    # stem, ragged 32x32 tiles, two channel chunks
    (1, 64, 96, 16, 3, 9, 1, True),       # head (input-gradient form walks the taps backwards)
    (2, 67, 64, 3, 16, 9, 1, False),      # head's input gradient: 3 -> 16
    (1, 65, 130, 64, 3, 3, 1, False),     # loss model conv1_1 input gradient: 64 -> 3, four channel chunks
    (1, 64, 64, 64, 3, 3, 1, True),
    # pointwise layers of the predictor: plain GEMM kernel (1x1, stride 1, Ci % 4 == 0, at least 4096 pixels)
    (2, 64, 64, 16, 72, 1, 1, False),
    (1, 64, 80, 72, 24, 1, 1, False),     # K = 72: three chunks, the last one partial
    (1, 70, 61, 8, 6, 1, 1, False),       # ragged pixel tile, Co % 4 != 0
])
def test_op_conv2d_fp32(cuda_device, b, h, w, ci, co, k, s, transposed):
    rng = np.random.default_rng(b * 1000 + h * 10 + k)
    x = rng.standard_normal((b, h, w, ci)).astype(np.float32)
    kern = (rng.standard_normal((k, k, co, ci) if transposed else (k, k, ci, co)) * 0.1).astype(np.float32)
    bias = rng.standard_normal(co).astype(np.float32)
    tx, tk, tb = torch.as_tensor(x), torch.as_tensor(kern), torch.as_tensor(bias)
    ref = (O.conv2d_transpose_same(tx, tk, tb, s) if transposed else O.conv2d_same(tx, tk, tb, s))
    ref = torch.relu(ref).numpy()
    d_y = torch.full(ref.shape, float("nan"), device=cuda_device)
    d_x, d_k, d_b = dev(x, cuda_device), dev(kern, cuda_device), dev(bias, cuda_device)
    _native.op_conv2d(d_x.data_ptr(), d_k.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), b, h, w, ci, co, k, k, s,
                      transposed, _native.ACT_RELU, _native.PRECISION_FP32, stream())
    got = d_y.cpu().numpy()
    assert got.shape == ref.shape
    if b:
        assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("b,h,w,ci,co,relu,transposed", [
    (2, 16, 32, 64, 64, True, False),        # VGG block1_conv2
    (1, 30, 60, 128, 256, True, False),      # two 128-column blocks, partial tiles (30 rows, 60 columns)
    (1, 9, 21, 512, 512, True, False),       # 16 channel groups, four column blocks
    (2, 12, 20, 64, 128, False, False),      # no activation
    (2, 12, 20, 128, 64, False, True),       # input gradient of a 64 -> 128 layer (= stride-1 Conv2DTranspose)
    (1, 15, 30, 512, 256, False, True),
])
def test_op_conv2d_tf32(cuda_device, b, h, w, ci, co, relu, transposed):
    """tcgen05 kind::tf32 3x3 conv against (a) the same conv on tf32-rounded operands in fp64: only accumulation order differs;
    (b) the exact fp32 conv: within tf32 operand rounding (2^-11 relative per operand)."""
    rng = np.random.default_rng(ci + co + h)
    x = rng.standard_normal((b, h, w, ci)).astype(np.float32)
    kern = (rng.standard_normal((3, 3, co, ci) if transposed else (3, 3, ci, co)) * np.sqrt(2.0 / (9 * ci))).astype(np.float32)
    bias = rng.standard_normal(co).astype(np.float32)

    def tf32(a):                                  # round to nearest, ties away (cvt.rna.tf32.f32)
        u = a.view(np.uint32).astype(np.uint64)
        return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)

    def conv(xx, kk):
        f = O.conv2d_transpose_same if transposed else O.conv2d_same
        y = f(torch.as_tensor(xx, dtype=torch.float64), torch.as_tensor(kk, dtype=torch.float64),
              torch.as_tensor(bias, dtype=torch.float64), 1)
        return (torch.relu(y) if relu else y).numpy()

    exact, rounded = conv(x, kern), conv(tf32(x), tf32(kern))
    d_y = torch.full(exact.shape, float("nan"), device=cuda_device)
    d_x, d_k, d_b = dev(x, cuda_device), dev(kern, cuda_device), dev(bias, cuda_device)
    _native.op_conv2d(d_x.data_ptr(), d_k.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), b, h, w, ci, co, 3, 3, 1, transposed,
                      _native.ACT_RELU if relu else _native.ACT_NONE, _native.PRECISION_TF32, stream())
    got = d_y.cpu().numpy().astype(np.float64)
    scale = np.abs(exact).max()
    e_rounded, e_exact = np.abs(got - rounded).max() / scale, np.abs(got - exact).max() / scale
    print(f"tf32 conv {ci}->{co}: vs tf32-rounded operands {e_rounded:.2e}, vs exact fp32 {e_exact:.2e}")
    assert e_rounded < 1e-3        # the MMA truncates x (only the weights are pre-rounded) and the output is stored in tf32
    assert e_exact < 2e-3


@pytest.mark.parametrize("b,h,w,ci,co,relu,transposed", [
    (2, 16, 32, 64, 64, True, False),
    (1, 30, 60, 128, 256, True, False),
    (1, 9, 21, 512, 512, True, False),
    (2, 12, 20, 128, 64, False, True),
    (1, 15, 30, 512, 256, False, True),
])
def test_op_conv2d_split_tf32(cuda_device, b, h, w, ci, co, relu, transposed):
    """Error-compensated split tf32 (three tensor-core products per fp32 product) with per-channel-group accumulation chains
    summed in fp32 by the epilogue: fp32-level accuracy (7e-7 measured).  With ONE long chain in the tensor core the same
    kernel is 10-60x less accurate (1e-5 .. 4.4e-5, growing with the number of MMAs): the accumulate step truncates."""
    rng = np.random.default_rng(ci + co + h + 1)
    x = rng.standard_normal((b, h, w, ci)).astype(np.float32)
    kern = (rng.standard_normal((3, 3, co, ci) if transposed else (3, 3, ci, co)) * np.sqrt(2.0 / (9 * ci))).astype(np.float32)
    bias = rng.standard_normal(co).astype(np.float32)
    f = O.conv2d_transpose_same if transposed else O.conv2d_same
    ref = f(torch.as_tensor(x, dtype=torch.float64), torch.as_tensor(kern, dtype=torch.float64), torch.as_tensor(bias, dtype=torch.float64), 1)
    ref = (torch.relu(ref) if relu else ref).numpy()
    d_y = torch.full(ref.shape, float("nan"), device=cuda_device)
    d_x, d_k, d_b = dev(x, cuda_device), dev(kern, cuda_device), dev(bias, cuda_device)
    _native.op_conv2d(d_x.data_ptr(), d_k.data_ptr(), d_b.data_ptr(), d_y.data_ptr(), b, h, w, ci, co, 3, 3, 1, transposed,
                      _native.ACT_RELU if relu else _native.ACT_NONE, _native.PRECISION_TF32X3, stream())
    err = np.abs(d_y.cpu().numpy() - ref).max() / np.abs(ref).max()
    print(f"split tf32 conv {ci}->{co}: max err relative to max |y| {err:.2e}")
    assert err < 5e-6


def test_op_apply_style_weights_known_answer(cuda_device):
    """The reference's own known-answer test, run through the CUDA operator."""
    g = np.load(os.path.join(GOLDEN, "apply_style_weights_known_answer.npz"))
    d_w, d_p = dev(g["style_weights"], cuda_device), dev(g["style_params"], cuda_device)
    d_o = torch.empty((2, 10, 20, 6), device=cuda_device)
    _native.op_apply_style_weights(d_w.data_ptr(), d_p.data_ptr(), d_o.data_ptr(), 2, 10, 20, 6, stream())
    np.testing.assert_almost_equal(d_o.cpu().numpy(), g["expected"], decimal=5)
    got = styleTransfer._apply_style_weights(g["style_weights"], g["style_params"])      # python surface
    np.testing.assert_almost_equal(got, g["expected"], decimal=5)


@pytest.mark.parametrize("styles", [1, 2])
def test_op_cin(cuda_device, styles):
    rng = np.random.default_rng(styles)
    b, h, w, f = 2, 9, 11, 12
    x = (rng.standard_normal((b, h, w, f)) * 2 + 5).astype(np.float32)
    params = rng.standard_normal((b, styles, 2 * f)).astype(np.float32)
    weights = rng.uniform(-0.5, 1.5, (b, h, w, 2)).astype(np.float32) if styles == 2 else None
    p4 = torch.as_tensor(params).unsqueeze(1)
    tw = torch.as_tensor(weights) if weights is not None else None
    ref = O.cin(torch.as_tensor(x), O.apply_style_weights(tw, p4[..., :f]), O.apply_style_weights(tw, p4[..., f:]))
    d_y = torch.empty((b, h, w, f), device=cuda_device)
    d_w = dev(weights, cuda_device) if weights is not None else None
    d_x, d_p = dev(x, cuda_device), dev(params, cuda_device)     # keep alive across the call
    _native.op_cin(d_x.data_ptr(), d_p.data_ptr(),
                   d_w.data_ptr() if d_w is not None else 0, d_y.data_ptr(), b, h, w, f, styles, _native.ACT_NONE,
                   stream())
    assert np.abs(d_y.cpu().numpy() - ref.numpy()).max() < 2e-5


@pytest.mark.parametrize("shape", [
    (2, 24, 40, 64),       # block1_conv2's channel count: M = 128 with a zero-filled upper half, N = 64
    (1, 30, 61, 128),      # ragged pixel count (1830 = 57 items + 6 pixels: the last TMA box is zero filled past the sample)
    (2, 16, 24, 256),      # 2 x 2 channel blocks (off-diagonal blocks load two operands)
    (1, 15, 30, 512),      # block4_conv3: 4 x 4 blocks
    (3, 120, 240, 128),    # long reduction (28800 pixels: several accumulator flushes per CTA), three samples
    (2, 9, 11, 24),        # not a tensor-core shape: generic CUDA-core kernel
])
def test_op_gram(cuda_device, shape):
    """get_gram_matrix_model (styleLoss.py:11-18).  Channel counts of the VGG16 style taps (64, 128, 256, 512) run on tcgen05
    (gram_tf32.cu: MN-major kind::tf32 operands, error-compensated hi/lo split = fp32-level accuracy); post-ReLU-like inputs
    (non-negative, so every Gram entry is a long sum of same-sign products: the worst case for truncating accumulation)."""
    x = np.abs(np.random.default_rng(0).standard_normal(shape)).astype(np.float32)
    ref = O.gram_matrix(torch.as_tensor(x, dtype=torch.float64)).numpy()
    got = styleLoss.gram_matrix(x)
    assert got.shape == ref.shape
    rel = np.abs(got - ref).max() / np.abs(ref).max()
    print(f"gram {shape}: max rel err {rel:.2e}")
    assert rel < 1e-5
    assert np.abs(got - got.transpose(0, 2, 1)).max() / np.abs(ref).max() < 1e-5
    # second restatement (numpy einsum in fp64) on the same input
    from oracle import naive_np as NP
    assert np.abs(NP.gram(x) - ref).max() / np.abs(ref).max() < 1e-12


def test_op_gram_tensor_core_matches_cuda_core(cuda_device, monkeypatch):
    """Same input through the tcgen05 kernel and through the CUDA-core kernel it replaces (RST_GRAM_CUDA_CORE=1)."""
    x = np.abs(np.random.default_rng(1).standard_normal((2, 60, 120, 256))).astype(np.float32)
    tc = styleLoss.gram_matrix(x)
    monkeypatch.setenv("RST_GRAM_CUDA_CORE", "1")
    cc = styleLoss.gram_matrix(x)
    ref = O.gram_matrix(torch.as_tensor(x, dtype=torch.float64)).numpy()
    e_tc, e_cc = np.abs(tc - ref).max() / np.abs(ref).max(), np.abs(cc - ref).max() / np.abs(ref).max()
    print(f"gram 256 channels: tensor cores {e_tc:.2e}, CUDA cores {e_cc:.2e}")
    assert e_tc < 1e-5 and e_cc < 1e-5


# ---- whole transfer network ---------------------------------------------------------------------
def run_transfer(shape_in, shape_out, res_y, filters, styles, batch, weights, content, params, sw=None, taps=None,
                 precision=_native.PRECISION_FP32):
    ctx = _native.NativeContext(in_shape=shape_in, out_shape=shape_out, bottleneck_res_y=res_y,
                                bottleneck_num_filters=filters, num_styles=styles, max_batch=batch, precision=precision)
    ctx.set_weights(weights)
    if taps is not None:
        ctx.enable_taps(True)
    out = ctx.transfer_forward_host(content, params, sw)
    got_taps = {}
    if taps is not None:
        for name, t in taps.items():
            if name.endswith("conv1/cin"):
                continue        # fused with the residual add natively; covered by the "residual_block_<b>" tap
            got_taps[name] = ctx.tap(name, t.shape)
    launches = ctx.last_launch_count()
    ctx.close()
    return out, got_taps, launches


@pytest.mark.parametrize("trained_like", [False, True])
@pytest.mark.parametrize("styles", [1, 2])
def test_transfer_small_layer_by_layer(cuda_device, styles, trained_like):
    shape_in, shape_out = (32, 64, 17), (32, 64, 3)
    spec = O.TransferSpec(shape_in, shape_out, 8, 16, styles)
    weights = O.init_transfer_weights(spec, seed=1, trained_like=trained_like)
    cfg = ShapeConfig(num_channels=17)
    content = O.synthetic_content(2, 32, 64, cfg.channels, seed=0, unit_depth=True)
    params = np.random.default_rng(2).uniform(0.2, 1.2, (2, styles, spec.num_style_parameters)).astype(np.float32)
    sw = O.synthetic_style_weights(2, 32, 64) if styles == 2 else None
    ref_taps = {}
    ref = O.transfer_forward(spec, weights, content, params, sw, taps=ref_taps).numpy()
    out, taps, launches = run_transfer(shape_in, shape_out, 8, 16, styles, 2, weights, content, params, sw, ref_taps)
    assert launches > 0
    assert len(taps) >= 25
    for name, got in taps.items():
        t = ref_taps[name]
        scale = max(1.0, float(t.abs().max()))
        assert np.abs(got - t.numpy()).max() <= 2e-4 * scale, name
    assert np.abs(out - ref).max() <= FP32_TOL


def test_transfer_golden_fixture(cuda_device):
    g = np.load(os.path.join(GOLDEN, "tiny_transfer_fp64.npz"))
    weights = {k[3:]: g[k] for k in g.files if k.startswith("w::")}
    out, _, _ = run_transfer((16, 32, 5), (16, 32, 3), 4, 8, 2, 2, weights, g["content"], g["style_params"],
                             g["style_weights"])
    assert np.abs(out - g["output"]).max() <= FP32_TOL


def test_transfer_reference_test_geometries(cuda_device):
    """Shapes of the reference's own model tests, scaled down 8x: super-resolution with 4 expand blocks
    (styleTransferInferenceModelTest.py:18-44) and 3-contract/4-expand (styleTransferTrainingModelTest.py:15-44)."""
    for shape_in, shape_out, res_y, f, styles in [((60, 120, 3), (240, 480, 3), 15, 16, 2),
                                                  ((48, 96, 3), (96, 192, 3), 6, 4, 1)]:
        spec = O.TransferSpec(shape_in, shape_out, res_y, f, styles)
        weights = O.init_transfer_weights(spec, seed=4, trained_like=True)
        rng = np.random.default_rng(5)
        content = rng.uniform(0, 1, (1,) + shape_in).astype(np.float32)
        params = rng.uniform(0.2, 1.2, (1, styles, spec.num_style_parameters)).astype(np.float32)
        sw = O.synthetic_style_weights(1, shape_out[0], shape_out[1]) if styles == 2 else None
        ref = O.transfer_forward(spec, weights, content, params, sw).numpy()
        out, _, _ = run_transfer(shape_in, shape_out, res_y, f, styles, 1, weights, content, params, sw)
        assert out.shape == (1,) + shape_out
        assert np.abs(out - ref).max() <= FP32_TOL


def test_config1_rst_960_120_32_3_fp32(cuda_device):
    """BASELINE.json configs[0]: single-style forward, batch 1, synthetic RGB frame, FP32."""
    cfg = ShapeConfig.from_spec("rst-960-120-32-3")
    spec = O.TransferSpec(cfg.input_shape["content"], cfg.output_shape, 120, 32, 1)
    weights = O.init_transfer_weights(spec, seed=1)
    content = O.synthetic_content(1, 480, 960, [("FinalImage", 3)], seed=0)
    params = np.random.default_rng(1).uniform(0.3, 1.0, (1, 1, 742)).astype(np.float32)
    ref = O.transfer_forward(spec, weights, content, params).numpy()
    out, _, _ = run_transfer(cfg.input_shape["content"], cfg.output_shape, 120, 32, 1, 1, weights, content, params)
    err = np.abs(out - ref).max()
    print("config1 fp32 max abs err", err)
    assert err <= FP32_TOL
    assert out.min() > 0 and out.max() < 1


@pytest.mark.parametrize("styles", [1, 2])
def test_reference_builder_blocks_match_the_oracle(cuda_device, styles):
    """contract / residual_block / expand (styleTransfer.py:95-205) as stand-alone callables against the oracle's primitives."""
    from realtime_style_transfer_b200.models import styleTransfer as ST
    rng = np.random.default_rng(7 + styles)
    t = torch.as_tensor
    x = rng.standard_normal((2, 12, 20, 8)).astype(np.float32)
    # contract: ReLU(BN(ReLU(conv + b))), stride 2, non-trivial moving statistics
    c = ST.contract((12, 20, 8), 16, 3, 2, "0", seed=3)
    c.variables["conv/bias"] = rng.normal(0, 0.1, 16).astype(np.float32)
    c.variables["bn/moving_mean"] = rng.normal(0, 0.05, 16).astype(np.float32)
    c.variables["bn/moving_variance"] = rng.uniform(0.5, 1.5, 16).astype(np.float32)
    c.variables["bn/gamma"] = rng.uniform(0.5, 1.5, 16).astype(np.float32)
    v = {k: t(a) for k, a in c.variables.items()}
    ref = torch.relu(O.batchnorm(torch.relu(O.conv2d_same(t(x), v["conv/kernel"], v["conv/bias"], 2)), v["bn/gamma"], v["bn/beta"],
                                 v["bn/moving_mean"], v["bn/moving_variance"]))
    got = c(x)
    assert got.shape == (2,) + c.output_shape and np.abs(got - ref.numpy()).max() < 1e-5
    # residual block with skip, and expand with sigmoid
    sw = rng.uniform(0, 1, (2, 12, 20, 2)).astype(np.float32) if styles == 2 else None
    sw2 = rng.uniform(0, 1, (2, 24, 40, 2)).astype(np.float32) if styles == 2 else None

    def cin_ref(y, params, weights, f):
        p = t(params)
        w = t(weights) if weights is not None else None
        return O.cin(y, O.apply_style_weights(w, p[..., :f]), O.apply_style_weights(w, p[..., f:]))

    r = ST.residual_block((12, 20, 8), styles, 8, 3, 1, "1", seed=4)
    params = rng.uniform(0.3, 1.2, (2, 1, styles, 32)).astype(np.float32)
    inputs = {"content": x, "style_params": params}
    if styles == 2:
        inputs["style_weights"] = sw
    v = {k: t(a) for k, a in r.variables.items()}
    fx = torch.relu(cin_ref(torch.relu(O.conv2d_same(t(x), v["conv0/kernel"], v["conv0/bias"], 1)), params[..., :16], sw, 8))
    fx = cin_ref(torch.relu(O.conv2d_same(fx, v["conv1/kernel"], v["conv1/bias"], 1)), params[..., 16:], sw, 8)
    assert np.abs(r(inputs) - (t(x) + fx).numpy()).max() < 2e-5
    e = ST.expand((12, 20, 8), styles, 4, 3, 2, "0", activation="sigmoid", seed=5)
    eparams = rng.uniform(0.3, 1.2, (2, 1, styles, 8)).astype(np.float32)
    einputs = {"content": x, "style_params": eparams}
    if styles == 2:
        einputs["style_weights"] = sw2
    v = {k: t(a) for k, a in e.variables.items()}
    ref = torch.sigmoid(cin_ref(O.conv2d_transpose_same(t(x), v["conv/kernel"], v["conv/bias"], 2), eparams, sw2, 4))
    got = e(einputs)
    assert got.shape == (2,) + e.output_shape and np.abs(got - ref.numpy()).max() < 2e-5


@pytest.mark.parametrize("filters", [32, 128])
def test_fp32_residual_blocks_run_on_the_tensor_cores(cuda_device, monkeypatch, filters):
    """RST_PRECISION_FP32: the residual 3x3 convolutions are split-tf32 tcgen05 GEMMs (rst_api.cu::fp32_tensor_commit); the
    CUDA-core kernels (RST_FP32_TENSOR=0) give the same image to fp32 rounding, and both meet the fp32 bar against the oracle."""
    shape_in, shape_out = (64, 128, 17), (64, 128, 3)
    spec = O.TransferSpec(shape_in, shape_out, 16, filters, 1)
    weights = O.init_transfer_weights(spec, seed=1, trained_like=True)
    content = O.synthetic_content(2, 64, 128, ShapeConfig(num_channels=17).channels, seed=5, unit_depth=True)
    params = np.random.default_rng(2).uniform(0.3, 1.2, (2, 1, spec.num_style_parameters)).astype(np.float32)
    ref = O.transfer_forward(spec, weights, content, params).numpy()
    outs, groups = {}, {}
    for mode in ("1", "0"):
        monkeypatch.setenv("RST_FP32_TENSOR", mode)
        ctx = _native.NativeContext(in_shape=shape_in, out_shape=shape_out, bottleneck_res_y=16, bottleneck_num_filters=filters,
                                    num_styles=1, max_batch=2)
        ctx.set_weights(weights)
        ctx.profile(True)
        outs[mode] = ctx.transfer_forward_host(content, params)
        groups[mode] = ctx.profile_groups()
        ctx.close()
    assert groups["1"]["conv_tf32x3"][1] == 10 and "conv_tf32x3" not in groups["0"]        # ten residual convs
    assert groups["1"]["conv_fp32"][1] == groups["0"]["conv_fp32"][1] - 10                 # stem / strided / transposed layers stay
    print("fp32 tensor vs CUDA-core max abs", np.abs(outs["1"] - outs["0"]).max(), "vs oracle", np.abs(outs["1"] - ref).max())
    assert np.abs(outs["1"] - outs["0"]).max() <= 5e-5                                      # measured 3e-6 (F = 32) / 1.5e-5 (F = 128)
    assert np.abs(outs["1"] - ref).max() <= FP32_TOL and np.abs(outs["0"] - ref).max() <= FP32_TOL


def test_frames_are_independent_and_deterministic(cuda_device):
    """Size-independent properties: instance norm is per sample, so a batch equals its frames run alone."""
    shape_in, shape_out = (64, 128, 17), (64, 128, 3)
    spec = O.TransferSpec(shape_in, shape_out, 16, 32, 1)
    weights = O.init_transfer_weights(spec, seed=1)
    content = O.synthetic_content(3, 64, 128, ShapeConfig(num_channels=17).channels, seed=3)
    params = np.random.default_rng(1).uniform(0.3, 1.0, (3, 1, spec.num_style_parameters)).astype(np.float32)
    ctx = _native.NativeContext(in_shape=shape_in, out_shape=shape_out, bottleneck_res_y=16, bottleneck_num_filters=32,
                                num_styles=1, max_batch=3)
    ctx.set_weights(weights)
    full = ctx.transfer_forward_host(content, params)
    again = ctx.transfer_forward_host(content, params)
    np.testing.assert_array_equal(full, again)
    for i in range(3):
        single = ctx.transfer_forward_host(content[i:i + 1], params[i:i + 1])
        np.testing.assert_array_equal(single[0], full[i])
    assert ctx.transfer_forward_host(content[:0], params[:0]).shape == (0, 64, 128, 3)
    ctx.close()


# ---- predictor and the composed inference model ---------------------------------------------------
@pytest.mark.parametrize("extractor", ["DUMMY", "MOBILE_NET"])
def test_style_predictor(cuda_device, extractor):
    w = O.init_predictor_weights(extractor, 742, seed=2)
    style = np.random.default_rng(0).uniform(0, 1, (2, 96, 160, 3)).astype(np.float32)
    ref = O.predictor_forward(extractor, w, style).numpy()
    model = stylePrediction.create_style_prediction_model((96, 160, 3), extractor, 742)
    model.set_weights(w)
    got = model.predict(style)
    assert got.shape == (2, 742)
    assert np.abs(got - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    got_dev = model(torch.as_tensor(style).to(cuda_device)).cpu().numpy()      # device-pointer entry
    np.testing.assert_allclose(got_dev, got, atol=1e-6)
    model.close()


@pytest.mark.parametrize("styles", [1, 2])
def test_inference_model_through_reference_surface(cuda_device, styles):
    """make_style_transfer_inference_model(...).inference.predict(element) vs the oracle
    (predict_using_checkpoint.py:57-66, :99)."""
    in_shape, out_shape = (64, 128, 17), (64, 128, 3)
    models = styleTransferInferenceModel.make_style_transfer_inference_model(
        num_styles=styles,
        style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model(
            out_shape, stylePrediction.StyleFeatureExtractor.MOBILE_NET, n),
        style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model(in_shape, out_shape, 16, 32,
                                                                                      styles))
    spec = O.TransferSpec(in_shape, out_shape, 16, 32, styles)
    tw = O.init_transfer_weights(spec, seed=1, trained_like=True)
    pw = O.init_predictor_weights("MOBILE_NET", spec.num_style_parameters, seed=2)
    models.transfer.set_weights(tw)
    models.style_predictor.set_weights(pw)
    for m in (models.style_predictor, models.transfer, models.inference):
        m.trainable = False
        m.compile(run_eagerly=False)
    rng = np.random.default_rng(7)
    element = {"content": O.synthetic_content(2, 64, 128, ShapeConfig(num_channels=17).channels, seed=1, unit_depth=True),
               "style": rng.uniform(0, 1, (2, styles) + out_shape).astype(np.float32)}
    if styles == 2:
        element["style_weights"] = O.synthetic_style_weights(2, 64, 128)
    ref = O.inference_forward(spec, tw, "MOBILE_NET", pw, element["content"], element["style"],
                              element.get("style_weights")).numpy()
    got = models.inference.predict(element)
    assert got.shape == (2,) + out_shape and got.dtype == np.float32
    assert np.abs(got - ref).max() <= FP32_TOL
    # the video loop's split: style params once, then transfer.predict per frame (predict_video_using_checkpoint.py:77-96)
    sp = np.stack([models.style_predictor.predict(element["style"][:, s]) for s in range(styles)], axis=1)
    el2 = {"content": element["content"], "style_params": sp}
    if styles == 2:
        el2["style_weights"] = element["style_weights"]
    got2 = models.transfer.predict(el2, batch_size=1, verbose=0)
    assert np.abs(got2 - ref).max() <= FP32_TOL
    with pytest.raises(ValueError):
        models.transfer.predict({"content": element["content"][:, :32], "style_params": sp})
    models.inference.close(); models.transfer.close(); models.style_predictor.close()


def test_reference_inference_model_test_at_its_own_size(cuda_device, tmp_path):
    """realtime_style_transfer/models/styleTransferInferenceModelTest.py at the reference's OWN sizes (the largest geometry the
    reference exercises): content (480,960,3), output (1920,3840,3) = four expand blocks, two styles with a (1920,3840,1) weight
    map, DUMMY predictor.  test_output_shape_matches, test_inference (all-zero element, as the reference feeds it) and
    test_save_transfer_model (here: the ONNX file of export.py); plus what the reference could not check -- the numbers,
    against the oracle on a random element."""
    num_styles = 2
    in_shape, out_shape = (480, 960, 3), (1920, 3840, 3)
    models = styleTransferInferenceModel.make_style_transfer_inference_model(
        num_styles=num_styles,
        style_transfer_factory_func=lambda: styleTransfer.create_style_transfer_model(
            input_shape=in_shape, output_shape=out_shape, bottleneck_res_y=120, bottleneck_num_filters=128,
            num_styles=num_styles, name="StyleTransferTestModel"),
        style_predictor_factory_func=lambda n: stylePrediction.create_style_prediction_model(
            in_shape, stylePrediction.StyleFeatureExtractor.DUMMY, n),
        name="StyleTransferInferenceTestModel")
    models.inference.compile()
    assert models.inference.output_shape == (None,) + out_shape                       # test_output_shape_matches
    shapes = {"style": (num_styles,) + in_shape, "style_weights": (1920, 3840, num_styles - 1), "content": in_shape}
    y = models.inference.predict({name: np.zeros((1,) + shape, np.float32) for name, shape in shapes.items()})
    assert y.shape == (1,) + out_shape and np.isfinite(y).all()                       # test_inference
    path = models.transfer.save(str(tmp_path / "transfer.onnx"))                      # test_save_transfer_model
    assert (tmp_path / "transfer.onnx").stat().st_size > 1_000_000 and path.endswith(".onnx")
    # numbers: random element against the oracle (fp32 path, 1e-4)
    spec = O.TransferSpec(in_shape, out_shape, 120, 128, num_styles)
    assert (spec.n_contract, spec.n_expand) == (2, 4)
    tw = O.init_transfer_weights(spec, seed=8)
    pw = O.init_predictor_weights("DUMMY", spec.num_style_parameters, seed=9)
    models.transfer.set_weights(tw)
    models.style_predictor.set_weights(pw)
    rng = np.random.default_rng(10)
    element = {"content": rng.uniform(0, 1, (1,) + in_shape).astype(np.float32),
               "style": rng.uniform(0, 1, (1, num_styles) + in_shape).astype(np.float32),
               "style_weights": O.synthetic_style_weights(1, 1920, 3840)}
    ref = O.inference_forward(spec, tw, "DUMMY", pw, element["content"], element["style"], element["style_weights"]).numpy()
    got = models.inference.predict(element)
    print("reference test geometry 480x960x3 -> 1920x3840x3, 2 styles: max abs err", float(np.abs(got - ref).max()))
    assert got.shape == ref.shape and np.abs(got - ref).max() <= FP32_TOL
    models.inference.close(); models.transfer.close(); models.style_predictor.close()
