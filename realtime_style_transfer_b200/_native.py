"""ctypes binding of librst_sm100.so (include/rst_b200.h).

The product path has no CPU fallback: if the shared library is missing or no B200 is visible,
every compute entry point raises.  Loading the library itself needs no GPU (symbol checks in the
CPU test-suite rely on that).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "librst_sm100.so")

PRECISION_FP32 = 0
PRECISION_BF16 = 1
PRECISION_TF32 = 2
PRECISION_TF32X3 = 3
EXTRACTOR_NONE, EXTRACTOR_DUMMY, EXTRACTOR_MOBILE_NET = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SIGMOID = 0, 1, 4
DTYPE_F32, DTYPE_F16, DTYPE_U8 = 0, 1, 2          # enum rst_dtype


def content_dtype_code(dtype) -> int:
    """numpy dtype of a G-buffer array -> rst_dtype (float32, or float16 = the reduced-byte ingest)."""
    dtype = np.dtype(dtype)
    if dtype == np.float16:
        return DTYPE_F16
    if dtype == np.float32:
        return DTYPE_F32
    raise ValueError(f"content dtype must be float32 or float16, got {dtype}")


def out_dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.uint8:
        return DTYPE_U8
    if dtype == np.float32:
        return DTYPE_F32
    raise ValueError(f"output dtype must be float32 or uint8, got {dtype}")


class RstError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[rst {code}] {message}")
        self.code = code


class RstConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "in_h", "in_w", "in_c", "out_h", "out_w", "bottleneck_res_y", "bottleneck_num_filters",
        "num_styles", "max_batch", "precision", "extractor", "style_h", "style_w", "predictor_num_params")]


_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_vp = C.c_void_p

# name -> (restype, argtypes); every symbol include/rst_b200.h declares
SIGNATURES = {
    "rst_version": (C.c_char_p, []),
    "rst_host_crc32c": (C.c_uint32, [_vp, C.c_uint64]),
    "rst_host_alloc": (C.c_int, [C.POINTER(_vp), C.c_uint64, C.c_int]),
    "rst_host_free": (C.c_int, [_vp]),
    "rst_create": (C.c_int, [C.POINTER(RstConfig), C.c_int, C.POINTER(_vp)]),
    "rst_destroy": (C.c_int, [_vp]),
    "rst_last_error": (C.c_char_p, [_vp]),
    "rst_num_style_params": (C.c_int, [_vp]),
    "rst_num_contract_blocks": (C.c_int, [_vp]),
    "rst_num_expand_blocks": (C.c_int, [_vp]),
    "rst_weight_count": (C.c_int, [_vp]),
    "rst_weight_name": (C.c_char_p, [_vp, C.c_int]),
    "rst_weight_shape": (C.c_int, [_vp, C.c_int, _i64p, C.POINTER(C.c_int)]),
    "rst_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, _i64p, C.c_int]),
    "rst_get_weight": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64]),
    "rst_commit_weights": (C.c_int, [_vp]),
    "rst_transfer_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rst_transfer_forward_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int]),
    "rst_transfer_submit_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _i64p]),
    "rst_transfer_wait": (C.c_int, [_vp, C.c_int64]),
    "rst_transfer_forward_typed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "rst_transfer_forward_host_typed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int]),
    "rst_transfer_submit_host_typed": (C.c_int, [_vp, _vp, C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _i64p]),
    "rst_predict_style": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "rst_predict_style_host": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "rst_inference_forward_host": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int]),
    "rst_debug_enable_taps": (C.c_int, [_vp, C.c_int]),
    "rst_debug_tap": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64, _i64p]),
    "rst_last_launch_count": (C.c_int64, [_vp]),
    "rst_profile_enable": (C.c_int, [_vp, C.c_int]),
    "rst_profile_reset": (C.c_int, [_vp]),
    "rst_profile_get": (C.c_int, [_vp, C.c_char_p, C.POINTER(C.c_double), _i64p]),
    "rst_profile_group_count": (C.c_int, [_vp]),
    "rst_profile_group_name": (C.c_char_p, [_vp, C.c_int]),
    "rst_op_last_error": (C.c_char_p, []),
    "rst_op_conv2d": (C.c_int, [_vp, _vp, _vp, _vp] + [C.c_int] * 11 + [_vp]),
    "rst_op_cin": (C.c_int, [_vp, _vp, _vp, _vp] + [C.c_int] * 6 + [_vp]),
    "rst_op_apply_style_weights": (C.c_int, [_vp, _vp, _vp] + [C.c_int] * 4 + [_vp]),
    "rst_op_gram": (C.c_int, [_vp, _vp] + [C.c_int] * 4 + [_vp]),
    "rst_loss_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "rst_loss_destroy": (C.c_int, [_vp]),
    "rst_loss_last_error": (C.c_char_p, [_vp]),
    "rst_loss_set_weight": (C.c_int, [_vp, C.c_char_p, _vp, _i64p, C.c_int]),
    "rst_loss_commit": (C.c_int, [_vp]),
    "rst_loss_set_factors": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float]),
    "rst_loss_set_math": (C.c_int, [_vp, C.c_int]),
    "rst_loss_forward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp]),
    "rst_loss_backward": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp]),
    "rst_train_create": (C.c_int, [C.POINTER(RstConfig), C.c_int, C.POINTER(_vp)]),
    "rst_train_destroy": (C.c_int, [_vp]),
    "rst_train_last_error": (C.c_char_p, [_vp]),
    "rst_train_set_math": (C.c_int, [_vp, C.c_int]),
    "rst_train_model": (_vp, [_vp]),
    "rst_train_loss": (_vp, [_vp]),
    "rst_train_forward_backward": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, C.c_int]),
    "rst_train_stream": (_vp, [_vp]),
    "rst_train_gradients": (_vp, [_vp]),
    "rst_train_num_gradient_elements": (C.c_int64, [_vp]),
    "rst_train_variable_range": (C.c_int, [_vp, C.c_char_p, _i64p, _i64p]),
    "rst_train_apply_gradients": (C.c_int, [_vp, C.c_float, C.c_float, C.c_float]),
    "rst_train_sync_weights": (C.c_int, [_vp]),
    "rst_train_read_gradient": (C.c_int, [_vp, C.c_char_p, _vp, C.c_int64]),
    "rst_train_read_prediction": (C.c_int, [_vp, _vp, C.c_int64]),
    "rst_train_prediction": (_vp, [_vp]),
    "rst_train_debug_read": (C.c_int, [_vp, C.c_char_p, C.c_int, _vp, C.c_int64, _i64p]),
}

_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RstError(-1, f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header/library drift, surface it loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _host_f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _host_content(a) -> np.ndarray:
    """float16 G-buffers stay float16 (half the PCIe bytes); everything else becomes float32 like the reference's tensors."""
    if isinstance(a, np.ndarray) and a.dtype == np.float16:
        return np.ascontiguousarray(a)
    return _host_f32(a)


def _ptr(a):
    """numpy array / int / None -> void*."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(_vp)
    return _vp(int(a))


class PinnedArray:
    """A page-locked host array from rst_host_alloc (numpy view in ``.array``); free() or garbage collection releases it.
    write_combined=True for frame buffers the host only writes before they are uploaded."""

    def __init__(self, shape, dtype=np.float32, write_combined: bool = False):
        self.lib = load_library()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        ptr = _vp()
        rc = self.lib.rst_host_alloc(C.byref(ptr), n, int(write_combined))
        if rc != 0:
            raise RstError(rc, (self.lib.rst_last_error(None) or b"").decode())
        self.ptr = ptr
        buf = (C.c_uint8 * max(n, 1)).from_address(ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if getattr(self, "ptr", None):
            self.array = None
            self.lib.rst_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class NativeContext:
    """Owns one rst_ctx (one GPU).  Thin, no arithmetic on this side."""

    def __init__(self, *, in_shape=None, out_shape=None, bottleneck_res_y=0, bottleneck_num_filters=0,
                 num_styles=1, max_batch=1, precision=PRECISION_FP32, extractor=EXTRACTOR_NONE,
                 style_shape=None, predictor_num_params=0, device: int = 0):
        self.lib = load_library()
        cfg = RstConfig()
        if in_shape is not None:
            cfg.in_h, cfg.in_w, cfg.in_c = (int(v) for v in in_shape)
            cfg.out_h, cfg.out_w = int(out_shape[0]), int(out_shape[1])
        cfg.bottleneck_res_y = int(bottleneck_res_y)
        cfg.bottleneck_num_filters = int(bottleneck_num_filters)
        cfg.num_styles = int(num_styles)
        cfg.max_batch = int(max_batch)
        cfg.precision = int(precision)
        cfg.extractor = int(extractor)
        if style_shape is not None:
            cfg.style_h, cfg.style_w = int(style_shape[0]), int(style_shape[1])
        cfg.predictor_num_params = int(predictor_num_params)
        self.cfg = cfg
        self.device = device
        self._owns = True
        handle = _vp()
        rc = self.lib.rst_create(C.byref(cfg), device, C.byref(handle))
        if rc != 0:
            raise RstError(rc, (self.lib.rst_last_error(None) or b"").decode())
        self.handle = handle
        self.num_style_params = self.lib.rst_num_style_params(self.handle)

    @classmethod
    def _view(cls, handle, cfg: "RstConfig", device: int) -> "NativeContext":
        """Non-owning wrapper of a context that belongs to a trainer."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.cfg, self.device, self.handle, self._owns = cfg, device, _vp(handle), False
        self.num_style_params = self.lib.rst_num_style_params(self.handle)
        return self

    # -- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != 0:
            raise RstError(rc, (self.lib.rst_last_error(self.handle) or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            if getattr(self, "_owns", True):
                self.lib.rst_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- weights ----------------------------------------------------------------------------
    def weight_specs(self) -> Dict[str, tuple]:
        out = {}
        shape = (C.c_int64 * 4)()
        nd = C.c_int()
        for i in range(self.lib.rst_weight_count(self.handle)):
            name = self.lib.rst_weight_name(self.handle, i).decode()
            self._check(self.lib.rst_weight_shape(self.handle, i, shape, C.byref(nd)))
            out[name] = tuple(int(shape[k]) for k in range(nd.value))
        return out

    def set_weights(self, weights: Dict[str, np.ndarray], commit: bool = True):
        for name, value in weights.items():
            arr = _host_f32(value)
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            self._check(self.lib.rst_set_weight(self.handle, name.encode(), _ptr(arr), shape, arr.ndim))
        if commit:
            self._check(self.lib.rst_commit_weights(self.handle))

    def get_weight(self, name: str, shape: Sequence[int]) -> np.ndarray:
        out = np.empty(tuple(shape), np.float32)
        self._check(self.lib.rst_get_weight(self.handle, name.encode(), _ptr(out), out.size))
        return out

    # -- hot path ---------------------------------------------------------------------------
    def transfer_forward_host(self, content, style_params, style_weights=None, out_dtype=np.float32) -> np.ndarray:
        """content float32 or float16 (kept as is); out_dtype float32, or uint8 = trunc(255 * y) computed on the device."""
        content = _host_content(content)
        style_params = _host_f32(style_params)
        sw = _host_f32(style_weights) if style_weights is not None else None
        b = content.shape[0]
        out = np.empty((b, self.cfg.out_h, self.cfg.out_w, 3), np.dtype(out_dtype))
        self._check(self.lib.rst_transfer_forward_host_typed(self.handle, _ptr(content), content_dtype_code(content.dtype),
                                                             _ptr(style_params), _ptr(sw), _ptr(out), out_dtype_code(out.dtype), b))
        return out

    def transfer_forward_device(self, d_content: int, d_style_params: int, d_style_weights: Optional[int],
                                d_out: int, batch: int, stream: int = 0, content_dtype: int = DTYPE_F32,
                                out_dtype: int = DTYPE_F32):
        self._check(self.lib.rst_transfer_forward_typed(self.handle, _vp(d_content), content_dtype, _vp(d_style_params),
                                                        _vp(d_style_weights) if d_style_weights else None, _vp(d_out), out_dtype,
                                                        batch, _vp(stream) if stream else None))

    def transfer_submit_host(self, content: np.ndarray, style_params: np.ndarray, style_weights, out: np.ndarray) -> int:
        """Asynchronous submit; the caller keeps the (ideally pinned) arrays alive until transfer_wait(ticket).
        content: float32 or float16; out: float32 or uint8 (the array dtypes select the entry point's element types)."""
        for a in (content, style_params, out):
            assert a.flags["C_CONTIGUOUS"]
        assert style_params.dtype == np.float32 and (style_weights is None or style_weights.dtype == np.float32)
        ticket = C.c_int64()
        self._check(self.lib.rst_transfer_submit_host_typed(self.handle, _ptr(content), content_dtype_code(content.dtype),
                                                            _ptr(style_params), _ptr(style_weights), _ptr(out),
                                                            out_dtype_code(out.dtype), content.shape[0], C.byref(ticket)))
        return int(ticket.value)

    def transfer_wait(self, ticket: int):
        self._check(self.lib.rst_transfer_wait(self.handle, ticket))

    def predict_style_host(self, style) -> np.ndarray:
        style = _host_f32(style)
        b = style.shape[0]
        out = np.empty((b, self.num_style_params), np.float32)
        self._check(self.lib.rst_predict_style_host(self.handle, _ptr(style), _ptr(out), b))
        return out

    def predict_style_device(self, d_style: int, d_params: int, batch: int, stream: int = 0):
        self._check(self.lib.rst_predict_style(self.handle, _vp(d_style), _vp(d_params), batch,
                                               _vp(stream) if stream else None))

    def inference_forward_host(self, content, style, style_weights=None) -> np.ndarray:
        content = _host_f32(content)
        style = _host_f32(style)
        sw = _host_f32(style_weights) if style_weights is not None else None
        b = content.shape[0]
        out = np.empty((b, self.cfg.out_h, self.cfg.out_w, 3), np.float32)
        self._check(self.lib.rst_inference_forward_host(self.handle, _ptr(content), _ptr(style), _ptr(sw),
                                                        _ptr(out), b))
        return out

    # -- debug / accounting -----------------------------------------------------------------
    def enable_taps(self, enable: bool = True):
        self._check(self.lib.rst_debug_enable_taps(self.handle, int(enable)))

    def tap(self, name: str, shape: Sequence[int]) -> np.ndarray:
        n = C.c_int64()
        self._check(self.lib.rst_debug_tap(self.handle, name.encode(), None, 0, C.byref(n)))
        out = np.empty(int(n.value), np.float32)
        self._check(self.lib.rst_debug_tap(self.handle, name.encode(), _ptr(out), out.size, C.byref(n)))
        return out.reshape(tuple(shape))

    def last_launch_count(self) -> int:
        return int(self.lib.rst_last_launch_count(self.handle))

    def profile(self, enable: bool):
        self._check(self.lib.rst_profile_enable(self.handle, int(enable)))

    def profile_reset(self):
        self._check(self.lib.rst_profile_reset(self.handle))

    def profile_groups(self) -> Dict[str, tuple]:
        out = {}
        n = self.lib.rst_profile_group_count(self.handle)
        for i in range(n):
            name = self.lib.rst_profile_group_name(self.handle, i)
            ms, cnt = C.c_double(), C.c_int64()
            self._check(self.lib.rst_profile_get(self.handle, name, C.byref(ms), C.byref(cnt)))
            out[name.decode()] = (ms.value, int(cnt.value))
        return out


# ---- stand-alone operators on device pointers (torch tensors give .data_ptr()) -----------------
def _op_check(rc: int):
    if rc != 0:
        raise RstError(rc, (load_library().rst_op_last_error() or b"").decode())


def op_conv2d(d_x, d_kernel, d_bias, d_y, batch, h, w, ci, co, kh, kw, stride, transposed=False, act=ACT_NONE,
              precision=PRECISION_FP32, stream=0):
    _op_check(load_library().rst_op_conv2d(_vp(d_x), _vp(d_kernel), _vp(d_bias) if d_bias else None, _vp(d_y),
                                           batch, h, w, ci, co, kh, kw, stride, int(transposed), act, precision,
                                           _vp(stream) if stream else None))


def op_cin(d_x, d_params, d_weights, d_y, batch, h, w, f, num_styles, act=ACT_NONE, stream=0):
    _op_check(load_library().rst_op_cin(_vp(d_x), _vp(d_params), _vp(d_weights) if d_weights else None, _vp(d_y),
                                        batch, h, w, f, num_styles, act, _vp(stream) if stream else None))


def op_apply_style_weights(d_weights, d_params, d_out, batch, h, w, f, stream=0):
    _op_check(load_library().rst_op_apply_style_weights(_vp(d_weights), _vp(d_params), _vp(d_out), batch, h, w, f,
                                                        _vp(stream) if stream else None))


def op_gram(d_x, d_gram, batch, h, w, c, stream=0):
    _op_check(load_library().rst_op_gram(_vp(d_x), _vp(d_gram), batch, h, w, c, _vp(stream) if stream else None))


class NativeLoss:
    """Owns one rst_loss (VGG16 content / Gram style / total-variation loss, forward and backward)."""

    def __init__(self, h: int, w: int, max_batch: int, device: int = 0):
        self.lib = load_library()
        handle = _vp()
        rc = self.lib.rst_loss_create(int(h), int(w), int(max_batch), device, C.byref(handle))
        if rc != 0:
            raise RstError(rc, (self.lib.rst_loss_last_error(None) or b"").decode())
        self.handle, self.h, self.w, self.max_batch, self.device = handle, h, w, max_batch, device
        self._owns = True

    @classmethod
    def _view(cls, handle, h: int, w: int, max_batch: int, device: int) -> "NativeLoss":
        self = cls.__new__(cls)
        self.lib = load_library()
        self.handle, self.h, self.w, self.max_batch, self.device, self._owns = _vp(handle), h, w, max_batch, device, False
        return self

    def _check(self, rc):
        if rc != 0:
            raise RstError(rc, (self.lib.rst_loss_last_error(self.handle) or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            if getattr(self, "_owns", True):
                self.lib.rst_loss_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_weights(self, weights: Dict[str, np.ndarray]):
        for name, value in weights.items():
            arr = _host_f32(value)
            shape = (C.c_int64 * arr.ndim)(*arr.shape)
            self._check(self.lib.rst_loss_set_weight(self.handle, name.encode(), _ptr(arr), shape, arr.ndim))
        self._check(self.lib.rst_loss_commit(self.handle))

    def set_math(self, precision: int):
        """PRECISION_FP32 (default: split tf32, fp32-level accuracy) or PRECISION_TF32 (plain tf32) for the VGG16 convolutions;
        call before set_weights (commit)."""
        self._check(self.lib.rst_loss_set_math(self.handle, int(precision)))

    def set_factors(self, content: float, style: float, tv: float):
        self._check(self.lib.rst_loss_set_factors(self.handle, content, style, tv))

    def forward(self, d_pred: int, d_content: int, d_style: int, d_losses: int, batch: int, stream: int = 0):
        self._check(self.lib.rst_loss_forward(self.handle, _vp(d_pred), _vp(d_content), _vp(d_style), _vp(d_losses), batch,
                                              _vp(stream) if stream else None))

    def backward(self, d_pred: int, d_grad: int, batch: int, stream: int = 0):
        self._check(self.lib.rst_loss_backward(self.handle, _vp(d_pred), _vp(d_grad), batch, _vp(stream) if stream else None))


class NativeTrainer:
    """Owns one rst_trainer: predictor + transfer network in training mode, VGG loss model, RMSprop state."""

    def __init__(self, *, in_shape, out_shape, bottleneck_res_y, bottleneck_num_filters, max_batch, extractor, style_shape,
                 device: int = 0):
        self.lib = load_library()
        cfg = RstConfig()
        cfg.in_h, cfg.in_w, cfg.in_c = (int(v) for v in in_shape)
        cfg.out_h, cfg.out_w = int(out_shape[0]), int(out_shape[1])
        cfg.bottleneck_res_y, cfg.bottleneck_num_filters = int(bottleneck_res_y), int(bottleneck_num_filters)
        cfg.num_styles, cfg.max_batch, cfg.precision, cfg.extractor = 1, int(max_batch), PRECISION_FP32, int(extractor)
        cfg.style_h, cfg.style_w = int(style_shape[0]), int(style_shape[1])
        self.cfg, self.device = cfg, device
        handle = _vp()
        rc = self.lib.rst_train_create(C.byref(cfg), device, C.byref(handle))
        if rc != 0:
            raise RstError(rc, (self.lib.rst_train_last_error(None) or b"").decode())
        self.handle = handle
        self.model = NativeContext._view(self.lib.rst_train_model(handle), cfg, device)
        self.loss = NativeLoss._view(self.lib.rst_train_loss(handle), cfg.out_h, cfg.out_w, cfg.max_batch, device)
        self.num_gradient_elements = int(self.lib.rst_train_num_gradient_elements(handle))

    def _check(self, rc):
        if rc != 0:
            raise RstError(rc, (self.lib.rst_train_last_error(self.handle) or b"").decode())

    def close(self):
        if getattr(self, "handle", None):
            self.model.handle = None
            self.loss.handle = None
            self.lib.rst_train_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_math(self, precision: int):
        """PRECISION_TF32: tf32 tensor cores for the residual trunk and the loss model; call before the weights are set."""
        self._check(self.lib.rst_train_set_math(self.handle, int(precision)))

    def forward_backward(self, d_content: int, d_style: int, d_gt_content: int, d_gt_style: int, d_losses: int, batch: int):
        self._check(self.lib.rst_train_forward_backward(self.handle, _vp(d_content), _vp(d_style), _vp(d_gt_content),
                                                        _vp(d_gt_style), _vp(d_losses), batch))

    def stream_ptr(self) -> int:
        """cudaStream_t of the trainer's kernels (wrap with torch.cuda.ExternalStream to record timing events on it)."""
        return int(self.lib.rst_train_stream(self.handle) or 0)

    def gradients_ptr(self) -> int:
        return int(self.lib.rst_train_gradients(self.handle) or 0)

    def variable_range(self, name: str):
        off, n = C.c_int64(), C.c_int64()
        self._check(self.lib.rst_train_variable_range(self.handle, name.encode(), C.byref(off), C.byref(n)))
        return int(off.value), int(n.value)

    def apply_gradients(self, learning_rate: float = 1e-3, rho: float = 0.9, epsilon: float = 1e-7):
        self._check(self.lib.rst_train_apply_gradients(self.handle, learning_rate, rho, epsilon))

    def sync_weights(self):
        self._check(self.lib.rst_train_sync_weights(self.handle))

    def read_gradient(self, name: str, shape: Sequence[int]) -> np.ndarray:
        out = np.empty(tuple(shape), np.float32)
        self._check(self.lib.rst_train_read_gradient(self.handle, name.encode(), _ptr(out), out.size))
        return out

    def debug_read(self, name: str, want_grad: bool = False) -> np.ndarray:
        n = C.c_int64()
        self._check(self.lib.rst_train_debug_read(self.handle, name.encode(), int(want_grad), None, 0, C.byref(n)))
        out = np.empty((int(n.value),), np.float32)
        self._check(self.lib.rst_train_debug_read(self.handle, name.encode(), int(want_grad), _ptr(out), out.size, None))
        return out

    def read_prediction(self, batch: int) -> np.ndarray:
        out = np.empty((batch, self.cfg.out_h, self.cfg.out_w, 3), np.float32)
        self._check(self.lib.rst_train_read_prediction(self.handle, _ptr(out), out.size))
        return out
