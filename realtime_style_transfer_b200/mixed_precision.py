"""Global numeric policy, in the spirit of ``tf.keras.mixed_precision.set_global_policy`` which the
reference keeps as a commented-out switch (train_network.py:26, predict_using_checkpoint.py:40).

'float32'          -> CUDA-core fp32 kernels (bar: max abs err <= 1e-4 vs the fp32 oracle)
'mixed_bfloat16'   -> tcgen05 bf16 operands, fp32 accumulation and statistics (bar: <= 2e-2 relative)
"""
from ._native import PRECISION_BF16, PRECISION_FP32

_policy = "float32"


def set_global_policy(name: str):
    global _policy
    if name not in ("float32", "mixed_bfloat16"):
        raise ValueError(f"unknown policy {name!r}; expected 'float32' or 'mixed_bfloat16'")
    _policy = name


def global_policy() -> str:
    return _policy


def native_precision() -> int:
    return PRECISION_BF16 if _policy == "mixed_bfloat16" else PRECISION_FP32
