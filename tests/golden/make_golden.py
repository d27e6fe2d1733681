"""Regenerates the committed golden fixtures.  Run from the repo root: python tests/golden/make_golden.py

1. apply_style_weights_known_answer.npz -- the inputs and expected output of the reference's only
   numeric test (realtime_style_transfer/models/styleTransferTest.py:28-49).  TensorFlow cannot be
   imported here, so the reference test's input construction (its vertical-gradient helper, :12-24)
   and its explicit 4-deep expected-value loop (:41-47) are restated in numpy (oracle/naive_np.py).
2. tiny_transfer_fp64.npz -- a small end-to-end case (weights, inputs, fp64 oracle output) that
   freezes the oracle's behaviour so later edits to oracle/ cannot drift silently.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import naive_np, rst_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def known_answer():
    w = np.stack([naive_np.vertical_gradient((0, 1), (2, 10, 20)),
                  naive_np.vertical_gradient((1, 0), (2, 10, 20))], axis=-1)
    p = np.asarray([[[[10, 20, 30, 40, 50, 60], [70, 80, 90, 100, 110, 120]]]] * 2, np.float32)
    expected = naive_np.apply_style_weights_loop(w, p)
    np.savez(os.path.join(HERE, "apply_style_weights_known_answer.npz"), style_weights=w, style_params=p,
             expected=expected)


def tiny_transfer():
    spec = rst_oracle.TransferSpec((16, 32, 5), (16, 32, 3), 4, 8, 2)
    weights = rst_oracle.init_transfer_weights(spec, seed=11, trained_like=True)
    g = torch.Generator().manual_seed(12)
    content = torch.rand((2, 16, 32, 5), generator=g).numpy()
    params = (torch.rand((2, 2, spec.num_style_parameters), generator=g) + 0.5).numpy()
    sw = torch.rand((2, 16, 32, 1), generator=g).numpy()
    out = rst_oracle.transfer_forward(spec, weights, content, params, sw, dtype=torch.float64).numpy()
    np.savez_compressed(os.path.join(HERE, "tiny_transfer_fp64.npz"), content=content, style_params=params,
                        style_weights=sw, output=out, **{"w::" + k: v for k, v in weights.items()})


if __name__ == "__main__":
    known_answer()
    tiny_transfer()
    print("golden fixtures written to", HERE)
