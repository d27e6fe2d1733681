"""Checkpoint ingestion: TF2 object-based checkpoint -> named numpy arrays.

Not built yet; ``load_weights`` accepts the ``.npz`` written by ``save_weights`` today.
"""
from __future__ import annotations


def read_checkpoint_variables(prefix: str):
    raise FileNotFoundError(f"{prefix}: TF2 checkpoint bundles are not readable yet; use a .npz written by save_weights")


def match_checkpoint_to_model(ckpt_vars, model):
    raise NotImplementedError
