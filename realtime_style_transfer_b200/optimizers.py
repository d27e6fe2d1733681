"""Optimizer configuration objects -- the slice of ``tf.keras.optimizers`` the reference uses (train_network.py:102).

The arithmetic runs in librst_sm100.so (rst_train_apply_gradients); these objects only carry hyper-parameters.
"""
from __future__ import annotations


class RMSprop:
    """``tf.keras.optimizers.RMSprop()`` of TF 2.9: learning_rate 1e-3, rho 0.9, momentum 0, epsilon 1e-7, centered False."""

    def __init__(self, learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7, centered=False, name="RMSprop", **kwargs):
        if momentum != 0.0 or centered:
            raise NotImplementedError("only the plain RMSprop the reference trains with (momentum=0, centered=False) is built")
        self.learning_rate, self.rho, self.momentum, self.epsilon, self.centered, self.name = \
            float(learning_rate), float(rho), 0.0, float(epsilon), False, name
        self.iterations = 0

    def get_config(self):
        return {"name": self.name, "learning_rate": self.learning_rate, "rho": self.rho, "momentum": self.momentum,
                "epsilon": self.epsilon, "centered": self.centered}
