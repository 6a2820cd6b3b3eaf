"""ORACLE: fannypack.nn.resblocks (Appendix A.8).  y = act(block2(act(block1(x))) + x).
Used inside every hot MLP of the reference, e.g. ref: crossmodal/push_models/layers.py:20-24,
crossmodal/push_models/dynamics.py:25-31.  state_dict keys: ``block1.*``, ``block2.*``."""
import torch.nn as nn


def _activation(name):
    table = {"relu": nn.ReLU, "selu": nn.SELU, "none": nn.Identity}
    return table[name]


class _Residual(nn.Module):
    def __init__(self, block1, block2, activation):
        super().__init__()
        self.block1 = block1
        self.block2 = block2
        self.activation = _activation(activation)()

    def forward(self, x):
        hidden = self.activation(self.block1(x))
        return self.activation(self.block2(hidden) + x)


class Linear(_Residual):
    def __init__(self, units, bottleneck_units=None, activation="relu"):
        inner = units if bottleneck_units is None else bottleneck_units
        super().__init__(nn.Linear(units, inner), nn.Linear(inner, units), activation)


class Conv2d(_Residual):
    def __init__(self, channels, bottleneck_channels=None, kernel_size=3, activation="relu"):
        inner = channels if bottleneck_channels is None else bottleneck_channels
        pad = kernel_size // 2
        super().__init__(
            nn.Conv2d(channels, inner, kernel_size, padding=pad),
            nn.Conv2d(inner, channels, kernel_size, padding=pad),
            activation,
        )
