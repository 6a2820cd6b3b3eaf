"""Residual blocks with fannypack's parameter names (``block1``, ``block2``), so reference
checkpoints load unchanged.  y = relu(block2(relu(block1(x))) + x)
(used at e.g. ref: crossmodal/push_models/layers.py:23, crossmodal/push_models/dynamics.py:27-29)."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class Linear(nn.Module):
    def __init__(self, units: int, bottleneck_units: int = None, activation: str = "relu"):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("only the relu residual block is used by the filtering models")
        hidden = bottleneck_units or units
        self.block1 = nn.Linear(units, hidden)
        self.block2 = nn.Linear(hidden, units)
        self.activation = nn.ReLU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.relu(self.block2(F.relu(self.block1(x))) + x)


class Conv2d(nn.Module):
    def __init__(self, channels: int, bottleneck_channels: int = None, kernel_size: int = 3, activation: str = "relu"):
        super().__init__()
        if activation != "relu":
            raise NotImplementedError("only the relu residual block is used by the filtering models")
        hidden = bottleneck_channels or channels
        self.block1 = nn.Conv2d(channels, hidden, kernel_size, padding=kernel_size // 2)
        self.block2 = nn.Conv2d(hidden, channels, kernel_size, padding=kernel_size // 2)
        self.activation = nn.ReLU()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return F.relu(self.block2(F.relu(self.block1(x))) + x)
