"""Deterministic synthetic parameters and inputs (no datasets / checkpoints are reachable).

Everything is derived from numpy ``default_rng`` streams keyed by tensor *name*, so any module
tree with the reference's ``state_dict`` key names (the reference's own classes, the oracle port,
this package's drop-in classes) receives bit-identical parameters.  Shapes follow the reference's
datasets: ref: crossmodal/tasks/_push.py:187,201,214,249 (image (T,N,32,32), gripper_pos (T,N,3),
gripper_sensors (T,N,7), controls (T,N,7)); inputs are z-normalised there (:364-399).
"""
import zlib

import numpy as np
import torch


def _stream(seed: int, name: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def fill_parameters(module: torch.nn.Module, seed: int = 0, scale: float = 1.0) -> torch.nn.Module:
    """U(-b, b), b = scale / sqrt(fan_in) for every trainable tensor; fixed noise parameters
    (``Q_scale_tril*``, ``requires_grad=False``) are left at their constructor values."""
    with torch.no_grad():
        for name, p in module.named_parameters():
            if not p.requires_grad or "Q_scale_tril" in name:
                continue
            if p.dim() > 1:
                fan_in = int(np.prod(p.shape[1:]))
            else:
                # bias: use the fan-in of the sibling weight when there is one
                sibling = dict(module.named_parameters()).get(name[: -len("bias")] + "weight")
                fan_in = int(np.prod(sibling.shape[1:])) if sibling is not None else p.numel()
            bound = scale / np.sqrt(max(fan_in, 1))
            values = _stream(seed, name).uniform(-bound, bound, size=tuple(p.shape)).astype(np.float32)
            p.copy_(torch.from_numpy(values))
    return module


def synthetic_trajectories(T: int, N: int, state_dim: int, seed: int = 0, blackout_fraction: float = 0.0):
    """(states (T,N,sd), observations dict of (T,N,...), controls (T,N,7)), all fp32 CPU tensors."""
    rng = np.random.default_rng([seed, 7919])
    states = rng.standard_normal((T, N, state_dim)).astype(np.float32)
    controls = rng.standard_normal((T, N, 7)).astype(np.float32)
    # last control column is a z-normalised binary contact flag (ref: crossmodal/tasks/_push.py:249-260)
    controls[..., 6] = np.where(rng.random((T, N)) < 0.78, -0.53, 1.88).astype(np.float32)
    image = rng.uniform(-1.0, 1.0, size=(T, N, 32, 32)).astype(np.float32)
    if blackout_fraction > 0:
        dark = rng.random((T, N)) < blackout_fraction
        image[dark] = 0.0
    observations = {
        "image": image,
        "gripper_pos": rng.standard_normal((T, N, 3)).astype(np.float32),
        "gripper_sensors": rng.standard_normal((T, N, 7)).astype(np.float32),
    }
    return (
        torch.from_numpy(states),
        {k: torch.from_numpy(v) for k, v in observations.items()},
        torch.from_numpy(controls),
    )
