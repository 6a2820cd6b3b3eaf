"""``SliceWrapper`` and the tensor converters the reference's helpers call at the filter boundary
(ref: crossmodal/eval_helpers.py:88-110,121,140,152)."""
from typing import Any, Callable

import numpy as np
import torch


class SliceWrapper:
    """Index every leaf of a (possibly dict-valued) batch with one expression."""

    def __init__(self, data: Any):
        self.data = data

    @property
    def _is_dict(self):
        return isinstance(self.data, dict)

    def __getitem__(self, index):
        return {k: v[index] for k, v in self.data.items()} if self._is_dict else self.data[index]

    def __len__(self):
        if self._is_dict:
            return len(next(iter(self.data.values()))) if self.data else 0
        return len(self.data)

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    @property
    def shape(self):
        leaves = list(self.data.values()) if self._is_dict else [self.data]
        if not leaves:
            raise ValueError("shape of an empty container")
        lead = tuple(leaves[0].shape)
        for leaf in leaves[1:]:
            n = 0
            for a, b in zip(lead, tuple(leaf.shape)):
                if a != b:
                    break
                n += 1
            lead = lead[:n]
        return lead

    def map(self, fn: Callable):
        return {k: fn(v) for k, v in self.data.items()} if self._is_dict else fn(self.data)

    def append(self, other):
        if self._is_dict:
            for k, v in other.items():
                self.data.setdefault(k, []).append(v)
        else:
            self.data.append(other)


def _walk(x, fn):
    if isinstance(x, dict):
        return {k: _walk(v, fn) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(_walk(v, fn) for v in x)
    return fn(x)


def to_torch(x, device="cpu", convert_doubles_to_floats=True):
    def conv(a):
        t = torch.as_tensor(np.asarray(a))
        if convert_doubles_to_floats and t.dtype == torch.float64:
            t = t.float()
        return t.to(device)

    return _walk(x, conv)


def to_numpy(x):
    return _walk(x, lambda t: t.detach().cpu().numpy())


def to_device(x, device, detach=False):
    return _walk(x, lambda t: (t.detach() if detach else t).to(device))


def freeze_module(module, recurse=True):
    for p in module.parameters(recurse=recurse):
        p.requires_grad = False


def unfreeze_module(module, recurse=True):
    for p in module.parameters(recurse=recurse):
        p.requires_grad = True


class Buddy:
    """Placeholder: experiment management (checkpoints, TensorBoard) is outside the hot path."""
