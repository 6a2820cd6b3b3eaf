"""ORACLE: fannypack.utils pieces used at the API boundary (Appendix A.8):
SliceWrapper (ref: crossmodal/eval_helpers.py:88-110,121,140), to_torch/to_numpy
(ref: crossmodal/eval_helpers.py:100-102,152), freeze/unfreeze.  Buddy is a name-only stub
(module-level annotation at ref: crossmodal/eval_helpers.py:11)."""
import numpy as np
import torch


class SliceWrapper:
    """Apply one index expression to every leaf of a tensor / array / list / dict-of-those."""

    def __init__(self, data):
        self.data = data

    def _leaves(self):
        if isinstance(self.data, dict):
            return list(self.data.values())
        return [self.data]

    def __getitem__(self, index):
        if isinstance(self.data, dict):
            return {key: value[index] for key, value in self.data.items()}
        return self.data[index]

    def __len__(self):
        leaves = self._leaves()
        if isinstance(self.data, dict) and not leaves:
            return 0
        return len(leaves[0])

    @property
    def shape(self):
        leaves = self._leaves()
        assert leaves, "empty container has no shape"
        shapes = [tuple(leaf.shape) for leaf in leaves]
        common = []
        for dims in zip(*shapes):
            if any(d != dims[0] for d in dims):
                break
            common.append(dims[0])
        return tuple(common)

    def map(self, fn):
        if isinstance(self.data, dict):
            return {key: fn(value) for key, value in self.data.items()}
        return fn(self.data)

    def append(self, other):
        if isinstance(self.data, dict):
            assert isinstance(other, dict)
            for key, value in other.items():
                self.data.setdefault(key, []).append(value)
        else:
            self.data.append(other)

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]


def to_torch(x, device="cpu", convert_doubles_to_floats=True):
    if isinstance(x, dict):
        return {k: to_torch(v, device, convert_doubles_to_floats) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_torch(v, device, convert_doubles_to_floats) for v in x)
    out = torch.from_numpy(np.asarray(x))
    if convert_doubles_to_floats and out.dtype == torch.float64:
        out = out.float()
    return out.to(device)


def to_numpy(x):
    if isinstance(x, dict):
        return {k: to_numpy(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_numpy(v) for v in x)
    return x.detach().cpu().numpy()


def to_device(x, device, detach=False):
    if isinstance(x, dict):
        return {k: to_device(v, device, detach) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return type(x)(to_device(v, device, detach) for v in x)
    x = x.detach() if detach else x
    return x.to(device)


def freeze_module(module, recurse=True):
    for p in module.parameters(recurse=recurse):
        p.requires_grad = False


def unfreeze_module(module, recurse=True):
    for p in module.parameters(recurse=recurse):
        p.requires_grad = True


class Buddy:  # name-only stub: experiment management is out of scope
    pass
