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
    """The slice of ``fannypack.utils.Buddy`` the reference's training / evaluation helpers touch
    (ref: crossmodal/train_helpers.py:15-26,155-162, crossmodal/eval_helpers.py:10-38,
    scripts/push_task/train_push.py:30-51,80,109-116): device choice, one Adam optimiser per ``optimizer_name``
    (lr 1e-3), ``minimize``, scalar logging hooks, ``state_dict`` checkpoints.  TensorBoard / YAML metadata are not
    written (outside the hot path); scalars are kept in ``self.scalars``.

    Multi-GPU: when ``torch.distributed`` is initialised, ``minimize`` all-reduces (averages) the gradients over the
    ranks between backward and the optimiser step -- the BPTT step's only collective."""

    def __init__(self, experiment_name, model=None, *, device=None, checkpoint_dir="checkpoints", lr=1e-3,
                 verbose=False, **_ignored):
        self.experiment_name = experiment_name
        self.checkpoint_dir = checkpoint_dir
        self.lr = lr
        self.verbose = verbose
        self.device = torch.device(device) if device is not None else torch.device(
            "cuda" if torch.cuda.is_available() else "cpu")
        self.model = None
        self.optimizers = {}
        self.optimizer_steps = 0
        self.scalars = {}
        self.metadata = {}
        self._scope = []
        if model is not None:
            self.attach_model(model)

    # ---- model / optimisers ----------------------------------------------------------------------------------------
    def attach_model(self, model):
        self.model = model.to(self.device)

    def get_optimizer(self, optimizer_name="primary"):
        opt = self.optimizers.get(optimizer_name)
        if opt is None:
            assert self.model is not None, "attach a model first"
            opt = torch.optim.Adam(self.model.parameters(), lr=self.lr)
            self.optimizers[optimizer_name] = opt
        return opt

    def set_learning_rate(self, value, optimizer_name="primary"):
        for group in self.get_optimizer(optimizer_name).param_groups:
            group["lr"] = value

    def minimize(self, loss, optimizer_name="primary", *, retain_graph=False, checkpoint_interval=None):
        opt = self.get_optimizer(optimizer_name)
        opt.zero_grad(set_to_none=True)
        loss.backward(retain_graph=retain_graph)
        from ...distributed import allreduce_gradients  # no-op unless torch.distributed is initialised

        allreduce_gradients(self.model)
        opt.step()
        self.optimizer_steps += 1
        if checkpoint_interval and self.optimizer_steps % checkpoint_interval == 0:
            self.save_checkpoint()

    # ---- logging ---------------------------------------------------------------------------------------------------
    def log_scope(self, name):
        buddy = self

        class _Scope:
            def __enter__(self_inner):
                buddy._scope.append(name)

            def __exit__(self_inner, *exc):
                buddy._scope.pop()

        return _Scope()

    def log_scope_push(self, name):
        self._scope.append(name)

    def log_scope_pop(self, name=None):
        self._scope.pop()

    def log_scalar(self, name, value):
        key = "/".join(self._scope + [name])
        self.scalars.setdefault(key, []).append((self.optimizer_steps, float(value)))

    log = log_scalar

    def set_metadata(self, metadata):
        self.metadata = dict(metadata)

    def add_metadata(self, metadata):
        self.metadata.update(metadata)

    # ---- checkpoints (torch state_dict of the whole filter + optimiser states) ---------------------------------------
    def _path(self, label, experiment_name=None):
        import os

        name = experiment_name or self.experiment_name
        return os.path.join(self.checkpoint_dir, f"{name}-{label}.ckpt")

    def save_checkpoint(self, label=None):
        import os

        label = label if label is not None else f"{self.optimizer_steps:016d}"
        os.makedirs(self.checkpoint_dir, exist_ok=True)
        torch.save({"state_dict": self.model.state_dict(), "steps": self.optimizer_steps,
                    "optimizers": {k: o.state_dict() for k, o in self.optimizers.items()}}, self._path(label))

    def load_checkpoint(self, label=None, *, experiment_name=None, path=None):
        ckpt = torch.load(path or self._path(label, experiment_name), map_location=self.device)
        self.model.load_state_dict(ckpt["state_dict"])
        self.optimizer_steps = ckpt.get("steps", 0)
        for k, sd in ckpt.get("optimizers", {}).items():
            self.get_optimizer(k).load_state_dict(sd)

    def load_checkpoint_module(self, source, target=None, label=None, *, experiment_name=None, path=None):
        """Load the sub-tree ``source`` of a checkpoint into the sub-module ``target`` of the attached model."""
        target = source if target is None else target
        ckpt = torch.load(path or self._path(label, experiment_name), map_location=self.device)
        prefix = source + "."
        sub = {k[len(prefix):]: v for k, v in ckpt["state_dict"].items() if k.startswith(prefix)}
        assert sub, f"no entries under {source!r} in the checkpoint"
        module = self.model
        for part in target.split("."):
            module = getattr(module, part) if not part.isdigit() else module[int(part)]
        module.load_state_dict(sub)
