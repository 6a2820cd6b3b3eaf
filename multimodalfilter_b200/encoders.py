"""Image observation encoder with a fused sm_100a convolutional trunk.

``ImageEncoder`` IS the reference's ``nn.Sequential`` (same children, same ``state_dict`` keys:
ref: crossmodal/push_models/layers.py:93-104, crossmodal/door_models/layers.py:43-63); only its forward
changes: under ``torch.no_grad()`` on a CUDA device the Conv2d layers run through ``mmf_enc_trunk`` (one launch,
tensor cores, bf16 hi/lo split operands, fp32 accumulation) and the Flatten/Linear tail stays with torch.  With autograd enabled (encoder training / pre-training) the plain module path runs, so
gradients are untouched.  The trunk's scratch maps are kept per (device, stream), so the same encoder instance may run on several streams.  There is no CPU variant of the fused trunk: on a CPU tensor the module is the
ordinary torch Sequential, exactly as in the reference.
"""
import torch
import torch.nn as nn

from . import ops
from .fannypack.nn import resblocks

CHUNK_IMAGES = 16384  # images per launch of the trunk: bounds the (n, 8, 32, 32) fp32 output to 0.5 GB


def _is(m, cls, **attrs):
    return isinstance(m, cls) and all(getattr(m, k) == v for k, v in attrs.items())


class ImageEncoder(nn.Sequential):
    fused_trunk = True  # set False (class or instance) to force the torch path

    def _trunk(self):
        """The Conv2d layers if this Sequential has the default (non-pooled) reference architecture, else None."""
        m = list(self.children())
        ok = (
            len(m) >= 8
            and _is(m[0], nn.Conv2d, in_channels=1, out_channels=32, kernel_size=(5, 5), padding=(2, 2))
            and isinstance(m[1], nn.ReLU)
            and isinstance(m[2], resblocks.Conv2d)
            and _is(m[2].block1, nn.Conv2d, in_channels=32, out_channels=32, kernel_size=(3, 3), padding=(1, 1))
            and _is(m[2].block2, nn.Conv2d, in_channels=32, out_channels=32, kernel_size=(3, 3), padding=(1, 1))
            and _is(m[3], nn.Conv2d, in_channels=32, out_channels=16, kernel_size=(3, 3), padding=(1, 1))
            and isinstance(m[4], nn.ReLU)
            and _is(m[5], nn.Conv2d, in_channels=16, kernel_size=(3, 3), padding=(1, 1))
            and m[5].out_channels <= 16
            and isinstance(m[6], nn.Flatten)
        )
        if not ok:
            return None
        convs = [m[0], m[2].block1, m[2].block2, m[3], m[5]]
        if any(c.stride != (1, 1) or c.dilation != (1, 1) or c.groups != 1 or c.bias is None for c in convs):
            return None
        return convs

    def _packed(self, convs, device):
        key = tuple((p.data_ptr(), p._version) for c in convs for p in (c.weight, c.bias)) + (str(device),)
        cache = self.__dict__.get("_mmf_packed")
        if cache is None or cache[0] != key:
            cache = (key, ops.enc_pack_trunk(convs))
            self.__dict__["_mmf_packed"] = cache
        return cache[1]

    def _scratch(self, device):
        """Scratch maps of the fused trunk, one set per (device, stream): launches on one stream are ordered, and two
        streams running the same encoder instance never share maps.  Zeroed once: the kernels never write the guards."""
        key = (str(device), torch.cuda.current_stream(device).cuda_stream)
        cache = self.__dict__.setdefault("_mmf_ws", {})
        ws = cache.get(key)
        if ws is None:
            if len(cache) >= 4:  # streams come and go: keep the footprint bounded
                cache.clear()
            ws = cache[key] = ops.enc_trunk_scratch(device)
        return ws

    @staticmethod
    def _split_tail(tail):
        return len(tail) >= 2 and isinstance(tail[0], nn.Flatten) and isinstance(tail[1], nn.Linear)

    @staticmethod
    def _flatten_linear(h, linear):
        """Flatten + Linear(C*1024 -> units) as one batched GEMM over the C channel slabs: the plain (n, 8192) x
        (8192, 64) product gives the library only n/128 thread blocks; batching over channels gives C times more."""
        n, C = h.shape[0], h.shape[1]
        w = linear.weight.view(linear.out_features, C, 1024).permute(1, 2, 0)  # (C, 1024, units)
        part = torch.bmm(h.view(n, C, 1024).transpose(0, 1), w)                 # (C, n, units)
        return part.sum(dim=0) + linear.bias

    def forward(self, x):
        convs = self._trunk() if self.fused_trunk else None
        use_fused = (
            convs is not None
            and x.is_cuda
            and x.dtype == torch.float32
            and x.dim() == 4
            and tuple(x.shape[1:]) == (1, 32, 32)
            and not torch.is_grad_enabled()
        )
        if not use_fused:
            return super().forward(x)
        weights = self._packed(convs, x.device)
        scratch = self._scratch(x.device)
        cout = convs[4].out_channels
        tail = list(self.children())[6:]
        outs = []
        images = x.reshape(-1, 32, 32).contiguous()
        for lo in range(0, images.shape[0], CHUNK_IMAGES):
            # one launch: a persistent CTA carries each image through all five layers (maps stay in L2)
            h = ops.enc_trunk(images[lo:lo + CHUNK_IMAGES], weights, scratch, cout)
            h = self._flatten_linear(h, tail[1]) if self._split_tail(tail) else tail[1](tail[0](h))
            for layer in tail[2:]:
                h = layer(h)
            outs.append(h)
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)
