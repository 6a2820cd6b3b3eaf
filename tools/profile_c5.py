"""Per-kernel times of one C5 sweep point (PushUnimodalParticleFilter, M particles per trajectory, a tile of trajectories).
Usage: python tools/profile_c5.py M [n_tile] [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200 import ops
from multimodalfilter_b200.crossmodal import models as M_
from multimodalfilter_b200.synthetic import fill_parameters

Mp = int(sys.argv[1])
n_tile = int(sys.argv[2]) if len(sys.argv) > 2 else max(1, (1 << 27) // Mp)
mode = sys.argv[3] if len(sys.argv) > 3 else "multinomial"
sd, T = 2, 4
dev = torch.device("cuda:0")
filt = fill_parameters(M_.PushUnimodalParticleFilter(), seed=0).to(dev).eval()
filt.num_particles = Mp
filt.resample_mode = mode
g = torch.Generator(device=dev).manual_seed(0)
mean = torch.randn(n_tile, sd, device=dev, generator=g)
cov = (torch.eye(sd, device=dev) * 0.1)[None].expand(n_tile, sd, sd).contiguous()
feats = [torch.randn(T, n_tile, 64, device=dev, generator=g), torch.randn(T, n_tile, 128, device=dev, generator=g)]
controls = torch.randn(T, n_tile, 7, device=dev, generator=g)


def one_pass():
    with torch.no_grad():
        filt.initialize_beliefs(mean=mean, covariance=cov)
        for t in range(T):
            filt.forward(observations=None, controls=controls[t], _hoisted=([f[t] for f in feats], None))


one_pass()
torch.cuda.synchronize()
ops.PROFILE.reset(enabled=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
one_pass()
b.record()
prof = ops.PROFILE.collect()
ms = a.elapsed_time(b)
print(f"M={Mp} tile={n_tile} {mode}: {ms:.2f} ms for {T} steps = {n_tile * Mp * T / ms / 1e6:.3f} G particle-steps/s")
for k, v in sorted(prof["kernels"].items(), key=lambda kv: -kv[1]["total_ms"]):
    print(f"   {k:28s} {v['count']:3d} x {v['avg_ms']:9.3f} ms = {v['total_ms']:9.2f} ms ({100 * v['total_ms'] / ms:5.1f} %)")
