"""A few fused particle-filter steps at BASELINE config C3's shape (N=4096, M=1000, sd=2), for ncu.
Usage:  ncu ... python tools/profile_step.py [precision] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters

precision = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, Mp, sd = 4096, 1000, 2
dev = torch.device("cuda:0")
filt = fill_parameters(M.PushUnimodalParticleFilter(), seed=0).to(dev).eval()
filt.num_particles = Mp
filt.precision = precision
plan = filt.fused_plan()
g = torch.Generator(device=dev).manual_seed(0)
mean = torch.randn(N, sd, device=dev, generator=g)
cov = (torch.eye(sd, device=dev) * 0.1)[None].expand(N, sd, sd).contiguous()
feats = [torch.randn(steps, N, 64, device=dev, generator=g), torch.randn(steps, N, 128, device=dev, generator=g)]
controls = torch.randn(steps, N, 7, device=dev, generator=g)
with torch.no_grad():
    filt.initialize_beliefs(mean=mean, covariance=cov)
    for t in range(steps):
        est = filt.forward(observations=None, controls=controls[t], _hoisted=([f[t] for f in feats], None))
torch.cuda.synchronize()
print("estimate checksum", float(est.double().sum()))
