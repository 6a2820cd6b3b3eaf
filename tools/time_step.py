"""Times the fused PF step kernels at config C3's shape.  Usage: python tools/time_step.py [precisions...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200 import _lib, ops

if os.environ.get("MMF_LIB"):  # measurement builds (make -C multimodalfilter_b200/csrc ablate)
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters

precisions = sys.argv[1:] or ["bf16x3", "bf16"]
N, Mp, sd, steps = 4096, 1000, 2, 12
dev = torch.device("cuda:0")
filt = fill_parameters(M.PushUnimodalParticleFilter(), seed=0).to(dev).eval()
filt.num_particles = Mp
g = torch.Generator(device=dev).manual_seed(0)
mean = torch.randn(N, sd, device=dev, generator=g)
cov = (torch.eye(sd, device=dev) * 0.1)[None].expand(N, sd, sd).contiguous()
feats = [torch.randn(steps, N, 64, device=dev, generator=g), torch.randn(steps, N, 128, device=dev, generator=g)]
controls = torch.randn(steps, N, 7, device=dev, generator=g)
for prec in precisions:
    filt.precision = prec
    with torch.no_grad():
        filt.initialize_beliefs(mean=mean, covariance=cov)
        for t in range(steps):
            if t == 4:
                torch.cuda.synchronize()
                ops.PROFILE.reset(enabled=True)
            est = filt.forward(observations=None, controls=controls[t], _hoisted=([f[t] for f in feats], None))
    prof = ops.PROFILE.collect()
    ops.PROFILE.reset(enabled=False)
    k = prof["kernels"]
    pm = k["pf_predict_measure"]["avg_ms"]
    print(f"{prec:7s} predict_measure {pm:7.3f} ms  ({189824.0 * N * Mp / pm / 1e9:7.1f} TFLOP/s alg)  "
          f"normalize_resample {k['pf_normalize_resample']['avg_ms']:.3f} ms  traj_rows {k['pf_traj_rows']['avg_ms']:.3f} ms  "
          f"checksum {float(est.double().sum()):.6f}")
