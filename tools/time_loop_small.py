"""Times mmf_pf_forward_loop alone at config C1's shape (PushCrossmodalParticleFilter, 32 x 30 particles, 50 steps).
Usage: [MMF_PF_LOOP_SMALL=0|1] [MMF_LS_NT=1|2] [PREC=bf16x3|bf16|fp32] python tools/time_loop_small.py [N M T [mode]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200 import _lib, fused, ops
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters

N, Mp, T = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (32, 30, 50)))
mode = sys.argv[4] if len(sys.argv) > 4 else "multinomial"
prec = os.environ.get("PREC", "bf16x3")
sd = 2
dev = torch.device("cuda:0")
filt = fill_parameters(M.PushCrossmodalParticleFilter(), seed=0).to(dev).eval()
plan = fused.PFPlan.build(filt)
plan.refresh(dev)
g = torch.Generator(device=dev).manual_seed(0)
states0 = torch.randn(N, Mp, sd, device=dev, generator=g)
logw0 = torch.full((N, Mp), -3.4, device=dev)
eps = torch.randn(T, N * Mp, sd, device=dev, generator=g)
controls = torch.randn(T, N, 7, device=dev, generator=g)
feats = [torch.randn(T, N, 64, device=dev, generator=g), torch.randn(T, N, 128, device=dev, generator=g)]
modw = torch.randn(T, N, 2, device=dev, generator=g)
m = ops.RESAMPLE_MODES[mode]
u = torch.rand((T, N) if ops.is_systematic(m) else (T, N, Mp), device=dev, dtype=torch.float64, generator=g)


def once():
    s, l = states0.clone(), logw0.clone()
    return ops.pf_forward_loop(plan.struct, s, l, controls, feats, modw, 3, eps, precision=ops.PRECISIONS[prec],
                               estimation=ops.ESTIMATION["weighted_average"], mode=m, uniforms=u)


for _ in range(3):
    est = once()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    est = once()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
persistent = _lib.load().mmf_pf_forward_loop_persistent(N, Mp)
print(f"N={N} M={Mp} T={T} {mode:18s} {prec:7s} LOOP_SMALL={os.environ.get('MMF_PF_LOOP_SMALL', 'auto'):4s} persistent={persistent} "
      f"nt={os.environ.get('MMF_LS_NT', 'default'):7s}: {ms:7.3f} ms per pass, {ms * 1e3 / T:6.1f} us per filter step, "
      f"checksum {float(est.double().sum()):.6f}")
