"""Times the chain kernel alone at config C3's shape (used with the measurement builds, whose outputs are garbage).
Usage: [MMF_LIB=...] python tools/time_chain.py [precisions...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200 import _lib, fused, ops

if os.environ.get("MMF_LIB"):
    _lib.LIB_PATH = os.path.abspath(os.environ["MMF_LIB"])
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters

N, Mp, sd = 4096, 1000, 2
dev = torch.device("cuda:0")
filt = fill_parameters(M.PushUnimodalParticleFilter(), seed=0).to(dev).eval()
plan = fused.PFPlan.build(filt)
plan.refresh(dev)
g = torch.Generator(device=dev).manual_seed(0)
states = torch.randn(N, Mp, sd, device=dev, generator=g)
logw = torch.full((N, Mp), -6.9, device=dev)
eps = torch.randn(N * Mp, sd, device=dev, generator=g)
controls = torch.randn(N, 7, device=dev, generator=g)
feats = [torch.randn(N, 64, device=dev, generator=g), torch.randn(N, 128, device=dev, generator=g)]
rb = ops.pf_traj_rows(plan.struct, 2, controls, feats)
for prec in sys.argv[1:] or ["bf16x3", "bf16"]:
    for _ in range(3):
        ops.pf_predict_measure(plan.struct, states, eps, rb, logw, None, 3, precision=ops.PRECISIONS[prec])
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.pf_predict_measure(plan.struct, states, eps, rb, logw, None, 3, precision=ops.PRECISIONS[prec])
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"{os.environ.get('MMF_LIB', 'product'):40s} variant {os.environ.get('MMF_TC_VARIANT', 'default'):8s} {prec:7s} chain {ms:7.3f} ms "
          f"({189824.0 * N * Mp / ms / 1e9:7.1f} TFLOP/s alg)")
