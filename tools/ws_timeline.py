"""Hand-over timeline of the warp-specialised chain kernel (measurement build 3: make -C multimodalfilter_b200/csrc ablate).
Prints, for CTA 0 / chain 0 / second tile, per layer and group: issuer [wait start, a_ready seen, committed] and worker
[published, d_ready seen, epilogue done], in cycles relative to the first stamp.
Usage: MMF_LIB=tools/ubench/libmmf_ablate3.so python tools/ws_timeline.py [bf16x3|bf16]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multimodalfilter_b200 import _lib, fused, ops

_lib.LIB_PATH = os.path.abspath(os.environ.get("MMF_LIB", "tools/ubench/libmmf_ablate3.so"))
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
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
for _ in range(3):
    ops.pf_predict_measure(plan.struct, states, eps, rb, logw, None, 3, precision=ops.PRECISIONS[prec])
torch.cuda.synchronize()
buf = (C.c_ulonglong * (2 * 16 * 4 * 4))()
lib = _lib.load()
lib.mmf_debug_ws_stamps.restype = C.c_int
assert lib.mmf_debug_ws_stamps(buf) == 0
v = list(buf)
t0 = min(x for x in v if x)
at = lambda role, layer, grp, k: v[((role * 16 + layer) * 4 + grp) * 4 + k]
print(f"precision {prec}; cycles relative to the first stamp")
print("layer grp | issuer: wait  seen  committed | worker: published  d_ready  epilogue_done | commit->d_ready  epilogue  publish->seen")
for layer in range(10):
    for grp in range(4):
        i = [at(0, layer, grp, k) for k in range(3)]
        w = [at(1, layer, grp, k) for k in range(3)]
        if not i[0]:
            continue
        r = lambda x: x - t0 if x else -1
        print(f"{layer:5d} {grp:3d} | {r(i[0]):7d} {r(i[1]):7d} {r(i[2]):7d} | {r(w[0]):7d} {r(w[1]):7d} {r(w[2]):7d} | "
              f"{w[1] - i[2] if w[1] else -1:7d} {w[2] - w[1] if w[2] else -1:7d} {i[1] - w[0] if w[0] else -1:7d}")
