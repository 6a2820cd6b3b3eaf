"""Phase timestamps of one group of the tc chain kernel (debug).  MMF_TC_TIMESTAMPS=1 python tools/stamps.py [precision]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MMF_TC_TIMESTAMPS"] = "1"
import torch, numpy as np
from multimodalfilter_b200 import _lib
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
N, Mp, sd = 4096, 1000, 2
dev = torch.device("cuda:0")
f = fill_parameters(M.PushUnimodalParticleFilter(), seed=0).to(dev).eval(); f.num_particles = Mp; f.precision = prec
g = torch.Generator(device=dev).manual_seed(0)
with torch.no_grad():
    f.initialize_beliefs(mean=torch.randn(N, sd, device=dev, generator=g), covariance=(torch.eye(sd, device=dev) * 0.1)[None].expand(N, sd, sd).contiguous())
    feats = [torch.randn(N, 64, device=dev, generator=g), torch.randn(N, 128, device=dev, generator=g)]
    lib = _lib.load()
    buf = (C.c_ulonglong * (5 * 1600))()
    for t in range(3):
        f.forward(observations=None, controls=torch.randn(N, 7, device=dev, generator=g), _hoisted=(feats, None))
        lib.mmf_debug_tc_timestamps.restype = C.c_int
        n = lib.mmf_debug_tc_timestamps(buf, 1600)
a = np.array(buf[: 5 * n], dtype=np.int64).reshape(n, 5)
a = a[np.argsort(a[:, 0])]
print("layers stamped", n)
print("layer | st-wait | bar.sync | issue+commit | mma wait | epilogue(next ts0 - ts4)")
for i in range(min(n, 12)):
    nxt = a[i + 1, 0] - a[i, 4] if i + 1 < n else -1
    print(f"{i:5d} | {a[i,1]-a[i,0]:7d} | {a[i,2]-a[i,1]:8d} | {a[i,3]-a[i,2]:12d} | {a[i,4]-a[i,3]:8d} | {nxt}")
d = np.diff(a[:, 0])
print("mean cycles per layer (this group):", d[:9].mean())
