"""Times mmf_pf_normalize_resample at config C3's shape for every mode (phase cost breakdown)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalfilter_b200 import ops, _lib
N, M, sd = int(os.environ.get("N", 4096)), int(os.environ.get("M", 1000)), 2
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
states = torch.randn(N, M, sd, device=dev, generator=g)
logw = torch.randn(N, M, device=dev, generator=g) * float(os.environ.get('SPREAD', 1.0))
u_m = torch.rand(N, M, device=dev, dtype=torch.float64, generator=g)
u_s = torch.rand(N, device=dev, dtype=torch.float64, generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
modes = [("none", _lib.RESAMPLE_NONE, None), ("multinomial_strict", _lib.RESAMPLE_MULTINOMIAL_STRICT, u_m),
         ("multinomial_fast", _lib.RESAMPLE_MULTINOMIAL_FAST, u_m), ("systematic_strict", _lib.RESAMPLE_SYSTEMATIC_STRICT, u_s),
         ("systematic_fast", _lib.RESAMPLE_SYSTEMATIC_FAST, u_s)]
for name, mode, u in modes:
    ts = []
    for rep in range(6):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); ops.pf_normalize_resample(states, logw, mode=mode, uniforms=u); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = sorted(ts[1:])[len(ts[1:]) // 2] * 1e3
    byts = N * M * (4 + 4 * sd + (0 if mode == _lib.RESAMPLE_NONE else (8 if u is u_m else 0) + 4 * sd) + 4)
    print(f"{name:20s} {t:7.1f} us   {byts / t / 1e3:7.1f} GB/s algorithmic")
