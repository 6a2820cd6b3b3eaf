"""Where does the end-to-end forward_loop time go at config C3?  (H2D, hoisted encoders, recursion, D2H)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories
N, Mp, sd, T = 4096, 1000, 2, 100
dev = torch.device("cuda:0")
f = fill_parameters(M.PushUnimodalParticleFilter(), seed=0).to(dev).eval(); f.num_particles = Mp
states, obs, controls = synthetic_trajectories(T + 1, N, sd, seed=0)
host_obs = {k: v[1:].contiguous().pin_memory() for k, v in obs.items()}
host_controls = controls[1:].contiguous().pin_memory()
def ev(): e = torch.cuda.Event(enable_timing=True); e.record(); return e
for rep in range(2):
    torch.cuda.synchronize()
    e0 = ev()
    o = {k: v.to(dev, non_blocking=True) for k, v in host_obs.items()}; c = host_controls.to(dev, non_blocking=True)
    e1 = ev()
    plan = f.fused_plan()
    with torch.no_grad():
        feats, modw = f.hoist_observations(plan, o, T, N)
    e2 = ev()
    with torch.no_grad():
        f.initialize_beliefs(mean=states[0].to(dev), covariance=(torch.eye(sd, device=dev) * 0.1)[None].expand(N, sd, sd).contiguous())
        for t in range(T):
            est = f.forward(observations=None, controls=c[t], _hoisted=([None if x is None else x[t] for x in feats], None))
    e3 = ev()
    torch.cuda.synchronize()
    print(f"rep {rep}: H2D {e0.elapsed_time(e1):.1f} ms | hoisted encoders {e1.elapsed_time(e2):.1f} ms | recursion {e2.elapsed_time(e3):.1f} ms")
for fmt in ("default", "channels_last"):
    enc = plan.heads[0].head.observation_image_layers
    x = o["image"].reshape(T * N, 1, 32, 32)[:16384]
    if fmt == "channels_last":
        enc = enc.to(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32; torch.backends.cuda.matmul.allow_tf32 = tf32
        with torch.no_grad():
            enc(x); torch.cuda.synchronize(); a = ev(); enc(x); b = ev(); torch.cuda.synchronize()
        print(f"image encoder, 16384 images, {fmt}, tf32={tf32}: {a.elapsed_time(b):.2f} ms  ({a.elapsed_time(b)*25:.0f} ms for 409600)")
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        enc(x); torch.cuda.synchronize(); a = ev(); enc(x); b = ev(); torch.cuda.synchronize()
    print(f"image encoder, 16384 images, {fmt}, bf16 autocast: {a.elapsed_time(b):.2f} ms")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
for fmt in ("default", "channels_last"):
    enc = plan.heads[0].head.observation_image_layers
    x = o["image"].reshape(T * N, 1, 32, 32)[:16384]
    if fmt == "channels_last":
        enc = enc.to(memory_format=torch.channels_last); x = x.contiguous(memory_format=torch.channels_last)
    else:
        enc = enc.to(memory_format=torch.contiguous_format)
    with torch.no_grad():
        ref = enc(x); enc(x); torch.cuda.synchronize(); a = ev(); y = enc(x); b = ev(); torch.cuda.synchronize()
    print(f"cudnn.benchmark=True fp32 {fmt}: {a.elapsed_time(b):.2f} ms per 16384 images ({a.elapsed_time(b)*25:.0f} ms for 409600)")
