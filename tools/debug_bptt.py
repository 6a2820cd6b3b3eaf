import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, numpy as np
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from oracle import crossmodal_port as port
from oracle.noise import RecordedNoise
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories
from util import ReplayNoise, draw_noise
name, sd, N, Mp, T = "PushCrossmodalParticleFilter", 2, 6, 30, int(sys.argv[1]) if len(sys.argv) > 1 else 5
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
init, eps, _ = draw_noise(T, N, Mp, sd, seed=13)
states, obs, controls = synthetic_trajectories(T + 1, N, sd, seed=14)
cov = (torch.eye(sd) * 0.1)[None].expand(N, sd, sd)
res = []
for side in ("oracle", "product"):
    if side == "oracle":
        f = fill_parameters(getattr(port, name)(), seed=27); f.noise = RecordedNoise(init_eps=init, process_eps=eps); dev = "cpu"
    else:
        f = fill_parameters(getattr(M, name)(), seed=27).to("cuda:0"); f.noise = ReplayNoise(init_eps=init, process_eps=eps); dev = "cuda:0"; f.precision = prec
    f.train(); f.num_particles = Mp
    for prm in f.dynamics_model.parameters(): prm.requires_grad_(False)
    f.initialize_beliefs(mean=states[0].to(dev), covariance=cov.to(dev).contiguous())
    est = f.forward_loop(observations={k: v[1:].to(dev) for k, v in obs.items()}, controls=controls[1:].to(dev))
    loss = torch.mean((est - states[1:].to(dev)) ** 2); loss.backward()
    res.append((loss.item(), {k: p.grad.detach().cpu().double() for k, p in f.named_parameters() if p.grad is not None}))
(lo, go), (lp, gp) = res
print("loss", lo, lp)
for k in go:
    e = go[k]; a = gp[k]; scale = e.pow(2).mean().sqrt().item()
    print(f"{k:85s} scale {scale:9.2e} max|err|/scale {((a-e).abs().max().item()/max(scale,1e-30)):9.2e}")
