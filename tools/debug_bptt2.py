import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters, synthetic_trajectories
from util import ReplayNoise, draw_noise
name, sd, N, Mp = "PushCrossmodalParticleFilter", 2, 6, 30
T = int(sys.argv[1]) if len(sys.argv) > 1 else 5
init, eps, _ = draw_noise(T, N, Mp, sd, seed=13)
states, obs, controls = synthetic_trajectories(T + 1, N, sd, seed=14)
cov = (torch.eye(sd) * 0.1)[None].expand(N, sd, sd)
res = []
dev = "cuda:0"
for prec in ("fp32", "bf16x3"):
    f = fill_parameters(getattr(M, name)(), seed=27).to(dev); f.noise = ReplayNoise(init_eps=init, process_eps=eps); f.precision = prec
    f.train(); f.num_particles = Mp
    for prm in f.dynamics_model.parameters(): prm.requires_grad_(False)
    f.initialize_beliefs(mean=states[0].to(dev), covariance=cov.to(dev).contiguous())
    ests = []
    o = {k: v[1:].to(dev) for k, v in obs.items()}; c = controls[1:].to(dev)
    for t in range(T):
        ests.append(f(observations={k: v[t] for k, v in o.items()}, controls=c[t]))
    est = torch.stack(ests)
    loss = torch.mean((est - states[1:].to(dev)) ** 2); loss.backward()
    res.append((loss.item(), est.detach(), {k: p.grad.detach().double() for k, p in f.named_parameters() if p.grad is not None}))
(lo, eo, go), (lp, ep, gp) = res
print("T", T, "loss", lo, lp, "est maxdiff", (eo - ep).abs().max().item())
for k in go:
    if "measurement_models.1" not in k: continue
    e = go[k]; a = gp[k]; scale = e.pow(2).mean().sqrt().item()
    print(f"{k:85s} scale {scale:9.2e} max|err|/scale {((a-e).abs().max().item()/max(scale,1e-30)):9.2e}")
