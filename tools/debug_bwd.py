import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
torch.backends.cuda.matmul.allow_tf32 = False
from multimodalfilter_b200 import fused, ops, training
from multimodalfilter_b200.crossmodal import models as M
from multimodalfilter_b200.synthetic import fill_parameters
dev = torch.device("cuda:0")
f = fill_parameters(M.PushCrossmodalParticleFilter(), seed=27).to(dev)
plan = fused.PFPlan.build(f); plan.refresh(dev, backward=True)
N, Mp, sd = 6, 30, 2
g = torch.Generator(device=dev).manual_seed(1)
states = torch.randn(N, Mp, sd, device=dev, generator=g); eps = torch.randn(N * Mp, sd, device=dev, generator=g)
rows = torch.randn(1 + plan.K, N, 64, device=dev, generator=g)
moved, ll, act = ops.pf_heads_forward_train(plan.struct, states, eps, rows, 3)
d_ll = torch.randn(plan.K, N, Mp, device=dev, generator=g)
delta = ops.pf_heads_backward(plan.struct, N, Mp, act, d_ll, 3)
x = moved.reshape(-1, sd)
for k, spec in enumerate(plan.heads):
    (in_lin, pre), (mid, post, out) = spec.state, spec.shared
    zs = []
    def lin(w, b, a):
        z = F.linear(a, w, b); z.retain_grad(); zs.append(z); return z
    a = torch.relu(lin(in_lin.weight, in_lin.bias, x)); acts = [a]
    for r in pre:
        t = torch.relu(lin(r.block1.weight, r.block1.bias, a)); acts.append(t)
        a = torch.relu(lin(r.block2.weight, r.block2.bias, t) + a); acts.append(a)
    rb = rows[1 + k].repeat_interleave(Mp, dim=0)
    z = F.linear(a, mid.weight[:, spec.feat_dim:]) + rb; z.retain_grad(); zs.append(z)
    a = torch.relu(z); acts.append(a)
    for r in post:
        t = torch.relu(lin(r.block1.weight, r.block1.bias, a)); acts.append(t)
        a = torch.relu(lin(r.block2.weight, r.block2.bias, t) + a); acts.append(a)
    llk = F.linear(a, out.weight, out.bias)[:, 0]
    print(f"head {k}: ll err {(llk.detach() - ll[k].reshape(-1)).abs().max().item():.2e}")
    (llk * d_ll[k].reshape(-1)).sum().backward()
    L = len(zs) - 1
    for i, a_ref in enumerate(acts):
        print(f"   act[{i}] err {(act[k, i] - a_ref.detach()).abs().max().item():.2e}", end="")
    print()
    # zs[0] = input layer pre-act -> delta plane L ; zs[1+l] -> delta plane l
    for l in range(L):
        ref = zs[1 + l].grad; got = delta[k, l]
        print(f"   delta[{l}] relerr {((got - ref).abs().max() / ref.abs().max()).item():.2e}", end="")
    ref = zs[0].grad; got = delta[k, L]
    print(f"   delta_in relerr {((got - ref).abs().max() / ref.abs().max()).item():.2e}")
