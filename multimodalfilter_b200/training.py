"""BPTT support: the fused training step of the particle filter (BASELINE config C4).

In train mode torchfilter does not resample (A.3) and every curriculum of the reference freezes the
dynamics before end-to-end training (ref: scripts/push_task/train_push.py:154,213), so the particle
states carry no gradient to a trainable leaf; the loss reaches the parameters only through the
log-weights.  One training step is therefore

    moved, ll_k = kernel(states, eps, rows)            forward:  mmf_pf_heads_forward_train
    d ll_k -> delta of every head layer                backward: mmf_pf_heads_backward
    dW = delta^T a, db = sum delta, d rows_k = sum_m delta_mid      (reductions over the N*M rows)

wrapped in one ``torch.autograd.Function``; the fusion log-sum-exp, the normalisation and the
estimate are (N, M)-sized torch ops that autograd differentiates itself, and the per-trajectory
pieces (observation encoders, the observation half of the first shared Linear, the modality weight
model) stay ordinary torch modules, so their gradients come from autograd as well.
"""
from typing import List

import torch

from . import _lib, ops

U = _lib.UNITS


def head_parameters(spec) -> List[torch.Tensor]:
    """Trainable tensors of one head's per-particle chain, in a fixed order (see ``_grads_for_head``)."""
    (in_lin, pre), (mid, post, out) = spec.state, spec.shared
    ps = [in_lin.weight, in_lin.bias]
    for r in pre:
        ps += [r.block1.weight, r.block1.bias, r.block2.weight, r.block2.bias]
    ps += [mid.weight]
    for r in post:
        ps += [r.block1.weight, r.block1.bias, r.block2.weight, r.block2.bias]
    ps += [out.weight, out.bias]
    return ps


def mid_layer_index(spec) -> int:
    return 2 * len(spec.state[1])


def _grads_for_head(spec, dW, db, g_in, g_out, d_ll_sum):
    """dW (L, 64, 64), db (L+1, 64), g_in (64, sd), g_out (64,) from mmf_pf_heads_weight_grads; d_ll_sum (1,) = the sum
    of d ll over the particles (gradient of the output bias).  Returns grads in head_parameters order and the index of
    the mid layer."""
    (in_lin, pre), (mid, post, out) = spec.state, spec.shared
    L = 2 * len(pre) + 1 + 2 * len(post)
    grads = [g_in, db[L]]
    layer = 0
    for _ in pre:
        for _half in range(2):
            grads += [dW[layer], db[layer]]
            layer += 1
    g_mid = torch.zeros_like(mid.weight)
    g_mid[:, spec.feat_dim:] = dW[layer]  # state half; the observation half flows through the rows
    grads += [g_mid]
    mid_layer = layer
    layer += 1
    for _ in post:
        for _half in range(2):
            grads += [dW[layer], db[layer]]
            layer += 1
    grads += [g_out[None, :], d_ll_sum.reshape(1)]
    return grads, mid_layer


class HeadGradToken(torch.autograd.Function):
    """All head parameters of a plan -> ONE flat tensor whose only purpose is to carry gradients: ``FusedHeads`` takes the
    token instead of the ~30 parameters per head and hands its backward the per-step parameter gradients as one flat
    tensor [dW | db | g_in | g_out | sum d_ll] per head.  Over a T-step sequence autograd then sums T flat tensors (T - 1
    adds) and this node distributes the total to the parameters ONCE, instead of 61 ``AccumulateGrad`` additions per
    filter step (915 small kernels per C4 training step).  Heads that no step of the sequence enabled get ``None``."""

    @staticmethod
    def forward(ctx, plan, info, *params):
        ctx.plan, ctx.info = plan, info
        return params[0].new_zeros(token_size(plan))

    @staticmethod
    def backward(ctx, d_token):
        plan, info = ctx.plan, ctx.info
        per_head = token_size(plan) // plan.K
        grads = []
        for k, spec in enumerate(plan.heads):
            n_params = len(head_parameters(spec))
            if d_token is None or not (info["used"] >> k) & 1:
                grads += [None] * n_params
                continue
            L, sd = _layers(spec), spec.sd
            part = d_token[k * per_head:(k + 1) * per_head]
            o = 0
            dW = part[o:o + L * U * U].view(L, U, U); o += L * U * U
            db = part[o:o + (L + 1) * U].view(L + 1, U); o += (L + 1) * U
            g_in = part[o:o + U * sd].view(U, sd); o += U * sd
            g_out = part[o:o + U]; o += U
            g, _ = _grads_for_head(spec, dW, db, g_in, g_out, part[o:o + 1])
            grads += g
        return (None, None, *grads)


def _layers(spec) -> int:
    return 2 * len(spec.state[1]) + 1 + 2 * len(spec.shared[1])


def token_size(plan) -> int:
    spec = plan.heads[0]
    L, sd = _layers(spec), spec.sd
    return plan.K * (L * U * U + (L + 1) * U + U * sd + U + 1)


class FusedHeads(torch.autograd.Function):
    """(states, eps, dynamics row, head rows, gradient token of the head parameters) -> (moved, ll[K, N, M])."""

    @staticmethod
    def forward(ctx, plan, states, eps, dyn_row, head_rows, enabled_mask, precision, token, info):
        info["used"] |= enabled_mask
        N, M, sd = states.shape
        dev = states.device
        plan.refresh(dev, backward=True)
        rowbias = torch.cat([dyn_row[None], head_rows.detach()]).contiguous()
        moved, ll, act = ops.pf_heads_forward_train(plan.struct, states.detach(), eps, rowbias, enabled_mask,
                                                    precision=precision)
        if enabled_mask != (1 << plan.K) - 1:
            # planes of disabled heads are never written by the kernel: clear them here, once, so that backward can
            # treat the saved tensor as read-only (retain_graph / a second backward see the same values)
            for k in range(plan.K):
                if not (enabled_mask >> k) & 1:
                    act[k].zero_()
        ctx.plan, ctx.mask, ctx.shape = plan, enabled_mask, (N, M, sd)
        ctx.save_for_backward(act, moved)
        ctx.mark_non_differentiable(moved)
        return moved, ll

    @staticmethod
    def backward(ctx, _d_moved, d_ll):
        plan, mask = ctx.plan, ctx.mask
        N, M, sd = ctx.shape
        act, moved = ctx.saved_tensors
        d_ll = torch.nan_to_num(d_ll.contiguous(), nan=0.0)
        delta = ops.pf_heads_backward(plan.struct, N, M, act, d_ll, mask)
        if mask != (1 << plan.K) - 1:  # planes of disabled heads are never written by the kernels
            for k in range(plan.K):
                if not (mask >> k) & 1:
                    delta[k].zero_()
        moved_flat = moved.reshape(N * M, sd)
        dW, db, g_in, g_out = ops.pf_heads_weight_grads(act, delta, moved_flat, d_ll.reshape(plan.K, N * M))
        K = plan.K
        # per-trajectory row gradients: the mid layer's delta summed over the particles (chunk-major planes -> (N, 64))
        mid_layer = mid_layer_index(plan.heads[0])
        d_rows = delta[:, mid_layer].view(K, U // 4, N, M, 4).sum(dim=3).permute(0, 2, 1, 3).reshape(K, N, U)
        # this step's parameter gradients as ONE flat tensor, per head [dW | db | g_in | g_out | sum d_ll] (HeadGradToken)
        d_token = torch.cat([dW.reshape(K, -1), db.reshape(K, -1), g_in.reshape(K, -1), g_out.reshape(K, -1),
                             d_ll.reshape(K, -1).sum(dim=1, keepdim=True)], dim=1).reshape(-1)
        return (None, None, None, None, d_rows, None, None, d_token, None)


class Reweight(torch.autograd.Function):
    """(ll[K, N, M], modality log-weights (N, K) | None, logw_in (N, M), moved states) -> (normalised logw (N, M), estimate
    (N, sd)): fusion over the enabled heads, reweighting, normalisation and the weighted-average estimate of A.3 in one
    kernel, and their reverse mode in one more (``mmf_pf_reweight_train_fwd / _bwd``; the backward recomputes instead of
    saving intermediates).  The particle states carry no gradient (frozen dynamics)."""

    @staticmethod
    def forward(ctx, ll, modw, logw_in, states, enabled_mask):
        ll, logw_in, states = ll.contiguous(), logw_in.contiguous(), states.detach().contiguous()
        modw = None if modw is None else modw.contiguous()
        logw, est = ops.pf_reweight_train_fwd(ll.detach(), None if modw is None else modw.detach(), logw_in.detach(), states,
                                              enabled_mask)
        ctx.mask = enabled_mask
        ctx.has_modw = modw is not None
        ctx.save_for_backward(ll, logw_in, states, *([modw] if modw is not None else []))
        ctx.set_materialize_grads(False)
        return logw, est

    @staticmethod
    def backward(ctx, d_logw, d_est):
        ll, logw_in, states, *rest = ctx.saved_tensors
        modw = rest[0] if ctx.has_modw else None
        d_ll, d_w, d_in = ops.pf_reweight_train_bwd(ll, modw, logw_in, states, ctx.mask, d_est, d_logw)
        return d_ll, d_w, d_in, None, None


def fused_train_applicable(filt, plan, resample: bool) -> bool:
    """The fused training step covers exactly the reference's setting: no resampling, frozen dynamics,
    particle states without gradient, tensor-core precision."""
    if plan is None or resample or filt.precision not in ("bf16x3", "bf16"):
        return False
    if filt.particle_states.requires_grad:
        return False
    return not any(p.requires_grad for p in filt.dynamics_model.parameters())
