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


def _grads_for_head(spec, dW, db, g_in, g_out, d_ll):
    """dW (L, 64, 64), db (L+1, 64), g_in (64, sd), g_out (64,) from mmf_pf_heads_weight_grads; d_ll (P,).
    Returns grads in head_parameters order and the index of the mid layer."""
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
    grads += [g_out[None, :], d_ll.sum().reshape(1)]
    return grads, mid_layer


class FusedHeads(torch.autograd.Function):
    """(states, eps, dynamics row, head rows, head parameters...) -> (moved, ll[K_enabled, N, M])."""

    @staticmethod
    def forward(ctx, plan, states, eps, dyn_row, head_rows, enabled_mask, precision, *params):
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
        d_rows = torch.zeros((plan.K, N, U), device=act.device, dtype=torch.float32)
        grads = []
        for k, spec in enumerate(plan.heads):
            n_params = len(head_parameters(spec))
            if not (mask >> k) & 1:
                grads += [None] * n_params
                continue
            g, mid_layer = _grads_for_head(spec, dW[k], db[k], g_in[k], g_out[k], d_ll[k].reshape(-1))
            grads += g
            d_rows[k] = delta[k, mid_layer].view(U // 4, N, M, 4).sum(dim=2).permute(1, 0, 2).reshape(N, U)
        return (None, None, None, None, d_rows, None, None, *grads)


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
