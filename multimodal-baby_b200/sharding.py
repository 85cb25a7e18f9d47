"""Global-batch InfoNCE sharded by pairs over the ranks of a torch.distributed group
(the reference itself is single-process -- its oracle for this path is the
single-process loss on the concatenated batch).

Rank r owns pairs [r*b, (r+1)*b).  One exchange step: all-gather of the L2-normalised bf16
features.  Each rank then evaluates its ROW block  S[R, :] = e^s I_R T_all^T  (image->text CE of
its b images) and its COLUMN block  S[:, R]^T = e^s T_R I_all^T  (text->image CE of its b texts);
the positives sit at column offset r*b.  Backward: all-gather the 2*B fp32 log-sum-exps, then
  dI_R = e^s G[R, :] T_all,   dT_R = e^s G[:, R]^T I_all,   G = (P_row + P_col - 2 I) / (2B)
so no reduce-scatter of feature gradients is needed; d s is all-reduced.

The collectives live here; the arithmetic is injected (`compute_fwd` / `compute_bwd`): the
product passes the CUDA ops of `ops.py`, the CPU (gloo) tests pass a torch restatement so the
orchestration (offsets, gathers, reductions) is covered without a GPU.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def group_info(group):
    if group is None or not dist.is_available() or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def all_gather_rows(t: torch.Tensor, group, world: int) -> torch.Tensor:
    """[b, ...] per rank -> [world*b, ...] in rank order (identity when world == 1)."""
    if world == 1:
        return t
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


def infonce_forward(img, txt, log_scale, group, compute_fwd):
    """img, txt: local [b, E] operands.  Returns (out5 global [>=5], saved-for-backward tuple,
    local argmax pair)."""
    world, rank = group_info(group)
    b = img.shape[0]
    img_all = all_gather_rows(img, group, world)
    txt_all = all_gather_rows(txt, group, world)
    Bg = world * b
    out5, lse0, lse1, a0, a1 = compute_fwd(img, txt_all, txt, img_all, log_scale, rank * b, 1.0 / Bg)
    out5 = out5.clone()
    if world > 1:
        dist.all_reduce(out5, group=group)       # partial sums already scaled by 1/B_global
    saved = (img, txt, img_all, txt_all, lse0, lse1, log_scale, rank, b, world)
    return out5, saved, (a0, a1)


def infonce_backward(saved, group, compute_bwd):
    """-> (dimg [b,E], dtxt [b,E], dscale [1]) for upstream gradient 1."""
    img, txt, img_all, txt_all, lse0, lse1, log_scale, rank, b, world = saved
    lse0_all = all_gather_rows(lse0, group, world)
    lse1_all = all_gather_rows(lse1, group, world)
    Bg = world * b
    dimg, dtxt, ds = compute_bwd(img, txt_all, txt, img_all, log_scale, rank * b, 0.5 / Bg,
                                 lse0, lse1_all, lse1, lse0_all)
    if world > 1:
        dist.all_reduce(ds, group=group)
    return dimg, dtxt, ds


def allreduce_gradients(params, group=None):
    """Sum the per-rank partial gradients of replicated parameters (head weight/bias, embedding
    table): the loss is the GLOBAL-batch loss on every rank, so the correct reduction is SUM, not
    the mean a DDP wrapper would apply.  One flat all-reduce (about 9 MB for CVCL)."""
    world, _ = group_info(group)
    grads = [p.grad for p in params if p.grad is not None]
    if world == 1 or not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
