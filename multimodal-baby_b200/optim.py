"""Fused AdamW for the trainable head parameters (the optimiser of
multimodal_lit.py:112-128).  Same update rule and defaults as `torch.optim.AdamW`; one kernel launch
per parameter tensor (`cvcl_adamw_step`), moments kept in fp32 next to the parameters."""
from __future__ import annotations

import torch

from . import _cabi, ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("FusedAdamW handles fp32 CUDA parameters only (no CPU path)")
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.contiguous()
                if not p.is_contiguous():
                    raise RuntimeError("FusedAdamW needs contiguous parameters")
                _cabi.call("cvcl_adamw_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                           st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1), float(b2),
                           float(group["eps"]), float(group["weight_decay"]), int(st["step"]), 1.0, None,
                           ops._stream())
        return loss
