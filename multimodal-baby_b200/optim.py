"""Fused AdamW for the trainable head parameters (the optimiser of
multimodal_lit.py:112-128).  Same update rule and defaults as `torch.optim.AdamW`, moments kept in fp32
next to the parameters.

Two entry points:
  step()        one launch per parameter tensor (`cvcl_adamw_step`), step number on the host: the eager
                optimizer.step() of a training loop;
  graph_step()  ONE launch for all tensors (`cvcl_adamw_multi_step`) with the step number kept on the
                device, so the launch can be captured in a CUDA graph and replayed
                (`GraphedContrastiveStep(..., optimizer=opt)` makes forward + backward + update one graph).

Both refresh the bf16 shadow of every parameter registered with `attach_shadow` in the same pass, so the
head GEMM never needs a separate fp32 -> bf16 cast of W (ops.weight_shadow).
"""
from __future__ import annotations

import ctypes

import torch

from . import _cabi, ops


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._shadows = {}
        self._multi = None
        self._dev_counters = None            # (step_dev, ticket): created once, outside any capture

    # ------------------------------------------------------------------------------- shadows
    def attach_shadow(self, p):
        """keep a bf16 copy of parameter `p` (the projection weight) up to date inside the update kernel
        and hand it to the ops (`ops.weight_shadow(p)` then returns it without casting)."""
        if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
            raise RuntimeError("FusedAdamW.attach_shadow needs a contiguous fp32 CUDA parameter")
        w16 = p.detach().to(torch.bfloat16).contiguous()
        self._shadows[id(p)] = w16
        ops.register_weight_shadow(p, w16)
        self._multi = None
        return w16

    def _state(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
        return st

    @staticmethod
    def _check(p):
        if not p.is_cuda or p.dtype != torch.float32:
            raise RuntimeError("FusedAdamW handles fp32 CUDA parameters only (no CPU path)")
        if not p.is_contiguous():
            raise RuntimeError("FusedAdamW needs contiguous parameters")

    # ------------------------------------------------------------------------------- eager
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
                self._check(p)
                st = self._state(p)
                st["step"] += 1
                g = p.grad.contiguous()
                sh = self._shadows.get(id(p))
                _cabi.call("cvcl_adamw_step", p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(),
                           st["exp_avg_sq"].data_ptr(), p.numel(), float(group["lr"]), float(b1), float(b2),
                           float(group["eps"]), float(group["weight_decay"]), int(st["step"]), 1.0,
                           None if sh is None else sh.data_ptr(), ops._stream())
        return loss

    # ------------------------------------------------------------------------------- graph-capturable
    def _build_multi(self):
        ps, lrs, wds = [], [], []
        betas = eps = None
        for group in self.param_groups:
            if betas is None:
                betas, eps = tuple(group["betas"]), float(group["eps"])
            elif tuple(group["betas"]) != betas or float(group["eps"]) != eps:
                raise RuntimeError("FusedAdamW.graph_step needs the same betas / eps in every group")
            for p in group["params"]:
                if p.grad is None:
                    continue
                self._check(p)
                if not p.grad.is_contiguous():
                    raise RuntimeError("FusedAdamW.graph_step needs contiguous .grad tensors")
                ps.append(p); lrs.append(float(group["lr"])); wds.append(float(group["weight_decay"]))
        if not ps or len(ps) > 8:
            raise RuntimeError("FusedAdamW.graph_step handles 1..8 tensors with gradients (got %d)" % len(ps))
        n = len(ps)
        dev = ps[0].device
        sts = [self._state(p) for p in ps]
        P = ctypes.c_void_p * n
        m = dict(
            n=n, params=ps, grads=[p.grad for p in ps],
            p=P(*[p.data_ptr() for p in ps]), g=P(*[p.grad.data_ptr() for p in ps]),
            m=P(*[s["exp_avg"].data_ptr() for s in sts]), v=P(*[s["exp_avg_sq"].data_ptr() for s in sts]),
            numel=(ctypes.c_longlong * n)(*[p.numel() for p in ps]),
            shadow=P(*[(self._shadows[id(p)].data_ptr() if id(p) in self._shadows else None) for p in ps]),
            lr=(ctypes.c_float * n)(*lrs), wd=(ctypes.c_float * n)(*wds), betas=betas, eps=eps,
            step_dev=self._counters(dev, max(s["step"] for s in sts))[0], ticket=self._counters(dev, 0)[1])
        return m

    def _counters(self, dev, step0):
        if self._dev_counters is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdamW.graph_step must run once eagerly (warm-up) before it is captured")
            self._dev_counters = (torch.tensor([step0], dtype=torch.int32, device=dev),
                                  torch.zeros(1, dtype=torch.int32, device=dev))
        return self._dev_counters

    @torch.no_grad()
    def reset_state(self):
        """zero the moments and the step counters, re-derive the bf16 shadows from the fp32 masters."""
        for group in self.param_groups:
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["exp_avg"].zero_(); st["exp_avg_sq"].zero_(); st["step"] = 0
                if id(p) in self._shadows:
                    self._shadows[id(p)].copy_(p.detach())
        if self._dev_counters is not None:
            self._dev_counters[0].zero_(); self._dev_counters[1].zero_()

    @torch.no_grad()
    def graph_step(self):
        """one launch for every tensor that has a .grad; capturable (static pointers: the .grad tensors
        must stay the same objects, as GraphedContrastiveStep's flat gradient views do)."""
        m = self._multi
        if m is None or any(p.grad is None or p.grad.data_ptr() != g.data_ptr() for p, g in zip(m["params"], m["grads"])):
            m = self._multi = self._build_multi()          # host-side pointer tables only
        _cabi.call("cvcl_adamw_multi_step", m["n"], m["p"], m["g"], m["m"], m["v"], m["numel"], m["shadow"],
                   m["lr"], m["wd"], float(m["betas"][0]), float(m["betas"][1]), float(m["eps"]), 1.0,
                   m["step_dev"].data_ptr(), m["ticket"].data_ptr(), ops._stream())

    def device_step_count(self):
        """steps performed through graph_step (device counter; synchronises)."""
        return 0 if self._multi is None else int(self._multi["step_dev"].item())
