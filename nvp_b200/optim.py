"""Fused AdamW over a flat parameter buffer.

Drop-in for the reference's `torch.optim.AdamW(lr, weight_decay=1e-3)` + `CosineAnnealingLR(T_max, eta_min=1e-5)` +
`optim.zero_grad()` (training.py:13-14,73-76): one kernel reads p/g/m/v once, writes p/m/v and clears g.  All model
parameters are re-seated as views into ONE flat fp32 buffer (and their gradients into another), which is also what
the multi-GPU all-reduce wants.
"""
from __future__ import annotations

import math

import torch

from . import _lib


def flatten_parameters(model: torch.nn.Module, align: int = 64, last=()):
    """Re-seat every parameter (and its .grad) as a view into flat fp32 buffers.  Returns (flat_params, flat_grads).
    Parameters in `last` go to the end (see dist.attach_flat_grads)."""
    last_ids = {id(p) for p in last}
    ps = [p for p in model.parameters() if id(p) not in last_ids] + list(last)
    offs, total = [], 0
    for p in ps:
        offs.append(total)
        total += (p.numel() + align - 1) // align * align
    dev = ps[0].device
    flat_p = torch.zeros(total, dtype=torch.float32, device=dev)
    flat_g = torch.zeros(total, dtype=torch.float32, device=dev)
    for p, o in zip(ps, offs):
        flat_p[o:o + p.numel()].copy_(p.data.reshape(-1))
        p.data = flat_p[o:o + p.numel()].view_as(p)
        p.grad = flat_g[o:o + p.numel()].view_as(p)
    n_last = sum(1 for _ in last)
    flat_g.replicated_numel = offs[len(ps) - n_last] if n_last else total
    return flat_p, flat_g


class FusedAdamW:
    """AdamW(betas=(0.9,0.999), eps=1e-8, weight_decay) with an optional cosine schedule, one kernel per step."""

    def __init__(self, flat_params, flat_grads, lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-3,
                 t_max=None, eta_min=1e-5):
        assert flat_params.is_cuda and flat_params.dtype == torch.float32 and flat_params.shape == flat_grads.shape
        self.p, self.g = flat_params, flat_grads
        self.m = torch.zeros_like(flat_params)
        self.v = torch.zeros_like(flat_params)
        self.base_lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.t_max, self.eta_min = t_max, eta_min
        self.t = 0

    def current_lr(self) -> float:
        """Closed form of CosineAnnealingLR after self.t scheduler steps."""
        if self.t_max is None:
            return self.base_lr
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.t / self.t_max)) / 2

    def step(self, zero_grad: bool = True) -> None:
        lr = self.current_lr()
        self.t += 1
        with torch.cuda.device(self.p.device):
            rc = _lib.load().nvp_adamw_step(self.p.data_ptr(), self.g.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                            self.p.numel(), lr, self.betas[0], self.betas[1], self.eps, self.wd, self.t,
                                            1 if zero_grad else 0, torch.cuda.current_stream(self.p.device).cuda_stream)
        _lib.check(rc, "nvp_adamw_step")
