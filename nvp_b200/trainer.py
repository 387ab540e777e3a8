"""The per-step device work of the reference's training loop (training.py:42-76) as one object, on one or several GPUs.

    trainer = FusedTrainer(model, lr=1e-2, total_steps=epochs, distributed=True, fused_optimizer=True)
    loss = trainer.step(model_input, gt_u8)        # fwd + image_mse + bwd (+ all-reduce) + AdamW + cosine schedule
    sd = trainer.model_state_dict()                # complete on every rank (owned t-slabs are gathered first)

Multi-GPU scheme (the reference is single-GPU): one process per GPU, the GLOBAL batch of the reference's
sampler is split by ownership of the 3-D grid's frames.  SparseGrid.forward reads the grid at the NEAREST frame only
(sparsegrid.py:43-46,65), so rank r owns frames t_slab(T, r, G) of `sparse_grid.embeddings` -- parameters, gradient and
AdamW moments -- and takes exactly the samples whose nearest frame lies in its slab.  The union over ranks is the
reference's batch (same sample set as one GPU: PSNR at equal steps is unchanged), the sparse grid needs no collective
and 1/G of the optimiser traffic, and the one all-reduce per step carries the keyframe planes and the MLP only.
`model_state_dict()` / `sync_slabs()` make every rank's copy of the grid complete again (checkpoints: training.py:36,66,90).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

from . import dist as nvp_dist
from .optim import FusedAdamW, flatten_parameters


def nearest_frame(t_coord: torch.Tensor, t_resolution: int) -> torch.Tensor:
    """Frame index SparseGrid.forward reads for a temporal coordinate (sparsegrid.py:43-46): clamp(trunc((T-1) c + 0.5)),
    fp32 multiply then add, as the reference computes it."""
    f = (t_coord.to(torch.float32) * float(t_resolution - 1)) + 0.5
    return f.to(torch.int64).clamp_(0, t_resolution - 1)


def route_to_slab(model_input: Dict[str, torch.Tensor], gt_u8: torch.Tensor, t_resolution: int, rank: int, world: int):
    """The samples of a global batch whose nearest grid frame belongs to `rank`'s slab (order preserved)."""
    coords = model_input["all_coords"].reshape(-1, 3)
    lo, hi = nvp_dist.t_slab(t_resolution, rank, world)
    frame = nearest_frame(coords[:, 0], t_resolution)
    keep = (frame >= lo) & (frame < hi)
    return ({"all_coords": coords[keep][None], "temporal_steps": model_input["temporal_steps"].reshape(-1)[keep][None]},
            gt_u8.reshape(-1, 3)[keep][None])


class FusedTrainer:
    def __init__(self, model, lr: float, total_steps: int, distributed: Optional[bool] = None, fused_optimizer: bool = True,
                 weight_decay: float = 1e-3, eta_min: float = 1e-5, group=None):
        self.model, self.group = model, group
        ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        self.distributed = ddp if distributed is None else (distributed and ddp)
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.t_resolution = int(model.sparse_grid.t_resolution)
        emb = model.sparse_grid.embeddings
        if self.distributed:
            nvp_dist.broadcast_parameters(model, 0, group)
        # one flat parameter / gradient buffer; the slab-owned sparse grid sits at its end, outside the all-reduce
        self.flat_params, self.flat_grads = flatten_parameters(model, last=[emb] if self.distributed else ())
        n_rep = self.flat_grads.replicated_numel
        self.reduce_view = self.flat_grads[:n_rep]
        self.slab = nvp_dist.t_slab(self.t_resolution, self.rank, self.world)
        per_frame = emb.numel() // self.t_resolution
        self._emb_off = emb.data_ptr() - self.flat_params.data_ptr()
        assert self._emb_off % 4 == 0
        self._emb_off //= 4
        self._owned = (self._emb_off + self.slab[0] * per_frame, self._emb_off + self.slab[1] * per_frame)
        self.loss_sum = torch.zeros(1, dtype=torch.float32, device=self.flat_params.device)
        self.fused_optimizer = fused_optimizer
        self.total_steps = total_steps
        if fused_optimizer:
            kw = dict(lr=lr, weight_decay=weight_decay, t_max=total_steps, eta_min=eta_min)
            if self.distributed:
                # replicated parameters on every rank + the owned slab only: 1/G of the sparse grid's optimiser traffic
                a, b = self._owned
                self.optimizers = [FusedAdamW(self.flat_params[:n_rep], self.flat_grads[:n_rep], **kw),
                                   FusedAdamW(self.flat_params[a:b], self.flat_grads[a:b], **kw)]
            else:
                self.optimizers = [FusedAdamW(self.flat_params, self.flat_grads, **kw)]
            self.torch_optim = self.scheduler = None
        else:
            self.optimizers = []
            self.torch_optim = torch.optim.AdamW(lr=lr, params=model.parameters(), weight_decay=weight_decay)
            self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(self.torch_optim, T_max=total_steps, eta_min=eta_min)
            self.flat_grads.zero_()
        self._synced = True

    # ------------------------------------------------------------------------------------------
    def current_lr(self) -> float:
        return self.optimizers[0].current_lr() if self.fused_optimizer else float(self.scheduler.get_last_lr()[0])

    def zero_grads(self) -> None:
        """Clear what this rank accumulates into: everything on one GPU; with slab ownership the replicated gradients and
        the owned frames only (the other frames never receive a gradient on this rank: the routing and the kernels compute
        the same nearest frame)."""
        if self.distributed:
            self.reduce_view.zero_()
            a, b = self._owned
            self.flat_grads[a:b].zero_()
        else:
            self.flat_grads.zero_()

    def local_batch(self, model_input, gt_u8):
        """This rank's share of a GLOBAL batch (identity on one GPU)."""
        if not self.distributed:
            return model_input, gt_u8
        return route_to_slab(model_input, gt_u8, self.t_resolution, self.rank, self.world)

    def step(self, model_input, gt_u8, n_global: Optional[int] = None, routed: bool = False) -> torch.Tensor:
        """One optimisation step on a GLOBAL batch (or on this rank's already routed share, with n_global given).
        Returns the global image_mse as a 1-element device tensor (no host synchronisation)."""
        if n_global is None:
            n_global = model_input["all_coords"].reshape(-1, 3).shape[0]
        if self.distributed and not routed:
            model_input, gt_u8 = self.local_batch(model_input, gt_u8)
        self.loss_sum.zero_()
        if model_input["all_coords"].numel() > 0:
            self.model.fwd_loss_bwd(model_input, gt_u8, n_global=n_global, loss_sum=self.loss_sum)
        if self.distributed:
            dist.all_reduce(self.reduce_view, group=self.group)
            dist.all_reduce(self.loss_sum, group=self.group)
        if self.fused_optimizer:
            # the gradient clear is folded into the update.  (Frames this rank does not own never receive a gradient: the
            # routing above and the kernels compute the same nearest frame, so nothing there needs clearing.)
            for o in self.optimizers:
                o.step(zero_grad=True)
        else:
            self.torch_optim.step()
            self.scheduler.step()
            self.flat_grads.zero_()
        self._synced = not self.distributed
        return self.loss_sum / (3.0 * n_global)

    # ------------------------------------------------------------------------------------------
    def sync_slabs(self) -> None:
        """Every rank's copy of the 3-D grid becomes complete: each rank broadcasts the frames it owns."""
        if not self.distributed or self._synced:
            return
        emb = self.model.sparse_grid.embeddings.data
        for r in range(self.world):
            lo, hi = nvp_dist.t_slab(self.t_resolution, r, self.world)
            if hi > lo:
                dist.broadcast(emb[lo:hi], r, group=self.group)
        self._synced = True

    def model_state_dict(self):
        self.sync_slabs()
        return self.model.state_dict()

    def optimizer_state_dict(self) -> dict:
        if not self.fused_optimizer:
            return {"optimizer": self.torch_optim.state_dict(), "scheduler": self.scheduler.state_dict()}
        o = self.optimizers[0]
        return {"optimizer": {"fused_adamw": True, "step": o.t, "lr": o.base_lr, "betas": o.betas, "eps": o.eps,
                              "weight_decay": o.wd, "exp_avg": [x.m for x in self.optimizers],
                              "exp_avg_sq": [x.v for x in self.optimizers], "slab": self.slab, "world": self.world},
                "scheduler": {"T_max": o.t_max, "eta_min": o.eta_min, "last_epoch": o.t, "_last_lr": [o.current_lr()]}}


def psnr_from_loss(loss_value: float) -> float:
    """training.py:58 (peak^2 = 4 for a signal in [-1, 1])."""
    return 10 * math.log10(4 / loss_value)
