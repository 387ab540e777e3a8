"""Multi-GPU plumbing: one process per GPU, coordinate batches sharded across ranks, ONE
all-reduce (sum) of a flat fp32 gradient buffer per step.  The reference has no distributed code; this is the
data-parallel scheme the north star asks for.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the
only communication layer; the kernels take `n_global` so each shard's gradients are already scaled for the
global mean and a plain sum reproduces the single-GPU gradient.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, near-equal split of n samples; the first n % world ranks take one extra."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def attach_flat_grads(model: torch.nn.Module, align: int = 64, last=()) -> torch.Tensor:
    """Make every parameter's .grad a view into one zero-filled flat buffer (offsets aligned to `align` floats so
    vector reductions stay 16-byte aligned).  Returns the flat buffer: zero it once per step, all-reduce it once.
    Parameters listed in `last` are placed at the end of the buffer (see replicated_numel)."""
    last_ids = {id(p) for p in last}
    ps = [p for p in model.parameters() if p.requires_grad and id(p) not in last_ids]
    ps += [p for p in last if p.requires_grad]
    offs, total = [], 0
    for p in ps:
        offs.append(total)
        total += (p.numel() + align - 1) // align * align
    flat = torch.zeros(total, dtype=torch.float32, device=ps[0].device)
    for p, o in zip(ps, offs):
        p.grad = flat[o:o + p.numel()].view_as(p)
    flat.replicated_numel = offs[len(ps) - len([p for p in last if p.requires_grad])] if last else total
    return flat


def t_slab(t_resolution: int, rank: int, world: int) -> Tuple[int, int]:
    """Frames [lo, hi) of the 3-D sparse grid owned by `rank`.  SparseGrid.forward indexes the grid by the NEAREST
    frame only (sparsegrid.py:43-46,65), so a rank whose samples all have t inside its slab reads and updates that
    slab alone: the sparse grid (80 % of all parameters) needs no collective, only keyframes + MLP are all-reduced."""
    return shard_range(t_resolution, rank, world)


def all_reduce_grads(flat: torch.Tensor, group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)


def broadcast_parameters(model: torch.nn.Module, src: int = 0, group=None) -> None:
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        for p in model.parameters():
            dist.broadcast(p.data, src, group=group)


class GridFirstAllReduce:
    """The step's ONE logical gradient all-reduce, issued in two pieces so that most of it overlaps compute.

    The flat gradient buffer starts with the grid gradients (keyframe planes, and the sparse grid unless it is slab-owned).
    The fused step runs the grid scatter-add BEFORE the weight-gradient kernel and records `event` in between
    (nvp_record_grid_grads_event), so the grid piece (99.6 % of the bytes) is reduced while wgrad computes; the MLP
    piece follows.  Opt-in (bench.py --overlap): measured on 2 GPUs only (3.84 -> 3.73 ms/step).

        ar = GridFirstAllReduce(flat[:flat.replicated_numel], grid_numel)
        model.fwd_loss_bwd(x, gt, n_global=N, loss_sum=ls, grid_event=ar.event)
        ar.run()                     # returns with the main stream ordered after both pieces
    """

    def __init__(self, reduce_view: torch.Tensor, grid_numel: int, group=None):
        self.grid = reduce_view[:grid_numel]
        self.rest = reduce_view[grid_numel:]
        self.group = group
        self.cuda = reduce_view.is_cuda          # CPU tensors (gloo tests): same two pieces, no streams / events
        self.event = torch.cuda.Event() if self.cuda else None
        self.side = torch.cuda.Stream(reduce_view.device) if self.cuda else None

    def run(self) -> None:
        # Both pieces are issued with async_op=True so that both run on the process group's own NCCL stream, in this
        # order on every rank.  (A synchronous collective may be enqueued on the *calling* stream instead; mixed with an
        # asynchronous one that lets two kernels of one communicator run concurrently, which NCCL does not allow - the
        # first version of this class did that and hung at 8 GPUs once the grid piece outlasted the wgrad kernel.)
        if self.cuda:
            self.side.wait_event(self.event)
            with torch.cuda.stream(self.side):
                w_grid = dist.all_reduce(self.grid, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            w_grid = dist.all_reduce(self.grid, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        w_rest = dist.all_reduce(self.rest, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self.rest.numel() else None
        w_grid.wait()
        if w_rest is not None:
            w_rest.wait()


def grid_grad_numel(model: torch.nn.Module, flat: torch.Tensor) -> int:
    """Number of leading elements of a flat gradient buffer (attach_flat_grads / optim.flatten_parameters order) that
    belong to grid parameters: everything before the first non-grid parameter's gradient."""
    grid_ids = {id(model.keyframes_xy.params), id(model.keyframes_yt.params), id(model.keyframes_xt.params),
                id(model.sparse_grid.embeddings)}
    base = flat.data_ptr()
    firsts = [(p.grad.data_ptr() - base) // 4 for p in model.parameters() if id(p) not in grid_ids and p.grad is not None]
    return min(firsts) if firsts else flat.numel()


class ShardedStep:
    """Runs the fused step on this rank's shard of a GLOBAL batch that every rank holds (or can slice).

    step(model_input, gt_u8): inputs are the global batch [1, N, ...]; the rank processes samples
    shard_range(N, rank, world), accumulates into the flat gradient buffer and all-reduces it (and the loss sum).
    """

    def __init__(self, model, group=None):
        self.model, self.group = model, group
        self.flat = attach_flat_grads(model)
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.loss_sum: Optional[torch.Tensor] = None

    def step(self, model_input: Dict[str, torch.Tensor], gt_u8: torch.Tensor) -> torch.Tensor:
        coords = model_input["all_coords"].reshape(-1, 3)
        tsteps = model_input["temporal_steps"].reshape(-1)
        gt = gt_u8.reshape(-1, 3)
        n = coords.shape[0]
        lo, hi = shard_range(n, self.rank, self.world)
        if self.loss_sum is None:
            self.loss_sum = torch.zeros(1, dtype=torch.float32, device=coords.device)
        self.flat.zero_()
        self.loss_sum.zero_()
        if hi > lo:
            self.model.fwd_loss_bwd({"all_coords": coords[lo:hi], "temporal_steps": tsteps[lo:hi]}, gt[lo:hi],
                                    n_global=n, loss_sum=self.loss_sum)
        all_reduce_grads(self.flat, self.group)
        all_reduce_grads(self.loss_sum, self.group)
        return self.loss_sum / (3.0 * n)
