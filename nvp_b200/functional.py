"""Autograd bridge between torch tensors and the C ABI (include/nvp_b200.h).

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic of the path
runs in libnvp_b200.so.  There is no CPU / eager fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib

# Order of the flat parameter list handed to NvpFunction (matches nvp_params field order).
PARAM_ORDER = (
    "kf_xy", "kf_yt", "kf_xt", "sparse",
    "siren_w0", "siren_w1", "siren_w2", "siren_b0", "siren_b1", "siren_b2",
    "last_w", "last_b",
    "mod_w0", "mod_w1", "mod_w2", "mod_b0", "mod_b1", "mod_b2",
)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def pack_ptrs(ts: Sequence[Optional[torch.Tensor]]) -> _lib.NvpPtrs:
    p = _lib.NvpPtrs()
    p.kf_xy, p.kf_yt, p.kf_xt, p.sparse = _ptr(ts[0]), _ptr(ts[1]), _ptr(ts[2]), _ptr(ts[3])
    for i in range(3):
        p.siren_w[i] = _ptr(ts[4 + i])
        p.siren_b[i] = _ptr(ts[7 + i])
        p.mod_w[i] = _ptr(ts[12 + i])
        p.mod_b[i] = _ptr(ts[15 + i])
    p.last_w, p.last_b = _ptr(ts[10]), _ptr(ts[11])
    return p


def _require_cuda(name: str, t: torch.Tensor, dtype) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} is a CPU tensor: nvp_b200 has no CPU path (move the model and inputs to CUDA)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _check_out(name: str, t: Optional[torch.Tensor], like: torch.Tensor, numel: int) -> None:
    """Output / accumulation buffers are handed to the library as raw pointers: validate what it cannot."""
    if t is None:
        return
    if not t.is_cuda or t.device != like.device:
        raise RuntimeError(f"{name} must live on {like.device} (got {t.device}); nvp_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")
    if t.numel() != numel:
        raise ValueError(f"{name} has {t.numel()} elements, expected {numel}")


def _check_grads(params: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]]) -> None:
    for i, (p, g) in enumerate(zip(params, grads)):
        _check_out(f"gradient of {PARAM_ORDER[i]}", g, p, p.numel())


class _Workspace:
    """Grow-only per-device scratch handed to the library (the library never allocates)."""

    def __init__(self):
        self.buf: Dict[torch.device, torch.Tensor] = {}

    def get(self, device, nbytes: int) -> torch.Tensor:
        nbytes = max(int(nbytes), 256)
        b = self.buf.get(device)
        if b is None or b.numel() < nbytes:
            self.buf[device] = b = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return b


WORKSPACE = _Workspace()
_last_launches = 0


def last_launch_count() -> int:
    return _last_launches


def _note_launches():
    global _last_launches
    _last_launches = _lib.load().nvp_last_launch_count()


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def forward(desc: _lib.NvpDesc, params: Sequence[torch.Tensor], coords: torch.Tensor, tsteps: torch.Tensor,
            mode: int) -> torch.Tensor:
    coords = _require_cuda("all_coords", coords, torch.float32)
    tsteps = _require_cuda("temporal_steps", tsteps, torch.float32)
    params = [_require_cuda(f"parameter {PARAM_ORDER[i]}", p.detach(), torch.float32) for i, p in enumerate(params)]
    n = coords.shape[0]
    out = torch.empty((n, 3), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(coords.device):
        ws = WORKSPACE.get(coords.device, _lib.workspace_bytes(desc, n, mode, 0))
        pp = pack_ptrs(params)
        rc = _lib.load().nvp_forward(C.byref(desc), C.byref(pp), coords.data_ptr(), tsteps.data_ptr(), n,
                                     out.data_ptr(), ws.data_ptr(), ws.numel(), mode, _stream_ptr(coords.device))
    _lib.check(rc, "nvp_forward")
    _note_launches()
    return out


def backward(desc: _lib.NvpDesc, params: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]],
             coords: torch.Tensor, tsteps: torch.Tensor, dout: torch.Tensor, mode: int) -> None:
    """grads[i] += d(sum(dout*rgb))/d(params[i]) for every non-None grads[i]."""
    coords = _require_cuda("all_coords", coords, torch.float32)
    tsteps = _require_cuda("temporal_steps", tsteps, torch.float32)
    dout = _require_cuda("grad_output", dout, torch.float32)
    params = [_require_cuda(f"parameter {PARAM_ORDER[i]}", p.detach(), torch.float32) for i, p in enumerate(params)]
    _check_grads(params, grads)
    n = coords.shape[0]
    with torch.cuda.device(coords.device):
        ws = WORKSPACE.get(coords.device, _lib.workspace_bytes(desc, n, mode, 1))
        pp, gg = pack_ptrs(params), pack_ptrs(grads)
        rc = _lib.load().nvp_backward(C.byref(desc), C.byref(pp), coords.data_ptr(), tsteps.data_ptr(),
                                      dout.data_ptr(), n, C.byref(gg), ws.data_ptr(), ws.numel(), mode,
                                      _stream_ptr(coords.device))
    _lib.check(rc, "nvp_backward")
    _note_launches()


def fwd_loss_bwd(desc: _lib.NvpDesc, params: Sequence[torch.Tensor], grads: Sequence[Optional[torch.Tensor]],
                 coords: torch.Tensor, tsteps: torch.Tensor, gt_u8: torch.Tensor, n_global: int,
                 loss_sum: torch.Tensor, mode: int, out_rgb: Optional[torch.Tensor] = None,
                 grid_event: Optional[torch.cuda.Event] = None) -> None:
    """One fused step: loss_sum[0] += sum((rgb-gt)^2); grads += d(mean over 3*n_global)/d(params).
    grid_event (optional) is recorded on the current stream as soon as the grid gradients are final, before the MLP
    weight gradients are computed (nvp_record_grid_grads_event): the multi-GPU host starts its collective there."""
    coords = _require_cuda("all_coords", coords, torch.float32)
    tsteps = _require_cuda("temporal_steps", tsteps, torch.float32)
    gt_u8 = _require_cuda("img", gt_u8, torch.uint8)
    params = [_require_cuda(f"parameter {PARAM_ORDER[i]}", p.detach(), torch.float32) for i, p in enumerate(params)]
    n = coords.shape[0]
    _check_grads(params, grads)
    _check_out("loss_sum", loss_sum, coords, 1)
    _check_out("out_rgb", out_rgb, coords, 3 * n)
    with torch.cuda.device(coords.device):
        ws = WORKSPACE.get(coords.device, _lib.workspace_bytes(desc, n, mode, 1))
        pp, gg = pack_ptrs(params), pack_ptrs(grads)
        if grid_event is not None:
            if not grid_event.cuda_event:                    # torch creates the handle lazily
                grid_event.record(torch.cuda.current_stream(coords.device))
            _lib.check(_lib.load().nvp_record_grid_grads_event(grid_event.cuda_event), "nvp_record_grid_grads_event")
        rc = _lib.load().nvp_fwd_loss_bwd(C.byref(desc), C.byref(pp), coords.data_ptr(), tsteps.data_ptr(),
                                          gt_u8.data_ptr(), n, n_global, C.byref(gg), loss_sum.data_ptr(),
                                          _ptr(out_rgb), ws.data_ptr(), ws.numel(), mode, _stream_ptr(coords.device))
    _lib.check(rc, "nvp_fwd_loss_bwd")
    _note_launches()


def encode_latent(desc: _lib.NvpDesc, params: Sequence[torch.Tensor], coords: torch.Tensor) -> torch.Tensor:
    coords = _require_cuda("all_coords", coords, torch.float32)
    n = coords.shape[0]
    zdim = _lib.load().nvp_latent_dim(C.byref(desc))
    z = torch.empty((n, zdim), dtype=torch.float32, device=coords.device)
    ps: List[Optional[torch.Tensor]] = [_require_cuda("grid parameter", p.detach(), torch.float32) for p in params[:4]]
    ps += [None] * (len(PARAM_ORDER) - 4)
    with torch.cuda.device(coords.device):
        pp = pack_ptrs(ps)
        rc = _lib.load().nvp_encode_latent(C.byref(desc), C.byref(pp), coords.data_ptr(), n, z.data_ptr(),
                                           _stream_ptr(coords.device))
    _lib.check(rc, "nvp_encode_latent")
    _note_launches()
    return z


class NvpFunction(torch.autograd.Function):
    """rgb = NVP.forward(coords, tsteps; params) with the reference's autograd contract
    (gradients for the 18 parameter tensors, none for the inputs — SURVEY.md 3.3).

    direct=True (the module's default, NVP.direct_grad_accumulation): backward adds each parameter's gradient straight into
    its `.grad` (created zero-filled when missing) and returns None for it, which is what `loss.backward()` would end up
    with -- without 18 temporary tensors (543 MB for config S), their clearing and autograd's second accumulation pass.
    Set the module flag to False when gradients must be RETURNED (torch.autograd.grad, hooks, higher-order use)."""

    @staticmethod
    def forward(ctx, desc, mode, direct, coords, tsteps, *params):
        ctx.desc, ctx.mode, ctx.direct = desc, mode, direct
        ctx.params = params if direct else None          # the leaf Parameters themselves (their .grad is the target)
        ctx.save_for_backward(coords, tsteps, *params)
        return forward(desc, params, coords, tsteps, mode)

    @staticmethod
    def backward(ctx, dout):
        coords, tsteps, *params = ctx.saved_tensors
        needs = ctx.needs_input_grad[5:]
        if ctx.direct:
            grads = []
            for p, need in zip(ctx.params, needs):
                if not need:
                    grads.append(None)
                    continue
                g = p.grad
                if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.shape != p.shape or g.device != p.device:
                    new = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    if g is not None:
                        new.copy_(g)
                    p.grad = g = new
                grads.append(g)
            backward(ctx.desc, params, grads, coords, tsteps, dout.reshape(-1, 3), ctx.mode)
            return (None, None, None, None, None) + (None,) * len(params)
        grads = [torch.zeros_like(p) if need else None for p, need in zip(params, needs)]
        backward(ctx.desc, params, grads, coords, tsteps, dout.reshape(-1, 3), ctx.mode)
        return (None, None, None, None, None, *grads)


def scatter_latent(desc: _lib.NvpDesc, grads: Sequence[Optional[torch.Tensor]], coords: torch.Tensor, dz: torch.Tensor) -> None:
    """grads[0..3] (kf_xy, kf_yt, kf_xt, sparse; None = skip) += backward of encode_latent for upstream dz [N, Z]."""
    coords = _require_cuda("all_coords", coords, torch.float32)
    dz = _require_cuda("grad of the latent", dz, torch.float32)
    gs: List[Optional[torch.Tensor]] = list(grads[:4]) + [None] * (len(PARAM_ORDER) - 4)
    with torch.cuda.device(coords.device):
        gg = pack_ptrs(gs)
        rc = _lib.load().nvp_scatter_latent(C.byref(desc), coords.data_ptr(), coords.shape[0], dz.data_ptr(), C.byref(gg),
                                            _stream_ptr(coords.device))
    _lib.check(rc, "nvp_scatter_latent")
    _note_launches()


class LatentFunction(torch.autograd.Function):
    """z = [DG_xy | DG_yt | DG_xt | SG](coords) with gradients for the four grid tensors (none for coords)."""

    @staticmethod
    def forward(ctx, desc, coords, kf_xy, kf_yt, kf_xt, sparse):
        ctx.desc = desc
        ctx.save_for_backward(coords, kf_xy, kf_yt, kf_xt, sparse)
        return encode_latent(desc, [kf_xy, kf_yt, kf_xt, sparse], coords)

    @staticmethod
    def backward(ctx, dz):
        coords, *grids = ctx.saved_tensors
        grads = [torch.zeros_like(p) if ctx.needs_input_grad[2 + i] else None for i, p in enumerate(grids)]
        scatter_latent(ctx.desc, grads, coords, dz.contiguous())
        return (None, None, *grads)
