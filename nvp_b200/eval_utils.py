"""Inference helpers — host code kept from the reference's eval script (experiment_scripts/eval.py:19-109,220-256).

  quantize_keyframes / quantize_sparse_grid   8-bit min/max quantise -> de-quantise of the learnable grids, per level /
                                              per feature, exactly as eval.py does before evaluating the codec
  render_frame                                full-frame forward in `n_slices` slices (eval.py:220-240), optionally with
                                              temporal / spatial interpolation (temporal_interp -> forward_inter kernel path)
  psnr                                        eval.py:256
All arithmetic of the model itself runs in the CUDA path; these are torch tensor ops around it.
"""
from __future__ import annotations

import math

import torch

from . import dataio

UNIT_MULTIPLIER = 2.0 ** 8 - 1.0   # eval.py: unit_multiplier


def _quant_dequant(x: torch.Tensor) -> torch.Tensor:
    lo, hi = torch.min(x), torch.max(x)
    q = (x - lo) / (hi - lo)
    q = (UNIT_MULTIPLIER * q + 0.5).to(torch.uint8)          # torch.tensor(x + 0.5, dtype=uint8): truncation
    q = torch.clamp(q, 0, 255).to(torch.float32) / UNIT_MULTIPLIER
    return (hi - lo) * q + lo


@torch.no_grad()
def quantize_keyframes(params: torch.Tensor, config: dict) -> torch.nn.Parameter:
    """eval.py:19-81: per level and per feature min/max 8-bit round trip of a flat tcnn parameter vector."""
    n_levels, dim, scale = config["n_levels"], config["n_features_per_level"], config["per_level_scale"]
    offs, total = [], 0
    for i in range(n_levels):
        a = math.exp(i * math.log(scale)) * 16 - 1           # eval.py:29 hard-codes base 16
        b = int(math.ceil(a) + 1)
        offs.append(total)
        total += b * b
    offs.append(total)
    feats = params.detach().clone().reshape(-1, dim)
    assert feats.shape[0] == total, "parameter vector does not match the level layout (no padding, compression.py:77)"
    for d in range(dim):
        for i in range(n_levels):
            feats[offs[i]:offs[i + 1], d] = _quant_dequant(feats[offs[i]:offs[i + 1], d])
    return torch.nn.Parameter(feats.reshape(-1))


@torch.no_grad()
def quantize_sparse_grid(params: torch.Tensor, config: dict) -> torch.nn.Parameter:
    """eval.py:83-109: per feature min/max 8-bit round trip of the [T,X,Y,F] grid."""
    g = params.detach().clone()
    for d in range(config["n_features_per_level"]):
        g[..., d] = _quant_dequant(g[..., d])
    return torch.nn.Parameter(g)


@torch.no_grad()
def quantize_model(model) -> None:
    """eval.py:163-179: replace the four grids by their quantised versions (attribute assignment, as the reference does)."""
    cfg = model.encoding_config
    model.keyframes_xy.params = quantize_keyframes(model.keyframes_xy.params, cfg["2d_encoding_xy"])
    model.keyframes_xt.params = quantize_keyframes(model.keyframes_xt.params, cfg["2d_encoding_xt"])
    model.keyframes_yt.params = quantize_keyframes(model.keyframes_yt.params, cfg["2d_encoding_yt"])
    model.sparse_grid.embeddings = quantize_sparse_grid(model.sparse_grid.embeddings, cfg["3d_encoding"])


@torch.no_grad()
def render_frame(model, f: int, nframes: int, resolution, org_nframes=None, temporal_interp=False, n_slices=100):
    """eval.py:220-245: RGB frame [3,H,W] in [0,1] for frame index f of an nframes-long rendering."""
    org_nframes = org_nframes or nframes
    h, w = resolution
    total = h * w
    dev = next(model.parameters()).device
    spatial = dataio.get_mgrid((h, w), dim=2).to(dev)
    half_dt = 0.5 / org_nframes
    tstep = (torch.linspace(half_dt, 1 - half_dt, nframes)[f] * torch.ones(total)).to(dev)
    tcoord = (torch.linspace(0, 1, nframes)[f] * torch.ones(total)).to(dev)
    coords = torch.cat((tcoord.unsqueeze(1), spatial), dim=1)
    out = torch.zeros(total, 3, device=dev)
    split = int(total / n_slices)
    for i in range(n_slices):
        sl = slice(i * split, (i + 1) * split)
        out[sl] = model({"all_coords": coords[None, sl], "temporal_steps": tstep[None, sl]},
                        temporal_interp=temporal_interp)["model_out"][0]
    img = out.view(h, w, 3).permute(2, 0, 1)
    return torch.clamp((img + 1) / 2, 0, 1)


def psnr(img: torch.Tensor, gt: torch.Tensor) -> float:
    """eval.py:256 (both in [0,1])."""
    return float(10 * torch.log10(1 / torch.mean((img - gt) ** 2)))
