"""CPU oracle for NVP's per-coordinate hot path.  TEST INFRASTRUCTURE ONLY.

This module is a CPU restatement (numpy index math + torch-CPU fp32/fp64 tensor math) of the
algorithm the reference runs on its hot path.  It exists to CHECK the CUDA product path; only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it.  Nothing under `nvp_b200/` imports it and the product never falls back to it.

What each function restates (citations relative to /root/reference):

  level_table            eval.py:28-35, compression.py:26-33,72,77   (level resolutions / offsets /
                         no padding; scale arithmetic in fp32 as tiny-cuda-nn grid.h does it)
  dense_grid_forward     modules.py:14-23,65-67 -> tcnn.Encoding(DenseGrid, 2-D).  The arithmetic
                         lives in the un-vendored, un-pinned fork github.com/subin-kim-cv/tiny-cuda-nn
                         (README.md:30-32); restated from upstream NVlabs/tiny-cuda-nn
                         include/tiny-cuda-nn/encodings/grid.h semantics (pos = fma(scale,x,0.5),
                         floor/fract, 4-corner bilinear, flat index modulo level size, dim-0 fastest).
  sparse_grid_forward    sparsegrid.py:23-72   (nearest voxel + 3x3 (x,y) neighbourhood, clamped)
  sparse_grid_forward_inter  sparsegrid.py:76-156 (eval-only temporal blend of two t-slices)
  modulator_forward      modulation.py:96-121  (LeakyReLU(0.01) MLP with skip-concat of the latent)
  siren_forward          modulation.py:20-56,60-92 (sin(w0*(Wx+b)), in-place gating by the modulator)
  nvp_forward            modules.py:51-84      (concat order xy, yt, xt, sparse; plane inputs
                         xy=(x,y), xt=(t,x), yt=(t,y))
  image_mse / psnr       loss_functions.py:1-5, training.py:47-48,55,58
  nvp_loss_and_grads     training.py:50-52,74  (autograd of the restated forward)
  sample_batch           dataio.py:85-99,104-120 (the coordinate sampler)

PARITY PINNING.  `oracle/check_vs_reference.py` imports the reference's own sparsegrid.py,
modulation.py, loss_functions.py, modules.py and dataio.py (unmodified, from /root/reference) and
checks this restatement against them.  The reference ships no tests, fixtures or golden vectors
(SURVEY.md section 4), and tiny-cuda-nn is absent, so:
  * SparseGrid / Modulator / SirenNet / NVP.forward glue / image_mse / sampler: PINNED against the
    reference code run in the build container (outputs and gradients compared, see that script).
  * DenseGrid interpolation arithmetic (0.5 offset, corner order, edge aliasing): PARITY UNPINNED —
    only the parameter layout is pinned by the reference (eval.py:28-35, compression.py:72,77).
    The golden vectors under tests/golden/ are generated from this restatement and become the pin.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

N_SAMPLES_PER_STEP = 1245184  # dataio.py:91


# --------------------------------------------------------------------------------------------
# DenseGrid level table
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class LevelTable:
    n_levels: int
    scales: np.ndarray   # float32 [L]
    res: np.ndarray      # int64   [L]
    offsets: np.ndarray  # int64   [L+1], in cells (multiply by F for floats)

    @property
    def n_cells(self) -> int:
        return int(self.offsets[-1])


def level_table(n_levels: int = 16, base_resolution: int = 16, per_level_scale: float = 1.35) -> LevelTable:
    """scale_l = exp2f(l*log2f(pls))*base - 1 ; res_l = ceil(scale_l)+1 ; offsets cumulative res^2.

    fp32 arithmetic as in tcnn's grid_scale()/grid_resolution(); eval.py:29-30 computes the same
    resolutions in double (exp(i*log(pls))*16-1) — both give [16,22,30,...,1443] for the shipped configs.
    """
    # fp32 steps; log2/exp2 evaluated in double and rounded once (deterministic across libms; the
    # C-ABI library's nvp_level_table() uses the identical recipe and tests compare them bit-for-bit).
    log2_pls = np.float32(np.log2(np.float64(np.float32(per_level_scale))))
    scales, res, offs = [], [], [0]
    for l in range(n_levels):
        arg = np.float32(np.float32(l) * log2_pls)
        e = np.float32(np.exp2(np.float64(arg)))
        s = np.float32(np.float32(e * np.float32(base_resolution)) - np.float32(1.0))
        r = int(math.ceil(float(s))) + 1
        scales.append(s)
        res.append(r)
        offs.append(offs[-1] + r * r)
    return LevelTable(n_levels, np.asarray(scales, np.float32), np.asarray(res, np.int64), np.asarray(offs, np.int64))


def _fmaf(a: np.ndarray, b: np.ndarray, c: float) -> np.ndarray:
    """fp32 fused multiply-add emulated through fp64 (product of two fp32 is exact in fp64)."""
    return (a.astype(np.float64) * b.astype(np.float64) + np.float64(c)).astype(np.float32)


def dense_grid_indices(u: np.ndarray, table: LevelTable):
    """Index/weight computation of the DenseGrid for inputs u [N,2] (fp32, dim 0 fastest).

    Returns idx int64 [N,L,4] (cell index INCLUDING the level offset) and w float32 [N,L,4].
    Corner c: bit0 -> +1 in dim 0, bit1 -> +1 in dim 1.  No clamping: flat index then mod res^2.
    """
    u = np.ascontiguousarray(u, np.float32)
    n = u.shape[0]
    L = table.n_levels
    idx = np.empty((n, L, 4), np.int64)
    w = np.empty((n, L, 4), np.float32)
    for l in range(L):
        s = table.scales[l]
        r = int(table.res[l])
        pos0 = _fmaf(np.full(n, s, np.float32), u[:, 0], 0.5)
        pos1 = _fmaf(np.full(n, s, np.float32), u[:, 1], 0.5)
        g0 = np.floor(pos0)
        g1 = np.floor(pos1)
        f0 = (pos0 - g0).astype(np.float32)
        f1 = (pos1 - g1).astype(np.float32)
        g0 = g0.astype(np.int64)
        g1 = g1.astype(np.int64)
        one = np.float32(1.0)
        for c in range(4):
            c0, c1 = c & 1, (c >> 1) & 1
            w0 = f0 if c0 else (one - f0)
            w1 = f1 if c1 else (one - f1)
            w[:, l, c] = w0 * w1
            flat = (g0 + c0) + (g1 + c1) * r
            idx[:, l, c] = table.offsets[l] + np.mod(flat, r * r)
    return idx, w


def dense_grid_forward(params: torch.Tensor, u: torch.Tensor, n_features: int, table: LevelTable) -> torch.Tensor:
    """params flat [n_cells*F] (level-major, cell, feature-minor); u [N,2] -> [N, L*F] (col = l*F+f)."""
    idx, w = dense_grid_indices(u.detach().cpu().numpy(), table)
    n, L = idx.shape[0], table.n_levels
    tab = params.view(-1, n_features)
    idx_t = torch.from_numpy(idx).reshape(-1)
    w_t = torch.from_numpy(w).to(params.dtype)
    vals = tab[idx_t].view(n, L, 4, n_features)
    out = (vals * w_t.unsqueeze(-1)).sum(dim=2)
    return out.reshape(n, L * n_features)


# --------------------------------------------------------------------------------------------
# SparseGrid (sparsegrid.py:23-72)
# --------------------------------------------------------------------------------------------
def sparse_grid_indices(coords: np.ndarray, t_res: int, x_res: int, y_res: int):
    """idx_d = clamp(trunc((res_d-1)*c_d + 0.5), 0, res_d-1) in fp32 (mul then add, unfused)."""
    c = np.ascontiguousarray(coords, np.float32)

    def nearest(col, res):
        f = (np.float32(res - 1) * c[:, col]).astype(np.float32)
        i = (f + np.float32(0.5)).astype(np.float32).astype(np.int64)  # trunc toward zero
        return np.clip(i, 0, res - 1)

    return nearest(0, t_res), nearest(1, x_res), nearest(2, y_res)


def sparse_grid_forward(emb: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """emb [T,X,Y,F]; coords [N,3]=(t,x,y) -> [N,9F]; col = ((i+1)*3+(j+1))*F+f, i over x, j over y."""
    T, X, Y, F = emb.shape
    ti, xi, yi = sparse_grid_indices(coords.detach().cpu().numpy(), T, X, Y)
    ti = torch.from_numpy(ti)
    feats = []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            vx = torch.from_numpy(np.clip(xi + i, 0, X - 1))
            vy = torch.from_numpy(np.clip(yi + j, 0, Y - 1))
            feats.append(emb[ti, vx, vy, :])
    return torch.cat(feats, dim=1)


def sparse_grid_forward_inter(emb: torch.Tensor, coords: torch.Tensor) -> torch.Tensor:
    """Eval-time temporal interpolation (sparsegrid.py:76-156): blend of the t-slices below / above the sample.

    Restated with the reference's quirk (sparsegrid.py:108-109): upper is normalised first and the lower coefficient
    is then divided by (NEW upper + lower); at the last frame both raw coefficients are 0 and the result is NaN,
    exactly as in the reference."""
    T, X, Y, F = emb.shape
    c = np.ascontiguousarray(coords.detach().cpu().numpy(), np.float32)
    tf = (np.float32(T - 1) * c[:, 0]).astype(np.float32)
    lower = tf.astype(np.int64)
    upper = np.clip((tf + np.float32(1)).astype(np.float32).astype(np.int64), 0, T - 1)
    up = (tf - lower.astype(np.float32)).astype(np.float32)
    lo = (upper.astype(np.float32) - tf).astype(np.float32)
    with np.errstate(invalid="ignore", divide="ignore"):
        up = (up / (up + lo)).astype(np.float32)
        lo = (lo / (up + lo)).astype(np.float32)
    _, xi, yi = sparse_grid_indices(c, T, X, Y)
    lo_t, up_t = torch.from_numpy(lo).to(emb.dtype)[:, None], torch.from_numpy(up).to(emb.dtype)[:, None]
    lower_t, upper_t = torch.from_numpy(lower), torch.from_numpy(upper)
    feats = []
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            vx = torch.from_numpy(np.clip(xi + i, 0, X - 1))
            vy = torch.from_numpy(np.clip(yi + j, 0, Y - 1))
            feats.append(emb[lower_t, vx, vy, :] * lo_t + emb[upper_t, vx, vy, :] * up_t)
    return torch.cat(feats, dim=1)


# --------------------------------------------------------------------------------------------
# Modulated SIREN (modulation.py)
# --------------------------------------------------------------------------------------------
def _operand(x: torch.Tensor, mma_dtype) -> torch.Tensor:
    """Round a GEMM operand to the tensor-core operand type (straight-through for autograd).

    mma_dtype=None is the reference arithmetic.  mma_dtype=torch.float16 restates the product's
    NVP_MODE_TC_F16 forward (fp16 operands, wide accumulation) so tests can separate arithmetic-mode
    effects (LeakyReLU sign flips of near-zero pre-activations) from kernel bugs."""
    if mma_dtype is None:
        return x
    return x + (x.to(mma_dtype).to(x.dtype) - x).detach()


def modulator_forward(z: torch.Tensor, p: Dict[str, torch.Tensor], n_layers: int = 3, mma_dtype=None):
    hs = []
    x = z
    for i in range(n_layers):
        W = p[f"wrapper.modulator.layers.{i}.0.weight"]
        b = p[f"wrapper.modulator.layers.{i}.0.bias"]
        h = torch.nn.functional.leaky_relu(_operand(x, mma_dtype) @ _operand(W, mma_dtype).t() + b, 0.01)
        hs.append(h)
        x = torch.cat((h, z), dim=1)
    return hs


def siren_forward(tau: torch.Tensor, mods, p: Dict[str, torch.Tensor], n_layers: int = 3, w0_initial: float = 30.0,
                  mma_dtype=None):
    x = tau
    for i in range(n_layers):
        W = p[f"net.layers.{i}.weight"]
        b = p[f"net.layers.{i}.bias"]
        w0 = w0_initial if i == 0 else 1.0
        if i == 0:
            x = torch.sin(w0 * (x @ W.t() + b))  # K = 1: CUDA cores in every mode
        else:
            x = torch.sin(w0 * (_operand(x, mma_dtype) @ _operand(W, mma_dtype).t() + b))
        x = x * mods[i]
    return x @ p["net.last_layer.weight"].t() + p["net.last_layer.bias"]


@dataclass
class NVPConfig:
    """The slice of config_nvp_*.json["nvp"] the hot path reads."""
    n_features: int = 2           # F of the three 2-D encodings
    n_levels: int = 16
    base_resolution: int = 16
    per_level_scale: float = 1.35
    sparse_features: int = 2      # F of the 3-D grid
    t_resolution: int = 600
    x_resolution: int = 300
    y_resolution: int = 300
    n_neurons: int = 128
    n_hidden_layers: int = 3

    @staticmethod
    def from_json(cfg: dict) -> "NVPConfig":
        e = cfg["2d_encoding_xy"]
        s = cfg["3d_encoding"]
        n = cfg["network"]
        return NVPConfig(e["n_features_per_level"], e["n_levels"], e["base_resolution"], e["per_level_scale"],
                         s["n_features_per_level"], s["t_resolution"], s["x_resolution"], s["y_resolution"],
                         n["n_neurons"], n["n_hidden_layers"])

    def to_json(self) -> dict:
        enc = {"otype": "DenseGrid", "n_levels": self.n_levels, "n_features_per_level": self.n_features,
               "log2_hashmap_size": 32, "base_resolution": self.base_resolution, "per_level_scale": self.per_level_scale}
        return {"2d_encoding_xy": dict(enc), "2d_encoding_xt": dict(enc), "2d_encoding_yt": dict(enc),
                "3d_encoding": {"otype": "SparseGrid", "n_features_per_level": self.sparse_features,
                                "x_resolution": self.x_resolution, "y_resolution": self.y_resolution,
                                "t_resolution": self.t_resolution, "upsample": False},
                "network": {"n_neurons": self.n_neurons, "n_hidden_layers": self.n_hidden_layers}}

    @property
    def table(self) -> LevelTable:
        return level_table(self.n_levels, self.base_resolution, self.per_level_scale)

    @property
    def latent_dim(self) -> int:
        return 3 * self.n_levels * self.n_features + 9 * self.sparse_features


def latent_forward(p: Dict[str, torch.Tensor], coords: torch.Tensor, cfg: NVPConfig, temporal_interp: bool = False) -> torch.Tensor:
    """z = [DG_xy(x,y) | DG_yt(t,y) | DG_xt(t,x) | SG(t,x,y)]   (modules.py:61-78; temporal_interp: modules.py:72-73)."""
    tab = cfg.table
    c = coords.reshape(-1, 3)
    xy = dense_grid_forward(p["keyframes_xy.params"], c[:, [1, 2]], cfg.n_features, tab)
    xt = dense_grid_forward(p["keyframes_xt.params"], c[:, [0, 1]], cfg.n_features, tab)
    yt = dense_grid_forward(p["keyframes_yt.params"], c[:, [0, 2]], cfg.n_features, tab)
    sg = (sparse_grid_forward_inter if temporal_interp else sparse_grid_forward)(p["sparse_grid.embeddings"], c)
    return torch.cat((xy, yt, xt, sg), dim=1)


def nvp_forward(p: Dict[str, torch.Tensor], coords: torch.Tensor, tsteps: torch.Tensor, cfg: NVPConfig,
                mma_dtype=None, temporal_interp: bool = False) -> torch.Tensor:
    """coords [N,3]=(t,x,y), tsteps [N] -> rgb [N,3]."""
    z = latent_forward(p, coords, cfg, temporal_interp)
    mods = modulator_forward(z, p, cfg.n_hidden_layers, mma_dtype)
    return siren_forward(tsteps.reshape(-1, 1).to(z.dtype), mods, p, cfg.n_hidden_layers, mma_dtype=mma_dtype)


def normalise_gt(img_u8: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """training.py:47-48."""
    return (img_u8.to(dtype) - 127.5) / 127.5


def image_mse(rgb: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """loss_functions.py:3 with mask=None."""
    return ((rgb - gt) ** 2).mean()


def psnr_from_mse(mse: float) -> float:
    """training.py:58 (peak^2 = 4 for a [-1,1] signal)."""
    return 10.0 * math.log10(4.0 / mse)


PARAM_KEYS_GRID = ("keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params", "sparse_grid.embeddings")


def mlp_param_keys(n_layers: int = 3):
    keys = []
    for i in range(n_layers):
        keys += [f"net.layers.{i}.weight", f"net.layers.{i}.bias"]
    keys += ["net.last_layer.weight", "net.last_layer.bias"]
    for i in range(n_layers):
        keys += [f"wrapper.modulator.layers.{i}.0.weight", f"wrapper.modulator.layers.{i}.0.bias"]
    return keys


def init_params(cfg: NVPConfig, seed: int = 0, grid_std: float = 1e-4, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Reference initialisation under one torch CPU generator.

    grids U(-grid_std, grid_std) (sparsegrid.py:19-21; tcnn default 1e-4); SIREN layer 0 W,b~U(-1,1),
    other SIREN layers W,b~U(+-sqrt(6/dim)) (modulation.py:44-51); modulator W kaiming-normal fan_in/relu,
    b nn.Linear default U(+-1/sqrt(fan_in)) (modulation.py:109-110,151-154).
    """
    g = torch.Generator().manual_seed(seed)
    H, L, Fz = cfg.n_neurons, cfg.n_hidden_layers, cfg.latent_dim
    ncell = cfg.table.n_cells

    def U(shape, a):
        return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * a

    p = {}
    for k in ("keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params"):
        p[k] = U((ncell * cfg.n_features,), grid_std)
    p["sparse_grid.embeddings"] = U((cfg.t_resolution, cfg.x_resolution, cfg.y_resolution, cfg.sparse_features), grid_std)
    for i in range(L):
        din = 1 if i == 0 else H
        a = (1.0 / din) if i == 0 else math.sqrt(6.0 / din)
        p[f"net.layers.{i}.weight"] = U((H, din), a)
        p[f"net.layers.{i}.bias"] = U((H,), a)
    a = math.sqrt(6.0 / H)
    p["net.last_layer.weight"] = U((3, H), a)
    p["net.last_layer.bias"] = U((3,), a)
    for i in range(L):
        din = Fz if i == 0 else H + Fz
        p[f"wrapper.modulator.layers.{i}.0.weight"] = torch.randn((H, din), generator=g) * math.sqrt(2.0 / din)
        p[f"wrapper.modulator.layers.{i}.0.bias"] = U((H,), 1.0 / math.sqrt(din))
    return {k: v.to(dtype) for k, v in p.items()}


def nvp_loss_and_grads(p: Dict[str, torch.Tensor], coords: torch.Tensor, tsteps: torch.Tensor,
                       gt_u8: torch.Tensor, cfg: NVPConfig, n_global: Optional[int] = None,
                       dtype=torch.float32, mma_dtype=None) -> Tuple[torch.Tensor, float, Dict[str, torch.Tensor]]:
    """Forward + MSE + backward of the restated path.  Returns (rgb, loss, grads).

    n_global: denominator of the mean is 3*n_global (for a shard of a larger batch); default N.
    dtype=float64 gives a high-precision reference for tolerance budgeting; mma_dtype=torch.float16
    restates the tensor-core mode's forward operand rounding (see _operand).
    """
    q = {k: v.detach().to(dtype).requires_grad_(True) for k, v in p.items()}
    rgb = nvp_forward(q, coords.to(dtype), tsteps.to(dtype), cfg, mma_dtype)
    gt = normalise_gt(gt_u8.reshape(-1, 3), dtype)
    n = rgb.shape[0] if n_global is None else n_global
    loss = ((rgb - gt) ** 2).sum() / (3.0 * n)
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    return rgb.detach(), float(loss.detach()), grads


def leaky_relu_margin(p: Dict[str, torch.Tensor], coords: torch.Tensor, cfg: NVPConfig, dtype=torch.float64, mma_dtype=None) -> torch.Tensor:
    """min over the modulator's 3x128 units of |pre-activation|, per sample [N].  LeakyReLU's derivative jumps from 0.01
    to 1 at zero (modulation.py:109-117), so a sample with a unit within round-off of zero has no well-defined gradient
    for a comparison across arithmetics; tests use this to keep such samples out of max-norm gradient checks."""
    with torch.no_grad():
        z = latent_forward(p, coords.reshape(-1, 3), cfg).to(dtype)
        q = {k: v.to(dtype) for k, v in p.items() if k not in PARAM_KEYS_GRID}
        hs = modulator_forward(z, q, cfg.n_hidden_layers, mma_dtype)
        margin = None
        for h in hs:
            m = torch.where(h >= 0, h, -h / 0.01).min(dim=1).values     # |m| recovered from h = lrelu(m)
            margin = m if margin is None else torch.minimum(margin, m)
    return margin


def nvp_loss_and_grads_touched(p: Dict[str, torch.Tensor], coords: torch.Tensor, tsteps: torch.Tensor, gt_u8: torch.Tensor,
                              cfg: NVPConfig, n_global: Optional[int] = None, dtype=torch.float64, mma_dtype=None):
    """The same forward + MSE + backward as nvp_loss_and_grads for FULL-SIZE grids (600x300x300xF): the dense layers go
    through autograd, the grid gradients are returned on the touched cells only (autograd of the gathers would build nine
    dense 0.9-1.7 GB tensors: sparsegrid.py:61-69 backward = index_put(accumulate) per neighbour).

    Returns (rgb, loss, mlp_grads, grid_grads, dz) with grid_grads[key] = (flat cell indices int64 [K] sorted unique,
    values [K, F] in `dtype`): values[k] = sum over samples / corners of weight * dz, i.e. tcnn's kernel_grid_backward
    (dparams[index] += w * dout) and the transpose of the 3x3 neighbourhood gather (sparsegrid.py:61-69).
    dz [N, Z] is dL/dz (for size-independent checks: bilinear weights sum to one)."""
    c = coords.reshape(-1, 3)
    with torch.no_grad():
        z0 = latent_forward(p, c, cfg)                  # gather + interpolation in the parameters' own precision
    z = z0.to(dtype).requires_grad_(True)
    q = {k: v.detach().to(dtype).requires_grad_(True) for k, v in p.items() if k not in PARAM_KEYS_GRID}
    mods = modulator_forward(z, q, cfg.n_hidden_layers, mma_dtype)
    rgb = siren_forward(tsteps.reshape(-1, 1).to(dtype), mods, q, cfg.n_hidden_layers, mma_dtype=mma_dtype)
    gt = normalise_gt(gt_u8.reshape(-1, 3), dtype)
    n = rgb.shape[0] if n_global is None else n_global
    loss = ((rgb - gt) ** 2).sum() / (3.0 * n)
    loss.backward()
    dz = z.grad.detach()
    mlp_grads = {k: v.grad for k, v in q.items()}
    cn = c.detach().cpu().numpy()
    tab, F2, F3, L = cfg.table, cfg.n_features, cfg.sparse_features, cfg.n_levels
    pw = L * F2
    grid_grads = {}

    def accumulate(flat_idx: np.ndarray, contrib: np.ndarray):
        uniq, inv = np.unique(flat_idx, return_inverse=True)
        acc = np.zeros((uniq.shape[0], contrib.shape[1]), np.float64)
        np.add.at(acc, inv, contrib)
        return torch.from_numpy(uniq), torch.from_numpy(acc).to(dtype)

    dzn = dz.double().numpy()
    # concat order xy, yt, xt (modules.py:69); plane inputs xy=(x,y), yt=(t,y), xt=(t,x) (modules.py:61-63)
    for k, (key, cols) in enumerate((("keyframes_xy.params", [1, 2]), ("keyframes_yt.params", [0, 2]), ("keyframes_xt.params", [0, 1]))):
        idx, w = dense_grid_indices(cn[:, cols], tab)                         # [N, L, 4]
        d = dzn[:, k * pw:(k + 1) * pw].reshape(-1, L, 1, F2)                  # [N, L, 1, F]
        contrib = (w.astype(np.float64)[..., None] * d).reshape(-1, F2)        # [N*L*4, F]
        grid_grads[key] = accumulate(idx.reshape(-1), contrib)
    ti, xi, yi = sparse_grid_indices(cn, cfg.t_resolution, cfg.x_resolution, cfg.y_resolution)
    X, Y = cfg.x_resolution, cfg.y_resolution
    flat, contrib = [], []
    v = 0
    for i in (-1, 0, 1):
        for j in (-1, 0, 1):
            vx, vy = np.clip(xi + i, 0, X - 1), np.clip(yi + j, 0, Y - 1)
            flat.append((ti * X + vx) * Y + vy)
            contrib.append(dzn[:, 3 * pw + v * F3: 3 * pw + (v + 1) * F3])
            v += 1
    grid_grads["sparse_grid.embeddings"] = accumulate(np.concatenate(flat), np.concatenate(contrib))
    return rgb.detach(), float(loss.detach()), mlp_grads, grid_grads, dz


# --------------------------------------------------------------------------------------------
# Sampler (dataio.py:85-120) and synthetic video
# --------------------------------------------------------------------------------------------
def get_mgrid_2d(h: int, w: int) -> torch.Tensor:
    """dataio.py:11-19: row-major (row/(H-1), col/(W-1))."""
    rows, cols = np.mgrid[:h, :w]
    g = np.stack((rows, cols), axis=-1).astype(np.float32)
    g[..., 0] = g[..., 0] / (h - 1)
    g[..., 1] = g[..., 1] / (w - 1)
    return torch.from_numpy(g.reshape(-1, 2))


def synthetic_video(t: int, h: int, w: int, seed: int = 0) -> np.ndarray:
    """Smooth-plus-noise uint8 [T,H,W,3] stand-in for UVG (not available offline)."""
    rng = np.random.default_rng(seed)
    tt = np.linspace(0, 1, t, dtype=np.float32)[:, None, None]
    yy = np.linspace(0, 1, h, dtype=np.float32)[None, :, None]
    xx = np.linspace(0, 1, w, dtype=np.float32)[None, None, :]
    vid = np.empty((t, h, w, 3), np.uint8)
    for c in range(3):
        base = 0.5 + 0.25 * np.sin(2 * np.pi * ((c + 1) * xx + 0.5 * tt)) * np.cos(2 * np.pi * ((c + 2) * yy - 0.3 * tt))
        noise = rng.normal(0.0, 0.03, size=(t, h, w)).astype(np.float32)
        vid[..., c] = np.clip((base + noise) * 255.0, 0, 255).astype(np.uint8)
    return vid


def sample_batch(vid_flat: torch.Tensor, mgrid: torch.Tensor, n_samples: int, generator: Optional[torch.Generator] = None):
    """dataio.py:104-120.  vid_flat uint8 [T, H*W, 3]; returns coords [N,3], tsteps [N], img u8 [N,3]."""
    T, HW = vid_flat.shape[0], vid_flat.shape[1]
    t_idx = torch.randint(0, T, (n_samples,), generator=generator)
    p_idx = torch.randint(0, HW, (n_samples,), generator=generator)
    img = vid_flat[t_idx, p_idx, :]
    half_dt = 0.5 / T
    temporal_steps = torch.linspace(half_dt, 1 - half_dt, T)[t_idx]
    temporal_coords = torch.linspace(0, 1, T)[t_idx]
    coords = torch.cat((temporal_coords.unsqueeze(1), mgrid[p_idx, :]), dim=1)
    return coords, temporal_steps, img
