"""Pin the oracle against the reference's own Python modules (build container only).

TEST INFRASTRUCTURE.  Imports /root/reference/{sparsegrid,modulation,loss_functions,modules,dataio}.py
UNMODIFIED (read-only, via sys.path) with sys.modules stubs for packages that are absent here:
  skvideo / skvideo.io / pytorch_msssim : unused on the hot path
  tinycudann                            : stub whose Encoding is oracle.dense_grid_forward (the
                                          real fork is CUDA-only, un-vendored and un-pinned)
and checks oracle/nvp_oracle.py against them: forward outputs, loss, all gradients, the sampler.

/root/reference does not exist on the GPU box; this script is run here (it is also invoked by
tests/test_oracle_vs_reference.py when the reference tree is present, skipped otherwise).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import nvp_oracle as O  # noqa: E402

REF = os.environ.get("NVP_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF, "modules.py"))


def import_reference():
    """Returns the reference modules namespace (modules, sparsegrid, modulation, loss_functions, dataio)."""
    if "tinycudann" not in sys.modules:
        tcnn = types.ModuleType("tinycudann")

        class Encoding(torch.nn.Module):
            """tcnn.Encoding stand-in: same ctor/call/.params/.dtype surface (modules.py:14-23,65-67)."""

            def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=None):
                super().__init__()
                assert n_input_dims == 2 and encoding_config["otype"] == "DenseGrid"
                self.F = encoding_config["n_features_per_level"]
                self.table = O.level_table(encoding_config["n_levels"], encoding_config["base_resolution"],
                                           encoding_config["per_level_scale"])
                g = torch.Generator().manual_seed(seed)
                self.params = torch.nn.Parameter((torch.rand(self.table.n_cells * self.F, generator=g) * 2 - 1) * 1e-4)
                self.dtype = torch.float32

            def forward(self, x):
                return O.dense_grid_forward(self.params, x, self.F, self.table)

        tcnn.Encoding = Encoding
        sys.modules["tinycudann"] = tcnn
    for name in ("skvideo", "skvideo.io", "skvideo.datasets", "pytorch_msssim"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "pytorch_msssim":
                m.ms_ssim = lambda *a, **k: None
            sys.modules[name] = m
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import dataio, loss_functions, modulation, modules, sparsegrid  # noqa: E401
    return types.SimpleNamespace(modules=modules, sparsegrid=sparsegrid, modulation=modulation,
                                 loss_functions=loss_functions, dataio=dataio)


def small_cfg(F=2, T=6, X=20, Y=24, levels=16):
    return O.NVPConfig(n_features=F, n_levels=levels, sparse_features=F, t_resolution=T, x_resolution=X, y_resolution=Y)


def edge_coords(n, T, H, W, seed):
    """Coordinates drawn like the sampler, with the edge cases (0, 1, exact pixel centres) included."""
    g = torch.Generator().manual_seed(seed)
    t = torch.linspace(0, 1, T)[torch.randint(0, T, (n,), generator=g)]
    x = torch.randint(0, H, (n,), generator=g).float() / (H - 1)
    y = torch.randint(0, W, (n,), generator=g).float() / (W - 1)
    c = torch.stack((t, x, y), dim=1)
    c[:8] = torch.tensor([[0, 0, 0], [1, 1, 1], [1, 0, 1], [0, 1, 0], [1, 1, 0], [0, 0, 1], [0.5, 0.5, 0.5], [1, 0.5, 1]],
                         dtype=torch.float32)[: min(8, n)]
    return c


def build_reference_model(ref, cfg: O.NVPConfig, p):
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref.modules.NVP(type="nvp", out_features=3, encoding_config=cfg.to_json())
    sd = m.state_dict()
    for k, v in p.items():
        assert k in sd and sd[k].shape == v.shape, (k, sd.get(k, None) is not None and sd[k].shape, v.shape)
    missing = m.load_state_dict({**p, **{"wrapper." + k: v for k, v in p.items() if k.startswith("net.")}}, strict=True)
    return m


def check(cfg: O.NVPConfig, n=512, seed=0, grid_std=0.5, verbose=True):
    ref = import_reference()
    p = O.init_params(cfg, seed=seed, grid_std=grid_std)
    coords = edge_coords(n, cfg.t_resolution, 37, 53, seed + 1)
    tsteps = (coords[:, 0] * (cfg.t_resolution - 1) + 0.5) / cfg.t_resolution
    g = torch.Generator().manual_seed(seed + 2)
    gt_u8 = torch.randint(0, 256, (n, 3), generator=g, dtype=torch.uint8)

    model = build_reference_model(ref, cfg, p)
    out = model({"all_coords": coords[None], "temporal_steps": tsteps[None]})["model_out"]
    gt = {"img": ((gt_u8.float() - 127.5) / 127.5)[None]}
    loss = ref.loss_functions.image_mse(None, {"model_out": out}, gt)["img_loss"]
    loss.backward()
    ref_grads = {k: v.grad for k, v in model.named_parameters()}

    rgb, oloss, grads = O.nvp_loss_and_grads(p, coords, tsteps, gt_u8, cfg)
    res = {"fwd_max_abs": float((rgb - out[0].detach()).abs().max()), "loss_abs": abs(oloss - float(loss))}
    for k, gk in grads.items():
        rk = ref_grads[k]
        denom = float(rk.abs().max()) + 1e-30
        res["grad:" + k] = float((gk - rk).abs().max()) / denom
    # sparse grid and the sampler directly
    sg = ref.sparsegrid.SparseGrid(cfg.sparse_features, cfg.x_resolution, cfg.y_resolution, cfg.t_resolution, False)
    sg.embeddings.data.copy_(p["sparse_grid.embeddings"])
    res["sparse_max_abs"] = float((sg(coords) - O.sparse_grid_forward(p["sparse_grid.embeddings"], coords)).abs().max())
    # eval-time temporal interpolation (sparsegrid.py:76-156), fractional frame positions; NaN at t == 1 in both
    ci = coords.clone()
    ci[:, 0] = torch.rand(n, generator=g) * 0.999
    ci[:4, 0] = torch.tensor([0.0, 1.0, 0.5, 1.0 / (cfg.t_resolution - 1)])
    a, b = sg.forward_inter(ci).detach(), O.sparse_grid_forward_inter(p["sparse_grid.embeddings"], ci)
    assert torch.equal(torch.isnan(a), torch.isnan(b))
    res["sparse_inter_max_abs"] = float((torch.nan_to_num(a) - torch.nan_to_num(b)).abs().max())
    with torch.no_grad():
        oi = model({"all_coords": ci[None], "temporal_steps": tsteps[None]}, temporal_interp=True)["model_out"][0]
    mine = O.nvp_forward(p, ci, tsteps, cfg, temporal_interp=True)
    res["fwd_inter_max_abs"] = float((torch.nan_to_num(oi) - torch.nan_to_num(mine)).abs().max())
    if verbose:
        for k, v in res.items():
            print(f"  {k:55s} {v:.3e}")
    return res


def check_sampler(T=5, H=12, W=16, n=1000, seed=3):
    ref = import_reference()
    vid = O.synthetic_video(T, H, W, seed=1)

    class DS:
        nframes, channels, shape = T, 3, (H, W)

        def __len__(self):
            return 1

        def __getitem__(self, i):
            return vid

    w = ref.dataio.VideoTimeWrapper(DS(), sidelength=(H, W))
    w.N_samples = n
    torch.manual_seed(seed)
    a, b = w[0]
    torch.manual_seed(seed)
    c, ts, img = O.sample_batch(torch.from_numpy(vid).view(T, -1, 3), O.get_mgrid_2d(H, W), n)
    ok = torch.equal(a["all_coords"], c) and torch.equal(a["temporal_steps"], ts) and torch.equal(b["img"], img)
    return ok


def main():
    assert reference_available(), f"reference tree not found at {REF}"
    worst = 0.0
    for F in (2, 4):
        for std in (1e-4, 0.5):
            print(f"config F={F} grid_std={std}")
            r = check(small_cfg(F=F), n=512, seed=F, grid_std=std)
            worst = max(worst, *r.values())
    print("sampler identical:", check_sampler())
    print(f"worst deviation {worst:.3e}")
    assert worst < 2e-5 and check_sampler()
    print("ORACLE PINNED against the reference modules (DenseGrid arithmetic: layout only, see header).")


if __name__ == "__main__":
    main()
