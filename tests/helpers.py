"""Shared test helpers: oracle config <-> product model plumbing and golden loading."""
import os

import numpy as np
import torch

from oracle import nvp_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ("s_init", "s_trained", "l_trained")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    F, L, base, F3, T, X, Y = [int(v) for v in g["cfg"]]
    cfg = O.NVPConfig(n_features=F, n_levels=L, base_resolution=base, per_level_scale=float(g["per_level_scale"]),
                      sparse_features=F3, t_resolution=T, x_resolution=X, y_resolution=Y)
    return g, cfg


def golden_params(g, cfg):
    p = O.init_params(cfg, seed=int(g["seed"]), grid_std=float(g["grid_std"]))
    chk = np.asarray([float(v.double().sum()) for v in p.values()] + [float(v.double().abs().sum()) for v in p.values()])
    np.testing.assert_allclose(chk, g["param_checksum"], rtol=1e-12, atol=1e-12,
                               err_msg="torch CPU RNG stream differs from the one that made the golden vectors")
    return p


def make_model(cfg: O.NVPConfig, params=None, mode="fp32", device="cuda"):
    """Product model loaded with the oracle's parameter dict."""
    import nvp_b200
    m = nvp_b200.NVP(type="nvp", out_features=3, encoding_config=cfg.to_json(), mode=mode)
    if params is not None:
        sd = {k: v.clone() for k, v in params.items()}
        sd.update({"wrapper." + k: v.clone() for k, v in params.items() if k.startswith("net.")})
        m.load_state_dict(sd, strict=True)
    return m.to(device)


def model_grads(m):
    out = {}
    for k, v in m.named_parameters():
        out[k] = v.grad.detach().cpu() if v.grad is not None else torch.zeros_like(v).cpu()
    return out


def sampler_like_inputs(cfg, n, seed, H=1080, W=1920):
    g = torch.Generator().manual_seed(seed)
    T = cfg.t_resolution
    t_idx = torch.randint(0, T, (n,), generator=g)
    coords = torch.stack((torch.linspace(0, 1, T)[t_idx],
                          torch.randint(0, H, (n,), generator=g).float() / (H - 1),
                          torch.randint(0, W, (n,), generator=g).float() / (W - 1)), dim=1)
    tsteps = torch.linspace(0.5 / T, 1 - 0.5 / T, T)[t_idx]
    gt = torch.randint(0, 256, (n, 3), generator=g, dtype=torch.uint8)
    return coords, tsteps, gt


def rel_err(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
