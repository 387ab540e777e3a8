"""Generate the golden vectors under tests/golden/ from the oracle (run in the build container).

TEST INFRASTRUCTURE.  The oracle is first pinned against the reference's own modules
(oracle/check_vs_reference.py, which must pass here), then evaluated on seeded inputs; inputs and
expected outputs are stored compactly (parameters are regenerated from the seed by
oracle.init_params, with a checksum stored to detect RNG drift; grid gradients are stored sparsely
at the touched cells).  Usage:  python oracle/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import nvp_oracle as O  # noqa: E402
from oracle import check_vs_reference as R  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (F, T, X, Y, n, seed, grid_std)
    "s_init": (2, 6, 20, 24, 192, 11, 1e-4),
    "s_trained": (2, 6, 20, 24, 192, 12, 0.5),
    "l_trained": (4, 5, 18, 16, 160, 13, 0.5),
}


def param_checksum(p):
    return np.asarray([float(v.double().sum()) for v in p.values()] + [float(v.double().abs().sum()) for v in p.values()])


def make_case(name):
    F, T, X, Y, n, seed, std = CASES[name]
    cfg = O.NVPConfig(n_features=F, sparse_features=F, t_resolution=T, x_resolution=X, y_resolution=Y)
    p = O.init_params(cfg, seed=seed, grid_std=std)
    coords = R.edge_coords(n, T, 1080, 1920, seed + 100)
    tsteps = (torch.round(coords[:, 0] * (T - 1)) + 0.5) / T
    g = torch.Generator().manual_seed(seed + 200)
    gt = torch.randint(0, 256, (n, 3), generator=g, dtype=torch.uint8)
    rgb, loss, grads = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg)
    rgb64, loss64, grads64 = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg, dtype=torch.float64)
    z = O.latent_forward(p, coords, cfg)
    out = {
        "cfg": np.asarray([F, cfg.n_levels, cfg.base_resolution, F, T, X, Y], np.int64),
        "per_level_scale": np.float32(cfg.per_level_scale), "seed": np.int64(seed), "grid_std": np.float64(std),
        "param_checksum": param_checksum(p),
        "coords": coords.numpy(), "tsteps": tsteps.numpy(), "gt": gt.numpy(),
        "z": z.numpy(), "rgb": rgb.numpy(), "rgb64": rgb64.numpy().astype(np.float32), "loss": np.float64(loss), "loss64": np.float64(loss64),
    }
    for k, v in grads64.items():
        v = v.reshape(-1)
        if k in O.PARAM_KEYS_GRID:
            nz = torch.nonzero(v).reshape(-1)
            out["gidx:" + k] = nz.numpy().astype(np.int64)
            out["gval:" + k] = v[nz].numpy().astype(np.float32)
        else:
            out["grad:" + k] = v.numpy().astype(np.float32)
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss, "fp32-vs-fp64 rgb", float((rgb.double() - rgb64).abs().max()))


if __name__ == "__main__":
    if R.reference_available():
        R.main()
    else:
        print("WARNING: reference tree absent; golden vectors come from an unpinned oracle")
    for c in CASES:
        make_case(c)
