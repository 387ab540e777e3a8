"""CPU tests of the oracle: golden vectors, the reference modules (when present), domain properties."""
import numpy as np
import pytest
import torch

from oracle import check_vs_reference as R
from oracle import nvp_oracle as O
from tests.helpers import GOLDEN_CASES, golden_params, load_golden


def test_level_table_matches_reference_layout():
    # eval.py:28-35 / compression.py:26-33: res = ceil(exp(i*log(1.35))*16-1)+1, cumulative res^2, no padding
    t = O.level_table()
    assert list(t.res) == [16, 22, 30, 40, 54, 72, 97, 131, 177, 239, 322, 435, 587, 792, 1069, 1443]
    assert t.n_cells == 4616112  # SURVEY A.2; README bpp 0.901 depends on it
    import math
    for i in range(16):
        a = math.exp(i * math.log(1.35)) * 16 - 1
        assert int(math.ceil(a) + 1) == t.res[i]


def test_parameter_counts():
    for F, total in ((2, 135807267), (4, 271547715)):
        cfg = O.NVPConfig(n_features=F, sparse_features=F)
        n_mlp = (128 * 1 + 128) + 2 * (128 * 128 + 128) + (3 * 128 + 3)
        z = cfg.latent_dim
        n_mlp += (128 * z + 128) + 2 * (128 * (128 + z) + 128)
        n = 3 * cfg.table.n_cells * F + 600 * 300 * 300 * F + n_mlp
        assert n == total


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_reproduces_golden(name):
    g, cfg = load_golden(name)
    p = golden_params(g, cfg)
    coords, tsteps, gt = torch.from_numpy(g["coords"]), torch.from_numpy(g["tsteps"]), torch.from_numpy(g["gt"])
    rgb, loss, grads = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg)
    np.testing.assert_allclose(O.latent_forward(p, coords, cfg).numpy(), g["z"], rtol=0, atol=1e-9)
    np.testing.assert_allclose(rgb.numpy(), g["rgb"], rtol=0, atol=2e-7)
    np.testing.assert_allclose(rgb.numpy(), g["rgb64"], rtol=0, atol=5e-7)
    assert abs(loss - float(g["loss64"])) < 1e-6
    for k, v in grads.items():
        v = v.reshape(-1)
        if k in O.PARAM_KEYS_GRID:
            ref = torch.zeros_like(v)
            ref[torch.from_numpy(g["gidx:" + k])] = torch.from_numpy(g["gval:" + k])
        else:
            ref = torch.from_numpy(g["grad:" + k])
        assert float((v - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 1e-12, k


@pytest.mark.skipif(not R.reference_available(), reason="reference tree only exists in the build container")
def test_oracle_pinned_against_reference_modules():
    for F in (2, 4):
        res = R.check(R.small_cfg(F=F), n=256, seed=5 + F, grid_std=0.3, verbose=False)
        assert max(res.values()) < 2e-5, res
    assert R.check_sampler()


def test_sparse_grid_edges_and_neighbourhood_order():
    T, X, Y, F = 3, 4, 5, 2
    emb = torch.arange(T * X * Y * F, dtype=torch.float32).reshape(T, X, Y, F)
    c = torch.tensor([[0.0, 0.0, 0.0], [1.0, 1.0, 1.0], [0.5, 0.5, 0.5]])
    out = O.sparse_grid_forward(emb, c)
    assert out.shape == (3, 9 * F)
    # corner (0,0,0): neighbours clamp to 0 -> first four 3x3 entries touching x=-1 or y=-1 repeat voxel (0,0)/(0,1)...
    assert torch.equal(out[0, 0:2], emb[0, 0, 0]) and torch.equal(out[0, 8:10], emb[0, 0, 0])  # (i,j)=(-1,-1),(0,0)
    assert torch.equal(out[0, 16:18], emb[0, 1, 1])  # (i,j)=(1,1)
    assert torch.equal(out[1, 16:18], emb[T - 1, X - 1, Y - 1])
    # centre: t=0.5*(T-1)+0.5 -> 1 ; x=0.5*3+0.5=2 ; y=0.5*4+0.5=2 (trunc)
    assert torch.equal(out[2, 8:10], emb[1, 2, 2])


def test_dense_grid_edge_aliasing_and_partition_of_unity():
    tab = O.level_table()
    u = torch.tensor([[1.0, 1.0], [0.0, 0.0], [1.0, 0.3], [0.123, 1.0]])
    idx, w = O.dense_grid_indices(u.numpy(), tab)
    assert np.allclose(w.sum(axis=2), 1.0, atol=1e-6)
    for l in range(tab.n_levels):
        lo, hi = tab.offsets[l], tab.offsets[l + 1]
        assert (idx[:, l] >= lo).all() and (idx[:, l] < hi).all()
    # level 0 (scale 15.0, res 16) at u=1: cell 15 + 0.5 -> corner +1 = 16 aliases to the next row (A.2)
    r = 16
    assert idx[0, 0, 1] == (16 + 15 * r) % (r * r)
    assert idx[0, 0, 3] == (16 + 16 * r) % (r * r)


def test_forward_is_linear_in_last_layer_and_loss_scales_with_n_global():
    cfg = O.NVPConfig(t_resolution=4, x_resolution=6, y_resolution=6, n_levels=4)
    p = O.init_params(cfg, seed=3, grid_std=0.2)
    g = torch.Generator().manual_seed(1)
    coords, tau = torch.rand(64, 3, generator=g), torch.rand(64, generator=g)
    gt = torch.randint(0, 256, (64, 3), generator=g, dtype=torch.uint8)
    a = O.nvp_forward(p, coords, tau, cfg)
    q = dict(p)
    q["net.last_layer.weight"] = 2 * p["net.last_layer.weight"]
    q["net.last_layer.bias"] = 2 * p["net.last_layer.bias"]
    assert torch.allclose(O.nvp_forward(q, coords, tau, cfg), 2 * a, atol=1e-6)
    _, l1, g1 = O.nvp_loss_and_grads(p, coords, tau, gt, cfg)
    _, l2, g2 = O.nvp_loss_and_grads(p, coords, tau, gt, cfg, n_global=128)
    assert abs(l1 - 2 * l2) < 1e-7
    assert torch.allclose(g1["net.last_layer.weight"], 2 * g2["net.last_layer.weight"], atol=1e-9)


def test_touched_cells_gradients_equal_dense_autograd():
    """nvp_loss_and_grads_touched (used by the full-size GPU parity tests) == nvp_loss_and_grads on a grid small enough for
    dense autograd: same loss, same 14 dense-layer gradients, same grid gradients on the touched cells, zero elsewhere."""
    from tests.helpers import sampler_like_inputs
    cfg = O.NVPConfig(t_resolution=7, x_resolution=33, y_resolution=29)
    p = {k: v.double() for k, v in O.init_params(cfg, seed=3, grid_std=0.3).items()}
    coords, tsteps, gt = sampler_like_inputs(cfg, 2000, seed=1)
    rgb, loss, grads = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg, dtype=torch.float64)
    rgb2, loss2, mg, gg, dz = O.nvp_loss_and_grads_touched(p, coords, tsteps, gt, cfg)
    assert abs(loss - loss2) < 1e-15 and float((rgb - rgb2).abs().max()) < 1e-14
    for k, v in mg.items():
        assert float((v - grads[k]).abs().max()) <= 1e-14 * float(grads[k].abs().max()), k
    for k, (idx, val) in gg.items():
        F = cfg.n_features if "keyframes" in k else cfg.sparse_features
        dense = grads[k].reshape(-1, F)
        assert float((dense[idx] - val).abs().max()) <= 1e-13 * float(dense.abs().max()), k
        rest = torch.ones(dense.shape[0], dtype=torch.bool)
        rest[idx] = False
        assert float(dense[rest].abs().sum()) == 0.0, k
    # bilinear weights sum to one: a plane's gradient mass per feature equals the sum of its dz columns
    pw = cfg.n_levels * cfg.n_features
    for k, key in enumerate(("keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params")):
        mass = gg[key][1].sum(dim=0)
        want = dz[:, k * pw:(k + 1) * pw].reshape(-1, cfg.n_levels, cfg.n_features).sum(dim=(0, 1))
        assert float((mass - want).abs().max()) <= 1e-8 * float(dz.abs().sum())   # the fp32 weights sum to one to 1e-7
