"""GPU tests (-m gpu) of the tile-binned keyframe gather / scatter-add (csrc/grid_binned.cuh) behind the tensor-core
path: edge aliasing in the backward, coordinates outside [0,1] (direct-access fallback inside the binned kernels),
crowded buckets that split into several tasks, and agreement with the direct (unbinned) kernels, which stay
selectable with NVP_GRID_BINNED=0.
"""
import os

import pytest
import torch

from oracle import nvp_oracle as O
from tests.helpers import make_model, model_grads, rel_err, sampler_like_inputs
from tests.test_gpu_parity import FWD_TOL, assert_grads

pytestmark = pytest.mark.gpu


def run_step(cfg, p, coords, tsteps, gt, binned=True, mode="tc"):
    old = os.environ.get("NVP_GRID_BINNED")
    os.environ["NVP_GRID_BINNED"] = "1" if binned else "0"
    try:
        m = make_model(cfg, p, mode=mode)
        n = coords.shape[0]
        rgb = torch.empty(n, 3, device="cuda")
        ls = m.fwd_loss_bwd({"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}, gt.cuda(), out_rgb=rgb)
        torch.cuda.synchronize()
        return rgb.cpu(), float(ls) / (3 * n), model_grads(m)
    finally:
        if old is None:
            os.environ.pop("NVP_GRID_BINNED", None)
        else:
            os.environ["NVP_GRID_BINNED"] = old


def oracle_refs(cfg, p, coords, tsteps, gt):
    rgb, loss, grads = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg, dtype=torch.float64)
    grads16 = O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg, dtype=torch.float64, mma_dtype=torch.float16)[2]
    return rgb, loss, grads, grads16


def test_edge_coordinates_backward_matches_oracle():
    """u == 1.0 (flat-index aliasing into the next row / wrap to cell 0), u == 0 and tile-boundary values:
    the window keeps the aliased 'virtual' cells and must flush them onto their aliases."""
    cfg = O.NVPConfig(t_resolution=5, x_resolution=9, y_resolution=11)
    p = O.init_params(cfg, seed=23, grid_std=0.5)
    vals = torch.tensor([0.0, 1.0, 0.5, 0.25, 63.0 / 64, 1.0 / 1079, 1078.0 / 1079, 1.0 / 128])
    coords = torch.cartesian_prod(vals, vals, vals)
    n = coords.shape[0]
    g = torch.Generator().manual_seed(3)
    tsteps = torch.rand(n, generator=g)
    gt = torch.randint(0, 256, (n, 3), generator=g, dtype=torch.uint8)
    rgb_ref, loss_ref, grads, grads16 = oracle_refs(cfg, p, coords, tsteps, gt)
    rgb, loss, got = run_step(cfg, p, coords, tsteps, gt)
    assert float((rgb.double() - rgb_ref).abs().max()) <= FWD_TOL["tc"]
    assert abs(loss - loss_ref) <= 2e-3
    assert_grads(got, grads, "tc", grads16, "edges")


@pytest.mark.parametrize("n", [900, 6000])
def test_crowded_bucket_splits_into_tasks(n):
    """All samples in one frame and a narrow band of x: a handful of tiles hold everything, so buckets exceed the
    task chunk and several warps accumulate the same window (their flushes must add up)."""
    cfg = O.NVPConfig(t_resolution=6, x_resolution=20, y_resolution=24)
    p = O.init_params(cfg, seed=11, grid_std=0.3)
    g = torch.Generator().manual_seed(n)
    coords = torch.rand(n, 3, generator=g)
    coords[:, 0] = 0.4
    coords[:, 1] = 0.30 + 0.01 * coords[:, 1]
    tsteps = torch.full((n,), 0.41)
    gt = torch.randint(0, 256, (n, 3), generator=g, dtype=torch.uint8)
    rgb_ref, loss_ref, grads, grads16 = oracle_refs(cfg, p, coords, tsteps, gt)
    rgb, loss, got = run_step(cfg, p, coords, tsteps, gt)
    assert float((rgb.double() - rgb_ref).abs().max()) <= FWD_TOL["tc"]
    assert_grads(got, grads, "tc", grads16, f"crowded n={n}")


def test_binned_agrees_with_direct_kernels_including_out_of_range_coordinates():
    """Same call through both grid implementations.  Coordinates outside [0,1] land in the border tiles and take the
    direct-access branch inside the binned kernels (negative / beyond-the-table flat indices wrap modulo the level)."""
    cfg = O.NVPConfig(t_resolution=8, x_resolution=30, y_resolution=26)
    p = O.init_params(cfg, seed=77, grid_std=0.4)
    n = 20000
    coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=5)
    g = torch.Generator().manual_seed(6)
    k = 400
    coords[:k] = torch.rand(k, 3, generator=g) * 1.6 - 0.3      # in [-0.3, 1.3]
    coords[k:2 * k, 0] = 1.0
    coords[2 * k:3 * k, 2] = 1.0
    rgb_b, loss_b, g_b = run_step(cfg, p, coords, tsteps, gt, binned=True)
    rgb_d, loss_d, g_d = run_step(cfg, p, coords, tsteps, gt, binned=False)
    assert torch.isfinite(rgb_b).all()
    # the two gathers round their 4-corner sums in a different order: the latent differs by fp32 round-off, which the
    # fp16 operand conversion can amplify to one fp16 ulp
    assert float((rgb_b - rgb_d).abs().max()) <= 1e-3
    assert abs(loss_b - loss_d) <= 1e-4 * abs(loss_d)
    for name in g_d:
        assert rel_err(g_b[name], g_d[name]) <= 5e-3, name


@pytest.mark.parametrize("t_res,n", [(600, 1 << 17), (300, 1 << 17), (600, O.N_SAMPLES_PER_STEP)])
def test_full_size_tables_binned_vs_direct(t_res, n):
    """Config S tables (16 levels up to 1443^2 cells, T x 300 x 300 voxels; T = 600 Jockey / 300 ShakeNDry, BASELINE
    configs[1..2]) at a reduced batch and at the full 1,245,184-sample batch: binned and direct paths agree, and the
    per-plane gradient mass equals the direct kernels' (nothing lost at window borders)."""
    cfg = O.NVPConfig(t_resolution=t_res)
    torch.manual_seed(0)
    m = make_model(cfg, None, mode="tc")
    p = {k: v.detach().cpu() for k, v in m.state_dict().items() if not k.startswith("wrapper.net.")}
    del m
    coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=8)
    rgb_b, loss_b, g_b = run_step(cfg, p, coords, tsteps, gt, binned=True)
    rgb_d, loss_d, g_d = run_step(cfg, p, coords, tsteps, gt, binned=False)
    assert float((rgb_b - rgb_d).abs().max()) <= 1e-3
    for name in ("keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params", "sparse_grid.embeddings"):
        assert rel_err(g_b[name], g_d[name]) <= 5e-3, name
        assert abs(float(g_b[name].double().sum()) - float(g_d[name].double().sum())) <= 1e-3 * float(g_d[name].double().abs().sum())


def test_grid_grads_event_is_recorded_between_scatter_and_wgrad():
    """nvp_record_grid_grads_event: the event handed to the fused step completes, results are unchanged, and a side stream
    that waits on it sees the final grid gradients (what the multi-GPU host all-reduces under the wgrad kernel)."""
    cfg = O.NVPConfig(t_resolution=6, x_resolution=20, y_resolution=24)
    p = O.init_params(cfg, seed=3, grid_std=0.3)
    n = 5000
    coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=12)
    _, _, ref = run_step(cfg, p, coords, tsteps, gt)
    m = make_model(cfg, p, mode="tc")
    ev = torch.cuda.Event()
    side = torch.cuda.Stream()
    m.fwd_loss_bwd({"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}, gt.cuda(), grid_event=ev)
    side.wait_event(ev)
    with torch.cuda.stream(side):
        snap = {k: v.grad.detach().clone() for k, v in m.named_parameters() if "keyframes" in k or "sparse_grid" in k}
    torch.cuda.synchronize()
    assert ev.query()
    got = model_grads(m)
    for k in ref:
        assert rel_err(got[k], ref[k]) <= 1e-4, k
    for k, v in snap.items():
        assert rel_err(v.cpu(), ref[k]) <= 1e-4, k


@pytest.mark.parametrize("levels,ilv", [(16, "0"), (12, "1"), (24, "1"), (8, "1")])
def test_scatter_layouts_and_level_counts_match_oracle(levels, ilv):
    """The scatter-add's two window layouts and its lane mapping for level counts other than 16: the packed layout forced
    on the default shape (NVP_BIN_ILV=0), idle lanes (12 and 8 levels: 24 / 16 of the 32 lanes have a level) and two
    passes of 16 levels (24 levels; latent 162 wide -> the unfused forward / backward kernels).  Includes coordinates
    outside [0,1] (revisited by the direct path after the branch-free batch loop)."""
    cfg = O.NVPConfig(n_levels=levels, t_resolution=7, x_resolution=21, y_resolution=18)
    p = O.init_params(cfg, seed=100 + levels, grid_std=0.4)
    n = 5000
    coords, tsteps, gt = sampler_like_inputs(cfg, n, seed=levels)
    g = torch.Generator().manual_seed(levels)
    coords[:200] = torch.rand(200, 3, generator=g) * 1.4 - 0.2
    rgb_ref, loss_ref, grads, _ = oracle_refs(cfg, p, coords, tsteps, gt)
    old = os.environ.get("NVP_BIN_ILV")
    os.environ["NVP_BIN_ILV"] = ilv
    try:
        rgb, loss, got = run_step(cfg, p, coords, tsteps, gt)
        _, _, direct = run_step(cfg, p, coords, tsteps, gt, binned=False)
    finally:
        if old is None:
            os.environ.pop("NVP_BIN_ILV", None)
        else:
            os.environ["NVP_BIN_ILV"] = old
    assert float((rgb.double() - rgb_ref).abs().max()) <= FWD_TOL["tc"]
    assert abs(loss - loss_ref) <= 2e-3
    from tests.test_gpu_parity import GRAD_L2_TOL_EXACT, l2_rel
    for name, ref in grads.items():
        # the index arithmetic: same MLP arithmetic on both sides, so the grid gradients agree to accumulation-order noise
        assert rel_err(got[name], direct[name]) <= 5e-3, (name, levels, ilv, rel_err(got[name], direct[name]))
        # the arithmetic mode against the exact oracle (a max-norm bar is meaningless for cells touched by a single sample
        # whose LeakyReLU unit flips under fp16 rounding)
        assert l2_rel(got[name], ref) <= GRAD_L2_TOL_EXACT, (name, levels, ilv, l2_rel(got[name], ref))
