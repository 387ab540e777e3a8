"""GPU parity at the BENCHMARKED grid sizes (VERDICT r1 item 2 / ADVICE r1): the fused step at full config S and L
(600x300x300 sparse grid, 16-level 1443^2 keyframe planes; also the 300-frame variant of README.md:55) against the
fp64 oracle -- all 18 gradients, the grid gradients compared on the touched cells and required to be zero elsewhere --
plus the size-independent mass-conservation property of the scatter-add.

The batch is the sampler's, minus the samples that have a modulator unit within 2e-5 of LeakyReLU's kink in either
arithmetic (a few per cent): there the derivative jumps 100x on a round-off difference, and at these sizes most voxels and
fine-level cells carry the gradient of ONE sample, so a max-norm comparison across arithmetics is only meaningful away
from the kink (the L2 comparison over the full batch lives in test_gpu_parity.py).  Everything else -- index arithmetic,
bucketing, windows, aliasing, collisions, the flush -- is compared cell by cell.

Tolerances (relative to the largest reference entry of each tensor; measured values are printed):
  fp32 mode: 1e-4 for the dense layers, 2e-5 for the grids (measured 1.4e-6 / 7e-7)
  tc mode:   against the oracle restated with fp16 GEMM operands: 1e-3 for the dense layers (measured 2.5e-4); grids: L2 3e-3
             and at most 1e-4 of the touched entries off by more than 1e-2 of the largest (measured 9e-4 / 1.2e-5: the
             residual kink flips, see below); 2e-2 L2-relative against the exact oracle (measured 1.3e-2)
"""
import pytest
import torch

from oracle import nvp_oracle as O
from tests.helpers import make_model, sampler_like_inputs

pytestmark = pytest.mark.gpu

N = 1 << 17
GRAD_TOL = {"fp32": 1e-4, "tc": 1e-3}      # dense layers; measured 1.4e-6 / 2.5e-4 (r2, profiles/r02_parity_fullsize.txt)
GRID_TOL = {"fp32": 2e-5, "tc": 1e-2}      # grids, max-norm; measured 7e-7 / 2.6e-3 (keyframes)
KINK_MARGIN = 2e-5
L2_TOL_EXACT = 2e-2                          # tc vs the exact oracle, L2; measured 1.3e-2
CASES = {
    "s_600": dict(n_features=2, sparse_features=2, t_resolution=600),
    "s_300": dict(n_features=2, sparse_features=2, t_resolution=300),
    "l_600": dict(n_features=4, sparse_features=4, t_resolution=600),
}


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def l2(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module", params=list(CASES))
def case(request):
    cfg = O.NVPConfig(**CASES[request.param])
    p = O.init_params(cfg, seed=11, grid_std=0.25)      # trained-scale features: errors are not hidden by 1e-4 grids
    coords, tsteps, gt = sampler_like_inputs(cfg, N, seed=12)
    keep = (O.leaky_relu_margin(p, coords, cfg) >= KINK_MARGIN) & (O.leaky_relu_margin(p, coords, cfg, mma_dtype=torch.float16) >= KINK_MARGIN)
    assert float(keep.float().mean()) > 0.9
    coords, tsteps, gt = coords[keep].contiguous(), tsteps[keep].contiguous(), gt[keep].contiguous()
    exact = O.nvp_loss_and_grads_touched(p, coords, tsteps, gt, cfg)
    f16 = O.nvp_loss_and_grads_touched(p, coords, tsteps, gt, cfg, mma_dtype=torch.float16)
    return request.param, cfg, p, coords, tsteps, gt, exact, f16


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_fused_step_full_size_all_gradients_vs_oracle(case, mode):
    name, cfg, p, coords, tsteps, gt, exact, f16 = case
    rgb_ref, loss_ref, mlp_ref, grid_ref, dz_ref = exact
    N = coords.shape[0]
    m = make_model(cfg, p, mode=mode)
    x = {"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}
    rgb = torch.empty(N, 3, device="cuda")
    ls = m.fwd_loss_bwd(x, gt.cuda(), out_rgb=rgb)
    torch.cuda.synchronize()
    ferr = float((rgb.cpu().double() - rgb_ref).abs().max())
    assert ferr <= (2e-5 if mode == "fp32" else 1e-3), ferr
    assert abs(float(ls) / (3 * N) - loss_ref) <= (1e-5 if mode == "fp32" else 2e-3)
    named = dict(m.named_parameters())
    worst = {"mlp": 0.0, "grid": 0.0, "l2": 0.0}
    ref_cmp = exact if mode == "fp32" else f16
    for k, ref in ref_cmp[2].items():
        got = named[k].grad.detach().cpu()
        e = rel(got, ref)
        worst["mlp"] = max(worst["mlp"], e)
        assert e <= GRAD_TOL[mode], (name, mode, k, e)
        if mode == "tc":
            e2 = l2(got, mlp_ref[k])
            worst["l2"] = max(worst["l2"], e2)
            assert e2 <= L2_TOL_EXACT, (name, mode, k, "l2 vs exact", e2)
    for k, (idx, val) in ref_cmp[3].items():
        F = cfg.n_features if "keyframes" in k else cfg.sparse_features
        g = named[k].grad.detach().reshape(-1, F)
        got = g[idx.cuda()].cpu()
        e = rel(got, val)
        worst["grid"] = max(worst["grid"], e)
        if mode == "tc":
            # In tc mode the kink cannot be kept out completely: an fp16 operand that rounds the other way (h, a = sin h
            # land within 1e-6 of a rounding boundary for ~1 in 500 elements) moves the next layer's pre-activations by
            # ~1e-5, so a few samples still flip a unit, and on the 3-D grid / fine levels a cell is one sample.  The
            # criterion there: L2 over all touched cells, and the share of cells off by more than the max-norm bar.
            d = (got.double() - val.double()).abs().reshape(-1) / float(val.abs().max())
            bad = float((d > GRID_TOL[mode]).double().mean())
            e_l2 = l2(got, val)
            print(f"\n[fullsize {name} tc] {k}: max-norm {e:.2e}, L2 {e_l2:.2e}, share of entries off by > {GRID_TOL[mode]:.0e}: {bad:.2e}, "
                  f"99.99th percentile {float(torch.quantile(d[:: max(1, d.numel() // 4000000)], 0.9999)):.2e}")
            assert e_l2 <= 3e-3 and bad <= 1e-4, (name, mode, k, e, e_l2, bad)   # measured 9e-4 / 1.2e-5
        else:
            assert e <= GRID_TOL[mode], (name, mode, k, e)
        # nothing outside the touched cells: total |grad| mass == mass on the touched cells
        total, touched = float(g.abs().sum(dtype=torch.float64)), float(g[idx.cuda()].abs().sum(dtype=torch.float64))
        assert abs(total - touched) <= 1e-6 * total, (name, mode, k, "gradient outside the touched cells")
        if mode == "tc":
            e2 = l2(got, exact[3][k][1])
            worst["l2"] = max(worst["l2"], e2)
            assert e2 <= L2_TOL_EXACT, (name, mode, k, "l2 vs exact", e2)
    print(f"\n[fullsize {name} {mode}] n = {N}; forward max|err| {ferr:.2e}; worst grad rel err: dense layers {worst['mlp']:.2e}, grids {worst['grid']:.2e}"
          + (f"; worst L2 vs exact oracle {worst['l2']:.2e}" if mode == "tc" else ""))


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_scatter_mass_conservation_full_size(case, mode):
    """Bilinear weights sum to one per (sample, level) and the 3x3 neighbourhood has unit weights, so per feature the
    gradient mass of a plane / of the 3-D grid equals the column sums of dL/dz -- whatever the cell collisions, the
    tile-binning or the flush order did (tests/test_gpu_parity.py promised this check in round 1)."""
    name, cfg, p, coords, tsteps, gt, exact, f16 = case
    dz = (exact if mode == "fp32" else f16)[4]
    m = make_model(cfg, p, mode=mode)
    m.fwd_loss_bwd({"all_coords": coords.cuda()[None], "temporal_steps": tsteps.cuda()[None]}, gt.cuda())
    named = dict(m.named_parameters())
    L, F2, F3 = cfg.n_levels, cfg.n_features, cfg.sparse_features
    pw = L * F2
    scale = float(dz.abs().sum(dim=0).max())     # mass is a signed sum: compare against the absolute column mass
    tol = 1e-5 if mode == "fp32" else 2e-3
    for k, key in enumerate(("keyframes_xy.params", "keyframes_yt.params", "keyframes_xt.params")):
        mass = named[key].grad.detach().reshape(-1, F2).sum(dim=0, dtype=torch.float64).cpu()
        want = dz[:, k * pw:(k + 1) * pw].reshape(-1, L, F2).sum(dim=(0, 1))
        assert float((mass - want).abs().max()) <= tol * scale * L, (name, mode, key)
    mass = named["sparse_grid.embeddings"].grad.detach().reshape(-1, F3).sum(dim=0, dtype=torch.float64).cpu()
    want = dz[:, 3 * pw:3 * pw + 9 * F3].reshape(-1, 9, F3).sum(dim=(0, 1))
    assert float((mass - want).abs().max()) <= tol * scale * 9, (name, mode, "sparse")
