"""GPU test: same PSNR at equal step count on identical seeds (north star), config[0]-style plumbing case.

A tiny synthetic video is fitted for a few AdamW steps three times from identical initial weights and the
identical sampler stream: (a) the oracle on CPU (torch autograd + torch.optim.AdamW: the reference's arithmetic),
(b) the product in fp32 mode, (c) the product in tc mode, all through nvp_b200.training.train's fused step.
"""
import math

import numpy as np
import pytest
import torch
from torch.utils.data import DataLoader

from nvp_b200 import dataio, training
from oracle import nvp_oracle as O
from tests.helpers import make_model

pytestmark = pytest.mark.gpu

STEPS, N, LR = 24, 8192, 1e-2


def oracle_training(cfg, p0, batches):
    p = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt = torch.optim.AdamW(list(p.values()), lr=LR, weight_decay=0.001)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=STEPS, eta_min=1e-5)
    losses = []
    for coords, tsteps, img in batches:
        opt.zero_grad()
        rgb = O.nvp_forward(p, coords, tsteps, cfg)
        loss = O.image_mse(rgb, O.normalise_gt(img))
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss.detach()))
    return losses


class Replay(torch.utils.data.Dataset):
    def __init__(self, batches):
        self.b = batches

    def __len__(self):
        return 1

    def __getitem__(self, i):
        c, t, g = self.b.pop(0)
        return {"all_coords": c, "temporal_steps": t}, {"img": g}


def product_training(cfg, p0, batches, mode, tmp_path):
    m = make_model(cfg, p0, mode=mode)
    dl = DataLoader(Replay(list(batches)), batch_size=1, shuffle=False)
    _, losses = training.train(m, dl, epochs=STEPS, lr=LR, steps_til_summary=10 ** 9, epochs_til_checkpoint=10 ** 9,
                               model_dir=str(tmp_path / mode), fused=True)
    return losses


def test_loss_curve_and_psnr_match_the_oracle_at_equal_steps(tmp_path):
    T, Hh, Ww = 8, 64, 64
    cfg = O.NVPConfig(t_resolution=T, x_resolution=32, y_resolution=32)
    vid = O.synthetic_video(T, Hh, Ww, seed=0)
    ds = dataio.VideoTime(vid)
    w = dataio.VideoTimeWrapper(ds, sidelength=ds.shape, n_samples=N)
    torch.manual_seed(0)
    batches = []
    for _ in range(STEPS):
        a, b = w[0]
        batches.append((a["all_coords"], a["temporal_steps"], b["img"]))
    p0 = O.init_params(cfg, seed=0)
    ref = oracle_training(cfg, p0, batches)
    assert ref[-1] < 0.5 * ref[0], "the oracle run must actually learn for this test to mean anything"
    psnr_ref = 10 * math.log10(4 / ref[-1])
    for mode, rtol, dpsnr in (("fp32", 2e-3, 0.02), ("tc", 3e-2, 0.15)):
        got = product_training(cfg, p0, batches, mode, tmp_path)
        assert len(got) == STEPS
        np.testing.assert_allclose(got[0], ref[0], rtol=1e-5 if mode == "fp32" else 1e-3)
        np.testing.assert_allclose(got, ref, rtol=rtol, err_msg=mode)
        assert abs(10 * math.log10(4 / got[-1]) - psnr_ref) <= dpsnr, mode


def test_eval_helpers_render_quantise_and_psnr():
    """eval.py path: quantise the grids to 8 bit (per level / feature min-max), render a frame in slices, PSNR."""
    from nvp_b200 import eval_utils
    T, Hh, Ww = 8, 40, 50
    cfg = O.NVPConfig(t_resolution=T, x_resolution=16, y_resolution=16)
    p = O.init_params(cfg, seed=2, grid_std=0.3)
    m = make_model(cfg, p, mode="tc")
    img = eval_utils.render_frame(m, 3, T, (Hh, Ww), n_slices=10)
    assert img.shape == (3, Hh, Ww) and float(img.min()) >= 0 and float(img.max()) <= 1
    # oracle for the same frame
    coords = torch.cat((torch.linspace(0, 1, T)[3] * torch.ones(Hh * Ww, 1), dataio.get_mgrid((Hh, Ww), 2)), dim=1)
    tsteps = torch.linspace(0.5 / T, 1 - 0.5 / T, T)[3] * torch.ones(Hh * Ww)
    ref = torch.clamp((O.nvp_forward(p, coords, tsteps, cfg).view(Hh, Ww, 3).permute(2, 0, 1) + 1) / 2, 0, 1)
    assert float((img.cpu() - ref).abs().max()) <= 1e-3
    assert eval_utils.psnr(img.cpu(), ref) > 60
    before = m.keyframes_xy.params.detach().clone()
    eval_utils.quantize_model(m)
    err = (m.keyframes_xy.params.detach() - before).abs().max()
    assert 0 < float(err) <= 0.3 * 2 / 255 + 1e-6        # at most one 8-bit step of the widest level range
    q = eval_utils.render_frame(m, 3, T, (Hh, Ww), n_slices=10)
    assert eval_utils.psnr(q.cpu(), ref) > 30
    # temporal interpolation rendering (eval.py --t_interp): frames between key frames are finite away from t=1
    mid = eval_utils.render_frame(m, 5, 2 * T, (Hh, Ww), org_nframes=T, temporal_interp=True, n_slices=10)
    assert torch.isfinite(mid).all()
