"""GPU tests of nvp_b200.trainer / training.train: the product training step with the fused
optimiser, the device-side sampler, and -- with two GPUs -- the t-slab data-parallel scheme against the one-GPU run."""
import math
import os
import socket

import numpy as np
import pytest
import torch

from oracle import nvp_oracle as O
from tests.helpers import make_model

pytestmark = pytest.mark.gpu

T, Hh, Ww, N, STEPS, LR = 12, 48, 64, 16384, 6, 1e-2


def close_after_adam(a, b):
    """Two runs of a few AdamW steps that differ only in summation order: the update lr * m / sqrt(v) is discontinuous in
    the gradient's sign, so a handful of entries with gradients at round-off level move differently; everything else agrees."""
    a, b = torch.as_tensor(a).double().reshape(-1), torch.as_tensor(b).double().reshape(-1)
    scale = float(b.abs().max()) + 1e-12
    l2 = float((a - b).norm() / (b.norm() + 1e-30))
    off = float(((a - b).abs() > 1e-3 * scale).double().mean())
    return l2 <= 2e-3 and off <= 2e-3, (l2, off)


def _batches():
    vid = torch.from_numpy(O.synthetic_video(T, Hh, Ww, seed=0)).reshape(T, Hh * Ww, 3)
    g = torch.Generator().manual_seed(3)
    return [O.sample_batch(vid, O.get_mgrid_2d(Hh, Ww), N, generator=g) for _ in range(STEPS)]


def _cfg():
    return O.NVPConfig(t_resolution=T, x_resolution=24, y_resolution=32)


def _run(mode, device, fused_optimizer=True):
    from nvp_b200.trainer import FusedTrainer
    cfg = _cfg()
    m = make_model(cfg, O.init_params(cfg, seed=0, grid_std=0.05), mode=mode, device=device)
    tr = FusedTrainer(m, lr=LR, total_steps=STEPS, fused_optimizer=fused_optimizer)
    losses = []
    for c, t, g in _batches():
        x = {"all_coords": c.to(device)[None], "temporal_steps": t.to(device)[None]}
        losses.append(float(tr.step(x, g.to(device)[None])))
    return losses, {k: v.detach().cpu().clone() for k, v in tr.model_state_dict().items()}


def test_fused_trainer_matches_torch_adamw_on_one_gpu():
    """Same step with the fused AdamW + closed-form cosine schedule vs torch.optim.AdamW + CosineAnnealingLR (training.py:13-14)."""
    la, sa = _run("fp32", "cuda:0", fused_optimizer=True)
    lb, sb = _run("fp32", "cuda:0", fused_optimizer=False)
    np.testing.assert_allclose(la, lb, rtol=2e-4)
    assert la[-1] < la[0]
    for k in sa:
        ok, info = close_after_adam(sa[k], sb[k])
        assert ok, (k, info)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _ddp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    losses, sd = _run("fp32", f"cuda:{rank}")
    out[rank] = (losses, {k: v.numpy() for k, v in sd.items()})
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_slab_training_equals_one_gpu():
    """t-slab data parallelism on 2 GPUs: same global batches, each rank takes the samples of its frames; losses and the
    gathered state_dict (every rank's, after sync_slabs) equal the one-GPU run up to summation order."""
    import torch.multiprocessing as mp
    ref_losses, ref_sd = _run("fp32", "cuda:0")
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_ddp_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        for r in range(world):
            losses, sd = out[r]
            np.testing.assert_allclose(losses, ref_losses, rtol=2e-4, err_msg=f"rank {r}")
            for k, v in ref_sd.items():
                ok, info = close_after_adam(sd[k], v)
                assert ok, (r, k, info)


def test_training_loop_with_device_sampler_learns(tmp_path):
    """training.train(device_sampler=True): video resident in HBM, batches drawn by the sampler kernel, fused optimiser;
    writes the reference's checkpoint files (training.py:90-94) with its state_dict keys."""
    from torch.utils.data import DataLoader
    from nvp_b200 import dataio, training
    cfg = _cfg()
    vid = O.synthetic_video(T, Hh, Ww, seed=0)
    ds = dataio.VideoTime(vid)
    w = dataio.VideoTimeWrapper(ds, sidelength=ds.shape, n_samples=N)
    dl = DataLoader(w, batch_size=1, shuffle=True, pin_memory=True, num_workers=0)
    torch.manual_seed(0)
    m = make_model(cfg, O.init_params(cfg, seed=0), mode="tc")
    _, losses = training.train(m, dl, epochs=40, lr=LR, steps_til_summary=10 ** 9, epochs_til_checkpoint=10 ** 9,
                               model_dir=str(tmp_path / "run"), device_sampler=True)
    assert len(losses) == 40 and losses[-1] < 0.5 * losses[0]
    ck = torch.load(tmp_path / "run" / "checkpoints" / "model_final.pth", weights_only=False)
    assert set(ck) >= {"epoch", "model", "optimizer", "scheduler"}
    assert "wrapper.net.layers.0.weight" in ck["model"] and ck["model"]["sparse_grid.embeddings"].shape == (T, 24, 32, 2)
    assert 10 * math.log10(4 / losses[-1]) > 10 * math.log10(4 / losses[0])
