"""Golden loss curve for the PSNR-at-equal-steps parity test at BASELINE configs[0]: synthetic 64x64x8 video, config_nvp_s,
N = 1,245,184 samples per step (dataio.py:91), 300 AdamW steps with the reference's schedule (training.py:13-14) --
the ORACLE on CPU (fp32 torch autograd over the restated forward: the reference's arithmetic).

  python oracle/make_golden_training.py [steps]     -> tests/golden/train_curve_s_64x64x8.npz   (about 15 s per step on 8 cores)

tests/test_gpu_training_long.py replays the identical sampler stream and initial weights through the CUDA path (fp32 and
tensor-core modes) and compares loss / PSNR step by step.  TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nvp_oracle as O  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
LR, SEED = 1e-2, 0
T, H, W = 8, 64, 64


def batches(steps):
    """The reference sampler stream (dataio.py:104-120) under torch.manual_seed(SEED)."""
    vid = torch.from_numpy(O.synthetic_video(T, H, W, seed=0)).reshape(T, H * W, 3)
    mgrid = O.get_mgrid_2d(H, W)
    g = torch.Generator().manual_seed(SEED)
    for _ in range(steps):
        yield O.sample_batch(vid, mgrid, O.N_SAMPLES_PER_STEP, generator=g)


def main():
    torch.set_num_threads(max(1, (os.cpu_count() or 2) - 2))
    cfg = O.NVPConfig()
    p = {k: v.clone().requires_grad_(True) for k, v in O.init_params(cfg, seed=SEED).items()}
    opt = torch.optim.AdamW(list(p.values()), lr=LR, weight_decay=0.001)
    sched = torch.optim.lr_scheduler.CosineAnnealingLR(opt, T_max=STEPS, eta_min=1e-5)
    losses = []
    t0 = time.time()
    for i, (coords, tsteps, img) in enumerate(batches(STEPS)):
        opt.zero_grad(set_to_none=True)
        loss = O.image_mse(O.nvp_forward(p, coords, tsteps, cfg), O.normalise_gt(img))
        loss.backward()
        opt.step()
        sched.step()
        losses.append(float(loss.detach()))
        if i % 10 == 0 or i == STEPS - 1:
            print(f"step {i}: loss {losses[-1]:.6f} psnr {O.psnr_from_mse(losses[-1]):.3f} dB  ({time.time() - t0:.0f} s)", flush=True)
            np.savez(os.path.join(ROOT, "tests", "golden", "train_curve_s_64x64x8.npz"), losses=np.asarray(losses, np.float64),
                     steps=STEPS, lr=LR, seed=SEED, video=np.asarray([T, H, W]), n_samples=O.N_SAMPLES_PER_STEP)


if __name__ == "__main__":
    main()
