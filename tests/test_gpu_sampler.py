"""GPU tests of the device-resident sampler against the kept host sampler (dataio.py:104-120)."""
import numpy as np
import pytest
import torch

from nvp_b200 import dataio
from oracle import nvp_oracle as O

pytestmark = pytest.mark.gpu


def test_parity_mode_reproduces_the_reference_batch_bit_for_bit():
    T, H, W, n = 7, 33, 45, 5000
    vid = O.synthetic_video(T, H, W, seed=3)
    host = dataio.VideoTimeWrapper(dataio.VideoTime(vid), sidelength=(H, W), n_samples=n)
    torch.manual_seed(11)
    a, b = host[0]
    torch.manual_seed(11)                       # the same two randint calls the host sampler makes (dataio.py:106-107)
    t_idx = torch.randint(0, T, (n,))
    p_idx = torch.randint(0, H * W, (n,))
    dev = dataio.DeviceSampler(vid, n_samples=n)
    x, y = dev.sample_indices(t_idx, p_idx)
    assert torch.equal(x["all_coords"][0].cpu(), a["all_coords"])
    assert torch.equal(x["temporal_steps"][0].cpu(), a["temporal_steps"])
    assert torch.equal(y["img"][0].cpu(), b["img"])


def test_throughput_mode_is_uniform_deterministic_and_consistent():
    T, H, W, n = 10, 24, 32, 200000
    vid = O.synthetic_video(T, H, W, seed=4)
    dev = dataio.DeviceSampler(vid, n_samples=n, seed=5, t_range=(2, 8))
    x, y, (ti, pi) = dev.sample(3, want_indices=True)
    x2, y2 = dev.sample(3)
    x3, _ = dev.sample(4)
    assert torch.equal(x["all_coords"], x2["all_coords"]) and torch.equal(y["img"], y2["img"])
    assert not torch.equal(x["all_coords"], x3["all_coords"])
    ti, pi = ti.cpu().long(), pi.cpu().long()
    assert int(ti.min()) >= 2 and int(ti.max()) <= 7 and int(pi.min()) >= 0 and int(pi.max()) < H * W
    # uniformity: chi-square of the frame histogram (6 bins) and of 16 pixel bins stays within 5 sigma
    for idx, bins in ((ti - 2, 6), (pi * 16 // (H * W), 16)):
        cnt = torch.bincount(idx, minlength=bins).double()
        chi2 = float(((cnt - n / bins) ** 2 / (n / bins)).sum())
        assert chi2 < bins + 5 * (2 * bins) ** 0.5 + 10, (chi2, bins)
    # the batch equals what the host formulas give for the drawn indices
    grid = dataio.get_mgrid((H, W), 2)
    assert torch.equal(x["all_coords"][0, :, 1:].cpu(), grid[pi])
    assert torch.equal(x["all_coords"][0, :, 0].cpu(), torch.linspace(0, 1, T)[ti])
    assert torch.equal(x["temporal_steps"][0].cpu(), torch.linspace(0.5 / T, 1 - 0.5 / T, T)[ti])
    assert torch.equal(y["img"][0].cpu(), torch.from_numpy(vid).view(T, -1, 3)[ti, pi])


def test_device_prefetcher_preserves_order_and_content():
    """Double-buffered H2D staging: batches arrive in order, bit-identical, while a consumer kernel is still running on
    the previous slot (the consumer below is slow on purpose)."""
    from nvp_b200.dataio import DevicePrefetcher
    g = torch.Generator().manual_seed(0)
    host = [(torch.rand(50000, 3, generator=g).pin_memory(), torch.randint(0, 255, (50000, 3), generator=g, dtype=torch.uint8))
            for _ in range(7)]
    pf = DevicePrefetcher(iter(host), device="cuda")
    sums, seen = [], 0
    for i, (c, u) in enumerate(pf):
        acc = c.double().sum() + u.double().sum()
        for _ in range(20):
            acc = acc + (c.double() * 1e-9).sum()      # keep the slot busy
        sums.append(acc)
        pf.release()
        seen += 1
    assert seen == len(host)
    for (c, u), got in zip(host, sums):
        ref = c.double().sum() + u.double().sum() + 20 * (c.double() * 1e-9).sum()
        assert abs(float(got) - float(ref)) <= 1e-6 * abs(float(ref))
