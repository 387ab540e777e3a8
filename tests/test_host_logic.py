"""CPU tests of the kept host code (sampler, training glue) and of the multi-GPU plumbing (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nvp_b200 import dataio
from nvp_b200.dist import attach_flat_grads, shard_range
from oracle import check_vs_reference as R
from oracle import nvp_oracle as O


def test_sampler_matches_oracle_and_reference_stream():
    T, H, W, n = 5, 12, 16, 1000
    vid = O.synthetic_video(T, H, W, seed=1)
    ds = dataio.VideoTime(vid)
    w = dataio.VideoTimeWrapper(ds, sidelength=ds.shape, n_samples=n)
    torch.manual_seed(7)
    a, b = w[0]
    torch.manual_seed(7)
    c, ts, img = O.sample_batch(torch.from_numpy(vid).view(T, -1, 3), O.get_mgrid_2d(H, W), n)
    assert torch.equal(a["all_coords"], c) and torch.equal(a["temporal_steps"], ts) and torch.equal(b["img"], img)
    assert a["all_coords"].dtype == torch.float32 and b["img"].dtype == torch.uint8
    assert w.N_samples == n and dataio.VideoTimeWrapper(ds, sidelength=ds.shape).N_samples == 1245184  # dataio.py:91
    if R.reference_available():
        ref = R.import_reference()

        class DS:
            nframes, channels, shape = T, 3, (H, W)

            def __len__(self):
                return 1

            def __getitem__(self, i):
                return vid

        rw = ref.dataio.VideoTimeWrapper(DS(), sidelength=(H, W))
        rw.N_samples = n
        torch.manual_seed(7)
        ra, rb = rw[0]
        assert torch.equal(ra["all_coords"], a["all_coords"]) and torch.equal(rb["img"], b["img"])
        assert torch.equal(ref.dataio.get_mgrid((H, W), 2), dataio.get_mgrid((H, W), 2))


@pytest.mark.parametrize("n,world", [(1245184, 8), (10, 4), (3, 8), (0, 2), (1000, 3)])
def test_shard_ranges_partition_the_batch(n, world):
    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a, b), (c, d) in zip(spans, spans[1:]):
        assert b == c and a <= b and c <= d
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_flat_gradient_buffer_aliases_param_grads():
    m = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    flat = attach_flat_grads(m)
    assert flat.numel() % 64 == 0
    for p in m.parameters():
        assert p.grad.data_ptr() >= flat.data_ptr() and (p.grad.data_ptr() - flat.data_ptr()) % 256 == 0
        p.grad.add_(1.0)
    assert float(flat.sum()) == sum(p.numel() for p in m.parameters())
    flat.zero_()
    assert all(float(p.grad.abs().sum()) == 0.0 for p in m.parameters())


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvp_b200.dist import all_reduce_grads, broadcast_parameters
    torch.manual_seed(rank)  # different init per rank -> broadcast must equalise
    m = torch.nn.Linear(6, 4)
    broadcast_parameters(m, 0)
    flat = attach_flat_grads(m)
    # emulate the sharded step on CPU: each rank contributes the gradient of ITS shard of a global batch with the
    # global-mean scaling (what the kernels do with n_global), then one all-reduce
    g = torch.Generator().manual_seed(123)
    x, y = torch.randn(10, 6, generator=g), torch.randn(10, 4, generator=g)
    lo, hi = shard_range(10, rank, world)
    loss = ((m(x[lo:hi]) - y[lo:hi]) ** 2).sum() / (10 * 4)
    gw, gb = torch.autograd.grad(loss, [m.weight, m.bias])
    m.weight.grad.add_(gw)
    m.bias.grad.add_(gb)
    all_reduce_grads(flat)
    full = ((m(x) - y) ** 2).mean()
    fw, fb = torch.autograd.grad(full, [m.weight, m.bias])
    ok = torch.allclose(m.weight.grad, fw, atol=1e-6) and torch.allclose(m.bias.grad, fb, atol=1e-6)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_shards_sum_to_the_global_gradient():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def _grid_first_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from nvp_b200.dist import GridFirstAllReduce, grid_grad_numel
    import types
    # a stand-in with the attribute names grid_grad_numel looks at: three keyframe planes + the sparse grid first,
    # then two MLP tensors, all with gradients in one flat buffer (attach_flat_grads order = parameters() order)
    m = torch.nn.Module()
    for name, n in (("keyframes_xy", 70), ("keyframes_yt", 70), ("keyframes_xt", 70)):
        enc = torch.nn.Module(); enc.params = torch.nn.Parameter(torch.zeros(n)); setattr(m, name, enc)
    m.sparse_grid = torch.nn.Module(); m.sparse_grid.embeddings = torch.nn.Parameter(torch.zeros(3, 4, 5, 2))
    m.net = torch.nn.Linear(5, 3)
    flat = attach_flat_grads(m)
    k = grid_grad_numel(m, flat)
    g = torch.Generator().manual_seed(100 + rank)
    flat.copy_(torch.randn(flat.numel(), generator=g))
    expect = sum(torch.randn(flat.numel(), generator=torch.Generator().manual_seed(100 + r)) for r in range(world))
    ar = GridFirstAllReduce(flat, k)
    ar.run()
    ok = k == m.net.weight.grad.data_ptr() // 4 - flat.data_ptr() // 4 and k >= 3 * 70 + 120
    ok = ok and torch.allclose(flat, expect, atol=1e-6)
    # degenerate split: everything in the grid piece
    flat2 = torch.full((64,), float(rank + 1))
    GridFirstAllReduce(flat2, 64).run()
    ok = ok and bool((flat2 == sum(range(1, world + 1))).all())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_two_rank_gloo_grid_first_all_reduce_equals_one_all_reduce():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_grid_first_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_slab_routing_partitions_the_reference_batch():
    """trainer.route_to_slab: every sample goes to exactly one rank, the one owning the frame SparseGrid.forward reads
    (sparsegrid.py:43-46, restated by the oracle), for video lengths equal to and different from the grid's t_resolution."""
    from nvp_b200 import trainer
    from nvp_b200.dist import t_slab
    for T_video, t_res, world in ((600, 600, 8), (8, 600, 4), (300, 300, 3), (17, 5, 2)):
        g = torch.Generator().manual_seed(T_video)
        n = 20000
        t_idx = torch.randint(0, T_video, (n,), generator=g)
        coords = torch.stack((torch.linspace(0, 1, T_video)[t_idx], torch.rand(n, generator=g), torch.rand(n, generator=g)), dim=1)
        coords[:3, 0] = torch.tensor([0.0, 1.0, 0.5])
        x = {"all_coords": coords[None], "temporal_steps": torch.rand(1, n, generator=g)}
        gt = torch.randint(0, 256, (1, n, 3), generator=g, dtype=torch.uint8)
        frames = torch.from_numpy(O.sparse_grid_indices(coords.numpy(), t_res, 4, 4)[0])
        assert torch.equal(trainer.nearest_frame(coords[:, 0], t_res), frames)
        total, seen = 0, torch.zeros(n, dtype=torch.int32)
        for r in range(world):
            xi, gi = trainer.route_to_slab(x, gt, t_res, r, world)
            lo, hi = t_slab(t_res, r, world)
            fr = trainer.nearest_frame(xi["all_coords"][0, :, 0], t_res)
            assert bool(((fr >= lo) & (fr < hi)).all())
            keep = (frames >= lo) & (frames < hi)
            assert torch.equal(xi["all_coords"][0], coords[keep]) and torch.equal(gi[0], gt[0][keep])
            assert torch.equal(xi["temporal_steps"][0], x["temporal_steps"][0][keep])
            seen += keep.int()
            total += int(keep.sum())
        assert total == n and bool((seen == 1).all())


def _slab_sync_worker(rank, world, port, out):
    import types
    import torch.distributed as dist
    from nvp_b200 import trainer
    from nvp_b200.dist import t_slab
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    T = 7
    truth = torch.arange(T * 3 * 2 * 2, dtype=torch.float32).reshape(T, 3, 2, 2)
    emb = torch.full_like(truth, -1.0)                    # stale copy everywhere ...
    lo, hi = t_slab(T, rank, world)
    emb[lo:hi] = truth[lo:hi]                             # ... except the frames this rank owns and keeps up to date
    tr = object.__new__(trainer.FusedTrainer)
    tr.distributed, tr._synced, tr.world, tr.group, tr.t_resolution = True, False, world, None, T
    tr.model = types.SimpleNamespace(sparse_grid=types.SimpleNamespace(embeddings=torch.nn.Parameter(emb)))
    tr.sync_slabs()
    out[rank] = bool(torch.equal(tr.model.sparse_grid.embeddings.data, truth))
    dist.destroy_process_group()


def test_two_rank_gloo_slab_gather_completes_the_grid():
    """Checkpoints of a slab-trained model (training.py:36,66,90): after sync_slabs every rank holds every owner's frames."""
    import torch.multiprocessing as mp
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_slab_sync_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
