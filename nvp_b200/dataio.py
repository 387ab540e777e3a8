"""Coordinate sampler and video container — host code kept from the reference (dataio.py:11-120).

Same classes, same tensors, same torch-RNG consumption (two `torch.randint` calls per batch, t first), so
a seeded run draws exactly the reference's sample stream.  `.mp4` input needs scikit-video (absent in this
image) and is gated; `.npy` and PNG directories work.
"""
import glob
import os

import numpy as np
import torch
from torch.utils.data import Dataset


def get_mgrid(sidelen, dim=2):
    """Flattened grid of coordinates in [0, 1] (dataio.py:11-29)."""
    if isinstance(sidelen, int):
        sidelen = dim * (sidelen,)
    if dim == 2:
        pixel_coords = np.stack(np.mgrid[:sidelen[0], :sidelen[1]], axis=-1)[None, ...].astype(np.float32)
        pixel_coords[0, :, :, 0] = pixel_coords[0, :, :, 0] / (sidelen[0] - 1)
        pixel_coords[0, :, :, 1] = pixel_coords[0, :, :, 1] / (sidelen[1] - 1)
    elif dim == 3:
        pixel_coords = np.stack(np.mgrid[:sidelen[0], :sidelen[1], :sidelen[2]], axis=-1)[None, ...].astype(np.float32)
        pixel_coords[..., 0] = pixel_coords[..., 0] / max(sidelen[0] - 1, 1)
        pixel_coords[..., 1] = pixel_coords[..., 1] / (sidelen[1] - 1)
        pixel_coords[..., 2] = pixel_coords[..., 2] / (sidelen[2] - 1)
    else:
        raise NotImplementedError('Not implemented for dim=%d' % dim)
    return torch.Tensor(pixel_coords).view(-1, dim)


class VideoTime(Dataset):
    """uint8 video [T,H,W,3] from a .npy file, an in-memory array, or a directory of PNGs (dataio.py:33-71)."""

    def __init__(self, path_to_video, split_num=300):
        super().__init__()
        self.split_num = split_num
        if isinstance(path_to_video, np.ndarray):
            self.vid = path_to_video
        elif 'npy' in path_to_video:
            self.vid = np.load(path_to_video)
        elif 'mp4' in path_to_video:
            raise NotImplementedError("mp4 input needs scikit-video, which is not installed; convert to .npy or PNGs")
        else:
            from PIL import Image
            files = sorted(glob.glob(os.path.join(path_to_video, "*.png")))[:self.split_num]
            first = np.array(Image.open(files[0]))
            self.vid = np.zeros((self.split_num,) + first.shape, dtype=np.uint8)
            for idx, f in enumerate(files):
                self.vid[idx] = np.array(Image.open(f))
        self.shape = self.vid.shape[1:-1]
        self.nframes = self.vid.shape[0]
        self.channels = self.vid.shape[-1]

    def __len__(self):
        return 1

    def __getitem__(self, idx):
        return self.vid


class VideoTimeWrapper(torch.utils.data.Dataset):
    """Random (t, pixel) sampler, N_samples per item (dataio.py:75-120)."""

    def __init__(self, dataset, sidelength=None, n_samples=1245184):
        self.dataset = dataset
        nframes = self.dataset.nframes
        self.sidelength = sidelength
        self.mgrid = get_mgrid(sidelength, dim=2)
        data = torch.from_numpy(self.dataset[0])
        self.data = data.view(self.dataset.nframes, -1, self.dataset.channels)
        self.N_samples = n_samples
        half_dt = 0.5 / nframes
        self.temporal_steps = torch.linspace(half_dt, 1 - half_dt, self.dataset.nframes)
        self.temporal_coords = torch.linspace(0, 1, nframes)

    def __len__(self):
        return len(self.dataset)

    def __getitem__(self, idx):
        temporal_coord_idx = torch.randint(0, self.data.shape[0], (self.N_samples,))
        spatial_coord_idx = torch.randint(0, self.data.shape[1], (self.N_samples,))
        data = self.data[temporal_coord_idx, spatial_coord_idx, :]
        spatial_coords = self.mgrid[spatial_coord_idx, :]
        temporal_coords = self.temporal_coords[temporal_coord_idx]
        temporal_steps = self.temporal_steps[temporal_coord_idx]
        all_coords = torch.cat((temporal_coords.unsqueeze(1), spatial_coords), dim=1)
        return {'all_coords': all_coords, "temporal_steps": temporal_steps}, {'img': data}


class DeviceSampler:
    """VideoTimeWrapper with the video resident in HBM and the batch produced by one kernel.

    sample(step)                      throughput mode: Philox stream keyed by (seed, step); no host work, no H2D.
    sample_indices(t_idx, p_idx)      parity mode: the reference's CPU index stream (dataio.py:106-107) uploaded;
                                      the batch is bit-identical to VideoTimeWrapper.__getitem__.
    Both return ({'all_coords': [1,N,3], 'temporal_steps': [1,N]}, {'img': uint8 [1,N,3]}) on the device.
    """

    def __init__(self, video, n_samples=1245184, device="cuda", seed=0, t_range=None):
        from . import _lib
        self._lib = _lib
        vid = torch.as_tensor(video)
        assert vid.dtype == torch.uint8 and vid.dim() == 4 and vid.shape[-1] == 3
        self.T, self.H, self.W = int(vid.shape[0]), int(vid.shape[1]), int(vid.shape[2])
        self.video = vid.reshape(self.T, self.H * self.W, 3).to(device).contiguous()
        half_dt = 0.5 / self.T
        self.temporal_steps = torch.linspace(half_dt, 1 - half_dt, self.T).to(device)
        self.temporal_coords = torch.linspace(0, 1, self.T).to(device)
        self.N_samples, self.seed = n_samples, seed
        self.t_range = t_range or (0, self.T)
        self.device = self.video.device

    def _run(self, n, t_idx, p_idx, step, want_indices=False):
        dev = self.device
        coords = torch.empty(n, 3, dtype=torch.float32, device=dev)
        tsteps = torch.empty(n, dtype=torch.float32, device=dev)
        gt = torch.empty(n, 3, dtype=torch.uint8, device=dev)
        ti = torch.empty(n, dtype=torch.int32, device=dev) if want_indices else None
        pi = torch.empty(n, dtype=torch.int32, device=dev) if want_indices else None
        ptr = lambda t: None if t is None else t.data_ptr()  # noqa: E731
        with torch.cuda.device(dev):
            rc = self._lib.load().nvp_sample_batch(
                self.video.data_ptr(), self.T, self.H, self.W, self.temporal_coords.data_ptr(), self.temporal_steps.data_ptr(),
                n, ptr(t_idx), ptr(p_idx), self.seed, step, self.t_range[0], self.t_range[1], coords.data_ptr(),
                tsteps.data_ptr(), gt.data_ptr(), ptr(ti), ptr(pi), torch.cuda.current_stream(dev).cuda_stream)
        self._lib.check(rc, "nvp_sample_batch")
        out = ({"all_coords": coords[None], "temporal_steps": tsteps[None]}, {"img": gt[None]})
        return out + ((ti, pi),) if want_indices else out

    def sample(self, step: int, want_indices=False):
        return self._run(self.N_samples, None, None, step, want_indices)

    def sample_indices(self, t_idx: torch.Tensor, p_idx: torch.Tensor):
        t_idx = t_idx.to(self.device, torch.int64).contiguous()
        p_idx = p_idx.to(self.device, torch.int64).contiguous()
        return self._run(t_idx.numel(), t_idx, p_idx, 0)


class DevicePrefetcher:
    """Host -> device double buffering for the training loop (replaces the reference's per-step `.cuda()` copies,
    training.py:45-46): batch i+1 is copied from pinned host memory on a side stream while step i computes.

        pf = DevicePrefetcher(iterable_of_host_batches)          # each batch: tuple of CPU tensors
        for dev_batch in pf: step(*dev_batch)

    The consumer's stream waits on the copy's event before using a slot, and a slot is only overwritten after the
    step that read it has been enqueued and recorded, so no host synchronisation is needed beyond the caller's own.
    """

    def __init__(self, batches, device="cuda", depth=2):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(self.device)
        self.depth = depth
        self.slots = [None] * depth
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.done = [None] * depth
        self.head = 0       # next slot to fill
        self.tail = 0       # next slot to hand out
        self.inflight = 0

    def _fill(self):
        try:
            host = next(self.it)
        except StopIteration:
            return False
        k = self.head
        with torch.cuda.stream(self.copy_stream):
            if self.done[k] is not None:
                self.copy_stream.wait_event(self.done[k])      # the step that read this slot has finished
            if self.slots[k] is None or any(d.shape != t.shape or d.dtype != t.dtype for d, t in zip(self.slots[k], host)):
                # (batch sizes vary when a global batch is routed to t-slab owners: trainer.route_to_slab)
                self.slots[k] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in host)
            for dst, src in zip(self.slots[k], host):
                dst.copy_(src if src.is_pinned() else src.pin_memory(), non_blocking=True)
            self.ready[k].record(self.copy_stream)
        self.head = (k + 1) % self.depth
        self.inflight += 1
        return True

    def __iter__(self):
        return self

    def __next__(self):
        while self.inflight < self.depth - 1 and self._fill():
            pass
        if self.inflight == 0 and not self._fill():
            raise StopIteration
        k = self.tail
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(self.ready[k])
        self._fill()                                           # next batch's copy overlaps this step
        self.tail = (k + 1) % self.depth
        self.inflight -= 1
        self._pending = k
        return self.slots[k]

    def release(self):
        """Call after enqueueing the step that consumes the batch last returned."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.done[self._pending] = ev
