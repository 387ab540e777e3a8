#!/usr/bin/env python
"""Benchmark of the NVP per-coordinate hot path (BASELINE.json metric: Mpixels/s fwd+bwd @1920x1080x600).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config s|l]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...        (N > 1)

A "step" is one pass of the hot path over one batch of N=1,245,184 sampled coordinates (dataio.py:91)
of a synthetic 1920x1080x600 video: zero the gradient buffer, sample bucketing, positional-feature gather,
ONE fused kernel for modulator+SIREN forward, L2 loss and backward, grid scatter-add, weight gradients and — with
more than one GPU — one NCCL all-reduce of the keyframe + MLP gradients (the 3-D grid is owned per rank by t-slab,
nvp_b200.trainer.FusedTrainer; --scaling strong splits the reference's ONE batch over the GPUs instead of giving each
its own).  Prints ONE JSON line (rank 0).

  value      Mpixels/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e        same metric through the public module API with HOST (pinned) inputs: H2D copies of
             coords/tsteps/gt and the D2H read of every step's loss are inside the timed region
  with_optimizer / unfused_api   the same step + fused AdamW; the reference's model() / loss / backward() sequence
  roofline   the dominant kernel's achieved algorithmic rate vs the measured peak (MEASURED_PEAKS.json),
             timed live with CUDA events on the launch stream; all kernels listed under "kernels"
  cpu_baseline  the oracle (CPU port of the reference arithmetic) timed on this box's host cores on a
             bounded sample of the same workload
--impl reference times that CPU port alone (the reference itself is CPU-runnable only through the
oracle here: its sole native dependency, a tiny-cuda-nn fork, is un-vendored and the reference tree is
not present on the GPU box).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SAMPLES = 1245184            # dataio.py:91
VIDEOS = {"1080p": (600, 1080, 1920), "4k": (600, 2160, 3840)}   # T, H, W (BASELINE.json configs[2..4])
VIDEO = VIDEOS["1080p"]
METRIC = "Mpixels/sec fwd+bwd @1920x1080x600"
UNIT = "Mpixels/s"


# ------------------------------------------------------------------------------------------
def load_config(name: str) -> dict:
    with open(os.path.join(ROOT, "config", f"config_nvp_{name}.json")) as f:
        return json.load(f)["nvp"]


def synth_batch(n: int, seed: int, t_range=None, video=None):
    """One sampler batch (dataio.py:104-120) over a virtual synthetic video: the pixel value is a smooth
    pattern plus hash noise evaluated at the sampled (t,row,col) — the 3.7 GB video is never materialised.
    t_range=(lo,hi) restricts the frame index to a rank's t-slab (stratified version of the uniform sampler)."""
    T, Hh, Ww = video or VIDEO
    g = torch.Generator().manual_seed(seed)
    lo, hi = t_range if t_range is not None else (0, T)
    t_idx = torch.randint(lo, hi, (n,), generator=g)
    p_idx = torch.randint(0, Hh * Ww, (n,), generator=g)
    row, col = p_idx // Ww, p_idx % Ww
    coords = torch.stack((torch.linspace(0, 1, T)[t_idx], row.float() / (Hh - 1), col.float() / (Ww - 1)), dim=1)
    half_dt = 0.5 / T
    tsteps = torch.linspace(half_dt, 1 - half_dt, T)[t_idx]
    x, y, t = coords[:, 2], coords[:, 1], coords[:, 0]
    chans = []
    for c in range(3):
        base = 0.5 + 0.25 * torch.sin(6.2831853 * ((c + 1) * x + 0.5 * t)) * torch.cos(6.2831853 * ((c + 2) * y - 0.3 * t))
        noise = (((p_idx * 2654435761 + t_idx * 40503 + c * 97) % 1024).float() / 1024.0 - 0.5) * 0.06
        chans.append(((base + noise) * 255.0).clamp(0, 255).to(torch.uint8))
    return coords.contiguous(), tsteps.contiguous(), torch.stack(chans, dim=1).contiguous()


class ClockSampler:
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        if self.nv is None or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"], "tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def measured_traffic(config: str):
    """DRAM bytes per step and kernel kind from the committed `ncu --set full` capture of prof_step.py (same workload)."""
    import glob
    paths = sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_traffic_{config}.json")))
    if paths:
        with open(paths[-1]) as f:           # newest capture (file names sort by round / session)
            return json.load(f)["dram_bytes_per_step"]
    return {}


# Algorithmic work per pixel (every gathered/scattered element counted once, no cache credit).
def algorithmic_work(F: int):
    G = (3 * 16 * 4 + 9) * F * 4          # grid gather bytes / px  (1608 B for S)
    Z = 57 * F
    fwd_mac = Z * 128 + 2 * (128 + Z) * 128 + (128 + 2 * 128 * 128 + 384)
    return {
        "grid_gather": ("hbm", G + 12),                    # + coords
        "grid_scatter": ("hbm", G + 12),
        "mlp_forward": ("tensor", 2 * fwd_mac),
        "mlp_backward": ("tensor", 2 * (fwd_mac - 128)),   # dgrad (none to the scalar SIREN input)
        "mlp_wgrad": ("tensor", 2 * fwd_mac),
        "mlp_fused": ("tensor", 2 * fwd_mac + 2 * (fwd_mac - 128)),   # forward + dgrad in one kernel
        "total_bytes": 2 * G + 31, "total_flop": 3 * 2 * fwd_mac - 2 * 128,
    }


# ------------------------------------------------------------------------------------------
def cpu_port_step_time(cfg_json: dict, n_sample: int, steps: int, warmup: int, seed: int = 0):
    """fwd + loss + bwd of the oracle (CPU restatement of the reference) on n_sample coordinates of the workload."""
    from oracle import nvp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.NVPConfig.from_json(cfg_json)
    p = O.init_params(cfg, seed=seed)
    coords, tsteps, gt = synth_batch(n_sample, seed + 1)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.nvp_loss_and_grads(p, coords, tsteps, gt, cfg, n_global=N_SAMPLES)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return sum(times) / len(times)


def run_reference(args):
    """--impl reference: the CPU port alone, same metric/config keys, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg_json = load_config(args.config)
    n_sample = N_SAMPLES // 4
    # --steps / --warmup are honoured up to a cap that keeps the run within a few minutes (a step is ~5-15 s of CPU work)
    steps, warmup = max(1, min(args.steps, 12)), max(1, min(args.warmup, 3))
    t = cpu_port_step_time(cfg_json, n_sample, steps, warmup)
    v = n_sample / t / 1e6
    cores = os.cpu_count() or 1
    sample = (f"{n_sample} of {N_SAMPLES} coordinates per step (1/4 batch), {steps} timed steps after {warmup} warm-up"
              + (f" (asked: --steps {args.steps} --warmup {args.warmup}; capped at 12 / 3)" if (steps, warmup) != (args.steps, args.warmup) else ""))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": dict(bench_config(args, "cpu oracle port"), reference_sample=sample),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = pure-PyTorch CPU path restated by oracle/ (tiny-cuda-nn DenseGrid restated; see DESIGN.md)",
    }))


def bench_config(args, mode):
    T, Hh, Ww = VIDEOS[args.video]
    strong = args.gpus > 1 and args.scaling == "strong"
    per = (f"{N_SAMPLES} sampled coordinates per step GLOBALLY (the reference's batch, dataio.py:91), split over the GPUs by "
           "ownership of the 3-D grid's frames" if strong else f"{N_SAMPLES} sampled coordinates per step per GPU")
    return {"workload": f"synthetic {Ww}x{Hh}x{T} ({'UVG Jockey stand-in' if args.video == '1080p' else 'synthetic 4K'}), "
                        f"config_nvp_{args.config}, {per}",
            "nvp_config": f"config_nvp_{args.config}", "video": args.video, "scaling": args.scaling if args.gpus > 1 else "n/a (1 GPU)",
            "samples_per_step_per_gpu": N_SAMPLES // args.gpus if strong else N_SAMPLES, "mode": mode,
            "step": "grad-buffer zero + sample bucketing + grid gather + fused MLP fwd + L2 loss + fused bwd + grid scatter-add + wgrad"
                    + ((" + NCCL all-reduce of the whole flat gradient buffer" if args.full_allreduce else
                        " + NCCL all-reduce of keyframe+MLP gradients (sparse 3-D grid owned per rank by t-slab, samples "
                        "stratified by slab)") + (", grid piece overlapped with wgrad" if getattr(args, "overlap", False) else "")
                       if args.gpus > 1 else ""),
            "l2": f"working set >> L2 ({'543' if args.config == 's' else '1086'} MB params, as much again in gradients, and GBs of "
                  "activation tiles per step); 8 rotating input batches",
            "parallelism": f"dp{args.gpus}"}


# ------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import nvp_b200
    from nvp_b200 import _lib, functional
    from nvp_b200.trainer import FusedTrainer, route_to_slab

    t_start = time.time()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # a collective that does not complete within 3 minutes aborts the job (NCCL watchdog) instead of spinning on the GPUs
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N>1)"

    cfg_json = load_config(args.config)
    video = VIDEOS[args.video]
    torch.manual_seed(0)
    model = nvp_b200.NVP(type="nvp", out_features=3, encoding_config=cfg_json, mode=args.mode).to(dev)
    # The product's training step (nvp_b200.trainer): flat parameter / gradient buffers, and with several GPUs t-slab
    # ownership of the sparse grid: its gradient sits at the end of the flat buffer, outside the
    # all-reduce, and each rank takes the samples of its own frames.  --full-allreduce: replicated everything instead.
    slab = world > 1 and not args.full_allreduce
    trainer = FusedTrainer(model, lr=1e-2, total_steps=100000, distributed=slab)
    if world > 1 and not slab:
        from nvp_b200.dist import broadcast_parameters
        broadcast_parameters(model, 0)
    flat_params, flat = trainer.flat_params, trainer.flat_grads
    reduce_view = trainer.reduce_view if slab else flat
    t_res = cfg_json["3d_encoding"]["t_resolution"]
    strong = world > 1 and args.scaling == "strong"
    n_global = N_SAMPLES if strong else N_SAMPLES * world
    F = cfg_json["2d_encoding_xy"]["n_features_per_level"]

    n_pool = 8
    host = []
    for i in range(n_pool):
        if strong:
            # the SAME global batch on every rank (the reference's sampler stream), each rank keeps the samples whose
            # nearest grid frame it owns: the union over ranks is the one-GPU batch
            c, t, g = synth_batch(N_SAMPLES, i, None, video)
            if slab:
                x, gg = route_to_slab({"all_coords": c[None], "temporal_steps": t[None]}, g[None], t_res, rank, world)
                c, t, g = x["all_coords"][0].contiguous(), x["temporal_steps"][0].contiguous(), gg[0].contiguous()
            else:
                from nvp_b200.dist import shard_range
                lo, hi = shard_range(N_SAMPLES, rank, world)
                c, t, g = c[lo:hi].contiguous(), t[lo:hi].contiguous(), g[lo:hi].contiguous()
        else:
            assert not slab or t_res == video[0]
            c, t, g = synth_batch(N_SAMPLES, 1000 * rank + i, trainer.slab if slab else None, video)
        host.append((c.pin_memory(), t.pin_memory(), g.pin_memory()))
    resident = [(c.to(dev), t.to(dev), g.to(dev)) for c, t, g in host]
    n_local = sum(c.shape[0] for c, _, _ in host) / n_pool

    def trace(msg):
        if os.environ.get("NVP_BENCH_TRACE"):
            print(f"[bench rank {rank}] {msg} ({time.time() - t_start:.1f} s)", file=sys.stderr, flush=True)
    trace(f"pool ready, {n_local:.0f} samples per step on this rank")
    loss_sum = torch.zeros(1, device=dev)
    launches = [0]

    overlap = None
    if world > 1 and args.overlap:
        from nvp_b200.dist import GridFirstAllReduce, grid_grad_numel
        overlap = GridFirstAllReduce(reduce_view, min(grid_grad_numel(model, flat), reduce_view.numel()))

    def step(c, t, g):
        trainer.zero_grads()     # (with slab ownership: the replicated gradients + the owned frames only)
        loss_sum.zero_()
        model.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g, n_global=n_global, loss_sum=loss_sum,
                           grid_event=overlap.event if overlap else None)
        launches[0] += functional.last_launch_count()
        if overlap:
            overlap.run()            # grid gradients reduce on a side stream under the wgrad kernel, MLP gradients after it
        elif world > 1:
            dist.all_reduce(reduce_view)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- device-resident measurement (value) with live per-kernel timing
    for i in range(args.warmup):
        step(*resident[i % n_pool])
    trace("warm-up done")
    sampler = ClockSampler(local)
    launches[0] = 0
    _lib.profile_enable(True)
    sampler.start()
    ms_total = timed(lambda i: step(*resident[i % n_pool]), args.steps)
    kern = _lib.profile_read()
    _lib.profile_enable(False)
    gpu_launches = launches[0]
    loss_last = float(loss_sum) / (3.0 * n_local)
    trace("timed region done")

    # ---- end-to-end through the public API with host buffers
    # every step's inputs start in pinned HOST memory; nvp_b200.dataio.DevicePrefetcher -- the host->device stage of
    # nvp_b200.training.train's loop -- copies batch i+1 on a side stream while step i computes.  All K copies, K steps
    # and K loss read-backs happen inside the timed region; the first copy is not overlapped with anything.
    from nvp_b200.dataio import DevicePrefetcher

    def host_batches():
        i = 0
        while True:
            yield host[i % n_pool]
            i += 1

    pf_box = [None]
    # the step's result (the loss sum, 4 bytes) is read back EVERY step through a pinned host buffer: the copy is
    # enqueued behind the step and the host looks at it one step later, while the next step is already running -- the
    # read of every step's loss is inside the timed region (the last one at its end), without the per-step stall of a
    # synchronous .item().  --e2e-sync-loss restores the reference's `.item()` (training.py:78).
    loss_host = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ev = [torch.cuda.Event(), torch.cuda.Event()]
    seen = []

    def e2e_step(i):
        if pf_box[0] is None:
            pf_box[0] = DevicePrefetcher(host_batches(), device=dev)
        batch = next(pf_box[0])
        step(*batch)
        pf_box[0].release()
        if args.e2e_sync_loss:
            seen.append(loss_sum.item())
            return
        k = i & 1
        loss_host[k:k + 1].copy_(loss_sum, non_blocking=True)      # D2H read of this step's result ...
        loss_ev[k].record()
        if i > 0:
            loss_ev[k ^ 1].synchronize()                            # ... consumed by the host one step later
            seen.append(float(loss_host[k ^ 1]))

    def e2e_drain(k_steps):
        if not args.e2e_sync_loss and k_steps > 0:
            loss_ev[(k_steps - 1) & 1].synchronize()
            seen.append(float(loss_host[(k_steps - 1) & 1]))

    for i in range(min(args.warmup, 3)):
        e2e_step(i)
    e2e_drain(min(args.warmup, 3))
    torch.cuda.synchronize()
    pf_box[0] = None                 # the timed loop starts with nothing prefetched
    seen.clear()

    def e2e_loop(i):
        e2e_step(i)
        if i == args.steps - 1:
            e2e_drain(args.steps)    # the last step's loss is read inside the timed region too

    ms_e2e = timed(e2e_loop, args.steps)
    assert len(seen) == args.steps, (len(seen), args.steps)
    # keep the clock sampler fed for at least ~1.5 s of the same loop (short --steps runs give NVML no samples)
    t_end = time.time() + max(0.0, 1.5 - (ms_total + ms_e2e) / 1e3)
    i = 0
    while time.time() < t_end:
        step(*resident[i % n_pool]); i += 1
        if i % 16 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    trace("end-to-end leg done")

    # ---- second line: the whole training step of nvp_b200.training.train -- FusedTrainer.step: the same
    # fwd + loss + bwd (+ all-reduce) followed by the fused AdamW update, which also clears the gradients.  With t-slabs
    # each rank updates the replicated parameters and the frames it owns only.
    flat.zero_()
    if slab or world == 1:
        def step_opt(i):
            c, t, g = resident[i % n_pool]
            trainer.step({"all_coords": c[None], "temporal_steps": t[None]}, g[None], n_global=n_global, routed=True)
    else:
        from nvp_b200.optim import FusedAdamW
        fopt = FusedAdamW(flat_params, flat, lr=1e-2, weight_decay=1e-3, t_max=100000, eta_min=1e-5)

        def step_opt(i):
            c, t, g = resident[i % n_pool]
            loss_sum.zero_()
            model.fwd_loss_bwd({"all_coords": c, "temporal_steps": t}, g, n_global=n_global, loss_sum=loss_sum)
            dist.all_reduce(flat)
            fopt.step(zero_grad=True)

    for i in range(3):
        step_opt(i)
    ms_opt = timed(step_opt, args.steps)
    trace("optimiser leg done")

    # ---- third line: the reference's own call sequence (training.py:47-52,74), unfused: model(x) -> image_mse in torch ->
    # loss.backward().  The forward runs the inference kernels, the backward recomputes it inside the fused kernel; gradients
    # are accumulated straight into .grad (functional.NvpFunction).
    ms_api = None
    if world == 1:
        from nvp_b200.loss_functions import image_mse

        def step_api(i):
            c, t, g = resident[i % n_pool]
            flat.zero_()
            out = model({"all_coords": c[None], "temporal_steps": t[None]})
            gt = {"img": ((g.float() - 127.5) / 127.5)[None]}
            image_mse(None, out, gt)["img_loss"].mean().backward()

        for i in range(3):
            step_api(i)
        ms_api = timed(step_api, args.steps)
        flat.zero_()
        trace("unfused API leg done")

    if rank == 0:
        ms_step = ms_total / args.steps
        value = n_global / (ms_step * 1e-3) / 1e6
        e2e_v = n_global / (ms_e2e / args.steps * 1e-3) / 1e6
        peaks = measured_peaks()
        work = algorithmic_work(F)
        kernels = {}
        traffic = measured_traffic(args.config)
        for name, (ms, cnt) in kern.items():
            if cnt == 0:
                continue
            per_step = ms / args.steps            # all launches of this kind in one step (gather / scatter launch two kernels)
            ent = {"launches_per_step": cnt / args.steps, "ms_per_step": per_step, "share_of_step": ms / ms_total}
            if name in work:
                bound, per_px = work[name]
                per_step_work = per_px * n_local
                if bound == "hbm":
                    ent.update(bound="hbm", achieved=per_step_work / (per_step * 1e-3) / 1e9, peak=peaks["hbm_gbs"], unit="GB/s")
                else:
                    ent.update(bound="tensor", achieved=per_step_work / (per_step * 1e-3) / 1e12, peak=peaks["tflops_sustained"], unit="TFLOP/s")
                ent["frac"] = ent["achieved"] / ent["peak"]
                ent["algorithmic_per_step"] = per_step_work
                ent["traffic"] = traffic.get(name)
            kernels[name] = ent
        dom = max((k for k in kernels if "bound" in kernels[k]), key=lambda k: kernels[k]["ms_per_step"])
        d = kernels[dom]
        roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"], "unit": d["unit"],
                    "frac": d["frac"], "traffic": d["traffic"], "peak_source": peaks["source"],
                    "note": "achieved = algorithmic bytes / flops (SURVEY 8(d): every gathered/scattered element once; dense layers 2 "
                            "FLOP per MAC) / CUDA-event time of the kind per step; traffic = ncu dram bytes per step from profiles/ "
                            "(null if no capture for this config)",
                    "whole_step": {"hbm_frac": work["total_bytes"] * n_local / (ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                   "tensor_frac": work["total_flop"] * n_local / (ms_step * 1e-3) / 1e12 / peaks["tflops_sustained"]}}
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (tcgen05)" if args.mode == "tc" else "f32",
            "data": "synthetic", "config": bench_config(args, "tc_f16" if args.mode == "tc" else "fp32_simt"),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(19 * n_local), "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps,
                    "loss_readback": "synchronous .item() every step" if args.e2e_sync_loss else
                                     "every step, through pinned memory, consumed one step later"},
            "with_optimizer": {"value": n_global / (ms_opt / args.steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_opt / args.steps,
                               "what": "nvp_b200.trainer.FusedTrainer.step: the same step + fused AdamW (exact torch.optim.AdamW "
                                       "semantics, gradient clear folded in)" + (f"; each rank updates the {reduce_view.numel()} replicated "
                                       f"parameters and its own 1/{world} of the sparse grid" if slab else
                                       f" over all {flat_params.numel()} parameters")},
            "unfused_api": None if ms_api is None else {
                "value": n_global / (ms_api / args.steps * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_api / args.steps,
                "what": "model(x) -> loss_functions.image_mse (torch) -> loss.backward(): the reference's own sequence through the "
                        "drop-in nn.Module and autograd (forward kernels + recomputing fused backward + torch's loss kernels)"},
            "gpu_launches": gpu_launches, "clocks": clocks, "roofline": roofline, "kernels": kernels,
            "loss_last_step": loss_last,
        }
        if world == 1 and not args.no_cpu_baseline:
            n_sample = N_SAMPLES // 4
            t = cpu_port_step_time(cfg_json, n_sample, 2, 1)
            out["cpu_baseline"] = {"value": n_sample / t / 1e6, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"{n_sample} of {N_SAMPLES} coordinates (1/4 batch), 2 timed steps after 1 warm-up; "
                                             "oracle/ CPU port of the reference arithmetic (torch CPU, all host threads)"}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="s", choices=["s", "l"])
    ap.add_argument("--mode", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--video", default="1080p", choices=sorted(VIDEOS), help="coordinate grid of the synthetic video (configs[4]: 4k)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = 1,245,184 samples per GPU; strong = the reference's 1,245,184 samples per step split over "
                         "the GPUs (same sample set as one GPU)")
    ap.add_argument("--e2e-sync-loss", action="store_true", help="end-to-end leg: read the loss with a blocking .item() every step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--full-allreduce", action="store_true", help="N>1: all-reduce the whole gradient (no t-slab ownership)")
    ap.add_argument("--overlap", action="store_true",
                    help="N>1: all-reduce the grid gradients under the wgrad kernel (dist.GridFirstAllReduce; validated on 2 GPUs "
                         "only) instead of one all-reduce after the step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
