"""Training loop — host code kept from the reference (training.py:11-100) with the device work of each step
(gt normalise, forward, image_mse, backward: training.py:47-52,74) replaced by ONE fused call when `fused=True`.

Same optimiser (AdamW lr, wd 1e-3), cosine schedule (eta_min 1e-5), PSNR definition (peak^2 = 4), checkpoint
names and cadence.  TensorBoard scalars are written when `torch.utils.tensorboard` is importable; the loss is
read back once per step (the reference's `.item()`), nothing else synchronises.
"""
import math
import os

import numpy as np
import torch


def cond_mkdir(path):
    if not os.path.exists(path):
        os.makedirs(path)


def _make_writer(path):
    try:
        from torch.utils.tensorboard import SummaryWriter
        return SummaryWriter(path)
    except Exception:  # noqa: BLE001 - tensorboard is optional host tooling
        return None


def _device_batches(train_dataloader, epochs, device, device_sampler, trainer):
    """Yields (model_input, gt) on the device, one per step.  device_sampler=True: the video lives in HBM and a kernel draws
    the batch (dataio.DeviceSampler; on several GPUs every rank draws its own frames only); otherwise the reference's host
    sampler feeds a double-buffered host->device prefetcher (replaces the per-step .cuda() copies of training.py:45-46)."""
    from . import dataio
    if device_sampler:
        wrapper = train_dataloader.dataset
        video = wrapper.data.view(wrapper.dataset.nframes, *wrapper.sidelength, wrapper.dataset.channels)
        t_range, n = None, wrapper.N_samples
        if trainer is not None and trainer.distributed and wrapper.dataset.nframes == trainer.t_resolution:
            t_range = trainer.slab                            # stratified: this rank's frames, 1/G of the batch
            n = wrapper.N_samples // trainer.world
        sampler = dataio.DeviceSampler(video, n_samples=n, device=device, seed=torch.initial_seed() % (1 << 31), t_range=t_range)
        for step in range(epochs * len(train_dataloader)):
            yield sampler.sample(step) + (t_range is not None,)
        return

    def host_batches():
        for _ in range(epochs):
            for model_input, gt in train_dataloader:
                yield (model_input["all_coords"], model_input["temporal_steps"], gt["img"])

    pf = dataio.DevicePrefetcher(host_batches(), device=device)
    for c, t, g in pf:
        yield {"all_coords": c, "temporal_steps": t}, {"img": g}, False
        pf.release()


def train(model, train_dataloader, epochs, lr, steps_til_summary, epochs_til_checkpoint, model_dir, loss_fn=None,
          summary_fn=None, fused=True, log=print, distributed=None, fused_optimizer=True, device_sampler=False):
    """training.py:11-100.  fused=True (default) runs every step through nvp_b200.trainer.FusedTrainer: one fused
    forward+loss+backward call, the fused AdamW + cosine schedule (fused_optimizer=False: torch.optim.AdamW on the same
    flat buffers), prefetched or device-side sampling, and -- when torch.distributed is initialised with more than one
    rank (distributed=None/True) -- the t-slab data-parallel scheme.  Checkpoints keep the reference's names, cadence and
    dict layout; with several ranks the owned slabs are gathered first and rank 0 writes.  fused=False is the reference's
    own sequence (model(x) -> loss_fn -> backward -> torch AdamW) on one GPU."""
    import torch.distributed as dist
    if fused:
        from .trainer import FusedTrainer
        trainer = FusedTrainer(model, lr, epochs, distributed=distributed, fused_optimizer=fused_optimizer)
        optim = scheduler = None
    else:
        trainer = None
        optim = torch.optim.AdamW(lr=lr, params=model.parameters(), weight_decay=0.001)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optim, T_max=epochs, eta_min=1e-5)
    rank0 = not (dist.is_available() and dist.is_initialized()) or dist.get_rank() == 0
    summaries_dir = os.path.join(model_dir, 'summaries')
    checkpoints_dir = os.path.join(model_dir, 'checkpoints')
    if rank0:
        cond_mkdir(summaries_dir)
        cond_mkdir(checkpoints_dir)
    writer = _make_writer(summaries_dir) if rank0 else None
    device = next(model.parameters()).device

    def model_state():
        return trainer.model_state_dict() if trainer is not None else model.state_dict()

    def full_checkpoint(step):
        sd = model_state()                                   # collective with several ranks: every rank calls it
        if not rank0:
            return None
        extra = trainer.optimizer_state_dict() if trainer is not None else {'optimizer': optim.state_dict(), 'scheduler': scheduler.state_dict()}
        return {'epoch': step, 'model': sd, **extra}

    total_steps, best_psnr, psnr = 0, 0.0, 0.0
    train_losses = []
    model_input = gt = None
    steps_per_epoch = len(train_dataloader)
    for model_input, gt, routed in _device_batches(train_dataloader, epochs, device, device_sampler, trainer):
        epoch = total_steps // steps_per_epoch
        if not total_steps % steps_per_epoch and not epoch % epochs_til_checkpoint and epoch:
            sd = model_state()
            if rank0:
                torch.save(sd, os.path.join(checkpoints_dir, 'model_epoch_%04d.pth' % epoch))
                np.savetxt(os.path.join(checkpoints_dir, 'train_losses_epoch_%04d.txt' % epoch), np.array(train_losses))
        if trainer is not None:
            n = gt['img'].numel() // 3
            n_global = n * trainer.world if routed else n
            train_loss = trainer.step(model_input, gt['img'], n_global=n_global, routed=routed).squeeze(0)
            lr_now = trainer.current_lr()
        else:
            optim.zero_grad(set_to_none=False)
            gt = dict(gt)
            gt['img'] = (gt['img'].float() - 127.5) / 127.5
            train_loss = loss_fn(model(model_input), gt)['img_loss'].mean()
            train_loss.backward()
            lr_now = float(scheduler.get_last_lr()[0])
        loss_value = float(train_loss.detach())          # the reference's train_loss.item()
        tmp_psnr = 10 * math.log10(4 / loss_value)
        if writer is not None:
            writer.add_scalar('img_loss', loss_value, total_steps)
            writer.add_scalar('img_loss_psnr', tmp_psnr, total_steps)
            writer.add_scalar('lr', lr_now, total_steps)
            writer.add_scalar('total_train_loss', loss_value, total_steps)
        if tmp_psnr > best_psnr and not (total_steps + 1) % 200:
            # (the reference saves before this step's update, training.py:64-76; with the update fused into the step the
            # checkpoint holds the parameters after it)
            ck = full_checkpoint(total_steps)
            if rank0:
                torch.save(ck, os.path.join(checkpoints_dir, 'model_best.pth'))
            best_psnr = tmp_psnr
        if trainer is None:
            optim.step()
            scheduler.step()
        train_losses.append(loss_value)
        if summary_fn is not None and not total_steps % steps_til_summary:
            if trainer is not None:
                trainer.sync_slabs()
            psnr = summary_fn(model, model_input, gt, writer, total_steps)
            log("Epoch %d, Total loss %0.6f, psnr: %0.6f" % (epoch, loss_value, psnr))
        total_steps += 1
    ck = full_checkpoint(total_steps)
    if rank0:
        torch.save(ck, os.path.join(checkpoints_dir, 'model_final.pth'))
    if summary_fn is not None and model_input is not None:
        psnr = summary_fn(model, model_input, gt, writer, total_steps)
    if writer is not None:
        writer.close()
    if rank0:
        np.savetxt(os.path.join(checkpoints_dir, 'train_losses_final.txt'), np.array(train_losses))
    return psnr, train_losses
