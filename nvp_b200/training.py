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


def train(model, train_dataloader, epochs, lr, steps_til_summary, epochs_til_checkpoint, model_dir, loss_fn=None,
          summary_fn=None, fused=True, log=print):
    optim = torch.optim.AdamW(lr=lr, params=model.parameters(), weight_decay=0.001)
    scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optim, T_max=epochs, eta_min=1e-5)
    summaries_dir = os.path.join(model_dir, 'summaries')
    checkpoints_dir = os.path.join(model_dir, 'checkpoints')
    cond_mkdir(summaries_dir)
    cond_mkdir(checkpoints_dir)
    writer = _make_writer(summaries_dir)
    total_steps, best_psnr, psnr = 0, 0.0, 0.0
    train_losses = []
    model_input = gt = None
    for epoch in range(epochs):
        if not epoch % epochs_til_checkpoint and epoch:
            torch.save(model.state_dict(), os.path.join(checkpoints_dir, 'model_epoch_%04d.pth' % epoch))
            np.savetxt(os.path.join(checkpoints_dir, 'train_losses_epoch_%04d.txt' % epoch), np.array(train_losses))
        for step, (model_input, gt) in enumerate(train_dataloader):
            model_input = {k: v.cuda(non_blocking=True) for k, v in model_input.items()}
            gt = {k: v.cuda(non_blocking=True) for k, v in gt.items()}
            optim.zero_grad(set_to_none=False)
            if fused:
                n = gt['img'].numel() // 3
                train_loss = (model.fwd_loss_bwd(model_input, gt['img']) / (3.0 * n)).squeeze(0)
            else:
                gt['img'] = (gt['img'].float() - 127.5) / 127.5
                train_loss = loss_fn(model(model_input), gt)['img_loss'].mean()
                train_loss.backward()
            loss_value = float(train_loss.detach())          # the reference's train_loss.item()
            tmp_psnr = 10 * math.log10(4 / loss_value)
            if writer is not None:
                writer.add_scalar('img_loss', loss_value, total_steps)
                writer.add_scalar('img_loss_psnr', tmp_psnr, total_steps)
                writer.add_scalar('lr', float(scheduler.get_last_lr()[0]), total_steps)
                writer.add_scalar('total_train_loss', loss_value, total_steps)
            if tmp_psnr > best_psnr and not (total_steps + 1) % 200:
                torch.save({'epoch': total_steps, 'model': model.state_dict(), 'optimizer': optim.state_dict(),
                            'scheduler': scheduler.state_dict()}, os.path.join(checkpoints_dir, 'model_best.pth'))
                best_psnr = tmp_psnr
            optim.step()
            scheduler.step()
            train_losses.append(loss_value)
            if summary_fn is not None and not total_steps % steps_til_summary:
                psnr = summary_fn(model, model_input, gt, writer, total_steps)
                log("Epoch %d, Total loss %0.6f, psnr: %0.6f" % (epoch, loss_value, psnr))
            total_steps += 1
    torch.save({'epoch': total_steps, 'model': model.state_dict(), 'optimizer': optim.state_dict(),
                'scheduler': scheduler.state_dict()}, os.path.join(checkpoints_dir, 'model_final.pth'))
    if summary_fn is not None and model_input is not None:
        psnr = summary_fn(model, model_input, gt, writer, total_steps)
    if writer is not None:
        writer.close()
    np.savetxt(os.path.join(checkpoints_dir, 'train_losses_final.txt'), np.array(train_losses))
    return psnr, train_losses
