"""train_video.py — the reference CLI (experiment_scripts/train_video.py:22-110) on the B200-native path.

Same flags (--config, --logging_root, --experiment_name, --lr, --num_epochs, --epochs_til_ckpt,
--steps_til_summary, --dataset, --num_frames); stdlib argparse instead of configargparse (not installed),
and an existing log directory is replaced without the interactive prompt (train_video.py:57-62) when
--overwrite is given.  Extra: --mode tc|fp32, --unfused (drive the model through autograd like the reference).
"""
import argparse
import json
import os
import shutil
import sys

sys.path.append(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from functools import partial  # noqa: E402

from torch.utils.data import DataLoader  # noqa: E402

from nvp_b200 import dataio, loss_functions, modules, training  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument('-c', '--config', required=True)
    p.add_argument('--logging_root', type=str, default='./logs_nvp')
    p.add_argument('--experiment_name', type=str, default="")
    p.add_argument('--lr', type=float, default=1e-2)
    p.add_argument('--num_epochs', type=int, default=100000)
    p.add_argument('--epochs_til_ckpt', type=int, default=25000)
    p.add_argument('--steps_til_summary', type=int, default=1000)
    p.add_argument('--dataset', type=str, required=True)
    p.add_argument('--num_frames', type=int, default=600)
    p.add_argument('--mode', default=None, choices=[None, 'tc', 'fp32'])
    p.add_argument('--unfused', action='store_true')
    p.add_argument('--overwrite', action='store_true')
    opt = p.parse_args()

    with open(opt.config, 'r') as f:
        config = json.load(f)
    model = modules.NVP(type='nvp', out_features=3, encoding_config=config["nvp"], mode=opt.mode)
    model.cuda()
    vid_dataset = dataio.VideoTime(opt.dataset, split_num=opt.num_frames)
    coord_dataset = dataio.VideoTimeWrapper(vid_dataset, sidelength=vid_dataset.shape)
    dataloader = DataLoader(coord_dataset, shuffle=True, batch_size=1, pin_memory=True, num_workers=0)

    n_params = sum(p.numel() for p in model.parameters())
    n_pix = vid_dataset.vid.shape[0] * vid_dataset.vid.shape[1] * vid_dataset.vid.shape[2]
    root_path = os.path.join(opt.logging_root, opt.experiment_name)
    if os.path.exists(root_path):
        if not opt.overwrite:
            raise FileExistsError("The model directory %s exists (pass --overwrite)" % root_path)
        shutil.rmtree(root_path)
    os.makedirs(os.path.join(root_path, 'results'), exist_ok=True)
    with open(os.path.join(root_path, "config.json"), "w") as f:
        json.dump(config, f, indent=4)
    n_mlp = sum(p.numel() for p in model.wrapper.parameters())
    with open(os.path.join(root_path, 'results', 'results.txt'), 'w') as f:
        f.write(f" - video shape [f, w, h, c]: {vid_dataset.vid.shape}\n - cur bpp: {n_params * 32 / n_pix}\n"
                f" - quantized bpp: {(n_params * 8 + n_mlp * 24) / n_pix}\n - epochs: {opt.num_epochs}\n - learning rate: {opt.lr}\n")
    psnr, _ = training.train(model=model, train_dataloader=dataloader, epochs=opt.num_epochs, lr=opt.lr,
                             steps_til_summary=opt.steps_til_summary, epochs_til_checkpoint=opt.epochs_til_ckpt,
                             model_dir=root_path, loss_fn=partial(loss_functions.image_mse, None), summary_fn=None,
                             fused=not opt.unfused)
    print("final psnr (last summary):", psnr)


if __name__ == "__main__":
    main()
