"""`modules.NVP` drop-in: same constructor, attributes, state_dict keys and forward contract as
/root/reference/modules.py:8-84, with the whole per-coordinate path (3 keyframe planes + sparse grid
gather, modulator, SIREN, and their backward) executed by libnvp_b200.so on a B200.

    model = NVP(type='nvp', out_features=3, encoding_config=config["nvp"]).cuda()
    out = model({'all_coords': [b,t,3], 'temporal_steps': [b,t]})['model_out']      # [b,t,3]

Extra (not in the reference): `mode=` kwarg / NVP_B200_MODE env ("tc" = tcgen05 fp16 tensor cores,
"fp32" = CUDA-core fp32), and `fwd_loss_bwd()` = the fused training step used by nvp_b200.training.
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
from torch import nn

from . import _lib, functional, modulation
from .encoding import Encoding
from .sparsegrid import SparseGrid

DEFAULT_MODE = os.environ.get("NVP_B200_MODE", "tc")


class NVP(nn.Module):
    def __init__(self, out_features=3, encoding_config=None, mode: Optional[str] = None, **kwargs):
        # **kwargs swallows the junk the reference's callers pass (type='nvp', in_features=2):
        # train_video.py:46, eval.py:148.
        super().__init__()
        if out_features != 3:
            raise NotImplementedError("the fused path is built for RGB output (out_features=3)")
        self.encoding_config = encoding_config
        # model(x) -> loss.backward(): gradients are added straight into .grad (functional.NvpFunction); set to False when
        # they must be returned instead (torch.autograd.grad, tensor hooks)
        self.direct_grad_accumulation = True
        self.desc = _lib.make_desc(encoding_config)
        self.mode_name = mode or DEFAULT_MODE
        if self.mode_name not in _lib.MODES:
            raise ValueError(f"mode must be one of {sorted(_lib.MODES)}")

        # same construction order as modules.py:13-47 so torch's RNG stream lines up with the reference
        self.keyframes_xy = Encoding(n_input_dims=2, encoding_config=encoding_config["2d_encoding_xy"])
        assert self.keyframes_xy.dtype == torch.float32
        self.keyframes_yt = Encoding(n_input_dims=2, encoding_config=encoding_config["2d_encoding_yt"])
        assert self.keyframes_yt.dtype == torch.float32
        self.keyframes_xt = Encoding(n_input_dims=2, encoding_config=encoding_config["2d_encoding_xt"])
        assert self.keyframes_xt.dtype == torch.float32

        e3 = encoding_config["3d_encoding"]
        self.sparse_grid = SparseGrid(level_dim=e3["n_features_per_level"], x_resolution=e3["x_resolution"],
                                      y_resolution=e3["y_resolution"], t_resolution=e3["t_resolution"],
                                      upsample=e3["upsample"])
        self.net = modulation.SirenNet(dim_in=1, dim_hidden=encoding_config["network"]["n_neurons"],
                                       dim_out=out_features, num_layers=encoding_config["network"]["n_hidden_layers"],
                                       w0_initial=30.0)
        latent_dim = sum(encoding_config[k]["n_levels"] * encoding_config[k]["n_features_per_level"]
                         for k in ("2d_encoding_xy", "2d_encoding_yt", "2d_encoding_xt"))
        latent_dim += e3["n_features_per_level"] * 9
        self.latent_dim = latent_dim
        self.wrapper = modulation.SirenWrapper(self.net, latent_dim=latent_dim)

    # ------------------------------------------------------------------------------------------
    @property
    def mode(self) -> int:
        return _lib.MODES[self.mode_name]

    def hot_path_parameters(self) -> List[nn.Parameter]:
        """The 18 parameter tensors in functional.PARAM_ORDER."""
        m = self.wrapper.modulator.layers
        s = self.net.layers
        return [self.keyframes_xy.params, self.keyframes_yt.params, self.keyframes_xt.params,
                self.sparse_grid.embeddings,
                s[0].weight, s[1].weight, s[2].weight, s[0].bias, s[1].bias, s[2].bias,
                self.net.last_layer.weight, self.net.last_layer.bias,
                m[0][0].weight, m[1][0].weight, m[2][0].weight, m[0][0].bias, m[1][0].bias, m[2][0].bias]

    def forward(self, model_input, temporal_interp=False, params=None):
        timesteps = model_input["temporal_steps"]
        b, t = timesteps.size(0), timesteps.size(1)
        tsteps = timesteps.reshape(b * t)
        coords = model_input["all_coords"].reshape(-1, 3)  # t, x, y
        if temporal_interp:
            # eval-only path (eval.py:239 with --t_interp): SparseGrid.forward_inter, forward kernel only
            with torch.no_grad():
                out = functional.forward(self.desc, self.hot_path_parameters(), coords, tsteps,
                                         self.mode | _lib.FLAG_TEMPORAL_INTERP)
            return {"model_out": out.reshape(b, t, 3)}
        out = functional.NvpFunction.apply(self.desc, self.mode, self.direct_grad_accumulation and torch.is_grad_enabled(),
                                           coords, tsteps, *self.hot_path_parameters())
        return {"model_out": out.reshape(b, t, 3)}

    def encode(self, all_coords: torch.Tensor) -> torch.Tensor:
        """Positional feature vector z [N, latent_dim] (modules.py:61-78)."""
        return functional.encode_latent(self.desc, self.hot_path_parameters(), all_coords.reshape(-1, 3))

    def fwd_loss_bwd(self, model_input, gt_u8: torch.Tensor, n_global: Optional[int] = None,
                     loss_sum: Optional[torch.Tensor] = None, out_rgb: Optional[torch.Tensor] = None,
                     grid_event: Optional[torch.cuda.Event] = None) -> torch.Tensor:
        """Fused training step (training.py:47-52,74): accumulates d(image_mse)/d(params) into `.grad`
        (allocated zero-filled when None) and returns sum((rgb-gt)^2) as a 1-element device tensor
        (divide by 3*n_global for the loss).  gt_u8 is the raw uint8 `img` from the sampler."""
        coords = model_input["all_coords"].reshape(-1, 3)
        tsteps = model_input["temporal_steps"].reshape(-1)
        gt = gt_u8.reshape(-1, 3)
        n = coords.shape[0]
        ps = self.hot_path_parameters()
        grads = []
        for p in ps:
            if not p.requires_grad:
                grads.append(None)
                continue
            g = p.grad
            if g is None or g.dtype != p.dtype or g.device != p.device or not g.is_contiguous() or g.shape != p.shape:
                # the library accumulates through raw pointers: an incompatible .grad is replaced (keeping its value)
                new = torch.zeros_like(p, memory_format=torch.contiguous_format)
                if g is not None:
                    new.copy_(g)
                p.grad = g = new
            grads.append(g)
        if loss_sum is None:
            loss_sum = torch.zeros(1, dtype=torch.float32, device=coords.device)
        functional.fwd_loss_bwd(self.desc, ps, grads, coords, tsteps, gt, n_global or n, loss_sum, self.mode, out_rgb,
                                grid_event)
        return loss_sum
