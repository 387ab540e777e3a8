"""Parameter containers of the modulated SIREN, mirroring the reference's module tree.

Mirrors /root/reference/modulation.py (Siren :30-56, SirenNet :60-92, Modulator :96-121,
SirenWrapper :124-145): same attribute names, state_dict keys and initialisation (same torch RNG
calls in the same order, so a given torch.manual_seed yields the reference's initial weights).
These classes hold parameters only; the arithmetic runs in the fused CUDA path (functional.py).
"""
from __future__ import annotations

import math

import torch
from torch import nn


class Siren(nn.Module):
    """One SIREN layer's parameters (modulation.py:30-51)."""

    def __init__(self, dim_in, dim_out, w0=1.0, c=6.0, is_first=False):
        super().__init__()
        self.dim_in, self.is_first, self.w0 = dim_in, is_first, w0
        weight = torch.zeros(dim_out, dim_in)
        bias = torch.zeros(dim_out)
        w_std = (1 / dim_in) if is_first else (math.sqrt(c / dim_in) / w0)
        weight.uniform_(-w_std, w_std)
        bias.uniform_(-w_std, w_std)
        self.weight = nn.Parameter(weight)
        self.bias = nn.Parameter(bias)


class SirenNet(nn.Module):
    """modulation.py:60-81."""

    def __init__(self, dim_in, dim_hidden, dim_out, num_layers, w0=1.0, w0_initial=30.0):
        super().__init__()
        self.num_layers, self.dim_hidden = num_layers, dim_hidden
        self.w0_initial = w0_initial
        self.layers = nn.ModuleList([])
        for ind in range(num_layers):
            first = ind == 0
            self.layers.append(Siren(dim_in if first else dim_hidden, dim_hidden,
                                     w0=w0_initial if first else w0, is_first=first))
        self.last_layer = Siren(dim_hidden, dim_out, w0=w0)


class Modulator(nn.Module):
    """modulation.py:96-110 (nn.Linear default init, then kaiming-normal weights :151-154)."""

    def __init__(self, dim_in, dim_hidden, num_layers):
        super().__init__()
        self.layers = nn.ModuleList([])
        for ind in range(num_layers):
            dim = dim_in if ind == 0 else (dim_hidden + dim_in)
            self.layers.append(nn.Sequential(nn.Linear(dim, dim_hidden), nn.LeakyReLU()))

        def init_weights_normal(m):
            if type(m) == nn.Linear:
                nn.init.kaiming_normal_(m.weight, a=0.0, nonlinearity="relu", mode="fan_in")

        self.layers.apply(init_weights_normal)


class SirenWrapper(nn.Module):
    """modulation.py:124-138.  `net` is the same object the model also registers as `.net`."""

    def __init__(self, net, latent_dim=None):
        super().__init__()
        self.net = net
        self.modulator = None
        if latent_dim is not None:
            self.modulator = Modulator(dim_in=latent_dim, dim_hidden=net.dim_hidden, num_layers=net.num_layers)
