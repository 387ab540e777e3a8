"""Multi-resolution dense 2-D keyframe grid: the `tcnn.Encoding(n_input_dims=2, DenseGrid)` surface.

Mirrors the operator boundary the reference uses (modules.py:14-23): ctor `(n_input_dims, encoding_config)`,
attributes `.params` (flat fp32 nn.Parameter, level-major / cell / feature-minor, no padding — the layout
eval.py:28-35 and compression.py:72,77 rely on), `.dtype`, `.n_output_dims`.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib


class Encoding(nn.Module):
    def __init__(self, n_input_dims, encoding_config, seed=1337, dtype=torch.float32):
        super().__init__()
        if n_input_dims != 2 or encoding_config.get("otype", "DenseGrid") != "DenseGrid":
            raise NotImplementedError("only the 2-D DenseGrid encoding used by NVP is built")
        if dtype != torch.float32:
            raise NotImplementedError("the reference fork runs fp32 parameters (modules.py:15)")
        self.n_input_dims = 2
        self.encoding_config = dict(encoding_config)
        self.n_levels = encoding_config["n_levels"]
        self.n_features = encoding_config["n_features_per_level"]
        self.n_output_dims = self.n_levels * self.n_features
        self.dtype = torch.float32
        self.seed = seed
        desc = _lib.NvpDesc(self.n_features, self.n_levels, encoding_config["base_resolution"],
                            encoding_config["per_level_scale"], 1, 1, 1, 1, 128, 3, 30.0)
        self.level_scales, self.level_res, self.level_offsets = _lib.level_table(desc)
        self._op_desc = desc   # xy plane + 1x1x1 dummy 3-D grid with one feature: latent = [this plane | 2 dummies | 9]
        n_params = self.level_offsets[-1] * self.n_features
        # tcnn initialises U(-1e-4, 1e-4) from its own pcg32 (seed 1337, identical for the three planes) and
        # does not touch torch's global RNG; a private generator keeps both properties.
        g = torch.Generator().manual_seed(seed)
        self.params = nn.Parameter((torch.rand(n_params, generator=g) * 2 - 1) * 1e-4)

    # ------------------------------------------------------------------------------------------
    def padded_level_offsets(self):
        """Level offsets (in cells) of upstream tiny-cuda-nn's layout, which pads every level to a multiple of 8 entries.
        The reference's fork does not pad (compression.py:72,77 would fail otherwise); checkpoints written with an upstream
        build do."""
        off, out = 0, [0]
        for res in self.level_res:
            off += (res * res + 7) // 8 * 8
            out.append(off)
        return out

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        """Accept a `params` tensor in upstream tcnn's padded layout: the per-level padding entries are dropped (they are
        never addressed by this implementation, whose flat index wraps modulo res^2 - SURVEY A.2); any other size
        mismatch gets a message that names both layouts.  PARITY UNPINNED for that remap: no tcnn build is available to
        check an upstream checkpoint against."""
        key = prefix + "params"
        t = state_dict.get(key)
        if t is not None and t.numel() != self.params.numel():
            padded = self.padded_level_offsets()
            if t.numel() == padded[-1] * self.n_features:
                flat = t.reshape(-1, self.n_features)
                state_dict[key] = torch.cat([flat[padded[l]: padded[l] + self.level_res[l] ** 2]
                                             for l in range(self.n_levels)]).reshape(-1).to(t.dtype)
            else:
                raise RuntimeError(
                    f"{key}: got {t.numel()} values; expected {self.params.numel()} (level-major res^2*F entries, no padding: "
                    f"the reference fork's layout) or {padded[-1] * self.n_features} (upstream tiny-cuda-nn, levels padded "
                    f"to multiples of 8 entries)")
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """[N,2] in [0,1] -> [N, n_levels*F] (level-major), differentiable w.r.t. `.params` — the standalone
        tcnn.Encoding.__call__ of modules.py:65-67.  Runs the same gather / scatter kernels as the fused model:
        the inputs are presented as the xy plane of a latent whose other planes and 3-D grid are dummies."""
        from . import functional
        if not x.is_cuda:
            raise RuntimeError("Encoding input is a CPU tensor: nvp_b200 has no CPU path")
        n = x.shape[0]
        coords = torch.cat((torch.zeros(n, 1, device=x.device, dtype=torch.float32), x.to(torch.float32)), dim=1)
        dummy_plane = self.params.detach()
        dummy_grid = torch.zeros(1, 1, 1, 1, device=x.device, dtype=torch.float32)
        z = functional.LatentFunction.apply(self._op_desc, coords, self.params, dummy_plane, dummy_plane, dummy_grid)
        return z[:, : self.n_output_dims]

    def extra_repr(self):
        return f"DenseGrid levels={self.n_levels} F={self.n_features} params={self.params.numel()}"
