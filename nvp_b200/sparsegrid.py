"""3-D sparse positional feature grid: parameter container mirroring /root/reference/sparsegrid.py:4-21.

The lookup itself (nearest voxel + 3x3 neighbourhood, sparsegrid.py:23-72) runs inside the fused CUDA
path; this class owns `embeddings` [T,X,Y,F] and the attributes eval.py / compression.py read.
"""
import torch
from torch import nn


class SparseGrid(nn.Module):
    def __init__(self, level_dim=2, x_resolution=300, y_resolution=300, t_resolution=600, upsample=False):
        super().__init__()
        if upsample:
            raise NotImplementedError("SparseGrid(upsample=True) is disabled in both reference configs and not built")
        self.level_dim = level_dim
        self.x_resolution, self.y_resolution, self.t_resolution = x_resolution, y_resolution, t_resolution
        self.embeddings = nn.Parameter(torch.empty(t_resolution, x_resolution, y_resolution, level_dim))
        self.upsample = upsample
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)  # sparsegrid.py:19-21

    def forward(self, inputs: torch.Tensor) -> torch.Tensor:
        """[N,3] (t,x,y) -> [N, 9*level_dim]: nearest voxel + 3x3 neighbourhood (sparsegrid.py:23-72), differentiable
        w.r.t. `embeddings`.  Standalone use of the fused model's gather / scatter kernels (keyframe planes are dummies)."""
        from . import _lib, functional
        if not inputs.is_cuda:
            raise RuntimeError("SparseGrid input is a CPU tensor: nvp_b200 has no CPU path")
        desc = _lib.NvpDesc(1, 1, 2, 1.35, self.level_dim, self.t_resolution, self.x_resolution, self.y_resolution, 128, 3, 30.0)
        dummy_plane = torch.zeros(4, device=inputs.device, dtype=torch.float32)   # 1 level of 2x2 cells, 1 feature
        z = functional.LatentFunction.apply(desc, inputs.to(torch.float32).contiguous(), dummy_plane, dummy_plane,
                                            dummy_plane, self.embeddings)
        return z[:, 3:]
