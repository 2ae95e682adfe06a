"""WarpFieldNet -- parameter container with the reference's names (models/warp_field_net.py:6-15):
Linear(positions_dim + pose_dim -> width) -> ReLU -> Linear(width -> 3).  Evaluated inside the fused
kernel by SmplNerfPipeline; the stand-alone ``forward`` (models/warp_field_net.py:17-21) runs the two layers
as tcgen05 GEMM launches of the library (smpl_nerf_b200/mlp.py): inference only, CUDA only."""
import torch.nn as nn


class WarpFieldNet(nn.Module):

    def __init__(self, n_layers=8, width=256, positions_dim=60, pose_dim=24):
        super().__init__()
        self.positions_dim = positions_dim
        self.direcions_dim = pose_dim            # (sic) the reference stores pose_dim under this name
        self.linear1 = nn.Linear(positions_dim + pose_dim, width)
        self.linear2 = nn.Linear(width, 3)

    def forward(self, x):
        from ..mlp import warp_field_net_forward
        return warp_field_net_forward(self, x)

    @property
    def is_cuda(self):
        return next(self.parameters()).is_cuda
