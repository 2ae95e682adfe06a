"""RenderRayNet -- parameter container with the reference's constructor, attribute and state_dict
names (models/render_ray_net.py:6-40), so checkpoints (`model_coarse.pt`, `model_fine.pt`) load
unchanged and the fused engine can read the hyper-parameters off the module.

The network is evaluated inside the fused kernel (csrc/nrf_fused.cu) when a pipeline is called; the
stand-alone ``forward`` on pre-encoded features (models/render_ray_net.py:42-61) runs every layer as one
tcgen05 GEMM launch of the library (smpl_nerf_b200/mlp.py): inference only, CUDA only."""
import torch.nn as nn


class RenderRayNet(nn.Module):

    def __init__(self, n_layers=8, width=256, positions_dim=60, directions_dim=24, additional_input_dim=0,
                 skips=(4,), use_directional_input=1):
        super().__init__()
        self.n_layers = n_layers
        self.width = width
        self.positions_dim = positions_dim
        self.direcions_dim = directions_dim      # (sic) the reference's attribute name
        self.skips = list(skips)
        self.additional_input_dim = additional_input_dim
        self.use_directional_input = use_directional_input
        first_in = positions_dim + additional_input_dim
        # creation order == the reference's, so the same torch.manual_seed gives the same init
        self.positions_pose_input = nn.Linear(first_in, width)
        self.positional_net = nn.ModuleList()
        for i in range(n_layers - 1):
            self.positional_net.append(nn.Linear(width + first_in if i in self.skips else width, width))
        self.additional_linear_layer = nn.Linear(width, width)
        self.sigma_out_layer = nn.Linear(width, 1)
        half = width // 2
        self.directional_input = nn.Linear(width + directions_dim if use_directional_input else width, half)
        self.directional_net = nn.ModuleList([nn.Linear(half, half)])
        self.rgb_out_layer = nn.Linear(half, 3)

    def forward(self, x):
        from ..mlp import render_ray_net_forward
        return render_ray_net_forward(self, x)

    @property
    def is_cuda(self):
        return next(self.parameters()).is_cuda
