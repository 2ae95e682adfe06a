"""AppendToNerfPipeline -- drop-in for models/append_to_nerf_pipeline.py:7-90."""
from .nerf_pipeline import NerfPipeline


class AppendToNerfPipeline(NerfPipeline):
    """``AppendToNerfPipeline(model_coarse, model_fine, args, position_encoder, direction_encoder,
    human_pose_encoder)``; ``data`` carries ``goal_pose[B,69]`` before ``rgb``.  The two pose angles
    (columns 38 and 41), encoded or raw per ``args.human_pose_encoding``, are prepended to the MLP
    input; being constant along a ray they become a per-ray bias inside the kernel."""

    kind = 'append'

    def __init__(self, model_coarse, model_fine, args, position_encoder, direction_encoder, human_pose_encoder):
        super().__init__(model_coarse, model_fine, args, position_encoder, direction_encoder)
        self.human_pose_encoder = human_pose_encoder

    def forward(self, data):
        if len(data) < 6:
            raise ValueError('data must be [ray_samples, ray_translation, ray_direction, z_vals, goal_pose, rgb]')
        o = self._render(data)
        return o['rgb'], o['rgb_fine'], o['samples_out'], o['alpha_out']
