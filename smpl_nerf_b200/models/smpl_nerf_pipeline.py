"""SmplNerfPipeline -- drop-in for models/smpl_nerf_pipeline.py:7-100."""
from .nerf_pipeline import NerfPipeline


class SmplNerfPipeline(NerfPipeline):
    """``SmplNerfPipeline(model_coarse, model_fine, model_warp_field, args, position_encoder,
    direction_encoder, human_pose_encoder)``.

    Returns ``(rgb, rgb_fine, warp, ray_samples, warped_samples, densities)`` for the last pass that ran
    (fine when ``args.run_fine`` else coarse), like the reference."""

    kind = 'smpl'

    def __init__(self, model_coarse, model_fine, model_warp_field, args, position_encoder, direction_encoder,
                 human_pose_encoder):
        super().__init__(model_coarse, model_fine, args, position_encoder, direction_encoder)
        self.human_pose_encoder = human_pose_encoder
        self.model_warp_field = model_warp_field
        self.args = args

    def forward(self, data):
        if len(data) < 6:
            raise ValueError('data must be [ray_samples, ray_translation, ray_direction, z_vals, goal_pose, rgb]')
        o = self._render(data)
        return o['rgb'], o['rgb_fine'], o['warp_out'], o['samples_out'], o['warped_out'], o['alpha_out']
