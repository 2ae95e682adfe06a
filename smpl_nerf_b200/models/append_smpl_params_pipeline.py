"""AppendSmplParamsPipeline -- drop-in for models/append_smpl_params_pipeline.py:7-91 (the paper's model and
inference.py's default --inf_model_type, inference.py:227)."""
from .append_to_nerf_pipeline import AppendToNerfPipeline


class AppendSmplParamsPipeline(AppendToNerfPipeline):
    """``AppendSmplParamsPipeline(model_coarse, model_fine, args, position_encoder, direction_encoder,
    human_pose_encoder)``: all 69 SMPL pose parameters (positionally encoded: 1380 features when
    ``args.human_pose_encoding``) are prepended to the MLP input.  They are constant along a ray, so the engine
    folds them into per-ray bias vectors of the first and the skip layer (one SGEMM per net, ``nrf_ray_bias``)
    and runs the same fused kernel as ``AppendToNerfPipeline``.  Returns
    ``(rgb, rgb_fine, ray_samples_fine, densities)`` like the reference."""

    kind = 'append_full'
