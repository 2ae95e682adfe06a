"""Drop-in pipeline classes (same constructor and ``forward(data)`` contracts as the reference's
``models/*_pipeline.py``)."""
from .nerf_pipeline import NerfPipeline  # noqa: F401
from .smpl_nerf_pipeline import SmplNerfPipeline  # noqa: F401
from .append_to_nerf_pipeline import AppendToNerfPipeline  # noqa: F401
from .append_smpl_params_pipeline import AppendSmplParamsPipeline  # noqa: F401
from .render_ray_net import RenderRayNet  # noqa: F401
from .warp_field_net import WarpFieldNet  # noqa: F401
