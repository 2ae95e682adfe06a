"""NerfPipeline -- drop-in for models/nerf_pipeline.py:7-67, computed by the fused sm_100a kernel."""
from .. import engine
from .singe_sample_pipeline import SmplPipeline


class NerfPipeline(SmplPipeline):
    """``NerfPipeline(model_coarse, model_fine, args, position_encoder, direction_encoder)``.

    ``forward(data)`` with ``data = [ray_samples, ray_translation, ray_direction, z_vals, rgb]`` returns
    ``(rgb, rgb_fine, ray_samples_fine, densities)``; with ``args.run_fine == 0`` it returns
    ``(rgb, rgb, ray_samples, densities)`` exactly like the reference (``densities`` is alpha).
    """

    kind = 'nerf'

    def __init__(self, model_coarse, model_fine, args, position_encoder, direction_encoder):
        super().__init__(model_coarse, args, position_encoder, direction_encoder)
        self.model_fine = model_fine

    #: set to True to read the kernel's range flag after every call (costs one device synchronisation per call)
    strict_range = False
    #: 0 = parity (fp16 hi/lo split, 3 MMA passes, every parity claim), 1 = fast (one fp16 pass; misses the 1e-4 alpha bar),
    #: 2 = exact (bf16 x 3 planes, 6 passes, layer by layer: exact fp32 operands, no activation-range limit; ~8x slower, inference only)
    precision = 0

    def _render(self, data, **kw):
        kw.setdefault('precision', self.precision)
        out = engine.render(self.kind, self.model_coarse, self.model_fine, getattr(self, 'model_warp_field', None),
                            self.args, self.position_encoder, self.direction_encoder,
                            getattr(self, 'human_pose_encoder', None), data, **kw)
        self._status = out['status']
        if self.strict_range:
            self.check_range()
        return out

    def check_range(self):
        """Raise if an activation of the LAST call left the fp16 range (|x| > 65504): the fp16 hi/lo operand split
        saturates there, so the result would silently deviate from the fp32 reference.  Synchronises the device."""
        st = getattr(self, '_status', None)
        if st is not None and int(st.item()) & 3:          # bit 0: fused kernel (activation or packed weight), bit 1: layer-by-layer path
            raise FloatingPointError('smpl_nerf_b200: a hidden activation exceeded the fp16 range (65504), which the fp16 hi/lo '
                                     'operand split cannot represent -- set pipeline.precision = 2 (exact mode: bf16 x 3 planes '
                                     'with fp32\'s exponent range) for this net')

    def forward(self, data):
        if len(data) < 5:
            raise ValueError('data must be [ray_samples, ray_translation, ray_direction, z_vals, rgb]')
        o = self._render(data)
        return o['rgb'], o['rgb_fine'], o['samples_out'], o['alpha_out']
