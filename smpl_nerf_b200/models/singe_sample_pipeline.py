"""Base class of the pipelines -- mirrors the constructor contract of
models/singe_sample_pipeline.py:8-15 (attributes ``device``, ``args``, ``model_coarse``,
``position_encoder``, ``direction_encoder``; nets registered as sub-modules).  The reference's
single-sample ``forward`` is not on the hot path and is not provided."""
import torch
from torch import nn


class SmplPipeline(nn.Module):

    def __init__(self, model_coarse, args, position_encoder, direction_encoder):
        super().__init__()
        self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.args = args
        self.model_coarse = model_coarse
        self.position_encoder = position_encoder
        self.direction_encoder = direction_encoder

    def forward(self, data):
        raise NotImplementedError("the single-sample pipeline is outside the accelerated hot path")
