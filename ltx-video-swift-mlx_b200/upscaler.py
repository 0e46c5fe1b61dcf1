"""SpatialUpscaler + the latent glue between the two stages of generateVideoTwoStage -- host mirror of
Models/Upscaler/SpatialUpscaler.swift and Pipeline/LatentUtils.swift:201-227 over libltxcuda."""
from __future__ import annotations

import numpy as np

from .context import LtxContext


class SpatialUpscaler:
    """Handle on the upscaler weights held by an LtxContext (loadSpatialUpscaler, SpatialUpscaler.swift:262)."""

    def __init__(self, ctx: LtxContext):
        self.ctx = ctx


def upsample_latents(latent, upscaler: SpatialUpscaler) -> np.ndarray:
    """upsampleLatents(_:upscaler:latentMean:latentStd:) (:360-383): [1,128,F,H,W] -> [1,128,F,2H,2W]; the per-channel
    statistics are the ones loaded with the VAE decoder."""
    return upscaler.ctx.upscale_latent(latent)[None]


def adain_filter_latent(latent, reference, ctx: LtxContext, factor: float = 1.0) -> np.ndarray:
    """adainFilterLatent(_:reference:factor:) (Pipeline/LatentUtils.swift:201-227) on [1,C,F,H,W] arrays."""
    lat = np.asarray(latent, dtype=np.float32)
    ref = np.asarray(reference, dtype=np.float32)
    return ctx.adain_filter(lat[0], ref[0], factor)[None]
