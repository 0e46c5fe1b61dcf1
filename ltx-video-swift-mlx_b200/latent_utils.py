"""Host-side shape helpers mirroring Pipeline/VideoLatentShape.swift and the validation in Configuration/LTXConfig.swift.
(patchify / unpatchify themselves are device kernels inside libltxcuda; the host versions here only serve callers that
hold numpy latents, e.g. the Swift seam emulation in bench.py / tests.)"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np

from ._lib import LtxError


@dataclass
class VideoLatentShape:
    """Pipeline/VideoLatentShape.swift:35-41,72,95."""
    batch: int
    channels: int
    frames: int
    height: int
    width: int

    @staticmethod
    def from_pixel_dimensions(batch: int, channels: int, frames: int, height: int, width: int) -> "VideoLatentShape":
        return VideoLatentShape(batch, channels, (frames - 1) // 8 + 1, height // 32, width // 32)

    @property
    def token_count(self) -> int:
        return self.frames * self.height * self.width

    @property
    def shape(self) -> Tuple[int, int, int, int, int]:
        return (self.batch, self.channels, self.frames, self.height, self.width)

    @property
    def fhw(self) -> Tuple[int, int, int]:
        return (self.frames, self.height, self.width)


@dataclass
class LTXVideoGenerationConfig:
    """Subset of LTXVideoGenerationConfig (Configuration/LTXConfig.swift:216-362) that reaches the hot path."""
    width: int = 768
    height: int = 512
    num_frames: int = 25
    num_steps: int = 8
    cfg_scale: float = 1.0
    guidance_rescale: float = 0.0
    stg_scale: float = 0.0
    stg_blocks: List[int] = field(default_factory=lambda: [29])
    ge_gamma: float = 0.0
    seed: int = 0

    def validate(self):
        """:310-362 -- same rules, raised as LtxError(1) = LTXError.invalidConfiguration."""
        def bad(msg):
            raise LtxError(1, msg)
        if self.width % 32 or self.height % 32:
            bad("width and height must be divisible by 32")
        if self.width <= 0 or self.height <= 0 or self.width > 2048 or self.height > 2048:
            bad("width/height out of range")
        if (self.num_frames - 1) % 8 != 0:
            bad("num_frames must be 8n+1")
        if not (9 <= self.num_frames <= 257):
            bad("num_frames must be in [9, 257]")
        if not (1 <= self.num_steps <= 100):
            bad("num_steps must be in [1, 100]")
        if not (1.0 <= self.cfg_scale <= 20.0):
            bad("cfg_scale must be in [1, 20]")


def patchify(latent: np.ndarray) -> np.ndarray:
    """(B,C,F,H,W) -> (B, F*H*W, C)   Pipeline/LatentUtils.swift:20-36"""
    B, C, F, H, W = latent.shape
    return np.ascontiguousarray(latent.transpose(0, 2, 3, 4, 1).reshape(B, F * H * W, C))


def unpatchify(x: np.ndarray, shape: VideoLatentShape) -> np.ndarray:
    """(B,T,C) -> (B,C,F,H,W)   :38-54"""
    B, T, C = x.shape
    return np.ascontiguousarray(x.reshape(B, shape.frames, shape.height, shape.width, C).transpose(0, 4, 1, 2, 3))
