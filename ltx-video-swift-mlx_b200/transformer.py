"""LTXTransformer -- host mirror of Models/Transformer/LTXTransformer.swift over libltxcuda (same call signature, same
runtime flags).  This is what the SwiftPM adapter does on the Swift side."""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np

from .context import LtxContext, make_flags


class LTXTransformer:
    def __init__(self, ctx: LtxContext):
        self.ctx = ctx
        self.config = ctx.config
        self._stg_blocks: Sequence[int] = ()
        self._skip_self_attn = False
        self._skip_ff = False
        self._cas_blocks: Sequence[int] = ()
        self._cas = 1.0

    # setSTGSkipFlags / clearSTGSkipFlags / setCrossAttentionScale (LTXTransformer.swift:497-526)
    def set_stg_skip_flags(self, skip_self_attention: bool = True, skip_feed_forward: bool = False,
                           block_indices: Sequence[int] = ()):
        self._stg_blocks, self._skip_self_attn, self._skip_ff = tuple(block_indices), skip_self_attention, skip_feed_forward

    def clear_stg_skip_flags(self):
        self._stg_blocks, self._skip_self_attn, self._skip_ff = (), False, False

    def set_cross_attention_scale(self, scale: float, for_blocks: Sequence[int]):
        self._cas, self._cas_blocks = float(scale), tuple(for_blocks)

    def clear_rope_cache(self):            # :202
        self.ctx.clear_caches()

    def __call__(self, latent, context, timesteps, context_mask=None, latent_shape: Tuple[int, int, int] = (1, 1, 1),
                 context_key: int = 0) -> np.ndarray:
        """callAsFunction(latent:context:timesteps:contextMask:latentShape:) (:235-241) -> velocity [B,N,C] fp32.
        `context_key` (extension): non-zero = the text embedding is step-invariant, cache its projections under this key."""
        flags = make_flags(self._stg_blocks, self._skip_self_attn, self._skip_ff, self._cas_blocks, self._cas, context_key)
        return self.ctx.dit_forward(latent, context, timesteps, context_mask, tuple(latent_shape), flags)


class LTX2Transformer:
    """Host mirror of LTX2Transformer (Models/Transformer/LTX2Transformer.swift): the dual audio / video model.  The call takes
    the Swift argument list (videoLatent, audioLatent, videoContext, audioContext, videoTimesteps, audioTimesteps, masks,
    videoLatentShape, audioNumFrames) and returns (video, audio) velocities."""

    def __init__(self, ctx):
        self.ctx = ctx

    def __call__(self, video_latent, audio_latent, video_context, audio_context, video_timesteps, audio_timesteps,
                 video_context_mask=None, audio_context_mask=None, video_latent_shape=None, audio_num_frames=None,
                 context_key: int = 0):
        import numpy as np
        vt, at = np.asarray(video_timesteps, dtype=np.float32).reshape(-1), np.asarray(audio_timesteps, dtype=np.float32).reshape(-1)
        n_video = int(np.asarray(video_latent.shape)[1])
        if vt.size not in (1, n_video) or at.size != 1:
            # videoTimesteps is [B] or [B, N] (image-to-video, Pipeline/LTXPipeline.swift:1293-1298); audioTimesteps is [B]
            raise ValueError("videoTimesteps must hold one sigma or one per video token, audioTimesteps one sigma")
        if audio_num_frames is not None and audio_num_frames != np.asarray(audio_latent.shape)[1]:
            raise ValueError("audio_num_frames must equal the audio latent length")
        return self.ctx.av_forward(video_latent, audio_latent, video_context, audio_context,
                                   float(vt[0]) if vt.size == 1 else vt.reshape(1, -1), float(at[0]),
                                   tuple(video_latent_shape), video_context_mask, audio_context_mask, context_key)
