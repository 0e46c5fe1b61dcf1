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
