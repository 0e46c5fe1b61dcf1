"""Shared helpers for the parity tests (tests may import the oracle; the product package may not)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ltx_oracle as O  # noqa: E402


def product():
    import ltx_video_swift_mlx_b200  # noqa: F401
    from ltx_video_swift_mlx_b200 import context
    return context


def small_dit_config(layers=2, heads=2, caption=192):
    ocfg = O.DiTConfig(num_layers=layers, num_heads=heads, head_dim=128, caption_channels=caption)
    ctxmod = product()
    pcfg = ctxmod.LTXTransformerConfig(num_layers=layers, num_attention_heads=heads, attention_head_dim=128,
                                       caption_channels=caption)
    return ocfg, pcfg


def make_ctx_with_dit(ocfg, pcfg, seed=1):
    ctxmod = product()
    w = O.make_dit_weights(ocfg, seed)
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights(w)
    ctx.finalize_weights()
    return ctx, w


def small_vae_config(base=512, blocks=1):
    ocfg = O.VAEConfig(base_channels=base, blocks_per_stage=blocks)
    ctxmod = product()
    pcfg = ctxmod.LTXTransformerConfig(num_layers=1, num_attention_heads=1, vae_base_channels=base,
                                       vae_blocks_per_stage=blocks)
    return ocfg, pcfg


def make_ctx_with_vae(ocfg, pcfg, seed=2, bf16_weights=True):
    ctxmod = product()
    w = O.make_vae_weights(ocfg, seed)
    if bf16_weights:  # conv kernels are stored as bf16 on the device; give the oracle the same values
        w = {k: (O.bf16_round(v) if (k.endswith(".weight") and v.ndim >= 2) else v) for k, v in w.items()}
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights(w, prefix="vae.")
    ctx.finalize_weights()
    return ctx, w


def rel_l2(a, b):
    a = torch.as_tensor(np.asarray(a)) if not torch.is_tensor(a) else a
    b = torch.as_tensor(np.asarray(b)) if not torch.is_tensor(b) else b
    return O.rel_l2(a.detach().cpu(), b.detach().cpu())
