"""The CUDA path against the COMMITTED vectors of tests/golden/ (made by tests/golden/make_golden.py from the oracle): the same
tolerances as the live-oracle parity tests, but nothing is recomputed on the CPU here except the seeded weights -- a drifting
oracle and a drifting kernel cannot cancel."""
import os

import numpy as np
import pytest
import torch

from helpers import O, ROOT, product, rel_l2

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


def test_dit_velocity_golden():
    g = np.load(os.path.join(GOLD, "dit_small.npz"))
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    ctx = ctxmod.LtxContext(ctxmod.LTXTransformerConfig(num_layers=2, num_attention_heads=2, caption_channels=192), 0)
    ctx.load_weights(O.make_dit_weights(ocfg, 1234))
    ctx.finalize_weights()
    lat, cx = torch.from_numpy(g["latent"]).bfloat16(), torch.from_numpy(g["context"]).bfloat16()   # fixtures are bf16-exact
    out = ctx.dit_forward(lat, cx, g["sigma"], g["mask"], (2, 4, 6))
    assert rel_l2(out, g["velocity"]) <= 1e-2, rel_l2(out, g["velocity"])
    stg = ctx.dit_forward(lat, cx, g["sigma"], g["mask"], (2, 4, 6), ctxmod.make_flags(stg_blocks=[1], skip_self_attn=True))
    assert rel_l2(stg, g["velocity_stg"]) <= 1e-2
    ctx.close()


def test_guided_euler_golden():
    g = np.load(os.path.join(GOLD, "guidance.npz"))
    ctxmod = product()
    ctx = ctxmod.LtxContext(ctxmod.LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
    x, vp = np.ascontiguousarray(g["x"]).copy(), np.ascontiguousarray(g["vp"]).copy()
    # make_golden.guidance_case: cfg 4.0, rescale 0.7, stg 0.5, GE gamma 0.3, sigma 0.8 -> 0.6
    ctx.guided_euler_step(x, g["vc"], g["vu"], g["vs"], vp, use_prev=True, cfg_scale=4.0, rescale_phi=0.7, stg_scale=0.5,
                          ge_gamma=0.3, sigma=0.8, sigma_next=0.6)
    assert rel_l2(x, g["out"]) <= 1e-5 and rel_l2(vp, g["v"]) <= 1e-5
    ctx.close()


def test_vae_frames_golden():
    ctxmod = product()
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    w = O.make_vae_weights(vcfg, 99)
    w = {k: (O.bf16_round(v) if k.endswith("conv.weight") else v) for k, v in w.items()}
    ctx = ctxmod.LtxContext(ctxmod.LTXTransformerConfig(num_layers=1, num_attention_heads=1, vae_base_channels=512,
                                                        vae_blocks_per_stage=1), 0)
    ctx.load_weights(w, prefix="vae.")
    ctx.finalize_weights()
    g = np.load(os.path.join(GOLD, "vae_small.npz"))
    fr = ctx.vae_decode(g["latent"][0])
    assert O.psnr(torch.from_numpy(fr), torch.from_numpy(g["frames"].astype(np.float32))) >= 40.0
    g2 = np.load(os.path.join(GOLD, "vae_tiled_small.npz"))
    ft = ctx.vae_decode_tiled(g2["latent"][0], 3, 1)
    assert ft.shape == g2["frames"].shape
    assert O.psnr(torch.from_numpy(ft), torch.from_numpy(g2["frames"].astype(np.float32))) >= 40.0
    ctx.close()


def test_av_velocities_golden():
    from ltx_video_swift_mlx_b200.transformer import LTX2Transformer
    g = np.load(os.path.join(GOLD, "av_small.npz"))
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    av = O.AVConfig(audio_heads=2)
    ctx = ctxmod.LtxContext(ctxmod.LTXTransformerConfig(num_layers=2, num_attention_heads=2, caption_channels=192,
                                                        audio_num_attention_heads=2), 0)
    ctx.load_weights(O.make_av_weights(ocfg, av, 2468))
    ctx.finalize_weights()
    model = LTX2Transformer(ctx)                       # the Swift argument list (T/LTX2Transformer.swift:240-251)
    b = lambda k: torch.from_numpy(g[k]).bfloat16()    # noqa: E731
    args = (b("video_latent"), b("audio_latent"), b("video_context"), b("audio_context"))
    v, a = model(*args, np.array([0.7], dtype=np.float32), np.array([0.55], dtype=np.float32), g["mask"], g["mask"], (2, 4, 6), 11)
    assert rel_l2(v, g["video_velocity"]) <= 1e-2 and rel_l2(a, g["audio_velocity"]) <= 1e-2
    v, a = model(*args, g["video_sigmas_tok"], np.array([0.55], dtype=np.float32), g["mask"], g["mask"], (2, 4, 6), 11)
    assert rel_l2(v, g["video_velocity_tok"]) <= 1e-2 and rel_l2(a, g["audio_velocity_tok"]) <= 1e-2
    with pytest.raises(ValueError):
        model(*args, np.zeros(7, dtype=np.float32), np.array([0.55], dtype=np.float32), None, None, (2, 4, 6), 11)
    ctx.close()
