"""Generates tests/golden/*.npz from the CPU oracle (seeded weights and inputs).

The reference ships no golden vectors for this path and cannot run here (parity unpinned), so these fixtures
pin OUR oracle: they catch accidental drift of the restatement, and let the GPU parity tests check against committed
numbers as well as against the live oracle.  The sigma tables are the values obtained by executing
Scheduler/LTXScheduler.swift:74-182 in float32.  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ltx_oracle as O  # noqa: E402


def dit_case():
    cfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    w = O.make_dit_weights(cfg, 1234)
    g = torch.Generator().manual_seed(4321)
    fhw = (2, 4, 6)
    lat = torch.randn(1, 48, 128, generator=g).bfloat16().float()
    ctx = torch.randn(1, 40, 192, generator=g)
    ctx = (ctx / ctx.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16().float()
    mask = torch.ones(1, 40, dtype=torch.int32)
    mask[:, :7] = 0
    sig = torch.tensor([0.7])
    vel, blocks = O.dit_forward(w, cfg, lat, ctx, sig, mask, fhw, return_blocks=True)
    vel_stg = O.dit_forward(w, cfg, lat, ctx, sig, mask, fhw, stg_blocks=[1], skip_self_attn=True)
    cos, sin = O.rope_table(cfg, *fhw)
    return dict(latent=lat.numpy(), context=ctx.numpy(), mask=mask.numpy(), sigma=sig.numpy(), velocity=vel.numpy(),
                velocity_stg=vel_stg.numpy(), block_means=np.array([float(b.mean()) for b in blocks], dtype=np.float64),
                rope_cos_head0=cos[0].numpy(), rope_sin_head1=sin[1].numpy())


def vae_case():
    cfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    w = O.make_vae_weights(cfg, 99)
    w = {k: (O.bf16_round(v) if k.endswith("conv.weight") else v) for k, v in w.items()}
    g = torch.Generator().manual_seed(98)
    z = torch.randn(1, 128, 2, 2, 3, generator=g)
    frames = O.decode_video(w, cfg, z)
    return dict(latent=z.numpy(), frames=frames.numpy().astype(np.float16))


def sigma_case():
    return dict(distilled_512=np.array(O.set_timesteps(8, True, 512)), distilled_1536=np.array(O.set_timesteps(8, True, 1536)),
                distilled_6144=np.array(O.set_timesteps(8, True, 6144)), dev40_1536=np.array(O.set_timesteps(40, False, 1536)))


def guidance_case():
    g = torch.Generator().manual_seed(7)
    shp = (1, 128, 2, 4, 6)
    x, vc, vu, vs, vp = [torch.randn(shp, generator=g) for _ in range(5)]
    out, v = O.guided_euler_step(x, vc, vu, vs, vp, 4.0, 0.7, 0.5, 0.3, 0.8, 0.6)
    return dict(x=x.numpy(), vc=vc.numpy(), vu=vu.numpy(), vs=vs.numpy(), vp=vp.numpy(), out=out.numpy(), v=v.numpy())


def av_case():
    """Dual audio / video forward (T/LTX2Transformer.swift:240-392): scalar sigmas, and per-token video sigmas (image-to-video)."""
    cfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    av = O.AVConfig(audio_heads=2)
    w = O.make_av_weights(cfg, av, 2468)
    g = torch.Generator().manual_seed(8642)
    fhw, Ta, S = (2, 4, 6), 11, 40
    vl = torch.randn(1, 48, 128, generator=g).bfloat16().float()
    al = torch.randn(1, Ta, 128, generator=g).bfloat16().float()

    def text():
        t = torch.randn(1, S, 192, generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16().float()
    vc, ac = text(), text()
    mask = torch.ones(1, S, dtype=torch.int32)
    mask[:, :5] = 0
    sv, sa = torch.tensor([0.7]), torch.tensor([0.55])
    v, a = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, sv, sa, mask, mask, fhw, Ta)
    ts = torch.full((1, 48), 0.7)
    ts[:, :24] = 0.0
    v_tok, a_tok = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, ts, sa, mask, mask, fhw, Ta)
    return dict(video_latent=vl.numpy(), audio_latent=al.numpy(), video_context=vc.numpy(), audio_context=ac.numpy(),
                mask=mask.numpy(), video_velocity=v.numpy(), audio_velocity=a.numpy(), video_sigmas_tok=ts.numpy(),
                video_velocity_tok=v_tok.numpy(), audio_velocity_tok=a_tok.numpy())


def vae_tiled_case():
    """decodeWithTemporalTiling (V/VideoDecoder.swift:517-602) on the vae_small weights: 5 latent frames, tile 3, overlap 1."""
    cfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    w = O.make_vae_weights(cfg, 99)
    w = {k: (O.bf16_round(v) if k.endswith("conv.weight") else v) for k, v in w.items()}
    z = torch.randn(1, 128, 5, 2, 2, generator=torch.Generator().manual_seed(97))
    frames = O.decode_video(w, cfg, z, temporal_tile_size=3, temporal_tile_overlap=1)
    return dict(latent=z.numpy(), frames=frames.numpy().astype(np.float16))


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "dit_small.npz"), **dit_case())
    np.savez_compressed(os.path.join(HERE, "vae_small.npz"), **vae_case())
    np.savez_compressed(os.path.join(HERE, "sigmas.npz"), **sigma_case())
    np.savez_compressed(os.path.join(HERE, "guidance.npz"), **guidance_case())
    np.savez_compressed(os.path.join(HERE, "av_small.npz"), **av_case())
    np.savez_compressed(os.path.join(HERE, "vae_tiled_small.npz"), **vae_tiled_case())
    for f in sorted(os.listdir(HERE)):
        print(f, os.path.getsize(os.path.join(HERE, f)))
