"""GPU parity of the rows either side of the denoise loop: VAE encoder, latent upscaler, AdaIN and the
two-stage glue, each against the CPU oracle on the same seeded inputs.  Tolerances: bf16 tensor-core operands with fp32
accumulation vs the fp32 oracle -> rel-L2 <= 1e-2 (the north_star's bf16 bound); AdaIN / re-noise are fp32 elementwise."""
import numpy as np
import pytest
import torch

from helpers import O, product, rel_l2

pytestmark = pytest.mark.gpu


def _bf16_weights(w):
    return {k: (O.bf16_round(v) if (k.endswith(".weight") and v.ndim >= 2) else v) for k, v in w.items()}


def _ctx(**kw):
    ctxmod = product()
    return ctxmod.LtxContext(ctxmod.LTXTransformerConfig(num_layers=1, num_attention_heads=1, **kw), 0)


def _vae_stats(ctx, seed=3):
    g = torch.Generator().manual_seed(seed)
    mean, std = torch.randn(128, generator=g) * 0.1, 1.0 + 0.1 * torch.rand(128, generator=g)
    return mean, std


@pytest.mark.parametrize("T,H,W", [(1, 64, 96), (9, 64, 64), (5, 96, 64)])
def test_vae_encoder_matches_oracle(T, H, W):
    ecfg = O.EncoderConfig(base_channels=64)
    w = _bf16_weights(O.make_encoder_weights(ecfg, 11))
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    vw = _bf16_weights(O.make_vae_weights(vcfg, 12))
    ctx = _ctx(vae_base_channels=512, vae_blocks_per_stage=1)
    ctx.load_weights(w, prefix="vae_encoder.")
    ctx.load_weights(vw, prefix="vae.")
    ctx.finalize_weights()
    px = torch.rand(1, 3, T, H, W, generator=torch.Generator().manual_seed(T)) * 2 - 1
    ref = O.vae_encode(w, ecfg, px)
    got = ctx.vae_encode(px.numpy(), normalize=False)
    assert got.shape == tuple(ref.shape[1:])
    assert rel_l2(got, ref[0]) <= 1e-2
    refn = O.encode_image_latent(w, ecfg, px, vw["mean_of_means"], vw["std_of_means"])
    assert rel_l2(ctx.vae_encode(px.numpy(), normalize=True), refn[0]) <= 1e-2
    ctx.close()


def test_vae_encoder_ltx2_channel_plan_image():
    """The real channel plan (128 -> 2048, 4/6/6/2 res blocks, 129-channel conv_out) on one 128x192 image."""
    ecfg = O.EncoderConfig()
    w = _bf16_weights(O.make_encoder_weights(ecfg, 13))
    ctx = _ctx()
    ctx.load_weights(w, prefix="vae_encoder.")
    ctx.finalize_weights()
    px = torch.rand(1, 3, 1, 128, 192, generator=torch.Generator().manual_seed(5)) * 2 - 1
    ref = O.vae_encode(w, ecfg, px)
    got = ctx.vae_encode(px.numpy(), normalize=False)
    assert got.shape == (128, 1, 4, 6)
    assert rel_l2(got, ref[0]) <= 1e-2
    ctx.close()


@pytest.mark.parametrize("F,H,W,mid,blocks", [(3, 4, 6, 128, 1), (2, 8, 8, 256, 2)])
def test_upscaler_matches_oracle(F, H, W, mid, blocks):
    uw = _bf16_weights(O.make_upscaler_weights(mid=mid, blocks=blocks, seed=21))
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    vw = _bf16_weights(O.make_vae_weights(vcfg, 22))
    ctx = _ctx(vae_base_channels=512, vae_blocks_per_stage=1)
    ctx.load_weights(uw, prefix="upscaler.")
    ctx.load_weights(vw, prefix="vae.")
    ctx.finalize_weights()
    lat = torch.randn(1, 128, F, H, W, generator=torch.Generator().manual_seed(7))
    ref = O.upsample_latents(uw, lat, vw["mean_of_means"], vw["std_of_means"], blocks)
    got = ctx.upscale_latent(lat.numpy())
    assert got.shape == (128, F, 2 * H, 2 * W)
    assert rel_l2(got, ref[0]) <= 1e-2
    ctx.close()


def test_adain_and_renoise_match_oracle():
    ctx = _ctx()
    g = torch.Generator().manual_seed(9)
    lat = torch.randn(1, 128, 3, 8, 12, generator=g) * 1.7 + 0.3
    ref = torch.randn(1, 128, 3, 4, 6, generator=g) * 0.6 - 0.2
    for factor in (1.0, 0.4):
        want = O.adain_filter_latent(lat, ref, factor)
        got = ctx.adain_filter(lat[0].numpy(), ref[0].numpy(), factor)
        assert rel_l2(got, want[0]) <= 1e-5
    # factor 1 is idempotent: the result already has the reference statistics
    once = ctx.adain_filter(lat[0].numpy(), ref[0].numpy(), 1.0)
    twice = ctx.adain_filter(once, ref[0].numpy(), 1.0)
    assert rel_l2(twice, once) <= 1e-5
    ctx.close()


def test_two_stage_resident_matches_oracle():
    """Stage 1 (2 steps) -> upscale + AdaIN + re-noise -> stage 2 (2 steps), device-resident, against the oracle chain."""
    from ltx_video_swift_mlx_b200.pipeline import generate_two_stage_resident
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    pcfg = ctxmod.LTXTransformerConfig(num_layers=2, num_attention_heads=2, caption_channels=192, vae_base_channels=512,
                                       vae_blocks_per_stage=1)
    w = O.make_dit_weights(ocfg, 31)
    uw = _bf16_weights(O.make_upscaler_weights(mid=128, blocks=1, seed=32))
    vw = _bf16_weights(O.make_vae_weights(O.VAEConfig(base_channels=512, blocks_per_stage=1), 33))
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights(w)
    ctx.load_weights(uw, prefix="upscaler.")
    ctx.load_weights(vw, prefix="vae.")
    ctx.finalize_weights()
    g = torch.Generator().manual_seed(1)
    F, H, W = 2, 4, 6
    n1 = torch.randn(1, 128, F, H, W, generator=g)
    n2 = torch.randn(1, 128, F, 2 * H, 2 * W, generator=g)
    text = torch.randn(1, 40, 192, generator=g)
    text = (text / text.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    s1 = O.set_timesteps(8, True, F * H * W)[5:8]
    s2 = [0.909375, 0.725, 0.0]
    z1 = O.denoise_loop(w, ocfg, n1, text.float(), None, s1)
    up = O.upsample_latents(uw, z1, vw["mean_of_means"], vw["std_of_means"], 1)
    z = O.renoise(O.adain_filter_latent(up, z1), n2, s2[0])
    ref = O.denoise_loop(w, ocfg, n2, text.float(), None, s2, init_latent=z)
    got = generate_two_stage_resident(ctx, n1.numpy(), n2.numpy(), text, None, s1, s2)
    assert got.shape == (1, 128, F, 2 * H, 2 * W)
    assert rel_l2(got, ref) <= 2e-2
    # host-buffer variant of the stage-2 start (ltx_denoise_begin_from_latent) gives the same start point
    ctx.denoise_begin_from_latent(O.adain_filter_latent(up, z1)[0].numpy(), n2[0].numpy(), s2[0], (F, 2 * H, 2 * W), text)
    assert rel_l2(ctx.denoise_get_latent(), z[0]) <= 1e-5
    ctx.close()


def test_components_can_be_loaded_and_finalized_incrementally():
    """loadVAEEncoder() runs on demand, after loadModels() (P/LTXPipeline.swift:1870-1884): a second ltx_finalize_weights packs
    only the new component and leaves the finalized ones working."""
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    vw = _bf16_weights(O.make_vae_weights(vcfg, 61))
    ctx = _ctx(vae_base_channels=512, vae_blocks_per_stage=1)
    ctx.load_weights(vw, prefix="vae.")
    ctx.finalize_weights()
    z = torch.randn(1, 128, 2, 2, 3, generator=torch.Generator().manual_seed(2))
    before = ctx.vae_decode(z[0].numpy())
    ecfg = O.EncoderConfig(base_channels=64)
    we = _bf16_weights(O.make_encoder_weights(ecfg, 62))
    ctx.load_weights(we, prefix="vae_encoder.")
    ctx.finalize_weights()
    px = torch.rand(1, 3, 1, 64, 64, generator=torch.Generator().manual_seed(3)) * 2 - 1
    assert rel_l2(ctx.vae_encode(px.numpy(), normalize=False), O.vae_encode(we, ecfg, px)[0]) <= 1e-2
    assert np.array_equal(ctx.vae_decode(z[0].numpy()), before)
    ctx.close()


def test_error_paths_map_to_ltx_error_codes():
    """Bad shapes / missing weights fail loudly with the LTXError-equivalent codes (LTXVideo.swift:66-107): 2 = generationFailed
    (invalid argument), 4 = weightLoadingFailed, 5 = unsupported; nothing falls back to a CPU path."""
    from ltx_video_swift_mlx_b200._lib import LtxError
    ctx = _ctx()
    with pytest.raises(LtxError) as e:                       # encoder weights were never loaded
        ctx.vae_encode(np.zeros((3, 1, 64, 64), dtype=np.float32), normalize=False)
    assert e.value.code == 4
    with pytest.raises(LtxError) as e:                       # upscaler weights were never loaded
        ctx.upscale_latent(np.zeros((128, 1, 4, 4), dtype=np.float32))
    assert e.value.code == 4
    ecfg = O.EncoderConfig(base_channels=64)
    ctx.load_weights(_bf16_weights(O.make_encoder_weights(ecfg, 71)), prefix="vae_encoder.")
    ctx.finalize_weights()
    with pytest.raises(LtxError) as e:                       # H, W must be multiples of 32
        ctx.lib.ltx_vae_encode  # noqa: B018 (symbol exists)
        px = np.zeros((3, 1, 72, 64), dtype=np.float32)
        out = np.zeros((128, 1, 2, 2), dtype=np.float32)
        ctx._check(ctx.lib.ltx_vae_encode(ctx.handle, px.ctypes.data, 1, 72, 64, 0, out.ctypes.data))
    assert e.value.code == 2
    with pytest.raises(LtxError) as e:                       # normalisation needs the decoder's latent statistics
        ctx.vae_encode(np.zeros((3, 1, 64, 64), dtype=np.float32), normalize=True)
    assert e.value.code == 4
    with pytest.raises(LtxError) as e:                       # LoRA target must be loaded and not yet packed
        ctx.fuse_lora("transformer_blocks.0.attn1.to_q.weight", torch.zeros(8, 256).bfloat16(), torch.zeros(256, 8).bfloat16())
    assert e.value.code == 4
    ctx.close()
