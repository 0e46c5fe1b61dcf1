"""CPU checks of the oracle itself: known-answer constants from the reference's sources, closed-form cases, the
committed golden fixtures (regression pin of our restatement) and fp32-vs-fp64 self-consistency."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_sigma_tables_match_reference_constants():
    # S/LTXScheduler.swift:18-36 literal tables
    assert O.DISTILLED_SIGMA_VALUES == [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]
    assert O.STAGE_2_DISTILLED_SIGMA_VALUES == O.DISTILLED_SIGMA_VALUES[5:]
    # without a token count the distilled schedule is the raw table (:104-106)
    assert np.allclose(O.set_timesteps(8, True, None), O.DISTILLED_SIGMA_VALUES)


def test_sigma_kats_from_survey():
    # values obtained by executing S/LTXScheduler.swift:74-182 in float32
    kat = {512: [1, .99326, .986474, .979642, .972764, .897623, .653374, .1, 0],
           1536: [1, .994059, .988067, .982024, .975929, .908606, .680049, .1, 0],
           6144: [1, .995145, .990236, .985273, .980255, .923979, .720582, .1, 0]}
    for n, want in kat.items():
        assert np.allclose(O.set_timesteps(8, True, n), want, atol=2e-6)
    dev = O.set_timesteps(40, False, 1536)
    assert len(dev) == 41 and np.allclose(dev[:4], [1, .99204, .98381, .97528], atol=1e-5)
    assert np.allclose(dev[-3:], [.16485, .1, 0], atol=1e-5)
    g = np.load(os.path.join(GOLD, "sigmas.npz"))
    assert np.array_equal(g["distilled_1536"], np.array(O.set_timesteps(8, True, 1536)))
    assert np.array_equal(g["dev40_1536"], np.array(dev))


def test_latent_geometry():
    # P/VideoLatentShape.swift:35-41 ; frame formula 8(F'-1)+1 (V/VideoDecoder.swift:294)
    assert O.latent_shape(25, 512, 768) == (4, 16, 24)
    assert O.latent_shape(121, 512, 768) == (16, 16, 24)
    assert O.latent_shape(9, 512, 512) == (2, 16, 16)
    x = torch.arange(2 * 3 * 2 * 4 * 5, dtype=torch.float32).view(2, 3, 2, 4, 5)
    assert torch.equal(O.unpatchify(O.patchify(x), (2, 4, 5)), x)


def test_rope_table_structure():
    cfg = O.DiTConfig()
    cos, sin = O.rope_table(cfg, 2, 3, 4)
    assert cos.shape == (32, 24, 64) and sin.shape == (32, 24, 64)
    # the two left-pad entries are the identity rotation (T/LTXRoPE.swift:451-476) and land in head 0
    assert torch.all(cos[0, :, :2] == 1) and torch.all(sin[0, :, :2] == 0)
    assert torch.allclose(cos ** 2 + sin ** 2, torch.ones_like(cos), atol=1e-6)
    # first real frequency: idx_0 = pi/2 times the t position scaled to [-1, 1]
    grid = O.position_grid(2, 3, 4)
    st = 2 * (grid[0].double() / 20) - 1
    assert torch.allclose(cos[0, :, 2].double(), torch.cos(st * math.pi / 2), atol=1e-6)
    # rotation by the identity leaves padded channels untouched
    x = torch.randn(1, 24, 4096)
    y = O.apply_split_rope(x, cos, sin, 32)
    assert torch.equal(y[0, :, :2], x[0, :, :2]) and torch.equal(y[0, :, 64:66], x[0, :, 64:66])
    assert torch.allclose(y.norm(), x.norm(), rtol=1e-5)


def test_guidance_closed_forms():
    g = torch.Generator().manual_seed(0)
    x, vc, vu = [torch.randn(1, 8, 2, 3, 4, generator=g) for _ in range(3)]
    out, v = O.guided_euler_step(x, vc, None, None, None, 1.0, 0, 0, 0, 0.5, 0.0)
    assert torch.allclose(out, x - 0.5 * vc)                      # sigma' = 0 returns the denoised sample
    assert torch.equal(O.apply_cfg(vu, vc, 1.0), vc)              # scale 1 returns cond
    r = O.guidance_rescale(O.apply_cfg(vu, vc, 4.0), vc, 1.0)     # phi = 1: std of the result equals std of cond
    assert abs(float(r.std(unbiased=False)) - float(vc.std(unbiased=False))) < 1e-4
    out2, _ = O.guided_euler_step(x, vc, None, None, None, 1.0, 0, 0, 0, 0.5, 0.25)
    assert torch.allclose(out2, x + (0.25 - 0.5) * vc, atol=1e-6)  # Euler: x + (sigma' - sigma) v
    gd = np.load(os.path.join(GOLD, "guidance.npz"))
    o, vv = O.guided_euler_step(*[torch.from_numpy(gd[k]) for k in ("x", "vc", "vu", "vs", "vp")], 4.0, 0.7, 0.5, 0.3, 0.8, 0.6)
    assert np.allclose(o.numpy(), gd["out"], atol=1e-6) and np.allclose(vv.numpy(), gd["v"], atol=1e-6)


def test_conv3d_full_equals_three_conv2d_slices():
    # the reference executes the 3x3x3 conv as 3 conv2d calls over shifted temporal slices (V/VideoConvolution.swift:310-339)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 6, 4, 5, 7, generator=g)
    w = torch.randn(8, 6, 3, 3, 3, generator=g)
    b = torch.randn(8, generator=g)
    for causal in (False, True):
        ref = O.conv3d_full(x, w, b, causal)
        xp = torch.nn.functional.pad(x.reshape(1, 24, 5, 7), (1, 1, 1, 1), mode="reflect").view(1, 6, 4, 7, 9)
        xp = torch.cat([xp[:, :, :1]] * (2 if causal else 1) + [xp] + ([] if causal else [xp[:, :, -1:]]), 2)
        acc = 0
        for kt in range(3):
            sl = xp[:, :, kt:kt + 4].permute(0, 2, 1, 3, 4).reshape(4, 6, 7, 9)
            acc = acc + torch.nn.functional.conv2d(sl, w[:, :, kt])
        got = acc.view(1, 4, 8, 5, 7).permute(0, 2, 1, 3, 4) + b.view(1, -1, 1, 1, 1)
        assert torch.allclose(ref, got, atol=1e-4)


def test_d2s_and_unpatchify_are_permutations():
    x = torch.arange(1 * 16 * 2 * 3 * 4, dtype=torch.float32).view(1, 16, 2, 3, 4)
    y = O.depth_to_space(x, 2)
    assert y.shape == (1, 2, 4, 6, 8) and torch.equal(y.flatten().sort().values, x.flatten())
    # channel c*8 + p1*4 + p2*2 + p3 lands at (2t+p1, 2h+p2, 2w+p3)
    assert y[0, 1, 2 * 1 + 1, 2 * 2 + 0, 2 * 3 + 1] == x[0, 1 * 8 + 1 * 4 + 0 * 2 + 1, 1, 2, 3]
    z = torch.arange(1 * 48 * 2 * 3 * 4, dtype=torch.float32).view(1, 48, 2, 3, 4)
    u = O.vae_unpatchify(z, 4)
    assert u.shape == (1, 3, 2, 12, 16)
    # channel c*16 + pa*4 + pb lands at row 4h + pb, column 4w + pa (pW before pH, V/VideoDecoder.swift:270-272)
    assert u[0, 2, 1, 4 * 2 + 3, 4 * 1 + 2] == z[0, 2 * 16 + 2 * 4 + 3, 1, 2, 1]


def test_dit_oracle_golden_and_fp64():
    g = np.load(os.path.join(GOLD, "dit_small.npz"))
    cfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    w = O.make_dit_weights(cfg, 1234)
    args = (torch.from_numpy(g["latent"]), torch.from_numpy(g["context"]), torch.from_numpy(g["sigma"]),
            torch.from_numpy(g["mask"]), (2, 4, 6))
    vel = O.dit_forward(w, cfg, *args)
    assert O.rel_l2(vel, torch.from_numpy(g["velocity"])) < 1e-5
    v64 = O.dit_forward(w, cfg, *args, dtype=torch.float64)
    assert O.rel_l2(vel, v64) < 1e-4                     # fp32 oracle is within the fp32-mode tolerance of fp64
    stg = O.dit_forward(w, cfg, *args, stg_blocks=[1], skip_self_attn=True)
    assert O.rel_l2(stg, torch.from_numpy(g["velocity_stg"])) < 1e-5
    assert O.rel_l2(stg, vel) > 1e-2                     # the perturbed pass really differs
    # mask semantic: additive -10000 on padded keys == dropping those keys
    m0 = args[3].clone()
    keep = m0[0].bool()
    dropped = O.dit_forward(w, cfg, args[0], args[1][:, keep], args[2], None, (2, 4, 6))
    assert O.rel_l2(vel, dropped) < 1e-5


def test_vae_oracle_golden():
    g = np.load(os.path.join(GOLD, "vae_small.npz"))
    cfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    w = O.make_vae_weights(cfg, 99)
    w = {k: (O.bf16_round(v) if k.endswith("conv.weight") else v) for k, v in w.items()}
    fr = O.decode_video(w, cfg, torch.from_numpy(g["latent"]))
    assert fr.shape == (9, 64, 96, 3) and float(fr.min()) >= 0 and float(fr.max()) <= 1
    assert O.psnr(fr, torch.from_numpy(g["frames"].astype(np.float32))) > 55      # fixture stored as fp16
    # frame count formula and the causal switch
    cfg_c = O.VAEConfig(base_channels=512, blocks_per_stage=1, causal=True)
    assert O.psnr(O.decode_video(w, cfg_c, torch.from_numpy(g["latent"])), fr) < 50


def test_encoder_shapes_and_space_to_depth():
    """V/VideoEncoder.swift: 8k+1 frames -> k+1 latent frames, /32 spatially; space-to-depth front-pads odd T with frame 0
    and is the inverse of the decoder's depth-to-space on even T."""
    ecfg = O.EncoderConfig(base_channels=64)
    w = O.make_encoder_weights(ecfg, 1)
    assert w["conv_in.conv.weight"].shape == (64, 48, 3, 3, 3) and w["conv_out.conv.weight"].shape[0] == 129
    assert w["down_blocks_0.downsamplers.conv.conv.weight"].shape[0] == 128 // 4
    assert w["down_blocks_1.downsamplers.conv.conv.weight"].shape[0] == 256 // 2
    for T, Tl in ((1, 1), (9, 2), (17, 3)):
        z = O.vae_encode(w, ecfg, torch.randn(1, 3, T, 64, 64, generator=torch.Generator().manual_seed(T)))
        assert z.shape == (1, 128, Tl, 2, 2)
    x = torch.randn(1, 16, 4, 6, 8)
    assert torch.equal(O.depth_to_space(O.space_to_depth(x, (2, 2, 2)), 16), x)
    odd = torch.randn(1, 4, 3, 2, 2)
    s = O.space_to_depth(odd, (2, 1, 1))
    assert s.shape == (1, 8, 2, 2, 2)
    assert torch.equal(s[:, 0::2, 0], odd[:, :, 0]) and torch.equal(s[:, 1::2, 0], odd[:, :, 0])   # (pad, frame 0)
    assert torch.equal(s[:, 0::2, 1], odd[:, :, 1]) and torch.equal(s[:, 1::2, 1], odd[:, :, 2])
    # patchify puts pW before pH in the channel index (:13-32)
    img = torch.arange(3 * 8 * 8, dtype=torch.float32).view(1, 3, 1, 8, 8)
    p = O.encoder_patchify(img)
    assert p.shape == (1, 48, 1, 2, 2)
    c, pw, ph = 1, 2, 3
    assert p[0, (c * 4 + pw) * 4 + ph, 0, 1, 0] == img[0, c, 0, 4 + ph, pw]


def test_upscaler_and_adain_properties():
    """SpatialUpscaler doubles H and W only; GroupNorm statistics span (D,H,W,C/32); AdaIN transplants per-channel
    mean / population std (P/LatentUtils.swift:201-227); re-noise is the stage-2 mix (:2644-2647)."""
    uw = O.make_upscaler_weights(mid=64, blocks=1, seed=2)
    lat = torch.randn(1, 128, 3, 4, 6, generator=torch.Generator().manual_seed(3))
    up = O.spatial_upscaler(uw, lat, 1)
    assert up.shape == (1, 128, 3, 8, 12)
    x = torch.randn(1, 64, 2, 3, 3)
    y = O._group_norm(x, torch.ones(64), torch.zeros(64))
    g = y.reshape(1, 32, -1)
    assert g.mean(-1).abs().max() < 1e-5 and (g.var(-1, unbiased=False) - 1).abs().max() < 1e-3
    ref = torch.randn(1, 128, 3, 2, 3) * 0.5 + 1.0
    a = O.adain_filter_latent(up, ref)
    assert torch.allclose(a.mean((2, 3, 4)), ref.mean((2, 3, 4)), atol=1e-5)
    assert torch.allclose(a.var((2, 3, 4), unbiased=False).sqrt(), ref.var((2, 3, 4), unbiased=False).sqrt(), atol=1e-4)
    assert torch.equal(O.adain_filter_latent(up, ref, 0.0), up)
    half = O.adain_filter_latent(up, ref, 0.5)
    assert torch.allclose(half, 0.5 * a + 0.5 * up, atol=1e-6)
    n = torch.randn_like(up)
    assert torch.allclose(O.renoise(up, n, 0.909375), 0.909375 * n + (1 - 0.909375) * up)


def test_av_oracle_structure():
    """Dual audio/video restatement (T/LTX2Transformer.swift, T/LTX2TransformerBlock.swift): the generic RoPE builder
    reproduces the video table, audio positions follow createAudioPositionGrid (T/LTXRoPE.swift:627-655), the two streams are
    coupled only through the cross-modal attentions, and zero cross-modal gates decouple them exactly."""
    cfg = O.DiTConfig(num_layers=1, num_heads=2, caption_channels=64)
    av = O.AVConfig(audio_heads=2)
    c1, s1 = O.rope_table(cfg, 2, 3, 4)
    c2, s2 = O.rope_table_nd(O.position_grid(2, 3, 4), cfg.inner_dim, cfg.num_heads, cfg.rope_theta, cfg.max_pos)
    assert torch.equal(c1, c2) and torch.equal(s1, s2)
    g = O.audio_position_grid(4)
    assert g.shape == (1, 4)
    # frame 0 covers mel [0, 1), frame i >= 1 covers [4i-3, 4i+1): mid-points * hop / sr
    want = torch.tensor([0.5, 3.0, 7.0, 11.0]) * 160.0 / 16000.0
    assert torch.allclose(g[0], want)
    w = O.make_av_weights(cfg, av, 4)
    gen = torch.Generator().manual_seed(1)
    fhw, Ta, S = (2, 2, 3), 5, 7
    vl, al = torch.randn(1, 12, 128, generator=gen), torch.randn(1, Ta, 128, generator=gen)
    vc, ac = torch.randn(1, S, 64, generator=gen), torch.randn(1, S, 64, generator=gen)
    sg = torch.tensor([0.6])
    v0, a0 = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, sg, sg, None, None, fhw, Ta)
    v1, a1 = O.av_dit_forward(w, cfg, av, vl, al * 2.0, vc, ac, sg, sg, None, None, fhw, Ta)
    assert v0.shape == (1, 12, 128) and a0.shape == (1, Ta, 128)
    assert O.rel_l2(v1, v0) > 1e-4                     # audio reaches the video stream through a2v
    # zero a2v gate (table row 4 and the gate embedder's output layer): the video stream no longer sees the audio
    wz = dict(w)
    wz["transformer_blocks.0.scale_shift_table_a2v_ca_video"] = w["transformer_blocks.0.scale_shift_table_a2v_ca_video"].clone()
    wz["transformer_blocks.0.scale_shift_table_a2v_ca_video"][4] = 0
    wz["av_ca_a2v_gate_adaln_single.linear.weight"] = torch.zeros_like(w["av_ca_a2v_gate_adaln_single.linear.weight"])
    wz["av_ca_a2v_gate_adaln_single.linear.bias"] = torch.zeros_like(w["av_ca_a2v_gate_adaln_single.linear.bias"])
    v2, _ = O.av_dit_forward(wz, cfg, av, vl, al, vc, ac, sg, sg, None, None, fhw, Ta)
    v3, _ = O.av_dit_forward(wz, cfg, av, vl, al * 2.0, vc, ac, sg, sg, None, None, fhw, Ta)
    assert torch.equal(v2, v3)


def test_av_oracle_per_token_sigmas_and_image_conditioned_loop():
    """videoTimesteps [1, N] (T/LTX2Transformer.swift:273-298): a constant per-token vector equals the scalar call; in the
    image-to-video loop (P/LTXPipeline.swift:1262-1298, 1381-1391) frame 0 keeps the image latent and the other frames move."""
    cfg = O.DiTConfig(num_layers=1, num_heads=2, caption_channels=64)
    av = O.AVConfig(audio_heads=2)
    w = O.make_av_weights(cfg, av, 6)
    gen = torch.Generator().manual_seed(2)
    fhw, Ta, S = (2, 2, 3), 5, 7
    N = 12
    vl, al = torch.randn(1, N, 128, generator=gen), torch.randn(1, Ta, 128, generator=gen)
    vc, ac = torch.randn(1, S, 64, generator=gen), torch.randn(1, S, 64, generator=gen)
    sg = torch.tensor([0.6])
    v0, a0 = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, sg, sg, None, None, fhw, Ta)
    v1, a1 = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, torch.full((1, N), 0.6), sg, None, None, fhw, Ta)
    assert O.rel_l2(v1, v0) < 1e-5 and O.rel_l2(a1, a0) < 1e-5
    ts = torch.full((1, N), 0.6)
    ts[:, :6] = 0.0
    v2, _ = O.av_dit_forward(w, cfg, av, vl, al, vc, ac, ts, sg, None, None, fhw, Ta)
    assert O.rel_l2(v2[:, :6], v0[:, :6]) > 1e-3
    vn, an = torch.randn(1, 128, *fhw, generator=gen), torch.randn(1, Ta, 128, generator=gen)
    img = torch.randn(1, 128, 1, 2, 3, generator=gen)
    sig = [1.0, 0.7, 0.3, 0.0]
    lv, la = O.av_denoise_loop(w, cfg, av, vn, an, vc, ac, None, sig, image_latent=img)
    assert torch.equal(lv[:, :, 0:1], img) and lv.shape == vn.shape and la.shape == an.shape
    assert not torch.allclose(lv[:, :, 1:], vn[:, :, 1:])
    inj = [torch.randn(1, 128, 1, 2, 3, generator=gen) for _ in range(3)]
    lv2, _ = O.av_denoise_loop(w, cfg, av, vn, an, vc, ac, None, sig, image_latent=img, inject_noise=inj, image_cond_noise_scale=0.1)
    assert torch.allclose(lv2[:, :, 0:1], img + 0.1 * inj[2] * 0.3 * 0.3, atol=1e-6)


def test_vae_temporal_tiling_oracle():
    """decodeWithTemporalTiling (V/VideoDecoder.swift:517-602): frame count follows the chunk arithmetic, frames outside every
    cross-fade equal the chunk's own decode, the first cross-faded frame (weight 0) is still the earlier chunk's."""
    cfg = O.VAEConfig(base_channels=64, blocks_per_stage=1)
    w = O.make_vae_weights(cfg, 3)
    z = torch.randn(1, 128, 5, 2, 2, generator=torch.Generator().manual_seed(4))
    tiled = O.decode_video(w, cfg, z, temporal_tile_size=3, temporal_tile_overlap=1)      # chunks [0,3) and [2,5): 17 + 17 - 8
    assert tiled.shape == (26, 64, 64, 3)
    a = O.decode_video(w, cfg, z[:, :, 0:3])
    b = O.decode_video(w, cfg, z[:, :, 2:5])
    assert torch.allclose(tiled[:10], a[:10], atol=1e-6) and torch.allclose(tiled[17:], b[8:], atol=1e-6)
    mid = torch.clamp((a[13] * 0.5 + b[4] * 0.5), 0, 1)                                   # frame 9 + 4: weight 4 / 8
    inside = (a[13] > 1e-3) & (a[13] < 1 - 1e-3) & (b[4] > 1e-3) & (b[4] < 1 - 1e-3)
    assert torch.allclose(tiled[13][inside], mid[inside], atol=1e-6)
    assert torch.equal(O.decode_video(w, cfg, z, temporal_tile_size=8), O.decode_video(w, cfg, z))   # fits one tile


def test_av_and_tiled_vae_golden():
    """Committed vectors of the dual forward (scalar and per-token video sigmas) and of the temporally tiled decode."""
    g = np.load(os.path.join(GOLD, "av_small.npz"))
    cfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=192)
    av = O.AVConfig(audio_heads=2)
    w = O.make_av_weights(cfg, av, 2468)
    t = lambda k: torch.from_numpy(g[k])                                     # noqa: E731
    args = (t("video_latent"), t("audio_latent"), t("video_context"), t("audio_context"))
    v, a = O.av_dit_forward(w, cfg, av, *args, torch.tensor([0.7]), torch.tensor([0.55]), t("mask"), t("mask"), (2, 4, 6), 11)
    assert O.rel_l2(v, t("video_velocity")) < 1e-5 and O.rel_l2(a, t("audio_velocity")) < 1e-5
    v, a = O.av_dit_forward(w, cfg, av, *args, t("video_sigmas_tok"), torch.tensor([0.55]), t("mask"), t("mask"), (2, 4, 6), 11)
    assert O.rel_l2(v, t("video_velocity_tok")) < 1e-5 and O.rel_l2(a, t("audio_velocity_tok")) < 1e-5
    g2 = np.load(os.path.join(GOLD, "vae_tiled_small.npz"))
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    vw = O.make_vae_weights(vcfg, 99)
    vw = {k: (O.bf16_round(x) if k.endswith("conv.weight") else x) for k, x in vw.items()}
    fr = O.decode_video(vw, vcfg, torch.from_numpy(g2["latent"]), temporal_tile_size=3, temporal_tile_overlap=1)
    assert fr.shape == (26, 64, 64, 3)
    assert O.psnr(fr, torch.from_numpy(g2["frames"].astype(np.float32))) > 55


def test_parameter_inventory_matches_the_reference_constructors():
    """Parameter count per block as a function of the width, from the constructors (T/LTXAttention.swift:117-157,
    T/LTXFeedForward.swift:39-52, T/LTXTransformerBlock.swift:140-160): 8 Linears (D x D + D) in the two attentions, four
    RMSNorm weights, the 4x FFN and the 6 x D table.  At D = 4096 that is 268 529 664 per block and, with the global tensors,
    13.04 B for the 48-block video model (26.1 GB bf16, the measured 27 GB mean RAM of the reference)."""
    def per_block(D, mult=4):
        return 8 * (D * D + D) + 4 * D + (D * mult * D + mult * D) + (mult * D * D + D) + 6 * D
    assert per_block(4096) == 268_529_664
    cfg = O.DiTConfig(num_layers=3, num_heads=2, head_dim=128, caption_channels=192)
    w = O.make_dit_weights(cfg, 0)
    D = cfg.inner_dim
    blk = sum(v.numel() for k, v in w.items() if k.startswith("transformer_blocks.1."))
    assert blk == per_block(D)
    glob = sum(v.numel() for k, v in w.items() if not k.startswith("transformer_blocks."))
    # patchify_proj, adaln_single (256 -> D -> D, D -> 6D), caption_projection (Cc -> D -> D), proj_out, scale_shift_table
    want = (128 * D + D) + (256 * D + D) + (D * D + D) + (D * 6 * D + 6 * D) + (192 * D + D) + (D * D + D) + (D * 128 + 128) + 2 * D
    assert glob == want
    full_glob = (128 * 4096 + 4096) + (256 * 4096 + 4096) + (4096 ** 2 + 4096) + (6 * 4096 ** 2 + 6 * 4096) + \
                (3840 * 4096 + 4096) + (4096 ** 2 + 4096) + (4096 * 128 + 128) + 2 * 4096
    total = 48 * per_block(4096) + full_glob
    assert abs(total / 1e9 - 13.04) < 0.01 and abs(full_glob / 1e6 - 152.1) < 0.1
