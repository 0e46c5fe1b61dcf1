"""Full DiT forward / denoise-loop parity against the CPU oracle (rel-L2 <= 1e-2 in bf16 mode, north_star)."""
import numpy as np
import pytest
import torch

from helpers import O, make_ctx_with_dit, rel_l2, small_dit_config, product

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _inputs(ocfg, fhw, S, seed, B=1, mask_prefix=0):
    g = torch.Generator().manual_seed(seed)
    N = fhw[0] * fhw[1] * fhw[2]
    lat = torch.randn(B, N, ocfg.in_channels, generator=g).bfloat16()
    ctx = torch.randn(B, S, ocfg.caption_channels, generator=g)
    ctx = (ctx / ctx.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()     # unit-RMS rows like the real connector
    mask = torch.ones(B, S, dtype=torch.int32)
    if mask_prefix:
        mask[:, :mask_prefix] = 0
    return lat, ctx, mask


@pytest.mark.parametrize("layers,heads,fhw,S,B,mask_prefix", [
    (2, 2, (2, 4, 6), 40, 1, 7), (3, 4, (4, 8, 10), 150, 1, 0), (2, 2, (1, 4, 4), 24, 2, 5),
    (1, 32, (2, 4, 6), 40, 1, 7),    # full LTX-2 width (D = 4096): exercises the D-specialised row kernels
    (1, 32, (2, 16, 16), 128, 1, 0), # BASELINE config 0: one LTX-2 block, latent 512x512x9 (N = 512), 128 text tokens
])
def test_dit_forward_matches_oracle(layers, heads, fhw, S, B, mask_prefix):
    ocfg, pcfg = small_dit_config(layers, heads)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=layers * 10 + heads)
    lat, cx, mask = _inputs(ocfg, fhw, S, 5, B, mask_prefix)
    sig = torch.tensor([0.7, 0.3][:B])
    ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), sig, mask if mask_prefix else None, fhw)
    out = ctx.dit_forward(lat, cx, sig.numpy(), mask if mask_prefix else None, fhw)
    assert np.isfinite(out).all()
    err = rel_l2(out, ref)
    assert err <= TOL, err
    # fp32 host inputs take the cast path and must agree with the bf16 call
    out2 = ctx.dit_forward(lat.float(), cx.float(), sig.numpy(), mask if mask_prefix else None, fhw)
    assert rel_l2(out2, out) <= 1e-6
    ctx.close()


def test_dit_stg_and_cross_scale_flags():
    ocfg, pcfg = small_dit_config(3, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=3)
    fhw, S = (2, 4, 6), 40
    lat, cx, mask = _inputs(ocfg, fhw, S, 9)
    sig = torch.tensor([0.5])
    ctxmod = product()
    for stg in ([0], [1], [2], [0, 1]):
        ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), sig, None, fhw, stg_blocks=stg, skip_self_attn=True)
        out = ctx.dit_forward(lat, cx, sig.numpy(), None, fhw, ctxmod.make_flags(stg_blocks=stg, skip_self_attn=True))
        assert rel_l2(out, ref) <= TOL
    ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), sig, None, fhw, stg_blocks=[1], skip_self_attn=True, skip_ff=True)
    out = ctx.dit_forward(lat, cx, sig.numpy(), None, fhw, ctxmod.make_flags(stg_blocks=[1], skip_self_attn=True, skip_ff=True))
    assert rel_l2(out, ref) <= TOL
    ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), sig, None, fhw, cross_attn_scale={1: 0.25})
    out = ctx.dit_forward(lat, cx, sig.numpy(), None, fhw, ctxmod.make_flags(cas_blocks=[1], cross_attn_scale=0.25))
    assert rel_l2(out, ref) <= TOL
    ctx.close()


def test_context_cache_key_is_equivalent():
    ocfg, pcfg = small_dit_config(2, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=4)
    fhw, S = (2, 4, 6), 40
    lat, cx, mask = _inputs(ocfg, fhw, S, 13, mask_prefix=3)
    sig = np.array([0.6], dtype=np.float32)
    ctxmod = product()
    a = ctx.dit_forward(lat, cx, sig, mask, fhw)
    b = ctx.dit_forward(lat, cx, sig, mask, fhw, ctxmod.make_flags(context_key=77))
    c = ctx.dit_forward(lat, cx, sig, mask, fhw, ctxmod.make_flags(context_key=77))    # served from the cache
    np.testing.assert_array_equal(a, b)
    np.testing.assert_array_equal(b, c)
    ctx.close()


@pytest.mark.parametrize("guided", [False, True])
def test_denoise_loop_matches_oracle(guided):
    ocfg, pcfg = small_dit_config(3, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=6)
    fhw, S = (2, 4, 6), 40
    g = torch.Generator().manual_seed(21)
    noise = torch.randn(1, 128, *fhw, generator=g)
    _, cx, mask = _inputs(ocfg, fhw, S, 17)
    _, ncx, _ = _inputs(ocfg, fhw, S, 18)
    sigmas = O.set_timesteps(4, False, 48) if guided else O.set_timesteps(8, True, 48)[4:]
    kw = dict(neg_context=ncx.float(), cfg_scale=4.0, phi=0.7, stg_scale=0.5, stg_blocks=(1,), ge_gamma=0.1) if guided else {}
    ref = O.denoise_loop(w, ocfg, noise, cx.float(), None, sigmas, **kw)
    ctx.denoise_begin(noise[0].numpy(), fhw, sigmas[0], cx, None, ncx if guided else None, None)
    for i in range(len(sigmas) - 1):
        ctx.denoise_step(sigmas[i], sigmas[i + 1], i, cfg_scale=4.0 if guided else 1.0, rescale_phi=0.7 if guided else 0.0,
                         stg_scale=0.5 if guided else 0.0, stg_blocks=(1,) if guided else (), ge_gamma=0.1 if guided else 0.0)
    out = ctx.denoise_get_latent()
    err = rel_l2(out, ref[0])
    assert err <= 2e-2, err      # several guided steps compound the per-step 1e-2 velocity tolerance
    ctx.close()


def test_stg_prefix_sharing_is_bit_identical():
    """resuming the STG pass from the conditional pass's stream at the first perturbed block changes nothing."""
    ocfg, pcfg = small_dit_config(4, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=8)
    fhw, S = (2, 4, 6), 40
    g = torch.Generator().manual_seed(23)
    noise = torch.randn(1, 128, *fhw, generator=g)
    _, cx, _ = _inputs(ocfg, fhw, S, 27)
    _, ncx, _ = _inputs(ocfg, fhw, S, 28)
    sigmas = O.set_timesteps(3, False, 48)
    outs = []
    for share in (True, False):
        ctx.denoise_begin(noise[0].numpy(), fhw, sigmas[0], cx, None, ncx, None)
        for i in range(len(sigmas) - 1):
            ctx.denoise_step(sigmas[i], sigmas[i + 1], i, cfg_scale=3.0, stg_scale=0.5, stg_blocks=(2,), share_stg_prefix=share)
        outs.append(ctx.denoise_get_latent())
    np.testing.assert_array_equal(outs[0], outs[1])
    ref = O.denoise_loop(w, ocfg, noise, cx.float(), None, sigmas, neg_context=ncx.float(), cfg_scale=3.0, stg_scale=0.5, stg_blocks=(2,))
    assert rel_l2(outs[0], ref[0]) <= 2e-2
    ctx.close()


def test_per_token_timesteps_and_i2v_loop():
    """LTXTransformer with timesteps [B, N] and the image-conditioned denoise() loop (frame 0 clean, slice Euler)."""
    ocfg, pcfg = small_dit_config(2, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=9)
    fhw, S = (3, 4, 6), 40
    N = 72
    lat, cx, _ = _inputs(ocfg, fhw, S, 31)
    g = torch.Generator().manual_seed(33)
    ts = torch.rand(1, N, generator=g)
    ts[:, :24] = 0.0
    ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), ts, None, fhw)
    out = ctx.dit_forward(lat, cx, ts.numpy(), None, fhw)
    assert rel_l2(out, ref) <= TOL
    # a constant per-token timestep equals the per-batch path
    same = ctx.dit_forward(lat, cx, np.full((1, N), 0.4, dtype=np.float32), None, fhw)
    base = ctx.dit_forward(lat, cx, np.array([0.4], dtype=np.float32), None, fhw)
    assert rel_l2(same, base) <= 3e-3      # bf16 tensor-core timestep MLP vs the fp32 GEMV path
    # I2V loop: frame 0 of the latent stays exactly what it was
    noise = torch.randn(1, 128, *fhw, generator=g)
    sigmas = O.set_timesteps(8, True, N)[4:]
    init = noise * sigmas[0]
    init[:, :, 0] = torch.randn(128, 4, 6, generator=g) * 0.5          # "encoded image" in frame 0
    ref_l = O.denoise_loop(w, ocfg, noise, cx.float(), None, sigmas, frame0_conditioned=True, init_latent=init)
    # denoise_begin scales its input by sigma0, so hand it init / sigma0
    ctx.denoise_begin((init[0] / sigmas[0]).numpy(), fhw, sigmas[0], cx, None)
    for i in range(len(sigmas) - 1):
        ctx.denoise_step(sigmas[i], sigmas[i + 1], i, i2v_frame0_conditioned=True)
    out_l = ctx.denoise_get_latent()
    assert np.allclose(out_l[:, 0], init[0, :, 0].numpy(), atol=1e-6)
    assert rel_l2(out_l, ref_l[0]) <= 2e-2
    ctx.close()


# ---------------------------------------------------------------------------------------------------------------------
# fp32 mode (ltx_set_precision(ctx, 32)): fp32 weights and activations, split-bf16 tensor-core GEMMs, fp32 attention.
# north_star tolerance: per-step velocity rel-L2 <= 1e-4 against the fp32 graph; the oracle runs in fp64 here so that
# its own rounding does not enter the comparison.
# ---------------------------------------------------------------------------------------------------------------------
TOL_F32 = 1e-4


@pytest.mark.parametrize("layers,heads,fhw,S,B,mask_prefix", [
    (2, 2, (2, 4, 6), 40, 1, 7), (3, 4, (2, 5, 7), 72, 2, 3),
    (1, 32, (2, 16, 16), 128, 1, 0),   # BASELINE config 0: one LTX-2 block, random-init fp32, latent 512x512x9 + 128 text tokens
])
def test_dit_forward_fp32_mode(layers, heads, fhw, S, B, mask_prefix):
    ocfg, pcfg = small_dit_config(layers, heads)
    w = O.make_dit_weights(ocfg, seed=layers * 7 + heads, bf16=False)
    ctx = product().LtxContext(pcfg, 0)
    ctx.set_precision(32)
    ctx.load_weights(w)
    ctx.finalize_weights()
    g = torch.Generator().manual_seed(99)
    N = fhw[0] * fhw[1] * fhw[2]
    lat = torch.randn(B, N, ocfg.in_channels, generator=g)
    cx = torch.randn(B, S, ocfg.caption_channels, generator=g)
    cx = cx / cx.pow(2).mean(-1, keepdim=True).sqrt()
    mask = torch.ones(B, S, dtype=torch.int32)
    if mask_prefix:
        mask[:, :mask_prefix] = 0
    sig = torch.tensor([0.7, 0.3][:B])
    ref = O.dit_forward(w, ocfg, lat.double(), cx.double(), sig, mask if mask_prefix else None, fhw,
                        dtype=torch.float64, mlx_bf16=False)
    out = ctx.dit_forward(lat, cx, sig.numpy(), mask if mask_prefix else None, fhw)
    assert np.isfinite(out).all()
    err = rel_l2(out, ref)
    assert err <= TOL_F32, err
    # and the bf16-mode tolerance is not what makes this pass: the same weights rounded to bf16 differ by >> 1e-4
    ctx.close()


def test_fp32_mode_rejects_late_switch_and_quantisation():
    ocfg, pcfg = small_dit_config(1, 1)
    w = O.make_dit_weights(ocfg, seed=3, bf16=False)
    ctx = product().LtxContext(pcfg, 0)
    ctx.load_weights(w)
    with pytest.raises(product().LtxError):
        ctx.set_precision(32)             # too late: the DiT matrices were already stored as bf16
    ctx.close()
    ctx = product().LtxContext(pcfg, 0)
    ctx.set_precision(32)
    ctx.load_weights(w)
    with pytest.raises(product().LtxError):
        ctx.finalize_weights(quant_bits=8)
    ctx.close()


def test_lora_fuse_matches_merged_weights():
    """ltx_fuse_lora (LoRA/LoRAAdapter.swift:64-166): forward with fused factors == oracle forward on W + scale * up @ down, for
    bf16 and for weights quantised after the merge; the key mapper follows LoRAKeyMapper.loraKeyToModelKey."""
    ctxmod = product()
    assert ctxmod.map_weight_key(6, "diffusion_model.transformer_blocks.0.attn1.to_out.0") == "transformer_blocks.0.attn1.to_out.weight"
    assert ctxmod.map_weight_key(6, "diffusion_model.transformer_blocks.3.ff.net.0.proj") == "transformer_blocks.3.ff.project_in.proj.weight"
    assert ctxmod.map_weight_key(6, "transformer_blocks.3.ff.net.2") == "transformer_blocks.3.ff.project_out.weight"
    ocfg, pcfg = small_dit_config(2, 2)
    w = O.make_dit_weights(ocfg, 21)
    g = torch.Generator().manual_seed(4)
    targets = ["transformer_blocks.0.attn1.to_q.weight", "transformer_blocks.0.attn1.to_out.weight",
               "transformer_blocks.1.attn2.to_k.weight", "transformer_blocks.1.ff.project_in.proj.weight"]
    rank, scale = 16, 0.7
    merged = dict(w)
    factors = {}
    for k in targets:
        out_f, in_f = w[k].shape
        down = (torch.randn(rank, in_f, generator=g) / in_f ** 0.5).bfloat16()
        up = (torch.randn(out_f, rank, generator=g) * 0.3).bfloat16()
        factors[k] = (down, up)
        merged[k] = O.bf16_round(w[k].float() + scale * (up.float() @ down.float()))
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights(w)
    for k, (down, up) in factors.items():
        ctx.fuse_lora(k, down, up, scale)
    ctx.finalize_weights()
    fhw, S = (2, 4, 6), 40
    lat, cx, _ = _inputs(ocfg, fhw, S, 5, 1, 0)
    sig = torch.tensor([0.6])
    ref = O.dit_forward(merged, ocfg, lat.float(), cx.float(), sig, None, fhw)
    base = O.dit_forward(w, ocfg, lat.float(), cx.float(), sig, None, fhw)
    out = ctx.dit_forward(lat, cx, sig.numpy(), None, fhw)
    assert rel_l2(ref, base) > 5e-2            # the adapters matter
    assert rel_l2(out, ref) <= TOL
    ctx.close()


@pytest.mark.parametrize("guided", [False, True])
def test_captured_step_replay_is_bit_identical(guided):
    """ltx_denoise_step runs eagerly until its launch sequence has been seen once with final buffer sizes, captures it into a
    CUDA graph and replays it afterwards (api.cu: run_graphed).  The replayed loop must equal the eager loop bit for bit, also
    across two sessions on one context (the second session's steps replay the first one's graph with fresh text caches)."""
    ocfg, pcfg = small_dit_config(3, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=16)
    fhw, S = (2, 4, 6), 40
    g = torch.Generator().manual_seed(23)
    noise = torch.randn(1, 128, *fhw, generator=g)
    _, cx, _ = _inputs(ocfg, fhw, S, 27)
    _, cx2, _ = _inputs(ocfg, fhw, S, 28)
    _, ncx, _ = _inputs(ocfg, fhw, S, 29)
    sigmas = O.set_timesteps(6, False, 48)

    def loop(context):
        ctx.denoise_begin(noise[0].numpy(), fhw, sigmas[0], context, None, ncx if guided else None, None)
        for i in range(len(sigmas) - 1):
            ctx.denoise_step(sigmas[i], sigmas[i + 1], i, cfg_scale=3.0 if guided else 1.0, rescale_phi=0.5 if guided else 0.0,
                             stg_scale=0.4 if guided else 0.0, stg_blocks=(1,) if guided else (), ge_gamma=0.1 if guided else 0.0)
        return ctx.denoise_get_latent()

    ctx.set_graphs(False)
    eager1, eager2 = loop(cx), loop(cx2)
    assert ctx.graph_stats() == (0, 0)
    ctx.set_graphs(True)
    graphed1, graphed2 = loop(cx), loop(cx2)
    captures, replays = ctx.graph_stats()
    assert captures >= 1 and replays >= 4, (captures, replays)
    np.testing.assert_array_equal(graphed1, eager1)
    np.testing.assert_array_equal(graphed2, eager2)
    assert rel_l2(eager1, eager2) > 1e-3                 # the two prompts really differ
    ctx.close()


def test_captured_forward_at_the_host_seam():
    """ltx_dit_forward with a context_key: from the third call on the device part is a graph replay; results stay identical and
    follow the inputs (latent and timestep are re-uploaded into the same staging buffers before every replay)."""
    ocfg, pcfg = small_dit_config(2, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=18)
    fhw, S = (2, 4, 6), 40
    lat, cx, mask = _inputs(ocfg, fhw, S, 31, mask_prefix=4)
    lat2, _, _ = _inputs(ocfg, fhw, S, 32)
    ctxmod = product()
    fl = ctxmod.make_flags(context_key=ctx.new_context_key())
    ref_a = ctx.dit_forward(lat, cx, np.array([0.6], dtype=np.float32), mask, fhw)          # un-keyed: always eager
    ref_b = ctx.dit_forward(lat2, cx, np.array([0.3], dtype=np.float32), mask, fhw)
    outs = [ctx.dit_forward(lat, cx, np.array([0.6], dtype=np.float32), mask, fhw, fl) for _ in range(4)]
    out_b = ctx.dit_forward(lat2, cx, np.array([0.3], dtype=np.float32), mask, fhw, fl)
    captures, replays = ctx.graph_stats()
    assert captures == 1 and replays >= 2, (captures, replays)
    for o in outs:
        np.testing.assert_array_equal(o, ref_a)
    np.testing.assert_array_equal(out_b, ref_b)
    ctx.close()


@pytest.mark.parametrize("i2v,masked,stg", [(False, False, True), (True, False, False), (False, True, True), (True, True, True)])
def test_batched_guidance_equals_separate_passes(i2v, masked, stg):
    """ltx_denoise_step runs the conditional and unconditional passes as ONE B = 2 forward on a single GPU (denoise() batches
    them the same way, P/LTXPipeline.swift:2234-2269; generateVideo issues two B = 1 calls, :829-848).  Per batch row the
    arithmetic is that of the separate passes: the two loops agree to bf16 round-off (bit-identical in practice), with masks
    on one or both prompts, per-token timesteps (image-to-video) and the STG pass resumed from batch row 0 of the shared prefix."""
    ocfg, pcfg = small_dit_config(3, 2)
    ctx, w = make_ctx_with_dit(ocfg, pcfg, seed=26)
    fhw, S = (3, 4, 8), 40          # 96 tokens: a multiple of 8, so the batched pass also takes the fused q|k|v projection
    g = torch.Generator().manual_seed(41)
    noise = torch.randn(1, 128, *fhw, generator=g)
    _, cx, mk = _inputs(ocfg, fhw, S, 43, mask_prefix=6)
    _, ncx, _ = _inputs(ocfg, fhw, S, 44)
    sigmas = O.set_timesteps(4, False, 96)

    def loop(batched):
        ctx.denoise_begin(noise[0].numpy(), fhw, sigmas[0], cx, mk if masked else None, ncx, None)
        for i in range(len(sigmas) - 1):
            ctx.denoise_step(sigmas[i], sigmas[i + 1], i, cfg_scale=3.0, rescale_phi=0.5, stg_scale=0.4 if stg else 0.0,
                             stg_blocks=(1,) if stg else (), ge_gamma=0.1, i2v_frame0_conditioned=i2v, batched_cfg=batched)
        return ctx.denoise_get_latent()

    sep = loop(False)
    bat = loop(True)
    assert np.isfinite(bat).all()
    assert rel_l2(bat, sep) <= 1e-5, rel_l2(bat, sep)
    if not i2v:
        kw = dict(neg_context=ncx.float(), cfg_scale=3.0, phi=0.5, stg_scale=0.4 if stg else 0.0, stg_blocks=(1,), ge_gamma=0.1)
        ref = O.denoise_loop(w, ocfg, noise, cx.float(), mk if masked else None, sigmas, **kw)
        assert rel_l2(bat, ref[0]) <= 2e-2
    ctx.close()
