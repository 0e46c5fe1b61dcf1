"""GPU parity of the dual audio / video transformer against the CPU oracle: per-stream velocity
rel-L2 <= 1e-2 (the north_star's bf16 bound), repeated calls bit-identical, cached text K/V == uncached."""
import numpy as np
import pytest
import torch

from helpers import O, product, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _setup(layers, heads, aheads, seed, caption=192):
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=layers, num_heads=heads, head_dim=128, caption_channels=caption)
    av = O.AVConfig(audio_heads=aheads)
    pcfg = ctxmod.LTXTransformerConfig(num_layers=layers, num_attention_heads=heads, caption_channels=caption,
                                       audio_num_attention_heads=aheads)
    w = O.make_av_weights(ocfg, av, seed)
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights(w)
    ctx.finalize_weights()
    return ocfg, av, w, ctx


def _inputs(fhw, Ta, S, caption, seed, mask_prefix=0):
    g = torch.Generator().manual_seed(seed)
    N = fhw[0] * fhw[1] * fhw[2]
    vl = torch.randn(1, N, 128, generator=g).bfloat16()
    al = torch.randn(1, Ta, 128, generator=g).bfloat16()

    def text():
        t = torch.randn(1, S, caption, generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    vc, ac = text(), text()
    mask = None
    if mask_prefix:
        mask = torch.ones(1, S, dtype=torch.int32)
        mask[:, :mask_prefix] = 0
    return vl, al, vc, ac, mask


@pytest.mark.parametrize("layers,heads,aheads,fhw,Ta,S,mask_prefix", [
    (2, 2, 2, (2, 4, 6), 11, 40, 0),
    (2, 2, 4, (3, 4, 4), 26, 24, 5),
    (1, 4, 2, (2, 8, 10), 130, 150, 0),     # more than one key tile on every attention, ragged tiles
])
def test_av_forward_matches_oracle(layers, heads, aheads, fhw, Ta, S, mask_prefix):
    ocfg, av, w, ctx = _setup(layers, heads, aheads, seed=layers * 100 + heads * 10 + aheads)
    vl, al, vc, ac, mask = _inputs(fhw, Ta, S, 192, 3, mask_prefix)
    rv, ra = O.av_dit_forward(w, ocfg, av, vl.float(), al.float(), vc.float(), ac.float(), torch.tensor([0.7]), torch.tensor([0.55]),
                              mask, mask, fhw, Ta)
    ov, oa = ctx.av_forward(vl, al, vc, ac, 0.7, 0.55, fhw, mask, mask)
    assert np.isfinite(ov).all() and np.isfinite(oa).all()
    assert rel_l2(ov, rv) <= TOL, rel_l2(ov, rv)
    assert rel_l2(oa, ra) <= TOL, rel_l2(oa, ra)
    # deterministic, and the cached text K/V path gives the same numbers
    ov2, oa2 = ctx.av_forward(vl, al, vc, ac, 0.7, 0.55, fhw, mask, mask, context_key=7)
    ov3, oa3 = ctx.av_forward(vl, al, vc, ac, 0.7, 0.55, fhw, mask, mask, context_key=7)
    assert np.array_equal(ov2, ov) and np.array_equal(oa2, oa)
    assert np.array_equal(ov3, ov) and np.array_equal(oa3, oa)
    ctx.close()


@pytest.mark.parametrize("layers,heads,aheads,fhw,Ta,S", [
    (2, 2, 2, (3, 4, 6), 11, 40),
    (1, 4, 2, (2, 8, 10), 130, 150),
])
def test_av_forward_per_token_sigmas_match_oracle(layers, heads, aheads, fhw, Ta, S):
    """videoTimesteps [1, N] (image-to-video: sigma * (1 - conditioningMask), P/LTXPipeline.swift:1293-1298): the main video
    AdaLN-single, the two cross-modal video embedders and the head's embedded timestep are per token
    (T/LTX2Transformer.swift:273-298).  Frame 0 at sigma 0, a ramp elsewhere so every token has its own modulation row."""
    ocfg, av, w, ctx = _setup(layers, heads, aheads, seed=layers * 100 + heads * 10 + aheads + 1)
    vl, al, vc, ac, _ = _inputs(fhw, Ta, S, 192, 4)
    N = fhw[0] * fhw[1] * fhw[2]
    ts = torch.linspace(0.35, 0.85, N).view(1, N)
    ts[:, :fhw[1] * fhw[2]] = 0.0
    rv, ra = O.av_dit_forward(w, ocfg, av, vl.float(), al.float(), vc.float(), ac.float(), ts, torch.tensor([0.55]), None, None,
                              fhw, Ta)
    ov, oa = ctx.av_forward(vl, al, vc, ac, ts.numpy(), 0.55, fhw)
    assert np.isfinite(ov).all() and np.isfinite(oa).all()
    assert rel_l2(ov, rv) <= TOL, rel_l2(ov, rv)
    assert rel_l2(oa, ra) <= TOL, rel_l2(oa, ra)
    ov2, oa2 = ctx.av_forward(vl, al, vc, ac, ts.numpy(), 0.55, fhw)
    assert np.array_equal(ov2, ov) and np.array_equal(oa2, oa)
    # a constant per-token vector is the scalar call up to the embedders' arithmetic (tensor-core GEMM vs fp32 GEMV)
    us, ua = ctx.av_forward(vl, al, vc, ac, 0.7, 0.55, fhw)
    ut, uat = ctx.av_forward(vl, al, vc, ac, np.full((1, N), 0.7, dtype=np.float32), 0.55, fhw)
    assert rel_l2(ut, us) <= 5e-3 and rel_l2(uat, ua) <= 5e-3, (rel_l2(ut, us), rel_l2(uat, ua))
    # the per-token values matter: the scalar result is far from the image-conditioned one
    assert rel_l2(us, ov) > 1e-2
    ctx.close()


def test_av_streams_are_coupled_and_video_only_model_still_works():
    """Changing the audio latent changes the video velocity (the a2v attention is live), and the same context still serves
    the video-only forward (LTXTransformer) from the shared video weights."""
    ocfg, av, w, ctx = _setup(2, 2, 2, seed=5)
    fhw, Ta, S = (2, 4, 6), 11, 40
    vl, al, vc, ac, _ = _inputs(fhw, Ta, S, 192, 9)
    ov, _ = ctx.av_forward(vl, al, vc, ac, 0.6, 0.6, fhw)
    ov_b, _ = ctx.av_forward(vl, (al.float() * 1.5).bfloat16(), vc, ac, 0.6, 0.6, fhw)
    assert rel_l2(ov_b, ov) > 1e-3
    wv = {k: v for k, v in w.items() if k in O.make_dit_weights(ocfg, 5)}
    ref = O.dit_forward(wv, ocfg, vl.float(), vc.float(), torch.tensor([0.6]), None, fhw)
    out = ctx.dit_forward(vl, vc, np.array([0.6], dtype=np.float32), None, fhw)
    assert rel_l2(out, ref) <= TOL
    ctx.close()


def test_av_random_init_full_width_block():
    """One block at the real widths (D = 4096 / 32 heads, Da = 2048 / 32 x 64) from the on-device random init: finite output,
    deterministic -- exercises the D-specialised kernels on the dual path."""
    ctxmod = product()
    pcfg = ctxmod.LTXTransformerConfig(num_layers=1)
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.init_random_weights(17, seed=3)
    ctx.finalize_weights()
    fhw, Ta, S = (2, 8, 8), 50, 128
    vl, al, vc, ac, _ = _inputs(fhw, Ta, S, 3840, 11)
    ov, oa = ctx.av_forward(vl, al, vc, ac, 0.8, 0.8, fhw)
    ov2, oa2 = ctx.av_forward(vl, al, vc, ac, 0.8, 0.8, fhw)
    assert np.isfinite(ov).all() and np.isfinite(oa).all() and ov.std() > 0.05 and oa.std() > 0.05
    assert np.array_equal(ov, ov2) and np.array_equal(oa, oa2)
    # per-token sigmas on the D = 4096 / 2048 row kernels: a constant vector reproduces the scalar call, a frame-0 mask does not
    N = fhw[0] * fhw[1] * fhw[2]
    ts = np.full((1, N), 0.8, dtype=np.float32)
    pv, pa = ctx.av_forward(vl, al, vc, ac, ts, 0.8, fhw)
    assert rel_l2(pv, ov) <= 5e-3 and rel_l2(pa, oa) <= 5e-3, (rel_l2(pv, ov), rel_l2(pa, oa))
    ts[:, :fhw[1] * fhw[2]] = 0.0
    qv, qa = ctx.av_forward(vl, al, vc, ac, ts, 0.8, fhw)
    qv2, _ = ctx.av_forward(vl, al, vc, ac, ts, 0.8, fhw)
    assert np.isfinite(qv).all() and np.isfinite(qa).all() and np.array_equal(qv, qv2)
    assert rel_l2(qv[:, :fhw[1] * fhw[2]], ov[:, :fhw[1] * fhw[2]]) > 1e-2
    ctx.close()


@pytest.mark.parametrize("bits,tol", [(8, 3e-2), (4, 0.4)])
def test_av_quantised_weights(bits, tol):
    """quantize(model: ltx2, groupSize: 64, bits:) (P/LTXPipeline.swift:491): every Linear of both streams through the
    dequant-fused GEMM.  As for the video model (test_gpu_quant.py) the MLX rounding rule is not in the reference tree, so the
    check is the drift of the quantised model against the bf16 model, plus determinism; per-token sigmas ride along."""
    ocfg, av, w, ctx16 = _setup(2, 2, 2, seed=21)
    ctxmod = product()
    pcfg = ctxmod.LTXTransformerConfig(num_layers=2, num_attention_heads=2, caption_channels=192, audio_num_attention_heads=2)
    ctxq = ctxmod.LtxContext(pcfg, 0)
    ctxq.load_weights(w)
    ctxq.finalize_weights(quant_bits=bits)
    for fhw, Ta, S in [((2, 4, 6), 11, 40), ((3, 10, 10), 30, 24)]:      # second case: N = 300 > 256 takes the panel path
        vl, al, vc, ac, _ = _inputs(fhw, Ta, S, 192, 13)
        N = fhw[0] * fhw[1] * fhw[2]
        for vs in (0.7, np.linspace(0.3, 0.8, N, dtype=np.float32).reshape(1, N)):
            a_v, a_a = ctx16.av_forward(vl, al, vc, ac, vs, 0.55, fhw)
            q_v, q_a = ctxq.av_forward(vl, al, vc, ac, vs, 0.55, fhw)
            q_v2, q_a2 = ctxq.av_forward(vl, al, vc, ac, vs, 0.55, fhw)
            assert np.isfinite(q_v).all() and np.isfinite(q_a).all()
            ev, ea = rel_l2(q_v, a_v), rel_l2(q_a, a_a)
            assert ev <= tol and ea <= tol, (ev, ea)
            assert ev > 1e-5 and ea > 1e-5                      # the quantised path really ran
            assert np.array_equal(q_v, q_v2) and np.array_equal(q_a, q_a2)
    ctx16.close()
    ctxq.close()


def test_av_restrictions_fail_loudly():
    """Calling the dual model without the audio tensors is a weight error (2 / 4), not a silent video-only fallback."""
    from ltx_video_swift_mlx_b200._lib import LtxError
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=1, num_heads=2, head_dim=128, caption_channels=192)
    av = O.AVConfig(audio_heads=2)
    pcfg = ctxmod.LTXTransformerConfig(num_layers=1, num_attention_heads=2, caption_channels=192, audio_num_attention_heads=2)
    w = O.make_av_weights(ocfg, av, 3)
    ctx = ctxmod.LtxContext(pcfg, 0)
    ctx.load_weights({k: v for k, v in w.items() if k in O.make_dit_weights(ocfg, 3)})   # video-only weights
    ctx.finalize_weights()
    vl, al, vc, ac, _ = _inputs((2, 4, 6), 11, 24, 192, 1)
    with pytest.raises(LtxError) as e:
        ctx.av_forward(vl, al, vc, ac, 0.5, 0.5, (2, 4, 6))
    assert e.value.code in (2, 4)
    ctx.close()


def test_av_denoise_loop_matches_oracle():
    """generateVideoWithAudio's step loop (P/LTXPipeline.swift:1277-1404) at the host seams, with CFG + rescale: both final
    latents within the multi-step bf16 bound of the oracle loop."""
    from ltx_video_swift_mlx_b200.pipeline import denoise_av_host_seam
    ocfg, av, w, ctx = _setup(2, 2, 2, seed=77)
    fhw, Ta, S = (2, 4, 6), 11, 40
    g = torch.Generator().manual_seed(5)
    vn, an = torch.randn(1, 128, *fhw, generator=g), torch.randn(1, Ta, 128, generator=g)

    def text():
        t = torch.randn(1, S, 192, generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    vc, ac, nvc, nac = text(), text(), text(), text()
    sig = O.set_timesteps(8, True, 48)[4:8]
    rv, ra = O.av_denoise_loop(w, ocfg, av, vn, an, vc.float(), ac.float(), None, sig, nvc.float(), nac.float(), None, cfg_scale=3.0, phi=0.5)
    ov, oa = denoise_av_host_seam(ctx, vn.numpy(), an.numpy(), vc, ac, None, sig, nvc, nac, None, cfg_scale=3.0, guidance_rescale=0.5)
    assert rel_l2(ov, rv) <= 2e-2 and rel_l2(oa, ra) <= 2e-2, (rel_l2(ov, rv), rel_l2(oa, ra))
    ctx.close()


def test_av_image_conditioned_denoise_loop_matches_oracle():
    """The image-to-video branch of generateVideoWithAudio (P/LTXPipeline.swift:1262-1298, 1381-1391): frame 0 = image latent
    (+ per-step injected noise passed in as data), per-token video timesteps, Euler on frames 1+ only."""
    from ltx_video_swift_mlx_b200.pipeline import denoise_av_host_seam
    ocfg, av, w, ctx = _setup(2, 2, 2, seed=78)
    fhw, Ta, S = (3, 4, 6), 11, 40
    g = torch.Generator().manual_seed(6)
    vn, an = torch.randn(1, 128, *fhw, generator=g), torch.randn(1, Ta, 128, generator=g)
    img = torch.randn(1, 128, 1, fhw[1], fhw[2], generator=g)

    def text():
        t = torch.randn(1, S, 192, generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    vc, ac = text(), text()
    sig = O.set_timesteps(8, True, 72)[3:7]
    inj = [torch.randn(1, 128, 1, fhw[1], fhw[2], generator=g) for _ in range(len(sig) - 1)]
    rv, ra = O.av_denoise_loop(w, ocfg, av, vn, an, vc.float(), ac.float(), None, sig, image_latent=img, inject_noise=inj,
                               image_cond_noise_scale=0.15)
    ov, oa = denoise_av_host_seam(ctx, vn.numpy(), an.numpy(), vc, ac, None, sig, image_latent=img.numpy(),
                                  inject_noise=[t.numpy() for t in inj], image_cond_noise_scale=0.15)
    assert rel_l2(ov, rv) <= 2e-2 and rel_l2(oa, ra) <= 2e-2, (rel_l2(ov, rv), rel_l2(oa, ra))
    # frame 0 left the loop as the (last) conditioned frame, untouched by the Euler update
    sg_last = float(sig[-2])
    expect0 = img + 0.15 * inj[-1] * (sg_last * sg_last) if sg_last > 0 else img
    assert np.allclose(ov[:, :, 0:1], expect0.numpy(), atol=1e-6)
    ctx.close()


@pytest.mark.parametrize("i2v,cfg", [(False, 3.0), (True, 1.0), (True, 2.5)])
def test_av_resident_session_matches_host_seam_loop_and_oracle(i2v, cfg):
    """ltx_av_denoise_begin / _step (latents resident in HBM) against the host-seam loop -- the same kernels behind both, so the
    video latent is bit-identical and the audio latent equal to fp32 rounding of (sigma' - sigma) -- and against the oracle."""
    from ltx_video_swift_mlx_b200.pipeline import denoise_av_host_seam, denoise_av_resident
    from ltx_video_swift_mlx_b200._lib import LtxError
    ocfg, av, w, ctx = _setup(2, 2, 2, seed=79)
    fhw, Ta, S = (3, 4, 6), 11, 40
    g = torch.Generator().manual_seed(8)
    vn, an = torch.randn(1, 128, *fhw, generator=g), torch.randn(1, Ta, 128, generator=g)
    img = torch.randn(1, 128, 1, fhw[1], fhw[2], generator=g) if i2v else None

    def text():
        t = torch.randn(1, S, 192, generator=g)
        return (t / t.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    vc, ac, nvc, nac = text(), text(), text(), text()
    mask = torch.ones(1, S, dtype=torch.int32)
    mask[:, :3] = 0
    use_neg = cfg > 1.0
    sig = O.set_timesteps(8, True, 72)[3:7]
    inj = [torch.randn(1, 128, 1, fhw[1], fhw[2], generator=g) for _ in range(len(sig) - 1)] if i2v else None
    kw = dict(cfg_scale=cfg, guidance_rescale=0.4 if use_neg else 0.0,
              image_latent=None if img is None else img.numpy(), inject_noise=None if inj is None else [t.numpy() for t in inj],
              image_cond_noise_scale=0.15 if i2v else 0.0)
    negs = (nvc, nac, mask) if use_neg else (None, None, None)
    hv, ha = denoise_av_host_seam(ctx, vn.numpy(), an.numpy(), vc, ac, mask, sig, *negs, **kw)
    rv, ra = denoise_av_resident(ctx, vn.numpy(), an.numpy(), vc, ac, mask, sig, *negs, **kw)
    assert np.array_equal(rv, hv)
    assert np.allclose(ra, ha, rtol=0, atol=1e-5)
    ov, oa = O.av_denoise_loop(w, ocfg, av, vn, an, vc.float(), ac.float(), mask, sig, nvc.float() if use_neg else None,
                               nac.float() if use_neg else None, mask if use_neg else None, cfg_scale=cfg,
                               phi=0.4 if use_neg else 0.0, image_latent=img, inject_noise=inj,
                               image_cond_noise_scale=0.15 if i2v else 0.0)
    assert rel_l2(rv, ov) <= 2e-2 and rel_l2(ra, oa) <= 2e-2, (rel_l2(rv, ov), rel_l2(ra, oa))
    with pytest.raises(LtxError) as e:                      # no STG / GE in this loop: refused, not ignored
        p = ctx.lib  # noqa: F841
        from ltx_video_swift_mlx_b200._lib import LtxStepParams
        import ctypes as C
        sp = LtxStepParams()
        sp.sigma, sp.sigma_next, sp.cfg_scale, sp.stg_scale = 0.5, 0.4, 1.0, 0.5
        ctx._check(ctx.lib.ltx_av_denoise_step(ctx.handle, C.byref(sp)))
    assert e.value.code == 5
    ctx.close()
