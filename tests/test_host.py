"""Host-side mirror of the reference interface (no GPU): scheduler, shapes, validation, flag marshalling."""
import numpy as np
import pytest

from helpers import O, product

product()
from ltx_video_swift_mlx_b200 import latent_utils, scheduler  # noqa: E402
from ltx_video_swift_mlx_b200._lib import LtxError  # noqa: E402
from ltx_video_swift_mlx_b200.context import make_flags  # noqa: E402


@pytest.mark.parametrize("tokens", [None, 512, 1536, 4096, 6144, 50688])
def test_scheduler_mirror_matches_oracle_distilled(tokens):
    s = scheduler.LTXScheduler()
    got = s.set_timesteps(8, distilled=True, latent_token_count=tokens)
    assert np.allclose(got, O.set_timesteps(8, True, tokens), rtol=0, atol=1e-6)   # float32 evaluation order / expf differ by a few ulp between libms
    assert s.total_steps == 8 and s.initial_sigma == 1.0


@pytest.mark.parametrize("steps,tokens", [(40, 1536), (20, 512), (8, None), (2, 1536)])
def test_scheduler_mirror_matches_oracle_dev(steps, tokens):
    got = scheduler.LTXScheduler().set_timesteps(steps, False, tokens)
    assert np.allclose(got, O.set_timesteps(steps, False, tokens), rtol=0, atol=1e-6)
    assert got[0] == 1.0 and got[-1] == 0.0 and all(a > b for a, b in zip(got, got[1:]))


def test_custom_sigmas_and_constants():
    s = scheduler.LTXScheduler()
    s.set_custom_sigmas(scheduler.STAGE_2_DISTILLED_SIGMA_VALUES[:-1])
    assert s.sigmas == scheduler.STAGE_2_DISTILLED_SIGMA_VALUES
    assert scheduler.get_sigma_schedule(8, True) == scheduler.DISTILLED_SIGMA_VALUES


def test_shapes_and_validation():
    sh = latent_utils.VideoLatentShape.from_pixel_dimensions(1, 128, 25, 512, 768)
    assert sh.fhw == (4, 16, 24) and sh.token_count == 1536
    x = np.random.RandomState(0).randn(1, 128, 2, 3, 4).astype(np.float32)
    sh2 = latent_utils.VideoLatentShape(1, 128, 2, 3, 4)
    assert np.array_equal(latent_utils.unpatchify(latent_utils.patchify(x), sh2), x)
    latent_utils.LTXVideoGenerationConfig().validate()
    for bad in (dict(width=770), dict(num_frames=24), dict(num_frames=265), dict(num_steps=0), dict(cfg_scale=0.5)):
        with pytest.raises(LtxError) as e:
            latent_utils.LTXVideoGenerationConfig(**bad).validate()
        assert e.value.code == 1


def test_flag_marshalling():
    f = make_flags(stg_blocks=[29], skip_self_attn=True, cas_blocks=[1, 2], cross_attn_scale=0.5, context_key=9)
    assert f.n_stg_blocks == 1 and f.stg_blocks[0] == 29 and f.skip_self_attn == 1 and f.skip_ff == 0
    assert f.n_cas_blocks == 2 and f.cas_blocks[1] == 2 and abs(f.cross_attn_scale - 0.5) < 1e-7 and f.context_key == 9


class _OracleBackedContext:
    """Stand-in for LtxContext on the GPU-less host: the three seams answer from the CPU oracle, so that the HOST logic of
    pipeline.py (loop order, patchify / unpatchify, flag handling, frame-0 handling, CFG on the audio stream) can be checked
    without a device.  Test scaffolding only -- the product has no such path."""

    def __init__(self, w, cfg, av=None):
        import torch
        self.torch, self.w, self.cfg, self.av = torch, w, cfg, av
        self.config = type("C", (), {"out_channels": 128, "in_channels": 128, "audio_in_channels": 128})()
        self.calls = []

    def dit_forward(self, latent, context, timesteps, mask, fhw, flags=None):
        t = self.torch
        stg = [flags.stg_blocks[i] for i in range(flags.n_stg_blocks)] if flags is not None else []
        self.calls.append(("dit", tuple(stg), int(flags.context_key) if flags is not None else 0))
        out = O.dit_forward(self.w, self.cfg, t.as_tensor(np.asarray(latent, dtype=np.float32)), t.as_tensor(np.asarray(context, dtype=np.float32)),
                            t.as_tensor(np.asarray(timesteps, dtype=np.float32)), None if mask is None else t.as_tensor(np.asarray(mask)),
                            tuple(fhw), stg_blocks=stg, skip_self_attn=bool(flags.skip_self_attn) if flags is not None else False)
        return out.numpy()

    def new_context_key(self):
        self.keys = getattr(self, "keys", 0) + 1
        return 100 + self.keys

    def av_forward(self, vl, al, vc, ac, vs, a_s, fhw, vm=None, am=None, context_key=0):
        t = self.torch
        self.calls.append(("av", np.size(vs), context_key))
        f = lambda x: t.as_tensor(np.asarray(x, dtype=np.float32))            # noqa: E731
        vts = f(vs).reshape(1, -1) if np.size(vs) > 1 else f([float(np.asarray(vs).reshape(-1)[0])])
        v, a = O.av_dit_forward(self.w, self.cfg, self.av, f(vl), f(al), f(vc), f(ac), vts, f([a_s]),
                                None if vm is None else t.as_tensor(np.asarray(vm)), None if am is None else t.as_tensor(np.asarray(am)),
                                tuple(fhw), np.asarray(al).shape[1])
        return v.numpy(), a.numpy()

    def guided_euler_step(self, latent, v_cond, v_uncond=None, v_stg=None, v_prev=None, use_prev=False, cfg_scale=1.0,
                          rescale_phi=0.0, stg_scale=0.0, ge_gamma=0.0, sigma=1.0, sigma_next=0.0):
        t = self.torch
        f = lambda x: None if x is None else t.as_tensor(np.asarray(x, dtype=np.float32))   # noqa: E731
        new, v = O.guided_euler_step(f(latent), f(v_cond), f(v_uncond), f(v_stg), f(v_prev) if use_prev else None, cfg_scale,
                                     rescale_phi, stg_scale, ge_gamma, sigma, sigma_next)
        latent[...] = new.numpy()
        if v_prev is not None:
            v_prev[...] = v.numpy()
        return latent


def test_host_seam_loops_follow_the_reference_order():
    """pipeline.denoise_host_seam / denoise_av_host_seam driven through an oracle-backed stand-in equal the oracle's own loops
    (P/LTXPipeline.swift:793-956, 1255-1404): same pass order, same text-cache keys per prompt, STG flags set only around the
    perturbed pass, frame 0 untouched and per-token video sigmas in the image-to-video mode."""
    import torch
    from ltx_video_swift_mlx_b200 import pipeline
    cfg = O.DiTConfig(num_layers=2, num_heads=2, head_dim=128, caption_channels=64)
    w = O.make_dit_weights(cfg, 3)
    g = torch.Generator().manual_seed(4)
    fhw, S = (2, 2, 3), 9
    noise = torch.randn(1, 128, *fhw, generator=g)
    cx, ncx = torch.randn(1, S, 64, generator=g), torch.randn(1, S, 64, generator=g)
    sig = [1.0, 0.8, 0.45, 0.0]
    ctx = _OracleBackedContext(w, cfg)
    out = pipeline.denoise_host_seam(ctx, noise.numpy(), cx.numpy(), None, sig, ncx.numpy(), None, cfg_scale=3.0,
                                     guidance_rescale=0.5, stg_scale=0.4, stg_blocks=(1,), ge_gamma=0.2)
    ref = O.denoise_loop(w, cfg, noise, cx, None, sig, ncx, None, 3.0, 0.5, 0.4, (1,), 0.2)
    assert O.rel_l2(torch.from_numpy(out), ref) < 1e-5
    kp, kn = ctx.calls[0][2], ctx.calls[1][2]
    assert kp != 0 and kn != 0 and kp != kn
    assert ctx.calls[:3] == [("dit", (), kp), ("dit", (), kn), ("dit", (1,), kp)] and len(ctx.calls) == 9
    assert all(c[2] == (kn if i % 3 == 1 else kp) for i, c in enumerate(ctx.calls))     # one key per prompt for the whole loop
    # a second generation on the same context must not reuse the first one's keys (stale text K/V under constant keys)
    pipeline.denoise_host_seam(ctx, noise.numpy(), ncx.numpy(), None, sig[:2])
    assert ctx.calls[-1][2] not in (0, kp, kn)

    av = O.AVConfig(audio_heads=2)
    wav = O.make_av_weights(cfg, av, 5)
    an = torch.randn(1, 5, 128, generator=g)
    acx, nacx = torch.randn(1, S, 64, generator=g), torch.randn(1, S, 64, generator=g)
    img = torch.randn(1, 128, 1, 2, 3, generator=g)
    inj = [torch.randn(1, 128, 1, 2, 3, generator=g) for _ in range(3)]
    for image, cfgs in ((None, 2.0), (img, 1.0), (img, 2.5)):
        ctx = _OracleBackedContext(wav, cfg, av)
        neg = (ncx.numpy(), nacx.numpy(), None) if cfgs > 1 else (None, None, None)
        v, a = pipeline.denoise_av_host_seam(ctx, noise.numpy(), an.numpy(), cx.numpy(), acx.numpy(), None, sig, *neg, cfg_scale=cfgs,
                                             guidance_rescale=0.3 if cfgs > 1 else 0.0,
                                             image_latent=None if image is None else image.numpy(),
                                             inject_noise=None if image is None else [t.numpy() for t in inj],
                                             image_cond_noise_scale=0.1 if image is not None else 0.0)
        rv, ra = O.av_denoise_loop(wav, cfg, av, noise, an, cx, acx, None, sig, ncx if cfgs > 1 else None, nacx if cfgs > 1 else None,
                                   None, cfg_scale=cfgs, phi=0.3 if cfgs > 1 else 0.0, image_latent=image, inject_noise=inj,
                                   image_cond_noise_scale=0.1 if image is not None else 0.0)
        assert O.rel_l2(torch.from_numpy(v), rv) < 1e-5 and O.rel_l2(torch.from_numpy(a), ra) < 1e-5
        n_tok = 12 if image is not None else 1
        assert all(c[0] == "av" and c[1] == n_tok for c in ctx.calls) and len(ctx.calls) == (6 if cfgs > 1 else 3)
