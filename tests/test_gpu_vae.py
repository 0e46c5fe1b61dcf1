"""Video-VAE decode parity against the CPU oracle: PSNR >= 40 dB on [0,1] frames (bf16 activations into the convs)."""
import numpy as np
import pytest
import torch

from helpers import O, make_ctx_with_vae, small_vae_config

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("base,blocks,fhw", [(512, 1, (2, 3, 4)), (512, 2, (3, 4, 6)), (1024, 1, (1, 2, 3))])
def test_vae_decode_matches_oracle(base, blocks, fhw):
    ocfg, pcfg = small_vae_config(base, blocks)
    ctx, w = make_ctx_with_vae(ocfg, pcfg, seed=base + blocks)
    g = torch.Generator().manual_seed(31)
    z = torch.randn(1, 128, *fhw, generator=g)
    ref = O.decode_video(w, ocfg, z)
    out = ctx.vae_decode(z[0].numpy())
    assert out.shape == tuple(ref.shape) == (8 * (fhw[0] - 1) + 1, 32 * fhw[1], 32 * fhw[2], 3)
    assert np.isfinite(out).all() and out.min() >= 0 and out.max() <= 1
    p = O.psnr(torch.from_numpy(out), ref)
    assert p >= 40.0, p
    ctx.close()


def test_vae_decode_timestep_conditioned():
    """decodeVideo(timestep: 0.05) -- noise injection + time-embedded scale/shift tables (V/VideoDecoder.swift:368-375,
    100-114, 419-436); the noise is passed in as data."""
    ocfg, pcfg = small_vae_config(512, 1)
    ctx, w = make_ctx_with_vae(ocfg, pcfg, seed=41)
    g = torch.Generator().manual_seed(43)
    z = torch.randn(1, 128, 2, 3, 4, generator=g)
    nz = torch.randn(1, 128, 2, 3, 4, generator=g)
    ref = O.decode_video(w, ocfg, z, timestep=0.05, decode_noise=nz)
    out = ctx.vae_decode(z[0].numpy(), timestep=0.05, decode_noise=nz[0].numpy())
    assert O.psnr(torch.from_numpy(out), ref) >= 40.0
    plain = ctx.vae_decode(z[0].numpy())
    assert O.psnr(torch.from_numpy(plain), ref) < 35.0      # the conditioning really changes the frames
    ctx.close()


def test_vae_decode_causal():
    """causal: true -- two leading frame copies instead of one leading + one trailing (V/VideoConvolution.swift:281-294); covers the
    fused conv1 -> conv2 hand-over and the halo fill with the causal frame offset."""
    ocfg, pcfg = small_vae_config(512, 2)
    ocfg.causal = True
    ctx, w = make_ctx_with_vae(ocfg, pcfg, seed=51)
    z = torch.randn(1, 128, 3, 3, 4, generator=torch.Generator().manual_seed(53))
    ref = O.decode_video(w, ocfg, z)
    out = ctx.vae_decode(z[0].numpy(), causal=True)
    assert O.psnr(torch.from_numpy(out), ref) >= 40.0
    noncausal = ctx.vae_decode(z[0].numpy(), causal=False)
    assert O.psnr(torch.from_numpy(noncausal), ref) < 35.0
    ctx.close()
