"""The oracle's primitives against INDEPENDENT implementations of the same definitions -- PyTorch's own kernels
(torch.nn.functional), einops patterns and complex arithmetic.  The reference's arithmetic lives in mlx-swift 0.30.6
(MLXFast.rmsNorm, MLXFast.scaledDotProductAttention, MLXNN.geluApproximate, MLX.conv2d ...), which is not in the tree and
cannot run here, so the restatement assumes the standard definitions of those calls; this file pins each assumed
definition to a second implementation that shares no code with oracle/ltx_oracle.py."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as Fn
from einops import rearrange

from helpers import O

G = torch.Generator().manual_seed(2024)


def rnd(*shape, dtype=torch.float64):
    return torch.randn(*shape, generator=G, dtype=torch.float32).to(dtype)


def test_rms_norm_gelu_silu_match_torch_functional():
    x, w = rnd(3, 7, 96), rnd(96)
    assert torch.allclose(O.rms_norm(x, w, 1e-6), Fn.rms_norm(x, (96,), w, 1e-6), atol=1e-12)      # MLXFast.rmsNorm
    assert torch.allclose(O.rms_norm(x, None, 1e-6), Fn.rms_norm(x, (96,), None, 1e-6), atol=1e-12)
    assert torch.allclose(O.gelu_tanh(x), Fn.gelu(x, approximate="tanh"), atol=1e-12)              # MLXNN.geluApproximate
    assert torch.allclose(O.silu(x), Fn.silu(x), atol=1e-12)                                       # MLXNN.silu
    z = rnd(2, 16, 3, 4, 5)
    ref = rearrange(Fn.rms_norm(rearrange(z, "b c t h w -> b t h w c"), (16,), None, 1e-8), "b t h w c -> b c t h w")
    assert torch.allclose(O.pixel_norm(z), ref, atol=1e-12)                                        # vaePixelNorm: RMS over channels


@pytest.mark.parametrize("heads,Nq,Nk,masked", [(2, 5, 9, False), (4, 12, 7, True)])
def test_sdpa_matches_torch_scaled_dot_product_attention(heads, Nq, Nk, masked):
    d = 16
    q, k, v = rnd(2, Nq, heads * d), rnd(2, Nk, heads * d), rnd(2, Nk, heads * d)
    bias = None
    if masked:
        m = torch.ones(2, Nk)
        m[:, :3] = 0
        bias = ((1.0 - m) * -10000.0).view(2, 1, 1, Nk).to(torch.float64)                          # prepareAttentionMask
    split = lambda t: rearrange(t, "b n (h d) -> b h n d", h=heads)                                # noqa: E731
    ref = Fn.scaled_dot_product_attention(split(q), split(k), split(v), attn_mask=bias, scale=1.0 / math.sqrt(d))
    assert torch.allclose(O.sdpa(q, k, v, heads, bias), rearrange(ref, "b h n d -> b n (h d)"), atol=1e-10)


def test_split_rope_is_a_complex_rotation():
    heads, d, N = 2, 8, 6
    x = rnd(1, N, heads * d)
    ang = rnd(heads, N, d // 2)
    got = O.apply_split_rope(x, torch.cos(ang), torch.sin(ang), heads)
    xh = rearrange(x, "b n (h d) -> b h n d", h=heads)
    z = torch.complex(xh[..., : d // 2], xh[..., d // 2:]) * torch.polar(torch.ones_like(ang), ang)     # (x1 + i x2) e^{i a}
    ref = rearrange(torch.cat([z.real, z.imag], -1), "b h n d -> b n (h d)")
    assert torch.allclose(got, ref, atol=1e-12)


def test_sinusoidal_embedding_closed_form():
    t = torch.tensor([0.0, 1.0, 250.0, 999.0])
    got = O.sinusoidal_embedding(t, 256).double().numpy()
    k = np.arange(128, dtype=np.float64)
    f = 10000.0 ** (-k / 128.0)
    ref = np.concatenate([np.cos(t.double().numpy()[:, None] * f), np.sin(t.double().numpy()[:, None] * f)], -1)   # [cos | sin]
    assert np.abs(got - ref).max() < 2e-4           # fp32 arguments up to 999 rad


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("pad", ["reflect", "zeros"])
def test_conv3d_full_matches_explicit_padding_plus_conv3d(causal, pad):
    x, w, b = rnd(2, 5, 4, 6, 7), rnd(3, 5, 3, 3, 3), rnd(3)
    # spatial padding first (per frame), then frame replication in time (V/VideoConvolution.swift:281-294, 296-316)
    xs = rearrange(x, "b c t h w -> (b t) c h w")
    xs = Fn.pad(xs, (1, 1, 1, 1), mode="reflect") if pad == "reflect" else Fn.pad(xs, (1, 1, 1, 1))
    xs = rearrange(xs, "(b t) c h w -> b c t h w", b=2)
    first, last = xs[:, :, :1], xs[:, :, -1:]
    xs = torch.cat([first, first, xs], 2) if causal else torch.cat([first, xs, last], 2)
    # the reference evaluates the 3-D kernel as three 2-D convolutions over shifted frames and sums them (:318-340)
    ref = sum(rearrange(Fn.conv2d(rearrange(xs[:, :, dt:dt + 4], "b c t h w -> (b t) c h w"), w[:, :, dt]), "(b t) c h w -> b c t h w", b=2)
              for dt in range(3)) + b.view(1, 3, 1, 1, 1)
    assert torch.allclose(O.conv3d_full(x, w, b, causal, pad), ref, atol=1e-10)


def test_shuffles_match_einops_patterns():
    x = rnd(2, 3 * 8, 2, 3, 4)
    # (b, c, 2, 2, 2, t, h, w) -> transpose(0, 1, 5, 2, 6, 3, 7, 4)  (V/VideoDecoder.swift:201-212)
    assert torch.equal(O.depth_to_space(x, 3), rearrange(x, "b (c p1 p2 p3) t h w -> b c (t p1) (h p2) (w p3)", p1=2, p2=2, p3=2))
    y = rnd(1, 3 * 16, 2, 3, 4)
    # unpatchify: transposed(0, 1, 5, 2, 6, 4, 7, 3): the channel index is (c, pW, pH) -- pW before pH (:257-275)
    assert torch.equal(O.vae_unpatchify(y), rearrange(y, "b (c q pw ph) t h w -> b c (t q) (h ph) (w pw)", q=1, pw=4, ph=4))
    p = rnd(1, 3, 2, 8, 12)
    assert torch.equal(O.encoder_patchify(p), rearrange(p, "b c t (h ph) (w pw) -> b (c pw ph) t h w", ph=4, pw=4))
    assert torch.equal(O.vae_unpatchify(O.encoder_patchify(p)), p)                     # the two are inverses
    s = rnd(1, 2, 4, 6, 8)
    assert torch.equal(O.space_to_depth(s, (2, 2, 2)), rearrange(s, "b c (t ft) (h fh) (w fw) -> b (c ft fh fw) t h w", ft=2, fh=2, fw=2))
    lat = rnd(2, 5, 2, 3, 4)
    assert torch.equal(O.patchify(lat), rearrange(lat, "b c f h w -> b (f h w) c"))
    assert torch.equal(O.unpatchify(O.patchify(lat), (2, 3, 4)), lat)


def test_group_norm_and_guidance_statistics_match_torch():
    x, w, b = rnd(2, 64, 3, 4, 5), rnd(64), rnd(64)
    assert torch.allclose(O._group_norm(x, w, b, 32, 1e-5), Fn.group_norm(x, 32, w, b, 1e-5), atol=1e-10)
    cond, unc = rnd(2, 4, 3, 5, 6), rnd(2, 4, 3, 5, 6)
    g = O.apply_cfg(unc, cond, 4.0)
    assert torch.allclose(g, unc + 4.0 * (cond - unc), atol=1e-12)                     # the textbook form of applyCFG
    r = O.guidance_rescale(g, cond, 0.7)
    sc = torch.sqrt(cond.flatten(1).var(1, unbiased=False) + 1e-8).view(2, 1, 1, 1, 1)
    sg = torch.sqrt(g.flatten(1).var(1, unbiased=False) + 1e-8).view(2, 1, 1, 1, 1)
    assert torch.allclose(r, 0.7 * g * sc / sg + 0.3 * g, atol=1e-12)
    x0, v = rnd(1, 4, 2, 3, 3), rnd(1, 4, 2, 3, 3)
    assert torch.allclose(O.euler_step(x0, v, 0.8, 0.5), x0 + (0.5 - 0.8) * v, atol=1e-12)    # flow-matching Euler in one line
    assert torch.allclose(O.euler_step(x0, v, 0.8, 0.0), x0 - 0.8 * v, atol=1e-12)


def test_linear_layernorm_and_adain_match_torch():
    w = {"l.weight": rnd(7, 5), "l.bias": rnd(7)}
    x = rnd(3, 4, 5)
    assert torch.allclose(O.linear(x, w, "l"), Fn.linear(x, w["l.weight"], w["l.bias"]), atol=1e-12)
    up, ref = rnd(1, 6, 2, 4, 4), rnd(1, 6, 2, 2, 2)
    out = O.adain_filter_latent(up, ref, 1.0)
    # AdaIN: per-channel statistics of the result equal the reference's (P/LatentUtils.swift:201-227)
    assert torch.allclose(out.flatten(2).mean(-1), ref.flatten(2).mean(-1), atol=1e-10)
    assert torch.allclose(out.flatten(2).std(-1, unbiased=False), ref.flatten(2).std(-1, unbiased=False), rtol=1e-4)


@pytest.mark.parametrize("heads,head_dim,fhw", [(2, 16, (2, 2, 3)), (4, 32, (3, 2, 2)), (2, 128, (2, 3, 2))])
def test_rope_table_against_scalar_loops(heads, head_dim, fhw):
    """precomputeFreqsCis (doublePrecision, split; T/LTXRoPE.swift:375-488) written out as the reference's own scalar loops --
    a second derivation of the vectorised oracle table, including the left identity padding and the head split."""
    cfg = O.DiTConfig(num_layers=1, num_heads=heads, head_dim=head_dim)
    cos, sin = O.rope_table(cfg, *fhw)
    D, theta, max_pos = heads * head_dim, cfg.rope_theta, cfg.max_pos
    F, H, W = fhw
    grid = []                                    # createPositionGrid (:552-610): (t, h, w) mid-points per token, fp32
    for f in range(F):
        s0, e0 = max(f * 8 + (1 - 8), 0), max((f + 1) * 8 + (1 - 8), 0)
        for h in range(H):
            for w in range(W):
                grid.append((np.float32((np.float32(s0) + np.float32(e0)) / np.float32(2.0)) / np.float32(24.0),
                             np.float32(h * 32 + 16.0), np.float32(w * 32 + 16.0)))
    n_idx = max(1, D // 6)
    idx = [theta ** ((i / (n_idx - 1)) if n_idx > 1 else 0.0) * (math.pi / 2.0) for i in range(n_idx)]
    pad = max(0, D // 2 - n_idx * 3)
    hd2 = (D // 2) // heads
    for t, pos in enumerate(grid):
        row_c, row_s = [1.0] * pad, [0.0] * pad
        for fi in range(n_idx):
            for d in range(3):
                a = idx[fi] * (float(pos[d]) / max_pos[d] * 2.0 - 1.0)
                row_c.append(math.cos(a))
                row_s.append(math.sin(a))
        for hh in range(heads):
            assert np.allclose(cos[hh, t].numpy(), np.float32(row_c[hh * hd2:(hh + 1) * hd2]), atol=1e-7)
            assert np.allclose(sin[hh, t].numpy(), np.float32(row_s[hh * hd2:(hh + 1) * hd2]), atol=1e-7)
