"""CPU oracle for the LTX-2 denoise hot path (DiT step, guidance+Euler, video-VAE decode).

TEST INFRASTRUCTURE ONLY.  This file is a CPU restatement (PyTorch, fp32 with an fp64 switch) of the
reference's algorithm, written from the Swift sources under /root/reference (read-only, not present on
the GPU box).  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs
of `bench.py` may import it -- as the checker or the timed CPU baseline, never as the product path.

PARITY UNPINNED: the reference (Swift + MLX 0.30.6) cannot be built or imported here (no swift, no mlx)
and ships no golden vectors for this path (its only test is `testVersion`).  The arithmetic lives in the
un-vendored third-party package `mlx-swift` (Package.swift:21, exact 0.30.6); the semantics assumed for
its primitives are the published ones:
  rmsNorm(x,w,eps) = x * rsqrt(mean(x^2)+eps) * w       (fp32 accumulation)
  scaledDotProductAttention = softmax_fp32(q k^T * scale + mask) v
  Linear = x W^T + b (W [out,in]);  LayerNorm(affine:false) uses the population variance
  geluApproximate = 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)));  variance = population (ddof 0)
  type promotion bf16 (x) f32 -> f32.
The only in-source known-answer data are the sigma tables (S/LTXScheduler.swift:18-36) and the shape
formulae; those are checked in tests/test_oracle.py.  Each assumed primitive is additionally checked
against a second, code-independent implementation (torch.nn.functional / einops / complex arithmetic)
in tests/test_oracle_primitives.py -- that pins the DEFINITIONS, not MLX's bits: parity stays unpinned.

File:line citations are relative to /root/reference/Sources/LTXVideo/ (T/ = Models/Transformer,
V/ = Models/VAE, P/ = Pipeline, S/ = Scheduler, C/ = Configuration).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# configuration (C/LTXConfig.swift:122-139)
# ----------------------------------------------------------------------------------------------
@dataclass
class DiTConfig:
    num_layers: int = 48
    num_heads: int = 32
    head_dim: int = 128
    in_channels: int = 128
    out_channels: int = 128
    caption_channels: int = 3840
    ffn_mult: int = 4                      # T/LTXFeedForward.swift:39
    rope_theta: float = 10000.0
    max_pos: Tuple[int, int, int] = (20, 2048, 2048)
    timestep_scale_multiplier: float = 1000.0
    norm_eps: float = 1e-6

    @property
    def inner_dim(self) -> int:
        return self.num_heads * self.head_dim


@dataclass
class VAEConfig:
    latent_channels: int = 128
    # channel plan V/VideoDecoder.swift:331-355 : res @c0, d2s, res @c0/2, d2s, res @c0/4, d2s, res @c0/8
    base_channels: int = 1024
    blocks_per_stage: int = 5
    patch_size: int = 4
    causal: bool = False                   # VideoDecoder() default, V/VideoDecoder.swift:320
    timestep_conditioning: bool = False

    @property
    def stage_channels(self) -> List[int]:
        c = self.base_channels
        return [c, c // 2, c // 4, c // 8]


def bf16_round(x: Tensor) -> Tensor:
    return x.to(torch.bfloat16).to(x.dtype)


# ----------------------------------------------------------------------------------------------
# seeded random-init weights under the reference's post-mapping key names (SURVEY Appendix C,
# U/ModelDownloader.swift:756-899).  Values are bf16-representable when bf16=True, mirroring the loader's
# fp32->bf16 cast (U/ModelDownloader.swift:1005-1012).
# ----------------------------------------------------------------------------------------------
def make_dit_weights(cfg: DiTConfig, seed: int = 0, bf16: bool = True) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    D = cfg.inner_dim
    w: Dict[str, Tensor] = {}

    def lin(name: str, out_f: int, in_f: int, wstd: Optional[float] = None):
        std = wstd if wstd is not None else 1.0 / math.sqrt(in_f)
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    lin("patchify_proj", D, cfg.in_channels)
    lin("adaln_single.emb.linear_1", D, 256)
    lin("adaln_single.emb.linear_2", D, D)
    lin("adaln_single.linear", 6 * D, D, wstd=0.5 / math.sqrt(D))
    lin("caption_projection.linear_1", D, cfg.caption_channels)
    lin("caption_projection.linear_2", D, D)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        w[p + "scale_shift_table"] = torch.randn(6, D, generator=g) * 0.1
        for a in ("attn1", "attn2"):
            for l in ("to_q", "to_k", "to_v", "to_out"):
                lin(p + f"{a}.{l}", D, D)
            w[p + f"{a}.q_norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            w[p + f"{a}.k_norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
        lin(p + "ff.project_in.proj", cfg.ffn_mult * D, D)
        lin(p + "ff.project_out", D, cfg.ffn_mult * D)
    w["scale_shift_table"] = torch.randn(2, D, generator=g) * 0.1
    lin("proj_out", cfg.out_channels, D)
    if bf16:
        w = {k: bf16_round(v) for k, v in w.items()}
    return w


def make_vae_weights(cfg: VAEConfig, seed: int = 0) -> Dict[str, Tensor]:
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, Tensor] = {}

    def conv(name: str, cout: int, cin: int):
        w[name + ".conv.weight"] = torch.randn(cout, cin, 3, 3, 3, generator=g) / math.sqrt(27 * cin)
        w[name + ".conv.bias"] = torch.randn(cout, generator=g) * 0.02

    def lin(name: str, out_f: int, in_f: int):
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) / math.sqrt(in_f)
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    C = cfg.latent_channels
    w["mean_of_means"] = torch.randn(C, generator=g) * 0.1
    w["std_of_means"] = 1.0 + 0.1 * torch.rand(C, generator=g)
    w["timestep_scale_multiplier"] = torch.tensor(1000.0)
    chans = cfg.stage_channels
    conv("conv_in", chans[0], C)
    for s, c in enumerate(chans):
        blk = f"up_blocks_{2 * s}"
        for j in range(cfg.blocks_per_stage):
            conv(f"{blk}.res_blocks.{j}.conv1", c, c)
            conv(f"{blk}.res_blocks.{j}.conv2", c, c)
            w[f"{blk}.res_blocks.{j}.scale_shift_table"] = torch.randn(4, c, generator=g) * 0.1
        lin(f"{blk}.time_embedder.timestep_embedder.linear_1", 256, 256)
        lin(f"{blk}.time_embedder.timestep_embedder.linear_2", 4 * c, 256)
        if s < len(chans) - 1:
            conv(f"up_blocks_{2 * s + 1}.conv", 4 * c, c)
    w["last_scale_shift_table"] = torch.randn(2, chans[-1], generator=g) * 0.1
    lin("last_time_embedder.timestep_embedder.linear_1", 256, 256)
    lin("last_time_embedder.timestep_embedder.linear_2", 2 * chans[-1], 256)
    conv("conv_out", 3 * cfg.patch_size * cfg.patch_size, chans[-1])
    return w


# ----------------------------------------------------------------------------------------------
# primitives (assumed MLX semantics, see header)
# ----------------------------------------------------------------------------------------------
def linear(x: Tensor, w: Dict[str, Tensor], name: str) -> Tensor:
    return x @ w[name + ".weight"].to(x.dtype).t() + w[name + ".bias"].to(x.dtype)


def rms_norm(x: Tensor, weight: Optional[Tensor], eps: float) -> Tensor:
    y = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return y if weight is None else y * weight.to(x.dtype)


def gelu_tanh(x: Tensor) -> Tensor:                     # T/LTXFeedForward.swift:10-13
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x.pow(3))))


def silu(x: Tensor) -> Tensor:
    return x * torch.sigmoid(x)


def sinusoidal_embedding(t: Tensor, dim: int = 256) -> Tensor:
    """T/LTXTimestepEmbedding.swift:17-54: [cos(t f_k), sin(t f_k)], f_k = exp(-ln(1e4) k/half); fp32."""
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32) / float(half)
    freqs = torch.exp(-math.log(10000.0) * k)
    args = t.reshape(-1, 1).to(torch.float32) * freqs.reshape(1, -1)
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


# ----------------------------------------------------------------------------------------------
# RoPE (T/LTXRoPE.swift)
# ----------------------------------------------------------------------------------------------
def position_grid(F: int, H: int, W: int, t_scale: int = 8, s_scale: int = 32, fps: float = 24.0) -> Tensor:
    """T/LTXRoPE.swift:552-610: pixel-space mid-points, causal fix on t, t divided by fps. [3, F*H*W] fp32."""
    f = torch.arange(F, dtype=torch.float32)
    start = torch.clamp(f * t_scale + (1 - t_scale), min=0)
    end = torch.clamp((f + 1) * t_scale + (1 - t_scale), min=0)
    tc = ((start + end) / 2.0) / fps
    hc = torch.arange(H, dtype=torch.float32) * s_scale + s_scale / 2.0
    wc = torch.arange(W, dtype=torch.float32) * s_scale + s_scale / 2.0
    tg = tc.view(F, 1, 1).expand(F, H, W).reshape(-1)
    hg = hc.view(1, H, 1).expand(F, H, W).reshape(-1)
    wg = wc.view(1, 1, W).expand(F, H, W).reshape(-1)
    return torch.stack([tg, hg, wg], 0)


def rope_table(cfg: DiTConfig, F: int, H: int, W: int) -> Tuple[Tensor, Tensor]:
    """T/LTXRoPE.swift:375-488 (doublePrecision path, split type): returns cos, sin [heads, N, head_dim/2] fp32."""
    D = cfg.inner_dim
    grid = position_grid(F, H, W).to(torch.float64)           # (3, N), fp32 values widened (:388-417)
    n_dims = 3
    num_idx = max(1, D // (2 * n_dims))                        # :396
    i = torch.arange(num_idx, dtype=torch.float64)
    t = i / (num_idx - 1) if num_idx > 1 else torch.zeros(1, dtype=torch.float64)
    idx = torch.pow(torch.tensor(cfg.rope_theta, dtype=torch.float64), t) * (math.pi / 2.0)   # :398-404
    max_pos = torch.tensor(cfg.max_pos, dtype=torch.float64).view(3, 1)
    scaled = (grid / max_pos) * 2.0 - 1.0                      # (3, N)  :419-427
    # freqs[n, fi*3 + d] = idx[fi] * scaled[d, n]               :432-441
    freqs = (idx.view(1, num_idx, 1) * scaled.t().reshape(-1, 1, n_dims)).reshape(-1, num_idx * n_dims)
    cos, sin = torch.cos(freqs), torch.sin(freqs)
    pad = max(0, D // 2 - num_idx * n_dims)                    # :451-476 left pad with identity
    N = freqs.shape[0]
    cos = torch.cat([torch.ones(N, pad, dtype=torch.float64), cos], 1).to(torch.float32)
    sin = torch.cat([torch.zeros(N, pad, dtype=torch.float64), sin], 1).to(torch.float32)
    hd2 = (D // 2) // cfg.num_heads
    cos = cos.view(N, cfg.num_heads, hd2).permute(1, 0, 2).contiguous()   # :484-488
    sin = sin.view(N, cfg.num_heads, hd2).permute(1, 0, 2).contiguous()
    return cos, sin


def apply_split_rope(x: Tensor, cos: Tensor, sin: Tensor, heads: int) -> Tensor:
    """T/LTXRoPE.swift:84-149: x [B,N,D] viewed [B,H,N,d]; halves (x1|x2); y1=x1 c - x2 s, y2 = x2 c + x1 s."""
    B, N, D = x.shape
    d = D // heads
    xh = x.view(B, N, heads, d).permute(0, 2, 1, 3)
    x1, x2 = xh[..., : d // 2], xh[..., d // 2:]
    c, s = cos.to(x.dtype).unsqueeze(0), sin.to(x.dtype).unsqueeze(0)
    y = torch.cat([x1 * c - x2 * s, x2 * c + x1 * s], dim=-1)
    return y.permute(0, 2, 1, 3).reshape(B, N, D)


# ----------------------------------------------------------------------------------------------
# attention (T/LTXAttention.swift:160-218)
# ----------------------------------------------------------------------------------------------
def sdpa(q: Tensor, k: Tensor, v: Tensor, heads: int, bias: Optional[Tensor]) -> Tensor:
    B, Nq, D = q.shape
    Nk = k.shape[1]
    d = D // heads
    qh = q.view(B, Nq, heads, d).permute(0, 2, 1, 3)
    kh = k.view(B, Nk, heads, d).permute(0, 2, 1, 3)
    vh = v.view(B, Nk, heads, d).permute(0, 2, 1, 3)
    s = (qh @ kh.transpose(-1, -2)) * (1.0 / math.sqrt(d))       # :208
    if bias is not None:
        s = s + bias.to(s.dtype)
    p = torch.softmax(s, dim=-1)
    o = p @ vh
    return o.permute(0, 2, 1, 3).reshape(B, Nq, D)


def attention(w: Dict[str, Tensor], prefix: str, x: Tensor, ctx: Optional[Tensor], cfg: DiTConfig,
              rope: Optional[Tuple[Tensor, Tensor]], bias: Optional[Tensor], bf16_kv: bool = False) -> Tensor:
    c = x if ctx is None else ctx
    q = linear(x, w, prefix + ".to_q")
    k = linear(c, w, prefix + ".to_k")
    v = linear(c, w, prefix + ".to_v")
    if bf16_kv:                                  # text branch stays bf16 in the reference
        k, v = bf16_round(k), bf16_round(v)
    q = rms_norm(q, w[prefix + ".q_norm.weight"], cfg.norm_eps)          # :179-180, across all heads
    k = rms_norm(k, w[prefix + ".k_norm.weight"], cfg.norm_eps)
    if bf16_kv:
        k = bf16_round(k)
    if rope is not None:                                                  # :185-189
        q = apply_split_rope(q, rope[0], rope[1], cfg.num_heads)
        k = apply_split_rope(k, rope[0], rope[1], cfg.num_heads)
    o = sdpa(q, k, v, cfg.num_heads, bias)
    return linear(o, w, prefix + ".to_out")


# ----------------------------------------------------------------------------------------------
# transformer block (T/LTXTransformerBlock.swift:187-232) and full forward (T/LTXTransformer.swift:235-486)
# ----------------------------------------------------------------------------------------------
def block_forward(w: Dict[str, Tensor], i: int, x: Tensor, ada: Tensor, ctx: Tensor, bias: Optional[Tensor],
                  rope: Tuple[Tensor, Tensor], cfg: DiTConfig, skip_self_attn: bool = False,
                  skip_ff: bool = False, cross_attn_scale: float = 1.0, bf16_kv: bool = False) -> Tensor:
    """x [B,N,D]; ada [B,1|N,6,D]; ctx [B,S,D] (already caption-projected)."""
    p = f"transformer_blocks.{i}"
    m = w[p + ".scale_shift_table"].to(x.dtype).view(1, 1, 6, -1) + ada         # :163-185
    shift_msa, scale_msa, gate_msa = m[:, :, 0], m[:, :, 1], m[:, :, 2]
    shift_mlp, scale_mlp, gate_mlp = m[:, :, 3], m[:, :, 4], m[:, :, 5]
    if not skip_self_attn:                                                        # :199
        h = rms_norm(x, None, cfg.norm_eps) * (1 + scale_msa) + shift_msa        # :72-83
        x = x + attention(w, p + ".attn1", h, None, cfg, rope, None) * gate_msa   # :202
    ca = attention(w, p + ".attn2", x, ctx, cfg, None, bias, bf16_kv=bf16_kv)     # :205-210, x NOT normalised
    if cross_attn_scale != 1.0:
        ca = ca * cross_attn_scale
    x = x + ca
    if not skip_ff:
        h = rms_norm(x, None, cfg.norm_eps) * (1 + scale_mlp) + shift_mlp
        ff = linear(gelu_tanh(linear(h, w, p + ".ff.project_in.proj")), w, p + ".ff.project_out")
        x = x + ff * gate_mlp                                                     # :225-229
    return x


def timestep_path(w: Dict[str, Tensor], sigma: Tensor, cfg: DiTConfig, dtype: torch.dtype) -> Tuple[Tensor, Tensor]:
    """T/LTXTransformer.swift:105-121 + T/LTXTimestepEmbedding.swift:62-124. sigma [B] or [B,N].
    Returns ada [B,1|N,6,D], emb [B,1|N,D]."""
    B = sigma.shape[0]
    t = sigma.to(torch.float32) * cfg.timestep_scale_multiplier
    se = sinusoidal_embedding(t.reshape(-1)).to(dtype)
    emb = linear(silu(linear(se, w, "adaln_single.emb.linear_1")), w, "adaln_single.emb.linear_2")
    ada = linear(silu(emb), w, "adaln_single.linear")
    D = cfg.inner_dim
    return ada.view(B, -1, 6, D), emb.view(B, -1, D)


def caption_projection(w: Dict[str, Tensor], context: Tensor, mlx_bf16: bool) -> Tensor:
    """T/LTXTimestepEmbedding.swift:146-151."""
    h = linear(context, w, "caption_projection.linear_1")
    if mlx_bf16:
        h = bf16_round(h)
    h = gelu_tanh(h)
    if mlx_bf16:
        h = bf16_round(h)
    h = linear(h, w, "caption_projection.linear_2")
    if mlx_bf16:
        h = bf16_round(h)
    return h


def dit_forward(w: Dict[str, Tensor], cfg: DiTConfig, latent: Tensor, context: Tensor, sigma: Tensor,
                mask: Optional[Tensor], fhw: Tuple[int, int, int], stg_blocks: Sequence[int] = (),
                skip_self_attn: bool = False, skip_ff: bool = False,
                cross_attn_scale: Optional[Dict[int, float]] = None,
                dtype: torch.dtype = torch.float32, mlx_bf16: bool = True,
                return_blocks: bool = False):
    """Velocity prediction.  latent [B,N,C_in], context [B,S,C_cap], sigma [B]|[B,N], mask [B,S] (1=attend).
    mlx_bf16=True mimics the reference's bf16 mode: bf16-rounded inputs, bf16 patchify/caption/text-KV branch,
    fp32 everywhere else.  mlx_bf16=False is the all-fp32 (or fp64) mode."""
    F, H, W = fhw
    latent = latent.to(dtype)
    context = context.to(dtype)
    if mlx_bf16:
        latent, context = bf16_round(latent), bf16_round(context)     # P/LTXPipeline.swift:815
    x = linear(latent, w, "patchify_proj")                            # T/LTXTransformer.swift:257
    if mlx_bf16:
        x = bf16_round(x)
    ada, emb = timestep_path(w, sigma, cfg, dtype)                    # :274
    c = caption_projection(w, context, mlx_bf16)                      # :303
    bias = None
    if mask is not None:                                              # :141-156
        bias = ((1.0 - mask.to(dtype)) * -10000.0).view(mask.shape[0], 1, 1, mask.shape[-1])
    rope = rope_table(cfg, F, H, W)                                   # :322
    blocks_out = []
    for i in range(cfg.num_layers):                                   # :446-465
        in_stg = i in stg_blocks
        x = block_forward(w, i, x, ada, c, bias, rope, cfg,
                          skip_self_attn=skip_self_attn and in_stg, skip_ff=skip_ff and in_stg,
                          cross_attn_scale=(cross_attn_scale or {}).get(i, 1.0), bf16_kv=mlx_bf16)
        if return_blocks:
            blocks_out.append(x.clone())
    # processOutput :208-224 -- LayerNorm(no affine, eps 1e-6), rows: 0 = shift, 1 = scale
    o = w["scale_shift_table"].to(dtype).view(1, 1, 2, -1) + emb.unsqueeze(2)
    mu = x.mean(-1, keepdim=True)
    var = (x - mu).pow(2).mean(-1, keepdim=True)
    y = (x - mu) * torch.rsqrt(var + cfg.norm_eps)
    y = y * (1 + o[:, :, 1]) + o[:, :, 0]
    vel = linear(y, w, "proj_out")
    if return_blocks:
        return vel, blocks_out
    return vel


def single_block_forward(w: Dict[str, Tensor], cfg: DiTConfig, x: Tensor, ctx: Tensor, sigma: Tensor,
                         fhw: Tuple[int, int, int], block: int = 0, dtype=torch.float32) -> Tensor:
    """BASELINE config 1: one BasicTransformerBlock, fp32, context already inner_dim wide, mask all ones."""
    ada, _ = timestep_path(w, sigma, cfg, dtype)
    rope = rope_table(cfg, *fhw)
    return block_forward(w, block, x.to(dtype), ada, ctx.to(dtype), None, rope, cfg)


# ----------------------------------------------------------------------------------------------
# dual audio / video transformer (T/LTX2Transformer.swift, T/LTX2TransformerBlock.swift) -- SURVEY 8f-1
# ----------------------------------------------------------------------------------------------
@dataclass
class AVConfig:
    """Audio-side constants of LTXTransformerConfig (C/LTXConfig.swift:134-173): 32 heads x 64, 128 latent channels,
    1-D positions with max_pos [20]."""
    audio_heads: int = 32
    audio_head_dim: int = 64
    audio_in_channels: int = 128
    audio_out_channels: int = 128
    audio_max_pos: int = 20

    @property
    def audio_dim(self) -> int:
        return self.audio_heads * self.audio_head_dim


def make_av_weights(cfg: DiTConfig, av: AVConfig, seed: int = 0, bf16: bool = True) -> Dict[str, Tensor]:
    """Random-init LTX2Transformer weights under the Swift module keys (T/LTX2Transformer.swift:20-43,
    T/LTX2TransformerBlock.swift:44-72).  The video-only keys are those of make_dit_weights plus the learned norms."""
    w = make_dit_weights(cfg, seed, bf16=False)
    g = torch.Generator().manual_seed(seed + 7919)
    D, Da = cfg.inner_dim, av.audio_dim

    def lin(name, out_f, in_f, wstd=None):
        std = wstd if wstd is not None else 1.0 / math.sqrt(in_f)
        w[name + ".weight"] = torch.randn(out_f, in_f, generator=g) * std
        w[name + ".bias"] = torch.randn(out_f, generator=g) * 0.02

    def adaln(name, dim, n):
        lin(name + ".emb.linear_1", dim, 256)
        lin(name + ".emb.linear_2", dim, dim)
        lin(name + ".linear", n * dim, dim, wstd=0.5 / math.sqrt(dim))

    def attn(name, qdim, cdim, inner):
        lin(name + ".to_q", inner, qdim)
        lin(name + ".to_k", inner, cdim)
        lin(name + ".to_v", inner, cdim)
        lin(name + ".to_out", qdim, inner)
        w[name + ".q_norm.weight"] = 1.0 + 0.1 * torch.randn(inner, generator=g)
        w[name + ".k_norm.weight"] = 1.0 + 0.1 * torch.randn(inner, generator=g)

    def norm(name, dim):
        w[name + ".weight"] = 1.0 + 0.1 * torch.randn(dim, generator=g)

    lin("audio_patchify_proj", Da, av.audio_in_channels)
    adaln("audio_adaln_single", Da, 6)
    lin("audio_caption_projection.linear_1", Da, cfg.caption_channels)
    lin("audio_caption_projection.linear_2", Da, Da)
    w["audio_scale_shift_table"] = torch.randn(2, Da, generator=g) * 0.1
    lin("audio_proj_out", av.audio_out_channels, Da)
    adaln("av_ca_video_scale_shift_adaln_single", D, 4)
    adaln("av_ca_a2v_gate_adaln_single", D, 1)
    adaln("av_ca_audio_scale_shift_adaln_single", Da, 4)
    adaln("av_ca_v2a_gate_adaln_single", Da, 1)
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}."
        for n in ("norm1", "norm2", "norm3", "audio_to_video_norm"):
            norm(p + n, D)
        for n in ("audio_norm1", "audio_norm2", "audio_norm3", "video_to_audio_norm"):
            norm(p + n, Da)
        attn(p + "audio_attn1", Da, Da, Da)
        attn(p + "audio_attn2", Da, Da, Da)            # context = audio caption projection (Da wide)
        lin(p + "audio_ff.project_in.proj", cfg.ffn_mult * Da, Da)
        lin(p + "audio_ff.project_out", Da, cfg.ffn_mult * Da)
        w[p + "audio_scale_shift_table"] = torch.randn(6, Da, generator=g) * 0.1
        attn(p + "audio_to_video_attn", D, Da, Da)     # Q from video, K/V from audio, audio head layout
        attn(p + "video_to_audio_attn", Da, D, Da)     # Q from audio, K/V from video
        w[p + "scale_shift_table_a2v_ca_video"] = torch.randn(5, D, generator=g) * 0.1
        w[p + "scale_shift_table_a2v_ca_audio"] = torch.randn(5, Da, generator=g) * 0.1
    if bf16:
        w = {k: bf16_round(v) for k, v in w.items()}
    return w


def audio_position_grid(frames: int, hop: int = 160, sr: int = 16000, scale: int = 4, causal_offset: int = 1) -> Tensor:
    """createAudioPositionGrid (T/LTXRoPE.swift:627-655): mid-point of the latent frame's mel span in seconds. [1, T]."""
    f = torch.arange(frames, dtype=torch.float32)
    start = torch.clamp(f * scale + causal_offset - scale, min=0)
    end = torch.clamp((f + 1) * scale + causal_offset - scale, min=0)
    return ((start + end) / 2.0 * hop / sr).view(1, -1)


def rope_table_nd(grid: Tensor, dim: int, heads: int, theta: float, max_pos: Sequence[int]) -> Tuple[Tensor, Tensor]:
    """precomputeFreqsCis, split type, doublePrecision (T/LTXRoPE.swift:375-488) for an [n_dims, T] position grid:
    returns cos, sin [heads, T, dim / (2 heads)] fp32.  rope_table() is the n_dims = 3 video case."""
    n_dims = grid.shape[0]
    g64 = grid.to(torch.float64)
    num_idx = max(1, dim // (2 * n_dims))
    i = torch.arange(num_idx, dtype=torch.float64)
    t = i / (num_idx - 1) if num_idx > 1 else torch.zeros(1, dtype=torch.float64)
    idx = torch.pow(torch.tensor(theta, dtype=torch.float64), t) * (math.pi / 2.0)
    mp = torch.tensor(list(max_pos), dtype=torch.float64).view(n_dims, 1)
    scaled = (g64 / mp) * 2.0 - 1.0
    freqs = (idx.view(1, num_idx, 1) * scaled.t().reshape(-1, 1, n_dims)).reshape(-1, num_idx * n_dims)
    cos, sin = torch.cos(freqs), torch.sin(freqs)
    pad = max(0, dim // 2 - num_idx * n_dims)
    T = freqs.shape[0]
    cos = torch.cat([torch.ones(T, pad, dtype=torch.float64), cos], 1).to(torch.float32)
    sin = torch.cat([torch.zeros(T, pad, dtype=torch.float64), sin], 1).to(torch.float32)
    hd2 = (dim // 2) // heads
    return (cos.view(T, heads, hd2).permute(1, 0, 2).contiguous(), sin.view(T, heads, hd2).permute(1, 0, 2).contiguous())


def _adaln_single(w, name: str, t_scaled: Tensor, dtype) -> Tuple[Tensor, Tensor]:
    """AdaLayerNormSingle (T/LTXTimestepEmbedding.swift:96-124): returns (linear(silu(emb)), emb)."""
    se = sinusoidal_embedding(t_scaled.reshape(-1)).to(dtype)
    emb = linear(silu(linear(se, w, name + ".emb.linear_1")), w, name + ".emb.linear_2")
    return linear(silu(emb), w, name + ".linear"), emb


def _av_attention(w, prefix: str, x: Tensor, ctx: Optional[Tensor], heads: int, eps: float,
                  q_rope=None, k_rope=None, bias=None, bf16_kv: bool = False) -> Tensor:
    """LTXAttention.callAsFunction (T/LTXAttention.swift:160-218) with separate query / key RoPE (pe / kPe)."""
    c = x if ctx is None else ctx
    q, k, v = linear(x, w, prefix + ".to_q"), linear(c, w, prefix + ".to_k"), linear(c, w, prefix + ".to_v")
    if bf16_kv:
        k, v = bf16_round(k), bf16_round(v)
    q = rms_norm(q, w[prefix + ".q_norm.weight"], eps)
    k = rms_norm(k, w[prefix + ".k_norm.weight"], eps)
    if bf16_kv:
        k = bf16_round(k)
    if q_rope is not None:
        q = apply_split_rope(q, q_rope[0], q_rope[1], heads)
        kr = k_rope if k_rope is not None else q_rope
        k = apply_split_rope(k, kr[0], kr[1], heads)
    return linear(sdpa(q, k, v, heads, bias), w, prefix + ".to_out")


def av_dit_forward(w, cfg: DiTConfig, av: AVConfig, v_latent: Tensor, a_latent: Tensor, v_context: Tensor, a_context: Tensor,
                   v_sigma: Tensor, a_sigma: Tensor, v_mask: Optional[Tensor], a_mask: Optional[Tensor],
                   fhw: Tuple[int, int, int], audio_frames: int, dtype=torch.float32, mlx_bf16: bool = True):
    """LTX2Transformer.callAsFunction (T/LTX2Transformer.swift:240-392) with LTX2TransformerBlock (:174-297).
    v_latent [B,N,128], a_latent [B,Ta,128], contexts [B,S,3840], sigmas [B] (v_sigma may be [B,N]: the image-to-video mode
    feeds sigma * (1 - mask) per video token, P/LTXPipeline.swift:1293-1298, and every video-side embedder then works per
    token).  Returns (video velocity [B,N,128], audio velocity [B,Ta,128])."""
    F, H, W = fhw
    D, Da, eps = cfg.inner_dim, av.audio_dim, cfg.norm_eps
    B = v_latent.shape[0]
    vl, al, vc, ac = v_latent.to(dtype), a_latent.to(dtype), v_context.to(dtype), a_context.to(dtype)
    if mlx_bf16:
        vl, al, vc, ac = bf16_round(vl), bf16_round(al), bf16_round(vc), bf16_round(ac)
    vx = linear(vl, w, "patchify_proj")                                              # :255
    ax = linear(al, w, "audio_patchify_proj")                                        # :262
    if mlx_bf16:
        vx, ax = bf16_round(vx), bf16_round(ax)
    tv = v_sigma.to(torch.float32) * cfg.timestep_scale_multiplier                   # :256, 263
    ta = a_sigma.to(torch.float32) * cfg.timestep_scale_multiplier
    v_ada, v_emb = _adaln_single(w, "adaln_single", tv, dtype)
    a_ada, a_emb = _adaln_single(w, "audio_adaln_single", ta, dtype)
    v_ada, a_ada = v_ada.view(B, -1, 6, D), a_ada.view(B, -1, 6, Da)      # [B, 1 | N, 6, D]: per-token sigmas in I2V mode (:275-283)
    pvc = caption_projection(w, vc, mlx_bf16)                                        # :259
    h = linear(ac, w, "audio_caption_projection.linear_1")                           # :266
    if mlx_bf16:
        h = bf16_round(h)
    h = gelu_tanh(h)
    if mlx_bf16:
        h = bf16_round(h)
    pac = linear(h, w, "audio_caption_projection.linear_2")
    if mlx_bf16:
        pac = bf16_round(pac)
    # cross-modal modulation: 4 scale/shift values + 1 gate per stream, from the stream's own timestep (:275-298)
    cv_ss, _ = _adaln_single(w, "av_ca_video_scale_shift_adaln_single", tv, dtype)
    cv_g, _ = _adaln_single(w, "av_ca_a2v_gate_adaln_single", tv, dtype)
    ca_ss, _ = _adaln_single(w, "av_ca_audio_scale_shift_adaln_single", ta, dtype)
    ca_g, _ = _adaln_single(w, "av_ca_v2a_gate_adaln_single", ta, dtype)
    cv = torch.cat([cv_ss.view(B, -1, 4, D), cv_g.view(B, -1, 1, D)], dim=2)
    ca = torch.cat([ca_ss.view(B, -1, 4, Da), ca_g.view(B, -1, 1, Da)], dim=2)
    vbias = None if v_mask is None else ((1.0 - v_mask.to(dtype)) * -10000.0).view(B, 1, 1, -1)   # :394-403
    abias = None if a_mask is None else ((1.0 - a_mask.to(dtype)) * -10000.0).view(B, 1, 1, -1)
    v_rope = rope_table(cfg, F, H, W)                                                # :138-160
    a_grid = audio_position_grid(audio_frames)
    a_rope = rope_table_nd(a_grid, Da, av.audio_heads, cfg.rope_theta, [av.audio_max_pos])      # :162-186
    v_tgrid = position_grid(F, H, W)[0:1]                                            # temporal coordinate only (:198-212)
    xv_rope = rope_table_nd(v_tgrid, Da, av.audio_heads, cfg.rope_theta, [av.audio_max_pos])
    xa_rope = rope_table_nd(a_grid, Da, av.audio_heads, cfg.rope_theta, [av.audio_max_pos])     # :216-229
    Hv, Ha = cfg.num_heads, av.audio_heads
    for i in range(cfg.num_layers):
        p = f"transformer_blocks.{i}"
        vs = w[p + ".scale_shift_table"].to(dtype).view(1, 1, 6, D) + v_ada          # T/LTX2TransformerBlock.swift:183-206
        as_ = w[p + ".audio_scale_shift_table"].to(dtype).view(1, 1, 6, Da) + a_ada
        # 1-2: self-attention on both streams (:208-216)
        n = rms_norm(vx, w[p + ".norm1.weight"], eps) * (1 + vs[:, :, 1]) + vs[:, :, 0]
        vx = vx + _av_attention(w, p + ".attn1", n, None, Hv, eps, v_rope) * vs[:, :, 2]
        n = rms_norm(ax, w[p + ".audio_norm1.weight"], eps) * (1 + as_[:, :, 1]) + as_[:, :, 0]
        ax = ax + _av_attention(w, p + ".audio_attn1", n, None, Ha, eps, a_rope) * as_[:, :, 2]
        # 3-4: text cross-attention, learned RMSNorm in front, no RoPE, no gate (:218-226)
        vx = vx + _av_attention(w, p + ".attn2", rms_norm(vx, w[p + ".norm2.weight"], eps), pvc, Hv, eps, bias=vbias,
                                bf16_kv=mlx_bf16)
        ax = ax + _av_attention(w, p + ".audio_attn2", rms_norm(ax, w[p + ".audio_norm2.weight"], eps), pac, Ha, eps, bias=abias,
                                bf16_kv=mlx_bf16)
        # 5-6: cross-modal attention; rows: a2v_scale, a2v_shift, v2a_scale, v2a_shift, gate (:228-271)
        vca = w[p + ".scale_shift_table_a2v_ca_video"].to(dtype).view(1, 1, 5, D) + cv
        aca = w[p + ".scale_shift_table_a2v_ca_audio"].to(dtype).view(1, 1, 5, Da) + ca
        nv = rms_norm(vx, w[p + ".audio_to_video_norm.weight"], eps)
        na = rms_norm(ax, w[p + ".video_to_audio_norm.weight"], eps)
        a2v = _av_attention(w, p + ".audio_to_video_attn", nv * (1 + vca[:, :, 0]) + vca[:, :, 1],
                            na * (1 + aca[:, :, 0]) + aca[:, :, 1], Ha, eps, xv_rope, xa_rope)
        v2a = _av_attention(w, p + ".video_to_audio_attn", na * (1 + aca[:, :, 2]) + aca[:, :, 3],
                            nv * (1 + vca[:, :, 2]) + vca[:, :, 3], Ha, eps, xa_rope, xv_rope)
        vx = vx + a2v * vca[:, :, 4]
        ax = ax + v2a * aca[:, :, 4]
        # 7-8: feed-forward (:273-281)
        n = rms_norm(vx, w[p + ".norm3.weight"], eps) * (1 + vs[:, :, 4]) + vs[:, :, 3]
        vx = vx + linear(gelu_tanh(linear(n, w, p + ".ff.project_in.proj")), w, p + ".ff.project_out") * vs[:, :, 5]
        n = rms_norm(ax, w[p + ".audio_norm3.weight"], eps) * (1 + as_[:, :, 4]) + as_[:, :, 3]
        ax = ax + linear(gelu_tanh(linear(n, w, p + ".audio_ff.project_in.proj")), w, p + ".audio_ff.project_out") * as_[:, :, 5]

    def head(x, table, emb, proj):                                                    # T/LTX2Transformer.swift:370-388
        o = w[table].to(dtype).view(1, 1, 2, -1) + emb.view(B, -1, 1, x.shape[-1])
        mu = x.mean(-1, keepdim=True)
        y = (x - mu) * torch.rsqrt((x - mu).pow(2).mean(-1, keepdim=True) + eps)
        return linear(y * (1 + o[:, :, 1]) + o[:, :, 0], w, proj)

    return head(vx, "scale_shift_table", v_emb, "proj_out"), head(ax, "audio_scale_shift_table", a_emb, "audio_proj_out")


def av_denoise_loop(w, cfg: DiTConfig, av: AVConfig, v_noise: Tensor, a_noise: Tensor, v_ctx: Tensor, a_ctx: Tensor,
                    mask: Optional[Tensor], sigmas: Sequence[float], neg_v_ctx: Optional[Tensor] = None,
                    neg_a_ctx: Optional[Tensor] = None, neg_mask: Optional[Tensor] = None, cfg_scale: float = 1.0,
                    phi: float = 0.0, image_latent: Optional[Tensor] = None, inject_noise: Optional[Sequence[Tensor]] = None,
                    image_cond_noise_scale: float = 0.0):
    """Audio + video denoise loop (P/LTXPipeline.swift:1277-1404): v_noise [1,C,F,H,W], a_noise [1,Ta,128] (packed audio
    latent).  Per step: one dual forward (two with CFG), video = CFG (+ rescale) + scheduler.step, audio = CFG + plain Euler
    a += (sigma' - sigma) v (:1402).  image_latent [1,C,1,H,W] selects the image-to-video branch: frame 0 := image latent
    (:1262-1274), optionally re-noised every step with inject_noise[step] * scale * sigma^2 (:1288-1292; the draws are passed in
    as data), video timesteps = sigma * (1 - mask) per token (:1294-1298), Euler only on frames 1+ (:1381-1391).
    Returns (video latent [1,C,F,H,W], audio latent [1,Ta,128])."""
    fhw = tuple(v_noise.shape[2:])
    Ta = a_noise.shape[1]
    v_lat = v_noise.float() * sigmas[0]                                   # :1255-1259
    a_lat = a_noise.float() * sigmas[0]
    cond_mask = None
    if image_latent is not None:
        v_lat = v_lat.clone()
        v_lat[:, :, 0:1] = image_latent.float()
        cond_mask = torch.zeros(1, fhw[0] * fhw[1] * fhw[2])
        cond_mask[:, :fhw[1] * fhw[2]] = 1.0
    use_cfg = cfg_scale > 1.0 and neg_v_ctx is not None
    for step in range(len(sigmas) - 1):
        sg, sn = sigmas[step], sigmas[step + 1]
        ts = torch.tensor([sg], dtype=torch.float32)
        if image_latent is not None and image_cond_noise_scale > 0 and sg > 0 and inject_noise is not None:
            v_lat[:, :, 0:1] = image_latent.float() + image_cond_noise_scale * inject_noise[step].float() * (sg * sg)
        tsv = ts if cond_mask is None else torch.tensor(sg, dtype=torch.float32) * (1 - cond_mask)
        tok = patchify(v_lat)
        pv, pa = av_dit_forward(w, cfg, av, tok, a_lat, v_ctx, a_ctx, tsv, ts, mask, mask, fhw, Ta)
        vv, va = unpatchify(pv, fhw).float(), pa.float()
        if use_cfg:
            nv, na = av_dit_forward(w, cfg, av, tok, a_lat, neg_v_ctx, neg_a_ctx, tsv, ts, neg_mask, neg_mask, fhw, Ta)
            nvv = unpatchify(nv, fhw).float()
            cond_v = vv
            vv = apply_cfg(nvv, cond_v, cfg_scale)
            va = apply_cfg(na.float(), va, cfg_scale)
            if phi > 0:
                vv = guidance_rescale(vv, cond_v, phi)
        if image_latent is not None:                                      # :1381-1391: frame 0 stays as it is
            v_lat = torch.cat([v_lat[:, :, 0:1], euler_step(v_lat[:, :, 1:], vv[:, :, 1:], sg, sn)], dim=2)
        else:
            v_lat = euler_step(v_lat, vv, sg, sn)
        a_lat = a_lat + (sn - sg) * va
    return v_lat, a_lat


# ----------------------------------------------------------------------------------------------
# latent utils, guidance, scheduler (P/LatentUtils.swift, S/LTXScheduler.swift, P/LTXPipeline.swift:800-956)
# ----------------------------------------------------------------------------------------------
def patchify(latent: Tensor) -> Tensor:                  # P/LatentUtils.swift:20-36  (B,C,F,H,W)->(B,N,C)
    B, C, F, H, W = latent.shape
    return latent.permute(0, 2, 3, 4, 1).reshape(B, F * H * W, C)


def unpatchify(x: Tensor, fhw: Tuple[int, int, int]) -> Tensor:   # :38-54
    B, N, C = x.shape
    F, H, W = fhw
    return x.view(B, F, H, W, C).permute(0, 4, 1, 2, 3).contiguous()


def latent_shape(frames: int, height: int, width: int) -> Tuple[int, int, int]:
    """P/VideoLatentShape.swift:35-41."""
    return (frames - 1) // 8 + 1, height // 32, width // 32


DISTILLED_SIGMA_VALUES = [1.0, 0.99375, 0.9875, 0.98125, 0.975, 0.909375, 0.725, 0.421875, 0.0]   # S/LTXScheduler.swift:18-28
STAGE_2_DISTILLED_SIGMA_VALUES = [0.909375, 0.725, 0.421875, 0.0]                                   # :31-36
BASE_SHIFT_ANCHOR, MAX_SHIFT_ANCHOR = 1024, 4096


def _f32(x: float) -> float:
    return float(torch.tensor(x, dtype=torch.float32))


def set_timesteps(num_steps: int, distilled: bool, token_count: Optional[int], max_shift: float = 2.05,
                  base_shift: float = 0.95, stretch: bool = True, terminal: float = 0.1) -> List[float]:
    """S/LTXScheduler.swift:74-182, evaluated in float32 like the Swift `Float` code."""
    f32 = torch.float32
    T = lambda v: torch.tensor(v, dtype=f32)
    if distilled:
        s = T([v for v in DISTILLED_SIGMA_VALUES if v > 0])
        if token_count is not None:
            tok = min(token_count, MAX_SHIFT_ANCHOR)
            mm = (T(max_shift) - T(base_shift)) / (T(float(MAX_SHIFT_ANCHOR)) - T(float(BASE_SHIFT_ANCHOR)))
            b = T(base_shift) - mm * T(float(BASE_SHIFT_ANCHOR))
            mu = T(float(tok)) * mm + b
            e = torch.exp(mu)
            shifted = e / (e + (1.0 / s - 1.0))
            s = torch.where((s == 0) | (s == 1.0), s, shifted)
            if stretch:
                om = 1.0 - s
                last = om[-1]
                if last > 0:
                    scale = last / (1.0 - T(terminal))
                    s = torch.where(s == 0, torch.zeros_like(s), 1.0 - (1.0 - s) / scale)
        return [float(v) for v in s] + [0.0]
    tok = min(token_count if token_count is not None else MAX_SHIFT_ANCHOR, MAX_SHIFT_ANCHOR)
    s = 1.0 - torch.arange(num_steps + 1, dtype=f32) / float(num_steps)
    mm = (T(max_shift) - T(base_shift)) / (T(float(MAX_SHIFT_ANCHOR)) - T(float(BASE_SHIFT_ANCHOR)))
    b = T(base_shift) - mm * T(float(BASE_SHIFT_ANCHOR))
    e = torch.exp(T(float(tok)) * mm + b)
    safe = torch.where(s == 0, torch.ones_like(s), s)
    s = torch.where(s == 0, torch.zeros_like(s), e / (e + (1.0 / safe - 1.0)))
    if stretch and num_steps > 0:
        om = 1.0 - s
        scale = om[num_steps - 1] / (1.0 - T(terminal))
        s = torch.where(s == 0, torch.zeros_like(s), 1.0 - om / scale)
    return [float(v) for v in s]


def apply_cfg(uncond: Tensor, cond: Tensor, scale: float) -> Tensor:      # P/LatentUtils.swift:131-141
    return cond + (scale - 1.0) * (cond - uncond)


def guidance_rescale(cfg_out: Tensor, cond: Tensor, phi: float) -> Tensor:   # :164-183
    if phi <= 0:
        return cfg_out
    dims = list(range(1, cfg_out.ndim))
    s_cfg = torch.sqrt(cfg_out.var(dim=dims, unbiased=False, keepdim=True) + 1e-8)
    s_cond = torch.sqrt(cond.var(dim=dims, unbiased=False, keepdim=True) + 1e-8)
    return phi * (cfg_out * (s_cond / s_cfg)) + (1.0 - phi) * cfg_out


def euler_step(latent: Tensor, velocity: Tensor, sigma: float, sigma_next: float) -> Tensor:   # S/LTXScheduler.swift:305-327
    den = latent - sigma * velocity
    if sigma_next > 0:
        return den + sigma_next * (latent - den) / sigma
    return den


def guided_euler_step(latent: Tensor, v_cond: Tensor, v_uncond: Optional[Tensor], v_stg: Optional[Tensor],
                      v_prev: Optional[Tensor], cfg_scale: float, phi: float, stg_scale: float, ge_gamma: float,
                      sigma: float, sigma_next: float) -> Tuple[Tensor, Tensor]:
    """P/LTXPipeline.swift:861-935 in fp32.  Returns (new latent, velocity used) ; velocity feeds GE next step."""
    v = v_cond.float()
    if v_uncond is not None:
        v = apply_cfg(v_uncond.float(), v_cond.float(), cfg_scale)
        v = guidance_rescale(v, v_cond.float(), phi)
    if v_stg is not None and stg_scale > 0:
        v = v + stg_scale * (v - v_stg.float())                # :920
    if ge_gamma > 0 and v_prev is not None:
        v = ge_gamma * (v - v_prev) + v_prev                   # :924-927
    return euler_step(latent.float(), v, sigma, sigma_next), v


def denoise_loop(w, cfg: DiTConfig, noise: Tensor, context: Tensor, mask: Optional[Tensor], sigmas: Sequence[float],
                 neg_context: Optional[Tensor] = None, neg_mask: Optional[Tensor] = None, cfg_scale: float = 1.0,
                 phi: float = 0.0, stg_scale: float = 0.0, stg_blocks: Sequence[int] = (29,), ge_gamma: float = 0.0,
                 mlx_bf16: bool = True, dtype=torch.float32, return_velocities: bool = False,
                 frame0_conditioned: bool = False, init_latent: Optional[Tensor] = None):
    """P/LTXPipeline.swift:793-956 (generateVideo step loop).  noise [1,C,F,H,W] fp32.
    frame0_conditioned = the image-to-video variant of denoise() (:2237-2252, 2344-2357): tokens of latent frame 0 are a
    clean conditioning frame -> per-token timesteps sigma * (1 - mask), Euler only on frames 1+.  init_latent (if given) is
    the starting latent (frame 0 already holding the encoded image)."""
    latent = noise.float() * sigmas[0] if init_latent is None else init_latent.float().clone()   # :793
    fhw = tuple(noise.shape[2:])
    v_prev = None
    vels = []
    for step in range(len(sigmas) - 1):
        sg, sn = sigmas[step], sigmas[step + 1]
        tok = patchify(latent)
        ts = torch.tensor([sg], dtype=torch.float32)
        if frame0_conditioned:
            m = torch.zeros(1, fhw[0] * fhw[1] * fhw[2])
            m[:, : fhw[1] * fhw[2]] = 1.0
            ts = sg * (1.0 - m)
        vc = unpatchify(dit_forward(w, cfg, tok, context, ts, mask, fhw, dtype=dtype, mlx_bf16=mlx_bf16), fhw).float()
        vu = vs = None
        if cfg_scale > 1.0 and neg_context is not None:
            vu = unpatchify(dit_forward(w, cfg, tok, neg_context, ts, neg_mask, fhw, dtype=dtype, mlx_bf16=mlx_bf16), fhw).float()
        if stg_scale > 0:
            vs = unpatchify(dit_forward(w, cfg, tok, context, ts, mask, fhw, stg_blocks=stg_blocks, skip_self_attn=True,
                                        dtype=dtype, mlx_bf16=mlx_bf16), fhw).float()
        new_latent, v_prev = guided_euler_step(latent, vc, vu, vs, v_prev, cfg_scale, phi, stg_scale, ge_gamma, sg, sn)
        if frame0_conditioned:                                   # slice Euler: frame 0 is re-attached unchanged
            new_latent[:, :, :1] = latent[:, :, :1]
        latent = new_latent
        vels.append(v_prev)
    return (latent, vels) if return_velocities else latent


# ----------------------------------------------------------------------------------------------
# video VAE decoder (V/VideoConvolution.swift:202-348, V/VideoDecoder.swift)
# ----------------------------------------------------------------------------------------------
def conv3d_full(x: Tensor, weight: Tensor, bias: Tensor, causal: bool = False, spatial_pad: str = "reflect") -> Tensor:
    """V/VideoConvolution.swift:238-347.  x [B,C,T,H,W]; 3x3x3 cross-correlation, pad H/W by 1 (reflect | zeros |
    replicate), pad T by frame replication: causal ? 2 x first : 1 x first + 1 x last."""
    mode = {"reflect": "reflect", "zeros": "constant", "replicate": "replicate"}[spatial_pad]
    B, C, T, H, W = x.shape
    x2 = torch.nn.functional.pad(x.reshape(B, C * T, H, W), (1, 1, 1, 1), mode=mode).view(B, C, T, H + 2, W + 2)
    if causal:
        x2 = torch.cat([x2[:, :, :1], x2[:, :, :1], x2], dim=2)
    else:
        x2 = torch.cat([x2[:, :, :1], x2, x2[:, :, -1:]], dim=2)
    return torch.nn.functional.conv3d(x2, weight.to(x.dtype), bias.to(x.dtype))


def pixel_norm(x: Tensor, eps: float = 1e-8) -> Tensor:          # V/VideoDecoder.swift:29-32
    return x / torch.sqrt(x.pow(2).mean(dim=1, keepdim=True) + eps)


def depth_to_space(x: Tensor, c_out: int) -> Tensor:            # :201-212
    B, _, T, H, W = x.shape
    o = x.reshape(B, c_out, 2, 2, 2, T, H, W).permute(0, 1, 5, 2, 6, 3, 7, 4)
    return o.reshape(B, c_out, 2 * T, 2 * H, 2 * W)


def vae_time_embed(w, prefix: str, t: Tensor) -> Tensor:         # :37-52, 11-24
    e = sinusoidal_embedding(t, 256).to(t.dtype if t.dtype.is_floating_point else torch.float32)
    h = e @ w[prefix + ".timestep_embedder.linear_1.weight"].t() + w[prefix + ".timestep_embedder.linear_1.bias"]
    h = silu(h)
    return h @ w[prefix + ".timestep_embedder.linear_2.weight"].t() + w[prefix + ".timestep_embedder.linear_2.bias"]


def vae_resblock(w, prefix: str, x: Tensor, causal: bool, time_emb: Optional[Tensor]) -> Tensor:   # :93-130
    C = x.shape[1]
    tbl = w[prefix + ".scale_shift_table"].to(x.dtype).unsqueeze(0)                 # rows shift1, scale1, shift2, scale2
    if time_emb is not None:
        tbl = tbl + time_emb.view(-1, 4, C).to(x.dtype)
    sh1, sc1, sh2, sc2 = [tbl[:, r].reshape(-1, C, 1, 1, 1) for r in range(4)]
    h = silu(pixel_norm(x) * (sc1 + 1) + sh1)
    h = conv3d_full(h, w[prefix + ".conv1.conv.weight"], w[prefix + ".conv1.conv.bias"], causal)
    h = silu(pixel_norm(h) * (sc2 + 1) + sh2)
    h = conv3d_full(h, w[prefix + ".conv2.conv.weight"], w[prefix + ".conv2.conv.bias"], causal)
    return h + x


def vae_d2s_up(w, prefix: str, x: Tensor, causal: bool) -> Tensor:                  # :214-251
    C = x.shape[1]
    r = depth_to_space(x, C // 8)[:, :, 1:]
    r = torch.cat([r, r, r, r], dim=1)
    h = conv3d_full(x, w[prefix + ".conv.conv.weight"], w[prefix + ".conv.conv.bias"], causal)
    h = depth_to_space(h, C // 2)[:, :, 1:]
    return h + r


def vae_unpatchify(x: Tensor, p: int = 4) -> Tensor:                                # :257-275
    B, CP, T, H, W = x.shape
    c = CP // (p * p)
    o = x.reshape(B, c, 1, p, p, T, H, W).permute(0, 1, 5, 2, 6, 4, 7, 3)
    return o.reshape(B, c, T, H * p, W * p)


def vae_decode(w, cfg: VAEConfig, latent: Tensor, timestep: Optional[float] = None,
               decode_noise: Optional[Tensor] = None, dtype=torch.float32, return_stages: bool = False):
    """V/VideoDecoder.swift:358-449.  latent [1,128,F',H',W'] -> [1,3,8(F'-1)+1,32H',32W']."""
    x = latent.to(dtype)
    w = {k: v.to(dtype) for k, v in w.items()}
    B = x.shape[0]
    t = None
    if timestep is not None:                                                         # :368-375
        assert decode_noise is not None, "noise must be passed in explicitly (SURVEY H7)"
        x = decode_noise.to(dtype) * 0.025 + (1.0 - 0.025) * x
        t = torch.full((B,), float(timestep), dtype=dtype) * w["timestep_scale_multiplier"]
    x = x * w["std_of_means"].view(1, -1, 1, 1, 1) + w["mean_of_means"].view(1, -1, 1, 1, 1)   # :379-381
    x = conv3d_full(x, w["conv_in.conv.weight"], w["conv_in.conv.bias"], cfg.causal)            # :385
    stages = []
    chans = cfg.stage_channels
    for s in range(len(chans)):
        blk = f"up_blocks_{2 * s}"
        te = vae_time_embed(w, blk + ".time_embedder", t) if t is not None else None
        for j in range(cfg.blocks_per_stage):
            x = vae_resblock(w, f"{blk}.res_blocks.{j}", x, cfg.causal, te)
        stages.append(x)
        if s < len(chans) - 1:
            x = vae_d2s_up(w, f"up_blocks_{2 * s + 1}", x, cfg.causal)
            stages.append(x)
    C = chans[-1]
    tbl = w["last_scale_shift_table"].unsqueeze(0)                                    # rows shift, scale :419-436
    if t is not None:
        tbl = tbl + vae_time_embed(w, "last_time_embedder", t).view(B, 2, C)
    x = pixel_norm(x) * (tbl[:, 1].reshape(-1, C, 1, 1, 1) + 1) + tbl[:, 0].reshape(-1, C, 1, 1, 1)
    x = silu(x)
    x = conv3d_full(x, w["conv_out.conv.weight"], w["conv_out.conv.bias"], cfg.causal)          # :439
    x = vae_unpatchify(x, cfg.patch_size)                                                        # :444
    return (x, stages) if return_stages else x


def decode_video(w, cfg: VAEConfig, latent: Tensor, timestep: Optional[float] = None,
                 decode_noise: Optional[Tensor] = None, dtype=torch.float32, temporal_tile_size: int = 0,
                 temporal_tile_overlap: int = 1) -> Tensor:
    """V/VideoDecoder.swift:466-508: frames [F,H,W,3] in [0,1]; temporal tiling (:482-494) when the latent has more frames
    than one tile."""
    if latent.ndim == 4:
        latent = latent.unsqueeze(0)
    if decode_noise is not None and decode_noise.ndim == 4:
        decode_noise = decode_noise.unsqueeze(0)
    if temporal_tile_size > 0 and latent.shape[2] > temporal_tile_size:
        x = decode_with_temporal_tiling(w, cfg, latent, timestep, decode_noise, temporal_tile_size, temporal_tile_overlap, dtype)
    else:
        x = vae_decode(w, cfg, latent, timestep, decode_noise, dtype)
    x = torch.clamp((x + 1.0) / 2.0, 0.0, 1.0)
    return x[0].permute(1, 2, 3, 0).contiguous()


def decode_with_temporal_tiling(w, cfg: VAEConfig, latent: Tensor, timestep: Optional[float], decode_noise: Optional[Tensor],
                                tile_size: int, overlap: int, dtype=torch.float32) -> Tensor:
    """decodeWithTemporalTiling (V/VideoDecoder.swift:517-602): chunks [start, start + tile) at stride tile - overlap decoded
    independently (:534-551); each next chunk's first 8 * overlap frames are cross-faded into the result's last ones with
    weights j / (8 * overlap) when both sides are longer than the overlap (:572-585), else concatenated (:586-589).  Raw
    decoder output [1,3,F,H,W] (the caller normalises and clips, :595-599).  The reference draws fresh decode noise per chunk;
    here each chunk takes its slice of the caller's noise (noise is data)."""
    total = latent.shape[2]
    stride = tile_size - overlap
    po = 8 * overlap
    chunks = []
    start = 0
    while start < total:
        end = min(start + tile_size, total)
        nz = None if decode_noise is None else decode_noise[:, :, start:end]
        chunks.append(vae_decode(w, cfg, latent[:, :, start:end], timestep, nz, dtype))
        if end >= total:
            break
        start += stride
    result = chunks[0]
    for nxt in chunks[1:]:
        rf, nf = result.shape[2], nxt.shape[2]
        if 0 < po < rf and po < nf:
            wts = (torch.arange(po, dtype=torch.float32) / float(po)).to(dtype).view(1, 1, po, 1, 1)
            blended = result[:, :, rf - po:] * (1 - wts) + nxt[:, :, :po] * wts
            result = torch.cat([result[:, :, :rf - po], blended, nxt[:, :, po:]], dim=2)
        else:
            result = torch.cat([result, nxt], dim=2)
    return result


# ----------------------------------------------------------------------------------------------
# video VAE encoder (V/VideoEncoder.swift) -- SURVEY 8f-4: image / video -> latent for image-to-video conditioning
# ----------------------------------------------------------------------------------------------
@dataclass
class EncoderConfig:
    """Channel plan of VideoEncoder.init (V/VideoEncoder.swift:222-268): conv_in 48->c0, four down blocks
    (resnets, space-to-depth factor, output channels), mid block, conv_out -> latent_channels + 1."""
    base_channels: int = 128
    resnets: Tuple[int, ...] = (4, 6, 6, 2)
    factors: Tuple[Tuple[int, int, int], ...] = ((1, 2, 2), (2, 1, 1), (2, 2, 2), (2, 2, 2))
    mid_resnets: int = 2
    latent_channels: int = 128
    causal: bool = True

    @property
    def stage_channels(self) -> List[int]:
        return [self.base_channels * (2 ** i) for i in range(5)]


def make_encoder_weights(cfg: EncoderConfig, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init encoder weights under the post-mapping names of mapVAEEncoderWeights
    (U/ModelDownloader.swift:1224-1280)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, Tensor] = {}

    def conv(name: str, cout: int, cin: int):
        w[name + ".conv.weight"] = torch.randn(cout, cin, 3, 3, 3, generator=g) / math.sqrt(27 * cin)
        w[name + ".conv.bias"] = torch.randn(cout, generator=g) * 0.02

    ch = cfg.stage_channels
    conv("conv_in", ch[0], 48)
    for i in range(4):
        for j in range(cfg.resnets[i]):
            conv(f"down_blocks_{i}.resnets.resnets.{j}.conv1", ch[i], ch[i])
            conv(f"down_blocks_{i}.resnets.resnets.{j}.conv2", ch[i], ch[i])
        ft, fh, fw = cfg.factors[i]
        conv(f"down_blocks_{i}.downsamplers.conv", ch[i + 1] // (ft * fh * fw), ch[i])
    for j in range(cfg.mid_resnets):
        conv(f"mid_block.resnets.{j}.conv1", ch[4], ch[4])
        conv(f"mid_block.resnets.{j}.conv2", ch[4], ch[4])
    conv("conv_out", cfg.latent_channels + 1, ch[4])
    return w


def encoder_patchify(x: Tensor) -> Tensor:                      # V/VideoEncoder.swift:13-32 (pW before pH)
    B, C, T, H, W = x.shape
    o = x.reshape(B, C, T, H // 4, 4, W // 4, 4).permute(0, 1, 6, 4, 2, 3, 5)
    return o.reshape(B, C * 16, T, H // 4, W // 4)


def space_to_depth(x: Tensor, factor: Tuple[int, int, int]) -> Tensor:   # :38-66 (front-pad T with the first frame)
    ft, fh, fw = factor
    B, C, T, H, W = x.shape
    if T % ft != 0:
        pad = ft - T % ft
        x = torch.cat([x[:, :, :1].expand(B, C, pad, H, W), x], dim=2)
        T = x.shape[2]
    o = x.reshape(B, C, T // ft, ft, H // fh, fh, W // fw, fw).permute(0, 1, 3, 5, 7, 2, 4, 6)
    return o.reshape(B, C * ft * fh * fw, T // ft, H // fh, W // fw)


def encoder_resblock(w, prefix: str, x: Tensor, causal: bool) -> Tensor:       # :72-101
    h = silu(pixel_norm(x))
    h = conv3d_full(h, w[prefix + ".conv1.conv.weight"], w[prefix + ".conv1.conv.bias"], causal, "zeros")
    h = silu(pixel_norm(h))
    h = conv3d_full(h, w[prefix + ".conv2.conv.weight"], w[prefix + ".conv2.conv.bias"], causal, "zeros")
    return h + x


def encoder_downsample(w, prefix: str, x: Tensor, factor, out_channels: int, causal: bool) -> Tensor:   # :127-166
    main = space_to_depth(conv3d_full(x, w[prefix + ".conv.conv.weight"], w[prefix + ".conv.conv.bias"], causal, "zeros"), factor)
    r = space_to_depth(x, factor)
    B, Cr, T2, H2, W2 = r.shape
    r = r.reshape(B, out_channels, Cr // out_channels, T2, H2, W2).mean(dim=2)
    return main + r


def vae_encode(w, cfg: EncoderConfig, pixels: Tensor, dtype=torch.float32, return_stages: bool = False):
    """VideoEncoder.callAsFunction (V/VideoEncoder.swift:270-312): pixels [B,3,T,H,W] -> latent mean
    [B,128,T',H/32,W/32] (the log-variance channel is dropped)."""
    w = {k: v.to(dtype) for k, v in w.items()}
    ch = cfg.stage_channels
    h = encoder_patchify(pixels.to(dtype))
    h = conv3d_full(h, w["conv_in.conv.weight"], w["conv_in.conv.bias"], cfg.causal, "zeros")
    stages = [h]
    for i in range(4):
        for j in range(cfg.resnets[i]):
            h = encoder_resblock(w, f"down_blocks_{i}.resnets.resnets.{j}", h, cfg.causal)
        h = encoder_downsample(w, f"down_blocks_{i}.downsamplers", h, cfg.factors[i], ch[i + 1], cfg.causal)
        stages.append(h)
    for j in range(cfg.mid_resnets):
        h = encoder_resblock(w, f"mid_block.resnets.{j}", h, cfg.causal)
    h = silu(pixel_norm(h))
    h = conv3d_full(h, w["conv_out.conv.weight"], w["conv_out.conv.bias"], cfg.causal, "zeros")
    h = h[:, :cfg.latent_channels]
    return (h, stages) if return_stages else h


def encode_image_latent(w, cfg: EncoderConfig, pixels: Tensor, mean: Tensor, std: Tensor) -> Tensor:
    """encodeImage (P/LTXPipeline.swift:1902-1932): encode, then normalise with the decoder's per-channel statistics."""
    z = vae_encode(w, cfg, pixels)
    return (z - mean.view(1, -1, 1, 1, 1)) / std.view(1, -1, 1, 1, 1)


# ----------------------------------------------------------------------------------------------
# latent spatial upscaler + AdaIN + re-noise (Models/Upscaler/SpatialUpscaler.swift, P/LatentUtils.swift:201-227,
# P/LTXPipeline.swift:2594-2647) -- SURVEY 8f-3: the glue between the two stages of the two-stage pipeline
# ----------------------------------------------------------------------------------------------
def make_upscaler_weights(mid: int = 1024, in_ch: int = 128, blocks: int = 4, seed: int = 0) -> Dict[str, Tensor]:
    """Random-init SpatialUpscaler weights in the checkpoint (PyTorch) layout the reference's loader reads
    (SpatialUpscaler.swift:262-300): Conv3d (O,I,3,3,3), Conv2d (O,I,3,3)."""
    g = torch.Generator().manual_seed(seed)
    w: Dict[str, Tensor] = {}

    def conv3(name, cout, cin):
        w[name + ".weight"] = torch.randn(cout, cin, 3, 3, 3, generator=g) / math.sqrt(27 * cin)
        w[name + ".bias"] = torch.randn(cout, generator=g) * 0.02

    def norm(name, c):
        w[name + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        w[name + ".bias"] = 0.1 * torch.randn(c, generator=g)

    conv3("initial_conv", mid, in_ch)
    norm("initial_norm", mid)
    for grp in ("res_blocks", "post_upsample_res_blocks"):
        for i in range(blocks):
            conv3(f"{grp}.{i}.conv1", mid, mid)
            norm(f"{grp}.{i}.norm1", mid)
            conv3(f"{grp}.{i}.conv2", mid, mid)
            norm(f"{grp}.{i}.norm2", mid)
    w["upsampler.conv.weight"] = torch.randn(4 * mid, mid, 3, 3, generator=g) / math.sqrt(9 * mid)
    w["upsampler.conv.bias"] = torch.randn(4 * mid, generator=g) * 0.02
    conv3("final_conv", in_ch, mid)
    return w


def _group_norm(x: Tensor, weight: Tensor, bias: Tensor, groups: int = 32, eps: float = 1e-5) -> Tensor:
    """UpscalerGroupNorm3D (SpatialUpscaler.swift:12-58): statistics over (D,H,W, C/groups), population variance."""
    B, C = x.shape[:2]
    y = x.reshape(B, groups, -1)
    mu = y.mean(dim=2, keepdim=True)
    var = y.var(dim=2, unbiased=False, keepdim=True)
    y = ((y - mu) / torch.sqrt(var + eps)).reshape(x.shape)
    return y * weight.view(1, C, 1, 1, 1) + bias.view(1, C, 1, 1, 1)


def _upscaler_resblock(w, p: str, x: Tensor) -> Tensor:           # SpatialUpscaler.swift:62-107
    h = torch.nn.functional.conv3d(x, w[p + ".conv1.weight"], w[p + ".conv1.bias"], padding=1)
    h = silu(_group_norm(h, w[p + ".norm1.weight"], w[p + ".norm1.bias"]))
    h = torch.nn.functional.conv3d(h, w[p + ".conv2.weight"], w[p + ".conv2.bias"], padding=1)
    h = _group_norm(h, w[p + ".norm2.weight"], w[p + ".norm2.bias"])
    return silu(h + x)


def spatial_upscaler(w, x: Tensor, blocks: int = 4, dtype=torch.float32) -> Tensor:
    """SpatialUpscaler.callAsFunction (SpatialUpscaler.swift:215-258): [B,128,F,H,W] -> [B,128,F,2H,2W].  Conv3d zero
    padding 1; per-frame Conv2d + PixelShuffle(2) (:111-163: channel = c*4 + i*2 + j -> (2h+i, 2w+j))."""
    w = {k: v.to(dtype) for k, v in w.items()}
    h = torch.nn.functional.conv3d(x.to(dtype), w["initial_conv.weight"], w["initial_conv.bias"], padding=1)
    h = silu(_group_norm(h, w["initial_norm.weight"], w["initial_norm.bias"]))
    for i in range(blocks):
        h = _upscaler_resblock(w, f"res_blocks.{i}", h)
    B, C, D, H, W = h.shape
    fr = h.permute(0, 2, 1, 3, 4).reshape(B * D, C, H, W)
    fr = torch.nn.functional.conv2d(fr, w["upsampler.conv.weight"], w["upsampler.conv.bias"], padding=1)
    fr = torch.nn.functional.pixel_shuffle(fr, 2)
    h = fr.reshape(B, D, C, 2 * H, 2 * W).permute(0, 2, 1, 3, 4)
    for i in range(blocks):
        h = _upscaler_resblock(w, f"post_upsample_res_blocks.{i}", h)
    return torch.nn.functional.conv3d(h, w["final_conv.weight"], w["final_conv.bias"], padding=1)


def upsample_latents(w, latent: Tensor, mean: Tensor, std: Tensor, blocks: int = 4) -> Tensor:
    """upsampleLatents (SpatialUpscaler.swift:360-383) / P/LTXPipeline.swift:2604-2619: denormalise, upscale, renormalise."""
    m, s = mean.view(1, -1, 1, 1, 1), std.view(1, -1, 1, 1, 1)
    return (spatial_upscaler(w, latent * s + m, blocks) - m) / s


def adain_filter_latent(latent: Tensor, reference: Tensor, factor: float = 1.0) -> Tensor:   # P/LatentUtils.swift:201-227
    if factor <= 0:
        return latent
    lm, ls = latent.mean(dim=(2, 3, 4), keepdim=True), latent.var(dim=(2, 3, 4), unbiased=False, keepdim=True).sqrt()
    rm, rs = reference.mean(dim=(2, 3, 4), keepdim=True), reference.var(dim=(2, 3, 4), unbiased=False, keepdim=True).sqrt()
    out = (latent - lm) / (ls + 1e-8) * rs + rm
    return out if factor >= 1.0 else factor * out + (1.0 - factor) * latent


def renoise(latent: Tensor, noise: Tensor, noise_scale: float) -> Tensor:    # P/LTXPipeline.swift:2644-2647
    return noise_scale * noise + (1.0 - noise_scale) * latent


# ----------------------------------------------------------------------------------------------
# metrics used by the parity tests
# ----------------------------------------------------------------------------------------------
def rel_l2(a: Tensor, b: Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def psnr(a: Tensor, b: Tensor, peak: float = 1.0) -> float:
    mse = float((a.double() - b.double()).pow(2).mean())
    return 99.0 if mse == 0 else 10.0 * math.log10(peak * peak / mse)
