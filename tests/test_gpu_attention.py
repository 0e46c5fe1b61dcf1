"""tcgen05 flash-attention parity (ltx_op_attention) against an fp32 torch softmax(QK^T)V."""
import math

import pytest
import torch

from helpers import product, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = product().LtxContext(product().LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
    yield c
    c.close()


CASES = [  # B, H, Nq, Nk, masked
    (1, 1, 128, 128, False), (1, 2, 48, 48, False), (1, 2, 300, 300, False), (1, 2, 200, 40, True),
    (2, 2, 130, 70, True), (1, 4, 1536, 1536, False), (1, 4, 1536, 1024, True), (3, 1, 257, 129, False),
]


@pytest.mark.parametrize("B,H,Nq,Nk,masked", CASES)
def test_attention(ctx, B, H, Nq, Nk, masked):
    _run(ctx, B, H, Nq, Nk, masked, 128)


@pytest.fixture(params=["0", "1"])
def pair_mode(request, monkeypatch):
    """LTX_ATT_PAIR (read by the launcher on every call): 0 = one CTA per query tile, 1 = CTA pairs sharing the K / V tiles
    whenever there are two query tiles.  The default picks pairs from Nq = 3072 up."""
    monkeypatch.setenv("LTX_ATT_PAIR", request.param)
    return request.param


@pytest.mark.parametrize("B,H,Nq,Nk,masked,HD", [(1, 2, 300, 300, False, 128), (2, 2, 130, 70, True, 128), (1, 4, 1536, 1024, True, 128),
                                                 (3, 1, 257, 129, False, 128), (1, 4, 1536, 126, False, 64), (2, 2, 300, 70, True, 64)])
def test_attention_both_forms(ctx, pair_mode, B, H, Nq, Nk, masked, HD):
    _run(ctx, B, H, Nq, Nk, masked, HD)


@pytest.mark.parametrize("B,H,Nq,Nk,masked,HD", [(1, 2, 3072, 3072, False, 128), (1, 2, 3100, 1030, True, 128), (1, 2, 3200, 200, False, 64)])
def test_attention_long_sequences_default_form(ctx, B, H, Nq, Nk, masked, HD):
    """Nq >= 3072: the default dispatch takes the pair form (odd tile counts and a ragged last tile included)."""
    _run(ctx, B, H, Nq, Nk, masked, HD)


# head_dim 64: audio self-attention and the audio<->video cross-modal attentions of the dual block (32 heads x 64)
@pytest.mark.parametrize("B,H,Nq,Nk,masked", [(1, 2, 126, 126, False), (1, 4, 1536, 126, False), (1, 4, 126, 1536, False),
                                              (2, 2, 300, 70, True), (1, 32, 256, 384, False)])
def test_attention_head_dim_64(ctx, B, H, Nq, Nk, masked):
    _run(ctx, B, H, Nq, Nk, masked, 64)


def test_attention_logit_range(ctx):
    """Large logits (row max far above the mean, keys masked with -10000) exercise the lazy rescale and the clamp of the
    polynomial exp2 path."""
    _run(ctx, 1, 2, 384, 640, True, 128, qscale=6.0)


def _run(ctx, B, H, Nq, Nk, masked, HD, qscale=1.0):
    D = H * HD
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + H * 100 + Nq + Nk)
    q = (qscale * torch.randn(B * Nq, D, device="cuda", generator=g)).bfloat16()
    k = torch.randn(B * Nk, D, device="cuda", generator=g).bfloat16()
    v = torch.randn(B * Nk, D, device="cuda", generator=g).bfloat16()
    ldv = (Nk + 7) // 8 * 8                      # per-batch pitch; padding columns hold NaN on purpose (never read)
    vt = torch.full((D, B * ldv), float("nan"), device="cuda", dtype=torch.bfloat16)
    for b in range(B):
        vt[:, b * ldv: b * ldv + Nk] = v[b * Nk:(b + 1) * Nk].t()
    bias = None
    if masked:
        m = (torch.rand(B, Nk, device="cuda", generator=g) > 0.3).float()
        m[:, 0] = 1
        bias = ((1 - m) * -10000.0).contiguous()
    o = torch.full((B * Nq, D), float("nan"), device="cuda", dtype=torch.bfloat16)
    scale = 1 / math.sqrt(HD)
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    if HD == 128:
        ctx._check(ctx.lib.ltx_op_attention(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), ldv,
                                            bias.data_ptr() if bias is not None else None, o.data_ptr(), B, H, Nq, Nk, scale))
    else:
        ctx._check(ctx.lib.ltx_op_attention_hd(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), ldv,
                                               bias.data_ptr() if bias is not None else None, o.data_ptr(), B, H, HD, Nq, Nk, scale))
    ctx.sync()
    qh = q.float().view(B, Nq, H, HD).permute(0, 2, 1, 3)
    kh = k.float().view(B, Nk, H, HD).permute(0, 2, 1, 3)
    vh = v.float().view(B, Nk, H, HD).permute(0, 2, 1, 3)
    s = qh @ kh.transpose(-1, -2) * scale
    if bias is not None:
        s = s + bias.view(B, 1, 1, Nk)
    ref = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B * Nq, D)
    assert torch.isfinite(o.float()).all()
    assert rel_l2(o.float(), ref) <= 1e-2   # P and O are rounded to bf16


@pytest.mark.parametrize("H,HD,Nq,Nk", [(32, 128, 512, 128), (32, 128, 512, 512), (32, 64, 1536, 126), (16, 128, 1536, 1536)])
def test_attention_is_deterministic(ctx, H, HD, Nq, Nk):
    """Repeated launches on the same inputs are bit-identical (a softmax warp reading O before the last PV retired, or a
    tensor-memory hazard between P and the next S, shows up as run-to-run noise long before it breaks the tolerance)."""
    D = H * HD
    g = torch.Generator(device="cuda").manual_seed(Nq + Nk)
    q = torch.randn(Nq, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(Nk, D, device="cuda", generator=g).bfloat16()
    ldv = (Nk + 7) // 8 * 8
    vt = torch.zeros(D, ldv, device="cuda", dtype=torch.bfloat16)
    vt[:, :Nk] = torch.randn(D, Nk, device="cuda", generator=g).bfloat16()
    outs = []
    torch.cuda.synchronize()
    for _ in range(6):
        o = torch.full((Nq, D), float("nan"), device="cuda", dtype=torch.bfloat16)
        torch.cuda.synchronize()
        ctx._check(ctx.lib.ltx_op_attention_hd(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), ldv, None, o.data_ptr(),
                                               1, H, HD, Nq, Nk, 1 / math.sqrt(HD)))
        ctx.sync()
        outs.append(o)
    for o in outs[1:]:
        assert torch.equal(o.view(torch.int16), outs[0].view(torch.int16))


def test_attention_block_scattered_output(ctx):
    """The Ulysses epilogue (output row blocks stored into per-destination buffers) equals the plain output, bit for bit."""
    import ctypes
    H, Nq, Nk, D = 4, 192, 320, 512
    g = torch.Generator(device="cuda").manual_seed(5)
    q = torch.randn(Nq, D, device="cuda", generator=g).bfloat16()
    k = torch.randn(Nk, D, device="cuda", generator=g).bfloat16()
    vt = torch.randn(D, Nk, device="cuda", generator=g).bfloat16()
    plain = torch.empty(Nq, D, device="cuda", dtype=torch.bfloat16)
    rows = 64
    blocks = [torch.full((rows, D), float("nan"), device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    torch.cuda.synchronize()
    ctx._check(ctx.lib.ltx_op_attention(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), Nk, None, plain.data_ptr(), 1, H, Nq, Nk,
                                        1 / math.sqrt(128)))
    arr = (ctypes.c_void_p * 3)(*[b.data_ptr() for b in blocks])
    ctx._check(ctx.lib.ltx_op_attention_blocks(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), Nk, H, Nq, Nk, 1 / math.sqrt(128),
                                               arr, 3, rows))
    ctx.sync()
    got = torch.cat(blocks, 0)
    assert torch.equal(got.view(torch.int16), plain.view(torch.int16))
