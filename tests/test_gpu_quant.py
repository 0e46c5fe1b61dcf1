"""Quantised weights (MLX-affine style, group 64, 8 / 4 bit): the dequant-fused GEMM equals the bf16 GEMM on the dequantised
weights (the exact MLX rounding rule is not in the reference tree, so parity is pinned this way), the quantiser
honours its error bound, and the end-to-end drift of a quantised DiT against the bf16 DiT stays small."""
import math

import numpy as np
import pytest
import torch

from helpers import O, make_ctx_with_dit, product, rel_l2, small_dit_config

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = product().LtxContext(product().LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
    yield c
    c.close()


def _quantize(ctx, W, bits):
    N, K = W.shape
    q = torch.empty(N, K * bits // 8, device="cuda", dtype=torch.uint8)
    s = torch.empty(K // 64, N, device="cuda")
    b = torch.empty(K // 64, N, device="cuda")
    torch.cuda.synchronize()
    ctx._check(ctx.lib.ltx_op_quantize(ctx.handle, W.data_ptr(), N, K, bits, q.data_ptr(), s.data_ptr(), b.data_ptr()))
    Wd = torch.empty_like(W)
    ctx._check(ctx.lib.ltx_op_dequantize(ctx.handle, q.data_ptr(), s.data_ptr(), b.data_ptr(), N, K, bits, Wd.data_ptr()))
    ctx.sync()
    return q, s, b, Wd


@pytest.mark.parametrize("bits", [8, 4])
def test_quantiser_error_bound_and_layout(ctx, bits):
    g = torch.Generator(device="cuda").manual_seed(bits)
    N, K = 96, 256
    W = (torch.randn(N, K, device="cuda", generator=g) * 0.05).bfloat16()
    q, s, b, Wd = _quantize(ctx, W, bits)
    levels = 2 ** bits - 1
    codes = q.long() if bits == 8 else torch.stack([q.long() & 15, q.long() >> 4], -1).reshape(N, K)
    assert int(codes.min()) >= 0 and int(codes.max()) <= levels
    srow = s.t().repeat_interleave(64, dim=1)       # [N, K]
    brow = b.t().repeat_interleave(64, dim=1)
    ref = (codes.float() * srow + brow)
    assert torch.allclose(Wd.float(), ref.bfloat16().float(), atol=0, rtol=2 ** -7)      # dequantiser == s*q+beta, bf16-rounded
    err = (W.float() - ref).abs()
    assert bool((err <= 0.5 * srow + 2 ** -8 * W.float().abs() + 1e-6).all())            # half a step + bf16 rounding of s, beta
    Wg = W.float().view(N, K // 64, 64)
    assert torch.allclose(b.t(), Wg.min(-1).values.bfloat16().float())                     # beta = group minimum


@pytest.mark.parametrize("bits", [8, 4])
@pytest.mark.parametrize("M,N,K,mode,bn", [(128, 128, 64, 3, 0), (300, 264, 256, 0, 0), (1536, 4096, 4096, 3, 0),
                                           (1536, 8192, 4096, 1, 0), (1536, 4096, 16384, 0, 0), (200, 520, 128, 3, 48)])
def test_dequant_fused_gemm_equals_gemm_on_dequantised_weights(ctx, bits, M, N, K, mode, bn):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + bits)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    q, s, b, Wd = _quantize(ctx, W, bits)
    dt = torch.float32 if mode == 3 else torch.bfloat16
    out_q = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
    out_d = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
    torch.cuda.synchronize()
    ctx._check(ctx.lib.ltx_op_gemm_q(ctx.handle, A.data_ptr(), q.data_ptr(), s.data_ptr(), b.data_ptr(), bits, bias.data_ptr(),
                                     out_q.data_ptr(), M, N, K, mode, bn))
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), Wd.data_ptr(), bias.data_ptr(), out_d.data_ptr(), M, N, K, mode, bn))
    ctx.sync()
    assert torch.isfinite(out_q.float()).all()
    # same bf16 operand values, same MMA sequence: results agree to fp32 accumulation noise (bit-identical in practice)
    assert rel_l2(out_q.float(), out_d.float()) <= 1e-6
    ref = A.float() @ Wd.float().t() + bias
    if mode == 1:
        ref = 0.5 * ref * (1 + torch.tanh(math.sqrt(2 / math.pi) * (ref + 0.044715 * ref ** 3)))
    assert rel_l2(out_q.float(), ref) <= (2e-5 if mode == 3 else 4e-3)


@pytest.mark.parametrize("bits,tol", [(8, 3e-2), (4, 0.35)])
def test_quantised_dit_drift(bits, tol):
    ocfg, pcfg = small_dit_config(2, 2)
    ctx16, w = make_ctx_with_dit(ocfg, pcfg, seed=12)
    ctxq = product().LtxContext(pcfg, 0)
    ctxq.load_weights(w)
    ctxq.finalize_weights(quant_bits=bits)
    gq = torch.Generator().manual_seed(3)
    fhw, S = (2, 4, 6), 40
    lat = torch.randn(1, 48, 128, generator=gq).bfloat16()
    cx = torch.randn(1, S, ocfg.caption_channels, generator=gq)
    cx = (cx / cx.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    sig = np.array([0.6], dtype=np.float32)
    a = ctx16.dit_forward(lat, cx, sig, None, fhw)
    b = ctxq.dit_forward(lat, cx, sig, None, fhw)
    assert np.isfinite(b).all()
    err = rel_l2(b, a)
    assert err <= tol, err          # drift of the quantised model against the bf16 model (not an oracle tolerance)
    assert err > 1e-5               # and the quantised path really ran
    ctx16.close()
    ctxq.close()


@pytest.mark.parametrize("bits", [8, 4])
def test_materialised_quant_storage_equals_codes(bits):
    """ltx_set_quant_storage(ctx, 1): the weights are quantised exactly as in the default mode, but their dequantised bf16 values
    replace them once at load time instead of being rebuilt inside every GEMM.  Same operand values into the same tensor-core
    products: the two contexts must agree bit for bit (token counts on both sides of the fused-kernel / panel switch at M = 256),
    and both must differ from the unquantised model."""
    ocfg, pcfg = small_dit_config(2, 2)
    ctx16, w = make_ctx_with_dit(ocfg, pcfg, seed=14)
    ctxc = product().LtxContext(pcfg, 0)
    ctxc.load_weights(w)
    ctxc.finalize_weights(quant_bits=bits)
    ctxm = product().LtxContext(pcfg, 0)
    ctxm.set_quant_storage(True)
    ctxm.load_weights(w)
    ctxm.finalize_weights(quant_bits=bits)
    gq = torch.Generator().manual_seed(5)
    for fhw in ((2, 4, 6), (3, 10, 12)):            # 48 and 360 tokens
        N, S = fhw[0] * fhw[1] * fhw[2], 40
        lat = torch.randn(1, N, 128, generator=gq).bfloat16()
        cx = torch.randn(1, S, ocfg.caption_channels, generator=gq)
        cx = (cx / cx.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
        sig = np.array([0.6], dtype=np.float32)
        a = ctxc.dit_forward(lat, cx, sig, None, fhw)
        b = ctxm.dit_forward(lat, cx, sig, None, fhw)
        ref = ctx16.dit_forward(lat, cx, sig, None, fhw)
        assert np.isfinite(b).all()
        assert rel_l2(b, a) <= 2e-3, rel_l2(b, a)     # same operand values; the GEMM kernels (and their summation order) may differ
        assert rel_l2(b, ref) > 1e-5                  # the quantisation is really in the weights
    for c in (ctx16, ctxc, ctxm):
        c.close()
