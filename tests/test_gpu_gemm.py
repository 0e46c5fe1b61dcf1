"""tcgen05 GEMM parity (through the C ABI, ltx_op_gemm / ltx_op_gemm_resid) against a plain fp32 torch matmul."""
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


def _gelu_tanh(x):
    return 0.5 * x * (1 + torch.tanh(math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)))


SHAPES = [
    (128, 128, 64, 0), (128, 256, 64, 256), (128, 256, 128, 128), (256, 512, 256, 0),
    (48, 520, 192, 0), (300, 264, 128, 0), (1536, 128, 4096, 0), (1000, 1000, 512, 256),
    (1536, 4096, 4096, 0), (1536, 4096, 4096, 128), (4096, 48, 256, 0), (130, 8192, 320, 0),
    (1536, 4096, 4096, 176), (1536, 4096, 4096, 240), (300, 1000, 256, 48), (1536, 8192, 4096, 0), (4096, 1536, 4096, 0), (1536, 16384, 512, 0), (2000, 2304, 1024, 0),
    (1536, 128, 4096, 128), (19000, 256, 128, 0), (1536, 4096, 4096, 1192), (300, 1000, 256, 1064), (256, 256, 64, 1256),
    (1536, 16384, 4096, 1000), (130, 264, 128, 1128), (1536, 4096, 4096, 1176), (4096, 1536, 4096, 1176),
    # 4-CTA cluster kernel (A multicast between two pairs): fitted / forced widths, ragged M, N, K, odd column-tile counts
    (1536, 4096, 4096, 2000), (1536, 4096, 4096, 2192), (1536, 8192, 4096, 2256), (4096, 1536, 4096, 2000), (256, 512, 64, 2256),
    (300, 1000, 256, 2064), (130, 264, 128, 2128), (1000, 1000, 512, 2176), (2000, 2304, 1024, 2000), (1536, 16384, 512, 2000),
    (520, 776, 192, 2256), (19000, 512, 128, 2000),
]


@pytest.mark.parametrize("M,N,K,bn", SHAPES)
@pytest.mark.parametrize("mode", [0, 1, 3])
def test_gemm(ctx, M, N, K, bn, mode):
    if mode != 3 and N % 8 != 0:
        pytest.skip("bf16 output needs N % 8 == 0")
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K + mode)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if mode == 3 else torch.bfloat16)
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    rc = ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, mode, bn)
    ctx._check(rc)
    ctx.sync()
    ref = A.float() @ B.float().t() + bias
    if mode == 1:
        ref = _gelu_tanh(ref)
    tol = 2e-5 if mode == 3 else 4e-3   # fp32 out: accumulation-order noise; bf16 out: one rounding (2^-9 rms)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) <= tol


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 264, 128), (1536, 4096, 4096), (1536, 4096, 16384)])
def test_gemm_gate_residual(ctx, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ga = torch.randn(N, device="cuda", generator=g)
    gb = torch.randn(N, device="cuda", generator=g)
    x0 = torch.randn(M, N, device="cuda", generator=g)
    x = x0.clone()
    shadow = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    ctx._check(ctx.lib.ltx_op_gemm_resid(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), x.data_ptr(), ga.data_ptr(),
                                         gb.data_ptr(), shadow.data_ptr(), M, N, K, 0.5))
    ctx.sync()
    ref = x0 + (A.float() @ B.float().t() + bias) * (ga + gb) * 0.5
    assert rel_l2(x, ref) <= 2e-5
    assert rel_l2(shadow.float(), ref) <= 4e-3
