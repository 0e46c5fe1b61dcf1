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


SKINNY = [(26, 2048, 2048), (26, 8192, 2048), (26, 2048, 8192), (11, 128, 128), (32, 48, 96), (17, 16, 32), (16, 2048, 4096),
          (1, 128, 4096), (5, 4096, 64), (24, 256, 192), (31, 272, 1056)]


@pytest.mark.parametrize("M,N,K", SKINNY)
@pytest.mark.parametrize("mode", [0, 1, 3, 4])
def test_gemm_skinny(ctx, M, N, K, mode):
    """The weight-streaming kernel for M <= 32 (gemm_skinny.cu, the dual model's audio-stream Linears), forced with bn = -1:
    same contract and tolerances as the tile kernels, and the two agree with each other to accumulation-order noise."""
    g = torch.Generator(device="cuda").manual_seed(M * 11 + N * 5 + K + mode)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    dt = torch.float32 if mode == 3 else torch.bfloat16
    out = torch.full((M + 1, N), float("nan"), device="cuda", dtype=dt)      # one guard row: nothing may be written past M
    tile = torch.full((M, N), float("nan"), device="cuda", dtype=dt)
    torch.cuda.synchronize()
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, mode, -1))
    if N % 8 == 0 or mode == 3:
        ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), tile.data_ptr(), M, N, K, mode, 0))
    ctx.sync()
    ref = A.float() @ B.float().t() + bias
    if mode == 1:
        ref = _gelu_tanh(ref)
    if mode == 4:
        ref = ref * torch.sigmoid(ref)
    tol = 2e-5 if mode == 3 else 4e-3
    assert torch.isfinite(out[:M].float()).all() and torch.isnan(out[M].float()).all()
    assert rel_l2(out[:M].float(), ref) <= tol
    if N % 8 == 0 or mode == 3:
        assert rel_l2(out[:M].float(), tile.float()) <= (2e-5 if mode == 3 else 3e-3)
    # deterministic
    out2 = torch.empty_like(out)
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out2.data_ptr(), M, N, K, mode, -1))
    ctx.sync()
    assert torch.equal(out[:M], out2[:M])


def test_gemm_skinny_refuses_ineligible_shapes(ctx):
    from ltx_video_swift_mlx_b200._lib import LtxError
    A = torch.zeros(40, 64, device="cuda", dtype=torch.bfloat16)
    B = torch.zeros(32, 64, device="cuda", dtype=torch.bfloat16)
    out = torch.zeros(40, 32, device="cuda", dtype=torch.bfloat16)
    for M, N, K in [(40, 32, 64), (8, 24, 64), (8, 32, 48)]:      # too many rows, N % 16, K % 32
        with pytest.raises(LtxError) as e:
            ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), None, out.data_ptr(), M, N, K, 0, -1))
        assert e.value.code == 2


SWAPAB = [(192, 4096, 4096), (192, 16384, 4096), (192, 4096, 16384), (384, 4096, 4096), (384, 8192, 4096), (512, 1024, 512),
          (1, 128, 64), (16, 256, 64), (48, 520, 192), (130, 264, 320), (200, 1000, 1024), (300, 264, 128), (257, 4096, 2048),
          (96, 128, 4096), (64, 48, 256), (192, 12288, 4096)]


@pytest.mark.parametrize("M,N,K", SWAPAB)
@pytest.mark.parametrize("mode", [0, 1, 3, 4])
@pytest.mark.parametrize("bn", [-2, -3])
def test_gemm_swapab(ctx, M, N, K, mode, bn):
    """The weight-streaming kernel for M <= 512 (gemm_swapab.cu: weights on the TMEM lanes, activation rows as the MMA's N),
    forced with bn = -2 (split-K workspace attached, as the DiT forward does) / -3 (no split): same contract and tolerances as
    the tile kernels, nothing written past row M or into neighbouring columns, and bit-reproducible run to run although the
    split-K partials arrive in any order."""
    if mode != 3 and N % 8 != 0:
        pytest.skip("bf16 output needs N % 8 == 0")
    g = torch.Generator(device="cuda").manual_seed(M * 13 + N * 7 + K + mode)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    dt = torch.float32 if mode == 3 else torch.bfloat16
    out = torch.full((M + 1, N), float("nan"), device="cuda", dtype=dt)      # one guard row
    torch.cuda.synchronize()
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, mode, bn))
    ctx.sync()
    ref = A.float() @ B.float().t() + bias
    if mode == 1:
        ref = _gelu_tanh(ref)
    if mode == 4:
        ref = ref * torch.sigmoid(ref)
    assert torch.isfinite(out[:M].float()).all() and torch.isnan(out[M].float()).all()
    assert rel_l2(out[:M].float(), ref) <= (2e-5 if mode == 3 else 4e-3)
    out2 = torch.full_like(out, float("nan"))
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out2.data_ptr(), M, N, K, mode, bn))
    ctx.sync()
    assert torch.equal(out[:M], out2[:M])


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 264, 128), (1536, 4096, 4096), (1536, 4096, 16384),
                                   (192, 4096, 4096), (384, 4096, 16384), (64, 4096, 4096), (33, 520, 192),   # 32 < M <= 512: swap-AB kernel
                                   (26, 2048, 2048), (11, 128, 128), (32, 2048, 8192)])      # M <= 32: weight-streaming kernel
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
