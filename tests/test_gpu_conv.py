"""Implicit-GEMM 3x3x3 convolution (ltx_op_conv3d) vs torch conv3d on the reference's padding."""
import math

import pytest
import torch

from helpers import O, product, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = product().LtxContext(product().LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
    yield c
    c.close()


@pytest.fixture(params=["single", "pair", "pair+slab"])
def conv_form(request, monkeypatch):
    """LTX_CONV_PAIR / LTX_CONV_SLAB (read by the launcher on every call): one CTA per voxel tile; CTA pairs (cta_group::2: two
    voxel tiles against one weight tile, each CTA staging half of it) wherever Cin comes in 128-channel stages; pairs whose
    stages hold an h-haloed activation slab shared by three taps (tiles up to 128 columns; the default)."""
    monkeypatch.setenv("LTX_CONV_PAIR", "0" if request.param == "single" else "1")
    monkeypatch.setenv("LTX_CONV_SLAB", "1" if request.param == "pair+slab" else "0")
    return request.param


@pytest.mark.parametrize("T,H,W,Cin,Cout,causal", [
    (2, 4, 6, 64, 64, 0), (3, 8, 8, 128, 128, 0), (4, 16, 24, 128, 1024, 0), (2, 5, 7, 64, 128, 1),
    (7, 32, 48, 256, 256, 0), (1, 16, 24, 1024, 256, 0), (3, 6, 10, 128, 48, 0),
    # odd numbers of voxel tiles (the pair's second tile lies past the volume), one tile only, 512 channels
    (5, 8, 8, 128, 128, 1), (1, 8, 16, 256, 128, 0), (3, 16, 24, 512, 512, 0),
])
def test_conv3d(ctx, conv_form, T, H, W, Cin, Cout, causal):
    g = torch.Generator().manual_seed(T * H * W + Cin + Cout)
    x = torch.randn(1, Cin, T, H, W, generator=g)
    w = O.bf16_round(torch.randn(Cout, Cin, 3, 3, 3, generator=g) / math.sqrt(27 * Cin))
    b = torch.randn(Cout, generator=g) * 0.1
    ref = O.conv3d_full(O.bf16_round(x), w, b, causal=bool(causal))[0].permute(1, 2, 3, 0)   # [T,H,W,Cout]
    x_cl = x[0].permute(1, 2, 3, 0).contiguous().cuda()                                       # [T,H,W,Cin]
    w_dev = w.permute(2, 3, 4, 0, 1).reshape(27, Cout, Cin).contiguous().cuda().bfloat16()
    b_dev = b.cuda()
    out = torch.full((T, H, W, Cout), float("nan"), device="cuda")
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    ctx._check(ctx.lib.ltx_op_conv3d(ctx.handle, x_cl.data_ptr(), w_dev.data_ptr(), b_dev.data_ptr(), out.data_ptr(), T, H, W,
                                     Cin, Cout, causal))
    ctx.sync()
    assert torch.isfinite(out).all()
    assert rel_l2(out, ref) <= 1e-4     # same bf16-rounded operands, fp32 accumulation on both sides
