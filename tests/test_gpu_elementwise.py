"""Row / elementwise kernels through the C ABI vs the oracle's fp32 formulas."""
import numpy as np
import pytest
import torch

from helpers import O, product, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = product().LtxContext(product().LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
    yield c
    c.close()


@pytest.mark.parametrize("M,D,layernorm", [(5, 256, 0), (48, 4096, 0), (48, 4096, 1), (1536, 4096, 0), (7, 512, 1),
                                                # streaming form: fewer rows than CTAs, ring refills (rows > 3 x 148 x 4), ragged last round
                                                (1, 4096, 0), (445, 4096, 1), (3077, 4096, 0), (3077, 4096, 1)])
def test_rmsnorm_mod(ctx, M, D, layernorm):
    g = torch.Generator(device="cuda").manual_seed(M + D)
    x = torch.randn(M, D, device="cuda", generator=g) * 3 + 0.5
    ts, tc, as_, ac = [torch.randn(D, device="cuda", generator=g) * 0.3 for _ in range(4)]
    out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    ctx._check(ctx.lib.ltx_op_rmsnorm_mod(ctx.handle, x.data_ptr(), out.data_ptr(), M, D, ts.data_ptr(), tc.data_ptr(),
                                          as_.data_ptr(), ac.data_ptr(), 1e-6, layernorm))
    ctx.sync()
    if layernorm:
        mu = x.mean(-1, keepdim=True)
        n = (x - mu) * torch.rsqrt(((x - mu) ** 2).mean(-1, keepdim=True) + 1e-6)
    else:
        n = O.rms_norm(x, None, 1e-6)
    ref = n * (1 + tc + ac) + ts + as_
    assert rel_l2(out.float(), ref) <= 4e-3


@pytest.mark.parametrize("heads,fhw,rope", [(2, (2, 4, 6), True), (32, (4, 16, 24), True), (4, (1, 3, 5), False),
                                               # streaming form at D = 4096: ring refills with and without the rotation tables
                                               (32, (1, 1, 3), True), (32, (2, 40, 41), True), (32, (1, 50, 53), False)])
def test_qknorm_rope(ctx, heads, fhw, rope):
    cfg = O.DiTConfig(num_heads=heads)
    D = cfg.inner_dim
    N = fhw[0] * fhw[1] * fhw[2]
    g = torch.Generator(device="cuda").manual_seed(N + D)
    x = torch.randn(N, D, device="cuda", generator=g).bfloat16()
    w = 1 + 0.1 * torch.randn(D, device="cuda", generator=g)
    cos, sin = O.rope_table(cfg, *fhw)                      # [H, N, 64]
    cos_t = cos.permute(1, 0, 2).reshape(N, D // 2).contiguous().cuda()
    sin_t = sin.permute(1, 0, 2).reshape(N, D // 2).contiguous().cuda()
    y = x.clone()
    torch.cuda.synchronize()   # inputs were produced on torch\'s stream; the library runs on its own
    ctx._check(ctx.lib.ltx_op_qknorm_rope(ctx.handle, y.data_ptr(), N, D, w.data_ptr(), cos_t.data_ptr() if rope else None,
                                          sin_t.data_ptr() if rope else None, N, 1e-6))
    ctx.sync()
    ref = O.rms_norm(x.float().cpu(), w.cpu(), 1e-6).unsqueeze(0)
    if rope:
        ref = O.apply_split_rope(ref, cos, sin, heads)
    assert rel_l2(y.float(), ref[0]) <= 4e-3


@pytest.mark.parametrize("use_cfg,phi,stg,gamma,last", [(False, 0, 0, 0, False), (True, 0.7, 0.5, 0.0, False),
                                                         (True, 0.0, 0.0, 0.3, True), (False, 0, 0.5, 0.2, False)])
def test_guided_euler(ctx, use_cfg, phi, stg, gamma, last):
    g = torch.Generator().manual_seed(11)
    shape = (1, 128, 2, 4, 6)
    lat = torch.randn(shape, generator=g)
    vc, vu, vs, vp = [torch.randn(shape, generator=g) for _ in range(4)]
    sigma, sn = 0.7, (0.0 if last else 0.4)
    ref_lat, ref_v = O.guided_euler_step(lat, vc, vu if use_cfg else None, vs if stg > 0 else None, vp if gamma > 0 else None,
                                         4.0, phi, stg, gamma, sigma, sn)
    x = lat.numpy().copy()
    vprev = vp.numpy().copy()
    ctx.guided_euler_step(x, vc.numpy(), vu.numpy() if use_cfg else None, vs.numpy() if stg > 0 else None, vprev,
                          use_prev=gamma > 0, cfg_scale=4.0, rescale_phi=phi, stg_scale=stg, ge_gamma=gamma, sigma=sigma,
                          sigma_next=sn)
    assert rel_l2(x, ref_lat) <= 1e-5
    assert rel_l2(vprev, ref_v) <= 1e-5


def test_guided_euler_closed_forms(ctx):
    # sigma' = 0 returns x - sigma v ; cfg scale 1 returns the conditional velocity (SURVEY section 4 KATs)
    x0 = np.random.RandomState(0).randn(4096).astype(np.float32)
    v = np.random.RandomState(1).randn(4096).astype(np.float32)
    u = np.random.RandomState(2).randn(4096).astype(np.float32)
    x = x0.copy()
    ctx.guided_euler_step(x, v, sigma=0.5, sigma_next=0.0)
    np.testing.assert_allclose(x, x0 - 0.5 * v, rtol=0, atol=1e-6)
    x = x0.copy()
    vprev = np.zeros_like(x)
    ctx.guided_euler_step(x, v, v_uncond=u, v_prev=vprev, cfg_scale=1.0, sigma=0.5, sigma_next=0.25)
    np.testing.assert_allclose(vprev, v, rtol=0, atol=1e-6)
