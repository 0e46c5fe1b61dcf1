"""Short program for `ncu --set full`: a few launches of every hot kernel at its DiT / VAE shape (cfg 2: N=1536, D=4096)."""
import os
import math, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
M, N, K = 1536, 16384, 4096
A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
# gate * residual epilogue (attn to_out shape)
Bo = (torch.randn(4096, K, device="cuda") / math.sqrt(K)).bfloat16(); xres = torch.randn(M, 4096, device="cuda")
gate = torch.ones(4096, device="cuda"); shadow = torch.empty(M, 4096, device="cuda", dtype=torch.bfloat16)
# int8 weights of the FFN-in shape
q8 = torch.empty(N, K, device="cuda", dtype=torch.uint8); qs = torch.empty(K // 64, N, device="cuda"); qb = torch.empty(K // 64, N, device="cuda")
H, Nq = 32, 1536; D = H * 128
q = torch.randn(Nq, D, device="cuda").bfloat16(); k = torch.randn(Nq, D, device="cuda").bfloat16()
vt = torch.randn(D, Nq, device="cuda").bfloat16(); o = torch.empty(Nq, D, device="cuda", dtype=torch.bfloat16)
T, Hh, W, Cin, Cout = 7, 32, 48, 512, 512
x = torch.randn(T, Hh, W, Cin, device="cuda"); w = (torch.randn(27, Cout, Cin, device="cuda") / math.sqrt(27 * Cin)).bfloat16()
cb = torch.zeros(Cout, device="cuda"); co = torch.empty(T, Hh, W, Cout, device="cuda")
# row kernels
xr = torch.randn(M, D, device="cuda"); hr = torch.empty(M, D, device="cuda", dtype=torch.bfloat16); tb = torch.randn(4, D, device="cuda") * 0.1
qkr = torch.randn(M, D, device="cuda").bfloat16(); wn = torch.randn(D, device="cuda"); cs = torch.randn(M, D // 2, device="cuda"); sn = torch.randn(M, D // 2, device="cuda")
# guided Euler at the cfg-2 latent size
n = 128 * 4 * 16 * 24
lat = torch.randn(n, device="cuda"); vc = torch.randn(n, device="cuda"); vu = torch.randn(n, device="cuda"); vs = torch.randn(n, device="cuda"); vp = torch.zeros(n, device="cuda")
torch.cuda.synchronize()
ctx._check(ctx.lib.ltx_op_quantize(ctx.handle, B.data_ptr(), N, K, 8, q8.data_ptr(), qs.data_ptr(), qb.data_ptr()))
for _ in range(3):
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 1, 0))
    ctx._check(ctx.lib.ltx_op_gemm_resid(ctx.handle, A.data_ptr(), Bo.data_ptr(), bias.data_ptr(), xres.data_ptr(), gate.data_ptr(), gate.data_ptr(), shadow.data_ptr(), M, 4096, K, 0.5))
    ctx._check(ctx.lib.ltx_op_gemm_q(ctx.handle, A.data_ptr(), q8.data_ptr(), qs.data_ptr(), qb.data_ptr(), 8, bias.data_ptr(), out.data_ptr(), M, N, K, 1, 0))
    ctx._check(ctx.lib.ltx_op_attention(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), Nq, None, o.data_ptr(), 1, H, Nq, Nq, 1 / math.sqrt(128)))
    ctx._check(ctx.lib.ltx_op_conv3d(ctx.handle, x.data_ptr(), w.data_ptr(), cb.data_ptr(), co.data_ptr(), T, Hh, W, Cin, Cout, 0))
    ctx._check(ctx.lib.ltx_op_rmsnorm_mod(ctx.handle, xr.data_ptr(), hr.data_ptr(), M, D, tb[0].data_ptr(), tb[1].data_ptr(), tb[2].data_ptr(), tb[3].data_ptr(), 1e-6, 0))
    ctx._check(ctx.lib.ltx_op_qknorm_rope(ctx.handle, qkr.data_ptr(), M, D, wn.data_ptr(), cs.data_ptr(), sn.data_ptr(), M, 1e-6))
    ctx._check(ctx.lib.ltx_guided_euler_step_dev(ctx.handle, lat.data_ptr(), vc.data_ptr(), vu.data_ptr(), vs.data_ptr(), vp.data_ptr(), 1, n, 4.0, 0.0, 0.5, 0.0, 0.7, 0.5))
# few-row GEMMs of an Ulysses sp = 8 shard (M = 192): split-K weight-streaming kernel, D x D and FFN-out shapes
A192 = torch.randn(192, 16384, device="cuda").bfloat16(); x192 = torch.randn(192, 4096, device="cuda")
Bfo = (torch.randn(4096, 16384, device="cuda") / 128).bfloat16(); o192 = torch.empty(192, 4096, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A192.data_ptr(), Bo.data_ptr(), bias.data_ptr(), o192.data_ptr(), 192, 4096, 4096, 0, -2))
    ctx._check(ctx.lib.ltx_op_gemm_resid(ctx.handle, A192.data_ptr(), Bfo.data_ptr(), bias.data_ptr(), x192.data_ptr(), gate.data_ptr(), gate.data_ptr(), None, 192, 4096, 16384, 0.5))
ctx.sync()
print("ok")
