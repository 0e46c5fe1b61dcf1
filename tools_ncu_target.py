"""Short program for `ncu --set full`: a few launches of the three tensor-core kernels at their DiT / VAE shapes."""
import math, sys, torch
sys.path.insert(0, ".")
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
M, N, K = 1536, 16384, 4096
A = torch.randn(M, K, device="cuda").bfloat16(); B = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda"); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
H, Nq = 32, 1536; D = H * 128
q = torch.randn(Nq, D, device="cuda").bfloat16(); k = torch.randn(Nq, D, device="cuda").bfloat16()
vt = torch.randn(D, Nq, device="cuda").bfloat16(); o = torch.empty(Nq, D, device="cuda", dtype=torch.bfloat16)
T, Hh, W, Cin, Cout = 7, 32, 48, 512, 512
x = torch.randn(T, Hh, W, Cin, device="cuda"); w = (torch.randn(27, Cout, Cin, device="cuda") / math.sqrt(27 * Cin)).bfloat16()
cb = torch.zeros(Cout, device="cuda"); co = torch.empty(T, Hh, W, Cout, device="cuda")
torch.cuda.synchronize()
for _ in range(3):
    ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, 1, 0))
    ctx._check(ctx.lib.ltx_op_attention(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), Nq, None, o.data_ptr(), 1, H, Nq, Nq, 1 / math.sqrt(128)))
    ctx._check(ctx.lib.ltx_op_conv3d(ctx.handle, x.data_ptr(), w.data_ptr(), cb.data_ptr(), co.data_ptr(), T, Hh, W, Cin, Cout, 0))
ctx.sync()
print("ok")
