"""Short program for an `ncu --set full` capture of the attention kernel alone (self-attention, N = 6144: 48 key tiles per CTA,
so the steady-state loop dominates the sampled stalls)."""
import os
import math, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
H, Nq = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 6144
D = H * 128
q = torch.randn(Nq, D, device="cuda").bfloat16(); k = torch.randn(Nq, D, device="cuda").bfloat16()
vt = torch.randn(D, Nq, device="cuda").bfloat16(); o = torch.empty(Nq, D, device="cuda", dtype=torch.bfloat16)
torch.cuda.synchronize()
for _ in range(3):
    ctx._check(ctx.lib.ltx_op_attention(ctx.handle, q.data_ptr(), k.data_ptr(), vt.data_ptr(), Nq, None, o.data_ptr(), 1, H, Nq, Nq, 1 / math.sqrt(128)))
ctx.sync()
print("ok")
