"""Times the M <= 32 Linears of the dual model's audio stream on the weight-streaming kernel (bn = -1) and on the tcgen05 tile
kernel (bn = 0); weights are rotated through a > L2 pool so every launch streams from HBM.  GPU box: python tools/skinny_bench.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ltx_video_swift_mlx_b200  # noqa: E402,F401
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig  # noqa: E402

ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
st = torch.cuda.ExternalStream(ctx.stream)
for M, N, K in [(26, 2048, 2048), (26, 8192, 2048), (26, 2048, 8192), (26, 2048, 4096)]:
    nw = max(2, int(400e6 // (N * K * 2)))           # > 126 MB of distinct weights in rotation
    Ws = [(torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16() for _ in range(nw)]
    A = torch.randn(M, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    torch.cuda.synchronize()
    res = {}
    for name, bn in (("skinny", -1), ("tile", 0)):
        for w in Ws[:2]:
            ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), w.data_ptr(), None, out.data_ptr(), M, N, K, 0, bn))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3 * nw
        a.record(st)
        for i in range(reps):
            ctx._check(ctx.lib.ltx_op_gemm(ctx.handle, A.data_ptr(), Ws[i % nw].data_ptr(), None, out.data_ptr(), M, N, K, 0, bn))
        b.record(st)
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / reps * 1e3
    gb = N * K * 2 / 1e9
    print(f"M={M} N={N} K={K}: skinny {res['skinny']:.1f} us ({gb / res['skinny'] * 1e6 / 1e3:.2f} TB/s)  tile {res['tile']:.1f} us "
          f"({gb / res['tile'] * 1e6 / 1e3:.2f} TB/s)")
ctx.close()
