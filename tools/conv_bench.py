"""Micro-benchmark of the implicit-GEMM conv (prologue + conv) on the VAE decoder's stage shapes (768x512x25 frames)."""
import os
import math, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
stream = torch.cuda.ExternalStream(ctx.stream)
SHAPES = [(4, 16, 24, 1024, 1024, 10), (4, 16, 24, 1024, 4096, 1), (7, 32, 48, 512, 512, 10), (7, 32, 48, 512, 2048, 1),
          (13, 64, 96, 256, 256, 10), (13, 64, 96, 256, 1024, 1), (25, 128, 192, 128, 128, 10), (25, 128, 192, 128, 48, 1)]
tot = 0.0
for T, H, W, Ci, Co, cnt in SHAPES:
    x = torch.randn(T, H, W, Ci, device="cuda")
    w = (torch.randn(27, Co, Ci, device="cuda") / math.sqrt(27 * Ci)).bfloat16()
    b = torch.zeros(Co, device="cuda"); o = torch.empty(T, H, W, Co, device="cuda")
    torch.cuda.synchronize()
    ctx.set_profiling(False)
    run = lambda: ctx._check(ctx.lib.ltx_op_conv3d(ctx.handle, x.data_ptr(), w.data_ptr(), b.data_ptr(), o.data_ptr(), T, H, W, Ci, Co, 0))
    for _ in range(2): run()
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(5): run()
    e1.record(stream); ctx.sync()
    t = e0.elapsed_time(e1) / 5
    fl = 2.0 * 27 * Ci * Co * T * H * W
    tot += t * cnt
    print(f"T{T} {H}x{W} {Ci}->{Co}: {t*1e3:8.1f} us (prep+conv) {fl/t/1e9:7.1f} TFLOP/s  x{cnt} = {t*cnt:6.2f} ms", flush=True)
print(f"sum over the decoder's 45 convs (x2 per res block counted in cnt*2?): {tot:.2f} ms")
