"""ncu target for the dual audio/video forward: a 3-block LTX2Transformer at the real widths (D = 4096, Da = 2048), N = 1536 video
+ 26 audio tokens, S = 1024 -- the per-launch durations of one block are what the 48-block forward repeats.
Usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/ncu_av.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ltx_video_swift_mlx_b200  # noqa: E402,F401
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig  # noqa: E402

L = int(os.environ.get("AV_LAYERS", "3"))
F, H, W, S, Ta, C = 4, 16, 24, 1024, 26, 128
N = F * H * W
ctx = LtxContext(LTXTransformerConfig(num_layers=L), 0)
ctx.init_random_weights(17, seed=101)
ctx.finalize_weights()
g = torch.Generator().manual_seed(1)
vl = torch.randn(1, N, C, generator=g).bfloat16().cuda()
al = torch.randn(1, Ta, C, generator=g).bfloat16().cuda()
tx = torch.randn(1, S, 3840, generator=g)
tx = (tx / tx.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16().cuda()
sg = torch.tensor([0.7, 0.7], device="cuda")
ov = torch.empty(1, N, C, device="cuda")
oa = torch.empty(1, Ta, C, device="cuda")
torch.cuda.synchronize()
for _ in range(int(os.environ.get("AV_REPS", "2"))):
    ctx._check(ctx.lib.ltx_av_forward_dev(ctx.handle, vl.data_ptr(), 1, al.data_ptr(), 1, tx.data_ptr(), tx.data_ptr(), 1,
                                          sg.data_ptr(), sg.data_ptr() + 4, None, None, N, Ta, S, F, H, W, 77,
                                          ov.data_ptr(), oa.data_ptr()))
ctx.sync()
if os.environ.get("AV_TIME"):   # plain timing of the forward (not under ncu): AV_TIME=1 AV_LAYERS=48 python tools/ncu_av.py
    st = torch.cuda.ExternalStream(ctx.stream)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(5):
        ctx._check(ctx.lib.ltx_av_forward_dev(ctx.handle, vl.data_ptr(), 1, al.data_ptr(), 1, tx.data_ptr(), tx.data_ptr(), 1,
                                              sg.data_ptr(), sg.data_ptr() + 4, None, None, N, Ta, S, F, H, W, 77,
                                              ov.data_ptr(), oa.data_ptr()))
    b.record(st)
    torch.cuda.synchronize()
    print("ms_per_forward", a.elapsed_time(b) / 5, "layers", L)
print("ok", float(ov.float().std()), float(oa.float().std()))
ctx.close()
