"""Micro-benchmark of the HBM-bound row kernels at the DiT shapes (device time via CUDA events on the library stream)."""
import os
import sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
stream = torch.cuda.ExternalStream(ctx.stream)
M, D = 1536, 4096
x = torch.randn(M, D, device="cuda"); out = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
tb = torch.randn(4, D, device="cuda") * 0.1
qk = torch.randn(M, D, device="cuda").bfloat16(); w = torch.randn(D, device="cuda")
cs = torch.randn(M, D // 2, device="cuda"); sn = torch.randn(M, D // 2, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
def t_norm():
    ctx._check(ctx.lib.ltx_op_rmsnorm_mod(ctx.handle, x.data_ptr(), out.data_ptr(), M, D, tb[0].data_ptr(), tb[1].data_ptr(), tb[2].data_ptr(), tb[3].data_ptr(), 1e-6, 0))
def t_qk():
    ctx._check(ctx.lib.ltx_op_qknorm_rope(ctx.handle, qk.data_ptr(), M, D, w.data_ptr(), cs.data_ptr(), sn.data_ptr(), M, 1e-6))
for name, fn, nbytes in [("rmsnorm_mod", t_norm, M * D * 6), ("qknorm_rope(1 seg)", t_qk, M * D * 4 + M * D * 4)]:
    for cold in (True, False):
        for _ in range(3): fn()
        ctx.sync(); ts = []
        for _ in range(20):
            if cold: flush.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); ctx.sync(); ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[len(ts) // 2]
        # 10 back-to-back launches: amortises the event/launch overhead
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10): fn()
        e1.record(stream); ctx.sync(); t10 = e0.elapsed_time(e1) / 10
        print(f"{name:20s} {'cold(L2 flushed)' if cold else 'warm':17s}: single {t*1e3:7.1f} us ({nbytes/t/1e6:7.0f} GB/s)  x10 avg {t10*1e3:7.1f} us ({nbytes/t10/1e6:7.0f} GB/s)", flush=True)
