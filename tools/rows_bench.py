"""Micro-benchmark of the HBM-bound row kernels of a DiT block (AdaLN RMSNorm, q/k RMSNorm + RoPE) at D = 4096.
R launches queued behind a 1 GB memset on the library's stream, cycling over NBUF distinct buffers (> L2 in total) so every
launch reads its rows from HBM; CUDA events around the R.  GB/s counts x read + written and the RoPE tables once per launch."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ltx_video_swift_mlx_b200  # noqa
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
stream = torch.cuda.ExternalStream(ctx.stream)
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
D, NBUF, R = 4096, 8, 16
tb = torch.randn(4, D, device="cuda") * 0.1
wn = torch.randn(D, device="cuda")


def timed(run):
    run(0); ctx.sync(); ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            big.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(R):
            run(i % NBUF)
        e1.record(stream); ctx.sync(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / R)
    return sorted(ts)[len(ts) // 2]


for M in (1536, 3072):
    xs = [torch.randn(M, D, device="cuda") for _ in range(NBUF)]
    hs = [torch.empty(M, D, device="cuda", dtype=torch.bfloat16) for _ in range(NBUF)]
    t = timed(lambda i: ctx._check(ctx.lib.ltx_op_rmsnorm_mod(ctx.handle, xs[i].data_ptr(), hs[i].data_ptr(), M, D, tb[0].data_ptr(),
                                                              tb[1].data_ptr(), tb[2].data_ptr(), tb[3].data_ptr(), 1e-6, 0)))
    print(f"rmsnorm_mod   M={M}: {t*1e3:7.1f} us  {6.0*M*D/t/1e6:7.0f} GB/s", flush=True)
    qs = [torch.randn(M, D, device="cuda").bfloat16() for _ in range(NBUF)]
    t = timed(lambda i: ctx._check(ctx.lib.ltx_op_qknorm_rope(ctx.handle, qs[i].data_ptr(), M, D, wn.data_ptr(), None, None, 1, 1e-6)))
    print(f"qknorm        M={M}: {t*1e3:7.1f} us  {4.0*M*D/t/1e6:7.0f} GB/s", flush=True)
    cs = [torch.randn(1536, D // 2, device="cuda") for _ in range(NBUF)]
    sn = [torch.randn(1536, D // 2, device="cuda") for _ in range(NBUF)]
    t = timed(lambda i: ctx._check(ctx.lib.ltx_op_qknorm_rope(ctx.handle, qs[i].data_ptr(), M, D, wn.data_ptr(), cs[i].data_ptr(),
                                                              sn[i].data_ptr(), 1536, 1e-6)))
    print(f"qknorm_rope   M={M}: {t*1e3:7.1f} us  {(4.0*M*D + 4.0*1536*D)/t/1e6:7.0f} GB/s", flush=True)
