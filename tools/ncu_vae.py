"""ncu target for the video-VAE decode: random-init decoder at the LTX-2 widths, one warm-up + one 25-frame decode (768x512).
Usage (GPU box): ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/ncu_vae.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ltx_video_swift_mlx_b200  # noqa: E402,F401
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig  # noqa: E402

F, H, W = int(os.environ.get("VAE_FRAMES", "4")), 16, 24
ctx = LtxContext(LTXTransformerConfig(num_layers=1, num_attention_heads=1), 0)
ctx.init_random_weights(2, seed=7)            # VAE decoder only
ctx.finalize_weights()
z = torch.randn(128, F, H, W, generator=torch.Generator().manual_seed(3)).cuda()
out = torch.empty(8 * (F - 1) + 1, 32 * H, 32 * W, 3, device="cuda")
torch.cuda.synchronize()
for _ in range(2):
    ctx.vae_decode_dev(z.data_ptr(), (F, H, W), out.data_ptr())
ctx.sync()
print("ok", float(out.mean()), float(out.std()))
ctx.close()
