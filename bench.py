#!/usr/bin/env python
"""bench.py -- DiT steps/s (+ VAE frames/s) of the LTX-2 denoise hot path on B200, the metric BASELINE.json names.

One "step" = one denoise step of the distilled LTX-2 video DiT at 768x512x25 frames (BASELINE config 2: N = 1536 video
tokens, S = 1024 text tokens, 48 blocks, D = 4096, bf16 weights, random-init on the device, synthetic inputs):
patchify -> 48-block forward -> unpatchify -> Euler update.  Step-invariant text projections are cached by the warm-up.

  value        steps/s, whole job, latent resident in HBM (ltx_denoise_step), CUDA events on the library's stream
  e2e          steps/s through the host-buffer C ABI the Swift pipeline binds (ltx_dit_forward + ltx_guided_euler_step):
               pinned-host -> device copies of latent/timestep (+ text on a cache miss) and device -> host copies of the
               velocity and the new latent are inside the timed region
  roofline     tensor-pipe roofline of the dominant kernel class (the tcgen05 GEMM): algorithmic FLOPs of the GEMM launches
               of one step / their summed device time (CUDA events around every launch, separate profiled pass)
  cpu_baseline the CPU oracle (a port of the reference's algorithm) on this box's host cores, bounded sample
  vae          secondary metric: frames/s of the video-VAE decode of the same clip (25 frames, 768x512)

`--impl reference` times the CPU port only (the Swift/MLX reference cannot be built in this image).
Multi-GPU (torchrun, one rank per GPU): independent replicas (whole-video data parallel), weak scaling.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG2 = dict(width=768, height=512, frames=25, F=4, H=16, W=24, N=1536, S=1024)
D, L, HEADS, CAP, CIN = 4096, 48, 32, 3840, 128


def dit_flops_per_step(N, S, cached_text=True):
    """per block 8ND^2 + 4N^2D + 4ND^2 (+ 4SD^2 text K/V when not cached) + 4NSD + 16ND^2."""
    blk = 8 * N * D * D + 4 * N * N * D + 4 * N * D * D + 4 * N * S * D + 16 * N * D * D
    if not cached_text:
        blk += 4 * S * D * D
    return L * blk + 2 * N * CIN * D * 2


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(tflops_sustained=j["bf16_tflops_sustained"], tflops_burst=j["bf16_tflops"], hbm_gbs=j["hbm_gbs"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(tflops_sustained=1400.0, tflops_burst=1590.0, hbm_gbs=6650.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower().startswith("active") for r in self.rows)]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None, reasons=reasons,
                    samples=len(sm))


# ------------------------------------------------------------------------------------------------ CPU port arm
def cpu_port_step_seconds(sample_blocks=1, repeats=1):
    """Times the oracle (CPU port of the reference algorithm) on a bounded sample of the config-2 step: `sample_blocks`
    transformer blocks at the full shapes (N=1536, S=1024, D=4096, fp32) plus the embedding / head work measured once,
    extrapolated to 48 blocks.  Returns (seconds per full step, cores, description)."""
    from oracle import ltx_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = O.DiTConfig()
    if not _CPU_CACHE:
        _CPU_CACHE.update(_cpu_port_inputs(O, cfg))
    w, x, ctx, ada, rope = (_CPU_CACHE[k] for k in ("w", "x", "ctx", "ada", "rope"))
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            y = x
            for _b in range(sample_blocks):
                y = O.block_forward(w, 0, y, ada, ctx, None, rope, cfg)
            best = min(best, (time.perf_counter() - t0) / sample_blocks)
    # text K/V are step-invariant in our arm (cached); the oracle recomputes them inside block_forward, which is what the
    # reference does every step as well (T/LTXAttention.swift:174-180), so the per-block time is used unchanged.
    return best * L, os.cpu_count() or 1, f"{sample_blocks} of {L} blocks at N=1536,S=1024,D=4096 fp32 (torch CPU), x{L}"


_CPU_CACHE = {}


def _cpu_port_inputs(O, cfg):
    g = torch.Generator().manual_seed(0)
    w = {}
    p = "transformer_blocks.0."
    w[p + "scale_shift_table"] = torch.randn(6, D, generator=g) * 0.1
    for a in ("attn1", "attn2"):
        for l in ("to_q", "to_k", "to_v", "to_out"):
            w[p + f"{a}.{l}.weight"] = torch.randn(D, D, generator=g) / math.sqrt(D)
            w[p + f"{a}.{l}.bias"] = torch.zeros(D)
        w[p + f"{a}.q_norm.weight"] = torch.ones(D)
        w[p + f"{a}.k_norm.weight"] = torch.ones(D)
    w[p + "ff.project_in.proj.weight"] = torch.randn(4 * D, D, generator=g) / math.sqrt(D)
    w[p + "ff.project_in.proj.bias"] = torch.zeros(4 * D)
    w[p + "ff.project_out.weight"] = torch.randn(D, 4 * D, generator=g) / math.sqrt(4 * D)
    w[p + "ff.project_out.bias"] = torch.zeros(D)
    x = torch.randn(1, CFG2["N"], D, generator=g)
    ctx = torch.randn(1, CFG2["S"], D, generator=g)
    ada = torch.randn(1, 1, 6, D, generator=g) * 0.1
    rope = O.rope_table(cfg, CFG2["F"], CFG2["H"], CFG2["W"])
    return dict(w=w, x=x, ctx=ctx, ada=ada, rope=rope)


def run_reference(args, rank):
    if rank != 0:
        return
    times = []
    for i in range(args.warmup + args.steps):
        t, cores, desc = cpu_port_step_seconds(1, 1)
        if i >= args.warmup:
            times.append(t)
    sec = float(np.mean(times))
    line = dict(metric="dit_steps_per_s", value=1.0 / sec, unit="steps/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=sec * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                impl="reference",
                config=dict(workload="LTX-2 distilled DiT denoise step, 768x512x25f (N=1536 tokens, S=1024 text), 48 blocks",
                            note="CPU port of the reference algorithm (oracle/); the Swift+MLX reference cannot be built here"),
                cpu_baseline=dict(value=1.0 / sec, unit="steps/s", cores=cores, kind="port", sample=desc),
                e2e=dict(value=1.0 / sec, unit="steps/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args, rank, world, local_rank):
    import ltx_video_swift_mlx_b200  # noqa: F401
    from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig, make_flags
    from ltx_video_swift_mlx_b200.scheduler import LTXScheduler

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    ctx = LtxContext(LTXTransformerConfig(), local_rank)
    ctx.init_random_weights(3, seed=1234)   # one model; the replicas differ in noise and text
    ctx.finalize_weights()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    F, H, W, N, S = CFG2["F"], CFG2["H"], CFG2["W"], CFG2["N"], CFG2["S"]
    g = torch.Generator().manual_seed(1236 + rank)
    noise = torch.randn(1, CIN, F, H, W, generator=g)
    text = torch.randn(1, S, CAP, generator=g)
    text = (text / text.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()          # unit-RMS rows, mask all ones
    sigmas = LTXScheduler().set_timesteps(8, distilled=True, latent_token_count=N)
    pairs = [(sigmas[i], sigmas[i + 1]) for i in range(len(sigmas) - 1)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- resident path (value)
    ctx.denoise_begin(noise[0].numpy(), (F, H, W), sigmas[0], text, None)
    step_no = [0]

    def resident_step():
        sg, sn = pairs[step_no[0] % len(pairs)]
        ctx.denoise_step(sg, sn, step_no[0] % len(pairs))
        step_no[0] += 1

    for _ in range(max(args.warmup, 3)):
        resident_step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        resident_step()
    e1.record(stream)
    barrier()
    launches = (ctx.launch_count - l0) // args.steps
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step

    # ---------------- profiled pass: per-kernel-class device time of one step (not part of the timed value)
    ctx.set_profiling(True)
    resident_step()
    prof = ctx.get_profile()
    ctx.set_profiling(False)

    # ---------------- e2e through the host-buffer ABI (the Swift seam), pinned host memory
    lat_host = torch.empty(1, N, CIN, dtype=torch.bfloat16).pin_memory()
    lat32 = (noise * sigmas[0]).contiguous().pin_memory()
    vel_lat = torch.empty_like(lat32).pin_memory()
    text_pin = text.pin_memory()
    ts = torch.zeros(1).pin_memory()
    from ltx_video_swift_mlx_b200 import latent_utils
    shape = latent_utils.VideoLatentShape(1, CIN, F, H, W)
    flags = make_flags(context_key=4242)

    def host_step(i):
        sg, sn = pairs[i % len(pairs)]
        lat_host.copy_(torch.from_numpy(latent_utils.patchify(lat32.numpy())))     # host patchify + bf16 cast (:815)
        ts[0] = sg
        v = ctx.dit_forward(lat_host, text_pin, ts.numpy(), None, (F, H, W), flags)
        vel_lat.copy_(torch.from_numpy(latent_utils.unpatchify(v, shape)))
        ctx.guided_euler_step(lat32.numpy(), vel_lat.numpy(), sigma=sg, sigma_next=sn)

    for i in range(max(args.warmup, 3)):
        host_step(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        host_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    if dist is not None:
        t = torch.tensor([e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = N * CIN * 2 + 4 + 2 * N * CIN * 4          # bf16 tokens + sigma ; latent + velocity for the Euler call
    d2h = N * CIN * 4 + N * CIN * 4                  # velocity ; updated latent

    # ---------------- VAE decode (secondary metric), resident + host ABI
    lat_dev = torch.randn(CIN, F, H, W, device="cuda")
    frames_dev = torch.empty(8 * (F - 1) + 1, 32 * H, 32 * W, 3, device="cuda")
    torch.cuda.synchronize()
    for _ in range(3):
        ctx.vae_decode_dev(lat_dev.data_ptr(), (F, H, W), frames_dev.data_ptr())
    barrier()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = max(3, min(args.steps, 10))
    v0.record(stream)
    for _ in range(reps):
        ctx.vae_decode_dev(lat_dev.data_ptr(), (F, H, W), frames_dev.data_ptr())
    v1.record(stream)
    barrier()
    vae_ms = v0.elapsed_time(v1) / reps
    ctx.set_profiling(True)
    ctx.vae_decode_dev(lat_dev.data_ptr(), (F, H, W), frames_dev.data_ptr())
    vprof = ctx.get_profile()
    ctx.set_profiling(False)
    lat_cpu = lat_dev.cpu().numpy()
    frames_host = ctx.pinned_empty((8 * (F - 1) + 1, 32 * H, 32 * W, 3))      # page-locked output buffer (ltx_host_alloc)
    ctx.vae_decode(lat_cpu, out=frames_host)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.vae_decode(lat_cpu, out=frames_host)
    vae_e2e_ms = (time.perf_counter() - t0) * 1e3 / 3
    n_frames = 8 * (F - 1) + 1

    # ---------------- extra single-GPU reference points for the multi-GPU modes (guided step, 121-frame decode)
    extras = {}
    _, ntext = None, torch.randn(1, S, CAP, generator=g)
    ntext = (ntext / ntext.pow(2).mean(-1, keepdim=True).sqrt()).bfloat16()
    dev_sig = LTXScheduler().set_timesteps(40, distilled=False, latent_token_count=N)

    def time_guided(nsteps=3):
        """BASELINE config 3: dev schedule, CFG 4.0 + STG 0.5 at block 29 -> 3 forwards per step."""
        ctx.denoise_begin(noise[0].numpy(), (F, H, W), dev_sig[0], text, None, ntext, None)
        for i in range(2):
            ctx.denoise_step(dev_sig[i], dev_sig[i + 1], i, cfg_scale=4.0, stg_scale=0.5, stg_blocks=(29,))
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(2, 2 + nsteps):
            ctx.denoise_step(dev_sig[i], dev_sig[i + 1], i, cfg_scale=4.0, stg_scale=0.5, stg_blocks=(29,))
        b.record(stream)
        barrier()
        t = a.elapsed_time(b) / nsteps
        if dist is not None:
            tt = torch.tensor([t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t

    def time_vae121(reps=2):
        """BASELINE config 4: 768x512x121 frames (latent 16x16x24)."""
        lat = torch.randn(CIN, 16, H, W, generator=torch.Generator().manual_seed(77)).cuda()
        out = torch.empty(121, 32 * H, 32 * W, 3, device="cuda")
        torch.cuda.synchronize()
        ctx.vae_decode_dev(lat.data_ptr(), (16, H, W), out.data_ptr())
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(reps):
            ctx.vae_decode_dev(lat.data_ptr(), (16, H, W), out.data_ptr())
        b.record(stream)
        barrier()
        t = a.elapsed_time(b) / reps
        if dist is not None:
            tt = torch.tensor([t], device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        del lat, out
        return t

    if world == 1:
        # qint8 weights (the reference's --transformer-quant qint8 path): same step, dequant-fused GEMMs
        try:
            cq = LtxContext(LTXTransformerConfig(), local_rank)
            cq.init_random_weights(1, seed=99)
            cq.finalize_weights(quant_bits=8)
            sq = torch.cuda.ExternalStream(cq.stream, device=torch.device("cuda", local_rank))
            cq.denoise_begin(noise[0].numpy(), (F, H, W), sigmas[0], text, None)
            for i in range(3):
                cq.denoise_step(pairs[i][0], pairs[i][1], i)
            cq.sync()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(sq)
            for i in range(6):
                cq.denoise_step(pairs[i][0], pairs[i][1], i)
            b.record(sq)
            cq.sync()
            tq = a.elapsed_time(b) / 6
            extras["qint8"] = dict(desc="same step with int8 group-64 weights (M > 256: weight converted once per GEMM into an L2-resident bf16 panel + bf16 pair kernel; M <= 256: dequant-fused tcgen05 GEMM)", ms_per_step=tq,
                                   steps_per_s=1e3 / tq)
            cq.close()
        except Exception as e:   # report, do not hide
            extras["qint8"] = dict(error=str(e))
        # ---- the rows either side of the denoise loop: dual audio/video model, VAE encoder, latent upscaler
        def ev_time(fn, strm, reps=3, warm=1):
            for _ in range(warm):
                fn()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(strm)
            for _ in range(reps):
                fn()
            b.record(strm)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        try:
            ca = LtxContext(LTXTransformerConfig(), local_rank)
            ca.init_random_weights(17, seed=101)          # video + audio / cross-modal tensors: the full LTX2Transformer
            ca.finalize_weights()
            sa = torch.cuda.ExternalStream(ca.stream, device=torch.device("cuda", local_rank))
            Ta = 26                                       # 25 frames at 24 fps = 1.04 s of audio at 25 latent frames / s
            vl = torch.randn(1, N, CIN, generator=g).bfloat16().cuda()
            al = torch.randn(1, Ta, CIN, generator=g).bfloat16().cuda()
            tx = text.cuda()
            sg2 = torch.tensor([0.7, 0.7], device="cuda")
            ov = torch.empty(1, N, CIN, device="cuda")
            oa = torch.empty(1, Ta, CIN, device="cuda")
            torch.cuda.synchronize()

            def av_step():
                ca._check(ca.lib.ltx_av_forward_dev(ca.handle, vl.data_ptr(), 1, al.data_ptr(), 1, tx.data_ptr(), tx.data_ptr(), 1,
                                                    sg2.data_ptr(), sg2.data_ptr() + 4, None, None, N, Ta, S, F, H, W, 77,
                                                    ov.data_ptr(), oa.data_ptr()))
            av_step()
            ta_ms = ev_time(av_step, sa, reps=4)
            ca.set_profiling(True)
            av_step()
            pa = ca.get_profile()
            ca.set_profiling(False)
            extras["av_dual_forward"] = dict(
                desc=f"LTX2Transformer (dual audio/video, 48 blocks, D=4096 + Da=2048) forward, N={N} video + {Ta} audio tokens, S={S}",
                ms_per_forward=ta_ms, forwards_per_s=1e3 / ta_ms,
                kernel_classes={k: v for k, v in pa.items() if v["launches"]})
            ca.close()
            del vl, al, tx, ov, oa
        except Exception as e:
            extras["av_dual_forward"] = dict(error=str(e))
        try:
            ce = LtxContext(LTXTransformerConfig(), local_rank)
            ce.init_random_weights(2 | 4 | 8, seed=103)   # VAE decoder (latent statistics), encoder, upscaler
            ce.finalize_weights()
            se_ = torch.cuda.ExternalStream(ce.stream, device=torch.device("cuda", local_rank))
            px = (torch.rand(3, 1, 32 * H, 32 * W, generator=g) * 2 - 1).cuda()
            zl = torch.empty(CIN, 1, H, W, device="cuda")
            px25 = (torch.rand(3, 8 * (F - 1) + 1, 32 * H, 32 * W, generator=g) * 2 - 1).cuda()
            zl25 = torch.empty(CIN, F, H, W, device="cuda")
            lat1 = torch.randn(CIN, 33, H, W, generator=g).cuda()     # stage-1 latent of BASELINE config 5 (768x512x257)
            lat2 = torch.empty(CIN, 33, 2 * H, 2 * W, device="cuda")
            torch.cuda.synchronize()
            t_img = ev_time(lambda: ce.vae_encode_dev(px.data_ptr(), (1, 32 * H, 32 * W), zl.data_ptr()), se_)
            t_clip = ev_time(lambda: ce.vae_encode_dev(px25.data_ptr(), (8 * (F - 1) + 1, 32 * H, 32 * W), zl25.data_ptr()), se_)
            t_up = ev_time(lambda: ce.upscale_latent_dev(lat1.data_ptr(), (33, H, W), lat2.data_ptr()), se_)
            ce.set_profiling(True)
            ce.upscale_latent_dev(lat1.data_ptr(), (33, H, W), lat2.data_ptr())
            pu = ce.get_profile()
            ce.set_profiling(False)
            extras["vae_encode"] = dict(desc="VideoEncoder + latent normalisation, 768x512", image_ms=t_img, clip_25f_ms=t_clip,
                                        clip_frames_per_s=(8 * (F - 1) + 1) * 1e3 / t_clip)
            extras["latent_upscale"] = dict(desc="upsampleLatents 33x16x24 -> 33x32x48 (stage switch of BASELINE config 5)",
                                            ms=t_up, kernel_classes={k: v for k, v in pu.items() if v["launches"]})
            ce.close()
            del px, zl, px25, zl25, lat1, lat2
        except Exception as e:
            extras["vae_encode"] = dict(error=str(e))
        tg = time_guided()
        tv = time_vae121()
        extras["guided_cfg3"] = dict(desc="dev CFG 4.0 + STG 0.5 (3 forwards/step), 1 GPU", ms_per_step=tg, steps_per_s=1e3 / tg)
        extras["vae_121f"] = dict(desc="VAE decode 768x512x121f, 1 GPU", ms_per_decode=tv, frames_per_s=121e3 / tv)
    else:
        from ltx_video_swift_mlx_b200 import dist as ltxdist
        # (1) Ulysses: ONE video's step strong-scaled over all ranks
        ltxdist.init_context(ctx, sp_size=world, pass_groups=1)
        ctx.denoise_begin(noise[0].numpy(), (F, H, W), sigmas[0], text, None)
        for i in range(3):
            ctx.denoise_step(pairs[i][0], pairs[i][1], i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for i in range(args.steps):
            sg, sn = pairs[i % len(pairs)]
            ctx.denoise_step(sg, sn, i % len(pairs))
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b) / args.steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ctx.set_profiling(True)
        ctx.denoise_step(pairs[0][0], pairs[0][1], 0)
        sp_prof = ctx.get_profile()
        ctx.set_profiling(False)
        extras["ulysses"] = dict(desc=f"one video, sequence-parallel sp={world} (strong scaling)", ms_per_step=float(t.item()),
                                 steps_per_s=1e3 / float(t.item()), kernel_classes={k: v for k, v in sp_prof.items() if v["launches"]})
        # temporally sharded VAE decode of the 121-frame clip on the same communicator
        tv = time_vae121()
        extras["vae_121f_sharded"] = dict(desc=f"VAE decode 768x512x121f, {world} temporal shards + halo exchange",
                                          ms_per_decode=tv, frames_per_s=121e3 / tv)
        ctx.dist_shutdown()
        # (2) pass-parallel guidance (BASELINE config 3): conditional (+ STG, sharing the prefix) on one GPU group, the
        # unconditional pass on the other; each group runs its forwards sequence-parallel over world/2 GPUs
        groups = 2
        ltxdist.init_context(ctx, sp_size=world // groups, pass_groups=groups)
        tg = time_guided()
        extras["guided_cfg3_pass_parallel"] = dict(
            desc=f"dev CFG 4.0 + STG 0.5, passes split over {groups} GPU groups x Ulysses sp={world // groups}",
            ms_per_step=tg, steps_per_s=1e3 / tg)
        ctx.dist_shutdown()
        # (3) BASELINE config 5, stage 2: 1536x1024x257 frames -> N = 33*32*48 = 50688 tokens, int8 weights, the 3-step
        # refinement schedule, Ulysses over all GPUs (the upscaler is out of scope: seeded random latent of that shape)
        if world >= 4 and not args.no_cfg5:
            try:
                from ltx_video_swift_mlx_b200.scheduler import STAGE_2_DISTILLED_SIGMA_VALUES as S2
                ctx.close()
                torch.cuda.empty_cache()
                F5, H5, W5 = 33, 32, 48
                cq = LtxContext(LTXTransformerConfig(), local_rank)
                cq.init_random_weights(1, seed=99)          # same seed on every rank: sequence parallelism shards one model
                cq.finalize_weights(quant_bits=8)
                ltxdist.init_context(cq, sp_size=world, pass_groups=1)
                sq = torch.cuda.ExternalStream(cq.stream, device=torch.device("cuda", local_rank))
                n5 = torch.randn(CIN, F5, H5, W5, generator=torch.Generator().manual_seed(77))
                cq.denoise_begin(n5.numpy(), (F5, H5, W5), S2[0], text, None)
                cq.denoise_step(S2[0], S2[1], 0)                # builds the RoPE table, the text cache and the peer buffers
                barrier()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(sq)
                for i in range(3):
                    cq.denoise_step(S2[i], S2[i + 1], i)
                b.record(sq)
                barrier()
                t5 = torch.tensor([a.elapsed_time(b) / 3], device="cuda")
                dist.all_reduce(t5, op=dist.ReduceOp.MAX)
                cq.set_profiling(True)
                cq.denoise_step(S2[0], S2[1], 0)
                p5 = cq.get_profile()
                cq.set_profiling(False)
                fl5 = dit_flops_per_step(F5 * H5 * W5, S)
                extras["cfg5_stage2_qint8_ulysses"] = dict(
                    desc=f"two-stage refinement step at 1536x1024x257f (N={F5 * H5 * W5}), int8 weights, Ulysses sp={world}, "
                         f"peer-memory exchange={bool(cq.lib.ltx_dist_p2p_active(cq.handle))}",
                    ms_per_step=float(t5.item()), steps_per_s=1e3 / float(t5.item()),
                    aggregate_tflops=fl5 / (float(t5.item()) * 1e-3) / 1e12,
                    kernel_classes={k: v for k, v in p5.items() if v["launches"]})
                cq.dist_shutdown()
                cq.close()
            except Exception as e:   # report, do not hide
                extras["cfg5_stage2_qint8_ulysses"] = dict(error=str(e))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    pk = peaks()
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (FFN-in launch), if present
    traffic, traffic_note = None, None
    try:
        import csv
        rows = list(csv.reader(open(os.path.join(ROOT, "profiles", "r01c_gemm_pair_ffn_in_ncu_raw.csv"))))
        hdr, units, last = rows[0], rows[1], rows[-1]

        def _bytes(k):
            v, u = float(last[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        traffic = _bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")
        traffic_note = "bytes of ONE FFN-in launch (M=1536,N=16384,K=4096; algorithmic 197.1e6) from profiles/r01c_gemm_pair_ffn_in_ncu_raw.csv"
    except Exception:
        pass
    gemm = prof["gemm"]
    achieved = gemm["flops"] / (gemm["ms"] * 1e-3) / 1e12 if gemm["ms"] > 0 else 0.0
    conv = vprof["conv3d"]
    cpu_sec, cores, desc = cpu_port_step_seconds(1, 1) if not args.no_cpu_baseline else (float("nan"), 0, "skipped")
    line = dict(
        metric="dit_steps_per_s", value=value, unit="steps/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
        ms_per_step=ms_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="bf16", data="synthetic",
        config=dict(workload="LTX-2 distilled 13B-video DiT denoise step (BASELINE config 2): 768x512x25f -> N=1536 tokens, "
                             "S=1024 text tokens, 48 blocks, D=4096, 32 heads, bf16 weights random-init, fp32 residual stream",
                    global_batch=world, l2="weights per step (26 GB) >> 126 MB L2; no flush needed",
                    parallelism="replicas" if world > 1 else "single",
                    flops_per_step=dit_flops_per_step(N, S), step_tflops=dit_flops_per_step(N, S) / (ms_per_step * 1e-3) / 1e12),
        e2e=dict(value=world * 1e3 / e2e_ms, unit="steps/s", ms_per_step=e2e_ms, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                 path="ltx_dit_forward + ltx_guided_euler_step with pinned host buffers (text cached by context_key)"),
        gpu_launches=int(launches),
        roofline=dict(bound="tensor", kernel="gemm_bf16_2cta / gemm_bf16_tcgen05 (all GEMM launches of one step)", achieved=achieved,
                      peak=pk["tflops_sustained"], unit="TFLOP/s", frac=achieved / pk["tflops_sustained"], traffic=traffic,
                      traffic_note=traffic_note,
                      peak_source=pk["source"] + ", sustained bf16", launches=gemm["launches"], ms=gemm["ms"]),
        kernel_classes={k: v for k, v in prof.items() if v["launches"]},
        cpu_baseline=dict(value=1.0 / cpu_sec if cpu_sec == cpu_sec else None, unit="steps/s", cores=cores, kind="port", sample=desc),
        vae=dict(metric="vae_frames_per_s", value=world * n_frames / (vae_ms * 1e-3), unit="frames/s", ms_per_decode=vae_ms,
                 frames=n_frames, e2e_value=n_frames / (vae_e2e_ms * 1e-3), e2e_ms=vae_e2e_ms,
                 conv_tflops=conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] > 0 else None,
                 conv_frac_of_peak=(conv["flops"] / (conv["ms"] * 1e-3) / 1e12) / pk["tflops_sustained"] if conv["ms"] > 0 else None,
                 kernel_classes={k: v for k, v in vprof.items() if v["launches"]}),
        clocks=clocks,
        extras=extras,
    )
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cfg5", action="store_true", help="skip the 50688-token int8 Ulysses extra of the >= 4-GPU runs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
