"""N-GPU == 1-GPU parity (pass-parallel guidance, temporally sharded VAE, Ulysses) -- needs >= 2 GPUs on the box."""
import os
import subprocess
import sys

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_multi_gpu_matches_single_gpu():
    n = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-2000:])
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_two_devices_in_one_process_match_single_gpu():
    """ltx_dist_init_local: two contexts on two devices inside ONE process (the Swift pipeline is one actor in one process,
    Pipeline/LTXPipeline.swift:117), driven by one host thread per context; Ulysses over same-process peer memory, then the
    temporally sharded VAE.  Same answers as a single context."""
    import threading

    import numpy as np

    from helpers import O, product
    ctxmod = product()
    ocfg = O.DiTConfig(num_layers=3, num_heads=4, head_dim=128, caption_channels=192)
    vcfg = O.VAEConfig(base_channels=512, blocks_per_stage=1)
    pcfg = ctxmod.LTXTransformerConfig(num_layers=3, num_attention_heads=4, caption_channels=192, vae_base_channels=512,
                                       vae_blocks_per_stage=1)
    w, vw = O.make_dit_weights(ocfg, 7), O.make_vae_weights(vcfg, 8)
    g = torch.Generator().manual_seed(5)
    fhw = (2, 4, 8)
    noise = torch.randn(1, 128, *fhw, generator=g)
    text = torch.randn(1, 40, 192, generator=g).bfloat16()
    sigmas = O.set_timesteps(4, False, 64)
    z = torch.randn(128, 5, 3, 4, generator=g).numpy()

    def make(dev):
        c = ctxmod.LtxContext(pcfg, dev)
        c.load_weights(w)
        c.load_weights(vw, prefix="vae.")
        c.finalize_weights()
        return c

    def work(c, out, key):
        try:
            c.denoise_begin(noise[0].numpy(), fhw, sigmas[0], text, None)
            for i in range(len(sigmas) - 1):
                c.denoise_step(sigmas[i], sigmas[i + 1], i)
            out[key] = (c.denoise_get_latent(), c.vae_decode(z), int(c.lib.ltx_dist_p2p_active(c.handle)))
        except Exception as e:   # surfaced by the assertions below
            out[key] = e

    single = make(0)
    ref = {}
    work(single, ref, "single")
    single.close()
    assert not isinstance(ref["single"], Exception), ref["single"]
    ctxs = [make(0), make(1)]
    ctxmod.LtxContext.dist_init_local(ctxs, sp_size=2, pass_groups=1)
    res = {}
    threads = [threading.Thread(target=work, args=(c, res, i)) for i, c in enumerate(ctxs)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=600)
    for i in range(2):
        assert not isinstance(res.get(i), Exception), res.get(i)
        lat, frames, p2p = res[i]
        assert p2p == 1, "same-process ranks should exchange over peer memory"
        assert O.rel_l2(torch.from_numpy(lat), torch.from_numpy(ref["single"][0])) <= 5e-3
        assert float(np.abs(frames - ref["single"][1]).max()) <= 1e-4

    def shut(c):
        c.dist_shutdown()
    ts = [threading.Thread(target=shut, args=(c,)) for c in ctxs]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    for c in ctxs:
        c.close()
