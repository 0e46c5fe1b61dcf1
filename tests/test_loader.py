"""Checkpoint key mapping (ltx_map_weight_key; CPU) and the safetensors loader (ltx_load_safetensors; GPU).

The mapping rules are the reference's mapTransformerKey / mapVAEWeights (Utils/ModelDownloader.swift:756-899) and the
filter of loadTransformerWeights (:617-629); the expected names are SURVEY Appendix C."""
import json
import os
import struct

import numpy as np
import pytest
import torch

from helpers import O, product, rel_l2, small_dit_config, small_vae_config


def mk(which, key):
    return product().map_weight_key(which, key)


def test_transformer_key_mapping():
    P = "model.diffusion_model."
    cases = {
        P + "proj_in.weight": "patchify_proj.weight",
        P + "time_embed.emb.timestep_embedder.linear_1.weight": "adaln_single.emb.linear_1.weight",
        P + "time_embed.linear.bias": "adaln_single.linear.bias",
        P + "adaln_single.emb.timestep_embedder.linear_2.bias": "adaln_single.emb.linear_2.bias",
        P + "caption_projection.linear_1.weight": "caption_projection.linear_1.weight",
        P + "transformer_blocks.7.attn1.norm_q.weight": "transformer_blocks.7.attn1.q_norm.weight",
        P + "transformer_blocks.7.attn2.norm_k.weight": "transformer_blocks.7.attn2.k_norm.weight",
        P + "transformer_blocks.7.attn1.to_out.0.weight": "transformer_blocks.7.attn1.to_out.weight",
        P + "transformer_blocks.7.ff.net.0.proj.bias": "transformer_blocks.7.ff.project_in.proj.bias",
        P + "transformer_blocks.7.ff.net.2.weight": "transformer_blocks.7.ff.project_out.weight",
        P + "transformer_blocks.7.scale_shift_table": "transformer_blocks.7.scale_shift_table",
        P + "scale_shift_table": "scale_shift_table",
        P + "proj_out.bias": "proj_out.bias",
        "transformer_blocks.0.attn1.to_q.weight": "transformer_blocks.0.attn1.to_q.weight",   # already-stripped key
    }
    for k, want in cases.items():
        assert mk(1, k) == want, k
    skipped = [
        P + "audio_proj_in.weight", P + "transformer_blocks.3.audio_attn1.to_q.weight", P + "av_cross_attn_video_scale_shift.weight",
        P + "transformer_blocks.3.video_to_audio_attn.to_q.weight", P + "transformer_blocks.3.scale_shift_table_a2v_ca_video",
        P + "video_embeddings_connector.transformer_1d_blocks.0.attn1.to_q.weight", P + "audio_embeddings_connector.x",
        P + "transformer_blocks.3.attn1.to_q.weight_scale", P + "transformer_blocks.3.attn1.to_q.input_scale",
        "vocoder.conv_pre.weight", P + "transformer_blocks.3.av_ca_a2v_gate_adaln_single.linear.weight",
    ]
    for k in skipped:
        assert mk(1, k) is None, k


def test_vae_key_mapping():
    cases = {
        "decoder.conv_in.conv.weight": "vae.conv_in.conv.weight",
        "decoder.mid_block.resnets.2.conv1.conv.weight": "vae.up_blocks_0.res_blocks.2.conv1.conv.weight",
        "decoder.mid_block.time_embedder.timestep_embedder.linear_1.weight": "vae.up_blocks_0.time_embedder.timestep_embedder.linear_1.weight",
        "decoder.up_blocks.0.upsamplers.0.conv.conv.weight": "vae.up_blocks_1.conv.conv.weight",
        "decoder.up_blocks.0.resnets.4.scale_shift_table": "vae.up_blocks_2.res_blocks.4.scale_shift_table",
        "decoder.up_blocks.1.upsamplers.0.conv.conv.bias": "vae.up_blocks_3.conv.conv.bias",
        "decoder.up_blocks.2.resnets.0.conv2.conv.bias": "vae.up_blocks_6.res_blocks.0.conv2.conv.bias",
        "decoder.up_blocks.5.conv.conv.weight": "vae.up_blocks_5.conv.conv.weight",               # legacy flat numbering
        "decoder.up_blocks.4.res_blocks.1.conv1.conv.weight": "vae.up_blocks_4.res_blocks.1.conv1.conv.weight",
        "vae.decoder.conv_out.conv.bias": "vae.conv_out.conv.bias",                                # unified checkpoint
        "vae.per_channel_statistics.mean-of-means": "vae.mean_of_means",
        "per_channel_statistics.std-of-means": "vae.std_of_means",
        "latents_mean": "vae.mean_of_means", "latents_std": "vae.std_of_means",
        "decoder.last_scale_shift_table": "vae.last_scale_shift_table",
        "decoder.timestep_scale_multiplier": "vae.timestep_scale_multiplier",
    }
    for k, want in cases.items():
        assert mk(2, k) == want, k
    for k in ("encoder.conv_in.conv.weight", "vae.encoder.down_blocks.0.resnets.0.conv1.conv.weight",
              "per_channel_statistics.channel"):
        assert mk(2, k) is None, k


# ---------------------------------------------------------------------------------------------------- safetensors writer
def write_safetensors(path, tensors):
    """Minimal writer of the published format: u64 header length, JSON header, raw little-endian data."""
    header, blobs, off = {"__metadata__": {"format": "pt", "note": "written by tests/test_loader.py"}}, [], 0
    for name, t in tensors.items():
        t = t.detach().cpu().contiguous()
        if t.dtype == torch.bfloat16:
            raw, dt = t.view(torch.uint16).numpy().tobytes(), "BF16"
        elif t.dtype == torch.float16:
            raw, dt = t.numpy().tobytes(), "F16"
        else:
            raw, dt = t.float().numpy().tobytes(), "F32"
        header[name] = {"dtype": dt, "shape": list(t.shape), "data_offsets": [off, off + len(raw)]}
        blobs.append(raw)
        off += len(raw)
    hj = json.dumps(header, separators=(",", ":")).encode()
    hj += b" " * ((8 - len(hj) % 8) % 8)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(hj)))
        f.write(hj)
        for b in blobs:
            f.write(b)


def _checkpoint_name_dit(k):
    """Inverse of mapTransformerKey: our post-mapping name -> the name in the unified checkpoint."""
    k = k.replace("patchify_proj.", "proj_in.").replace("adaln_single.emb.", "time_embed.emb.timestep_embedder.")
    k = k.replace("adaln_single.linear.", "time_embed.linear.").replace(".q_norm.", ".norm_q.").replace(".k_norm.", ".norm_k.")
    k = k.replace(".to_out.", ".to_out.0.").replace("ff.project_in.proj.", "ff.net.0.proj.").replace("ff.project_out.", "ff.net.2.")
    return "model.diffusion_model." + k


def _checkpoint_name_vae(k):
    """Inverse of mapVAEWeights (Diffusers layout: mid_block + up_blocks.{i}.{resnets,upsamplers.0})."""
    if k == "mean_of_means":
        return "latents_mean"
    if k == "std_of_means":
        return "latents_std"
    for i in range(7):
        p = f"up_blocks_{i}."
        if k.startswith(p):
            rest = k[len(p):].replace("res_blocks.", "resnets.")
            if i == 0:
                return "decoder.mid_block." + rest
            if i % 2 == 1:
                return f"decoder.up_blocks.{(i - 1) // 2}.upsamplers.0." + rest
            if rest.startswith("resnets."):
                return f"decoder.up_blocks.{(i - 2) // 2}." + rest
            return f"decoder.up_blocks.{i}." + rest          # time embedder of a res group: legacy flat numbering
    return "decoder." + k


@pytest.mark.gpu
def test_load_safetensors_dit_matches_direct_load(tmp_path):
    ocfg, pcfg = small_dit_config(2, 2)
    w = O.make_dit_weights(ocfg, seed=11)
    ckpt = {_checkpoint_name_dit(k): (v.bfloat16() if v.ndim >= 2 else v) for k, v in w.items()}
    # tensors the loader must ignore
    ckpt["model.diffusion_model.audio_proj_in.weight"] = torch.randn(8, 8)
    ckpt["model.diffusion_model.video_embeddings_connector.learnable_registers"] = torch.randn(4, 8)
    ckpt["vae.decoder.conv_in.conv.weight"] = torch.randn(4, 4, 3, 3, 3)
    ckpt["model.diffusion_model.transformer_blocks.0.attn1.to_q.weight_scale"] = torch.randn(1)
    path = os.path.join(tmp_path, "unified.safetensors")
    write_safetensors(path, ckpt)
    P = product()
    a = P.LtxContext(pcfg, 0)
    assert a.load_safetensors(path, 1) == len(w)
    a.finalize_weights()
    b = P.LtxContext(pcfg, 0)
    b.load_weights(w)
    b.finalize_weights()
    g = torch.Generator().manual_seed(4)
    fhw, S = (2, 4, 6), 40
    lat = torch.randn(1, 48, ocfg.in_channels, generator=g).bfloat16()
    cx = torch.randn(1, S, ocfg.caption_channels, generator=g).bfloat16()
    sig = np.array([0.6], dtype=np.float32)
    va = a.dit_forward(lat, cx, sig, None, fhw)
    vb = b.dit_forward(lat, cx, sig, None, fhw)
    assert np.array_equal(va, vb)                      # same tensors on the device -> bit-identical velocity
    ref = O.dit_forward(w, ocfg, lat.float(), cx.float(), torch.tensor([0.6]), None, fhw)
    assert rel_l2(va, ref) <= 1e-2
    with pytest.raises(P.LtxError):
        a.load_safetensors(os.path.join(tmp_path, "missing.safetensors"), 1)
    a.close(); b.close()


@pytest.mark.gpu
def test_load_safetensors_vae_matches_direct_load(tmp_path):
    ocfg, pcfg = small_vae_config(base=512, blocks=1)
    w = O.make_vae_weights(ocfg, seed=5)
    w = {k: (O.bf16_round(v) if (k.endswith(".weight") and v.ndim >= 2) else v) for k, v in w.items()}
    ckpt = {}
    for k, v in w.items():
        name = _checkpoint_name_vae(k)
        ckpt[name] = v.reshape(1, -1, 1, 1, 1) if name in ("latents_mean", "latents_std") else v
    ckpt["encoder.conv_in.conv.weight"] = torch.randn(4, 4, 3, 3, 3)
    path = os.path.join(tmp_path, "vae.safetensors")
    write_safetensors(path, ckpt)
    P = product()
    a = P.LtxContext(pcfg, 0)
    assert a.load_safetensors(path, 2) == len(w)
    a.finalize_weights()
    b = P.LtxContext(pcfg, 0)
    b.load_weights(w, prefix="vae.")
    b.finalize_weights()
    lat = torch.randn(128, 2, 4, 4, generator=torch.Generator().manual_seed(8)).numpy()
    fa, fb = a.vae_decode(lat), b.vae_decode(lat)
    assert np.array_equal(fa, fb)
    a.close(); b.close()


def test_vae_encoder_and_upscaler_key_mapping():
    """mapVAEEncoderWeights (Utils/ModelDownloader.swift:1224-1280) and loadSpatialUpscaler's key handling
    (Models/Upscaler/SpatialUpscaler.swift:262-300)."""
    enc = {
        "encoder.conv_in.conv.weight": "vae_encoder.conv_in.conv.weight",
        "encoder.down_blocks.0.resnets.3.conv1.conv.weight": "vae_encoder.down_blocks_0.resnets.resnets.3.conv1.conv.weight",
        "encoder.down_blocks.2.downsamplers.0.conv.conv.bias": "vae_encoder.down_blocks_2.downsamplers.conv.conv.bias",
        "encoder.mid_block.resnets.1.conv2.conv.weight": "vae_encoder.mid_block.resnets.1.conv2.conv.weight",
        "encoder.conv_out.conv.bias": "vae_encoder.conv_out.conv.bias",
        "vae.encoder.down_blocks.3.resnets.0.conv2.conv.bias": "vae_encoder.down_blocks_3.resnets.resnets.0.conv2.conv.bias",
    }
    for k, want in enc.items():
        assert mk(3, k) == want, k
    for k in ("decoder.conv_in.conv.weight", "latents_mean", "per_channel_statistics.mean-of-means"):
        assert mk(3, k) is None, k
    assert mk(2, "encoder.conv_in.conv.weight") is None          # the decoder loader still skips encoder tensors
    ups = {
        "initial_conv.weight": "upscaler.initial_conv.weight",
        "res_blocks.2.norm1.bias": "upscaler.res_blocks.2.norm1.bias",
        "upsampler.conv.weight": "upscaler.upsampler.conv.weight",
        "post_upsample_res_blocks.0.conv2.weight": "upscaler.post_upsample_res_blocks.0.conv2.weight",
        "final_conv.bias": "upscaler.final_conv.bias",
    }
    for k, want in ups.items():
        assert mk(4, k) == want, k
    assert mk(4, "upsampler.blur_down.kernel") is None


def test_audio_video_transformer_key_mapping():
    """loadTransformerWeights(includeAudio: true) (Utils/ModelDownloader.swift:605-639, 756-803): audio / cross-modal tensors are
    kept (with the general renames), vocoder and connector tensors are not transformer weights."""
    P = "model.diffusion_model."
    cases = {
        P + "audio_patchify_proj.weight": "audio_patchify_proj.weight",
        P + "audio_adaln_single.emb.timestep_embedder.linear_1.weight": "audio_adaln_single.emb.linear_1.weight",
        P + "av_ca_a2v_gate_adaln_single.emb.timestep_embedder.linear_2.bias": "av_ca_a2v_gate_adaln_single.emb.linear_2.bias",
        P + "transformer_blocks.3.audio_attn1.norm_q.weight": "transformer_blocks.3.audio_attn1.q_norm.weight",
        P + "transformer_blocks.3.audio_to_video_attn.to_out.0.weight": "transformer_blocks.3.audio_to_video_attn.to_out.weight",
        P + "transformer_blocks.3.audio_ff.net.0.proj.weight": "transformer_blocks.3.audio_ff.project_in.proj.weight",
        P + "transformer_blocks.3.audio_ff.net.2.bias": "transformer_blocks.3.audio_ff.project_out.bias",
        P + "transformer_blocks.3.scale_shift_table_a2v_ca_video": "transformer_blocks.3.scale_shift_table_a2v_ca_video",
        P + "transformer_blocks.3.attn1.to_q.weight": "transformer_blocks.3.attn1.to_q.weight",
        P + "proj_in.weight": "patchify_proj.weight",
    }
    for k, want in cases.items():
        assert mk(5, k) == want, k
    for k in ("vocoder.conv_pre.weight", P + "video_embeddings_connector.x", P + "audio_embeddings_connector.x",
              P + "transformer_blocks.3.attn1.to_q.weight_scale"):
        assert mk(5, k) is None, k
