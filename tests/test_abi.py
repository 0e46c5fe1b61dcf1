"""The C-ABI library loads without a GPU, exports every symbol include/ltxcuda.h declares, and fails loudly (no CPU
fallback) when no device is present."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

from helpers import ROOT, product

product()
from ltx_video_swift_mlx_b200 import _lib  # noqa: E402
from ltx_video_swift_mlx_b200.context import LtxContext, LTXTransformerConfig  # noqa: E402

HEADER = os.path.join(ROOT, "include", "ltxcuda.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ltx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    syms = declared_symbols()
    assert len(syms) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (ltx_[a-z0-9_]+)", out))
    assert set(syms) <= exported, sorted(set(syms) - exported)
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    assert lib.ltx_version().startswith(b"ltxcuda")


def test_struct_layouts_match_header():
    lib = _lib.load()
    cfg = _lib.LtxConfig()
    lib.ltx_config_default(ctypes.byref(cfg))
    assert (cfg.num_layers, cfg.num_heads, cfg.head_dim, cfg.caption_channels) == (48, 32, 128, 3840)
    assert list(cfg.max_pos) == [20, 2048, 2048] and cfg.vae_patch_size == 4 and abs(cfg.norm_eps - 1e-6) < 1e-12
    assert ctypes.sizeof(_lib.LtxDitFlags) == 4 * (1 + 64 + 2 + 1 + 64 + 1) + 4 + 8   # incl. padding before the u64
    assert ctypes.sizeof(_lib.LtxStepParams) == 4 * (6 + 1 + 64 + 4)


def test_only_sm100_sass_in_library():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a GPU-less host")
def test_no_cpu_fallback():
    with pytest.raises(_lib.LtxError) as e:
        LtxContext(LTXTransformerConfig(), 0)
    assert e.value.code == 3 and "no CPU fallback" in str(e.value)


def test_pure_host_entry_points_without_a_gpu():
    """ltx_vae_tiled_frames is pure chunk arithmetic (decodeWithTemporalTiling, V/VideoDecoder.swift:525-589): checked here, on the
    GPU-less host, against the same loop written out and against the untiled 8 (F' - 1) + 1."""
    lib = _lib.load()

    def frames(total, tile, overlap):      # the reference's loop, frame counts only
        if tile <= 0 or total <= tile:
            return 8 * (total - 1) + 1
        stride, po, chunks, start = tile - overlap, 8 * overlap, [], 0
        while start < total:
            end = min(start + tile, total)
            chunks.append(8 * (end - start - 1) + 1)
            if end >= total:
                break
            start += stride
        res = chunks[0]
        for nf in chunks[1:]:
            res = res + nf - po if 0 < po < res and po < nf else res + nf
        return res
    for total in (1, 2, 5, 8, 9, 16, 33):
        for tile in (0, 1, 2, 3, 4, 8, 16):
            for overlap in (0, 1, 2, 3):
                got = lib.ltx_vae_tiled_frames(total, tile, overlap)
                if 0 < tile < total and overlap >= tile:
                    assert got == -1, (total, tile, overlap, got)        # the stride would not advance
                else:
                    assert got == frames(total, tile, overlap), (total, tile, overlap, got)
    assert lib.ltx_vae_tiled_frames(16, 8, 1) == 107 and lib.ltx_vae_tiled_frames(0, 8, 1) == -1


def test_conv_plan_never_looks_at_the_frame_count(monkeypatch):
    """ltx_conv3d_plan (host-only): the two launcher decisions that change the ORDER in which the taps are summed -- slab stages
    and the tap split -- and the slab tile geometry must not depend on T, or a temporal shard of a clip would round differently
    from the whole clip (the 8-GPU decode is held to bit-identity with the 1-GPU one).  Swept over every conv shape of the
    decoder at 768 x 512 and a ragged size, T = 1 ... 123, all epilogue modes; plus the expected plans of the decoder's stages."""
    import ctypes as C
    lib = _lib.load()
    for env in ({}, {"LTX_CONV_SLAB": "0"}, {"LTX_CONV_PAIR": "0"}):
        for k in ("LTX_CONV_SLAB", "LTX_CONV_PAIR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)

        def plan(T, H, W, Cin, Cout, mode=0, ntaps=27, sms=148):
            out = (C.c_int32 * 7)()
            assert lib.ltx_conv3d_plan(T, H, W, Cin, Cout, mode, ntaps, sms, out) == 0
            return dict(zip(("bn", "pair", "slab", "bt", "bh", "bw", "ksplit"), list(out)))
        shapes = [(16, 24, 128, 1024), (16, 24, 1024, 1024), (32, 48, 1024, 4096), (32, 48, 512, 512), (64, 96, 512, 2048),
                  (64, 96, 256, 256), (128, 192, 256, 1024), (128, 192, 128, 128), (128, 192, 128, 48), (20, 30, 128, 128),
                  (7, 11, 256, 64), (33, 47, 64, 128)]
        for H, W, Cin, Cout in shapes:
            for mode in ((0, 3, 4) if 64 < Cout <= 256 else (0, 1, 2)):
                ref = plan(1, H, W, Cin, Cout, mode)
                for T in list(range(1, 34)) + [57, 65, 121, 123]:
                    p = plan(T, H, W, Cin, Cout, mode)
                    assert (p["slab"], p["ksplit"], p["pair"]) == (ref["slab"], ref["ksplit"], ref["pair"]), (env, T, H, W, Cin, Cout, mode, p, ref)
                    assert p["bt"] * p["bh"] * p["bw"] == 128 and p["bn"] in (64, 128, 256)
                    if p["slab"]:
                        assert (p["bt"], p["bh"], p["bw"]) == (1, ref["bh"], ref["bw"]) and p["bw"] % 8 == 0 and (p["bh"] + 2) * p["bw"] <= 192
                        assert p["pair"] == 1 and Cout < 256 and p["bn"] <= 128
        if not env:
            assert plan(25, 128, 192, 128, 128, 4) == dict(bn=128, pair=1, slab=1, bt=1, bh=16, bw=8, ksplit=1)    # last stage, hand-over
            assert plan(25, 128, 192, 128, 48, 2)["slab"] == 1 and plan(25, 128, 192, 128, 48, 2)["bn"] == 64      # output conv
            assert plan(4, 16, 24, 1024, 1024, 0)["ksplit"] == 3 and plan(4, 16, 24, 1024, 1024, 0)["slab"] == 0   # tile-starved stage
            assert plan(13, 64, 96, 256, 256, 3)["slab"] == 0 and plan(13, 64, 96, 256, 256, 3)["pair"] == 1
            assert plan(1, 33, 47, 64, 128)["pair"] == 0                                                           # Cin not in 128-channel stages
    bad = (C.c_int32 * 7)()
    assert lib.ltx_conv3d_plan(0, 16, 24, 128, 128, 0, 27, 148, bad) != 0 and lib.ltx_conv3d_plan(4, 16, 24, 128, 128, 0, 5, 148, bad) != 0
    assert lib.ltx_conv3d_plan(4, 16, 24, 128, 128, 0, 27, 148, None) != 0


def test_header_is_plain_c_and_links(tmp_path):
    """include/ltxcuda.h compiled as C99 with -Wall -Werror -pedantic by gcc, linked against the in-tree library and run:
    the boundary a SwiftPM C target would import."""
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", exe, "-L", libdir, "-l:libltxcuda.so",
                    "-Wl,-rpath," + libdir], check=True, capture_output=True, text=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert ("no CPU fallback" in r.stdout) != torch.cuda.is_available()
