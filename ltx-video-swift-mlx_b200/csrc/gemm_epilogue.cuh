// Fused GEMM epilogues shared by the bf16 and the dequant-fused kernels: thread = output row, 32 (or 16) columns per call.
#pragma once
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const uint32_t (&r)[32], int row, int col0, int ncols, int M, int N,
                                               const GemmEpi& ep) {
  // One thread owns `row`, columns [col0, col0+ncols), ncols = 32 (or 16 for the tail chunk of a tile).
  if (row >= M || col0 >= N) return;
  const float rowbias = (ep.bias && ep.bias_per_row) ? ep.bias[row] : 0.0f;
  const bool full = (col0 + ncols <= N);   // whole chunk in range -> vector path (ncols is 16 or 32)
  if (col0 + ncols < N) N = col0 + ncols;  // never touch the next tile's columns from a 16-wide tail chunk
  if (MODE == EPI_GATE_RESID) {
    float* xr = ep.resid + static_cast<int64_t>(row) * ep.ldr + col0;
    const float* ga = ep.gate_a ? ep.gate_a + static_cast<int64_t>(row / ep.rows_per_gate) * ep.gate_ld + col0 : nullptr;
    const float* gb = ep.gate_b ? ep.gate_b + col0 : nullptr;
    bf16* sh = ep.shadow ? ep.shadow + static_cast<int64_t>(row) * ep.lds + col0 : nullptr;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j >= ncols) break;
        float4 xv = *reinterpret_cast<const float4*>(xr + j);
        float4 bv = (ep.bias && !ep.bias_per_row) ? *reinterpret_cast<const float4*>(ep.bias + col0 + j)
                                                  : make_float4(rowbias, rowbias, rowbias, rowbias);
        float4 gv = make_float4(1.f, 1.f, 1.f, 1.f);
        if (ga) {
          gv = *reinterpret_cast<const float4*>(ga + j);
          if (gb) {
            float4 t = *reinterpret_cast<const float4*>(gb + j);
            gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w;
          }
        }
        xv.x += (__uint_as_float(r[j + 0]) + bv.x) * gv.x * ep.scale;
        xv.y += (__uint_as_float(r[j + 1]) + bv.y) * gv.y * ep.scale;
        xv.z += (__uint_as_float(r[j + 2]) + bv.z) * gv.z * ep.scale;
        xv.w += (__uint_as_float(r[j + 3]) + bv.w) * gv.w * ep.scale;
        *reinterpret_cast<float4*>(xr + j) = xv;
        if (sh) {
          uint2 pk = make_uint2(pack_bf16(xv.x, xv.y), pack_bf16(xv.z, xv.w));
          *reinterpret_cast<uint2*>(sh + j) = pk;
        }
      }
    } else {
      for (int j = 0; j < 32 && col0 + j < N; ++j) {
        float b = ep.bias ? (ep.bias_per_row ? rowbias : ep.bias[col0 + j]) : 0.f;
        float g = ga ? (ga[j] + (gb ? gb[j] : 0.f)) : 1.f;
        float v = xr[j] + (__uint_as_float(r[j]) + b) * g * ep.scale;
        xr[j] = v;
        if (sh) sh[j] = __float2bfloat16(v);
      }
    }
  } else if (MODE == EPI_F32) {
    float* o = reinterpret_cast<float*>(ep.out) + static_cast<int64_t>(row) * ep.ldo + col0;
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j >= ncols) break;
        float4 bv = (ep.bias && !ep.bias_per_row) ? *reinterpret_cast<const float4*>(ep.bias + col0 + j)
                                                  : make_float4(rowbias, rowbias, rowbias, rowbias);
        float4 v = make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                               __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
        *reinterpret_cast<float4*>(o + j) = v;
      }
    } else {
      for (int j = 0; j < 32 && col0 + j < N; ++j) {
        float b = ep.bias ? (ep.bias_per_row ? rowbias : ep.bias[col0 + j]) : 0.f;
        o[j] = __uint_as_float(r[j]) + b;
      }
    }
  } else {  // EPI_BF16 / EPI_GELU_BF16 / EPI_SILU_BF16
    bf16* o = reinterpret_cast<bf16*>(ep.out) + static_cast<int64_t>(row) * ep.ldo + col0;
    if (ep.col_block > 0)
      o = reinterpret_cast<bf16*>(ep.out) + static_cast<int64_t>(col0 / ep.col_block) * ep.col_block_stride +
          static_cast<int64_t>(row) * ep.ldo + (col0 % ep.col_block);
    if (full) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) {
        if (j >= ncols) break;
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; t += 4) {
          float4 bv = (ep.bias && !ep.bias_per_row) ? *reinterpret_cast<const float4*>(ep.bias + col0 + j + t)
                                                    : make_float4(rowbias, rowbias, rowbias, rowbias);
          v[t + 0] = __uint_as_float(r[j + t + 0]) + bv.x;
          v[t + 1] = __uint_as_float(r[j + t + 1]) + bv.y;
          v[t + 2] = __uint_as_float(r[j + t + 2]) + bv.z;
          v[t + 3] = __uint_as_float(r[j + t + 3]) + bv.w;
        }
        if (MODE == EPI_GELU_BF16) {
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = gelu_tanh(v[t]);
        }
        if (MODE == EPI_SILU_BF16) {
#pragma unroll
          for (int t = 0; t < 8; ++t) v[t] = silu(v[t]);
        }
        uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        *reinterpret_cast<uint4*>(o + j) = pk;
      }
    } else {
      for (int j = 0; j < 32 && col0 + j < N; ++j) {
        float b = ep.bias ? (ep.bias_per_row ? rowbias : ep.bias[col0 + j]) : 0.f;
        float v = __uint_as_float(r[j]) + b;
        if (MODE == EPI_GELU_BF16) v = gelu_tanh(v);
        if (MODE == EPI_SILU_BF16) v = silu(v);
        o[j] = __float2bfloat16(v);
      }
    }
  }
}

}  // namespace ltx
