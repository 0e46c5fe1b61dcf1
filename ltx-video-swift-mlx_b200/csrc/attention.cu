// Flash attention forward for sm_100a (head_dim 128, bf16 in/out, fp32 softmax), used for both the video-token
// self-attention and the text cross-attention of the LTX-2 DiT block.
//
// Replaces MLXFast.scaledDotProductAttention (T/LTXAttention.swift:192-211):
//   O = softmax(Q K^T * scale + key_bias) V,   key_bias = (1 - mask) * -10000 (T/LTXTransformer.swift:141-156) or none.
//
// One CTA per (128-query tile, head, batch), 192 threads:
//   warp 0   : TMA producer -- Q tile once, then K / V^T tiles (128 keys) through a 2-stage smem ring
//   warp 1   : MMA issuer   -- S = Q K^T (tcgen05.mma 128x128x16, accumulator in TMEM, double-buffered),
//                              O += P V  (A = P from smem, B = V^T tile, accumulator in TMEM)
//   warps 2-5: softmax      -- thread = query row: tcgen05.ld S row, online max/sum in fp32 (exp2 domain),
//                              rescale O in TMEM when the running max moves, write P (bf16) into 128B-swizzled smem
// V is consumed as V^T [head_dim, keys] (K-major for the PV product); the V-projection GEMM writes it in that layout.
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int ATT_THREADS = 192;
constexpr int TQ = 128, TK = 128, HD = 128;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column swizzled halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 64 * 2;
constexpr size_t ATT_SMEM = 1024 + 6 * TILE_BYTES + 16 * 8 + 16;
constexpr uint32_t ATT_TMEM_COLS = 512;  // S0 [0,128) S1 [128,256) O [256,384)

struct AttnParams {
  int B, H, Nq, Nk;
  float scale_log2;        // softmax scale * log2(e)
  const float* key_bias;   // nullable, [B, Nk]
  bf16* O;
  int64_t ldo;
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TILE_BYTES;       // 2 stages
  uint8_t* sV = sK + 2 * TILE_BYTES;   // 2 stages
  uint8_t* sP = sV + 2 * TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;   // [2]
  uint64_t* k_empty = bars + 3;  // [2]
  uint64_t* v_full = bars + 5;   // [2]
  uint64_t* v_empty = bars + 7;  // [2]
  uint64_t* s_full = bars + 9;   // [2]
  uint64_t* s_empty = bars + 11; // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* o_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Nk + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
    }
    mbar_init(p_full, 128);
    mbar_init(o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_O = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = b * p.Nq + q_tile * TQ;
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(sQ, &tmQ, q_full, h * HD, qrow);
      tma_load_2d(sQ + HALF_BYTES, &tmQ, q_full, h * HD + 64, qrow);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        const int krow = b * p.Nk + j * TK;
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TILE_BYTES);
        tma_load_2d(sK + s * TILE_BYTES, &tmK, &k_full[s], h * HD, krow);
        tma_load_2d(sK + s * TILE_BYTES + HALF_BYTES, &tmK, &k_full[s], h * HD + 64, krow);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TILE_BYTES);
        // V^T is a 3-D map (keys, features, batch): keys past Nk are zero-filled, never another batch's columns
        tma_load_3d(sV + s * TILE_BYTES, &tmV, &v_full[s], j * TK, h * HD, b);
        tma_load_3d(sV + s * TILE_BYTES + HALF_BYTES, &tmV, &v_full[s], j * TK + 64, h * HD, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
      mbar_wait(q_full, 0);
      auto issue_S = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_full[s], ph);
        mbar_wait(&s_empty[s], ph ^ 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(sK + s * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16(tmem_base + s * 128, umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), idesc, kk != 0);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      issue_S(0);
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) issue_S(j + 1);
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&v_full[s], ph);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(sV + s * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16(tmem_O, umma_desc_sw128(p_addr + off), umma_desc_sw128(v_addr + off), idesc, (j | kk) != 0);
        }
        umma_commit(&v_empty[s]);
        umma_commit(o_done);
      }
    }
  } else {
    const int q = warp & 3;
    const int r_in = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    const float LOG2E = 1.4426950408889634f;
    for (int j = 0; j < n_kv; ++j) {
      const int s = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      const int kv0 = j * TK;
      const int valid = min(TK, p.Nk - kv0);
      const float* kb = p.key_bias ? p.key_bias + static_cast<int64_t>(b) * p.Nk + kv0 : nullptr;
      const uint32_t s_addr = tmem_base + lane_addr + s * 128;
      mbar_wait(&s_full[s], ph);
      tc_fence_after();
      // pass 1: row max of the scaled (+biased) logits
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_addr + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          float v = __uint_as_float(r[i]) * p.scale_log2;
          if (kb) v += (col < valid ? kb[col] : 0.f) * LOG2E;
          if (col < valid) mx = fmaxf(mx, v);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f(m_run - m_new);
      // rescale the running output once the previous P V product has landed
      if (j > 0) {
        mbar_wait(o_done, (j - 1) & 1);
        tc_fence_after();
        if (__any_sync(0xffffffffu, alpha != 1.0f)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_O + lane_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(tmem_O + lane_addr + c * 32, r);
          }
          tmem_st_wait();
        }
      }
      // pass 2: P = exp2(S - m_new) -> bf16 -> swizzled smem (A operand of the P V product)
      float rowsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tmem_ld32(s_addr + c * 32, r);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pv[2];
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const int col = c * 32 + i + t;
            float v = __uint_as_float(r[i + t]) * p.scale_log2;
            if (kb) v += (col < valid ? kb[col] : 0.f) * LOG2E;
            pv[t] = (col < valid) ? exp2f(v - m_new) : 0.f;
          }
          rowsum += pv[0] + pv[1];
          pk[i >> 1] = pack_bf16(pv[0], pv[1]);
        }
        uint8_t* prow = sP + (c >> 1) * HALF_BYTES + r_in * 128;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int chunk = ((c & 1) * 4 + t) ^ (r_in & 7);
          *reinterpret_cast<uint4*>(prow + chunk * 16) = make_uint4(pk[4 * t], pk[4 * t + 1], pk[4 * t + 2], pk[4 * t + 3]);
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      tc_fence_before();
      mbar_arrive(&s_empty[s]);
      fence_proxy_async_smem();
      mbar_arrive(p_full);
    }
    // epilogue: O / l -> bf16
    mbar_wait(o_done, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    const int qi = q_tile * TQ + r_in;
    bf16* orow = p.O + (static_cast<int64_t>(b) * p.Nq + qi) * p.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_O + lane_addr + c * 32, r);
      tmem_ld_wait();
      if (qi < p.Nq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 pk = make_uint4(pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                                pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                                pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                                pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
          *reinterpret_cast<uint4*>(orow + c * 32 + i) = pk;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

}  // namespace

void launch_attention(const bf16* Q, int64_t ldq, const bf16* K, int64_t ldk, const bf16* Vt, int64_t ldvb,
                      const float* key_bias, bf16* O, int64_t ldo, int B, int H, int Nq, int Nk, int D, float scale,
                      cudaStream_t stream) {
  LTX_CHECK(D == H * HD, 2, "attention: head_dim must be 128");
  LTX_CHECK(Nq > 0 && Nk > 0 && B > 0, 2, "attention: empty problem");
  LTX_CHECK(ldvb % 8 == 0 && ldvb >= Nk, 2, "attention: V^T per-batch pitch must be a multiple of 8 and >= Nk");
  static bool configured = false;
  if (!configured) {
    LTX_CUDA(cudaFuncSetAttribute(attention_fwd_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(ATT_SMEM)));
    configured = true;
  }
  CUtensorMap tmQ = make_tmap_2d(Q, static_cast<uint64_t>(B) * Nq, D, ldq, 128);
  CUtensorMap tmK = make_tmap_2d(K, static_cast<uint64_t>(B) * Nk, D, ldk, 128);
  CUtensorMap tmV = make_tmap_3d(Vt, Nk, D, B, static_cast<uint64_t>(B) * ldvb, ldvb, 64, 128);
  AttnParams p;
  p.B = B; p.H = H; p.Nq = Nq; p.Nk = Nk;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.key_bias = key_bias;
  p.O = O;
  p.ldo = ldo;
  dim3 grid((Nq + TQ - 1) / TQ, H, B);
  attention_fwd_tcgen05<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  LTX_CUDA(cudaGetLastError());
}

}  // namespace ltx
