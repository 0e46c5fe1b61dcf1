// Flash attention forward for sm_100a (head_dim 128, bf16 in/out, fp32 softmax), used for both the video-token
// self-attention and the text cross-attention of the LTX-2 DiT block.
//
// Replaces MLXFast.scaledDotProductAttention (T/LTXAttention.swift:192-211):
//   O = softmax(Q K^T * scale + key_bias) V,   key_bias = (1 - mask) * -10000 (T/LTXTransformer.swift:141-156) or none.
//
// One CTA per (pair of 128-query tiles, head, batch), 320 threads; the two query tiles run in ping-pong so the tensor
// core works on one tile while the other tile's softmax runs:
//   warp 0    : TMA producer -- both Q tiles once, then K / V^T tiles (128 keys), single-buffered (the ping-pong
//                               schedule leaves a whole softmax phase of slack before the next tile is needed)
//   warp 1    : MMA issuer   -- S_i = Q_i K^T (tcgen05.mma 128x128x16 into TMEM), O_i += P_i V (A = P_i from smem),
//                               issued in the order  PV_0(j), S_0(j+1), PV_1(j), S_1(j+1), ...
//   warps 2-5 : softmax of query tile 0, warps 6-9: softmax of query tile 1 -- thread = query row: one tcgen05.ld pass
//               brings the 128 logits of the row into registers; running max / sum in fp32 (exp2 domain, ex2.approx);
//               the output accumulator in TMEM is rescaled only when the row max grows by more than 2^8 (lazy rescale,
//               exact after the final 1/l normalisation); P is written as bf16 into 128B-swizzled smem.
// tcgen05.commit tracks every earlier MMA of the issuing thread, so "S_i(j+1) landed" implies "PV_i(j) landed": one
// barrier per tile serves as S-ready, O-stable and P-buffer-free.  TMEM: S0 [0,128) S1 [128,256) O0 [256,384) O1 [384,512).
// V is consumed as V^T [head_dim, keys] (K-major for the PV product); the V-projection GEMM writes it in that layout.
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int ATT_THREADS = 320;
constexpr int TQ = 128, TK = 128, HD = 128;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column swizzled halves of 16 KB
constexpr uint32_t HALF_BYTES = 128 * 64 * 2;
constexpr size_t ATT_SMEM = 1024 + 6 * TILE_BYTES + 16 * 8 + 16;  // Q0 Q1 K V P0 P1
constexpr uint32_t ATT_TMEM_COLS = 512;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

struct AttnParams {
  int B, H, Nq, Nk;
  float scale_log2;        // softmax scale * log2(e)
  const float* key_bias;   // nullable, [B, Nk]
  bf16* O;
  int64_t ldo;
  int o_rows_per_block;   // > 0: output row r goes to o_blocks.p[r / o_rows_per_block] (Ulysses over peer memory)
  PeerTable o_blocks;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;                    // 2 tiles
  uint8_t* sK = sQ + 2 * TILE_BYTES;
  uint8_t* sV = sK + TILE_BYTES;
  uint8_t* sP = sV + TILE_BYTES;         // 2 tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * TILE_BYTES);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = bars + 2;
  uint64_t* v_full = bars + 3;
  uint64_t* v_empty = bars + 4;
  uint64_t* s_full = bars + 5;   // [2]
  uint64_t* p_full = bars + 7;   // [2]
  uint64_t* o_final = bars + 9;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_pair = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Nk + TK - 1) / TK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(k_full, 1);
    mbar_init(k_empty, 1);
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_final[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();   // PDL: the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      const int qrow = b * p.Nq + q_pair * 2 * TQ;
      mbar_arrive_expect_tx(q_full, 2 * TILE_BYTES);
      for (int i = 0; i < 2; ++i) {
        tma_load_2d(sQ + i * TILE_BYTES, &tmQ, q_full, h * HD, qrow + i * TQ);
        tma_load_2d(sQ + i * TILE_BYTES + HALF_BYTES, &tmQ, q_full, h * HD + 64, qrow + i * TQ);
      }
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        const int krow = b * p.Nk + j * TK;
        mbar_wait(k_empty, ph ^ 1);
        mbar_arrive_expect_tx(k_full, TILE_BYTES);
        tma_load_2d(sK, &tmK, k_full, h * HD, krow);
        tma_load_2d(sK + HALF_BYTES, &tmK, k_full, h * HD + 64, krow);
        mbar_wait(v_empty, ph ^ 1);
        mbar_arrive_expect_tx(v_full, TILE_BYTES);
        // V^T is a 3-D map (keys, features, batch): keys past Nk are zero-filled, never another batch's columns
        tma_load_3d(sV, &tmV, v_full, j * TK, h * HD, b);
        tma_load_3d(sV + HALF_BYTES, &tmV, v_full, j * TK + 64, h * HD, b);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      auto issue_S = [&](int i) {
        const uint32_t q_addr = smem_u32(sQ + i * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16(tmem_base + i * 128, umma_desc_sw128(q_addr + off), umma_desc_sw128(k_addr + off), idesc, kk != 0);
        }
        umma_commit(&s_full[i]);
      };
      auto issue_PV = [&](int i, int j) {
        const uint32_t p_addr = smem_u32(sP + i * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {
          const uint32_t off = (kk >> 2) * HALF_BYTES + (kk & 3) * 32;
          umma_bf16(tmem_base + 256 + i * 128, umma_desc_sw128(p_addr + off), umma_desc_sw128(v_addr + off), idesc,
                    (j | kk) != 0);
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(k_full, 0);
      tc_fence_after();
      issue_S(0);
      issue_S(1);
      umma_commit(k_empty);
      for (int j = 0; j < n_kv; ++j) {
        const uint32_t ph = j & 1;
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_full[i], ph);
          if (i == 0) mbar_wait(v_full, ph);
          tc_fence_after();
          issue_PV(i, j);
          if (i == 1) umma_commit(v_empty);
          if (j + 1 < n_kv) {
            if (i == 0) {
              mbar_wait(k_full, ph ^ 1);
              tc_fence_after();
            }
            issue_S(i);  // its commit also covers PV_i(j): "S_i(j+1) ready" implies "O_i stable, P_i free"
            if (i == 1) umma_commit(k_empty);
          } else {
            umma_commit(&o_final[i]);
          }
        }
      }
    }
  } else {
    const int wg = (warp - 2) >> 2;   // query tile 0 / 1
    const int q = warp & 3;           // TMEM lane quarter
    const int r_in = q * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + wg * 128;
    const uint32_t o_addr = tmem_base + lane_addr + 256 + wg * 128;
    uint8_t* prow = sP + wg * TILE_BYTES + r_in * 128;
    const float LOG2E = 1.4426950408889634f;
    float m_ref = -1.0e30f, l_run = 0.f;   // m_ref in raw-logit units (before the softmax scale)
    for (int j = 0; j < n_kv; ++j) {
      const int kv0 = j * TK;
      const int valid = min(TK, p.Nk - kv0);
      mbar_wait(&s_full[wg], j & 1);
      tc_fence_after();
      float s[128];
      {
        uint32_t r[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tmem_ld32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(r[i]);
        }
      }
      if (p.key_bias != nullptr) {
        // the additive bias is defined on the scaled logits: fold it in as bias / scale so one FFMA applies both later
        const float inv_scale = LOG2E / p.scale_log2;
        const float* kb = p.key_bias + static_cast<int64_t>(b) * p.Nk + kv0;
        if (valid == TK && (reinterpret_cast<uintptr_t>(kb) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < 128; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(kb + i));
            s[i] = fmaf(bv.x, inv_scale, s[i]);
            s[i + 1] = fmaf(bv.y, inv_scale, s[i + 1]);
            s[i + 2] = fmaf(bv.z, inv_scale, s[i + 2]);
            s[i + 3] = fmaf(bv.w, inv_scale, s[i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i < valid) s[i] = fmaf(__ldg(kb + i), inv_scale, s[i]);
        }
      }
      if (valid < TK) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < 128; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m_ref, mx);
      const bool need = (m_new - m_ref) * p.scale_log2 > RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx((m_ref - m_new) * p.scale_log2) : 1.0f;
        if (j > 0) {
          uint32_t r[32];
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            tmem_ld32(o_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(o_addr + c * 32, r);
          }
          tmem_st_wait();
        }
        l_run *= alpha;
        if (need) m_ref = m_new;
      }
      const float neg = -m_ref * p.scale_log2;
      float rowsum = 0.f;
#pragma unroll
      for (int c = 0; c < 16; ++c) {   // 16 chunks of 8 keys = one 16-byte smem store each
        uint32_t pk[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float p0 = ex2_approx(fmaf(s[c * 8 + 2 * t], p.scale_log2, neg));
          const float p1 = ex2_approx(fmaf(s[c * 8 + 2 * t + 1], p.scale_log2, neg));
          rowsum += p0 + p1;
          pk[t] = pack_bf16(p0, p1);
        }
        const int chunk = (c & 7) ^ (r_in & 7);
        *reinterpret_cast<uint4*>(prow + (c >> 3) * HALF_BYTES + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      l_run += rowsum;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(&p_full[wg]);
    }
    // epilogue: O / l -> bf16
    mbar_wait(&o_final[wg], 0);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    // O / l -> bf16, staged through this tile's P buffer (free now: o_final means every MMA has retired) so that the
    // global stores are whole 256-byte rows (16 lanes x 16 B) -- full lines, also over NVLink when the row lives in a
    // peer's buffer -- instead of 32 rows x 16 B per instruction.  Each warp reads back only the 32 rows it wrote.
    // 16-byte chunk c of row r sits at chunk c ^ (r & 7): conflict-free for the row-per-thread writes and the reads.
    uint8_t* otile = sP + wg * TILE_BYTES;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t r[32];
      tmem_ld32(o_addr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pk = make_uint4(pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                              pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                              pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                              pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
        const int chunk = c * 4 + (i >> 3);
        *reinterpret_cast<uint4*>(otile + r_in * 256 + ((chunk ^ (r_in & 7)) << 4)) = pk;
      }
    }
    __syncwarp();
    const int sub = lane >> 4, ch = lane & 15;   // 2 rows per instruction, 16 chunks of 16 B per row
#pragma unroll 4
    for (int i = 0; i < 16; ++i) {
      const int rr = q * 32 + 2 * i + sub;        // row inside the tile (this warp's quarter)
      const int qi = (q_pair * 2 + wg) * TQ + rr;
      if (qi >= p.Nq) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(otile + rr * 256 + ((ch ^ (rr & 7)) << 4));
      bf16* orow = p.O + (static_cast<int64_t>(b) * p.Nq + qi) * p.ldo + h * HD;
      if (p.o_rows_per_block > 0)
        orow = reinterpret_cast<bf16*>(p.o_blocks.p[qi / p.o_rows_per_block]) +
               static_cast<int64_t>(qi % p.o_rows_per_block) * p.ldo + h * HD;
      *reinterpret_cast<uint4*>(orow + ch * 8) = v;
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
                      cudaStream_t stream, const PeerTable* o_blocks, int rows_per_block) {
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
  p.o_rows_per_block = 0;
  p.o_blocks = PeerTable{};
  if (o_blocks) {
    LTX_CHECK(B == 1 && rows_per_block > 0, 2, "attention: peer-memory output needs B == 1");
    p.o_rows_per_block = rows_per_block;
    p.o_blocks = *o_blocks;
  }
  dim3 grid((Nq + 2 * TQ - 1) / (2 * TQ), H, B);
  launch_pdl(PDL_ATTN, attention_fwd_tcgen05, grid, dim3(ATT_THREADS), ATT_SMEM, stream, tmQ, tmK, tmV, p);
  LTX_CUDA(cudaGetLastError());
}

}  // namespace ltx
