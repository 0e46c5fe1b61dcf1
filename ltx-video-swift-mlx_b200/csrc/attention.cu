// Flash attention forward for sm_100a (head_dim 128 or 64, bf16 in/out, fp32 softmax), used for the video-token
// self-attention and the text cross-attention of the LTX-2 DiT block (head_dim 128) and for the audio / cross-modal
// attentions of the dual audio-video block (32 heads x 64, Models/Transformer/LTX2TransformerBlock.swift:124-139).
//
// Replaces MLXFast.scaledDotProductAttention (T/LTXAttention.swift:192-211):
//   O = softmax(Q K^T * scale + key_bias) V,   key_bias = (1 - mask) * -10000 (T/LTXTransformer.swift:141-156) or none.
//
// One CTA per (128-query tile, head, batch), 192 threads, TWO CTAs per SM (96 KB smem, 256 TMEM columns each): the second
// CTA's tensor work fills the gaps of the first, and 128-row work units quantise better over 148 SMs than pairs
// (N = 1536: 384 CTAs on 296 slots).  Keys are consumed 64 at a time so that the logits can be double-buffered in TMEM:
//   warp 0    : TMA producer -- the Q tile once, then K / V^T tiles (64 keys), two buffers each, fetched independently
//   warp 1    : MMA issuer   -- S(j) = Q K(j)^T (tcgen05.mma 128x64x16, operands from smem) into S buffer j % 2 and
//               O += P(j) V(j) with the A operand P read from TENSOR MEMORY (the probabilities never touch shared memory).
//               Issue order S(0) S(1) | PV(0) S(2) | PV(1) S(3) | ...: S(j+1) is computed while the softmax warps work on
//               S(j), so they never wait for the tensor core in steady state -- measured on the previous design (pairs of
//               query tiles in ping-pong, one S buffer per tile) the softmax warps idled 45 % of the time on that round trip.
//   warps 2-5 : softmax -- thread = query row: the 64 logits of the row come into registers with two tcgen05.ld in flight;
//               running max / sum in fp32 (exp2 domain, ex2.approx); the output accumulator in TMEM is rescaled only when
//               the row max grows by more than 2^8 (lazy rescale, exact after the final 1/l normalisation); P is packed to
//               bf16x2 and stored over the first 32 columns of the S buffer it came from.
// PAIR = true (Nq >= 3072, see the dispatch at the bottom): two CTAs of a cluster -- adjacent 128-query tiles of one head, on the two SMs of a TPC -- share every
// K / V^T tile: each CTA loads HALF of it (32 of the 64 keys of K, 64 of the 128 feature rows of V^T) and the leader issues
// tcgen05.mma.cta_group::2 (M = 256: rows 0-127 accumulate in the leader's tensor memory, 128-255 in the peer's).  Why: the
// one-CTA form is bound by shared-memory traffic (profiles/r02_attention_experiments.txt) -- per 64-key step an SM reads
// 8 x (4 KB Q + 2 KB K) + 4 x 4 KB V^T and takes 32 KB of TMA writes; in pair mode the B halves and the TMA writes halve: 96 -> 64 KB.
// Protocol as in gemm2.cu: 2-SM TMA loads report to the LEADER's full barriers (armed for both CTAs' bytes, the peer adds a remote
// arrive), tcgen05.commit multicasts to both CTAs (S landed, K / V stage free, PV retired), each CTA's softmax warps work on their
// own 128 rows and arrive -- one elected lane per warp -- on the leader's p_full.
//
// tcgen05.commit tracks every earlier MMA of the issuing thread: "S(j+2) landed" implies "PV(j) retired", i.e. the buffer's
// previous P has been consumed.  TMEM: S|P buffers [0,64) [64,128), O [128, 128 + HD).
// V is consumed as V^T [head_dim, keys] (K-major for the PV product); the V-projection GEMM writes it in that layout.
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int ATT_THREADS = 192;
constexpr int TQ = 128, TK = 64;
constexpr uint32_t ATT_TMEM_COLS = 256;
constexpr float RESCALE_THRESHOLD = 8.0f;  // log2 units

template <int HD, bool PAIR>
struct AttCfg {
  static constexpr int KROWS = PAIR ? TK / 2 : TK;       // keys of a K tile held by one CTA
  static constexpr int VROWS = PAIR ? HD / 2 : HD;       // feature rows of a V^T tile held by one CTA
  static constexpr uint32_t Q_TILE = 128 * HD * 2;       // [128 rows x HD]: HD/64 swizzled halves of 16 KB
  static constexpr uint32_t Q_HALF = 128 * 64 * 2;
  static constexpr uint32_t K_TILE = KROWS * HD * 2;     // [keys x HD]: HD/64 halves
  static constexpr uint32_t K_HALF = KROWS * 64 * 2;
  static constexpr uint32_t V_TILE = VROWS * TK * 2;     // [feature rows x 64 keys]
  static constexpr size_t SMEM = 1024 + Q_TILE + 2 * K_TILE + 2 * V_TILE + 16 * 8 + 16;
};

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

template <int HD, bool PAIR>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attention_fwd_tcgen05(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Cfg = AttCfg<HD, PAIR>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + Cfg::Q_TILE;        // 2 buffers
  uint8_t* sV = sK + 2 * Cfg::K_TILE;    // 2 buffers
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * Cfg::V_TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2] per S buffer
  uint64_t* p_full = bars + 11;   // [2] per S buffer
  // one tcgen05.commit per MMA group does all the signalling: s_full[st] = 'S(j) landed' = 'the K stage is free' (producer) =
  // 'PV(j-2) retired'; v_empty[st] = 'PV(j) retired' = 'the V stage is free' (producer) = 'O is stable up to step j' (rescale, epilogue).
  // No waiter can fall two phases behind on either (each phase needs a tile / probabilities the waiter itself has to supply first).
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int n_kv = (p.Nk + TK - 1) / TK;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;   // pair mode: CTAs 2i, 2i+1 of the x dimension = adjacent query tiles
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    // pair mode: the full barriers that count are the leader's (its expect_tx arrive + the peer's remote arrive); p_full takes one
    // arrive per softmax warp of both CTAs; everything a tcgen05.commit signals is multicast to both CTAs' copies
    mbar_init(q_full, PAIR ? 2 : 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], PAIR ? 2 : 1);
      mbar_init(&v_full[i], PAIR ? 2 : 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], PAIR ? 8 : 4);   // one arrive per softmax warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) tmem_alloc_2cta<ATT_TMEM_COLS>(tmem_slot);
    else tmem_alloc<ATT_TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (PAIR && !leader && threadIdx.x == 0) {
    // the leader's MMAs address both CTAs' tensor memory with ONE address: the paired allocation must have landed on the same
    // columns in both (it does -- cta_group::2 allocations are symmetric; this check makes a violation loud instead of wrong)
    uint32_t remote_addr, theirs;
    asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(remote_addr) : "r"(smem_u32(tmem_slot)));
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(theirs) : "r"(remote_addr) : "memory");
    if (theirs != tmem_base) {
      printf("ltxcuda: attention pair got asymmetric tensor-memory columns (%u vs %u)\n", theirs, tmem_base);
      __trap();
    }
  }
  griddep_launch();
  griddep_wait();   // PDL: the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      // pair mode: the leader arms the barrier for both CTAs' bytes, the peer adds its arrive; 2-SM loads report to the leader
      auto arm = [&](uint64_t* bar, uint32_t bytes) {
        if (!PAIR) mbar_arrive_expect_tx(bar, bytes);
        else if (leader) mbar_arrive_expect_tx(bar, 2 * bytes);
        else mbar_arrive_remote(bar, 0);
      };
      auto load2 = [&](void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
        if (PAIR) tma_load_2d_2sm(dst, m, bar, c0, c1);
        else tma_load_2d(dst, m, bar, c0, c1);
      };
      arm(q_full, Cfg::Q_TILE);
#pragma unroll
      for (int hh = 0; hh < HD / 64; ++hh)
        load2(sQ + hh * Cfg::Q_HALF, &tmQ, q_full, h * HD + hh * 64, b * p.Nq + q_tile * TQ);
      // K and V^T tiles are fetched independently, each as soon as one of its two buffers is free
      int nk = 0, nv = 0;
      uint32_t spins = 0;
      while (nk < n_kv || nv < n_kv) {
        bool progress = false;
        if (nk < n_kv && mbar_test(&s_full[nk & 1], ((nk >> 1) & 1) ^ 1)) {   // the stage is free once the S MMAs that read it have landed
          uint8_t* dst = sK + (nk & 1) * Cfg::K_TILE;
          arm(&k_full[nk & 1], Cfg::K_TILE);
#pragma unroll
          for (int hh = 0; hh < HD / 64; ++hh)   // pair mode: this CTA's half of the tile's keys
            load2(dst + hh * Cfg::K_HALF, &tmK, &k_full[nk & 1], h * HD + hh * 64, b * p.Nk + nk * TK + static_cast<int>(rank) * Cfg::KROWS);
          ++nk;
          progress = true;
        }
        if (nv < n_kv && mbar_test(&v_empty[nv & 1], ((nv >> 1) & 1) ^ 1)) {
          // V^T is a 3-D map (keys, features, batch): keys past Nk are zero-filled, never another batch's columns
          arm(&v_full[nv & 1], Cfg::V_TILE);
          if (PAIR)   // this CTA's half of the feature rows
            tma_load_3d_2sm(sV + (nv & 1) * Cfg::V_TILE, &tmV, &v_full[nv & 1], nv * TK, h * HD + static_cast<int>(rank) * Cfg::VROWS, b);
          else
            tma_load_3d(sV + (nv & 1) * Cfg::V_TILE, &tmV, &v_full[nv & 1], nv * TK, h * HD, b);
          ++nv;
          progress = true;
        }
        if (progress) {
          spins = 0;
        } else if (__nanosleep(128), ++spins > (1u << 23)) {   // back off: a spinning lane steals issue slots from the softmax warps
          printf("ltxcuda: attention producer stuck block(%d,%d,%d) k %d v %d\n", blockIdx.x, blockIdx.y, blockIdx.z, nk, nv);
          __trap();
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(PAIR ? 256 : 128, TK);
      constexpr uint32_t idesc_o = umma_idesc_bf16(PAIR ? 256 : 128, HD);
      auto commit = [&](uint64_t* bar) {
        if (PAIR) umma_commit_2cta(bar);   // both CTAs' copies
        else umma_commit(bar);
      };
      // every operand descriptor of the loop is a constant of the CTA: build them once, so that the issuing thread's
      // instruction stream per MMA is the tcgen05.mma itself (it is one thread; its issue rate is part of the step)
      const uint32_t q_addr = smem_u32(sQ);
      uint64_t qd[HD / 16], kd[2][HD / 16], vd[2][TK / 16];
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        qd[kk] = umma_desc_sw128(q_addr + (kk >> 2) * Cfg::Q_HALF + (kk & 3) * 32);
#pragma unroll
        for (int st = 0; st < 2; ++st)
          kd[st][kk] = umma_desc_sw128(smem_u32(sK + st * Cfg::K_TILE) + (kk >> 2) * Cfg::K_HALF + (kk & 3) * 32);
      }
#pragma unroll
      for (int kk = 0; kk < TK / 16; ++kk)
#pragma unroll
        for (int st = 0; st < 2; ++st) vd[st][kk] = umma_desc_sw128(smem_u32(sV + st * Cfg::V_TILE) + kk * 32);
      auto issue_S = [&](int j, auto st_tag) {
        constexpr int st = decltype(st_tag)::value;
        mbar_wait(&k_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          if (PAIR) umma_bf16_2cta(tmem_base + st * TK, qd[kk], kd[st][kk], idesc_s, kk != 0);
          else umma_bf16(tmem_base + st * TK, qd[kk], kd[st][kk], idesc_s, kk != 0);
        }
        commit(&s_full[st]);    // S(j) landed; also: PV(j-2) retired (the buffer's previous P has been consumed) and the K stage is free
      };
      auto issue_PV = [&](int j, auto st_tag) {
        constexpr int st = decltype(st_tag)::value;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(&p_full[st], (j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < TK / 16; ++kk) {  // A = P(j) from tensor memory: 16 keys = 8 packed columns per k-step
          if (PAIR) umma_bf16_ts_2cta(tmem_base + 2 * TK, tmem_base + st * TK + kk * 8, vd[st][kk], idesc_o, (j | kk) != 0);
          else umma_bf16_ts(tmem_base + 2 * TK, tmem_base + st * TK + kk * 8, vd[st][kk], idesc_o, (j | kk) != 0);
        }
        commit(&v_empty[st]);   // PV(j) retired: the V stage is free, O is stable up to step j (the softmax warps' rescale and the epilogue wait on it)
      };
      using S0 = std::integral_constant<int, 0>;
      using S1 = std::integral_constant<int, 1>;
      mbar_wait(q_full, 0);
      issue_S(0, S0{});
      if (n_kv > 1) issue_S(1, S1{});
      for (int j = 0; j < n_kv; j += 2) {   // two steps per trip: the stage index is a compile-time constant
        issue_PV(j, S0{});
        if (j + 2 < n_kv) issue_S(j + 2, S0{});   // into the S buffer PV(j) has just read P from (in-order tensor pipe)
        if (j + 1 < n_kv) {
          issue_PV(j + 1, S1{});
          if (j + 3 < n_kv) issue_S(j + 3, S1{});
        }
      }
    }
  } else {
    const int q = warp & 3;           // TMEM lane quarter
    const int r_in = q * 32 + lane;   // query row inside the tile == TMEM lane
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const uint32_t o_addr = tmem_base + lane_addr + 2 * TK;
    const float LOG2E = 1.4426950408889634f;
    float m_ref = -1.0e30f, l_run = 0.f;   // m_ref in raw-logit units (before the softmax scale)
    // one key tile; MASKED = the tile carries an additive key bias and / or is the ragged last tile (kept out of the main
    // loop: the per-element selects cost as much issue bandwidth as the exponentials)
    auto step = [&](int j, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int st = j & 1;
      const int kv0 = j * TK;
      const int valid = min(TK, p.Nk - kv0);
      const uint32_t s_addr = tmem_base + lane_addr + st * TK;   // S buffer; P (bf16x2) overwrites its first 32 columns
      mbar_wait(&s_full[st], (j >> 1) & 1);
      tc_fence_after();
      uint32_t sr[2][32];
      tmem_ld32(s_addr, sr[0]);   // both loads in flight, one wait
      tmem_ld32(s_addr + 32, sr[1]);
      tmem_ld_wait();
      float s[TK];
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int i = 0; i < 32; ++i) s[c * 32 + i] = __uint_as_float(sr[c][i]);
      if (MASKED && p.key_bias != nullptr) {
        // the additive bias is defined on the scaled logits: fold it in as bias / scale so one FFMA applies both later
        const float inv_scale = LOG2E / p.scale_log2;
        const float* kb = p.key_bias + static_cast<int64_t>(b) * p.Nk + kv0;
        if (valid == TK && (reinterpret_cast<uintptr_t>(kb) & 15) == 0) {
#pragma unroll
          for (int i = 0; i < TK; i += 4) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(kb + i));
            s[i] = fmaf(bv.x, inv_scale, s[i]);
            s[i + 1] = fmaf(bv.y, inv_scale, s[i + 1]);
            s[i + 2] = fmaf(bv.z, inv_scale, s[i + 2]);
            s[i + 3] = fmaf(bv.w, inv_scale, s[i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < TK; ++i)
            if (i < valid) s[i] = fmaf(__ldg(kb + i), inv_scale, s[i]);
        }
      }
      if (MASKED && valid < TK) {
#pragma unroll
        for (int i = 0; i < TK; ++i)
          if (i >= valid) s[i] = -INFINITY;
      }
      float mx8[8];   // eight independent max chains
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = s[i];
#pragma unroll
      for (int i = 8; i < TK; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], s[i]);
      const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      const float m_new = fmaxf(m_ref, mx);
      const bool need = (m_new - m_ref) * p.scale_log2 > RESCALE_THRESHOLD;
      if (__any_sync(0xffffffffu, need)) {
        const float alpha = need ? ex2_approx((m_ref - m_new) * p.scale_log2) : 1.0f;
        if (j > 0) {
          mbar_wait(&v_empty[(j - 1) & 1], ((j - 1) >> 1) & 1);   // O must be stable: PV(j-1) retired (PV(j) cannot start before P(j) below)
          tc_fence_after();
          uint32_t r[32];
#pragma unroll 1
          for (int c = 0; c < HD / 32; ++c) {
            tmem_ld32(o_addr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
            tmem_st32(o_addr + c * 32, r);
          }
        }
        l_run *= alpha;
        if (need) m_ref = m_new;
      }
      const float neg = -m_ref * p.scale_log2;
      float rs0 = 0.f, rs1 = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const float p0 = ex2_approx(fmaf(s[2 * t], p.scale_log2, neg));
        const float p1 = ex2_approx(fmaf(s[2 * t + 1], p.scale_log2, neg));
        rs0 += p0;
        rs1 += p1;
        pk[t] = pack_bf16(p0, p1);
      }
      tmem_st32(s_addr, pk);
      l_run += rs0 + rs1;
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();   // every lane's P columns are stored and fenced: one arrive per warp (pair mode: on the LEADER's barrier)
      if (lane == 0) {
        if (PAIR) mbar_arrive_remote(&p_full[st], 0);
        else mbar_arrive(&p_full[st]);
      }
    };
    {
      const bool ragged = (p.Nk % TK) != 0;
      if (p.key_bias != nullptr) {
        for (int j = 0; j < n_kv; ++j) step(j, std::true_type{});
      } else {
        const int n_full = ragged ? n_kv - 1 : n_kv;
        for (int j = 0; j < n_full; ++j) step(j, std::false_type{});
        if (ragged) step(n_kv - 1, std::true_type{});
      }
    }
    // epilogue: O / l -> bf16
    mbar_wait(&v_empty[(n_kv - 1) & 1], ((n_kv - 1) >> 1) & 1);   // the last PV has retired
    tc_fence_after();
    const float inv = 1.0f / l_run;
    // O / l -> bf16, staged through the Q buffer (free now: every MMA has retired) so that the global stores are whole rows
    // (HD * 2 bytes: 16 or 8 lanes x 16 B) -- full lines, also over NVLink when the row lives in a peer's buffer --
    // instead of 32 rows x 16 B per instruction.  Each warp reads back only the 32 rows it wrote.
    // 16-byte chunk c of row r sits at chunk c ^ (r & 7): conflict-free for the row-per-thread writes and the reads.
    constexpr int ROW_BYTES = HD * 2, CPR = HD / 8;   // chunks per row
    uint8_t* otile = sQ;
#pragma unroll 1
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(o_addr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        uint4 pkv = make_uint4(pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv),
                               pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv),
                               pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv),
                               pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv));
        const int chunk = c * 4 + (i >> 3);
        *reinterpret_cast<uint4*>(otile + r_in * ROW_BYTES + ((chunk ^ (r_in & 7)) << 4)) = pkv;
      }
    }
    __syncwarp();
    constexpr int RPI = 32 / CPR;                     // rows per store instruction (2 or 4)
    const int sub = lane / CPR, ch = lane % CPR;
#pragma unroll 4
    for (int i = 0; i < 32 / RPI; ++i) {
      const int rr = q * 32 + RPI * i + sub;        // row inside the tile (this warp's quarter)
      const int qi = q_tile * TQ + rr;
      if (qi >= p.Nq) continue;
      const uint4 v = *reinterpret_cast<const uint4*>(otile + rr * ROW_BYTES + ((ch ^ (rr & 7)) << 4));
      bf16* orow = p.O + (static_cast<int64_t>(b) * p.Nq + qi) * p.ldo + h * HD;
      if (p.o_rows_per_block > 0)
        orow = reinterpret_cast<bf16*>(p.o_blocks.p[qi / p.o_rows_per_block]) +
               static_cast<int64_t>(qi % p.o_rows_per_block) * p.ldo + h * HD;
      *reinterpret_cast<uint4*>(orow + ch * 8) = v;
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta<ATT_TMEM_COLS>(tmem_base);
    else tmem_dealloc<ATT_TMEM_COLS>(tmem_base);
  }
}

template <int HD, bool PAIR>
void attention_launch(const bf16* Q, int64_t ldq, const bf16* K, int64_t ldk, const bf16* Vt, int64_t ldvb, const AttnParams& p,
                      int D, cudaStream_t stream) {
  using Cfg = AttCfg<HD, PAIR>;
  auto kern = attention_fwd_tcgen05<HD, PAIR>;
  ensure_dyn_smem(kern, Cfg::SMEM);
  CUtensorMap tmQ = make_tmap_2d(Q, static_cast<uint64_t>(p.B) * p.Nq, D, ldq, 128);
  CUtensorMap tmK = make_tmap_2d(K, static_cast<uint64_t>(p.B) * p.Nk, D, ldk, Cfg::KROWS);
  CUtensorMap tmV = make_tmap_3d(Vt, p.Nk, D, p.B, static_cast<uint64_t>(p.B) * ldvb, ldvb, TK, Cfg::VROWS);
  int tiles = (p.Nq + TQ - 1) / TQ;
  if (PAIR) tiles = (tiles + 1) & ~1;   // whole pairs: an odd last tile gets a partner whose rows are all past Nq (never stored)
  dim3 grid(tiles, p.H, p.B);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (pdl_enabled(PDL_ATTN)) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (PAIR) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  LTX_CUDA(cudaLaunchKernelEx(&cfg, kern, tmQ, tmK, tmV, p));
}

}  // namespace

void launch_attention(const bf16* Q, int64_t ldq, const bf16* K, int64_t ldk, const bf16* Vt, int64_t ldvb,
                      const float* key_bias, bf16* O, int64_t ldo, int B, int H, int Nq, int Nk, int D, float scale,
                      cudaStream_t stream, const PeerTable* o_blocks, int rows_per_block) {
  LTX_CHECK(D == H * 128 || D == H * 64, 2, "attention: head_dim must be 128 or 64");
  LTX_CHECK(Nq > 0 && Nk > 0 && B > 0, 2, "attention: empty problem");
  LTX_CHECK(ldvb % 8 == 0 && ldvb >= Nk, 2, "attention: V^T per-batch pitch must be a multiple of 8 and >= Nk");
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
  // CTA pairs sharing the K / V tiles on LONG sequences.  Measured after the commits were merged (profiles/r02_attention_experiments.txt
  // (6)): one CTA per tile is 2-5 % faster at N = 1536 (56.2 vs 57.5 us self, 39.4 vs 41.6 cross), the pair 2-4 % faster from
  // N = 6144 up (569 vs 581 us; 12672: 2142 vs 2233 us).  LTX_ATT_PAIR=0 / 1: never / whenever there are two query tiles.
  const char* pair_str = getenv("LTX_ATT_PAIR");   // read per call: the parity tests run every case in both forms
  const int pair_env = pair_str ? atoi(pair_str) : -1;
  const bool pair = pair_env < 0 ? Nq >= 3072 : (pair_env != 0 && Nq > TQ);
  if (D == H * 128) {
    if (pair) attention_launch<128, true>(Q, ldq, K, ldk, Vt, ldvb, p, D, stream);
    else attention_launch<128, false>(Q, ldq, K, ldk, Vt, ldvb, p, D, stream);
  } else {
    if (pair) attention_launch<64, true>(Q, ldq, K, ldk, Vt, ldvb, p, D, stream);
    else attention_launch<64, false>(Q, ldq, K, ldk, Vt, ldvb, p, D, stream);
  }
}

}  // namespace ltx
