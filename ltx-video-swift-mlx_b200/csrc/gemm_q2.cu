// 2-CTA (cta_group::2) variant of the dequant-fused GEMM (gemm_q.cu): a cluster of two CTAs computes a 256 x BN tile.
//
// Why: the 1-CTA kernel is bound by the issue rate of its dequant stage -- every 128-row M tile converts the whole BN x 64
// code block of each k-block again (PRMT + FFMA + pack per pair of codes), ~450 TFLOP/s at the DiT shapes.  In pair mode each
// CTA converts only HALF of the B tile (BN/2 rows) into its own shared memory and tcgen05.mma.cta_group::2 (M = 256) reads
// both halves, so the conversion work per flop halves; the TMA traffic per CTA drops the same way (16 KB of A + BN/2 rows of
// codes) and the freed shared memory buys a 4-stage ring.
//
// Two rings: R (A tile + raw codes, 6 slots, filled by TMA) and D (converted B half, 3 slots): the TMA latency is covered by
// the deep ring R without paying 16 KB of converted operand per slot (a single 3-4 slot ring was latency-bound: 366 TFLOP/s).
// Pipeline per k-block (both CTAs run every role; only the leader, cluster rank 0, issues MMAs):
//   TMA producer   : A tile (this CTA's 128 rows, bf16, swizzled) + raw codes of this CTA's B half -> LOCAL full_raw barrier
//   8 dequant warps: wait full_raw (so this CTA's A has landed too), convert codes -> bf16 into the swizzled K-major B half,
//                    fence.proxy.async, arrive (release.cluster) on the LEADER's full_deq barrier (16 arrivals per stage)
//   MMA issuer     : waits full_deq (acquire.cluster), 4 x tcgen05.mma.cta_group::2, commit multicast -> frees the stage in
//                    both CTAs; accumulator hand-off to the epilogue warps of both CTAs as in gemm2.cu
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int Q2_BM = 128, Q2_BK = 64;
constexpr int Q2_THREADS = 448;      // 14 warps: TMA, MMA, 4 epilogue, 8 dequant
constexpr int Q2_DEQ_WARPS = 8;
constexpr int Q2_STAGES = 6;        // ring R: A tile + raw codes (filled by TMA: deep, it has to cover the L2 / HBM latency)
constexpr int Q2_DSTAGES = 3;       // ring D: dequantised B half (filled by the dequant warps, drained by the MMA)
constexpr int Q2_BN_MAX = 256;
constexpr uint32_t Q2_A_BYTES = Q2_BM * Q2_BK * 2;               // 16 KB
constexpr uint32_t Q2_R_STRIDE = (Q2_BN_MAX / 2) * Q2_BK;        // raw codes of the B half, up to 8 KB
constexpr uint32_t Q2_B_STRIDE = (Q2_BN_MAX / 2) * Q2_BK * 2;    // dequantised bf16 B half, up to 16 KB
constexpr size_t Q2_SMEM = 1024 + Q2_STAGES * (Q2_A_BYTES + Q2_R_STRIDE) + Q2_DSTAGES * Q2_B_STRIDE +
                           (2 * Q2_STAGES + 2 * Q2_DSTAGES + 4) * 8 + 16 + 128 + 4 * EPI_STAGE_BYTES;

__device__ __forceinline__ uint32_t deq2x(uint32_t word, int i0, int i1, float s, float bm) {
  // two codes -> s*q+beta -> bf16x2 (see deq2 in gemm_q.cu: PRMT into the mantissa of 2^15, one FFMA)
  const uint32_t sel0 = 0x7604u + (i0 << 4), sel1 = 0x7604u + (i1 << 4);
  const float f0 = __uint_as_float(__byte_perm(word, 0x47000000u, sel0));
  const float f1 = __uint_as_float(__byte_perm(word, 0x47000000u, sel1));
  return pack_bf16_alu(fmaf(f0, s, bm), fmaf(f1, s, bm));
}

// cluster-scope release arrive on the barrier at the same offset in CTA `cta` (publishes this thread's prior smem writes)
__device__ __forceinline__ void mbar_arrive_remote_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!ok && ++spins > (1u << 26)) {
      printf("ltxcuda: mbarrier (cluster) timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  } while (!ok);
}

template <int MODE, int BITS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Q2_THREADS, 1)
gemm_q_2cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmQ, int M, int N, int K, int BN,
            const float* __restrict__ scales, const float* __restrict__ biases, int a_kblock, const GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = sA + Q2_STAGES * Q2_A_BYTES;
  uint8_t* sR = sB + Q2_DSTAGES * Q2_B_STRIDE;
  uint64_t* full_raw = reinterpret_cast<uint64_t*>(sR + Q2_STAGES * Q2_R_STRIDE);
  uint64_t* empty = full_raw + Q2_STAGES;       // ring R slot consumed (A by the MMA; its codes were converted before that)
  uint64_t* full_deq = empty + Q2_STAGES;       // ring D slot converted (leader's copy counts both CTAs' dequant warps)
  uint64_t* empty_d = full_deq + Q2_DSTAGES;    // ring D slot consumed by the MMA
  uint64_t* tfull = empty_d + Q2_DSTAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~static_cast<uintptr_t>(127));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int num_mp = (M + 2 * Q2_BM - 1) / (2 * Q2_BM);
  const int num_n = (N + BN - 1) / BN;
  const int num_tiles = num_mp * num_n;
  const int num_k = (K + Q2_BK - 1) / Q2_BK;
  const int half_bn = BN >> 1;
  constexpr int ROW_BYTES = (BITS == 8) ? 64 : 32;       // raw bytes per row per k-block
  const uint32_t raw_bytes = static_cast<uint32_t>(half_bn) * ROW_BYTES;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmQ);
    for (int i = 0; i < Q2_STAGES; ++i) {
      mbar_init(&full_raw[i], 1);                   // local TMA producer (A + raw codes of this CTA)
      mbar_init(&empty[i], 1);                      // multicast tcgen05.commit
    }
    for (int i = 0; i < Q2_DSTAGES; ++i) {
      mbar_init(&full_deq[i], 2 * Q2_DEQ_WARPS);    // the dequant warps of BOTH CTAs (only the leader's copy is used)
      mbar_init(&empty_d[i], 1);                    // multicast tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);                     // 4 epilogue warps of each CTA (leader's copy)
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();   // PDL: everything above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int mp = tile % num_mp, n_blk = tile / num_mp;
        const int m0 = (mp * 2 + static_cast<int>(rank)) * Q2_BM;
        const int n0 = n_blk * BN + static_cast<int>(rank) * half_bn;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_raw[stage], Q2_A_BYTES + raw_bytes);
          if (a_kblock > 0)
            tma_load_3d(sA + stage * Q2_A_BYTES, &tmA, &full_raw[stage], (kb * Q2_BK) % a_kblock, m0, (kb * Q2_BK) / a_kblock);
          else
            tma_load_2d(sA + stage * Q2_A_BYTES, &tmA, &full_raw[stage], kb * Q2_BK, m0);
          tma_load_2d(sR + stage * Q2_R_STRIDE, &tmQ, &full_raw[stage], kb * ROW_BYTES, n0);
          if (++stage == Q2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * Q2_BM, BN);
      int stage = 0, ds = 0;
      uint32_t dphase = 0;
      int t = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
        const int as = t & 1;
        const uint32_t aphase = (t >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Q2_BN_MAX;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait_acquire_cluster(&full_deq[ds], dphase);   // both CTAs: A landed (ring R), B half converted (ring D)
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Q2_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + ds * Q2_B_STRIDE);
#pragma unroll
          for (int k = 0; k < Q2_BK / 16; ++k)
            umma_bf16_2cta(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(&empty[stage]);     // A slot (and its raw codes) may be refilled by TMA
          umma_commit_2cta(&empty_d[ds]);      // B slot may be overwritten by the dequant warps
          if (++stage == Q2_STAGES) stage = 0;
          if (++ds == Q2_DSTAGES) { ds = 0; dphase ^= 1; }
        }
        umma_commit_2cta(&tfull[as]);
      }
    }
  } else if (warp < 6) {
    const int q = warp & 3;
    int t = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
      const int mp = tile % num_mp, n_blk = tile / num_mp;
      const int as = t & 1;
      const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * Q2_BN_MAX;
      epilogue_tile<MODE>(taddr, BN, epi_stage + (warp - 2) * (EPI_STAGE_BYTES / 4), lane,
                          (mp * 2 + static_cast<int>(rank)) * Q2_BM + q * 32, n_blk * BN, M, N, ep);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty[as], 0);
    }
  } else {
    // ---- dequant warps: raw codes of this CTA's B half -> bf16 operand tile (K-major, 128 B per row, 128B swizzle)
    const int dt = threadIdx.x - 6 * 32;                 // 0..255
    constexpr int CH_PER_ROW = ROW_BYTES / 16;           // 16-byte raw chunks per row: 4 (int8) / 2 (int4)
    constexpr int MAX_IT = 2;                            // (BN/2) * CH_PER_ROW <= 512 chunks over 256 threads
    const int chunks = half_bn * CH_PER_ROW;
    int stage = 0, ds = 0;
    uint32_t phase = 0, dphase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      const int n0 = (tile / num_mp) * BN + static_cast<int>(rank) * half_bn;
      float sc[MAX_IT], bm[MAX_IT];
      auto fetch = [&](int kb, float (&so)[MAX_IT], float (&bo)[MAX_IT]) {
        const float* srow = scales + static_cast<int64_t>(kb) * N;
        const float* brow = biases + static_cast<int64_t>(kb) * N;
#pragma unroll
        for (int it = 0; it < MAX_IT; ++it) {
          const int c = dt + it * (Q2_DEQ_WARPS * 32);
          const int n = n0 + c / CH_PER_ROW;
          const bool ok = c < chunks && n < N;
          const float sv = ok ? __ldg(srow + n) : 0.f, bv = ok ? __ldg(brow + n) : 0.f;
          so[it] = sv;
          bo[it] = fmaf(-32768.0f, sv, bv);
        }
      };
      fetch(0, sc, bm);
      for (int kb = 0; kb < num_k; ++kb) {
        float sn[MAX_IT], bn[MAX_IT];
        if (kb + 1 < num_k) fetch(kb + 1, sn, bn);
        mbar_wait(&full_raw[stage], phase);
        mbar_wait(&empty_d[ds], dphase ^ 1);
        const uint8_t* raw = sR + stage * Q2_R_STRIDE;
        uint8_t* dst = sB + ds * Q2_B_STRIDE;
        uint4 u[MAX_IT];
#pragma unroll
        for (int it = 0; it < MAX_IT; ++it) {
          const int c = dt + it * (Q2_DEQ_WARPS * 32);
          if (c < chunks) u[it] = *reinterpret_cast<const uint4*>(raw + (c / CH_PER_ROW) * ROW_BYTES + (c % CH_PER_ROW) * 16);
        }
#pragma unroll
        for (int it = 0; it < MAX_IT; ++it) {
          const int c = dt + it * (Q2_DEQ_WARPS * 32);
          if (c >= chunks) continue;
          const int row = c / CH_PER_ROW, part = c % CH_PER_ROW;
          const float s = sc[it], b = bm[it];
          const uint32_t w[4] = {u[it].x, u[it].y, u[it].z, u[it].w};
          uint8_t* drow = dst + row * 128;
          if (BITS == 8) {
            uint32_t o[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              o[2 * i] = deq2x(w[i], 0, 1, s, b);
              o[2 * i + 1] = deq2x(w[i], 2, 3, s, b);
            }
            *reinterpret_cast<uint4*>(drow + (((2 * part) ^ (row & 7)) * 16)) = make_uint4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<uint4*>(drow + (((2 * part + 1) ^ (row & 7)) * 16)) = make_uint4(o[4], o[5], o[6], o[7]);
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const uint32_t lo = w[i] & 0x0F0F0F0Fu, hi = (w[i] >> 4) & 0x0F0F0F0Fu;
              const uint32_t e01 = __byte_perm(lo, hi, 0x5140);
              const uint32_t e23 = __byte_perm(lo, hi, 0x7362);
              const uint4 v = make_uint4(deq2x(e01, 0, 1, s, b), deq2x(e01, 2, 3, s, b), deq2x(e23, 0, 1, s, b), deq2x(e23, 2, 3, s, b));
              *reinterpret_cast<uint4*>(drow + (((4 * part + i) ^ (row & 7)) * 16)) = v;
            }
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(&full_deq[ds]);
          else mbar_arrive_remote_release(&full_deq[ds], 0);
        }
        if (++stage == Q2_STAGES) { stage = 0; phase ^= 1; }
        if (++ds == Q2_DSTAGES) { ds = 0; dphase ^= 1; }
        if (kb + 1 < num_k) {
#pragma unroll
          for (int it = 0; it < MAX_IT; ++it) { sc[it] = sn[it]; bm[it] = bn[it]; }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

template <int MODE, int BITS>
void launch_q2(const CUtensorMap& tmA, const CUtensorMap& tmQ, int M, int N, int K, int BN, const float* s, const float* b,
               int a_kblock, const GemmEpi& epi, cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_q_2cta<MODE, BITS>;
  if (!configured) {
    LTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Q2_SMEM)));
    configured = true;
  }
  const int tiles = ((M + 2 * Q2_BM - 1) / (2 * Q2_BM)) * ((N + BN - 1) / BN);
  const int clusters = device_sm_count() / 2;
  const int grid = 2 * (tiles < clusters ? tiles : clusters);
  launch_pdl(PDL_GEMM, kern, dim3(grid), dim3(Q2_THREADS), Q2_SMEM, stream, tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi);
}

template <int BITS>
void launch_q2_mode(const CUtensorMap& tmA, const CUtensorMap& tmQ, int M, int N, int K, int BN, const float* s, const float* b,
                    int a_kblock, const GemmEpi& epi, cudaStream_t stream) {
  switch (epi.mode) {
    case EPI_BF16: launch_q2<EPI_BF16, BITS>(tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi, stream); break;
    case EPI_GELU_BF16: launch_q2<EPI_GELU_BF16, BITS>(tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi, stream); break;
    case EPI_GATE_RESID: launch_q2<EPI_GATE_RESID, BITS>(tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi, stream); break;
    case EPI_F32: launch_q2<EPI_F32, BITS>(tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi, stream); break;
    case EPI_SILU_BF16: launch_q2<EPI_SILU_BF16, BITS>(tmA, tmQ, M, N, K, BN, s, b, a_kblock, epi, stream); break;
    default: LTX_CHECK(false, 2, "bad GEMM epilogue mode");
  }
}

}  // namespace

void launch_gemm_q_2cta(const bf16* A, int64_t lda, const QuantW& W, int M, int N, int K, const GemmEpi& epi, cudaStream_t stream,
                        int force_bn, int a_kblock, int64_t a_kblock_stride) {
  int bn = force_bn ? force_bn : gemm2_fit_tile_width(M, N);
  LTX_CHECK(bn >= 64 && bn <= 256 && bn % 16 == 0, 2, "2-CTA quantised GEMM: tile width must be a multiple of 16 in [64, 256]");
  const uint64_t row_bytes = W.bits == 8 ? K : K / 2;
  CUtensorMap tmA;
  if (a_kblock > 0) {
    LTX_CHECK(a_kblock % Q2_BK == 0 && K % a_kblock == 0 && lda == a_kblock, 2, "quantised GEMM: bad K-blocked A layout");
    tmA = make_tmap_3d(A, a_kblock, M, K / a_kblock, lda, a_kblock_stride, 64, Q2_BM);
  } else {
    tmA = make_tmap_2d(A, M, K, lda, Q2_BM);
  }
  CUtensorMap tmQ = make_tmap_u8(W.q, N, row_bytes, bn / 2, W.bits == 8 ? 64 : 32);
  if (W.bits == 8)
    launch_q2_mode<8>(tmA, tmQ, M, N, K, bn, W.scales, W.biases, a_kblock, epi, stream);
  else
    launch_q2_mode<4>(tmA, tmQ, M, N, K, bn, W.scales, W.biases, a_kblock, epi, stream);
}

}  // namespace ltx
