// 4-CTA cluster variant of the bf16 GEMM: two CTA pairs (cta_group::2) side by side along N share their A operand.
//
// EXPERIMENT, off by default (LTX_GEMM_4CTA=1 or force_bn >= 2000 selects it).  Hypothesis: the pair kernel (gemm2.cu) is
// bound by L2 -> SM operand traffic, so fetching A once per cluster and multicasting it should help.  A cluster of four
// CTAs computes a 256 x (2 BN) tile; pair p owns columns [(2 j + p) BN, +BN); each CTA fetches only a QUARTER of the
// shared 256 x 64 A block (64 rows, 8 KB) and TMA multicasts it to the CTA of the same row half in the other pair, which
// cuts the L2 reads per pair-tile from 32 + BN/8 KB to 16 + BN/8 KB per k-block.
// RESULT on B200 (profiles/r01_gemm_4cta_experiment.txt): bit-identical output, but the SAME time per k-block as the pair
// kernel at every width (e.g. M=1536 N=16384 K=4096: 154.4 vs 155.6 us), i.e. the limiter is the bytes ARRIVING in each
// SM's shared memory (~47 B/clk/SM, unchanged by multicast), not the L2 read traffic; and only 33 clusters of four fit on
// the 148 SMs (GPC boundaries), so large problems lose 6 % (8192^3: 1458 vs 1550 TFLOP/s).  Kept as the reference
// implementation of the 2-SM multicast protocol (used by nothing else).
//
// Cluster ranks: r = 2 p + h, p = pair (N half), h = row half (h = 0 is the pair's MMA leader).  Protocol on top of gemm2.cu:
//   * A quarter loads: cta_group::2 + multicast::cluster, mask {h, h + 2}; the completion bytes of every copy are reported
//     to the pair leader of the RECEIVING CTA (mbarrier address with the peer bit cleared), so each leader's full barrier
//     still expects everything that lands in its own pair: 2 x (16 KB + BN/2 x 128 B)
//   * a stage may be refilled only when BOTH pairs have consumed it (the refill writes into the other pair's smem):
//     empty barriers count 2 and every leader's tcgen05.commit is multicast to all four CTAs
//   * accumulator hand-off (tfull / tempty) stays inside each pair
#include <cstdlib>

#include "gemm_epilogue.cuh"
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int BM4 = 128;          // rows per CTA (256 per pair and per cluster)
constexpr int BK4 = 64;
constexpr int G4_THREADS = 192;
constexpr int G4_STAGES = 6;
constexpr int G4_BN_MAX = 256;
constexpr uint32_t G4_A_BYTES = BM4 * BK4 * 2;                 // 16 KB: this CTA's 128 rows (two multicast quarters)
constexpr uint32_t G4_AQ_BYTES = G4_A_BYTES / 2;               // 8 KB: the quarter this CTA fetches
constexpr uint32_t G4_B_STRIDE = (G4_BN_MAX / 2) * BK4 * 2;    // 16 KB: half of the pair's B tile
constexpr size_t G4_SMEM = 1024 + G4_STAGES * (G4_A_BYTES + G4_B_STRIDE) + (2 * G4_STAGES + 4) * 8 + 16 + 128 + 4 * EPI_STAGE_BYTES;

// 2-SM multicast TMA loads: the box lands at the same smem offset in every CTA of `mask`; completion bytes go to the pair
// leader of each receiving CTA
__device__ __forceinline__ void tma_load_2d_2sm_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, uint16_t mask) {
  const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}

template <int MODE>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(G4_THREADS, 1)
gemm_bf16_4cta(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, int K, int BN,
               int a_kblock, const GemmEpi ep) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + G4_STAGES * G4_A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + G4_STAGES * G4_B_STRIDE);
  uint64_t* empty = full + G4_STAGES;
  uint64_t* tfull = empty + G4_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* epi_stage = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 4) + 127) & ~static_cast<uintptr_t>(127));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = static_cast<int>(rank >> 1), half = static_cast<int>(rank & 1);
  const bool leader = half == 0;
  const uint32_t leader_rank = rank & ~1u;
  const int cluster_id = blockIdx.x >> 2, num_clusters = gridDim.x >> 2;
  const int num_mp = (M + 2 * BM4 - 1) / (2 * BM4);   // 256-row tiles
  const int num_n = (N + BN - 1) / BN;
  const int num_n2 = (num_n + 1) >> 1;                // column pairs: one per cluster tile
  const int num_tiles = num_mp * num_n2;
  const int num_k = (K + BK4 - 1) / BK4;
  const int half_bn = BN >> 1;
  const uint32_t b_bytes = static_cast<uint32_t>(half_bn) * BK4 * 2;
  const uint16_t a_mask = static_cast<uint16_t>((1u << half) | (1u << (half + 2)));   // same row half in both pairs
  const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * pair));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < G4_STAGES; ++i) {
      mbar_init(&full[i], 2);    // leader's expect_tx arrive + its peer's remote arrive (only the leaders' copies are used)
      mbar_init(&empty[i], 2);   // one multicast tcgen05.commit from each pair leader
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);   // pair-multicast tcgen05.commit
      mbar_init(&tempty[i], 8);  // 4 epilogue warps of each CTA of the pair (only the leaders' copies are used)
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; global memory is touched only from here on
  griddep_launch();
  griddep_wait();

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        const int mp = tile % num_mp, n_blk = 2 * (tile / num_mp) + pair;
        const int m0 = (mp * 2 + half) * BM4 + pair * (BM4 / 2);   // the 64-row quarter this CTA fetches for both pairs
        const int n0 = n_blk * BN + half * half_bn;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (G4_A_BYTES + b_bytes));
          else mbar_arrive_remote(&full[stage], leader_rank);
          uint8_t* a_dst = sA + stage * G4_A_BYTES + pair * G4_AQ_BYTES;
          if (a_kblock > 0)
            tma_load_3d_2sm_mc(a_dst, &tmA, &full[stage], (kb * BK4) % a_kblock, m0, (kb * BK4) / a_kblock, a_mask);
          else
            tma_load_2d_2sm_mc(a_dst, &tmA, &full[stage], kb * BK4, m0, a_mask);
          tma_load_2d_2sm(sB + stage * G4_B_STRIDE, &tmB, &full[stage], kb * BK4, n0);
          if (++stage == G4_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(2 * BM4, BN);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
        const int as = t & 1;
        const uint32_t aphase = (t >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * G4_BN_MAX;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * G4_A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * G4_B_STRIDE);
#pragma unroll
          for (int k = 0; k < BK4 / 16; ++k)
            umma_bf16_2cta(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                           (kb | k) != 0 ? 1u : 0u);
          umma_commit_2cta(&empty[stage], 0xF);   // frees the stage in all four CTAs (count 2: both pairs must be done)
          if (++stage == G4_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit_2cta(&tfull[as], pair_mask);
      }
    }
  } else {
    const int q = warp & 3;
    int t = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++t) {
      const int mp = tile % num_mp, n_blk = 2 * (tile / num_mp) + pair;
      const int as = t & 1;
      const uint32_t aphase = (t >> 1) & 1;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * G4_BN_MAX;
      epilogue_tile<MODE>(taddr, BN, epi_stage + (warp - 2) * (EPI_STAGE_BYTES / 4), lane, (mp * 2 + half) * BM4 + q * 32,
                          n_blk * BN, M, N, ep);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&tempty[as], leader_rank);
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta<512>(tmem_base);
  }
}

int g_clusters4 = -1;

template <int MODE>
void launch4(const CUtensorMap& tmA, const CUtensorMap& tmB, int M, int N, int K, int BN, int a_kblock, const GemmEpi& epi,
             cudaStream_t stream) {
  static bool configured = false;
  auto kern = gemm_bf16_4cta<MODE>;
  if (!configured) {
    LTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G4_SMEM)));
    configured = true;
  }
  const int tiles = ((M + 2 * BM4 - 1) / (2 * BM4)) * ((((N + BN - 1) / BN) + 1) / 2);
  const int clusters = gemm4_max_clusters();
  const int grid = 4 * (tiles < clusters ? tiles : clusters);
  static const int dbg = [] { const char* e = getenv("LTX_GEMM_DEBUG"); return e ? atoi(e) : 0; }();
  GemmEpi ep2 = epi;
  ep2.debug = dbg;
  launch_pdl(PDL_GEMM, kern, dim3(grid), dim3(G4_THREADS), G4_SMEM, stream, tmA, tmB, M, N, K, BN, a_kblock, ep2);
}

}  // namespace

// how many 4-CTA clusters of this kernel the device keeps resident (33 on a 148-SM B200: GPC boundaries)
int gemm4_max_clusters() {
  if (g_clusters4 > 0) return g_clusters4;
  auto kern = gemm_bf16_4cta<EPI_BF16>;
  LTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(G4_SMEM)));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(4 * 64);
  cfg.blockDim = dim3(G4_THREADS);
  cfg.dynamicSmemBytes = G4_SMEM;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  int n = 0;
  LTX_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
  g_clusters4 = n > 0 ? n : 1;
  return g_clusters4;
}

// tile width for the 4-CTA kernel: whole waves of resident clusters over 256 x (2 bn) cluster tiles.  Per pair-tile and
// k-block the operand traffic is 16 KB of A + 128 B x bn of B against 2 x bn tensor-pipe cycles: cost ~ max(2 bn, 172 + 1.34 bn)
int gemm4_fit_tile_width(int M, int N) {
  const int clusters = gemm4_max_clusters();
  const int num_mp = (M + 255) / 256;
  int best = 256;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 64; bn -= 16) {
    const long long num_n = (N + bn - 1) / bn;
    const long long tiles = static_cast<long long>(num_mp) * ((num_n + 1) / 2);
    const long long waves = (tiles + clusters - 1) / clusters;
    const double per_tile = 2.0 * bn > 172.0 + 1.34 * bn ? 2.0 * bn : 172.0 + 1.34 * bn;
    const double cost = static_cast<double>(waves) * (per_tile + 24.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

void launch_gemm_4cta(const bf16* A, int64_t lda, const bf16* B, int64_t ldb, int M, int N, int K, const GemmEpi& epi,
                      cudaStream_t stream, int force_bn, int a_kblock, int64_t a_kblock_stride) {
  int bn = force_bn ? force_bn : gemm4_fit_tile_width(M, N);
  LTX_CHECK(bn >= 64 && bn <= 256 && bn % 16 == 0, 2, "4-CTA GEMM: tile width must be a multiple of 16 in [64, 256]");
  CUtensorMap tmA;
  if (a_kblock > 0) {
    LTX_CHECK(a_kblock % BK4 == 0 && K % a_kblock == 0 && lda == a_kblock, 2, "GEMM: bad K-blocked A layout");
    tmA = make_tmap_3d(A, a_kblock, M, K / a_kblock, lda, a_kblock_stride, 64, BM4 / 2);
  } else {
    tmA = make_tmap_2d(A, M, K, lda, BM4 / 2);
  }
  CUtensorMap tmB = make_tmap_2d(B, N, K, ldb, bn / 2);
  switch (epi.mode) {
    case EPI_BF16: launch4<EPI_BF16>(tmA, tmB, M, N, K, bn, a_kblock, epi, stream); break;
    case EPI_GELU_BF16: launch4<EPI_GELU_BF16>(tmA, tmB, M, N, K, bn, a_kblock, epi, stream); break;
    case EPI_GATE_RESID: launch4<EPI_GATE_RESID>(tmA, tmB, M, N, K, bn, a_kblock, epi, stream); break;
    case EPI_F32: launch4<EPI_F32>(tmA, tmB, M, N, K, bn, a_kblock, epi, stream); break;
    case EPI_SILU_BF16: launch4<EPI_SILU_BF16>(tmA, tmB, M, N, K, bn, a_kblock, epi, stream); break;
    default: LTX_CHECK(false, 2, "bad GEMM epilogue mode");
  }
}

}  // namespace ltx
