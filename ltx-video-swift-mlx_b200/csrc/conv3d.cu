// 3x3x3 convolution of the video-VAE decoder as an implicit GEMM on tcgen05 (sm_100a), channels-last.
//
// Replaces Conv3dFull.callAsFunction (Models/VAE/VideoConvolution.swift:238-347: three MLX.conv2d calls over shifted
// temporal slices) and the ops fused around it in the decoder (Models/VAE/VideoDecoder.swift): pixel-norm +
// scale/shift + SiLU (:118-126, 419-436) as a padding prologue kernel, depth-to-space + first-frame trim + tiled
// residual (:201-251) and the final unpatchify + (x+1)/2 clip (:257-275, 501-505) as epilogue stores.
//
// Layout: activations fp32 [T, H, W, C]; the prologue writes a bf16 copy with the conv's own padding materialised
// ([T+2, H+2, W+2, C]: reflect in H/W, frame replication in T), so the implicit GEMM needs no boundary logic:
//   M = output voxels, tiled as TMA boxes bt x bh x bw = 128 voxels;  N = Cout;  K = 27 taps x Cin.
//   A tile (tap, k-chunk) = 4-D TMA box at (t0+dt, h0+dh, w0+dw, k0) of the padded volume -> [128 rows x 64 ch], 128B swizzle
//   B tile               = rows [tap*Cout + n0, +BN) x cols [k0, +64) of the [27*Cout, Cin] weight matrix.
// Warp roles and pipelines are those of gemm.cu (TMA producer / MMA issuer / 4 epilogue warps, 2 TMEM stages).
#include "ltx_internal.h"
#include "ptx.cuh"

namespace ltx {

namespace {

constexpr int CBM = 128, CBK = 64, CONV_THREADS = 192;

template <int BN>
struct ConvCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr uint32_t A_BYTES = CBM * CBK * 2;
  static constexpr uint32_t B_BYTES = BN * CBK * 2;
  static constexpr uint32_t TMEM_COLS = 2 * BN;
  static constexpr size_t SMEM = 1024 + STAGES * (A_BYTES + B_BYTES) + (2 * STAGES + 4) * 8 + 16;
};

struct ConvGeom {
  int T, H, W, Cin, Cout;
  int bt, bh, bw;     // voxel box of one M tile (bt*bh*bw == 128)
  int nt, nh, nw;     // tiles per axis
};

__device__ __forceinline__ float clip01(float v) { return fminf(fmaxf(v, 0.f), 1.f); }

template <int MODE>
__device__ __forceinline__ void conv_epilogue_chunk(const uint32_t (&r)[32], int t, int h, int w, int col0,
                                                    const ConvGeom& g, const ConvEpi& ep) {
  if (col0 >= g.Cout) return;
  const int64_t vox = (static_cast<int64_t>(t) * g.H + h) * g.W + w;
  if (MODE == 0) {
    float* o = ep.out + vox * g.Cout + col0;
    const float* rs = ep.resid ? ep.resid + vox * g.Cout + col0 : nullptr;
    if (col0 + 32 <= g.Cout) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
        float4 v = make_float4(__uint_as_float(r[j]) + b.x, __uint_as_float(r[j + 1]) + b.y,
                               __uint_as_float(r[j + 2]) + b.z, __uint_as_float(r[j + 3]) + b.w);
        if (rs) {
          float4 x = *reinterpret_cast<const float4*>(rs + j);
          v.x += x.x; v.y += x.y; v.z += x.z; v.w += x.w;
        }
        *reinterpret_cast<float4*>(o + j) = v;
      }
    } else {
      for (int j = 0; j < 32 && col0 + j < g.Cout; ++j)
        o[j] = __uint_as_float(r[j]) + ep.bias[col0 + j] + (rs ? rs[j] : 0.f);
    }
  } else if (MODE == 1) {
    // depth-to-space (2,2,2): conv channel co = c*8 + p1*4 + p2*2 + p3 -> out[2t+p1-1, 2h+p2, 2w+p3, c]
    // residual: x[t,h,w,(c mod Cin/8)*8 + p]   (VideoDecoder.swift:201-251)
    const int Cf = g.Cout >> 3;        // output channels
    const int Cr = g.Cin >> 3;         // residual d2s channels (tiled x4)
    const int Ho = 2 * g.H, Wo = 2 * g.W;
    const float* xr = ep.resid + vox * g.Cin;
    const int c0 = col0 >> 3;          // first of 4 output channels covered by these 32 columns
#pragma unroll
    for (int p = 0; p < 8; ++p) {
      const int p1 = p >> 2, p2 = (p >> 1) & 1, p3 = p & 1;
      const int to = 2 * t + p1 - 1 + ep.t_shift;
      if (to < 0) continue;
      float v[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int co = col0 + cc * 8 + p;
        const int c = c0 + cc;
        v[cc] = __uint_as_float(r[cc * 8 + p]) + ep.bias[co] + xr[(c % Cr) * 8 + p];
      }
      float* o = ep.out + ((static_cast<int64_t>(to) * Ho + (2 * h + p2)) * Wo + (2 * w + p3)) * Cf + c0;
      *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
    }
  } else {
    // unpatchify 4x4: conv channel co = c*16 + pa*4 + pb -> frames[t, 4h + pb, 4w + pa, c], (x+1)/2 clipped to [0,1]
    const int Ho = 4 * g.H, Wo = 4 * g.W;
    for (int j = 0; j < 32 && col0 + j < g.Cout; ++j) {
      const int co = col0 + j;
      const int c = co >> 4, pa = (co >> 2) & 3, pb = co & 3;
      const float v = (__uint_as_float(r[j]) + ep.bias[co] + 1.0f) * 0.5f;
      ep.out[((static_cast<int64_t>(t) * Ho + (4 * h + pb)) * Wo + (4 * w + pa)) * 3 + c] = clip01(v);
    }
  }
}

template <int BN, int MODE>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv3d_tcgen05(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const ConvGeom g,
               const ConvEpi ep) {
  using Cfg = ConvCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::B_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_m = g.nt * g.nh * g.nw;
  const int num_n = (g.Cout + BN - 1) / BN;
  const int num_tiles = num_m * num_n;
  const int kchunks = g.Cin / CBK;
  const int num_k = 27 * kchunks;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmW);
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  griddep_launch();
  griddep_wait();   // PDL: the prologue above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % num_m, n_blk = tile / num_m;
        const int iw = m_blk % g.nw, ih = (m_blk / g.nw) % g.nh, it = m_blk / (g.nw * g.nh);
        const int t0 = it * g.bt, h0 = ih * g.bh, w0 = iw * g.bw;
        for (int kb = 0; kb < num_k; ++kb) {
          const int tap = kb / kchunks, kc = kb % kchunks;
          const int dt = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], Cfg::A_BYTES + Cfg::B_BYTES);
          tma_load_4d(sA + stage * Cfg::A_BYTES, &tmX, &full[stage], kc * CBK, w0 + dw, h0 + dh, t0 + dt);
          tma_load_2d(sB + stage * Cfg::B_BYTES, &tmW, &full[stage], kc * CBK, tap * g.Cout + n_blk * BN);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(CBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int t = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
        const int as = t & 1;
        const uint32_t aphase = (t >> 1) & 1;
        mbar_wait(&tempty[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(sA + stage * Cfg::A_BYTES);
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_BYTES);
#pragma unroll
          for (int k = 0; k < CBK / 16; ++k)
            umma_bf16(d_tmem, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[as]);
      }
    }
  } else {
    const int q = warp & 3;
    int t = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
      const int m_blk = tile % num_m, n_blk = tile / num_m;
      const int iw = m_blk % g.nw, ih = (m_blk / g.nw) % g.nh, it = m_blk / (g.nw * g.nh);
      const int as = t & 1;
      const uint32_t aphase = (t >> 1) & 1;
      const int m = q * 32 + lane;  // row in tile -> voxel inside the box (w fastest)
      const int vw = iw * g.bw + m % g.bw;
      const int vh = ih * g.bh + (m / g.bw) % g.bh;
      const int vt = it * g.bt + m / (g.bw * g.bh);
      const bool valid = vw < g.W && vh < g.H && vt < g.T;
      mbar_wait(&tfull[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        tmem_ld_wait();
        if (valid) conv_epilogue_chunk<MODE>(r, vt, vh, vw, n_blk * BN + c * 32, g, ep);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------
// Padding prologue: one warp per padded voxel.
// mode 0: copy ; 1: x*a[c] + b[c] ; 2: silu(x / sqrt(mean_c x^2 + 1e-8) * (1 + a[c]) + b[c])
// ---------------------------------------------------------------------------------------------
// (x, a, b are produced by preceding kernels: no const __restrict__, see griddep_wait in ptx.cuh)
__global__ void __launch_bounds__(256) vae_prep_kernel(const float* x, bf16* out, int T, int H, int W, int C, int mode,
                                                        const float* a, const float* b, int tshift) {
  griddep_launch();
  griddep_wait();
  const int lane = threadIdx.x & 31;
  const int64_t nvox = static_cast<int64_t>(T + 2) * (H + 2) * (W + 2);
  const int64_t wid0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int nv = C >> 2;
  for (int64_t pv = wid0; pv < nvox; pv += nwarps) {
    const int pw = static_cast<int>(pv % (W + 2));
    const int ph = static_cast<int>((pv / (W + 2)) % (H + 2));
    const int pt = static_cast<int>(pv / (static_cast<int64_t>(W + 2) * (H + 2)));
    int ws = pw - 1; ws = ws < 0 ? -ws : (ws >= W ? 2 * W - 2 - ws : ws);   // reflect (VideoConvolution.swift:257-266)
    int hs = ph - 1; hs = hs < 0 ? -hs : (hs >= H ? 2 * H - 2 - hs : hs);
    int ts = pt - tshift; ts = ts < 0 ? 0 : (ts >= T ? T - 1 : ts);          // frame replication (:281-294)
    const float4* src = reinterpret_cast<const float4*>(x + ((static_cast<int64_t>(ts) * H + hs) * W + ws) * C);
    uint2* dst = reinterpret_cast<uint2*>(out + pv * C);
    float rs = 1.0f;
    if (mode == 2) {
      float ss = 0.f;
      for (int i = lane; i < nv; i += 32) {
        float4 v = src[i];
        ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      ss = warp_sum(ss);
      rs = rsqrtf(ss / C + 1e-8f);
    }
    for (int i = lane; i < nv; i += 32) {
      float4 v = src[i];
      if (mode == 1) {
        float4 av = reinterpret_cast<const float4*>(a)[i], bv = reinterpret_cast<const float4*>(b)[i];
        v.x = v.x * av.x + bv.x; v.y = v.y * av.y + bv.y; v.z = v.z * av.z + bv.z; v.w = v.w * av.w + bv.w;
      } else if (mode == 2) {
        float4 av = reinterpret_cast<const float4*>(a)[i], bv = reinterpret_cast<const float4*>(b)[i];
        v.x = silu(v.x * rs * (1.f + av.x) + bv.x);
        v.y = silu(v.y * rs * (1.f + av.y) + bv.y);
        v.z = silu(v.z * rs * (1.f + av.z) + bv.z);
        v.w = silu(v.w * rs * (1.f + av.w) + bv.w);
      }
      dst[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    }
  }
}

int best_pow2(int extent, int budget) {
  // power of two <= budget that wastes the least of `extent` when tiling; ties -> larger
  int best = 1;
  double best_u = 0.0;
  for (int b = 1; b <= budget; b <<= 1) {
    const int tiles = (extent + b - 1) / b;
    const double u = static_cast<double>(extent) / (static_cast<double>(tiles) * b);
    if (u >= best_u - 1e-9) { best_u = u; best = b; }
  }
  return best;
}

template <int BN, int MODE>
void conv_launch(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvGeom& g, const ConvEpi& ep, cudaStream_t s) {
  using Cfg = ConvCfg<BN>;
  static bool configured = false;
  auto kern = conv3d_tcgen05<BN, MODE>;
  if (!configured) {
    LTX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(Cfg::SMEM)));
    configured = true;
  }
  const int tiles = g.nt * g.nh * g.nw * ((g.Cout + BN - 1) / BN);
  const int grid = tiles < device_sm_count() ? tiles : device_sm_count();
  launch_pdl(PDL_VAE, kern, dim3(grid), dim3(CONV_THREADS), Cfg::SMEM, s, tmX, tmW, g, ep);
  LTX_CUDA(cudaGetLastError());
}

template <int BN>
void conv_launch_mode(const CUtensorMap& tmX, const CUtensorMap& tmW, const ConvGeom& g, const ConvEpi& ep, cudaStream_t s) {
  switch (ep.mode) {
    case 0: conv_launch<BN, 0>(tmX, tmW, g, ep, s); break;
    case 1: conv_launch<BN, 1>(tmX, tmW, g, ep, s); break;
    case 2: conv_launch<BN, 2>(tmX, tmW, g, ep, s); break;
    default: LTX_CHECK(false, 2, "bad conv epilogue mode");
  }
}

}  // namespace

void launch_conv3d(const bf16* x_pad, const bf16* w, int T, int H, int W, int Cin, int Cout, const ConvEpi& epi,
                   cudaStream_t s) {
  LTX_CHECK(T > 0 && H > 1 && W > 1, 2, "conv3d: bad volume (reflect padding needs H, W >= 2)");
  LTX_CHECK(Cin % 64 == 0, 2, "conv3d: Cin must be a multiple of 64");
  LTX_CHECK(epi.mode != 0 || Cout % 4 == 0, 2, "conv3d: Cout must be a multiple of 4");
  LTX_CHECK(epi.mode != 1 || (Cout % 32 == 0 && Cin % 8 == 0 && epi.resid != nullptr), 2, "conv3d: bad d2s configuration");
  ConvGeom g;
  g.T = T; g.H = H; g.W = W; g.Cin = Cin; g.Cout = Cout;
  g.bw = best_pow2(W, 128);
  g.bh = best_pow2(H, 128 / g.bw);
  g.bt = 128 / (g.bw * g.bh);
  g.nt = (T + g.bt - 1) / g.bt; g.nh = (H + g.bh - 1) / g.bh; g.nw = (W + g.bw - 1) / g.bw;
  const int bn = Cout >= 256 ? 256 : 128;
  CUtensorMap tmX = make_tmap_thwc(x_pad, T + 2, H + 2, W + 2, Cin, g.bt, g.bh, g.bw);
  CUtensorMap tmW = make_tmap_2d(w, static_cast<uint64_t>(27) * Cout, Cin, Cin, bn);
  if (bn == 256)
    conv_launch_mode<256>(tmX, tmW, g, epi, s);
  else
    conv_launch_mode<128>(tmX, tmW, g, epi, s);
}

void launch_vae_prep(const float* x, bf16* out, int T, int H, int W, int C, int mode, const float* a, const float* b,
                     int causal, cudaStream_t s) {
  LTX_CHECK(C % 4 == 0 && H > 1 && W > 1, 2, "vae_prep: bad shape");
  const int64_t nvox = static_cast<int64_t>(T + 2) * (H + 2) * (W + 2);
  int64_t blocks = (nvox + 7) / 8;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  launch_pdl(PDL_VAE, vae_prep_kernel, dim3(static_cast<int>(blocks)), dim3(256), 0, s, x, out, T, H, W, C, mode, a, b, causal ? 2 : 1);
  LTX_CUDA(cudaGetLastError());
}

}  // namespace ltx
