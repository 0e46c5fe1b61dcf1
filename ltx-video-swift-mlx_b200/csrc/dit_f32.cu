// fp32 mode of the LTX-2 video DiT forward (north_star: per-step velocity within rel-L2 1e-4 of the reference in fp32 mode;
// BASELINE config 0 is "one transformer block, random-init fp32").  Same graph as dit.cu (LTXTransformer.callAsFunction,
// Models/Transformer/LTXTransformer.swift:235-486; BasicTransformerBlock, LTXTransformerBlock.swift:187-232), but:
//   * weights stay fp32 in HBM (ltx_set_precision(ctx, 32) before loading), every activation is fp32;
//   * every Linear still runs on the tcgen05 bf16 tensor cores, as a SPLIT-bf16 product: an fp32 value is the exact sum
//     of three bf16 terms up to 2^-24 (v = t1 + t2 + t3), and bf16 x bf16 products are exact in the fp32 accumulator, so
//       A B^T ~= A3 B1 + A1 B3 + A2 B2 + A2 B1 + A1 B2 + A1 B1      (all terms of order <= 2^-16 kept, smallest first)
//     is one ordinary GEMM over the concatenated operands A' = [A3|A1|A2|A2|A1|A1], B' = [B1|B3|B2|B1|B2|B1] (K' = 6 K),
//     run by the unchanged kernel of gemm.cu / gemm2.cu with its fp32 or gate*residual epilogue.  The operands are split
//     by split3_kernel right before each GEMM (activations fused with their SiLU / GELU, with precise tanhf / expf);
//   * attention is an fp32 SIMT flash kernel (exact expf softmax), norms / RoPE are fp32 row kernels.
// A verification mode: about 6x the tensor work of bf16 mode plus the split passes; single GPU, no quantisation.
#include <cmath>

#include "ctx.h"
#include "ptx.cuh"

namespace ltx {

namespace {

__device__ __forceinline__ float act_apply(float v, int act) {
  if (act == 1) return v / (1.0f + expf(-v));                                                   // SiLU
  if (act == 2) return 0.5f * v * (1.0f + tanhf(0.7978845608028654f * (v + 0.044715f * v * v * v)));  // GELU-tanh
  return v;
}

// in [rows, K] fp32 (row pitch ld_in) -> out [rows, 6K] bf16: six K-wide segments holding the bf16 terms t1, t2, t3 of
// act(in) in the order given by `pat` (0: A operand [t3 t1 t2 t2 t1 t1], 1: B operand [t1 t3 t2 t1 t2 t1]).
__global__ void __launch_bounds__(256) split3_kernel(const float* in, int64_t ld_in, int64_t rows, int K, bf16* out, int pat,
                                                      int act) {
  const int k4 = K >> 2;
  const int64_t total = rows * k4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / k4;
    const int c = static_cast<int>(i - r * k4) * 4;
    const float4 v4 = *reinterpret_cast<const float4*>(in + r * ld_in + c);
    const float v[4] = {act_apply(v4.x, act), act_apply(v4.y, act), act_apply(v4.z, act), act_apply(v4.w, act)};
    bf16 t[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bf16 t1 = __float2bfloat16(v[e]);
      const float r1 = v[e] - __bfloat162float(t1);     // exact
      const bf16 t2 = __float2bfloat16(r1);
      const float r2 = r1 - __bfloat162float(t2);       // exact
      t[0][e] = t1; t[1][e] = t2; t[2][e] = __float2bfloat16(r2);
    }
    const int ordA[6] = {2, 0, 1, 1, 0, 0}, ordB[6] = {0, 2, 1, 0, 1, 0};
    bf16* orow = out + r * (6 * static_cast<int64_t>(K)) + c;
#pragma unroll
    for (int s = 0; s < 6; ++s) {
      const int w = pat == 0 ? ordA[s] : ordB[s];
      uint2 pk;
      pk.x = static_cast<uint32_t>(__bfloat16_as_ushort(t[w][0])) | (static_cast<uint32_t>(__bfloat16_as_ushort(t[w][1])) << 16);
      pk.y = static_cast<uint32_t>(__bfloat16_as_ushort(t[w][2])) | (static_cast<uint32_t>(__bfloat16_as_ushort(t[w][3])) << 16);
      *reinterpret_cast<uint2*>(orow + static_cast<int64_t>(s) * K) = pk;
    }
  }
}

__device__ __forceinline__ float block_sum_f32(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];
  return t;
}

// out = norm(x) * (1 + tbl_scale + ada_scale) + tbl_shift + ada_shift, fp32 in and out (one CTA per row)
__global__ void __launch_bounds__(256) rmsnorm_mod_f32_kernel(const float* x, float* out, int D, const float* tbl_shift,
                                                               const float* tbl_scale, const float* ada_shift,
                                                               const float* ada_scale, int64_t ada_ld, int rows_per_mod,
                                                               float eps, int layernorm) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float* xr = x + static_cast<int64_t>(row) * D;
  float s1 = 0.f, s2 = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) { const float v = xr[i]; s1 += v; s2 += v * v; }
  float mean = 0.f, rstd;
  if (layernorm) {
    mean = block_sum_f32(s1, red) / D;
    float sv = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) { const float d = xr[i] - mean; sv += d * d; }
    rstd = rsqrtf(block_sum_f32(sv, red) / D + eps);
  } else {
    rstd = rsqrtf(block_sum_f32(s2, red) / D + eps);
  }
  const int64_t aoff = static_cast<int64_t>(row / rows_per_mod) * ada_ld;
  float* orow = out + static_cast<int64_t>(row) * D;
  for (int i = threadIdx.x; i < D; i += blockDim.x)
    orow[i] = (xr[i] - mean) * rstd * (1.f + tbl_scale[i] + ada_scale[aoff + i]) + tbl_shift[i] + ada_shift[aoff + i];
}

// in place on fp32 rows: y = rms(x[:D]) * w, then optional split RoPE (T/LTXAttention.swift:179-189, T/LTXRoPE.swift:84-149)
__global__ void __launch_bounds__(256) qknorm_rope_f32_kernel(float* x, int64_t ld, int D, const float* __restrict__ w,
                                                               const float* __restrict__ cosb, const float* __restrict__ sinb,
                                                               int rows_per_rope, float eps) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  float* xr = x + static_cast<int64_t>(row) * ld;
  float ss = 0.f;
  for (int i = threadIdx.x; i < D; i += blockDim.x) { const float v = xr[i]; ss += v * v; }
  const float rstd = rsqrtf(block_sum_f32(ss, red) / D + eps);
  const int half = D >> 1;
  const float* cr = cosb ? cosb + static_cast<int64_t>(row % rows_per_rope) * half : nullptr;
  const float* sr = sinb ? sinb + static_cast<int64_t>(row % rows_per_rope) * half : nullptr;
  for (int p = threadIdx.x; p < half; p += blockDim.x) {   // pair p: head p / 64, element j = p % 64
    const int hh = p >> 6, j = p & 63;
    const int c1 = hh * 128 + j, c2 = c1 + 64;
    const float a = xr[c1] * rstd * w[c1], b = xr[c2] * rstd * w[c2];
    if (cr) {
      const float c = cr[p], s = sr[p];
      xr[c1] = a * c - b * s;
      xr[c2] = b * c + a * s;
    } else {
      xr[c1] = a;
      xr[c2] = b;
    }
  }
}

// fp32 flash attention, head_dim 128.  CTA = 8 warps x 4 queries of one (batch, head); keys in smem tiles of 32.
// Scores: lane = key; output: lane = 4 head dims.  Q [B*Nq, ldq], K / V [B*Nk, ld], O [B*Nq, ldo]; head h = columns h*128.
constexpr int AF_TK = 32, AF_QPW = 4, AF_WARPS = 8, AF_LD = 132;   // 132: float4-aligned row pitch, conflict-free
__global__ void __launch_bounds__(256) attention_f32_kernel(const float* Q, int64_t ldq, const float* Kp, int64_t ldk,
                                                             const float* Vp, int64_t ldv, const float* key_bias, float* O,
                                                             int64_t ldo, int Nq, int Nk, float scale) {
  extern __shared__ __align__(16) float af_smem[];
  float* sK = af_smem;                       // [AF_TK][AF_LD]
  float* sV = sK + AF_TK * AF_LD;            // [AF_TK][128]
  float* sQ = sV + AF_TK * 128;              // [AF_WARPS * AF_QPW][128]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * (AF_WARPS * AF_QPW) + warp * AF_QPW;
  // stage this warp's queries
  for (int qi = 0; qi < AF_QPW; ++qi) {
    const int q = q0 + qi;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q < Nq) v = *reinterpret_cast<const float4*>(Q + (static_cast<int64_t>(b) * Nq + q) * ldq + h * 128 + lane * 4);
    *reinterpret_cast<float4*>(&sQ[(warp * AF_QPW + qi) * 128 + lane * 4]) = v;
  }
  float m[AF_QPW], l[AF_QPW];
  float4 acc[AF_QPW];
#pragma unroll
  for (int qi = 0; qi < AF_QPW; ++qi) { m[qi] = -INFINITY; l[qi] = 0.f; acc[qi] = make_float4(0.f, 0.f, 0.f, 0.f); }
  for (int k0 = 0; k0 < Nk; k0 += AF_TK) {
    __syncthreads();
    for (int i = threadIdx.x; i < AF_TK * 32; i += blockDim.x) {   // 32 keys x 32 float4
      const int kr = i >> 5, c4 = i & 31;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + kr < Nk) {
        const int64_t grow = static_cast<int64_t>(b) * Nk + k0 + kr;
        kv = *reinterpret_cast<const float4*>(Kp + grow * ldk + h * 128 + c4 * 4);
        vv = *reinterpret_cast<const float4*>(Vp + grow * ldv + h * 128 + c4 * 4);
      }
      *reinterpret_cast<float4*>(&sK[kr * AF_LD + c4 * 4]) = kv;
      *reinterpret_cast<float4*>(&sV[kr * 128 + c4 * 4]) = vv;
    }
    __syncthreads();
    const int key = k0 + lane;
    const float kb = (key < Nk) ? (key_bias ? key_bias[static_cast<int64_t>(b) * Nk + key] : 0.f) : -INFINITY;
#pragma unroll
    for (int qi = 0; qi < AF_QPW; ++qi) {
      const float* qv = &sQ[(warp * AF_QPW + qi) * 128];
      const float* kr = &sK[lane * AF_LD];
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < 128; d += 4) {
        const float4 a = *reinterpret_cast<const float4*>(qv + d), k4 = *reinterpret_cast<const float4*>(kr + d);
        s += a.x * k4.x + a.y * k4.y + a.z * k4.z + a.w * k4.w;
      }
      s = s * scale + kb;
      const float mn = fmaxf(m[qi], warp_max(s));
      const float alpha = (m[qi] == -INFINITY) ? 0.f : expf(m[qi] - mn);
      const float p = (s == -INFINITY) ? 0.f : expf(s - mn);
      l[qi] = l[qi] * alpha + warp_sum(p);
      m[qi] = mn;
      float4 o = acc[qi];
      o.x *= alpha; o.y *= alpha; o.z *= alpha; o.w *= alpha;
#pragma unroll 8
      for (int j = 0; j < AF_TK; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
        const float4 v4 = *reinterpret_cast<const float4*>(&sV[j * 128 + lane * 4]);
        o.x += pj * v4.x; o.y += pj * v4.y; o.z += pj * v4.z; o.w += pj * v4.w;
      }
      acc[qi] = o;
    }
  }
#pragma unroll
  for (int qi = 0; qi < AF_QPW; ++qi) {
    const int q = q0 + qi;
    if (q >= Nq) continue;
    const float inv = 1.0f / l[qi];
    *reinterpret_cast<float4*>(O + (static_cast<int64_t>(b) * Nq + q) * ldo + h * 128 + lane * 4) =
        make_float4(acc[qi].x * inv, acc[qi].y * inv, acc[qi].z * inv, acc[qi].w * inv);
  }
}

__global__ void cast_any_f32_kernel(const bf16* in, float* out, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = __bfloat162float(in[i]);
}

int grid_1d(int64_t n, int per_block, int cap = 148 * 16) {
  int64_t b = (n + per_block - 1) / per_block;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

const float* wf32(ltx_ctx* c, const std::string& k, int64_t n) {
  const DevTensor& t = get_tensor(c, k);
  LTX_CHECK(t.dtype == LTX_F32 && t.numel() == n, LTX_ERR_WEIGHTS,
            "fp32 mode needs fp32 tensor '" + k + "' (call ltx_set_precision(ctx, 32) before loading weights)");
  return reinterpret_cast<const float*>(t.ptr);
}

// y = act_in(A) W^T (+ bias) through the split-bf16 tensor-core GEMM.  A [M, K] fp32 (row pitch lda), W [N, K] fp32.
void linear_f32(ltx_ctx* c, const float* A, int64_t lda, int M, int K, const float* W, int N, int act_in, const GemmEpi& e) {
  LTX_CHECK(K % 8 == 0, LTX_ERR_INVALID_CONFIGURATION, "fp32 mode: Linear input width must be a multiple of 8");
  const int64_t K6 = 6 * static_cast<int64_t>(K);
  c->f_asplit.reserve(static_cast<size_t>(M) * K6 * 2);
  c->f_wsplit.reserve(static_cast<size_t>(N) * K6 * 2);
  {
    ProfScope ps(c, PROF_OTHER, 0.0, 16.0 * (static_cast<double>(M) + N) * K, 2);
    split3_kernel<<<grid_1d(static_cast<int64_t>(M) * (K / 4), 256), 256, 0, c->stream>>>(A, lda, M, K, c->f_asplit.as<bf16>(), 0, act_in);
    split3_kernel<<<grid_1d(static_cast<int64_t>(N) * (K / 4), 256), 256, 0, c->stream>>>(W, K, N, K, c->f_wsplit.as<bf16>(), 1, 0);
    LTX_CUDA(cudaGetLastError());
  }
  ProfScope ps(c, PROF_GEMM, 2.0 * M * N * static_cast<double>(K6), 2.0 * (static_cast<double>(M) + N) * K6);
  launch_gemm(c->f_asplit.as<bf16>(), K6, c->f_wsplit.as<bf16>(), K6, M, N, static_cast<int>(K6), e, c->stream);
}

GemmEpi epi_f32(float* out, int64_t ldo, const float* bias) {
  GemmEpi e;
  e.mode = EPI_F32; e.out = out; e.ldo = ldo; e.bias = bias;
  return e;
}

void norm_f32(ltx_ctx* c, const float* x, float* out, int M, int D, const float* ts, const float* tsc, const float* as,
              const float* asc, int64_t ada_ld, int rows_per_mod, float eps, int ln) {
  ProfScope ps(c, PROF_ROW, 0.0, 8.0 * M * D);
  rmsnorm_mod_f32_kernel<<<M, 256, 0, c->stream>>>(x, out, D, ts, tsc, as, asc, ada_ld, rows_per_mod > 0 ? rows_per_mod : 1, eps, ln);
  LTX_CUDA(cudaGetLastError());
}
void qknorm_f32(ltx_ctx* c, float* x, int64_t ld, int M, int D, const float* w, const float* cs, const float* sn, int rpr, float eps) {
  ProfScope ps(c, PROF_ROW, 0.0, 8.0 * M * D);
  qknorm_rope_f32_kernel<<<M, 256, 0, c->stream>>>(x, ld, D, w, cs, sn, rpr > 0 ? rpr : 1, eps);
  LTX_CUDA(cudaGetLastError());
}
void attention_f32(ltx_ctx* c, const float* Q, int64_t ldq, const float* K, int64_t ldk, const float* V, int64_t ldv,
                   const float* bias, float* O, int64_t ldo, int B, int H, int Nq, int Nk, float scale) {
  ProfScope ps(c, PROF_ATTN, 4.0 * B * H * static_cast<double>(Nq) * Nk * 128.0, 4.0 * B * (2.0 * Nq + 2.0 * Nk) * H * 128.0);
  dim3 grid((Nq + AF_WARPS * AF_QPW - 1) / (AF_WARPS * AF_QPW), H, B);
  constexpr size_t smem = (AF_TK * AF_LD + AF_TK * 128 + AF_WARPS * AF_QPW * 128) * sizeof(float);
  ensure_dyn_smem(attention_f32_kernel, smem);
  attention_f32_kernel<<<grid, 256, smem, c->stream>>>(Q, ldq, K, ldk, V, ldv, bias, O, ldo, Nq, Nk, scale);
  LTX_CUDA(cudaGetLastError());
}

bool in_list32(int v, const int32_t* lst, int n) {
  for (int i = 0; i < n; ++i)
    if (lst[i] == v) return true;
  return false;
}

}  // namespace

// fp32 mode keeps the raw fp32 tensors as they were loaded (no packing); checks that everything the forward needs exists.
void dit_finalize_f32(ltx_ctx* c) {
  const ltx_config& g = c->cfg;
  LTX_CHECK(g.head_dim == 128, LTX_ERR_INVALID_CONFIGURATION, "head_dim must be 128");
  const int64_t D = static_cast<int64_t>(g.num_heads) * g.head_dim, FF = g.ffn_mult * D;
  wf32(c, "patchify_proj.weight", D * g.in_channels);
  wf32(c, "adaln_single.linear.weight", 6 * D * D);
  wf32(c, "caption_projection.linear_1.weight", D * g.caption_channels);
  wf32(c, "proj_out.weight", g.out_channels * D);
  for (int i = 0; i < g.num_layers; ++i) {
    const std::string p = "transformer_blocks." + std::to_string(i) + ".";
    for (const char* a : {"attn1", "attn2"})
      for (const char* l : {"to_q", "to_k", "to_v", "to_out"}) wf32(c, p + a + "." + l + ".weight", D * D);
    wf32(c, p + "ff.project_in.proj.weight", FF * D);
    wf32(c, p + "ff.project_out.weight", D * FF);
  }
  c->scratch.reserve(64 * sizeof(double));
  c->dit_ready = true;
}

void dit_forward_f32(ltx_ctx* c, const void* latent, int latent_dtype, const void* context, int context_dtype,
                     const float* timesteps_dev, int ts_per_token, const int32_t* mask_dev, int B, int N, int S, int F, int H,
                     int W, const ltx_dit_flags* flags, float* out_velocity_dev) {
  const ltx_config& g = c->cfg;
  const int D = g.num_heads * g.head_dim, FFD = g.ffn_mult * D, Hh = g.num_heads, L = g.num_layers;
  const int Cin = g.in_channels, Cout = g.out_channels, Cc = g.caption_channels;
  LTX_CHECK(!(c->dist.comm_world && c->dist.sp > 1), LTX_ERR_UNSUPPORTED, "fp32 mode does not support sequence parallelism");
  const int R = B * N, RS = B * S;
  const float eps = g.norm_eps;
  const float att_scale = 1.0f / sqrtf(static_cast<float>(g.head_dim));
  cudaStream_t st = c->stream;
  ltx_dit_flags noflags = {};
  noflags.cross_attn_scale = 1.0f;
  if (!flags) flags = &noflags;

  // ---- workspaces (all fp32)
  auto f32buf = [&](DevBuf& b, size_t n) { b.reserve(n * 4); return b.as<float>(); };
  float* x = f32buf(c->x, static_cast<size_t>(R) * D);
  float* h = f32buf(c->f_h, static_cast<size_t>(R) * D);
  float* q = f32buf(c->f_q, static_cast<size_t>(R) * D);
  float* k = f32buf(c->f_k, static_cast<size_t>(R) * D);
  float* v = f32buf(c->f_v, static_cast<size_t>(R) * D);
  float* att = f32buf(c->f_att, static_cast<size_t>(R) * D);
  float* ffh = f32buf(c->f_ffh, static_cast<size_t>(R) * FFD);
  const int TR = ts_per_token ? R : B;
  float* se = f32buf(c->se, static_cast<size_t>(TR) * 256);
  float* t1 = f32buf(c->t1, static_cast<size_t>(TR) * D);
  float* emb = f32buf(c->emb, static_cast<size_t>(TR) * D);
  float* ada = f32buf(c->ada, static_cast<size_t>(TR) * 6 * D);
  float* cx = f32buf(c->f_ctx, static_cast<size_t>(RS) * Cc);
  float* c1 = f32buf(c->f_c1, static_cast<size_t>(RS) * D);
  float* c2 = f32buf(c->f_c2, static_cast<size_t>(RS) * D);
  float* tk = f32buf(c->f_tk, static_cast<size_t>(RS) * D);
  float* tv = f32buf(c->f_tv, static_cast<size_t>(RS) * D);
  float* lat = f32buf(c->f_lat, static_cast<size_t>(R) * Cin);

  auto W_ = [&](const std::string& key, int64_t n) { return wf32(c, key, n); };
  auto to_f32 = [&](const void* src, int dtype, float* dst, int64_t n) {
    ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * n);
    if (dtype == LTX_F32) {
      LTX_CUDA(cudaMemcpyAsync(dst, src, static_cast<size_t>(n) * 4, cudaMemcpyDeviceToDevice, st));
    } else {
      LTX_CHECK(dtype == LTX_BF16, LTX_ERR_UNSUPPORTED, "input dtype must be bf16 or f32");
      cast_any_f32_kernel<<<grid_1d(n, 256), 256, 0, st>>>(reinterpret_cast<const bf16*>(src), dst, n);
      LTX_CUDA(cudaGetLastError());
    }
  };
  to_f32(latent, latent_dtype, lat, static_cast<int64_t>(R) * Cin);
  to_f32(context, context_dtype, cx, static_cast<int64_t>(RS) * Cc);

  // ---- RoPE table (shared builder lives in dit.cu), additive key bias
  dit_build_rope(c, F, H, W);
  const float* key_bias = nullptr;
  if (mask_dev) {
    c->f_bias.reserve(static_cast<size_t>(RS) * 4);
    ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * RS);
    launch_mask_to_bias(mask_dev, c->f_bias.as<float>(), RS, st);
    key_bias = c->f_bias.as<float>();
  }
  const float* cosb = c->rope_cos.as<float>();
  const float* sinb = c->rope_sin.as<float>();
  const int rows_per_b = ts_per_token ? 1 : N;

  // ---- patchify_proj, timestep path, caption projection
  linear_f32(c, lat, Cin, R, Cin, W_("patchify_proj.weight", static_cast<int64_t>(D) * Cin), D, 0,
             epi_f32(x, D, W_("patchify_proj.bias", D)));
  {
    ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * TR * 256);
    launch_sincos_embed(timesteps_dev, g.timestep_scale_multiplier, se, TR, 256, st);
  }
  linear_f32(c, se, 256, TR, 256, W_("adaln_single.emb.linear_1.weight", static_cast<int64_t>(D) * 256), D, 0,
             epi_f32(t1, D, W_("adaln_single.emb.linear_1.bias", D)));
  linear_f32(c, t1, D, TR, D, W_("adaln_single.emb.linear_2.weight", static_cast<int64_t>(D) * D), D, 1,
             epi_f32(emb, D, W_("adaln_single.emb.linear_2.bias", D)));
  linear_f32(c, emb, D, TR, D, W_("adaln_single.linear.weight", 6LL * D * D), 6 * D, 1,
             epi_f32(ada, 6 * D, W_("adaln_single.linear.bias", 6 * D)));
  linear_f32(c, cx, Cc, RS, Cc, W_("caption_projection.linear_1.weight", static_cast<int64_t>(D) * Cc), D, 0,
             epi_f32(c1, D, W_("caption_projection.linear_1.bias", D)));
  linear_f32(c, c1, D, RS, D, W_("caption_projection.linear_2.weight", static_cast<int64_t>(D) * D), D, 2,
             epi_f32(c2, D, W_("caption_projection.linear_2.bias", D)));

  const int64_t ada_ld = 6 * static_cast<int64_t>(D);
  const int64_t DD = static_cast<int64_t>(D) * D;
  for (int i = 0; i < L; ++i) {
    const std::string p = "transformer_blocks." + std::to_string(i) + ".";
    const float* sst = W_(p + "scale_shift_table", 6 * D);
    const bool flagged = in_list32(i, flags->stg_blocks, flags->n_stg_blocks);
    const bool skip_sa = flagged && flags->skip_self_attn;
    const bool skip_ff = flagged && flags->skip_ff;
    const float cas = in_list32(i, flags->cas_blocks, flags->n_cas_blocks) ? flags->cross_attn_scale : 1.0f;
    if (!skip_sa) {
      norm_f32(c, x, h, R, D, sst, sst + D, ada, ada + D, ada_ld, rows_per_b, eps, 0);
      linear_f32(c, h, D, R, D, W_(p + "attn1.to_q.weight", DD), D, 0, epi_f32(q, D, W_(p + "attn1.to_q.bias", D)));
      linear_f32(c, h, D, R, D, W_(p + "attn1.to_k.weight", DD), D, 0, epi_f32(k, D, W_(p + "attn1.to_k.bias", D)));
      linear_f32(c, h, D, R, D, W_(p + "attn1.to_v.weight", DD), D, 0, epi_f32(v, D, W_(p + "attn1.to_v.bias", D)));
      qknorm_f32(c, q, D, R, D, W_(p + "attn1.q_norm.weight", D), cosb, sinb, N, eps);
      qknorm_f32(c, k, D, R, D, W_(p + "attn1.k_norm.weight", D), cosb, sinb, N, eps);
      attention_f32(c, q, D, k, D, v, D, nullptr, att, D, B, Hh, N, N, att_scale);
      GemmEpi eo;
      eo.mode = EPI_GATE_RESID; eo.resid = x; eo.ldr = D; eo.bias = W_(p + "attn1.to_out.bias", D);
      eo.gate_a = ada + 2 * D; eo.gate_b = sst + 2 * D; eo.gate_ld = ada_ld; eo.rows_per_gate = rows_per_b;
      linear_f32(c, att, D, R, D, W_(p + "attn1.to_out.weight", DD), D, 0, eo);
    }
    {
      // cross-attention on the UN-normalised stream (T/LTXTransformerBlock.swift:205-214)
      linear_f32(c, x, D, R, D, W_(p + "attn2.to_q.weight", DD), D, 0, epi_f32(q, D, W_(p + "attn2.to_q.bias", D)));
      qknorm_f32(c, q, D, R, D, W_(p + "attn2.q_norm.weight", D), nullptr, nullptr, 1, eps);
      linear_f32(c, c2, D, RS, D, W_(p + "attn2.to_k.weight", DD), D, 0, epi_f32(tk, D, W_(p + "attn2.to_k.bias", D)));
      qknorm_f32(c, tk, D, RS, D, W_(p + "attn2.k_norm.weight", D), nullptr, nullptr, 1, eps);
      linear_f32(c, c2, D, RS, D, W_(p + "attn2.to_v.weight", DD), D, 0, epi_f32(tv, D, W_(p + "attn2.to_v.bias", D)));
      attention_f32(c, q, D, tk, D, tv, D, key_bias, att, D, B, Hh, N, S, att_scale);
      GemmEpi eo;
      eo.mode = EPI_GATE_RESID; eo.resid = x; eo.ldr = D; eo.bias = W_(p + "attn2.to_out.bias", D); eo.scale = cas;
      linear_f32(c, att, D, R, D, W_(p + "attn2.to_out.weight", DD), D, 0, eo);
    }
    if (!skip_ff) {
      norm_f32(c, x, h, R, D, sst + 3 * D, sst + 4 * D, ada + 3 * D, ada + 4 * D, ada_ld, rows_per_b, eps, 0);
      linear_f32(c, h, D, R, D, W_(p + "ff.project_in.proj.weight", static_cast<int64_t>(FFD) * D), FFD, 0,
                 epi_f32(ffh, FFD, W_(p + "ff.project_in.proj.bias", FFD)));
      GemmEpi eo;
      eo.mode = EPI_GATE_RESID; eo.resid = x; eo.ldr = D; eo.bias = W_(p + "ff.project_out.bias", D);
      eo.gate_a = ada + 5 * D; eo.gate_b = sst + 5 * D; eo.gate_ld = ada_ld; eo.rows_per_gate = rows_per_b;
      linear_f32(c, ffh, FFD, R, FFD, W_(p + "ff.project_out.weight", static_cast<int64_t>(D) * FFD), D, 2, eo);  // GELU fused in the split
    }
  }
  // ---- output head (T/LTXTransformer.swift:208-224)
  const float* sst_out = W_("scale_shift_table", 2 * D);
  norm_f32(c, x, h, R, D, sst_out, sst_out + D, emb, emb, D, rows_per_b, eps, 1);
  linear_f32(c, h, D, R, D, W_("proj_out.weight", static_cast<int64_t>(Cout) * D), Cout, 0,
             epi_f32(out_velocity_dev, Cout, W_("proj_out.bias", Cout)));
}

}  // namespace ltx
