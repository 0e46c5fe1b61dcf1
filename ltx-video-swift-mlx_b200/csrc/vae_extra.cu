// The callers either side of the denoise loop that share the VAE's implicit-GEMM convolution:
//   * video-VAE encoder   -- VideoEncoder.callAsFunction (Models/VAE/VideoEncoder.swift:270-312) + the latent normalisation of
//                            encodeImage (Pipeline/LTXPipeline.swift:1902-1932): image / clip -> conditioning latent;
//   * latent upscaler     -- SpatialUpscaler.callAsFunction (Models/Upscaler/SpatialUpscaler.swift:215-258) wrapped in
//                            upsampleLatents (:360-383): denormalise -> 2x spatial upscale -> renormalise;
//   * adainFilterLatent   -- Pipeline/LatentUtils.swift:201-227; re-noise -- Pipeline/LTXPipeline.swift:2644-2647.
// Every 3x3x3 / 3x3 convolution runs on conv3d_tcgen05 (conv3d.cu) through vae_conv(): the padding prologue materialises the
// layer's own padding (encoder: zero H/W + causal frame replication; upscaler: zeros on all three axes) together with the
// activation in front of the conv (pixel-norm + SiLU, or GroupNorm folded into a per-channel affine + SiLU).  The
// bandwidth-bound pieces around them (patchify, space-to-depth + group-mean residual, GroupNorm statistics, pixel shuffle,
// AdaIN) are plain coalesced kernels on channels-last fp32 volumes.
#include "ctx.h"

namespace ltx {

namespace {

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }

// pixels [3, T, H, W] fp32 -> channels-last [T, H/4, W/4, 64]: channel (c*4 + pw)*4 + ph (pW before pH,
// V/VideoEncoder.swift:13-32), channels 48..63 zero (the conv kernel's K granularity)
__global__ void enc_patchify_kernel(const float* __restrict__ px, float* __restrict__ out, int T, int H, int W) {
  const int H4 = H >> 2, W4 = W >> 2;
  const int64_t n = static_cast<int64_t>(T) * H4 * W4 * 64;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx & 63);
    const int64_t v = idx >> 6;
    const int w4 = static_cast<int>(v % W4), h4 = static_cast<int>((v / W4) % H4), t = static_cast<int>(v / (static_cast<int64_t>(W4) * H4));
    float val = 0.f;
    if (ch < 48) {
      const int c = ch >> 4, pw = (ch >> 2) & 3, ph = ch & 3;
      val = px[((static_cast<int64_t>(c) * T + t) * H + (h4 * 4 + ph)) * W + (w4 * 4 + pw)];
    }
    out[idx] = val;
  }
}

// VAESpaceToDepthDownsample3d (V/VideoEncoder.swift:127-166):
//   out[t', h', w', co] = conv[ts, hs, ws, co / P] + mean_{g < G} x[ts_g, hs_g, ws_g, ...]
// with P = ft*fh*fw, space-to-depth channel c*P + (it*fh + ih)*fw + iw (:38-66), time front-padded with copies of frame 0
// when T % ft != 0, G = Cin*P / Cout residual channels averaged per output channel.
__global__ void s2d_residual_kernel(const float* conv, const float* x, float* out, int T, int H, int W, int Cc, int Cin, int ft,
                                    int fh, int fw, int To, int Ho, int Wo, int Cout) {
  const int P = ft * fh * fw;
  const int G = Cin * P / Cout;
  const int padT = To * ft - T;
  const int64_t n = static_cast<int64_t>(To) * Ho * Wo * Cout;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % Cout);
    const int64_t v = idx / Cout;
    const int wo = static_cast<int>(v % Wo), ho = static_cast<int>((v / Wo) % Ho), to = static_cast<int>(v / (static_cast<int64_t>(Wo) * Ho));
    auto src = [&](int sc, const float* vol, int C) {   // space-to-depth channel sc of a [T,H,W,C] volume
      const int c = sc / P, r = sc % P;
      const int it = r / (fh * fw), ih = (r / fw) % fh, iw = r % fw;
      int ts = to * ft + it - padT;
      ts = ts < 0 ? 0 : ts;
      return vol[((static_cast<int64_t>(ts) * H + (ho * fh + ih)) * W + (wo * fw + iw)) * C + c];
    };
    float acc = 0.f;
    for (int g = 0; g < G; ++g) acc += src(co * G + g, x, Cin);
    out[idx] = src(co, conv, Cc) + acc / static_cast<float>(G);
  }
}

// tokens [V, C] -> latent [C, V] with the per-channel affine (x - sub[c]) / div[c] (sub/div nullable)
__global__ void to_channel_major_kernel(const float* in, float* out, int64_t V, int C, const float* sub, const float* div) {
  __shared__ float tile[32][33];
  const int64_t v0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t v = v0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (v < V && c < C) ? in[v * C + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i;
    const int64_t v = v0 + threadIdx.x;
    if (c < C && v < V) {
      float val = tile[threadIdx.x][i];
      if (sub) val = (val - sub[c]) / div[c];
      out[static_cast<int64_t>(c) * V + v] = val;
    }
  }
}

// ---- GroupNorm(32) over a channels-last volume (UpscalerGroupNorm3D, SpatialUpscaler.swift:12-58): statistics over
// (D, H, W, C/32) in fp64, folded into a per-channel affine  y = x * a[c] + b[c]  that the consumer applies.
// pass 1: block partial sums per group, each thread owns one float4 of channels; pass 2: fixed-order reduction (deterministic)
__global__ void __launch_bounds__(256) gn_partial_kernel(const float* x, int64_t V, int C, double* part /* [grid][32][2] */) {
  extern __shared__ double sm[];   // [nvb][C][2]
  const int qpv = C >> 2;                  // float4 per voxel
  const int nvb = 256 / qpv;               // voxels per block iteration
  const int q = threadIdx.x % qpv, vs = threadIdx.x / qpv;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, ss[4] = {0.f, 0.f, 0.f, 0.f};
  double ds[4] = {0, 0, 0, 0}, dss[4] = {0, 0, 0, 0};
  int cnt = 0;
  if (vs < nvb) {
    for (int64_t v = static_cast<int64_t>(blockIdx.x) * nvb + vs; v < V; v += static_cast<int64_t>(gridDim.x) * nvb) {
      const float4 f = reinterpret_cast<const float4*>(x + v * C)[q];
      s[0] += f.x; s[1] += f.y; s[2] += f.z; s[3] += f.w;
      ss[0] += f.x * f.x; ss[1] += f.y * f.y; ss[2] += f.z * f.z; ss[3] += f.w * f.w;
      if (++cnt == 64) {   // flush the fp32 partials into fp64 every 64 voxels
#pragma unroll
        for (int i = 0; i < 4; ++i) { ds[i] += s[i]; dss[i] += ss[i]; s[i] = 0.f; ss[i] = 0.f; }
        cnt = 0;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ds[i] += s[i]; dss[i] += ss[i];
      sm[(static_cast<size_t>(vs) * C + q * 4 + i) * 2] = ds[i];
      sm[(static_cast<size_t>(vs) * C + q * 4 + i) * 2 + 1] = dss[i];
    }
  }
  __syncthreads();
  const int cpg = C / 32;
  if (threadIdx.x < 32) {
    double a = 0, b = 0;
    for (int vv = 0; vv < nvb; ++vv)
      for (int cc = 0; cc < cpg; ++cc) {
        a += sm[(static_cast<size_t>(vv) * C + threadIdx.x * cpg + cc) * 2];
        b += sm[(static_cast<size_t>(vv) * C + threadIdx.x * cpg + cc) * 2 + 1];
      }
    part[(static_cast<size_t>(blockIdx.x) * 32 + threadIdx.x) * 2] = a;
    part[(static_cast<size_t>(blockIdx.x) * 32 + threadIdx.x) * 2 + 1] = b;
  }
}

__global__ void gn_finalize_kernel(const double* part, int nblk, int64_t V, int C, const float* w, const float* bias, float eps,
                                   float* a_out, float* b_out) {
  __shared__ float s_mean[32], s_rstd[32];
  const int cpg = C / 32;
  if (threadIdx.x < 32) {
    double a = 0, b = 0;
    for (int i = 0; i < nblk; ++i) {
      a += part[(static_cast<size_t>(i) * 32 + threadIdx.x) * 2];
      b += part[(static_cast<size_t>(i) * 32 + threadIdx.x) * 2 + 1];
    }
    const double n = static_cast<double>(V) * cpg;
    const double mean = a / n;
    double var = b / n - mean * mean;
    var = var < 0 ? 0 : var;
    s_mean[threadIdx.x] = static_cast<float>(mean);
    s_rstd[threadIdx.x] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  }
  __syncthreads();
  for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
    const int g = ch / cpg;
    const float sc = s_rstd[g] * w[ch];
    a_out[ch] = sc;
    b_out[ch] = bias[ch] - s_mean[g] * sc;
  }
}

// x_out = silu(y * a[c] + b[c] (+ resid)) on a channels-last fp32 volume (in place allowed)
__global__ void gn_apply_kernel(const float* y, const float* a, const float* b, const float* resid, float* out, int64_t n4, int C4) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c4 = static_cast<int>(i % C4);
    const float4 yv = reinterpret_cast<const float4*>(y)[i];
    const float4 av = reinterpret_cast<const float4*>(a)[c4], bv = reinterpret_cast<const float4*>(b)[c4];
    float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
    if (resid) r = reinterpret_cast<const float4*>(resid)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(silu_f(yv.x * av.x + bv.x + r.x), silu_f(yv.y * av.y + bv.y + r.y),
                                                     silu_f(yv.z * av.z + bv.z + r.z), silu_f(yv.w * av.w + bv.w + r.w));
  }
}

// PixelShuffle(2) on channels-last frames (SpatialUpscaler.swift:111-125): in [D, H, W, 4C], channel c*4 + i*2 + j ->
// out [D, 2H, 2W, C] at (2h + i, 2w + j, c)
__global__ void pixel_shuffle_kernel(const float* in, float* out, int D, int H, int W, int C) {
  const int64_t n = static_cast<int64_t>(D) * 2 * H * 2 * W * C;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < n;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const int64_t v = idx / C;
    const int wo = static_cast<int>(v % (2 * W)), ho = static_cast<int>((v / (2 * W)) % (2 * H));
    const int d = static_cast<int>(v / (static_cast<int64_t>(4) * W * H));
    const int h = ho >> 1, i = ho & 1, w = wo >> 1, j = wo & 1;
    out[idx] = in[((static_cast<int64_t>(d) * H + h) * W + w) * (4 * C) + c * 4 + i * 2 + j];
  }
}

// per-channel mean / std (population) of a channel-major [C, n] tensor: one block per channel, two passes in fp64
__global__ void __launch_bounds__(512) channel_stats_kernel(const float* x, int64_t n, float* mean_out, float* std_out) {
  __shared__ double red[512];
  __shared__ double s_mean;
  const float* p = x + static_cast<int64_t>(blockIdx.x) * n;
  double a = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) a += p[i];
  red[threadIdx.x] = a;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_mean = red[0] / static_cast<double>(n);
  __syncthreads();
  const double m = s_mean;
  double b = 0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const double d = p[i] - m;
    b += d * d;
  }
  __syncthreads();
  red[threadIdx.x] = b;
  __syncthreads();
  for (int o = 256; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    mean_out[blockIdx.x] = static_cast<float>(m);
    std_out[blockIdx.x] = static_cast<float>(sqrt(red[0] / static_cast<double>(n)));
  }
}

// adainFilterLatent (P/LatentUtils.swift:201-227): x = ((x - lm) / (ls + 1e-8) * rs + rm) blended with x by `factor`
__global__ void adain_apply_kernel(float* x, int64_t n, const float* lm, const float* ls, const float* rm, const float* rs,
                                   float factor) {
  const int c = blockIdx.y;
  const float m = lm[c], inv = 1.0f / (ls[c] + 1e-8f), rsd = rs[c], rmn = rm[c];
  float* p = x + static_cast<int64_t>(c) * n;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = p[i];
    const float r = (v - m) * inv * rsd + rmn;
    p[i] = factor >= 1.0f ? r : factor * r + (1.0f - factor) * v;
  }
}

__global__ void renoise_kernel(float* x, const float* nz, int64_t n, float s) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    x[i] = s * nz[i] + (1.0f - s) * x[i];
}

inline int grid_for(int64_t n, int per_block = 256, int cap = 148 * 16) {
  int64_t b = (n + per_block - 1) / per_block;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

const int kEncFactor[4][3] = {{1, 2, 2}, {2, 1, 1}, {2, 2, 2}, {2, 2, 2}};   // V/VideoEncoder.swift:239-258

}  // namespace

// ---------------------------------------------------------------------------------------------------------------- encoder
void vae_encoder_finalize(ltx_ctx* c) {
  EncWeights& e = c->enc;
  const std::string P = "vae_encoder.";
  const DevTensor& w0 = get_tensor(c, P + "conv_in.conv.weight");
  LTX_CHECK(w0.shape.size() == 5 && w0.shape[1] == 48, LTX_ERR_WEIGHTS, "vae_encoder.conv_in must take 48 patchified channels");
  const int64_t base = w0.shape[0];
  LTX_CHECK(base % 64 == 0, LTX_ERR_WEIGHTS, "vae_encoder base channels must be a multiple of 64");
  e.conv_in = vae_pack_conv_keys(c, P + "conv_in.conv.weight", P + "conv_in.conv.bias", base, 48, base, 27);
  auto res = [&](const std::string& name, int64_t ch) {
    EncResBlock r;
    r.c1 = vae_pack_conv_keys(c, name + ".conv1.conv.weight", name + ".conv1.conv.bias", ch, ch, ch, 27);
    r.c2 = vae_pack_conv_keys(c, name + ".conv2.conv.weight", name + ".conv2.conv.bias", ch, ch, ch, 27);
    return r;
  };
  int64_t ch = base;
  for (int i = 0; i < 4; ++i) {
    const std::string blk = P + "down_blocks_" + std::to_string(i);
    e.stage[i].res.clear();
    for (int j = 0; c->tensors.count(blk + ".resnets.resnets." + std::to_string(j) + ".conv1.conv.weight"); ++j)
      e.stage[i].res.push_back(res(blk + ".resnets.resnets." + std::to_string(j), ch));
    const int prod = kEncFactor[i][0] * kEncFactor[i][1] * kEncFactor[i][2];
    const int64_t cout = 2 * ch;
    LTX_CHECK(cout % prod == 0 && (cout / prod) % 4 == 0, LTX_ERR_WEIGHTS, "vae_encoder: bad downsample channel plan");
    e.stage[i].down = vae_pack_conv_keys(c, blk + ".downsamplers.conv.conv.weight", blk + ".downsamplers.conv.conv.bias", cout / prod,
                                         ch, cout / prod, 27);
    e.stage[i].cout = static_cast<int>(cout);
    ch = cout;
  }
  e.mid.clear();
  for (int j = 0; c->tensors.count(P + "mid_block.resnets." + std::to_string(j) + ".conv1.conv.weight"); ++j)
    e.mid.push_back(res(P + "mid_block.resnets." + std::to_string(j), ch));
  const DevTensor& wo = get_tensor(c, P + "conv_out.conv.weight");
  const int64_t C = c->cfg.vae_latent_channels;
  LTX_CHECK(wo.shape.size() == 5 && wo.shape[0] >= C, LTX_ERR_WEIGHTS, "vae_encoder.conv_out must have >= latent_channels outputs");
  // only the mean channels are used (V/VideoEncoder.swift:307-309): the log-variance row of the kernel is dropped at pack time
  e.conv_out = vae_pack_conv_keys(c, P + "conv_out.conv.weight", P + "conv_out.conv.bias", wo.shape[0], ch, C, 27);
  e.base = static_cast<int>(base);
  e.ready = true;
}

void vae_encode_dev(ltx_ctx* c, const float* pixels_dev, int T, int H, int W, int normalize, float* latent_dev) {
  EncWeights& e = c->enc;
  LTX_CHECK(e.ready, LTX_ERR_WEIGHTS, "VAE encoder weights not loaded / finalized");
  LTX_CHECK(pixels_dev && latent_dev && T >= 1 && H % 32 == 0 && W % 32 == 0 && H >= 64 && W >= 64, LTX_ERR_INVALID_ARGUMENT,
            "vae_encode: H and W must be multiples of 32 (>= 64)");
  LTX_CHECK(!normalize || c->vae.ready, LTX_ERR_WEIGHTS, "latent normalisation needs the decoder's per-channel statistics");
  cudaStream_t st = c->stream;
  const int pad = VAE_PAD_CAUSAL | VAE_PAD_ZERO_HW;   // VideoEncoder(causal: true), spatialPaddingMode .zeros (:222-227)
  int t = T, h = H / 4, w = W / 4;
  // the widest activation is the conv_in output / the patchified input (64 padded channels)
  size_t max_x = static_cast<size_t>(t) * h * w * std::max(64, e.base);
  {
    int tt = t, hh = h, ww = w;
    int64_t ch = e.base;
    for (int i = 0; i < 4; ++i) {
      tt = (tt + kEncFactor[i][0] - 1) / kEncFactor[i][0]; hh /= kEncFactor[i][1]; ww /= kEncFactor[i][2];
      ch *= 2;
      max_x = std::max(max_x, static_cast<size_t>(tt) * hh * ww * ch);
    }
  }
  c->v_a.reserve(max_x * 4);
  c->v_b.reserve(max_x * 4);
  c->v_h.reserve(max_x * 4);
  float* x = c->v_a.as<float>();
  float* y = c->v_b.as<float>();
  float* hb = c->v_h.as<float>();
  {
    const int64_t n = static_cast<int64_t>(t) * h * w * 64;
    ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * n + 4.0 * 3 * T * H * W);
    enc_patchify_kernel<<<grid_for(n), 256, 0, st>>>(pixels_dev, hb, T, H, W);
    LTX_CUDA(cudaGetLastError());
  }
  vae_conv(c, hb, 0, nullptr, nullptr, e.conv_in, t, h, w, pad, 0, x, nullptr);
  auto resblock = [&](const EncResBlock& r) {   // h = conv1(silu(pn(x))) ; x += conv2(silu(pn(h)))   (:72-101)
    vae_conv(c, x, 2, nullptr, nullptr, r.c1, t, h, w, pad, 0, hb, nullptr);
    vae_conv(c, hb, 2, nullptr, nullptr, r.c2, t, h, w, pad, 0, x, x);
  };
  int64_t ch = e.base;
  for (int i = 0; i < 4; ++i) {
    for (const EncResBlock& r : e.stage[i].res) resblock(r);
    const int ft = kEncFactor[i][0], fh = kEncFactor[i][1], fw = kEncFactor[i][2];
    LTX_CHECK(h % fh == 0 && w % fw == 0, LTX_ERR_INVALID_ARGUMENT, "vae_encode: spatial size not divisible by the stage factor");
    vae_conv(c, x, 0, nullptr, nullptr, e.stage[i].down, t, h, w, pad, 0, y, nullptr);
    const int to = (t + ft - 1) / ft, ho = h / fh, wo = w / fw;
    const int64_t n = static_cast<int64_t>(to) * ho * wo * e.stage[i].cout;
    {
      ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * (n + static_cast<double>(t) * h * w * (ch + e.stage[i].down.cout)));
      s2d_residual_kernel<<<grid_for(n), 256, 0, st>>>(y, x, hb, t, h, w, e.stage[i].down.cout, static_cast<int>(ch), ft, fh, fw, to,
                                                       ho, wo, e.stage[i].cout);
      LTX_CUDA(cudaGetLastError());
    }
    std::swap(x, hb);
    t = to; h = ho; w = wo; ch = e.stage[i].cout;
    LTX_CHECK(h > 1 && w > 1, LTX_ERR_INVALID_ARGUMENT, "vae_encode: input too small");
  }
  for (const EncResBlock& r : e.mid) resblock(r);
  vae_conv(c, x, 2, nullptr, nullptr, e.conv_out, t, h, w, pad, 0, y, nullptr);   // pn + SiLU + conv_out (:297-303)
  const int C = e.conv_out.cout;
  const int64_t V = static_cast<int64_t>(t) * h * w;
  ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * V * C);
  to_channel_major_kernel<<<dim3(static_cast<unsigned>((V + 31) / 32), (C + 31) / 32), dim3(32, 8), 0, st>>>(
      y, latent_dev, V, C, normalize ? c->vae.mean : nullptr, normalize ? c->vae.std : nullptr);
  LTX_CUDA(cudaGetLastError());
}

// --------------------------------------------------------------------------------------------------------------- upscaler
void upscaler_finalize(ltx_ctx* c) {
  UpsWeights& u = c->ups;
  const std::string P = "upscaler.";
  const DevTensor& w0 = get_tensor(c, P + "initial_conv.weight");
  LTX_CHECK(w0.shape.size() == 5, LTX_ERR_WEIGHTS, "upscaler.initial_conv.weight must be (O, I, 3, 3, 3)");
  const int64_t mid = w0.shape[0], cin = w0.shape[1];   // mid_channels detected from the weights (SpatialUpscaler.swift:268-272)
  LTX_CHECK((mid == 128 || mid == 256 || mid == 512 || mid == 1024) && cin % 64 == 0, LTX_ERR_WEIGHTS, "upscaler: mid channels must be 128, 256, 512 or 1024");
  auto norm = [&](const std::string& n) {
    GnW g;
    g.w = vae_vec(c, n + ".weight", mid);
    g.b = vae_vec(c, n + ".bias", mid);
    return g;
  };
  auto conv3 = [&](const std::string& n, int64_t co, int64_t ci) {
    return vae_pack_conv_keys(c, n + ".weight", n + ".bias", co, ci, co, 27);
  };
  u.initial = conv3(P + "initial_conv", mid, cin);
  u.initial_norm = norm(P + "initial_norm");
  auto blocks = [&](const std::string& grp, std::vector<UpsBlock>& out) {
    out.clear();
    for (int i = 0; c->tensors.count(grp + "." + std::to_string(i) + ".conv1.weight"); ++i) {
      const std::string b = grp + "." + std::to_string(i);
      UpsBlock k;
      k.c1 = conv3(b + ".conv1", mid, mid);
      k.n1 = norm(b + ".norm1");
      k.c2 = conv3(b + ".conv2", mid, mid);
      k.n2 = norm(b + ".norm2");
      out.push_back(k);
    }
  };
  blocks(P + "res_blocks", u.pre);
  blocks(P + "post_upsample_res_blocks", u.post);
  u.up2d = vae_pack_conv_keys(c, P + "upsampler.conv.weight", P + "upsampler.conv.bias", 4 * mid, mid, 4 * mid, 9);
  u.final_conv = conv3(P + "final_conv", cin, mid);
  u.mid = static_cast<int>(mid);
  u.cin = static_cast<int>(cin);
  u.ready = true;
}

void upscale_latent_dev(ltx_ctx* c, const float* latent_dev, int F, int H, int W, float* out_dev) {
  UpsWeights& u = c->ups;
  LTX_CHECK(u.ready, LTX_ERR_WEIGHTS, "upscaler weights not loaded / finalized");
  LTX_CHECK(c->vae.ready, LTX_ERR_WEIGHTS, "upsampleLatents needs the VAE's per-channel latent statistics");
  LTX_CHECK(latent_dev && out_dev && F >= 1 && H > 1 && W > 1, LTX_ERR_INVALID_ARGUMENT, "bad upscale_latent arguments");
  LTX_CHECK(u.cin == c->cfg.vae_latent_channels, LTX_ERR_WEIGHTS, "upscaler input channels != latent channels");
  cudaStream_t st = c->stream;
  const int C = u.mid, pad = VAE_PAD_ZERO_HW | VAE_PAD_ZERO_T;   // MLX Conv3d(padding: 1)
  const size_t max_x = static_cast<size_t>(F) * 2 * H * 2 * W * C;
  c->v_a.reserve(max_x * 4);
  c->v_b.reserve(max_x * 4);
  c->v_h.reserve(max_x * 4);
  float* x = c->v_a.as<float>();
  float* y = c->v_b.as<float>();
  float* z = c->v_h.as<float>();
  const int nblk = 148 * 2;
  c->u_part.reserve(static_cast<size_t>(nblk) * 32 * 2 * 8);
  c->u_ab.reserve(static_cast<size_t>(C) * 2 * 4);
  float* ga = c->u_ab.as<float>();
  float* gb = ga + C;
  // GroupNorm statistics of vol -> (ga, gb)
  auto gn_stats = [&](const float* vol, int64_t V, const GnW& g) {
    const int nvb = 256 / (C / 4);
    ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * V * C, 2);
    gn_partial_kernel<<<nblk, 256, static_cast<size_t>(nvb) * C * 2 * 8, st>>>(vol, V, C, c->u_part.as<double>());
    gn_finalize_kernel<<<1, 256, 0, st>>>(c->u_part.as<double>(), nblk, V, C, g.w, g.b, 1e-5f, ga, gb);
    LTX_CUDA(cudaGetLastError());
  };
  auto gn_apply = [&](const float* vol, const float* resid, float* out, int64_t V) {
    const int64_t n4 = V * C / 4;
    ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * V * C * (resid ? 3.0 : 2.0));
    gn_apply_kernel<<<grid_for(n4), 256, 0, st>>>(vol, ga, gb, resid, out, n4, C / 4);
    LTX_CUDA(cudaGetLastError());
  };
  ensure_dyn_smem(gn_partial_kernel, 64 * 1024);
  int h = H, w = W;
  int64_t V = static_cast<int64_t>(F) * h * w;
  // [C, F*H*W] -> channels-last; denormalise (x * std + mean, SpatialUpscaler.swift:370-371) inside the first padding prologue
  {
    ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * V * u.cin);
    launch_transpose_slice(latent_dev, V, u.cin, static_cast<int>(V), z, st);
  }
  vae_conv(c, z, 1, c->vae.std, c->vae.mean, u.initial, F, h, w, pad, 0, y, nullptr);
  gn_stats(y, V, u.initial_norm);
  gn_apply(y, nullptr, x, V);                                   // x = silu(norm(conv(x)))   (:225-229)
  auto resblock = [&](const UpsBlock& b) {                      // UpscalerResBlock3D (:62-107)
    vae_conv(c, x, 0, nullptr, nullptr, b.c1, F, h, w, pad, 0, y, nullptr);
    gn_stats(y, V, b.n1);
    vae_conv(c, y, 3, ga, gb, b.c2, F, h, w, pad, 0, z, nullptr);   // conv2(silu(norm1(.))): norm + SiLU in the prologue
    gn_stats(z, V, b.n2);
    gn_apply(z, x, x, V);                                        // x = silu(norm2(.) + x)
  };
  for (const UpsBlock& b : u.pre) resblock(b);
  // SpatialRationalResampler (:129-163): per-frame Conv2d mid -> 4*mid (the dt = 1 taps only) + PixelShuffle(2)
  vae_conv(c, x, 0, nullptr, nullptr, u.up2d, F, h, w, pad, 0, y, nullptr);
  {
    const int64_t n = V * 4 * C;
    ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * n);
    pixel_shuffle_kernel<<<grid_for(n), 256, 0, st>>>(y, x, F, h, w, C);
    LTX_CUDA(cudaGetLastError());
  }
  h *= 2; w *= 2; V *= 4;
  for (const UpsBlock& b : u.post) resblock(b);
  vae_conv(c, x, 0, nullptr, nullptr, u.final_conv, F, h, w, pad, 0, y, nullptr);
  // back to [C, F, 2H, 2W], renormalised ((x - mean) / std, :379-380)
  ProfScope ps(c, PROF_OTHER, 0.0, 8.0 * V * u.cin);
  to_channel_major_kernel<<<dim3(static_cast<unsigned>((V + 31) / 32), (u.cin + 31) / 32), dim3(32, 8), 0, st>>>(
      y, out_dev, V, u.cin, c->vae.mean, c->vae.std);
  LTX_CUDA(cudaGetLastError());
}

void adain_filter_dev(ltx_ctx* c, float* latent_dev, int64_t n, const float* ref_dev, int64_t n_ref, int C, float factor) {
  LTX_CHECK(latent_dev && ref_dev && n > 0 && n_ref > 0 && C > 0, LTX_ERR_INVALID_ARGUMENT, "bad adain arguments");
  if (factor <= 0.f) return;
  c->u_stats.reserve(static_cast<size_t>(C) * 4 * 4);
  float* s = c->u_stats.as<float>();
  ProfScope ps(c, PROF_OTHER, 0.0, 4.0 * C * (4.0 * n + 2.0 * n_ref), 3);
  channel_stats_kernel<<<C, 512, 0, c->stream>>>(latent_dev, n, s, s + C);
  channel_stats_kernel<<<C, 512, 0, c->stream>>>(ref_dev, n_ref, s + 2 * C, s + 3 * C);
  adain_apply_kernel<<<dim3(grid_for(n, 256, 64), C), 256, 0, c->stream>>>(latent_dev, n, s, s + C, s + 2 * C, s + 3 * C, factor);
  LTX_CUDA(cudaGetLastError());
}

void renoise_dev(ltx_ctx* c, float* latent_dev, const float* noise_dev, int64_t n, float noise_scale) {
  LTX_CHECK(latent_dev && noise_dev && n > 0, LTX_ERR_INVALID_ARGUMENT, "bad renoise arguments");
  ProfScope ps(c, PROF_OTHER, 0.0, 12.0 * n);
  renoise_kernel<<<grid_for(n), 256, 0, c->stream>>>(latent_dev, noise_dev, n, noise_scale);
  LTX_CUDA(cudaGetLastError());
}

}  // namespace ltx
