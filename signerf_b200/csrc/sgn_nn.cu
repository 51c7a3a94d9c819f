// K7-K9: the bandwidth-bound operators around the tensor-core contractions of the SDXL / ControlNet UNet
// (SURVEY §8a rows A12/A13): GroupNorm(+SiLU), LayerNorm, casts / nearest upsample / concat / stride-2 im2col feeding
// the fp16 GEMM operands, the small-channel direct convolutions (conv_in, ControlNet hint stack), the [B x K] embedding
// linears, the sinusoidal timestep embedding and the fused CFG + inpaint-blend + Euler-ancestral sampler update.
// Layout: NHWC.  The residual stream is fp32, everything handed to a tensor-core kernel is fp16.
#include <algorithm>
#include <cstdlib>

#include "sgn_common.cuh"

namespace sgn {

static inline int grid_1d(size_t n, int block, int per_sm = 8) {
  size_t want = (n + block - 1) / block;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sm_count() * per_sm));
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

// ------------------------------------------------------------------ GroupNorm
// x [B, HW, C] fp32.  Deterministic three-step reduction (no atomics, so CUDA-graph replays are bit-identical):
//   1. per 256-pixel chunk: (sum, sum of squares) per group          -> ws[b][chunk][g] (double2)
//   2. per image: chunks summed in order -> mean, rstd               -> ws tail [b][g] (float2 stored in a double)
//   3. normalise + affine (+ SiLU) -> fp16
constexpr int kGnPix = 256;  // max pixels per block (fewer on small images so that the grid still fills the GPU)
static inline int gn_pix_per_block(int B, int HW) {
  long long want = ((long long)B * HW + 148 * 4 - 1) / (148 * 4);      // ~4 blocks per SM
  int pix = (int)std::min<long long>(kGnPix, std::max<long long>(8, (want + 3) / 4 * 4));
  return pix;
}
__global__ void __launch_bounds__(256) k_gn_stats(const float* __restrict__ x, int HW, int C, int G, int pix_per_block,
                                                  double2* part) {
  extern __shared__ float2 s_part[];  // [4][C/2]
  const int b = blockIdx.y, chunks = gridDim.x;
  const int p0 = blockIdx.x * pix_per_block, p1 = min(HW, p0 + pix_per_block);
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int half_c = C >> 1, cpg2 = (C / G) >> 1;
  const float2* xb = reinterpret_cast<const float2*>(x + ((size_t)b * HW) * C);
  for (int cp = tx; cp < half_c; cp += 64) {
    float s = 0.f, q = 0.f;
    int pix = p0 + ty;
    for (; pix + 28 < p1; pix += 32) {  // 8 independent loads in flight per thread
      float2 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(xb + (size_t)(pix + 4 * u) * half_c + cp);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        s += v[u].x + v[u].y;
        q = fmaf(v[u].x, v[u].x, fmaf(v[u].y, v[u].y, q));
      }
    }
    for (; pix < p1; pix += 4) {
      float2 v = __ldg(xb + (size_t)pix * half_c + cp);
      s += v.x + v.y;
      q = fmaf(v.x, v.x, fmaf(v.y, v.y, q));
    }
    s_part[ty * half_c + cp] = make_float2(s, q);
  }
  __syncthreads();
  if (threadIdx.x < G) {
    double s = 0.0, q = 0.0;
    for (int t = 0; t < 4; ++t)
      for (int cp = threadIdx.x * cpg2; cp < (threadIdx.x + 1) * cpg2; ++cp) {
        float2 v = s_part[t * half_c + cp];
        s += (double)v.x, q += (double)v.y;
      }
    part[((size_t)b * chunks + blockIdx.x) * G + threadIdx.x] = make_double2(s, q);
  }
}

// `tpg` threads per group (32 for the UNet's 32 groups), each sums every tpg-th chunk with four loads in flight; combined in
// a fixed shuffle order (deterministic).  The kernel is pure latency (a 256^2 image has 293 chunks): with 8 threads per
// group and one dependent load at a time it took 11.5 us x 67 launches per step.
__global__ void __launch_bounds__(1024) k_gn_finalize(const double2* __restrict__ part, int chunks, int G, int tpg, double n,
                                                      float eps, float2* stats) {
  const int b = blockIdx.x, g = threadIdx.x / tpg, sub = threadIdx.x % tpg;
  double s0 = 0.0, q0 = 0.0, s1 = 0.0, q1 = 0.0, s2 = 0.0, q2 = 0.0, s3 = 0.0, q3 = 0.0;
  if (g < G) {
    const double2* p = part + (size_t)b * chunks * G + g;
    int c = sub;
    for (; c + 3 * tpg < chunks; c += 4 * tpg) {
      const double2 v0 = p[(size_t)c * G], v1 = p[(size_t)(c + tpg) * G], v2 = p[(size_t)(c + 2 * tpg) * G], v3 = p[(size_t)(c + 3 * tpg) * G];
      s0 += v0.x, q0 += v0.y, s1 += v1.x, q1 += v1.y, s2 += v2.x, q2 += v2.y, s3 += v3.x, q3 += v3.y;
    }
    for (; c < chunks; c += tpg) {
      const double2 v = p[(size_t)c * G];
      s0 += v.x, q0 += v.y;
    }
  }
  double s = (s0 + s1) + (s2 + s3), q = (q0 + q1) + (q2 + q3);
  for (int o = tpg >> 1; o; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (g < G && sub == 0) {
    const double mean = s / n, var = q / n - mean * mean;
    stats[b * G + g] = make_float2((float)mean, (float)(1.0 / sqrt(fmax(var, 0.0) + (double)eps)));
  }
}

// Pass 2: y = (x - mean) * rstd * gamma + beta, optional SiLU, fp16 out.  kSplit: rows of 2C halfs = [hi(y) | lo(y)] with
// y = hi + lo to 2^-22 (operands of the fp32-exact "doubled K" contractions of the VAE).
template <bool kSplit>
__global__ void __launch_bounds__(256) k_gn_apply(const float* __restrict__ x, int HW, int C, int G, float eps,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float2* __restrict__ stats, int act_silu, __half* out) {
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y;
  if (threadIdx.x < G) {
    float2 st = stats[b * G + threadIdx.x];
    s_mean[threadIdx.x] = st.x, s_rstd[threadIdx.x] = st.y;
  }
  __syncthreads();
  const int c4n = C >> 2, cpg = C / G;
  const size_t n4 = (size_t)HW * c4n;
  const float4* xb = reinterpret_cast<const float4*>(x + ((size_t)b * HW) * C);
  uint2* ob = reinterpret_cast<uint2*>(out + ((size_t)b * HW) * C * (kSplit ? 2 : 1));
  auto norm4 = [&](size_t i, float4 v) {
    const int c = 4 * (int)(i % c4n), g0 = c / cpg, g1 = (c + 2) / cpg;  // cpg is even: a pair never straddles groups
    const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c)), bt = __ldg(reinterpret_cast<const float4*>(beta + c));
    float a = (v.x - s_mean[g0]) * s_rstd[g0] * gm.x + bt.x, d = (v.y - s_mean[g0]) * s_rstd[g0] * gm.y + bt.y;
    float e = (v.z - s_mean[g1]) * s_rstd[g1] * gm.z + bt.z, f = (v.w - s_mean[g1]) * s_rstd[g1] * gm.w + bt.w;
    if (act_silu) a = silu(a), d = silu(d), e = silu(e), f = silu(f);
    __half2 h0 = __floats2half2_rn(a, d), h1 = __floats2half2_rn(e, f);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    if (!kSplit) {
      ob[i] = u;
    } else {
      const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
      __half2 l0 = __floats2half2_rn(a - f0.x, d - f0.y), l1 = __floats2half2_rn(e - f1.x, f - f1.y);
      uint2 ul;
      ul.x = *reinterpret_cast<uint32_t*>(&l0);
      ul.y = *reinterpret_cast<uint32_t*>(&l1);
      const size_t o = (i / c4n) * (size_t)(2 * c4n) + (i % c4n);
      ob[o] = u;
      ob[o + c4n] = ul;
    }
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + stride < n4; i += 2 * stride) {  // two independent 16-byte loads in flight per thread
    float4 v0 = __ldg(xb + i), v1 = __ldg(xb + i + stride);
    norm4(i, v0);
    norm4(i + stride, v1);
  }
  if (i < n4) norm4(i, __ldg(xb + i));
}

// ------------------------------------------------------------------ LayerNorm: one warp per token
__global__ void __launch_bounds__(256) k_layer_norm(const float* __restrict__ x, long long M, int C, float eps,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                    __half* out) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c4 = C >> 2;
  for (long long m = warp0; m < M; m += nwarps) {
    const float4* xr = reinterpret_cast<const float4*>(x + m * C);
    float s = 0.f;
    for (int i = lane; i < c4; i += 32) {
      float4 v = xr[i];
      s += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
    for (int i = lane; i < c4; i += 32) {
      float4 v = xr[i];
      float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + m * C);
    for (int i = lane; i < c4; i += 32) {
      float4 v = xr[i];
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + i), bb = __ldg(reinterpret_cast<const float4*>(beta) + i);
      __half2 h0 = __floats2half2_rn((v.x - mean) * rstd * g.x + bb.x, (v.y - mean) * rstd * g.y + bb.y);
      __half2 h1 = __floats2half2_rn((v.z - mean) * rstd * g.z + bb.z, (v.w - mean) * rstd * g.w + bb.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      orow[i] = u;
    }
  }
}

// Row held in registers (C = 128 * NV): one global read, two warp reductions.
template <int NV>
__global__ void __launch_bounds__(256) k_layer_norm_reg(const float* __restrict__ x, long long M, float eps,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        __half* out) {
  constexpr int C = NV * 128;
  const int lane = threadIdx.x & 31;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  // software pipeline over the warp's rows: the next row's loads are in flight while this one is reduced and stored,
  // so the read and the write streams of the launch overlap instead of running as two phases
  float4 v[NV], vn[NV];
  if (warp0 < M) {
    const float4* xr = reinterpret_cast<const float4*>(x + warp0 * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = xr[lane + 32 * i];
  }
  for (long long m = warp0; m < M; m += nwarps) {
    const bool more = m + nwarps < M;
    if (more) {
      const float4* xr = reinterpret_cast<const float4*>(x + (m + nwarps) * C);
#pragma unroll
      for (int i = 0; i < NV; ++i) vn[i] = xr[lane + 32 * i];
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / C + eps);
    uint2* orow = reinterpret_cast<uint2*>(out + m * C);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int j = lane + 32 * i;
      float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + j), bb = __ldg(reinterpret_cast<const float4*>(beta) + j);
      __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
      __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      orow[j] = u;
    }
    if (more) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = vn[i];
    }
  }
}

// ------------------------------------------------------------------ data movement feeding the GEMM operands
__global__ void k_cast_f16(const float* __restrict__ x, size_t n4, __half* out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

// F.interpolate(scale_factor=2, mode="nearest") + fp16 cast: x [B,H,W,C] fp32 -> out [B,2H,2W,C] fp16
__global__ void k_upsample2x_f16(const float* __restrict__ x, int B, int H, int W, int c4, __half* out) {
  const size_t n = (size_t)B * 2 * H * 2 * W * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % c4);
    size_t r = i / c4;
    int ox = (int)(r % (2 * W));
    r /= 2 * W;
    int oy = (int)(r % (2 * H));
    int b = (int)(r / (2 * H));
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)b * H + (oy >> 1)) * W + (ox >> 1)) * c4 + c);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

// torch.cat([a, b + scale * b2], dim=channels): a [P,Ca], b / b2 [P,Cb] -> out [P, Ca+Cb]   (fp32)
// out16 (optional): the same matrix as fp16, for the ResBlock's 1x1 shortcut GEMM (saves re-reading the fp32 result for a cast)
__global__ void k_concat(const float* __restrict__ a, int ca4, const float* __restrict__ b, const float* __restrict__ b2,
                         float scale, int cb4, size_t P, float* out, __half* out16) {
  const int c4 = ca4 + cb4;
  const size_t n = P * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % c4);
    size_t pix = i / c4;
    float4 v;
    if (c < ca4) {
      v = __ldg(reinterpret_cast<const float4*>(a) + pix * ca4 + c);
    } else {
      v = __ldg(reinterpret_cast<const float4*>(b) + pix * cb4 + (c - ca4));
      if (b2) {
        float4 w = __ldg(reinterpret_cast<const float4*>(b2) + pix * cb4 + (c - ca4));
        v.x = fmaf(scale, w.x, v.x), v.y = fmaf(scale, w.y, v.y), v.z = fmaf(scale, w.z, v.z), v.w = fmaf(scale, w.w, v.w);
      }
    }
    reinterpret_cast<float4*>(out)[i] = v;
    if (out16) {
      __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      reinterpret_cast<uint2*>(out16)[i] = u;
    }
  }
}

__global__ void k_axpy(const float* __restrict__ x, float a, size_t n4, float* y) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    float4 w = reinterpret_cast<float4*>(y)[i];
    w.x = fmaf(a, v.x, w.x), w.y = fmaf(a, v.y, w.y), w.z = fmaf(a, v.z, w.z), w.w = fmaf(a, v.w, w.w);
    reinterpret_cast<float4*>(y)[i] = w;
  }
}

// im2col of a 3x3 / stride 2 / pad 1 conv: x [B,H,W,C] fp32 -> out [B*Ho*Wo, 9*C] fp16, k = (ky*3+kx)*C + c
__global__ void k_im2col_s2(const float* __restrict__ x, int B, int H, int W, int c4, int Ho, int Wo, __half* out) {
  const size_t n = (size_t)B * Ho * Wo * 9 * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(i % c4);
    size_t r = i / c4;
    int tap = (int)(r % 9);
    r /= 9;
    int ox = (int)(r % Wo);
    r /= Wo;
    int oy = (int)(r % Ho);
    int b = (int)(r / Ho);
    int iy = 2 * oy + tap / 3 - 1, ix = 2 * ox + tap % 3 - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)b * H + iy) * W + ix) * c4 + c);
    __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0);
    u.y = *reinterpret_cast<uint32_t*>(&h1);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

// im2col of a 3x3 / pad 1 conv with hi/lo fp16 split of the fp32 input: out [M, 2*Kp]; one thread per (pixel, 8 k's)
// -> two 16-byte stores
__global__ void k_im2col_split(const float* __restrict__ x, int in_nchw, int B, int H, int W, int C, int stride, int Ho,
                               int Wo, int Kp, __half* out) {
  const int k8n = Kp >> 3;
  const size_t n = (size_t)B * Ho * Wo * k8n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k0 = (int)(i % k8n) * 8;
    size_t r = i / k8n;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho), b = (int)(r / Ho);
    __half hi[8], lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      float v = 0.f;
      if (k < 9 * C) {
        const int tap = k / C, c = k - tap * C;
        const int iy = stride * oy + tap / 3 - 1, ix = stride * ox + tap % 3 - 1;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
          v = in_nchw ? __ldg(x + (((size_t)b * C + c) * H + iy) * W + ix)
                      : __ldg(x + (((size_t)b * H + iy) * W + ix) * C + c);
      }
      hi[j] = __float2half_rn(v);
      lo[j] = __float2half_rn(v - __half2float(hi[j]));
    }
    __half* row = out + (i / k8n) * (size_t)(2 * Kp);
    *reinterpret_cast<uint4*>(row + k0) = *reinterpret_cast<uint4*>(hi);
    *reinterpret_cast<uint4*>(row + Kp + k0) = *reinterpret_cast<uint4*>(lo);
  }
}

// ------------------------------------------------------------------ direct 3x3 conv for small channel counts (fp32)
// 16x16 output pixels per block, kCoT output channels per block (grid.z), input channels staged through shared
// memory in chunks of kCiT.  w [Cout, 3, 3, Cin] fp32.
constexpr int kDcTile = 16, kCoT = 16, kCiT = 8;
struct DirectConvParams {
  const float* x;
  const float* w;
  const float* bias;
  const float* residual;  // NHWC [B,Ho,Wo,Cout] or null
  void* out;
  int B, H, W, Cin, Cout, stride, Ho, Wo;
  int in_nchw, act_silu, out_f16;
  int res_batch;  // residual holds res_batch images: output image b adds residual image b % res_batch
};

__global__ void __launch_bounds__(256) k_conv3x3_direct(const DirectConvParams p) {
  extern __shared__ float smem_f[];
  const int in_t = (kDcTile - 1) * p.stride + 3;  // input tile edge: 18 (stride 1) or 33 (stride 2)
  float* s_in = smem_f;                             // [kCiT][in_t][in_t]: lanes (adjacent pixels) hit adjacent banks
  float* s_w = smem_f + in_t * in_t * kCiT;         // [9][kCiT][kCoT]
  const int tiles_x = (p.Wo + kDcTile - 1) / kDcTile;
  const int tx0 = (blockIdx.x % tiles_x) * kDcTile, ty0 = (blockIdx.x / tiles_x) * kDcTile;
  const int b = blockIdx.y, co0 = blockIdx.z * kCoT;
  const int lx = threadIdx.x % kDcTile, ly = threadIdx.x / kDcTile;
  const int ox = tx0 + lx, oy = ty0 + ly;
  float acc[kCoT];
#pragma unroll
  for (int i = 0; i < kCoT; ++i) acc[i] = 0.f;
  const int ix0 = tx0 * p.stride - 1, iy0 = ty0 * p.stride - 1;
  for (int ci0 = 0; ci0 < p.Cin; ci0 += kCiT) {
    __syncthreads();
    for (int i = threadIdx.x; i < in_t * in_t * kCiT; i += 256) {
      int ci, px, py;
      if (p.in_nchw) {  // x fastest for coalescing
        px = i % in_t;
        py = (i / in_t) % in_t;
        ci = i / (in_t * in_t);
      } else {
        ci = i % kCiT;
        px = (i / kCiT) % in_t;
        py = i / (kCiT * in_t);
      }
      int iy = iy0 + py, ix = ix0 + px, c = ci0 + ci;
      float v = 0.f;
      if (c < p.Cin && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)
        v = p.in_nchw ? __ldg(p.x + (((size_t)b * p.Cin + c) * p.H + iy) * p.W + ix)
                      : __ldg(p.x + (((size_t)b * p.H + iy) * p.W + ix) * p.Cin + c);
      s_in[(ci * in_t + py) * in_t + px] = v;
    }
    for (int i = threadIdx.x; i < 9 * kCiT * kCoT; i += 256) {
      int co = i % kCoT, ci = (i / kCoT) % kCiT, tap = i / (kCoT * kCiT);
      float v = 0.f;
      if (co0 + co < p.Cout && ci0 + ci < p.Cin) v = __ldg(p.w + ((size_t)(co0 + co) * 9 + tap) * p.Cin + ci0 + ci);
      s_w[i] = v;
    }
    __syncthreads();
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float* ip = s_in + (ly * p.stride + tap / 3) * in_t + lx * p.stride + tap % 3;
      const float* wp = s_w + tap * kCiT * kCoT;
#pragma unroll
      for (int ci = 0; ci < kCiT; ++ci) {
        const float xv = ip[ci * in_t * in_t];
#pragma unroll
        for (int q = 0; q < kCoT / 4; ++q) {
          float4 w4 = *reinterpret_cast<const float4*>(wp + ci * kCoT + 4 * q);
          acc[4 * q] = fmaf(xv, w4.x, acc[4 * q]);
          acc[4 * q + 1] = fmaf(xv, w4.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(xv, w4.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(xv, w4.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (ox >= p.Wo || oy >= p.Ho) return;
  const size_t o = (((size_t)b * p.Ho + oy) * p.Wo + ox) * p.Cout + co0;
#pragma unroll
  for (int i = 0; i < kCoT; ++i) {
    if (co0 + i >= p.Cout) break;
    float v = acc[i] + (p.bias ? __ldg(p.bias + co0 + i) : 0.f);
    if (p.residual) v += p.residual[(((size_t)(b % p.res_batch) * p.Ho + oy) * p.Wo + ox) * p.Cout + co0 + i];
    if (p.act_silu) v = silu(v);
    if (p.out_f16) reinterpret_cast<__half*>(p.out)[o + i] = __float2half_rn(v);
    else reinterpret_cast<float*>(p.out)[o + i] = v;
  }
}

// ------------------------------------------------------------------ embedding path (M = batch rows)
// out[b][n] = sum_k act(x[b][k]) * W[n][k] + bias[n] (+ residual[b][n]); one warp per output feature, B <= 8.
__global__ void __launch_bounds__(256) k_linear_small(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, const float* __restrict__ residual,
                                                      int B, int N, int K, int silu_in, int silu_out, float* out) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  float acc[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) acc[b] = 0.f;
  const float* wr = w + (size_t)n * K;
  if ((K & 3) == 0) {  // 16-byte loads: one weight float4 feeds all rows
    const int k4 = K >> 2;
    for (int k = lane; k < k4; k += 32) {
      const float4 wv = __ldg(reinterpret_cast<const float4*>(wr) + k);
#pragma unroll
      for (int b = 0; b < 8; ++b)
        if (b < B) {
          float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * K) + k);
          if (silu_in) xv.x = silu(xv.x), xv.y = silu(xv.y), xv.z = silu(xv.z), xv.w = silu(xv.w);
          acc[b] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[b]))));
        }
    }
  } else {
    for (int k = lane; k < K; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int b = 0; b < 8; ++b)
        if (b < B) {
          float xv = __ldg(x + (size_t)b * K + k);
          if (silu_in) xv = silu(xv);
          acc[b] = fmaf(xv, wv, acc[b]);
        }
    }
  }
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
  if (lane == 0)
    for (int b = 0; b < B; ++b) {
      float v = acc[b] + (bias ? __ldg(bias + n) : 0.f);
      if (residual) v += residual[(size_t)b * N + n];
      if (silu_out) v = silu(v);
      out[(size_t)b * N + n] = v;
    }
}

// The emb_layers projections of ALL ResBlocks of a network in one launch (they only depend on emb): row n of the
// concatenated weight belongs to segment j (seg[j] <= n < seg[j+1]); segment j's output is its own contiguous [B, N_j]
// block at out + B * seg[j], i.e. exactly what the per-block call would have produced.
__global__ void __launch_bounds__(256) k_linear_small_seg(const float* __restrict__ x, const float* __restrict__ w,
                                                          const float* __restrict__ bias, int B, int N, int K, int silu_in,
                                                          const int* __restrict__ seg, int nseg, float* out) {
  const int lane = threadIdx.x & 31;
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  float acc[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) acc[b] = 0.f;
  const float4* wr = reinterpret_cast<const float4*>(w + (size_t)n * K);
  for (int k = lane; k < (K >> 2); k += 32) {
    const float4 wv = __ldg(wr + k);
#pragma unroll
    for (int b = 0; b < 8; ++b)
      if (b < B) {
        float4 xv = __ldg(reinterpret_cast<const float4*>(x + (size_t)b * K) + k);
        if (silu_in) xv.x = silu(xv.x), xv.y = silu(xv.y), xv.z = silu(xv.z), xv.w = silu(xv.w);
        acc[b] = fmaf(xv.x, wv.x, fmaf(xv.y, wv.y, fmaf(xv.z, wv.z, fmaf(xv.w, wv.w, acc[b]))));   // k_linear_small's order
      }
  }
#pragma unroll
  for (int b = 0; b < 8; ++b)
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[b] += __shfl_xor_sync(0xffffffffu, acc[b], o);
  if (lane == 0) {
    int j = 0;
    while (j + 1 < nseg && __ldg(seg + j + 1) <= n) ++j;
    const int s0 = __ldg(seg + j), nj = __ldg(seg + j + 1) - s0;
    for (int b = 0; b < B; ++b) out[(size_t)B * s0 + (size_t)b * nj + (n - s0)] = acc[b] + (bias ? __ldg(bias + n) : 0.f);
  }
}

// sgm timestep_embedding(t, dim): [cos(t f_i) | sin(t f_i)], f_i = exp(-ln(10000) i / half), fp32 like torch
__global__ void k_timestep_embedding(const float* __restrict__ t, int B, int dim, float* out) {
  const int half_d = dim / 2;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half_d) return;
  int b = i / half_d, j = i % half_d;
  float f = expf(-9.210340371976184f * (float)j / (float)half_d);
  float a = __ldg(t + b) * f;
  out[(size_t)b * dim + j] = cosf(a);
  out[(size_t)b * dim + half_d + j] = sinf(a);
}

// ------------------------------------------------------------------ K9: CFG + inpaint blend + Euler-ancestral update
// eps [2, n] = UNet output for (cond, uncond) on input x * c_in.  k-diffusion CompVisDenoiser: denoised = x - sigma*eps.
//   e      = eps_u + cfg * (eps_c - eps_u)
//   den    = x - sigma * e ;  den = init * mask + nmask * den            (A1111 CFGDenoiser inpaint blend)
//   d      = (x - den) / sigma ;  x' = x + d * (sigma_down - sigma) + noise * sigma_up
__global__ void k_cfg_euler_step(const float* __restrict__ x, const float* __restrict__ eps, const float* __restrict__ init,
                                 const float* __restrict__ mask, const float* __restrict__ noise, size_t n, size_t hw,
                                 int channels, float cfg, float sigma, float sigma_down, float sigma_up, float* x_out,
                                 float* denoised_out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    const float ec = __ldg(eps + i), eu = __ldg(eps + n + i);
    const float e = eu + cfg * (ec - eu);
    float den = xv - sigma * e;
    if (mask) {
      // latent mask [1,1,h,w] broadcast over channels: mask = 1 keeps the original latent (A1111: mask = 1 - latmask)
      const size_t pix = i % hw + (i / (hw * channels)) * hw;
      const float mk = __ldg(mask + pix);
      den = __ldg(init + i) * mk + (1.f - mk) * den;
    }
    if (denoised_out) denoised_out[i] = den;
    const float d = (xv - den) / sigma;
    float xn = xv + d * (sigma_down - sigma);
    if (noise) xn = fmaf(__ldg(noise + i), sigma_up, xn);
    x_out[i] = xn;
  }
}

// out[r] = scale * x for r in [0, repeats): the CFG pair (cond, uncond) shares one scaled latent x * c_in
__global__ void k_scale_repeat(const float* __restrict__ x, size_t n, float scale, int repeats, float* out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = __ldg(x + i) * scale;
    for (int r = 0; r < repeats; ++r) out[(size_t)r * n + i] = v;
  }
}

// Sheet -> denoiser conditioning.  hint [3,Hs,Ws] = uint8-truncated condition / 255 on three channels (the reference
// sends the condition sheet through tensor_to_image -> PNG; ControlNet preprocessor "none" divides by 255);
// lat_mask [Hs/8, Ws/8] = 1 - round(mean of the 8x8 mask block): 1 where the ORIGINAL latent is kept.
__global__ void k_hint_latmask(const float* __restrict__ cond, const float* __restrict__ mask, int Hs, int Ws, float* hint,
                               float* lat_mask) {
  const size_t n = (size_t)Hs * Ws;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float q = (float)(unsigned char)(int)(__ldg(cond + i) * 255.f) * (1.f / 255.f);   // torch CUDA `x / 255.0` = x * (1/255)
    hint[i] = q, hint[n + i] = q, hint[2 * n + i] = q;
    const int h8 = Hs >> 3, w8 = Ws >> 3;
    if (i < (size_t)h8 * w8) {
      const int ly = (int)(i / w8), lx = (int)(i % w8);
      float s = 0.f;
      for (int dy = 0; dy < 8; ++dy)
        for (int dx = 0; dx < 8; ++dx) s += __ldg(mask + (size_t)(ly * 8 + dy) * Ws + lx * 8 + dx);
      lat_mask[i] = 1.f - rintf(s * (1.f / 64.f));
    }
  }
}

}  // namespace sgn

using namespace sgn;

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int64_t sgn_group_norm_ws_doubles(int B, int HW, int groups) {
  const int pix = gn_pix_per_block(B, HW);
  const int64_t chunks = (HW + pix - 1) / pix;
  return (int64_t)B * groups * (2 * chunks + 1);
}

static int group_norm_impl(const float* d_x, int B, int HW, int C, int groups, float eps, const float* d_gamma,
                           const float* d_beta, int act_silu, double* d_ws, void* d_out, void* stream, bool split) {
  SGN_CHECK_ARG(B >= 0 && HW > 0 && C > 0 && groups > 0, "bad shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_gamma && d_beta && d_ws && d_out, "null pointer");
  SGN_CHECK_ARG(groups <= 64 && C % groups == 0 && (C / groups) % 2 == 0 && C % 4 == 0,
                "need groups <= 64, an even number of channels per group and C % 4 == 0");
  SGN_CHECK_ARG(C <= 2560, "GroupNorm kernel stages 4 x C/2 partials in 40 KB of shared memory (C <= 2560)");
  const int pix = gn_pix_per_block(B, HW);
  const int chunks = (HW + pix - 1) / pix;
  double2* part = reinterpret_cast<double2*>(d_ws);
  float2* stats = reinterpret_cast<float2*>(part + (size_t)B * chunks * groups);
  dim3 g1(chunks, B);
  k_gn_stats<<<g1, 256, (size_t)4 * (C / 2) * sizeof(float2), ST(stream)>>>(d_x, HW, C, groups, pix, part);
  SGN_LAUNCH_CHECK();
  int tpg = 32;                                   // threads per group: a power of two, groups * tpg <= 1024
  while (groups * tpg > 1024) tpg >>= 1;
  k_gn_finalize<<<B, ((groups * tpg + 31) / 32) * 32, 0, ST(stream)>>>(part, chunks, groups, tpg, (double)HW * (C / groups), eps, stats);
  SGN_LAUNCH_CHECK();
  size_t n4 = (size_t)HW * (C / 4);
  dim3 g2((unsigned)std::max<size_t>(1, std::min<size_t>((n4 + 511) / 512, (size_t)sm_count() * 8 / std::max(1, B) + 1)), B);
  if (split)
    k_gn_apply<true><<<g2, 256, 0, ST(stream)>>>(d_x, HW, C, groups, eps, d_gamma, d_beta, stats, act_silu,
                                                 reinterpret_cast<__half*>(d_out));
  else
    k_gn_apply<false><<<g2, 256, 0, ST(stream)>>>(d_x, HW, C, groups, eps, d_gamma, d_beta, stats, act_silu,
                                                  reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_group_norm_f16(const float* d_x, int B, int HW, int C, int groups, float eps, const float* d_gamma,
                                  const float* d_beta, int act_silu, double* d_ws, void* d_out, void* stream) {
  return group_norm_impl(d_x, B, HW, C, groups, eps, d_gamma, d_beta, act_silu, d_ws, d_out, stream, false);
}

extern "C" int sgn_group_norm_split_f16(const float* d_x, int B, int HW, int C, int groups, float eps, const float* d_gamma,
                                        const float* d_beta, int act_silu, double* d_ws, void* d_out, void* stream) {
  return group_norm_impl(d_x, B, HW, C, groups, eps, d_gamma, d_beta, act_silu, d_ws, d_out, stream, true);
}

extern "C" int sgn_layer_norm_f16(const float* d_x, int64_t M, int C, float eps, const float* d_gamma,
                                  const float* d_beta, void* d_out, void* stream) {
  SGN_CHECK_ARG(M >= 0 && C > 0 && C % 4 == 0, "C must be a multiple of 4");
  if (M == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_gamma && d_beta && d_out, "null pointer");
  int grid = grid_1d((size_t)M * 32, 256);
  __half* o = reinterpret_cast<__half*>(d_out);
  static const int ln_rows = [] { const char* e = getenv("SGN_LN_ROWS_PER_WARP"); return e ? atoi(e) : 4; }();
  if (C % 128 == 0 && ln_rows > 1)   // register kernels: ~ln_rows rows per warp (pipelined), at least 2 blocks per SM
    grid = std::max(std::min(grid, 2 * sm_count()), std::min(grid, (int)((M + 8 * ln_rows - 1) / (8 * ln_rows))));
  switch (C % 128 == 0 ? C / 128 : 0) {
    case 1: k_layer_norm_reg<1><<<grid, 256, 0, ST(stream)>>>(d_x, M, eps, d_gamma, d_beta, o); break;
    case 2: k_layer_norm_reg<2><<<grid, 256, 0, ST(stream)>>>(d_x, M, eps, d_gamma, d_beta, o); break;
    case 5: k_layer_norm_reg<5><<<grid, 256, 0, ST(stream)>>>(d_x, M, eps, d_gamma, d_beta, o); break;
    case 10: k_layer_norm_reg<10><<<grid, 256, 0, ST(stream)>>>(d_x, M, eps, d_gamma, d_beta, o); break;
    default: k_layer_norm<<<grid, 256, 0, ST(stream)>>>(d_x, M, C, eps, d_gamma, d_beta, o);
  }
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_cast_f16(const float* d_x, int64_t n, void* d_out, void* stream) {
  SGN_CHECK_ARG(n >= 0 && n % 4 == 0, "n must be a multiple of 4");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_cast_f16<<<grid_1d((size_t)n / 4, 256), 256, 0, ST(stream)>>>(d_x, (size_t)n / 4, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_upsample2x_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad shape (C % 4)");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  size_t n = (size_t)B * 4 * H * W * (C / 4);
  k_upsample2x_f16<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(d_x, B, H, W, C / 4, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_concat_f32(const float* d_a, int Ca, const float* d_b, const float* d_b2, float scale, int Cb,
                              int64_t P, float* d_out, void* stream) {
  SGN_CHECK_ARG(P >= 0 && Ca >= 0 && Cb > 0 && Ca % 4 == 0 && Cb % 4 == 0, "channel counts must be multiples of 4");
  if (P == 0) return SGN_OK;
  SGN_CHECK_ARG((Ca == 0 || d_a) && d_b && d_out, "null pointer");
  size_t n = (size_t)P * ((Ca + Cb) / 4);
  k_concat<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(d_a, Ca / 4, d_b, d_b2, scale, Cb / 4, (size_t)P, d_out, nullptr);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_concat_f32_f16(const float* d_a, int Ca, const float* d_b, const float* d_b2, float scale, int Cb,
                                  int64_t P, float* d_out, void* d_out16, void* stream) {
  SGN_CHECK_ARG(P >= 0 && Ca >= 0 && Cb > 0 && Ca % 4 == 0 && Cb % 4 == 0, "channel counts must be multiples of 4");
  if (P == 0) return SGN_OK;
  SGN_CHECK_ARG((Ca == 0 || d_a) && d_b && d_out && d_out16, "null pointer");
  size_t n = (size_t)P * ((Ca + Cb) / 4);
  k_concat<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(d_a, Ca / 4, d_b, d_b2, scale, Cb / 4, (size_t)P, d_out,
                                                    reinterpret_cast<__half*>(d_out16));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_axpy_f32(const float* d_x, float a, int64_t n, float* d_y, void* stream) {
  SGN_CHECK_ARG(n >= 0 && n % 4 == 0, "n must be a multiple of 4");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_y, "null pointer");
  k_axpy<<<grid_1d((size_t)n / 4, 256), 256, 0, ST(stream)>>>(d_x, a, (size_t)n / 4, d_y);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_im2col3x3_s2_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad shape (C % 4)");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
  size_t n = (size_t)B * Ho * Wo * 9 * (C / 4);
  k_im2col_s2<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(d_x, B, H, W, C / 4, Ho, Wo, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_conv3x3_direct(const float* d_x, int in_nchw, const float* d_w, const float* d_bias,
                                  const float* d_residual, int res_batch, int B, int H, int W, int Cin, int Cout,
                                  int stride, int act_silu, int out_f16, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && Cin > 0 && Cout > 0 && (stride == 1 || stride == 2), "bad conv shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_w && d_out, "null pointer");
  DirectConvParams p;
  p.x = d_x, p.w = d_w, p.bias = d_bias, p.residual = d_residual, p.out = d_out;
  p.B = B, p.H = H, p.W = W, p.Cin = Cin, p.Cout = Cout, p.stride = stride;
  p.Ho = (H - 1) / stride + 1, p.Wo = (W - 1) / stride + 1;
  p.in_nchw = in_nchw, p.act_silu = act_silu, p.out_f16 = out_f16;
  p.res_batch = res_batch > 0 ? res_batch : B;
  const int in_t = (kDcTile - 1) * stride + 3;
  size_t smem = ((size_t)in_t * in_t * kCiT + 9 * kCiT * kCoT) * sizeof(float);
  dim3 grid(((p.Wo + kDcTile - 1) / kDcTile) * ((p.Ho + kDcTile - 1) / kDcTile), B, (Cout + kCoT - 1) / kCoT);
  k_conv3x3_direct<<<grid, 256, smem, ST(stream)>>>(p);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_linear_small(const float* d_x, const float* d_w, const float* d_bias, const float* d_residual, int B,
                                int N, int K, int silu_in, int silu_out, float* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 1 && B <= 8 && N > 0 && K > 0, "sgn_linear_small handles 1..8 rows");
  SGN_CHECK_ARG(d_x && d_w && d_out, "null pointer");
  k_linear_small<<<(N * 32 + 255) / 256, 256, 0, ST(stream)>>>(d_x, d_w, d_bias, d_residual, B, N, K, silu_in, silu_out,
                                                               d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_linear_small_segments(const float* d_x, const float* d_w, const float* d_bias, int B, int N, int K, int silu_in,
                                         const int32_t* d_seg_offsets, int num_segments, float* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 1 && B <= 8 && N > 0 && K > 0 && K % 4 == 0 && num_segments >= 1, "sgn_linear_small_segments: 1..8 rows, K % 4 == 0");
  SGN_CHECK_ARG(d_x && d_w && d_seg_offsets && d_out, "null pointer");
  k_linear_small_seg<<<(N * 32 + 255) / 256, 256, 0, ST(stream)>>>(d_x, d_w, d_bias, B, N, K, silu_in, d_seg_offsets, num_segments,
                                                                   d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

// ---- CLIP text encoders (prompt conditioning, once per prompt)
// mode 0: quick_gelu x * sigmoid(1.702 x) (CLIP-L); mode 1: exact GELU (OpenCLIP bigG)
__global__ void k_act_f16(const float* __restrict__ x, size_t n, int mode, __half* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float y = mode == 0 ? v / (1.f + expf(-1.702f * v)) : 0.5f * v * (1.f + erff(v * 0.70710678118654752f));
    out[i] = __float2half_rn(y);
  }
}
// CLIPTextEmbeddings: token_embedding[ids] + position_embedding[position], rows of `width` floats
__global__ void k_embed_tokens(const int* __restrict__ ids, const float* __restrict__ tok, const float* __restrict__ pos,
                               int rows, int T, int width, int vocab, float* __restrict__ out) {
  const size_t n = (size_t)rows * width;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / width), c = (int)(i - (size_t)r * width);
    const int id = min(max(ids[r], 0), vocab - 1);
    out[i] = tok[(size_t)id * width + c] + pos[(size_t)(r % T) * width + c];
  }
}

extern "C" int sgn_act_f16(const float* d_x, int64_t n, int mode, void* d_out, void* stream) {
  SGN_CHECK_ARG(n >= 0 && (mode == 0 || mode == 1), "mode must be 0 (quick_gelu) or 1 (gelu)");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_act_f16<<<grid_1d((size_t)n, 256), 256, 0, ST(stream)>>>(d_x, (size_t)n, mode, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_embed_tokens(const int32_t* d_ids, const float* d_token_table, const float* d_position_table, int rows,
                                int T, int width, int vocab, float* d_out, void* stream) {
  SGN_CHECK_ARG(rows >= 0 && T > 0 && width > 0 && vocab > 0, "bad embedding shape");
  if (rows == 0) return SGN_OK;
  SGN_CHECK_ARG(d_ids && d_token_table && d_position_table && d_out, "null pointer");
  k_embed_tokens<<<grid_1d((size_t)rows * width, 256), 256, 0, ST(stream)>>>(d_ids, d_token_table, d_position_table, rows, T,
                                                                             width, vocab, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_timestep_embedding(const float* d_t, int B, int dim, float* d_out, void* stream) {
  SGN_CHECK_ARG(B > 0 && dim > 0 && dim % 2 == 0, "dim must be even");
  SGN_CHECK_ARG(d_t && d_out, "null pointer");
  int n = B * (dim / 2);
  k_timestep_embedding<<<(n + 127) / 128, 128, 0, ST(stream)>>>(d_t, B, dim, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_cfg_euler_step(const float* d_x, const float* d_eps, const float* d_init, const float* d_mask,
                                  const float* d_noise, int B, int C, int H, int W, float cfg_scale, float sigma,
                                  float sigma_down, float sigma_up, float* d_x_out, float* d_denoised, void* stream) {
  SGN_CHECK_ARG(B > 0 && C > 0 && H > 0 && W > 0, "bad latent shape");
  SGN_CHECK_ARG(d_x && d_eps && d_x_out, "null pointer");
  SGN_CHECK_ARG((d_mask == nullptr) == (d_init == nullptr), "mask and init latent go together");
  SGN_CHECK_ARG(sigma > 0.f, "sigma must be positive");
  size_t n = (size_t)B * C * H * W;
  k_cfg_euler_step<<<grid_1d(n, 256), 256, 0, ST(stream)>>>(d_x, d_eps, d_init, d_mask, d_noise, n, (size_t)H * W, C,
                                                            cfg_scale, sigma, sigma_down, sigma_up, d_x_out, d_denoised);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_scale_repeat_f32(const float* d_x, int64_t n, float scale, int repeats, float* d_out, void* stream) {
  SGN_CHECK_ARG(n >= 0 && repeats >= 1, "bad shape");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_scale_repeat<<<grid_1d((size_t)n, 256), 256, 0, ST(stream)>>>(d_x, (size_t)n, scale, repeats, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_sheet_to_conditioning(const float* d_cond, const float* d_mask, int Hs, int Ws, float* d_hint,
                                         float* d_lat_mask, void* stream) {
  SGN_CHECK_ARG(Hs > 0 && Ws > 0 && Hs % 8 == 0 && Ws % 8 == 0, "sheet size must be a multiple of 8");
  SGN_CHECK_ARG(d_cond && d_mask && d_hint && d_lat_mask, "null pointer");
  k_hint_latmask<<<grid_1d((size_t)Hs * Ws, 256), 256, 0, ST(stream)>>>(d_cond, d_mask, Hs, Ws, d_hint, d_lat_mask);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_im2col3x3_split_f16(const float* d_x, int in_nchw, int B, int H, int W, int C, int stride, void* d_out,
                                       void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && (stride == 1 || stride == 2), "bad shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1, Kp = (9 * C + 7) / 8 * 8;
  const size_t n = (size_t)B * Ho * Wo * (Kp / 8);
  k_im2col_split<<<grid_1d(n, 256, 16), 256, 0, ST(stream)>>>(d_x, in_nchw, B, H, W, C, stride, Ho, Wo, Kp,
                                                              reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}
