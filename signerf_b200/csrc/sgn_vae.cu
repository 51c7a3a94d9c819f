// what A1111's img2img does around the denoising loop for the request of reference
// signerf/diffuser/diffuser.py:132-169 -- the SDXL first-stage autoencoder (its contractions run on K5; the few operators
// the UNet does not have live here) and the integer image work: mask blur, latent mask, overlay compositing.
//   sgn_im2col3x3_s2_asym_f16   ldm Downsample: F.pad(x, (0,1,0,1)) + 3x3 / stride 2 / pad 0
//   sgn_softmax_rows_f16        AttnBlock softmax over 65 536 keys (scores from K5, probabilities back into K5)
//   sgn_pointwise_nchw          quant_conv / post_quant_conv (1x1, <= 16 channels)
//   sgn_vae_sample_latent       DiagonalGaussianDistribution.sample() * scale_factor
//   sgn_u8_to_vae_input / sgn_vae_output_to_u8
//   sgn_gaussian_blur_u8        cv2.GaussianBlur on CV_8U, bit-exact (Q8 taps with error diffusion, REFLECT_101)
//   sgn_pil_resize_bicubic_u8   PIL.Image.resize(BICUBIC) on an 8-bit band, bit-exact (22-bit fixed-point taps)
//   sgn_inpaint_masks_u8 / sgn_overlay_composite_u8   mask_for_overlay, latent mask, PIL paste + alpha_composite
// Everything integer here is bit-exact against cv2 4.13 / Pillow 12.2 (tests/golden/inpaint.npz).
#include <algorithm>
#include <cmath>
#include <vector>

#include "sgn_common.cuh"

namespace sgn {

static inline int grid_1d_v(size_t n, int block, int per_sm = 8) {
  size_t want = (n + block - 1) / block;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sm_count() * per_sm));
}
#define STV(s) reinterpret_cast<cudaStream_t>(s)

// ------------------------------------------------------------------ asymmetric stride-2 im2col
// x [B,H,W,C] fp32 -> out [B*Ho*Wo, 9*C] fp16, k = (ky*3+kx)*C + c, input pixel (2*oy + ky, 2*ox + kx); the single
// row / column of padding sits at the bottom / right.
__device__ __forceinline__ void split4(float4 v, uint2& hi, uint2& lo) {
  const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
  const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
  const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
  hi.x = *reinterpret_cast<const uint32_t*>(&h0), hi.y = *reinterpret_cast<const uint32_t*>(&h1);
  lo.x = *reinterpret_cast<const uint32_t*>(&l0), lo.y = *reinterpret_cast<const uint32_t*>(&l1);
}

// split: row = [hi(9C) | lo(9C)]
__global__ void k_im2col_s2_asym(const float* __restrict__ x, int B, int H, int W, int c4, int Ho, int Wo, int split,
                                 __half* out) {
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
    int iy = 2 * oy + tap / 3, ix = 2 * ox + tap % 3;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy < H && ix < W) v = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)b * H + iy) * W + ix) * c4 + c);
    uint2 hi, lo;
    split4(v, hi, lo);
    if (!split) {
      reinterpret_cast<uint2*>(out)[i] = hi;
    } else {
      const size_t row = i / (9 * (size_t)c4), k = i % (9 * (size_t)c4);
      reinterpret_cast<uint2*>(out)[row * 18 * c4 + k] = hi;
      reinterpret_cast<uint2*>(out)[row * 18 * c4 + 9 * c4 + k] = lo;
    }
  }
}

// fp32 [M, C] -> fp16 [M, 2C] = [hi | lo]
__global__ void k_split_f16(const float* __restrict__ x, size_t M, int c4, __half* out) {
  const size_t n = M * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint2 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>(x) + i), hi, lo);
    const size_t o = (i / c4) * (size_t)(2 * c4) + (i % c4);
    reinterpret_cast<uint2*>(out)[o] = hi;
    reinterpret_cast<uint2*>(out)[o + c4] = lo;
  }
}

// nearest x2 upsample of fp32 NHWC -> fp16 [B, 2H, 2W, 2C] = [hi | lo]
__global__ void k_upsample2x_split(const float* __restrict__ x, int B, int H, int W, int c4, __half* out) {
  const size_t n = (size_t)B * 2 * H * 2 * W * c4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % c4);
    size_t r = i / c4;
    const int ox = (int)(r % (2 * W));
    r /= 2 * W;
    const int oy = (int)(r % (2 * H));
    const int b = (int)(r / (2 * H));
    uint2 hi, lo;
    split4(__ldg(reinterpret_cast<const float4*>(x) + (((size_t)b * H + oy / 2) * W + ox / 2) * c4 + c), hi, lo);
    const size_t o = (i / c4) * (size_t)(2 * c4) + c;
    reinterpret_cast<uint2*>(out)[o] = hi;
    reinterpret_cast<uint2*>(out)[o + c4] = lo;
  }
}

// ------------------------------------------------------------------ row softmax
// One 256-thread block per row.  Pass 1: per-thread online (max, sum) over float4 loads, block reduction in a fixed
// order; pass 2 re-reads the row (an L2 hit: a 65 536-key row is 256 KB) and writes fp16 probabilities.
__global__ void __launch_bounds__(256) k_softmax_rows(const float* __restrict__ s, long long M, int N, float scale_log2e,
                                                      __half* __restrict__ out) {
  __shared__ float sm_m[8], sm_l[8];
  const int n4 = N >> 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long row = blockIdx.x; row < M; row += gridDim.x) {
    const float4* src = reinterpret_cast<const float4*>(s + row * (long long)N);
    float m = -INFINITY, l = 0.f;
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 v = src[i];
      float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)) * scale_log2e;
      if (mx > m) {
        l *= exp2f(m - mx);
        m = mx;
      }
      l += exp2f(v.x * scale_log2e - m) + exp2f(v.y * scale_log2e - m) + exp2f(v.z * scale_log2e - m) +
           exp2f(v.w * scale_log2e - m);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
      float mn = fmaxf(m, m2);
      l = (m == -INFINITY ? 0.f : l * exp2f(m - mn)) + (m2 == -INFINITY ? 0.f : l2 * exp2f(m2 - mn));
      m = mn;
    }
    if (lane == 0) sm_m[warp] = m, sm_l[warp] = l;
    __syncthreads();
    float M_ = sm_m[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) M_ = fmaxf(M_, sm_m[w]);
    float L_ = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) L_ += sm_m[w] == -INFINITY ? 0.f : sm_l[w] * exp2f(sm_m[w] - M_);
    const float inv = 1.f / L_;
    uint2* dst = reinterpret_cast<uint2*>(out + row * (long long)N);
    for (int i = threadIdx.x; i < n4; i += 256) {
      float4 v = src[i];
      __half2 h0 = __floats2half2_rn(exp2f(v.x * scale_log2e - M_) * inv, exp2f(v.y * scale_log2e - M_) * inv);
      __half2 h1 = __floats2half2_rn(exp2f(v.z * scale_log2e - M_) * inv, exp2f(v.w * scale_log2e - M_) * inv);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      dst[i] = u;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ 1x1 conv on a few channels, NCHW
constexpr int kPwMax = 16;
struct PointwiseParams {
  float w[kPwMax * kPwMax];
  float bias[kPwMax];
};
__global__ void k_pointwise_nchw(const float* __restrict__ x, const __grid_constant__ PointwiseParams pp, int B, int Cin,
                                 int Cout, long long HW, float in_scale, float* __restrict__ out) {
  const long long n = (long long)B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / HW, p = i - b * HW;
    float v[kPwMax];
    for (int c = 0; c < Cin; ++c) v[c] = in_scale * x[(b * Cin + c) * HW + p];
    for (int o = 0; o < Cout; ++o) {
      float acc = pp.bias[o];
      for (int c = 0; c < Cin; ++c) acc = fmaf(pp.w[o * Cin + c], v[c], acc);
      out[(b * Cout + o) * HW + p] = acc;
    }
  }
}

__global__ void k_vae_sample(const float* __restrict__ mom, const float* __restrict__ noise, int B, int Z, long long HW,
                             float scale, float* __restrict__ out) {
  const long long n = (long long)B * Z * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / (Z * HW), r = i - b * Z * HW;
    const float mean = mom[b * 2 * Z * HW + r];
    float z = mean;
    if (noise) {
      const float logvar = fminf(fmaxf(mom[b * 2 * Z * HW + Z * HW + r], -30.f), 20.f);
      z = mean + expf(0.5f * logvar) * noise[i];
    }
    out[i] = scale * z;
  }
}

__global__ void k_u8_to_vae_input(const uint8_t* __restrict__ img, long long HW, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) out[c * HW + i] = 2.f * ((float)img[i * 3 + c] / 255.f) - 1.f;
  }
}

__global__ void k_vae_output_to_u8(const float* __restrict__ x, long long HW, uint8_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = __fdiv_rn(__fadd_rn(x[c * HW + i], 1.f), 2.f);
      v = fminf(fmaxf(v, 0.f), 1.f);
      out[i * 3 + c] = (uint8_t)__fmul_rn(255.f, v);   // astype(uint8): truncation
    }
  }
}

// ------------------------------------------------------------------ cv2.GaussianBlur, CV_8U, one axis
constexpr int kMaxTaps = 63;
struct BlurTaps {
  int n;
  int q8[kMaxTaps];
};
__device__ __forceinline__ int reflect101(int i, int n) {
  if (n == 1) return 0;
  const int p = 2 * (n - 1);
  i = (i < 0 ? -i : i) % p;
  return i >= n ? p - i : i;
}
__global__ void k_gaussian_blur_u8(const uint8_t* __restrict__ in, int H, int W, const __grid_constant__ BlurTaps t,
                                   int horizontal, uint8_t* __restrict__ out) {
  const long long n = (long long)H * W;
  const int r = t.n >> 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / W), x = (int)(i - (long long)y * W);
    int acc = 0;
    if (horizontal) {
      const uint8_t* row = in + (long long)y * W;
      for (int k = 0; k < t.n; ++k) acc += t.q8[k] * (int)row[reflect101(x + k - r, W)];
    } else {
      for (int k = 0; k < t.n; ++k) acc += t.q8[k] * (int)in[(long long)reflect101(y + k - r, H) * W + x];
    }
    out[i] = (uint8_t)((acc + 128) >> 8);
  }
}

// ------------------------------------------------------------------ PIL bicubic resample of one axis
// coefficient tables (host-built like Resample.c precompute_coeffs) in the workspace: bounds [n_out][2], kk [n_out][ksize]
__global__ void k_pil_resample(const uint8_t* __restrict__ in, int H, int W, int horizontal, int n_out, int ksize,
                               const int* __restrict__ bounds, const int* __restrict__ kk, uint8_t* __restrict__ out) {
  const int Ho = horizontal ? H : n_out, Wo = horizontal ? n_out : W;
  const long long n = (long long)Ho * Wo;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / Wo), x = (int)(i - (long long)y * Wo);
    const int o = horizontal ? x : y;
    const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
    const int* k = kk + (long long)o * ksize;
    long long ss = 1ll << 21;   // 1 << (PRECISION_BITS - 1), PRECISION_BITS = 22
    if (horizontal) {
      const uint8_t* row = in + (long long)y * W + lo;
      for (int j = 0; j < cnt; ++j) ss += (long long)k[j] * row[j];
    } else {
      for (int j = 0; j < cnt; ++j) ss += (long long)k[j] * in[(long long)(lo + j) * W + x];
    }
    long long v = ss >> 22;
    out[i] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
  }
}

__global__ void k_inpaint_masks(const uint8_t* __restrict__ blurred, long long n, uint8_t* __restrict__ overlay) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int v = 2 * (int)blurred[i];
    overlay[i] = (uint8_t)(v > 255 ? 255 : v);
  }
}

__global__ void k_latent_keep_mask(const uint8_t* __restrict__ lat_u8, long long n, float* __restrict__ keep) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    keep[i] = lat_u8[i] >= 128 ? 0.f : 1.f;   // np.around(v / 255) = 1 -> repaint; sgn_cfg_euler_step wants 1 = keep
}

__device__ __forceinline__ int muldiv255(int a, int b) {
  int t = a * b + 128;
  return ((t >> 8) + t) >> 8;
}
__global__ void k_overlay_composite(const uint8_t* __restrict__ gen, const uint8_t* __restrict__ orig,
                                    const uint8_t* __restrict__ overlay_mask, long long HW, uint8_t* __restrict__ out_u8,
                                    float* __restrict__ out_f32) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
    const int a = 255 - (int)overlay_mask[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int g = gen[i * 3 + c];
      int o = g;
      if (a != 0) {
        const int pre = muldiv255((int)orig[i * 3 + c], a);
        int unp = pre;
        if (a != 255) unp = min(255, (255 * pre) / a);
        const unsigned tmp = (unsigned)(unp * a + g * (255 - a)) * 128u + (0x80u << 7);
        o = (int)((((tmp >> 8) + tmp) >> 8) >> 7);
      }
      if (out_u8) out_u8[i * 3 + c] = (uint8_t)o;
      if (out_f32) out_f32[i * 3 + c] = (float)o / 255.f;   // image_to_tensor (image_tensor_converter.py)
    }
  }
}

// host: OpenCV getGaussianKernelBitExact + getGaussianKernelFixedPoint_ED (Q8, error diffusion, centre = remainder)
static bool gaussian_taps_q8(int ksize, double sigma, BlurTaps& t) {
  if (ksize < 1 || ksize > kMaxTaps || (ksize & 1) == 0 || !(sigma > 0.0)) return false;
  t.n = ksize;
  if (ksize == 1) {
    t.q8[0] = 256;
    return true;
  }
  const int r = ksize / 2;
  std::vector<double> k(ksize);
  const double scale2x = -0.5 / (sigma * sigma);
  double sum = 0.0;
  for (int i = 0; i < ksize; ++i) {
    const double x = i - r;
    k[i] = std::exp(scale2x * x * x);
    sum += k[i];
  }
  double err = 0.0;
  long long total = 0;
  for (int i = 0; i < r; ++i) {
    const double adj = k[i] / sum * 256.0 + err;
    const long long v = (long long)std::nearbyint(adj);   // cvRound: round half to even
    err = adj - (double)v;
    t.q8[i] = t.q8[ksize - 1 - i] = (int)v;
    total += 2 * v;
  }
  t.q8[r] = (int)(256 - total);
  return true;
}

static double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

// Pillow Resample.c precompute_coeffs + normalize_coeffs_8bpc
static int pil_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk) {
  double support = 2.0, scale, filterscale;
  filterscale = scale = (double)in_size / out_size;
  if (filterscale < 1.0) filterscale = 1.0;
  support *= filterscale;
  const int ksize = (int)std::ceil(support) * 2 + 1;
  bounds.assign((size_t)out_size * 2, 0);
  kk.assign((size_t)out_size * ksize, 0);
  std::vector<double> w(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale, ss = 1.0 / filterscale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      w[x] = bicubic_filter((x + xmin - center + 0.5) * ss);
      ww += w[x];
    }
    for (int x = 0; x < xmax; ++x) {
      const double v = ww != 0.0 ? w[x] / ww : w[x];
      kk[(size_t)xx * ksize + x] = v < 0 ? (int)(-0.5 + v * (1 << 22)) : (int)(0.5 + v * (1 << 22));
    }
    bounds[2 * xx] = xmin, bounds[2 * xx + 1] = xmax;
  }
  return ksize;
}

static int64_t pil_ws_ints(int in_size, int out_size) {
  double fs = std::max(1.0, (double)in_size / out_size);
  const int ksize = (int)std::ceil(2.0 * fs) * 2 + 1;
  return (int64_t)out_size * (2 + ksize);
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_split_f16(const float* d_x, int64_t M, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(M >= 0 && C > 0 && C % 4 == 0, "bad shape (C % 4)");
  if (M == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_split_f16<<<grid_1d_v((size_t)M * (C / 4), 256), 256, 0, STV(stream)>>>(d_x, (size_t)M, C / 4, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_upsample2x_split_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && C % 4 == 0, "bad shape (C % 4)");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_upsample2x_split<<<grid_1d_v((size_t)B * 4 * H * W * (C / 4), 256), 256, 0, STV(stream)>>>(d_x, B, H, W, C / 4,
                                                                                              reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_im2col3x3_s2_asym_split_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H >= 2 && W >= 2 && C > 0 && C % 4 == 0, "bad shape (C % 4, H, W >= 2)");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
  const size_t n = (size_t)B * Ho * Wo * 9 * (C / 4);
  k_im2col_s2_asym<<<grid_1d_v(n, 256), 256, 0, STV(stream)>>>(d_x, B, H, W, C / 4, Ho, Wo, 1, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_im2col3x3_s2_asym_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H >= 2 && W >= 2 && C > 0 && C % 4 == 0, "bad shape (C % 4, H, W >= 2)");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  const int Ho = (H - 2) / 2 + 1, Wo = (W - 2) / 2 + 1;
  const size_t n = (size_t)B * Ho * Wo * 9 * (C / 4);
  k_im2col_s2_asym<<<grid_1d_v(n, 256), 256, 0, STV(stream)>>>(d_x, B, H, W, C / 4, Ho, Wo, 0, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_softmax_rows_f16(const float* d_scores, int64_t M, int N, float scale, void* d_out, void* stream) {
  SGN_CHECK_ARG(M >= 0 && N > 0 && N % 4 == 0, "bad softmax shape (N % 4)");
  if (M == 0) return SGN_OK;
  SGN_CHECK_ARG(d_scores && d_out, "null pointer");
  const int grid = (int)std::min<int64_t>(M, (int64_t)sm_count() * 8);
  k_softmax_rows<<<grid, 256, 0, STV(stream)>>>(d_scores, M, N, scale * 1.4426950408889634f, reinterpret_cast<__half*>(d_out));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_pointwise_nchw(const float* d_x, const float* h_w, const float* h_bias, int B, int Cin, int Cout,
                                  int64_t HW, float in_scale, float* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && Cin >= 1 && Cin <= kPwMax && Cout >= 1 && Cout <= kPwMax && HW > 0, "pointwise conv handles <= 16 channels");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && h_w && d_out, "null pointer");
  PointwiseParams pp = {};
  for (int i = 0; i < Cin * Cout; ++i) pp.w[i] = h_w[i];
  for (int i = 0; i < Cout; ++i) pp.bias[i] = h_bias ? h_bias[i] : 0.f;
  k_pointwise_nchw<<<grid_1d_v((size_t)B * HW, 256), 256, 0, STV(stream)>>>(d_x, pp, B, Cin, Cout, HW, in_scale, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_vae_sample_latent(const float* d_moments, const float* d_noise, int B, int Z, int64_t HW, float scale,
                                     float* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && Z > 0 && HW > 0, "bad latent shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_moments && d_out, "null pointer");
  k_vae_sample<<<grid_1d_v((size_t)B * Z * HW, 256), 256, 0, STV(stream)>>>(d_moments, d_noise, B, Z, HW, scale, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_u8_to_vae_input(const uint8_t* d_img, int H, int W, float* d_out, void* stream) {
  SGN_CHECK_ARG(H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(d_img && d_out, "null pointer");
  k_u8_to_vae_input<<<grid_1d_v((size_t)H * W, 256), 256, 0, STV(stream)>>>(d_img, (long long)H * W, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_vae_output_to_u8(const float* d_x, int H, int W, uint8_t* d_out, void* stream) {
  SGN_CHECK_ARG(H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(d_x && d_out, "null pointer");
  k_vae_output_to_u8<<<grid_1d_v((size_t)H * W, 256), 256, 0, STV(stream)>>>(d_x, (long long)H * W, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_gaussian_kernel_q8(int ksize, double sigma, int* h_taps) {
  SGN_CHECK_ARG(h_taps != nullptr, "null pointer");
  BlurTaps t;
  SGN_CHECK_ARG(gaussian_taps_q8(ksize, sigma, t), "ksize must be odd, 1..63, sigma > 0");
  for (int i = 0; i < ksize; ++i) h_taps[i] = t.q8[i];
  return SGN_OK;
}

extern "C" int sgn_gaussian_blur_u8(const uint8_t* d_in, int H, int W, int ksize, double sigma, int horizontal,
                                    uint8_t* d_out, void* stream) {
  SGN_CHECK_ARG(H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(d_in && d_out && d_in != d_out, "null or aliased pointer");
  BlurTaps t;
  SGN_CHECK_ARG(gaussian_taps_q8(ksize, sigma, t), "ksize must be odd, 1..63, sigma > 0");
  k_gaussian_blur_u8<<<grid_1d_v((size_t)H * W, 256), 256, 0, STV(stream)>>>(d_in, H, W, t, horizontal, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int64_t sgn_pil_resize_ws_bytes(int H, int W, int h, int w) {
  if (H <= 0 || W <= 0 || h <= 0 || w <= 0) return 0;
  return (int64_t)sizeof(int) * (pil_ws_ints(W, w) + pil_ws_ints(H, h)) + (int64_t)H * w + 64;
}

extern "C" int sgn_pil_resize_bicubic_u8(const uint8_t* d_in, int H, int W, int h, int w, void* d_ws, uint8_t* d_out,
                                         void* stream) {
  SGN_CHECK_ARG(H > 0 && W > 0 && h > 0 && w > 0, "bad resize shape");
  SGN_CHECK_ARG(d_in && d_ws && d_out, "null pointer");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_ws) & 3) == 0, "workspace must be 4-byte aligned");
  cudaStream_t st = STV(stream);
  int* ws_i = reinterpret_cast<int*>(d_ws);
  const int64_t n_h = pil_ws_ints(W, w), n_v = pil_ws_ints(H, h);
  uint8_t* tmp = reinterpret_cast<uint8_t*>(ws_i + n_h + n_v);
  const uint8_t* cur = d_in;
  int curW = W;
  std::vector<int> bounds, kk;
  if (w != W) {
    const int ks = pil_coeffs(W, w, bounds, kk);
    SGN_CUDA(cudaMemcpyAsync(ws_i, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SGN_CUDA(cudaMemcpyAsync(ws_i + bounds.size(), kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    uint8_t* dst = (h != H) ? tmp : d_out;
    k_pil_resample<<<grid_1d_v((size_t)H * w, 256), 256, 0, st>>>(cur, H, W, 1, w, ks, ws_i, ws_i + bounds.size(), dst);
    SGN_LAUNCH_CHECK();
    cur = dst, curW = w;
  }
  if (h != H) {
    const int ks = pil_coeffs(H, h, bounds, kk);
    int* base = ws_i + n_h;
    SGN_CUDA(cudaMemcpyAsync(base, bounds.data(), bounds.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    SGN_CUDA(cudaMemcpyAsync(base + bounds.size(), kk.data(), kk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    k_pil_resample<<<grid_1d_v((size_t)h * curW, 256), 256, 0, st>>>(cur, H, curW, 0, h, ks, base, base + bounds.size(), d_out);
    SGN_LAUNCH_CHECK();
  } else if (w == W) {
    SGN_CUDA(cudaMemcpyAsync(d_out, d_in, (size_t)H * W, cudaMemcpyDeviceToDevice, st));
  }
  return SGN_OK;
}

extern "C" int sgn_inpaint_overlay_mask_u8(const uint8_t* d_blurred, int64_t n, uint8_t* d_overlay, void* stream) {
  SGN_CHECK_ARG(n >= 0, "bad size");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_blurred && d_overlay, "null pointer");
  k_inpaint_masks<<<grid_1d_v((size_t)n, 256), 256, 0, STV(stream)>>>(d_blurred, n, d_overlay);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_latent_keep_mask(const uint8_t* d_lat_u8, int64_t n, float* d_keep, void* stream) {
  SGN_CHECK_ARG(n >= 0, "bad size");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_lat_u8 && d_keep, "null pointer");
  k_latent_keep_mask<<<grid_1d_v((size_t)n, 256), 256, 0, STV(stream)>>>(d_lat_u8, n, d_keep);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_overlay_composite_u8(const uint8_t* d_generated, const uint8_t* d_original, const uint8_t* d_overlay_mask,
                                        int H, int W, uint8_t* d_out_u8, float* d_out_f32, void* stream) {
  SGN_CHECK_ARG(H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(d_generated && d_original && d_overlay_mask && (d_out_u8 || d_out_f32), "null pointer");
  k_overlay_composite<<<grid_1d_v((size_t)H * W, 256), 256, 0, STV(stream)>>>(d_generated, d_original, d_overlay_mask,
                                                                             (long long)H * W, d_out_u8, d_out_f32);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

// ------------------------------------------------------------------ small-channel 3x3 conv on mma.sync, fp32-exact
// ControlNet's input_hint_block at the sheet resolution (16 -> 16 @2048^2, 32 -> 32 @1024^2): too narrow for the
// TMA / tcgen05 conv (Cin % 64) and, as a direct fp32 conv, 1.7 ms each on the CUDA cores.  Here a 16 x 16 pixel tile
// (+halo) is split once into fp16 hi + lo halves in shared memory (x = hi + lo to 2^-22), and every tap is two
// m16n8k16 MMAs (hi, lo) against the fp16 weights with fp32 accumulation: the fp32 convolution to fp32 rounding for
// fp16-representable weights, the same trick as sgn_im2col3x3_split_f16 without materialising the im2col matrix.
namespace sgn {

__device__ __forceinline__ void sc_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t sc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sc_ldsm_x4(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(sc_smem(p)));
}
__device__ __forceinline__ void sc_ldsm_x2(uint32_t (&r)[2], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(sc_smem(p)));
}

constexpr int kScTile = 16;

template <int CIN, int COUT, int STRIDE>
struct ScCfg {
  static constexpr int kHalo = (kScTile - 1) * STRIDE + 3;   // input tile edge: 18 (stride 1) / 33 (stride 2)
  static constexpr int kPix = CIN + 8;        // halfs per pixel in shared memory (48 / 80 B: conflict-free ldmatrix rows)
  static constexpr int kWRow = 9 * CIN + 8;   // halfs per weight row
  static constexpr size_t kSmem = (size_t)(2 * kHalo * kHalo * kPix + COUT * kWRow) * sizeof(__half);
};

// H, W: input size; Ho, Wo: output size ((H-1)/STRIDE + 1: pad 1)
template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(256)
k_conv3x3_small_tc(const float* __restrict__ x, const __half* __restrict__ w, const float* __restrict__ bias, int H, int W,
                   int Ho, int Wo, int act_silu, int out_f16, void* __restrict__ out) {
  using Cfg = ScCfg<CIN, COUT, STRIDE>;
  constexpr int kScHalo = Cfg::kHalo;
  extern __shared__ __align__(16) uint8_t sc_raw[];
  __half* s_hi = reinterpret_cast<__half*>(sc_raw);
  __half* s_lo = s_hi + kScHalo * kScHalo * Cfg::kPix;
  __half* s_w = s_lo + kScHalo * kScHalo * Cfg::kPix;
  const int tiles_x = (Wo + kScTile - 1) / kScTile;
  const int tx0 = (blockIdx.x % tiles_x) * kScTile, ty0 = (blockIdx.x / tiles_x) * kScTile;   // output coordinates
  const int b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // weights [COUT][9*CIN] fp16 -> padded rows (8 halfs = 16 B per copy)
  for (int i = tid; i < COUT * (9 * CIN / 8); i += 256) {
    const int r = i / (9 * CIN / 8), j = i % (9 * CIN / 8);
    *reinterpret_cast<uint4*>(s_w + r * Cfg::kWRow + j * 8) = __ldg(reinterpret_cast<const uint4*>(w + (size_t)r * 9 * CIN) + j);
  }
  // input halo tile: fp32 NHWC -> hi / lo fp16
  constexpr int c4n = CIN / 4;
  for (int i = tid; i < kScHalo * kScHalo * c4n; i += 256) {
    const int c4 = i % c4n, pix = i / c4n;
    const int py = pix / kScHalo, px = pix % kScHalo;
    const int iy = ty0 * STRIDE + py - 1, ix = tx0 * STRIDE + px - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const float4*>(x + (((size_t)b * H + iy) * W + ix) * CIN) + c4);
    const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    const __half2 l0 = __floats2half2_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    uint2 uh, ul;
    uh.x = *reinterpret_cast<const uint32_t*>(&h0), uh.y = *reinterpret_cast<const uint32_t*>(&h1);
    ul.x = *reinterpret_cast<const uint32_t*>(&l0), ul.y = *reinterpret_cast<const uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(s_hi + pix * Cfg::kPix + c4 * 4) = uh;
    *reinterpret_cast<uint2*>(s_lo + pix * Cfg::kPix + c4 * 4) = ul;
  }
  __syncthreads();
  // warp -> output rows 2*warp, 2*warp+1 of the tile (one m16 tile = 16 pixels of a row); all COUT channels
  float acc[2][COUT / 8][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < COUT / 8; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int dy = tap / 3, dx = tap % 3;
#pragma unroll
    for (int kc = 0; kc < CIN / 16; ++kc) {
      uint32_t bf[COUT / 8][2];
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt)
        sc_ldsm_x2(bf[nt], s_w + (nt * 8 + (lane & 7)) * Cfg::kWRow + tap * CIN + kc * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int pix = ((warp * 2 + mt) * STRIDE + dy) * kScHalo + dx + (lane & 15) * STRIDE;
        uint32_t ah[4], al[4];
        sc_ldsm_x4(ah, s_hi + pix * Cfg::kPix + kc * 16 + (lane >> 4) * 8);
        sc_ldsm_x4(al, s_lo + pix * Cfg::kPix + kc * 16 + (lane >> 4) * 8);
#pragma unroll
        for (int nt = 0; nt < COUT / 8; ++nt) {
          sc_mma(acc[mt][nt], ah, bf[nt]);
          sc_mma(acc[mt][nt], al, bf[nt]);
        }
      }
    }
  }
  const int g = lane >> 2, c0 = (lane & 3) * 2;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    const int oy = ty0 + warp * 2 + mt;
    if (oy >= Ho) continue;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int ox = tx0 + g + half * 8;
      if (ox >= Wo) continue;
      const size_t o = (((size_t)b * Ho + oy) * Wo + ox) * COUT;
#pragma unroll
      for (int nt = 0; nt < COUT / 8; ++nt) {
        float v0 = acc[mt][nt][half * 2] + (bias ? __ldg(bias + nt * 8 + c0) : 0.f);
        float v1 = acc[mt][nt][half * 2 + 1] + (bias ? __ldg(bias + nt * 8 + c0 + 1) : 0.f);
        if (act_silu) v0 = v0 / (1.f + __expf(-v0)), v1 = v1 / (1.f + __expf(-v1));
        if (out_f16) {
          const __half2 h = __floats2half2_rn(v0, v1);
          *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(out) + o + nt * 8 + c0) = h;
        } else {
          *reinterpret_cast<float2*>(reinterpret_cast<float*>(out) + o + nt * 8 + c0) = make_float2(v0, v1);
        }
      }
    }
  }
}

template <int CIN, int COUT, int STRIDE>
static int launch_small_tc(const float* x, const void* w, const float* bias, int B, int H, int W, int act_silu, int out_f16,
                           void* out, cudaStream_t st) {
  using Cfg = ScCfg<CIN, COUT, STRIDE>;
  static bool attr_set = false;
  if (!attr_set) {
    SGN_CUDA(cudaFuncSetAttribute(k_conv3x3_small_tc<CIN, COUT, STRIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::kSmem));
    attr_set = true;
  }
  const int Ho = (H - 1) / STRIDE + 1, Wo = (W - 1) / STRIDE + 1;
  dim3 grid(((Wo + kScTile - 1) / kScTile) * ((Ho + kScTile - 1) / kScTile), B);
  k_conv3x3_small_tc<CIN, COUT, STRIDE><<<grid, 256, Cfg::kSmem, st>>>(x, reinterpret_cast<const __half*>(w), bias, H, W, Ho, Wo,
                                                                      act_silu, out_f16, out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

}  // namespace sgn

extern "C" int sgn_conv3x3_small_tc(const float* d_x, const void* d_w16, const float* d_bias, int B, int H, int W, int Cin,
                                    int Cout, int stride, int act_silu, int out_f16, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && (stride == 1 || stride == 2), "bad conv shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_w16 && d_out, "null pointer");
  SGN_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_w16) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0,
                "operands must be 16-byte aligned");
  cudaStream_t st = STV(stream);
  if (stride == 1) {
    if (Cin == 16 && Cout == 16) return sgn::launch_small_tc<16, 16, 1>(d_x, d_w16, d_bias, B, H, W, act_silu, out_f16, d_out, st);
    if (Cin == 16 && Cout == 32) return sgn::launch_small_tc<16, 32, 1>(d_x, d_w16, d_bias, B, H, W, act_silu, out_f16, d_out, st);
    if (Cin == 32 && Cout == 32) return sgn::launch_small_tc<32, 32, 1>(d_x, d_w16, d_bias, B, H, W, act_silu, out_f16, d_out, st);
  } else if (Cin == 16 && Cout == 32) {
    return sgn::launch_small_tc<16, 32, 2>(d_x, d_w16, d_bias, B, H, W, act_silu, out_f16, d_out, st);
  }
  sgn::set_error("invalid argument: sgn_conv3x3_small_tc handles (Cin, Cout) in {(16,16), (16,32), (32,32)} at stride 1 and "
                 "(16,32) at stride 2");
  return SGN_ERR_INVALID_ARG;
}
