// SURVEY §8(f) row 4 - the NeRF fine-tune step that alternates with dataset generation (BASELINE config 5's
// "refinement rounds"): forward + BACKWARD of the main nerfacto field on a batch of rays, the SIGNeRF image loss
// (reference signerf/signerf.py:41-52: L1Loss or MSELoss on rgb) and the Adam update nerfstudio configures for the
// `fields` parameter group (signerf_config.py:43-50).
//
// The batch of a training step is small next to a reference-sheet render (16 patches of 32 x 32 rays x 48 samples =
// 0.8 M samples against 537 M), so this path is organised for exactness and simplicity, fp32 end to end on CUDA cores:
//   k_train_field      one thread per SAMPLE: hash features -> base / head MLP -> sigma, colour (kept for the backward)
//   k_train_composite  one thread per RAY: weights, rgb = sum w c + c_last (1 - sum w)  (training: no clamp)
//   k_train_ray_bwd    one thread per RAY: dL/dsigma_i, dL/dc_i from dL/drgb with the suffix sums of the transmittance
//   k_train_field_bwd  one thread per SAMPLE: recomputes the activations, back-propagates to the hash features
//                      (scatter-add into the table gradient, fp32 vector atomics) and writes the layer deltas /
//                      activations the weight gradients are outer products of
//   k_outer_reduce     dW[n][k] = sum_samples delta[n] * act[k]: skinny GEMM (<= 64 x 64 outputs, the sample axis is the
//                      reduction), 64 x 64 register tile per CTA, one atomicAdd per output per CTA
// Gradients follow torch autograd through nerfstudio's torch modules: trunc_exp's backward clamps its argument at 15,
// ReLU' = (h > 0), the selector multiplies the density only.  The appearance embedding enters as the folded bias of head
// layer 0 (eval semantics: mean embedding), i.e. it is a constant of the step.
#include <algorithm>

#include "sgn_device.cuh"

namespace sgn {

constexpr int kDeltaDim = 212;   // [da0 64 | dout1 16 | da1 64 | da2 64 | dpre 3 | pad 1]
constexpr int kActDim = 256;     // [feat 32 | h0 64 | hin 32 | h1 64 | h2 64]

struct TrainRays {
  const float* origins;
  const float* dirs;
  const float* bins;      // [S+1] shared euclidean edges, or
  const float* ray_bins;  // [N, S+1] per ray
  int64_t N;
  int S;
  const float* head_bias = nullptr;   // [N, 64] per-ray bias of head layer 0 (b + W_app . embedding[camera]), or null
};

__device__ __forceinline__ void sample_interval(const TrainRays& r, int64_t ray, int i, float& t0, float& t1) {
  if (r.ray_bins) {
    t0 = __ldg(r.ray_bins + ray * (r.S + 1) + i);
    t1 = __ldg(r.ray_bins + ray * (r.S + 1) + i + 1);
  } else {
    t0 = __ldg(r.bins + i);
    t1 = __ldg(r.bins + i + 1);
  }
}

__device__ __forceinline__ bool sample_position(const TrainRays& r, int64_t ray, int i, float& px, float& py, float& pz) {
  float t0, t1;
  sample_interval(r, ray, i, t0, t1);
  const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  const float* o = r.origins + 3 * ray;
  const float* d = r.dirs + 3 * ray;
  return contract_to_unit(__fadd_rn(__ldg(o), __fmul_rn(__ldg(d), mid)), __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(d + 1), mid)),
                          __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(d + 2), mid)), px, py, pz);
}

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(128) k_train_field(const GridDev grid, const MlpF32* __restrict__ w32, const TrainRays r,
                                                     float* __restrict__ sigma, float* __restrict__ color) {
  extern __shared__ __align__(16) unsigned char smem[];
  MlpF32* sw = reinterpret_cast<MlpF32*>(smem);
  for (int i = threadIdx.x; i < (int)(sizeof(MlpF32) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(sw)[i] = reinterpret_cast<const uint4*>(w32)[i];
  __syncthreads();
  const int64_t total = r.N * r.S;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ray = s / r.S;
    const int i = (int)(s - ray * r.S);
    float px, py, pz, sh[16], feat[32];
    const bool sel = sample_position(r, ray, i, px, py, pz);
    sh16(__ldg(r.dirs + 3 * ray), __ldg(r.dirs + 3 * ray + 1), __ldg(r.dirs + 3 * ray + 2), sh);
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      float2 f = encode_level(grid.table + (size_t)l * grid.size, grid.mask, grid.res[l], px, py, pz);
      feat[2 * l] = f.x;
      feat[2 * l + 1] = f.y;
    }
    float logit, c3[3];
    field_mlp_f32(sw, feat, sh, logit, c3, r.head_bias ? r.head_bias + 64 * ray : nullptr);
    sigma[s] = sel ? sw->avg_density * expf(logit) : 0.f;
    color[3 * s + 0] = sigmoidf_(c3[0]);
    color[3 * s + 1] = sigmoidf_(c3[1]);
    color[3 * s + 2] = sigmoidf_(c3[2]);
  }
}

__global__ void k_train_composite(const TrainRays r, const float* __restrict__ sigma, const float* __restrict__ color,
                                  float* __restrict__ rgb, float* __restrict__ acc_out) {
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < r.N; ray += (int64_t)gridDim.x * blockDim.x) {
    float cum = 0.f, acc = 0.f, cr = 0.f, cg = 0.f, cb = 0.f;
    const float* sg = sigma + ray * r.S;
    const float* cl = color + 3 * ray * r.S;
    for (int i = 0; i < r.S; ++i) {
      float t0, t1;
      sample_interval(r, ray, i, t0, t1);
      const float dd = __fsub_rn(t1, t0) * sg[i];
      float w = (1.f - expf(-dd)) * expf(-cum);
      cum += dd;
      if (w != w) w = 0.f;
      cr = fmaf(w, cl[3 * i], cr);
      cg = fmaf(w, cl[3 * i + 1], cg);
      cb = fmaf(w, cl[3 * i + 2], cb);
      acc += w;
    }
    const float rem = 1.f - acc;   // RGBRenderer("last_sample"); no clamp while training
    rgb[3 * ray + 0] = fmaf(cl[3 * (r.S - 1)], rem, cr);
    rgb[3 * ray + 1] = fmaf(cl[3 * (r.S - 1) + 1], rem, cg);
    rgb[3 * ray + 2] = fmaf(cl[3 * (r.S - 1) + 2], rem, cb);
    if (acc_out) acc_out[ray] = acc;
  }
}

// ---------------------------------------------------------------- backward, ray level
// C = sum_i w_i (c_i - c_L) + c_L with L the last sample, w_i = (1 - e^{-dd_i}) e^{-sum_{j<i} dd_j}, dd_i = delta_i sigma_i:
//   dL/dc_i  = g w_i (+ g (1 - sum w) for i = L)
//   dL/ddd_k = G_k T_k e^{-dd_k} - sum_{i>k} G_i w_i,   G_i = dL/dw_i = g . (c_i - c_L) + gweights_i
// gweights [N,S]: gradient of the terms that read the weights directly (distortion loss), or null.
__global__ void k_train_ray_bwd(const TrainRays r, const float* __restrict__ sigma, const float* __restrict__ color,
                                const float* __restrict__ grad_rgb, const float* __restrict__ gweights,
                                float* __restrict__ gsigma, float* __restrict__ gcolor) {
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < r.N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float* sg = sigma + ray * r.S;
    const float* cl = color + 3 * ray * r.S;
    float* gs = gsigma + ray * r.S;
    float* gc = gcolor + 3 * ray * r.S;
    const float g0 = grad_rgb[3 * ray], g1 = grad_rgb[3 * ray + 1], g2 = grad_rgb[3 * ray + 2];
    const float l0 = cl[3 * (r.S - 1)], l1 = cl[3 * (r.S - 1) + 1], l2 = cl[3 * (r.S - 1) + 2];
    // pass 1: weights (kept in gs as scratch) and their sum
    float cum = 0.f, acc = 0.f;
    for (int i = 0; i < r.S; ++i) {
      float t0, t1;
      sample_interval(r, ray, i, t0, t1);
      const float dd = __fsub_rn(t1, t0) * sg[i];
      float w = (1.f - expf(-dd)) * expf(-cum);
      cum += dd;
      if (w != w) w = 0.f;
      gs[i] = w;
      gc[3 * i] = expf(-cum);   // transmittance behind sample i (scratch; rewritten below)
      acc += w;
    }
    // pass 2, back to front: suffix sum of G_i w_i; T_k e^{-dd_k} = transmittance behind sample k
    float suffix = 0.f;
    for (int i = r.S - 1; i >= 0; --i) {
      float t0, t1;
      sample_interval(r, ray, i, t0, t1);
      const float delta = __fsub_rn(t1, t0);
      const float w = gs[i];
      const float G = g0 * (cl[3 * i] - l0) + g1 * (cl[3 * i + 1] - l1) + g2 * (cl[3 * i + 2] - l2) +
                      (gweights ? gweights[ray * r.S + i] : 0.f);
      const float t_after = gc[3 * i];
      gs[i] = delta * (G * t_after - suffix);
      suffix += G * w;
      const float wc = (i == r.S - 1) ? w + (1.f - acc) : w;
      gc[3 * i] = g0 * wc;
      gc[3 * i + 1] = g1 * wc;
      gc[3 * i + 2] = g2 * wc;
    }
  }
}

// ---------------------------------------------------------------- backward, sample level
// sum_k wrow[k] * x[k]: x in registers, wrow a 16-byte aligned row of the shared-memory weights (four per load)
template <int K>
__device__ __forceinline__ float dot_row(const float* __restrict__ wrow, const float (&x)[K]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
    a0 = fmaf(w4.x, x[k], a0), a1 = fmaf(w4.y, x[k + 1], a1), a2 = fmaf(w4.z, x[k + 2], a2), a3 = fmaf(w4.w, x[k + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
// acc[k] += wrow[k] * v
template <int K>
__device__ __forceinline__ void axpy_row(const float* __restrict__ wrow, float v, float (&acc)[K]) {
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 w4 = *reinterpret_cast<const float4*>(wrow + k);
    acc[k] = fmaf(w4.x, v, acc[k]), acc[k + 1] = fmaf(w4.y, v, acc[k + 1]);
    acc[k + 2] = fmaf(w4.z, v, acc[k + 2]), acc[k + 3] = fmaf(w4.w, v, acc[k + 3]);
  }
}

struct FieldBwd {
  GridDev grid;
  const MlpF32* w32;
  TrainRays rays;
  const float* gsigma;
  const float* gcolor;
  const float* ggeo;   // [samples, 15] d loss / d geo features from the normal-prediction branch, or null
  float* grad_table;   // [L * T, 2], accumulated
  float* deltas;       // [kDeltaDim, cap]: one row per delta component, samples of the chunk along the row, so that the
  float* acts;         // [kActDim, cap]    per-thread "arrays" below are coalesced across the warp
  int64_t first, count;   // samples [first, first + count) of the batch
  int64_t cap;            // row length of deltas / acts
};

__global__ void __launch_bounds__(128, 3) k_train_field_bwd(const __grid_constant__ FieldBwd p) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ float s_fgrad[32 * 128];   // [feature][thread]: d loss / d hash features, read back by the rolled level loop
  MlpF32* w = reinterpret_cast<MlpF32*>(smem);
  for (int i = threadIdx.x; i < (int)(sizeof(MlpF32) / 16); i += blockDim.x)
    reinterpret_cast<uint4*>(w)[i] = reinterpret_cast<const uint4*>(p.w32)[i];
  __syncthreads();
  const TrainRays& r = p.rays;
  const int lane = threadIdx.x & 31;
  // every lane of a warp runs the same number of iterations (the scatter-add below merges lanes with shuffles); lanes
  // past the end redo the last sample with zero gradients
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (p.count + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t e_raw = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = e_raw < p.count;
    const int64_t e = live ? e_raw : p.count - 1;
    const int64_t s = p.first + e;
    const int64_t ray = s / r.S;
    const int i = (int)(s - ray * r.S);
    // strided views: X(i) is element i of this thread's vector
    auto col = [&](float* base, int off) { return base + (size_t)off * p.cap + e; };
#define SGN_AT(ptr, i) (ptr)[(size_t)(i) * p.cap]
    float px, py, pz, sh[16];
    const bool sel = sample_position(r, ray, i, px, py, pz);
    sh16(__ldg(r.dirs + 3 * ray), __ldg(r.dirs + 3 * ray + 1), __ldg(r.dirs + 3 * ray + 2), sh);
    // Layer arithmetic: the input vector of a layer sits in registers (x32 / x64, statically indexed), the weights come
    // out of shared memory four at a time (one LDS.128 per four FMAs, the same address for the whole warp), the outputs
    // go to the activation / delta slabs (k_outer_reduce reads them) and are read back as the next layer's registers.
    // Transposed products (delta through W^T) run in accumulate form: acc[k] += W[n][k] * delta_n over a runtime n.
    float x32[32], x64[64];
    float* feat = col(p.acts, 0);
    float* h0 = col(p.acts, 32);
    float* hin = col(p.acts, 96);
    float* h1 = col(p.acts, 128);
    float* h2 = col(p.acts, 192);
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      float2 f = encode_level(p.grid.table + (size_t)l * p.grid.size, p.grid.mask, p.grid.res[l], px, py, pz);
      x32[2 * l] = f.x, x32[2 * l + 1] = f.y;
      SGN_AT(feat, 2 * l) = f.x;
      SGN_AT(feat, 2 * l + 1) = f.y;
    }
#pragma unroll 2
    for (int n = 0; n < 64; ++n) SGN_AT(h0, n) = fmaxf(w->b_base0[n] + dot_row<32>(w->w_base0 + n * 32, x32), 0.f);
#pragma unroll
    for (int k = 0; k < 64; ++k) x64[k] = SGN_AT(h0, k);
    float logit = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) x32[k] = sh[k];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float a = w->b_base1[n] + dot_row<64>(w->w_base1 + n * 64, x64);
      if (n == 0) logit = a;
      x32[16 + n] = n == 0 ? 0.f : a;
    }
#pragma unroll
    for (int k = 0; k < 32; ++k) SGN_AT(hin, k) = x32[k];
#pragma unroll 2
    for (int n = 0; n < 64; ++n) {
      const float b = r.head_bias ? __ldg(r.head_bias + 64 * ray + n) : w->b_head0[n];
      SGN_AT(h1, n) = fmaxf(b + dot_row<32>(w->w_head0 + n * 32, x32), 0.f);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) x64[k] = SGN_AT(h1, k);
#pragma unroll 1
    for (int n = 0; n < 64; ++n) SGN_AT(h2, n) = fmaxf(w->b_head1[n] + dot_row<64>(w->w_head1 + n * 64, x64), 0.f);
#pragma unroll
    for (int k = 0; k < 64; ++k) x64[k] = SGN_AT(h2, k);
    float dpre[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float cl = sigmoidf_(w->b_head2[c] + dot_row<64>(w->w_head2 + c * 64, x64));
      dpre[c] = p.gcolor[3 * s + c] * cl * (1.f - cl);
    }
    // ---- backward
    float* da0 = col(p.deltas, 0);
    float* dout1 = col(p.deltas, 64);
    float* da1 = col(p.deltas, 80);
    float* da2 = col(p.deltas, 144);
    float* dp = col(p.deltas, 208);
    SGN_AT(dp, 0) = dpre[0]; SGN_AT(dp, 1) = dpre[1]; SGN_AT(dp, 2) = dpre[2]; SGN_AT(dp, 3) = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {   // x64 = h2 -> da2 (stored for the n-loop below and for k_outer_reduce)
      const float g = w->w_head2[k] * dpre[0] + w->w_head2[64 + k] * dpre[1] + w->w_head2[128 + k] * dpre[2];
      SGN_AT(da2, k) = x64[k] > 0.f ? g : 0.f;
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) x64[k] = 0.f;
#pragma unroll 1
    for (int n = 0; n < 64; n += 4) {   // the four scalars of a trip are in flight together
      const float v0 = SGN_AT(da2, n), v1 = SGN_AT(da2, n + 1), v2 = SGN_AT(da2, n + 2), v3 = SGN_AT(da2, n + 3);
      axpy_row<64>(w->w_head1 + n * 64, v0, x64);
      axpy_row<64>(w->w_head1 + n * 64 + 64, v1, x64);
      axpy_row<64>(w->w_head1 + n * 64 + 128, v2, x64);
      axpy_row<64>(w->w_head1 + n * 64 + 192, v3, x64);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) SGN_AT(da1, k) = SGN_AT(h1, k) > 0.f ? x64[k] : 0.f;
    // trunc_exp backward: d exp(x) = exp(clamp(x, max = 15)); the selector multiplies the density only
    float d16[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) d16[n] = 0.f;
#pragma unroll 1
    for (int m = 0; m < 64; m += 4) {
      const float v0 = SGN_AT(da1, m), v1 = SGN_AT(da1, m + 1), v2 = SGN_AT(da1, m + 2), v3 = SGN_AT(da1, m + 3);
      axpy_row<16>(w->w_head0 + m * 32 + 16, v0, d16);
      axpy_row<16>(w->w_head0 + m * 32 + 48, v1, d16);
      axpy_row<16>(w->w_head0 + m * 32 + 80, v2, d16);
      axpy_row<16>(w->w_head0 + m * 32 + 112, v3, d16);
    }
    d16[0] = sel ? p.gsigma[s] * w->avg_density * expf(fminf(logit, 15.f)) : 0.f;
    if (p.ggeo) {
#pragma unroll
      for (int n = 1; n < 16; ++n) d16[n] += p.ggeo[15 * s + n - 1];
    }
#pragma unroll
    for (int n = 0; n < 16; ++n) SGN_AT(dout1, n) = d16[n];
#pragma unroll
    for (int k = 0; k < 64; ++k) x64[k] = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) axpy_row<64>(w->w_base1 + n * 64, d16[n], x64);
#pragma unroll
    for (int k = 0; k < 64; ++k) SGN_AT(da0, k) = SGN_AT(h0, k) > 0.f ? x64[k] : 0.f;
    // ---- hash features: d feat = W_base0^T da0, then the scatter-add into the table gradient
#pragma unroll
    for (int k = 0; k < 32; ++k) x32[k] = 0.f;
#pragma unroll 1
    for (int n = 0; n < 64; n += 4) {
      const float v0 = SGN_AT(da0, n), v1 = SGN_AT(da0, n + 1), v2 = SGN_AT(da0, n + 2), v3 = SGN_AT(da0, n + 3);
      axpy_row<32>(w->w_base0 + n * 32, v0, x32);
      axpy_row<32>(w->w_base0 + n * 32 + 32, v1, x32);
      axpy_row<32>(w->w_base0 + n * 32 + 64, v2, x32);
      axpy_row<32>(w->w_base0 + n * 32 + 96, v3, x32);
    }
    // the level loop stays rolled (its body is long): the 32 feature gradients go through a shared-memory column
#pragma unroll
    for (int k = 0; k < 32; ++k) s_fgrad[k * 128 + threadIdx.x] = live ? x32[k] : 0.f;
#pragma unroll 1
    for (int l = 0; l < 16; ++l) {
      const LevelCoords L = level_coords(p.grid.res[l], px, py, pz);
      float2* gt = reinterpret_cast<float2*>(p.grad_table) + (size_t)l * p.grid.size;
      scatter_level_merged(gt, L, p.grid.mask, s_fgrad[(2 * l) * 128 + threadIdx.x], s_fgrad[(2 * l + 1) * 128 + threadIdx.x], lane);
    }
#undef SGN_AT
  }
}

// grad_head_bias[ray][n] += sum over the ray's samples inside the chunk of da1[n] (rows 80..143 of the delta slab): the
// gradient of the per-ray bias of head layer 0, from which the appearance embedding's gradients follow.
__global__ void k_ray_bias_reduce(const float* __restrict__ deltas, int64_t first, int64_t count, int64_t cap, int S,
                                  float* __restrict__ grad_head_bias) {
  const int64_t ray0 = first / S, ray1 = (first + count - 1) / S;
  const int64_t total = (ray1 - ray0 + 1) * 64;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(e / (ray1 - ray0 + 1));            // consecutive threads: consecutive rays of one component row
    const int64_t ray = ray0 + (e - (int64_t)n * (ray1 - ray0 + 1));
    const int64_t s0 = max(ray * S, first), s1 = min((ray + 1) * S, first + count);
    const float* row = deltas + (size_t)(80 + n) * cap - first;
    float acc = 0.f;
    for (int64_t s = s0; s < s1; ++s) acc += row[s];
    grad_head_bias[ray * 64 + n] += acc;
  }
}

// Appearance embedding (NerfactoField while training: embedding_appearance(camera_indices) concatenated to the head's
// input): head_bias[ray] = b + W_app . E[cam[ray]];  W_app [64,32] = the head's columns 31..62, E [M,32].
__global__ void k_appearance_bias(const float* __restrict__ w_app, const float* __restrict__ b, const float* __restrict__ emb,
                                  const int32_t* __restrict__ cam, int64_t N, float* __restrict__ head_bias) {
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < N * 64; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ray = e >> 6;
    const int n = (int)(e & 63);
    const float* row = emb + (size_t)__ldg(cam + ray) * 32;
    float a = __ldg(b + n);
#pragma unroll
    for (int k = 0; k < 32; ++k) a = fmaf(__ldg(w_app + n * 32 + k), __ldg(row + k), a);
    head_bias[e] = a;
  }
}

// ... and its backward: gW_app[n][k] += sum_ray g[ray][n] E[cam][k], gb[n] += sum_ray g[ray][n], gE[cam][k] += sum_n W_app[n][k] g[ray][n].
// 256 threads own the 64 x 32 outputs (8 each) of one slab of rays.
__global__ void __launch_bounds__(256) k_appearance_bwd(const float* __restrict__ w_app, const float* __restrict__ emb,
                                                        const int32_t* __restrict__ cam, const float* __restrict__ g, int64_t N,
                                                        int per_cta, float* __restrict__ gw_app, float* __restrict__ gb,
                                                        float* __restrict__ gemb) {
  __shared__ float sw[64 * 32];
  __shared__ float sg[64];
  __shared__ float se[32];
  for (int i = threadIdx.x; i < 64 * 32; i += 256) sw[i] = w_app[i];
  const int n = threadIdx.x >> 2, k0 = (threadIdx.x & 3) * 8;
  float acc[8] = {}, bacc = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * per_cta, r1 = min(N, r0 + per_cta);
  for (int64_t ray = r0; ray < r1; ++ray) {
    __syncthreads();
    const int c = __ldg(cam + ray);
    if (threadIdx.x < 64) sg[threadIdx.x] = g[ray * 64 + threadIdx.x];
    else if (threadIdx.x < 96) se[threadIdx.x - 64] = emb[(size_t)c * 32 + threadIdx.x - 64];
    __syncthreads();
    const float gn = sg[n];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = fmaf(gn, se[k0 + q], acc[q]);
    if ((threadIdx.x & 3) == 0) bacc += gn;
    if (threadIdx.x < 32) {
      float ge = 0.f;
      for (int m = 0; m < 64; ++m) ge = fmaf(sw[m * 32 + threadIdx.x], sg[m], ge);
      atomicAdd(gemb + (size_t)c * 32 + threadIdx.x, ge);
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) atomicAdd(gw_app + n * 32 + k0 + q, acc[q]);
  if ((threadIdx.x & 3) == 0) atomicAdd(gb + n, bacc);
}

// dW[n][k] += sum_s D[s][doff + n] * A[s][aoff + k]  (n < N <= 64, k < K <= 64);  db[n] += sum_s D[s][doff + n]
// 256 threads own a 64 x 64 register tile (4 x 4 each); samples are staged 32 at a time through shared memory.
__global__ void __launch_bounds__(256) k_outer_reduce(const float* __restrict__ D, const float* __restrict__ A,
                                                      const __grid_constant__ OuterParams op, int64_t count, int64_t cap,
                                                      int per_cta) {
  const OuterLayer& ly = op.layer[blockIdx.y];
  const int doff = ly.doff, N = ly.N, aoff = ly.aoff, K = ly.K, ldw = ly.ldw;
  float* __restrict__ dW = ly.dW;
  float* __restrict__ db = ly.db;
  __shared__ __align__(16) float sD[32][64 + 4];   // row pitch 68 floats: 16-byte aligned rows for the float4 reads below
  __shared__ __align__(16) float sA[32][64 + 4];
  const int tn = (threadIdx.x >> 4) * 4, tk = (threadIdx.x & 15) * 4;
  float acc[4][4] = {};
  float bacc = 0.f;   // thread t < 64 sums column t of D
  const int64_t s0 = (int64_t)blockIdx.x * per_cta, s1 = min(count, s0 + per_cta);
  for (int64_t base = s0; base < s1; base += 32) {
    const int rows = (int)min((int64_t)32, s1 - base);
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int rr = e & 31, cc = e >> 5;     // consecutive threads read consecutive samples of one component row
      sD[rr][cc] = (rr < rows && cc < N) ? D[(size_t)(doff + cc) * cap + base + rr] : 0.f;
      sA[rr][cc] = (rr < rows && cc < K) ? A[(size_t)(aoff + cc) * cap + base + rr] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < 32; ++rr) {
      const float4 d4 = *reinterpret_cast<const float4*>(&sD[rr][tn]), a4 = *reinterpret_cast<const float4*>(&sA[rr][tk]);
      const float d[4] = {d4.x, d4.y, d4.z, d4.w}, a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int x = 0; x < 4; ++x)
#pragma unroll
        for (int y = 0; y < 4; ++y) acc[x][y] = fmaf(d[x], a[y], acc[x][y]);
    }
    if (threadIdx.x < 64)
      for (int rr = 0; rr < 32; ++rr) bacc += sD[rr][threadIdx.x];
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 4; ++x)
#pragma unroll
    for (int y = 0; y < 4; ++y)
      if (tn + x < N && tk + y < K) atomicAdd(dW + (tn + x) * ldw + tk + y, acc[x][y]);
  if (db && threadIdx.x < N) atomicAdd(db + threadIdx.x, bacc);
}

void launch_outer_reduce(const float* D, const float* A, const OuterParams& op, int layers, int64_t count, int64_t cap,
                         cudaStream_t st) {
  // 512-sample slabs keep ~1 000 CTAs in flight, which is what hides the latency of the slab loads (one 32-sample stage
  // at a time per CTA); blockIdx.y = layer
  const int per_cta = 512;
  const int ctas = (int)((count + per_cta - 1) / per_cta);
  k_outer_reduce<<<dim3(ctas, layers), 256, 0, st>>>(D, A, op, count, cap, per_cta);
}

// ---------------------------------------------------------------- loss + optimizer
// nerfstudio L1Loss / MSELoss = torch.nn.L1Loss / MSELoss (mean over all elements); one block, fixed summation order.
__global__ void __launch_bounds__(1024) k_rgb_loss(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                   int l1, float* __restrict__ loss, float* __restrict__ grad) {
  __shared__ double part[1024];
  double acc = 0.0;
  const float inv = 1.f / (float)n;
  for (int64_t i = threadIdx.x; i < n; i += 1024) {
    const float d = pred[i] - target[i];
    if (l1) {
      acc += fabsf(d);
      if (grad) grad[i] = d > 0.f ? inv : (d < 0.f ? -inv : 0.f);
    } else {
      acc += (double)d * d;
      if (grad) grad[i] = 2.f * d * inv;
    }
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 512; o; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(part[0] / (double)n);
}

// torch.optim.Adam (nerfstudio AdamOptimizerConfig: betas (0.9, 0.999), eps 1e-15, no weight decay, no amsgrad)
__global__ void k_adam(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       int64_t n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

static int blocks_for(int64_t n, int threads, int per_sm) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)sm_count() * per_sm));
}

}  // namespace sgn

using namespace sgn;

extern "C" int64_t sgn_mlp_param_count(void) { return (int64_t)(sizeof(MlpF32) / sizeof(float)); }

extern "C" int sgn_field_mlp_params(const SgnField* f, float** d_params) {
  SGN_CHECK_ARG(f && d_params, "null pointer");
  *d_params = reinterpret_cast<float*>(f->d_f32);
  return SGN_OK;
}

static int check_rays(const SgnField* f, const float* o, const float* d, int64_t N, const float* bins, const float* ray_bins, int S) {
  SGN_CHECK_ARG(f != nullptr, "null field");
  SGN_CHECK_ARG(N >= 0 && S >= 1 && S <= 1024 && N * (int64_t)S < ((int64_t)1 << 40), "bad batch shape");
  SGN_CHECK_ARG(N == 0 || (o && d), "null rays");
  SGN_CHECK_ARG((bins != nullptr) != (ray_bins != nullptr), "exactly one of the shared bins / per-ray bins must be given");
  return SGN_OK;
}

extern "C" int sgn_train_forward(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N,
                                 const float* d_bins, const float* d_ray_bins, int S, const float* d_head_bias, float* d_sigma,
                                 float* d_color, float* d_rgb, float* d_acc, void* stream) {
  int rc = check_rays(f, d_origins, d_directions, N, d_bins, d_ray_bins, S);
  if (rc) return rc;
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_sigma && d_color && d_rgb, "null output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TrainRays r{d_origins, d_directions, d_bins, d_ray_bins, N, S, d_head_bias};
  static bool attr = false;
  if (!attr) {
    SGN_CUDA(cudaFuncSetAttribute(k_train_field, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpF32)));
    SGN_CUDA(cudaFuncSetAttribute(k_train_field_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpF32)));
    attr = true;
  }
  k_train_field<<<blocks_for(N * S, 128, 8), 128, sizeof(MlpF32), st>>>(f->grid, f->d_f32, r, d_sigma, d_color);
  SGN_LAUNCH_CHECK();
  k_train_composite<<<blocks_for(N, 128, 8), 128, 0, st>>>(r, d_sigma, d_color, d_rgb, d_acc);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

constexpr int64_t kBwdChunk = 1 << 18;   // samples per back-propagation chunk (deltas + activations: 490 MB of scratch)

extern "C" int64_t sgn_train_ws_bytes(int64_t N, int S) {
  if (N <= 0 || S <= 0) return 0;
  const int64_t samples = N * S;
  return samples * 4 * (int64_t)sizeof(float) + std::min(samples, kBwdChunk) * (kDeltaDim + kActDim) * (int64_t)sizeof(float) + 256;
}

extern "C" int sgn_train_backward(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N,
                                  const float* d_bins, const float* d_ray_bins, int S, const float* d_head_bias,
                                  const float* d_sigma, const float* d_color, const float* d_grad_rgb,
                                  const float* d_grad_weights, const float* d_grad_geo, float* d_grad_table, float* d_grad_mlp,
                                  float* d_grad_head_bias, void* d_ws, int64_t ws_bytes, void* stream) {
  int rc = check_rays(f, d_origins, d_directions, N, d_bins, d_ray_bins, S);
  if (rc) return rc;
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_sigma && d_color && d_grad_rgb && d_grad_table && d_grad_mlp && d_ws, "null pointer");
  SGN_CHECK_ARG(ws_bytes >= sgn_train_ws_bytes(N, S), "workspace smaller than sgn_train_ws_bytes");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_ws) & 15) == 0, "workspace must be 16-byte aligned");
  SGN_CHECK_ARG((d_head_bias != nullptr) == (d_grad_head_bias != nullptr),
                "d_head_bias and d_grad_head_bias are given together (per-image appearance) or not at all");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const TrainRays r{d_origins, d_directions, d_bins, d_ray_bins, N, S, d_head_bias};
  const int64_t samples = N * S;
  float* gsigma = reinterpret_cast<float*>(d_ws);
  float* gcolor = gsigma + samples;
  float* deltas = gcolor + 3 * samples;
  float* acts = deltas + std::min(samples, kBwdChunk) * kDeltaDim;
  static bool attr = false;
  if (!attr) {
    SGN_CUDA(cudaFuncSetAttribute(k_train_field_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MlpF32)));
    attr = true;
  }
  k_train_ray_bwd<<<blocks_for(N, 128, 8), 128, 0, st>>>(r, d_sigma, d_color, d_grad_rgb, d_grad_weights, gsigma, gcolor);
  SGN_LAUNCH_CHECK();
  MlpF32* G = reinterpret_cast<MlpF32*>(d_grad_mlp);   // gradients in the parameter block's own layout
  for (int64_t first = 0; first < samples; first += kBwdChunk) {
    FieldBwd p;
    p.grid = f->grid; p.w32 = f->d_f32; p.rays = r; p.gsigma = gsigma; p.gcolor = gcolor; p.ggeo = d_grad_geo;
    p.grad_table = d_grad_table; p.deltas = deltas; p.acts = acts;
    p.first = first; p.count = std::min(kBwdChunk, samples - first);
    p.cap = std::min(samples, kBwdChunk);
    k_train_field_bwd<<<blocks_for(p.count, 128, 8), 128, sizeof(MlpF32), st>>>(p);
    SGN_LAUNCH_CHECK();
    OuterParams op;   // the five weight gradients in one launch
    //             deltas (offset, N)  activations (offset, K)  ld
    op.layer[0] = {0, 64, 0, 32, 32, G->w_base0, G->b_base0};
    op.layer[1] = {64, 16, 32, 64, 64, G->w_base1, G->b_base1};
    op.layer[2] = {80, 64, 96, 32, 32, G->w_head0, G->b_head0};
    op.layer[3] = {144, 64, 128, 64, 64, G->w_head1, G->b_head1};
    op.layer[4] = {208, 3, 192, 64, 64, G->w_head2, G->b_head2};
    launch_outer_reduce(deltas, acts, op, 5, p.count, p.cap, st);
    SGN_LAUNCH_CHECK();
    if (d_grad_head_bias) {
      k_ray_bias_reduce<<<blocks_for((p.count / S + 2) * 64, 128, 8), 128, 0, st>>>(deltas, first, p.count, p.cap, S, d_grad_head_bias);
      SGN_LAUNCH_CHECK();
    }
  }
  return SGN_OK;
}

extern "C" int sgn_appearance_bias(const float* d_w_app, const float* d_bias, const float* d_embedding, int num_images,
                                   const int32_t* d_camera_indices, int64_t N, float* d_head_bias, void* stream) {
  SGN_CHECK_ARG(N >= 0 && num_images >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_w_app && d_bias && d_embedding && d_camera_indices && d_head_bias, "null pointer");
  k_appearance_bias<<<blocks_for(N * 64, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_w_app, d_bias, d_embedding, d_camera_indices, N, d_head_bias);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_appearance_bias_backward(const float* d_w_app, const float* d_embedding, int num_images,
                                            const int32_t* d_camera_indices, const float* d_grad_head_bias, int64_t N,
                                            float* d_grad_w_app, float* d_grad_bias, float* d_grad_embedding, void* stream) {
  SGN_CHECK_ARG(N >= 0 && num_images >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_w_app && d_embedding && d_camera_indices && d_grad_head_bias && d_grad_w_app && d_grad_bias && d_grad_embedding,
                "null pointer");
  const int per_cta = (int)std::max<int64_t>(16, (N + 2 * sm_count() - 1) / (2 * sm_count()));
  k_appearance_bwd<<<(int)((N + per_cta - 1) / per_cta), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_w_app, d_embedding, d_camera_indices, d_grad_head_bias, N, per_cta, d_grad_w_app, d_grad_bias, d_grad_embedding);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_rgb_loss(const float* d_pred, const float* d_target, int64_t n, int l1, float* d_loss, float* d_grad,
                            void* stream) {
  SGN_CHECK_ARG(n > 0 && d_pred && d_target && d_loss, "bad loss arguments");
  k_rgb_loss<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_pred, d_target, n, l1, d_loss, d_grad);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1,
                             float beta2, float eps, int step, void* stream) {
  SGN_CHECK_ARG(n >= 0 && step >= 1 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f, "bad Adam arguments");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_param && d_grad && d_m && d_v, "null pointer");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = sqrtf(1.f - powf(beta2, (float)step));
  k_adam<<<blocks_for(n, 256, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_param, d_grad, d_m, d_v, n, lr, beta1,
                                                                                     beta2, eps, bc1, bc2);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}
