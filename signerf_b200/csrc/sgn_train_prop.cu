// SURVEY §8(f) row 4, second half - what nerfacto's training step does around the main field (models/nerfacto.py
// get_outputs / get_loss_dict while training, as SIGNeRFModel inherits them, signerf/signerf.py:62-82):
//   * ProposalNetworkSampler.generate_ray_samples in TRAINING mode: stratified initial bins (one draw per ray,
//     single_jitter), two proposal levels whose weights drive a jittered PDF re-sampling (256 -> 96 -> 48 bins);
//   * interlevel_loss (the final histogram against each proposal level's, gradient into the proposal networks only) and
//     distortion_loss (on the final level, gradient into the main field's densities), nerfstudio model_components/losses.py;
//   * the proposal networks' backward: weights -> densities -> 16-wide MLP -> 5-level hash grid (scatter-add).
// The random draws are INPUTS ([3, N] uniform numbers), so the oracle and this path sample the same bins.
// A training batch is small (16 384 rays x (256 + 96 + 48) samples): fp32 on CUDA cores, one thread per ray or sample.
#include <algorithm>
#include <vector>

#include "sgn_device.cuh"

namespace sgn {

void host_pdf_u(int nb, std::vector<float>& u);
void host_linspace01(int n, std::vector<float>& out);
int launch_pdf_resample(const float* weights, const float* spacing_in, const float* u, const float* jitter, float anneal,
                        float* cdf_scratch, float* spacing_out, float* euclid_out, float s_near, float s_far, int64_t rays, int S,
                        int nb, cudaStream_t st);

constexpr int kPropParams = 16 * 10 + 16 + 16 + 1;   // w0 | b0 | w1 | b1, contiguous in PropDev

struct RayList {
  const float* origins;   // [N,3]
  const float* dirs;      // [N,3]
  const float* eu;        // [N,S+1] euclidean bin edges
  int64_t N;
  int S;
};

__device__ __forceinline__ bool ray_sample_position(const RayList& r, int64_t ray, int i, float& px, float& py, float& pz) {
  const float t0 = __ldg(r.eu + ray * (r.S + 1) + i), t1 = __ldg(r.eu + ray * (r.S + 1) + i + 1);
  const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
  const float* o = r.origins + 3 * ray;
  const float* d = r.dirs + 3 * ray;
  return contract_to_unit(__fadd_rn(__ldg(o), __fmul_rn(__ldg(d), mid)), __fadd_rn(__ldg(o + 1), __fmul_rn(__ldg(d + 1), mid)),
                          __fadd_rn(__ldg(o + 2), __fmul_rn(__ldg(d + 2), mid)), px, py, pz);
}

// SpacedSampler.generate_ray_samples: spacing-domain bins linspace(0, 1, S+1), moved between the neighbouring bin centres
// by the ray's draw while training (jitter != NULL), and their euclidean images.
__global__ void k_initial_bins(const float* __restrict__ lin, const float* __restrict__ jitter, float s_near, float s_far,
                               int64_t N, int S, float* __restrict__ sp, float* __restrict__ eu) {
  const int64_t total = N * (S + 1);
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ray = e / (S + 1);
    const int i = (int)(e - ray * (S + 1));
    float b = __ldg(lin + i);
    if (jitter) {
      const float lower = i == 0 ? __ldg(lin) : __fmul_rn(__fadd_rn(__ldg(lin + i), __ldg(lin + i - 1)), 0.5f);
      const float upper = i == S ? __ldg(lin + S) : __fmul_rn(__fadd_rn(__ldg(lin + i + 1), __ldg(lin + i)), 0.5f);
      b = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), __ldg(jitter + ray)));
    }
    sp[e] = b;
    eu[e] = to_euclid(b, s_near, s_far);
  }
}

struct PropFwd {
  float out;        // pre-activation
  float a[16];      // hidden pre-activations
  float feat[10];
};

__device__ __forceinline__ void prop_forward(const PropDev& net, float px, float py, float pz, PropFwd& f) {
#pragma unroll
  for (int l = 0; l < 5; ++l) {
    const float2 v = encode_level(net.grid.table + (size_t)l * net.grid.size, net.grid.mask, net.grid.res[l], px, py, pz);
    f.feat[2 * l] = v.x;
    f.feat[2 * l + 1] = v.y;
  }
  f.out = net.b1;
#pragma unroll
  for (int n = 0; n < 16; ++n) {
    float a = net.b0[n];
#pragma unroll
    for (int k = 0; k < 10; ++k) a = fmaf(net.w0[n * 10 + k], f.feat[k], a);
    f.a[n] = a;
    f.out = fmaf(net.w1[n], fmaxf(a, 0.f), f.out);
  }
}

// HashMLPDensityField.density_fn for every sample: one thread per sample.
__global__ void __launch_bounds__(128) k_prop_sigma(const PropDev* __restrict__ netp, const RayList r, float* __restrict__ sigma) {
  __shared__ PropDev net;
  for (int i = threadIdx.x; i < (int)(sizeof(PropDev) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&net)[i] = reinterpret_cast<const uint32_t*>(netp)[i];
  __syncthreads();
  const int64_t total = r.N * r.S;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ray = s / r.S;
    float px, py, pz;
    const bool sel = ray_sample_position(r, ray, (int)(s - ray * r.S), px, py, pz);
    PropFwd f;
    prop_forward(net, px, py, pz, f);
    sigma[s] = sel ? net.avg_density * expf(f.out) : 0.f;
  }
}

// RaySamples.get_weights: one thread per ray.
__global__ void k_weights_fwd(const float* __restrict__ eu, const float* __restrict__ sigma, int64_t N, int S,
                              float* __restrict__ w) {
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float* b = eu + ray * (S + 1);
    float cum = 0.f;
    for (int i = 0; i < S; ++i) {
      const float dd = __fsub_rn(b[i + 1], b[i]) * sigma[ray * S + i];
      float v = (1.f - expf(-dd)) * expf(-cum);
      cum += dd;
      if (v != v) v = 0.f;
      w[ray * S + i] = v;
    }
  }
}

// dL/dw -> dL/dsigma through w_i = (1 - e^{-dd_i}) e^{-sum_{j<i} dd_j}, dd_i = delta_i sigma_i:
//   dL/ddd_k = g_k T_k e^{-dd_k} - sum_{i>k} g_i w_i     (T_k e^{-dd_k} = transmittance behind sample k)
__global__ void k_weights_bwd(const float* __restrict__ eu, const float* __restrict__ sigma, const float* __restrict__ gw,
                              int64_t N, int S, float* __restrict__ gsigma) {
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float* b = eu + ray * (S + 1);
    const float* sg = sigma + ray * S;
    const float* g = gw + ray * S;
    float* out = gsigma + ray * S;
    float cum = 0.f;
    for (int i = 0; i < S; ++i) cum += __fsub_rn(b[i + 1], b[i]) * sg[i];
    // back to front: cum = sum_{j<=i} dd_j
    float suffix = 0.f;
    for (int i = S - 1; i >= 0; --i) {
      const float delta = __fsub_rn(b[i + 1], b[i]);
      const float dd = delta * sg[i];
      const float t_after = expf(-cum);
      cum -= dd;
      float w = (1.f - expf(-dd)) * expf(-cum);
      if (w != w) w = 0.f;
      out[i] = delta * (g[i] * t_after - suffix);
      suffix += g[i] * w;
    }
  }
}

// losses.py lossfun_outer + its gradient with respect to the proposal weights, one thread per ray.
//   w_outer_i = sum_{k = lo_i}^{hi_i} wp_k,  lo_i = searchsorted(tp_starts, t_start_i, right) - 1,  hi_i = searchsorted(tp_ends,
//   t_end_i, right), both clamped to [0, Sp - 1];  loss_i = max(w_i - w_outer_i, 0)^2 / (w_i + eps)
// scale = interlevel_loss_mult / (N * Sf) (torch.mean); the gradient is accumulated as a difference array in gwp and
// integrated by a running sum.
__global__ void k_interlevel(const float* __restrict__ spf, const float* __restrict__ wf, const float* __restrict__ spp,
                             const float* __restrict__ wp, int64_t N, int Sf, int Sp, float scale, float* __restrict__ cy_scratch,
                             float* __restrict__ loss, float* __restrict__ gwp) {
  const float eps = 1.1920928955078125e-07f;   // torch.finfo(float32).eps
  float local = 0.f;
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float* t = spf + ray * (Sf + 1);
    const float* tp = spp + ray * (Sp + 1);
    const float* y = wp + ray * Sp;
    float* cy = cy_scratch + ray * (Sp + 1);
    float* g = gwp + ray * Sp;
    float run = 0.f;
    cy[0] = 0.f;
    for (int k = 0; k < Sp; ++k) {
      run = __fadd_rn(run, y[k]);
      cy[k + 1] = run;
      g[k] = 0.f;
    }
    int a = 0, b = 0;   // merge pointers: both edge sequences ascend along the ray
    for (int i = 0; i < Sf; ++i) {
      while (a < Sp && !(tp[a] > t[i])) ++a;          // searchsorted(tp[:-1], t_i, right)
      while (b < Sp && !(tp[b + 1] > t[i + 1])) ++b;  // searchsorted(tp[1:], t_{i+1}, right)
      const int lo = min(max(a - 1, 0), Sp - 1), hi = min(b, Sp - 1);
      const float w = wf[ray * Sf + i];
      const float diff = fmaxf(__fsub_rn(w, __fsub_rn(cy[hi + 1], cy[lo])), 0.f);
      local += diff * diff / (w + eps);
      if (diff > 0.f && hi >= lo) {
        const float gi = -2.f * diff / (w + eps) * scale;
        g[lo] += gi;
        if (hi + 1 < Sp) g[hi + 1] -= gi;
      }
    }
    run = 0.f;
    for (int k = 0; k < Sp; ++k) {
      run += g[k];
      g[k] = run;
    }
  }
  // block sum -> one atomic per block (a reported scalar: the order of the additions is not fixed)
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v * scale);
  }
}

// losses.py lossfun_distortion on the final level and its gradient with respect to the weights, one thread per ray:
//   L = sum_ij w_i w_j |u_i - u_j| + sum_i w_i^2 (t_{i+1} - t_i) / 3,  u = bin centres;  dL/dw_k = 2 sum_j w_j |u_k - u_j| + 2 w_k d_k / 3
__global__ void k_distortion(const float* __restrict__ sp, const float* __restrict__ w, int64_t N, int S, float scale,
                             float* __restrict__ loss, float* __restrict__ gw) {
  float local = 0.f;
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float* t = sp + ray * (S + 1);
    const float* y = w + ray * S;
    for (int k = 0; k < S; ++k) {
      const float uk = __fmul_rn(__fadd_rn(t[k + 1], t[k]), 0.5f);
      float inner = 0.f;
      for (int j = 0; j < S; ++j) inner = fmaf(y[j], fabsf(__fsub_rn(uk, __fmul_rn(__fadd_rn(t[j + 1], t[j]), 0.5f))), inner);
      const float dk = __fsub_rn(t[k + 1], t[k]);
      local += y[k] * inner + y[k] * y[k] * dk * (1.f / 3.f);
      if (gw) gw[ray * S + k] = scale * (2.f * inner + 2.f * y[k] * dk * (1.f / 3.f));
    }
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) atomicAdd(loss, v * scale);
  }
}

// Proposal network backward, one thread per sample: recompute the activations, back-propagate dL/dsigma to the MLP
// parameters (warp-reduced, then one shared-memory atomic per warp and value, one global atomic per block and value) and
// to the hash features (scatter-add, fp32 vector atomics).
__global__ void __launch_bounds__(128) k_prop_bwd(const PropDev* __restrict__ netp, const RayList r, const float* __restrict__ gsigma,
                                                  float* __restrict__ grad_table, float* __restrict__ grad_mlp) {
  __shared__ PropDev net;
  __shared__ float acc[kPropParams];
  for (int i = threadIdx.x; i < (int)(sizeof(PropDev) / 4); i += blockDim.x)
    reinterpret_cast<uint32_t*>(&net)[i] = reinterpret_cast<const uint32_t*>(netp)[i];
  for (int i = threadIdx.x; i < kPropParams; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t total = r.N * r.S;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // every lane of a warp runs the same number of iterations (the reductions below are warp-wide)
  const int64_t iters = (total + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t s = it * stride + blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = s < total;
    float d_out = 0.f, px = 0.f, py = 0.f, pz = 0.f;
    PropFwd f;
    if (live) {
      const int64_t ray = s / r.S;
      const bool sel = ray_sample_position(r, ray, (int)(s - ray * r.S), px, py, pz);
      prop_forward(net, px, py, pz, f);
      // trunc_exp backward: exp(clamp(x, max = 15)); the selector multiplies the density only
      d_out = sel ? gsigma[s] * net.avg_density * expf(fminf(f.out, 15.f)) : 0.f;
    } else {
#pragma unroll
      for (int n = 0; n < 16; ++n) f.a[n] = 0.f;
#pragma unroll
      for (int k = 0; k < 10; ++k) f.feat[k] = 0.f;
    }
    auto warp_add = [&](int slot, float v) {
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v != 0.f) atomicAdd(&acc[slot], v);
    };
    float dfeat[10];
#pragma unroll
    for (int k = 0; k < 10; ++k) dfeat[k] = 0.f;
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      const float da = f.a[n] > 0.f ? net.w1[n] * d_out : 0.f;
      warp_add(160 + 16 + n, d_out * fmaxf(f.a[n], 0.f));   // dw1[n]
      warp_add(160 + n, da);                                // db0[n]
#pragma unroll
      for (int k = 0; k < 10; ++k) {
        warp_add(n * 10 + k, da * f.feat[k]);               // dw0[n][k]
        dfeat[k] = fmaf(net.w0[n * 10 + k], da, dfeat[k]);
      }
    }
    warp_add(160 + 32, d_out);                              // db1
    // hash features: scatter-add, runs of samples in the same cell merged inside the warp (scatter_level_merged)
    const bool any = live && d_out != 0.f;
#pragma unroll 1
    for (int l = 0; l < 5; ++l) {
      const LevelCoords L = level_coords(net.grid.res[l], px, py, pz);
      float2* gt = reinterpret_cast<float2*>(grad_table) + (size_t)l * net.grid.size;
      scatter_level_merged(gt, L, net.grid.mask, any ? dfeat[2 * l] : 0.f, any ? dfeat[2 * l + 1] : 0.f, lane);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kPropParams; i += blockDim.x)
    if (acc[i] != 0.f) atomicAdd(grad_mlp + i, acc[i]);
}

static int blocks_of(int64_t n, int threads, int per_sm) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)sm_count() * per_sm));
}
static float h_spacing_fn(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }

}  // namespace sgn

using namespace sgn;

extern "C" int64_t sgn_prop_param_count(void) { return kPropParams; }

extern "C" int sgn_field_prop_params(const SgnField* f, int level, float** d_params) {
  SGN_CHECK_ARG(f && d_params, "null pointer");
  SGN_CHECK_ARG(level >= 0 && level < f->num_proposals, "no such proposal network");
  *d_params = reinterpret_cast<float*>(reinterpret_cast<char*>(f->d_prop[level]) + offsetof(PropDev, w0));
  return SGN_OK;
}

extern "C" int64_t sgn_train_sample_ws_bytes(int64_t N, int S0, int S1) {
  return N <= 0 ? 0 : N * (int64_t)(std::max(S0, S1) + 1) * 4 + 64 + (int64_t)(S0 + 1 + S1 + 1 + 64) * 8;
}

extern "C" int sgn_train_sample(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N, int S0, int S1,
                                int S2, float near_plane, float far_plane, const float* d_jitter, float anneal,
                                const SgnTrainSamples* out, void* d_ws, int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(f != nullptr && out != nullptr, "null pointer");
  SGN_CHECK_ARG(f->num_proposals == 2, "the training sampler needs a field created with 2 proposal networks");
  SGN_CHECK_ARG(N >= 0 && S0 >= 1 && S0 <= 1024 && S1 >= 1 && S1 <= 1024 && S2 >= 1 && S2 <= 1024, "bad sample counts");
  SGN_CHECK_ARG(anneal >= 0.f && anneal <= 1.f, "anneal must be in [0, 1]");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_origins && d_directions && d_ws, "null pointer");
  for (int l = 0; l < 3; ++l) SGN_CHECK_ARG(out->d_spacing[l] && out->d_euclid[l], "null output");
  for (int l = 0; l < 2; ++l) SGN_CHECK_ARG(out->d_sigma[l] && out->d_weights[l], "null output");
  SGN_CHECK_ARG(ws_bytes >= sgn_train_sample_ws_bytes(N, S0, S1) && (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0,
                "workspace smaller than sgn_train_sample_ws_bytes or misaligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const float s_near = h_spacing_fn(near_plane), s_far = h_spacing_fn(far_plane);
  // host tables: linspace(0, 1, S0+1) and the PDF sampler's u per level (torch's fp32 values)
  std::vector<float> lin, u1, u2;
  host_linspace01(S0 + 1, lin);
  host_pdf_u(S1 + 1, u1);   // eval: + 1 / (2 nb); training: the same linspace without it, see launch_pdf_resample
  host_pdf_u(S2 + 1, u2);
  float* cdf = reinterpret_cast<float*>(d_ws);
  float* d_lin = cdf + N * (int64_t)(std::max(S0, S1) + 1) + 16;
  float* d_u1 = d_lin + (S0 + 1);
  float* d_u2 = d_u1 + (S1 + 1);
  if (d_jitter) {   // u = linspace(0, 1 - 1/nb, nb) without the half-bin offset; the ray's draw / nb is added in the kernel
    for (int l = 0; l < 2; ++l) {
      std::vector<float>& u = l ? u2 : u1;
      const int nb = (int)u.size();
      const float end = (float)(1.0 - 1.0 / (double)nb);
      const float step = nb > 1 ? end / (float)(nb - 1) : 0.f;
      for (int i = 0; i < nb; ++i) u[i] = i < nb / 2 ? step * (float)i : end - step * (float)(nb - i - 1);
    }
  }
  SGN_CUDA(cudaMemcpyAsync(d_lin, lin.data(), (S0 + 1) * 4, cudaMemcpyHostToDevice, st));
  SGN_CUDA(cudaMemcpyAsync(d_u1, u1.data(), (S1 + 1) * 4, cudaMemcpyHostToDevice, st));
  SGN_CUDA(cudaMemcpyAsync(d_u2, u2.data(), (S2 + 1) * 4, cudaMemcpyHostToDevice, st));
  SGN_CUDA(cudaStreamSynchronize(st));   // the host vectors go out of scope
  k_initial_bins<<<blocks_of(N * (S0 + 1), 256, 8), 256, 0, st>>>(d_lin, d_jitter, s_near, s_far, N, S0, out->d_spacing[0],
                                                                  out->d_euclid[0]);
  SGN_LAUNCH_CHECK();
  const int S[3] = {S0, S1, S2};
  for (int l = 0; l < 2; ++l) {
    const RayList r{d_origins, d_directions, out->d_euclid[l], N, S[l]};
    k_prop_sigma<<<blocks_of(N * S[l], 128, 8), 128, 0, st>>>(f->d_prop[l], r, out->d_sigma[l]);
    SGN_LAUNCH_CHECK();
    k_weights_fwd<<<blocks_of(N, 128, 8), 128, 0, st>>>(out->d_euclid[l], out->d_sigma[l], N, S[l], out->d_weights[l]);
    SGN_LAUNCH_CHECK();
    int rc = launch_pdf_resample(out->d_weights[l], out->d_spacing[l], l ? d_u2 : d_u1, d_jitter ? d_jitter + (l + 1) * N : nullptr,
                                 anneal, cdf, out->d_spacing[l + 1], out->d_euclid[l + 1], s_near, s_far, N, S[l], S[l + 1] + 1, st);
    if (rc) return rc;
  }
  return SGN_OK;
}

extern "C" int sgn_weights_from_density(const float* d_euclid, const float* d_sigma, int64_t N, int S, float* d_weights,
                                        void* stream) {
  SGN_CHECK_ARG(N >= 0 && S >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_euclid && d_sigma && d_weights, "null pointer");
  k_weights_fwd<<<blocks_of(N, 128, 8), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_euclid, d_sigma, N, S, d_weights);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_interlevel_loss(const float* d_spacing_final, const float* d_weights_final, int S_final,
                                   const float* d_spacing_prop, const float* d_weights_prop, int S_prop, int64_t N, float mult,
                                   float* d_loss, float* d_grad_weights_prop, void* d_ws, int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(N >= 0 && S_final >= 1 && S_prop >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_spacing_final && d_weights_final && d_spacing_prop && d_weights_prop && d_loss && d_grad_weights_prop && d_ws,
                "null pointer");
  SGN_CHECK_ARG(ws_bytes >= N * (int64_t)(S_prop + 1) * 4, "workspace smaller than N * (S_prop + 1) floats");
  const float scale = mult / ((float)N * (float)S_final);
  k_interlevel<<<blocks_of(N, 128, 8), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_spacing_final, d_weights_final, d_spacing_prop, d_weights_prop, N, S_final, S_prop, scale, reinterpret_cast<float*>(d_ws),
      d_loss, d_grad_weights_prop);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_distortion_loss(const float* d_spacing, const float* d_weights, int64_t N, int S, float mult, float* d_loss,
                                   float* d_grad_weights, void* stream) {
  SGN_CHECK_ARG(N >= 0 && S >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_spacing && d_weights && d_loss, "null pointer");
  k_distortion<<<blocks_of(N, 128, 8), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_spacing, d_weights, N, S,
                                                                                          mult / (float)N, d_loss, d_grad_weights);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_prop_backward(const SgnField* f, int level, const float* d_origins, const float* d_directions, int64_t N, int S,
                                 const float* d_euclid, const float* d_sigma, const float* d_grad_weights, float* d_grad_table,
                                 float* d_grad_mlp, void* d_ws, int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(f != nullptr, "null field");
  SGN_CHECK_ARG(level >= 0 && level < f->num_proposals, "no such proposal network");
  SGN_CHECK_ARG(N >= 0 && S >= 1 && S <= 1024, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_origins && d_directions && d_euclid && d_sigma && d_grad_weights && d_grad_table && d_grad_mlp && d_ws, "null pointer");
  SGN_CHECK_ARG(ws_bytes >= N * (int64_t)S * 4, "workspace smaller than N * S floats");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* gsigma = reinterpret_cast<float*>(d_ws);
  k_weights_bwd<<<blocks_of(N, 128, 8), 128, 0, st>>>(d_euclid, d_sigma, d_grad_weights, N, S, gsigma);
  SGN_LAUNCH_CHECK();
  const RayList r{d_origins, d_directions, d_euclid, N, S};
  k_prop_bwd<<<blocks_of(N * S, 128, 8), 128, 0, st>>>(f->d_prop[level], r, gsigma, d_grad_table, d_grad_mlp);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}
