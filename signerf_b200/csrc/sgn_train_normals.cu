// SURVEY §8(f) row 4, third part - the normal regularisers of `predict_normals=True` (signerf_config.py:33; loss terms
// signerf/signerf.py:69-80 = NerfactoModel's `rendered_orientation_loss` / `rendered_pred_normal_loss`):
//   * analytic normals: -normalize(d density_logit / d p) at the contracted sample position p ([EXT] Field.get_normals:
//     torch.autograd.grad WITHOUT create_graph - the normals are constants of the step; their gradient through the hash
//     grid's trilinear weights and the base MLP is evaluated in closed form here);
//   * predicted normals: NeRFEncoding(2 frequencies) of the raw position | 15 geo features -> MLP 27 -> 64 -> 64 -> 64 ->
//     Linear(64, 3) -> tanh -> normalize ([EXT] NerfactoField.get_outputs, PredNormalsFieldHead);
//   * orientation_loss (value only: weights and normals carry no graph) and pred_normal_loss, whose gradient reaches the
//     prediction MLP and, through the geo features, the base MLP and the hash table (handed to sgn_train_backward as
//     d_grad_geo).
// One thread per sample, fp32 on CUDA cores, activations / deltas of the backward in [component][sample] slabs so that the
// weight gradients are the same skinny outer-product reductions as the main field's (sgn_train.cu, k_outer_reduce).
#include <algorithm>

#include "sgn_device.cuh"

namespace sgn {

// parameter block of the prediction MLP + head, nn.Linear row-major [out][in]; weights first so that every row is 16-byte
// aligned in shared memory except w0's 27-wide rows (read scalar)
struct alignas(16) PnParams {
  float w1[64 * 64];
  float w2[64 * 64];
  float wh[3 * 64];
  float w0[64 * 27];
  float b0[64];
  float b1[64];
  float b2[64];
  float bh[4];
};
constexpr int kPnParams = (int)(sizeof(PnParams) / sizeof(float));   // 10 308
constexpr int kPnActs = 27 + 64 + 64 + 64;                             // inp | l0 | l1 | x3
constexpr int kPnDeltas = 64 + 64 + 64 + 4;                            // d_l0 | d_l1 | d_x3 | d_h (3 + pad)

struct BaseW {   // what the normals need of the main field: base layer 0 and 1
  float w0[64 * 32];
  float w1[16 * 64];
  float b0[64];
  float b1[16];
};

struct NormalRays {
  const float* origins;
  const float* dirs;
  const float* ray_bins;   // [N, S+1]
  int64_t N;
  int S;
};

__device__ __forceinline__ bool raw_and_contracted(const NormalRays& r, int64_t ray, int i, float raw[3], float& px, float& py, float& pz) {
  const float t0 = __ldg(r.ray_bins + ray * (r.S + 1) + i), t1 = __ldg(r.ray_bins + ray * (r.S + 1) + i + 1);
  const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
#pragma unroll
  for (int k = 0; k < 3; ++k) raw[k] = __fadd_rn(__ldg(r.origins + 3 * ray + k), __fmul_rn(__ldg(r.dirs + 3 * ray + k), mid));
  return contract_to_unit(raw[0], raw[1], raw[2], px, py, pz);
}

__device__ __forceinline__ void load_base(const MlpF32* __restrict__ w32, BaseW* sb) {
  for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) sb->w0[i] = w32->w_base0[i];
  for (int i = threadIdx.x; i < 16 * 64; i += blockDim.x) sb->w1[i] = w32->w_base1[i];
  for (int i = threadIdx.x; i < 64; i += blockDim.x) sb->b0[i] = w32->b_base0[i];
  for (int i = threadIdx.x; i < 16; i += blockDim.x) sb->b1[i] = w32->b_base1[i];
}

// hash features of all levels; with kGrad also d logit / d p through the trilinear weights, given d logit / d feat
__device__ __forceinline__ void encode_all(const GridDev& g, float px, float py, float pz, float feat[32]) {
#pragma unroll
  for (int l = 0; l < 16; ++l) {
    const float2 f = encode_level(g.table + (size_t)l * g.size, g.mask, g.res[l], px, py, pz);
    feat[2 * l] = f.x, feat[2 * l + 1] = f.y;
  }
}

// base MLP: out1[0] = density logit, out1[1..15] = geo features; optionally gfeat = d logit / d feat
template <bool kGrad>
__device__ __forceinline__ void base_mlp(const BaseW* sb, const float feat[32], float out1[16], float gfeat[32]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) out1[j] = sb->b1[j];
  if (kGrad) {
#pragma unroll
    for (int k = 0; k < 32; ++k) gfeat[k] = 0.f;
  }
#pragma unroll 2
  for (int n = 0; n < 64; ++n) {
    const float4* row = reinterpret_cast<const float4*>(sb->w0 + n * 32);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float4 w = row[k];
      a0 = fmaf(w.x, feat[4 * k], a0), a1 = fmaf(w.y, feat[4 * k + 1], a1), a2 = fmaf(w.z, feat[4 * k + 2], a2), a3 = fmaf(w.w, feat[4 * k + 3], a3);
    }
    const float a = sb->b0[n] + ((a0 + a1) + (a2 + a3));
    const float h = fmaxf(a, 0.f);
#pragma unroll
    for (int j = 0; j < 16; ++j) out1[j] = fmaf(sb->w1[j * 64 + n], h, out1[j]);
    if (kGrad) {
      const float c = a > 0.f ? sb->w1[n] : 0.f;   // row 0 of layer 1 = the density logit
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float4 w = row[k];
        gfeat[4 * k] = fmaf(w.x, c, gfeat[4 * k]), gfeat[4 * k + 1] = fmaf(w.y, c, gfeat[4 * k + 1]);
        gfeat[4 * k + 2] = fmaf(w.z, c, gfeat[4 * k + 2]), gfeat[4 * k + 3] = fmaf(w.w, c, gfeat[4 * k + 3]);
      }
    }
  }
}

// d logit / d p = sum over levels of res_l * (gfeat . d feat / d offset): the interpolation of encode_level differentiated
__device__ __forceinline__ void position_gradient(const GridDev& g, float px, float py, float pz, const float gfeat[32], float gp[3]) {
  gp[0] = gp[1] = gp[2] = 0.f;
#pragma unroll 1
  for (int l = 0; l < 16; ++l) {
    const LevelCoords L = level_coords(g.res[l], px, py, pz);
    uint32_t idx[8];
    corner_rows(L, g.mask, idx);
    const float2* tab = g.table + (size_t)l * g.size;
    float2 f[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) f[c] = __ldg(tab + idx[c]);
    const float mx = 1.f - L.ox, my = 1.f - L.oy, mz = 1.f - L.oz;
    const float g0 = gfeat[2 * l], g1 = gfeat[2 * l + 1];
    auto dot2 = [&](float2 a, float2 b) { return g0 * (a.x - b.x) + g1 * (a.y - b.y); };
    // f03 = f0 ox + f3 mx, f12 = f1 ox + f2 mx, f47 = f4 ox + f7 mx, f56 = f5 ox + f6 mx;
    // f0312 = f03 oy + f12 my, f4756 = f47 oy + f56 my; out = f0312 oz + f4756 mz
    const float dx = (dot2(f[0], f[3]) * L.oy + dot2(f[1], f[2]) * my) * L.oz + (dot2(f[4], f[7]) * L.oy + dot2(f[5], f[6]) * my) * mz;
    const float2 f03 = lerp2(f[0], L.ox, f[3], mx), f12 = lerp2(f[1], L.ox, f[2], mx);
    const float2 f47 = lerp2(f[4], L.ox, f[7], mx), f56 = lerp2(f[5], L.ox, f[6], mx);
    const float dy = dot2(f03, f12) * L.oz + dot2(f47, f56) * mz;
    const float dz = dot2(lerp2(f03, L.oy, f12, my), lerp2(f47, L.oy, f56, my));
    gp[0] = fmaf(g.res[l], dx, gp[0]), gp[1] = fmaf(g.res[l], dy, gp[1]), gp[2] = fmaf(g.res[l], dz, gp[2]);
  }
}

// NeRFEncoding(in_dim 3, frequencies 1 and 2, no input): sin([2 pi x f | 2 pi x f + pi / 2])
__device__ __forceinline__ void nerf_enc12(const float raw[3], float* out) {
  const float two_pi = 6.283185307179586f, half_pi = 1.5707963267948966f;
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const float s = two_pi * raw[d];
    out[2 * d] = sinf(s), out[2 * d + 1] = sinf(s * 2.f);
    out[6 + 2 * d] = sinf(s + half_pi), out[6 + 2 * d + 1] = sinf(s * 2.f + half_pi);
  }
}

template <int K>
__device__ __forceinline__ float pn_dot4(const float* __restrict__ wrow, const float (&x)[K]) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + k);
    a0 = fmaf(w.x, x[k], a0), a1 = fmaf(w.y, x[k + 1], a1), a2 = fmaf(w.z, x[k + 2], a2), a3 = fmaf(w.w, x[k + 3], a3);
  }
  return (a0 + a1) + (a2 + a3);
}
template <int K>
__device__ __forceinline__ void pn_axpy4(const float* __restrict__ wrow, float v, float (&acc)[K]) {
#pragma unroll
  for (int k = 0; k < K; k += 4) {
    const float4 w = *reinterpret_cast<const float4*>(wrow + k);
    acc[k] = fmaf(w.x, v, acc[k]), acc[k + 1] = fmaf(w.y, v, acc[k + 1]), acc[k + 2] = fmaf(w.z, v, acc[k + 2]), acc[k + 3] = fmaf(w.w, v, acc[k + 3]);
  }
}

struct NormalsFwd {
  GridDev grid;
  const MlpF32* w32;
  const PnParams* pn;
  NormalRays rays;
  float* normals;   // [N,S,3]
  float* pred;      // [N,S,3]
};

// dynamic indexing of the 64-wide register vectors above would spill: the MLP layers therefore keep their outputs in
// shared-memory columns [component][thread] (conflict-free), one block = 64 threads
constexpr int kNThreads = 64;

__global__ void __launch_bounds__(kNThreads) k_normals_fwd(const __grid_constant__ NormalsFwd p) {
  extern __shared__ __align__(16) unsigned char smem[];
  PnParams* w = reinterpret_cast<PnParams*>(smem);
  BaseW* sb = reinterpret_cast<BaseW*>(smem + sizeof(PnParams));
  float* col = reinterpret_cast<float*>(smem + sizeof(PnParams) + sizeof(BaseW));   // [3 x 64 components][kNThreads]
  for (int i = threadIdx.x; i < kPnParams; i += blockDim.x) reinterpret_cast<float*>(w)[i] = reinterpret_cast<const float*>(p.pn)[i];
  load_base(p.w32, sb);
  __syncthreads();
  const NormalRays& r = p.rays;
  const int64_t total = r.N * r.S;
  float* c0 = col + threadIdx.x;                       // component k of this thread's vector at c0[k * kNThreads]
  float* c1 = c0 + 64 * kNThreads;
  for (int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; s < total; s += (int64_t)gridDim.x * blockDim.x) {
    const int64_t ray = s / r.S;
    float raw[3], px, py, pz;
    raw_and_contracted(r, ray, (int)(s - ray * r.S), raw, px, py, pz);
    float feat[32], out1[16], gfeat[32], gp[3];
    encode_all(p.grid, px, py, pz, feat);
    base_mlp<true>(sb, feat, out1, gfeat);
    position_gradient(p.grid, px, py, pz, gfeat, gp);
    const float gn = fmaxf(sqrtf(gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]), 1e-12f);   // F.normalize eps
    p.normals[3 * s] = -gp[0] / gn, p.normals[3 * s + 1] = -gp[1] / gn, p.normals[3 * s + 2] = -gp[2] / gn;
    // prediction MLP: layer outputs through the shared-memory columns
    float inp[27];
    nerf_enc12(raw, inp);
#pragma unroll
    for (int j = 0; j < 15; ++j) inp[12 + j] = out1[1 + j];
    for (int n = 0; n < 64; ++n) {
      float a = w->b0[n];
#pragma unroll
      for (int k = 0; k < 27; ++k) a = fmaf(w->w0[n * 27 + k], inp[k], a);
      c0[n * kNThreads] = fmaxf(a, 0.f);
    }
    float x[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = c0[k * kNThreads];
    for (int n = 0; n < 64; ++n) c1[n * kNThreads] = fmaxf(w->b1[n] + pn_dot4<64>(w->w1 + n * 64, x), 0.f);
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = c1[k * kNThreads];
    for (int n = 0; n < 64; ++n) c0[n * kNThreads] = w->b2[n] + pn_dot4<64>(w->w2 + n * 64, x);
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = c0[k * kNThreads];
    float t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = tanhf(w->bh[c] + pn_dot4<64>(w->wh + c * 64, x));
    const float tn = fmaxf(sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), 1e-12f);
    p.pred[3 * s] = t[0] / tn, p.pred[3 * s + 1] = t[1] / tn, p.pred[3 * s + 2] = t[2] / tn;
  }
}

// losses.py orientation_loss / pred_normal_loss per ray, means over the batch times the multipliers; d pred_normal_loss / d pred
__global__ void k_normal_losses(const float* __restrict__ w, const float* __restrict__ normals, const float* __restrict__ pred,
                                const float* __restrict__ dirs, int64_t N, int S, float s_orient, float s_pn,
                                float* __restrict__ loss_orient, float* __restrict__ loss_pn, float* __restrict__ gpred) {
  float lo = 0.f, lp = 0.f;
  for (int64_t ray = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ray < N; ray += (int64_t)gridDim.x * blockDim.x) {
    const float vx = -dirs[3 * ray], vy = -dirs[3 * ray + 1], vz = -dirs[3 * ray + 2];
    for (int i = 0; i < S; ++i) {
      const int64_t s = ray * S + i;
      const float wi = w[s];
      const float nx = normals[3 * s], ny = normals[3 * s + 1], nz = normals[3 * s + 2];
      const float ndv = fminf(0.f, nx * vx + ny * vy + nz * vz);
      lo += wi * ndv * ndv;
      lp += wi * (1.f - (nx * pred[3 * s] + ny * pred[3 * s + 1] + nz * pred[3 * s + 2]));
      gpred[3 * s] = -s_pn * wi * nx, gpred[3 * s + 1] = -s_pn * wi * ny, gpred[3 * s + 2] = -s_pn * wi * nz;
    }
  }
  __shared__ float red[2][32];
  for (int o = 16; o > 0; o >>= 1) {
    lo += __shfl_xor_sync(0xffffffffu, lo, o);
    lp += __shfl_xor_sync(0xffffffffu, lp, o);
  }
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = lo, red[1][threadIdx.x >> 5] = lp;
  __syncthreads();
  if (threadIdx.x < 32) {
    float a = threadIdx.x < (blockDim.x >> 5) ? red[0][threadIdx.x] : 0.f, b = threadIdx.x < (blockDim.x >> 5) ? red[1][threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (threadIdx.x == 0) {
      atomicAdd(loss_orient, a * s_orient);
      atomicAdd(loss_pn, b * s_pn);
    }
  }
}

struct NormalsBwd {
  GridDev grid;
  const MlpF32* w32;
  const PnParams* pn;
  NormalRays rays;
  const float* gpred;   // [N,S,3], or null in the fused step (then derived from the weights and the analytic normals)
  // fused step (sgn_train_normals_step): the analytic normals, both loss terms and d loss / d pred are evaluated here, in
  // the same pass as the backward - the separate forward + loss kernels gather and run the base MLP a second time
  const float* weights; // [N,S] detached final weights, or null
  float s_orient, s_pn; // multipliers / N
  float* loss_orient;
  float* loss_pn;
  float* ggeo;          // [N,S,15] out: d loss / d geo features
  float* acts;          // [kPnActs][cap]
  float* deltas;        // [kPnDeltas][cap]
  int64_t first, count, cap;
};

template <bool kFused>
__global__ void __launch_bounds__(kNThreads) k_normals_bwd(const __grid_constant__ NormalsBwd p) {
  extern __shared__ __align__(16) unsigned char smem[];
  PnParams* w = reinterpret_cast<PnParams*>(smem);
  BaseW* sb = reinterpret_cast<BaseW*>(smem + sizeof(PnParams));
  for (int i = threadIdx.x; i < kPnParams; i += blockDim.x) reinterpret_cast<float*>(w)[i] = reinterpret_cast<const float*>(p.pn)[i];
  load_base(p.w32, sb);
  __syncthreads();
  const NormalRays& r = p.rays;
  float lo = 0.f, lp = 0.f;   // fused: this thread's share of the two loss terms
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < p.count; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t s = p.first + e;
    const int64_t ray = s / r.S;
#define PN_AT(base, i) (base)[(size_t)(i) * p.cap + e]
    float* a_inp = p.acts;
    float* a_l0 = p.acts + (size_t)27 * p.cap;
    float* a_l1 = p.acts + (size_t)91 * p.cap;
    float* a_x3 = p.acts + (size_t)155 * p.cap;
    float* d_l0 = p.deltas;
    float* d_l1 = p.deltas + (size_t)64 * p.cap;
    float* d_x3 = p.deltas + (size_t)128 * p.cap;
    float* d_h = p.deltas + (size_t)192 * p.cap;
    float raw[3], px, py, pz;
    raw_and_contracted(r, ray, (int)(s - ray * r.S), raw, px, py, pz);
    float inp[27];
    float nrm[3] = {0.f, 0.f, 0.f};   // fused: analytic normal
    {
      float feat[32], out1[16];
      encode_all(p.grid, px, py, pz, feat);
      if (kFused) {
        float gfeat[32], gp[3];
        base_mlp<true>(sb, feat, out1, gfeat);
        position_gradient(p.grid, px, py, pz, gfeat, gp);
        const float gn = fmaxf(sqrtf(gp[0] * gp[0] + gp[1] * gp[1] + gp[2] * gp[2]), 1e-12f);
        nrm[0] = -gp[0] / gn, nrm[1] = -gp[1] / gn, nrm[2] = -gp[2] / gn;
      } else {
        base_mlp<false>(sb, feat, out1, nullptr);
      }
      nerf_enc12(raw, inp);
#pragma unroll
      for (int j = 0; j < 15; ++j) inp[12 + j] = out1[1 + j];
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) PN_AT(a_inp, k) = inp[k];
    // forward, layer outputs in the activation slab (read back as the next layer's registers)
    for (int n = 0; n < 64; ++n) {
      float a = w->b0[n];
#pragma unroll
      for (int k = 0; k < 27; ++k) a = fmaf(w->w0[n * 27 + k], inp[k], a);
      PN_AT(a_l0, n) = fmaxf(a, 0.f);
    }
    float x[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = PN_AT(a_l0, k);
    for (int n = 0; n < 64; ++n) PN_AT(a_l1, n) = fmaxf(w->b1[n] + pn_dot4<64>(w->w1 + n * 64, x), 0.f);
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = PN_AT(a_l1, k);
    for (int n = 0; n < 64; ++n) PN_AT(a_x3, n) = w->b2[n] + pn_dot4<64>(w->w2 + n * 64, x);
#pragma unroll
    for (int k = 0; k < 64; ++k) x[k] = PN_AT(a_x3, k);
    float t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = tanhf(w->bh[c] + pn_dot4<64>(w->wh + c * 64, x));
    // backward: pred = t / |t| -> tanh -> head -> x3 -> l1 -> l0 -> inp
    const float tn = fmaxf(sqrtf(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]), 1e-12f);
    const float ph[3] = {t[0] / tn, t[1] / tn, t[2] / tn};
    float g[3];
    if (kFused) {   // losses.py orientation_loss / pred_normal_loss on the detached weights; d pn / d pred = -w n
      const float wi = p.weights[s];
      const float ndv = fminf(0.f, -(nrm[0] * __ldg(r.dirs + 3 * ray) + nrm[1] * __ldg(r.dirs + 3 * ray + 1) + nrm[2] * __ldg(r.dirs + 3 * ray + 2)));
      lo += wi * ndv * ndv;
      lp += wi * (1.f - (nrm[0] * ph[0] + nrm[1] * ph[1] + nrm[2] * ph[2]));
      g[0] = -p.s_pn * wi * nrm[0], g[1] = -p.s_pn * wi * nrm[1], g[2] = -p.s_pn * wi * nrm[2];
    } else {
      g[0] = p.gpred[3 * s], g[1] = p.gpred[3 * s + 1], g[2] = p.gpred[3 * s + 2];
    }
    const float pg = ph[0] * g[0] + ph[1] * g[1] + ph[2] * g[2];
    float gh[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      gh[c] = (g[c] - ph[c] * pg) / tn * (1.f - t[c] * t[c]);
      PN_AT(d_h, c) = gh[c];
    }
    PN_AT(d_h, 3) = 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) {   // d x3 (no activation behind layer 2)
      x[k] = w->wh[k] * gh[0] + w->wh[64 + k] * gh[1] + w->wh[128 + k] * gh[2];
      PN_AT(d_x3, k) = x[k];
    }
    float acc[64];
#pragma unroll
    for (int k = 0; k < 64; ++k) acc[k] = 0.f;
#pragma unroll 1
    for (int n = 0; n < 64; n += 4) {
      const float v0 = PN_AT(d_x3, n), v1 = PN_AT(d_x3, n + 1), v2 = PN_AT(d_x3, n + 2), v3 = PN_AT(d_x3, n + 3);
      pn_axpy4<64>(w->w2 + n * 64, v0, acc), pn_axpy4<64>(w->w2 + n * 64 + 64, v1, acc);
      pn_axpy4<64>(w->w2 + n * 64 + 128, v2, acc), pn_axpy4<64>(w->w2 + n * 64 + 192, v3, acc);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) PN_AT(d_l1, k) = PN_AT(a_l1, k) > 0.f ? acc[k] : 0.f;
#pragma unroll
    for (int k = 0; k < 64; ++k) acc[k] = 0.f;
#pragma unroll 1
    for (int n = 0; n < 64; n += 4) {
      const float v0 = PN_AT(d_l1, n), v1 = PN_AT(d_l1, n + 1), v2 = PN_AT(d_l1, n + 2), v3 = PN_AT(d_l1, n + 3);
      pn_axpy4<64>(w->w1 + n * 64, v0, acc), pn_axpy4<64>(w->w1 + n * 64 + 64, v1, acc);
      pn_axpy4<64>(w->w1 + n * 64 + 128, v2, acc), pn_axpy4<64>(w->w1 + n * 64 + 192, v3, acc);
    }
#pragma unroll
    for (int k = 0; k < 64; ++k) PN_AT(d_l0, k) = PN_AT(a_l0, k) > 0.f ? acc[k] : 0.f;
    float gi[15];
#pragma unroll
    for (int j = 0; j < 15; ++j) gi[j] = 0.f;
    for (int n = 0; n < 64; ++n) {
      const float dn = PN_AT(d_l0, n);
#pragma unroll
      for (int j = 0; j < 15; ++j) gi[j] = fmaf(w->w0[n * 27 + 12 + j], dn, gi[j]);
    }
#pragma unroll
    for (int j = 0; j < 15; ++j) p.ggeo[15 * s + j] = gi[j];
#undef PN_AT
  }
  if (kFused) {   // block sums of the two loss terms (reported scalars: the order of the atomics is not fixed)
    __shared__ float red[2][kNThreads / 32];
    for (int o = 16; o > 0; o >>= 1) {
      lo += __shfl_xor_sync(0xffffffffu, lo, o);
      lp += __shfl_xor_sync(0xffffffffu, lp, o);
    }
    if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = lo, red[1][threadIdx.x >> 5] = lp;
    __syncthreads();
    if (threadIdx.x == 0) {
      float a = 0.f, b = 0.f;
      for (int i = 0; i < kNThreads / 32; ++i) a += red[0][i], b += red[1][i];
      if (a != 0.f) atomicAdd(p.loss_orient, a * p.s_orient);
      if (b != 0.f) atomicAdd(p.loss_pn, b * p.s_pn);
    }
  }
}

static int blocks_n(int64_t n, int threads, int per_sm) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, (int64_t)sm_count() * per_sm));
}
constexpr size_t kNormSmemFwd = sizeof(PnParams) + sizeof(BaseW) + (size_t)2 * 64 * kNThreads * sizeof(float);
constexpr size_t kNormSmemBwd = sizeof(PnParams) + sizeof(BaseW);
constexpr int64_t kNormChunk = 1 << 18;

}  // namespace sgn

using namespace sgn;

extern "C" int64_t sgn_pred_normals_param_count(void) { return kPnParams; }

static int check_normals_args(const SgnField* f, const float* pn, const float* o, const float* d, int64_t N, const float* bins, int S) {
  SGN_CHECK_ARG(f != nullptr && pn != nullptr, "null field / parameter block");
  SGN_CHECK_ARG(N >= 0 && S >= 1 && S <= 1024, "bad batch shape");
  SGN_CHECK_ARG(N == 0 || (o && d && bins), "null rays / bins");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(pn) & 15) == 0, "parameter block must be 16-byte aligned");
  return SGN_OK;
}

extern "C" int sgn_train_normals_forward(const SgnField* f, const float* d_pn_params, const float* d_origins,
                                         const float* d_directions, int64_t N, const float* d_ray_bins, int S, float* d_normals,
                                         float* d_pred_normals, void* stream) {
  int rc = check_normals_args(f, d_pn_params, d_origins, d_directions, N, d_ray_bins, S);
  if (rc) return rc;
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_normals && d_pred_normals, "null output");
  static bool attr = false;
  if (!attr) {
    SGN_CUDA(cudaFuncSetAttribute(k_normals_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNormSmemFwd));
    attr = true;
  }
  NormalsFwd p;
  p.grid = f->grid, p.w32 = f->d_f32, p.pn = reinterpret_cast<const PnParams*>(d_pn_params);
  p.rays = NormalRays{d_origins, d_directions, d_ray_bins, N, S};
  p.normals = d_normals, p.pred = d_pred_normals;
  k_normals_fwd<<<blocks_n(N * S, kNThreads, 2), kNThreads, kNormSmemFwd, reinterpret_cast<cudaStream_t>(stream)>>>(p);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_normal_losses(const float* d_weights, const float* d_normals, const float* d_pred_normals,
                                 const float* d_directions, int64_t N, int S, float orientation_mult, float pred_normal_mult,
                                 float* d_loss_orientation, float* d_loss_pred_normal, float* d_grad_pred, void* stream) {
  SGN_CHECK_ARG(N >= 0 && S >= 1, "bad shape");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_weights && d_normals && d_pred_normals && d_directions && d_loss_orientation && d_loss_pred_normal && d_grad_pred,
                "null pointer");
  k_normal_losses<<<blocks_n(N, 128, 8), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_weights, d_normals, d_pred_normals, d_directions, N, S, orientation_mult / (float)N, pred_normal_mult / (float)N,
      d_loss_orientation, d_loss_pred_normal, d_grad_pred);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int64_t sgn_train_normals_ws_bytes(int64_t N, int S) {
  if (N <= 0 || S <= 0) return 0;
  return std::min(N * S, kNormChunk) * (kPnActs + kPnDeltas) * (int64_t)sizeof(float) + 256;
}

static int normals_backward_impl(const SgnField* f, const float* d_pn_params, const float* d_origins, const float* d_directions,
                                 int64_t N, const float* d_ray_bins, int S, const float* d_grad_pred, const float* d_weights,
                                 float orientation_mult, float pred_normal_mult, float* d_loss_orientation,
                                 float* d_loss_pred_normal, float* d_grad_pn_params, float* d_grad_geo, void* d_ws,
                                 int64_t ws_bytes, void* stream) {
  int rc = check_normals_args(f, d_pn_params, d_origins, d_directions, N, d_ray_bins, S);
  if (rc) return rc;
  if (N == 0) return SGN_OK;
  const bool fused = d_grad_pred == nullptr;
  SGN_CHECK_ARG(d_grad_pn_params && d_grad_geo && d_ws, "null pointer");
  SGN_CHECK_ARG(!fused || (d_weights && d_loss_orientation && d_loss_pred_normal), "null pointer");
  SGN_CHECK_ARG(ws_bytes >= sgn_train_normals_ws_bytes(N, S) && (reinterpret_cast<uintptr_t>(d_ws) & 15) == 0,
                "workspace smaller than sgn_train_normals_ws_bytes or misaligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static bool attr = false;
  if (!attr) {
    SGN_CUDA(cudaFuncSetAttribute(k_normals_bwd<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNormSmemBwd));
    SGN_CUDA(cudaFuncSetAttribute(k_normals_bwd<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNormSmemBwd));
    attr = true;
  }
  const int64_t samples = N * S, cap = std::min(samples, kNormChunk);
  float* acts = reinterpret_cast<float*>(d_ws);
  float* deltas = acts + cap * kPnActs;
  PnParams* G = reinterpret_cast<PnParams*>(d_grad_pn_params);
  for (int64_t first = 0; first < samples; first += kNormChunk) {
    NormalsBwd p;
    p.grid = f->grid, p.w32 = f->d_f32, p.pn = reinterpret_cast<const PnParams*>(d_pn_params);
    p.rays = NormalRays{d_origins, d_directions, d_ray_bins, N, S};
    p.gpred = d_grad_pred, p.weights = d_weights, p.s_orient = orientation_mult / (float)N, p.s_pn = pred_normal_mult / (float)N;
    p.loss_orient = d_loss_orientation, p.loss_pn = d_loss_pred_normal;
    p.ggeo = d_grad_geo, p.acts = acts, p.deltas = deltas;
    p.first = first, p.count = std::min(kNormChunk, samples - first), p.cap = cap;
    if (fused) k_normals_bwd<true><<<blocks_n(p.count, kNThreads, 8), kNThreads, kNormSmemBwd, st>>>(p);
    else k_normals_bwd<false><<<blocks_n(p.count, kNThreads, 8), kNThreads, kNormSmemBwd, st>>>(p);
    SGN_LAUNCH_CHECK();
    OuterParams op;
    //             deltas (offset, N)  activations (offset, K)  ld
    op.layer[0] = {0, 64, 0, 27, 27, G->w0, G->b0};
    op.layer[1] = {64, 64, 27, 64, 64, G->w1, G->b1};
    op.layer[2] = {128, 64, 91, 64, 64, G->w2, G->b2};
    op.layer[3] = {192, 3, 155, 64, 64, G->wh, G->bh};
    op.layer[4] = op.layer[3];
    launch_outer_reduce(deltas, acts, op, 4, p.count, cap, st);
    SGN_LAUNCH_CHECK();
  }
  return SGN_OK;
}

extern "C" int sgn_train_normals_backward(const SgnField* f, const float* d_pn_params, const float* d_origins,
                                          const float* d_directions, int64_t N, const float* d_ray_bins, int S,
                                          const float* d_grad_pred, float* d_grad_pn_params, float* d_grad_geo, void* d_ws,
                                          int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(N == 0 || d_grad_pred != nullptr, "null gradient");
  return normals_backward_impl(f, d_pn_params, d_origins, d_directions, N, d_ray_bins, S, d_grad_pred, nullptr, 0.f, 0.f, nullptr,
                               nullptr, d_grad_pn_params, d_grad_geo, d_ws, ws_bytes, stream);
}

extern "C" int sgn_train_normals_step(const SgnField* f, const float* d_pn_params, const float* d_origins, const float* d_directions,
                                      int64_t N, const float* d_ray_bins, int S, const float* d_weights, float orientation_mult,
                                      float pred_normal_mult, float* d_loss_orientation, float* d_loss_pred_normal,
                                      float* d_grad_pn_params, float* d_grad_geo, void* d_ws, int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(N == 0 || d_weights != nullptr, "null weights");
  return normals_backward_impl(f, d_pn_params, d_origins, d_directions, N, d_ray_bins, S, nullptr, d_weights, orientation_mult,
                               pred_normal_mult, d_loss_orientation, d_loss_pred_normal, d_grad_pn_params, d_grad_geo, d_ws,
                               ws_bytes, stream);
}
