// K1 — fused reference-sheet renderer: ray-gen -> contraction -> hash-grid gather/trilerp -> field MLPs ->
// alpha-composite (last-sample background) + median depth, for V views per launch.
//
// Replaces (for all V views at once) reference DatasetGenerator.render_camera's
//   camera.generate_rays(...)                               signerf/datasetgenerator/datasetgenerator.py:691
//   graph.get_outputs_for_camera_ray_bundle(bundle)          signerf/datasetgenerator/datasetgenerator.py:694
// i.e. nerfstudio 1.0.x NerfactoModel.get_outputs in eval.
//
// Mapping: one warp owns an 8x4-pixel tile and marches all 32 rays in lock step, so the
// 32 lanes of every gather instruction hit neighbouring grid cells (shared 32 B sectors) instead of 32 cells
// strung along one ray.  The MLPs run as mma.sync m16n8k16 chains on register fragments; weights are staged
// once per CTA into shared memory by the TMA unit (cp.async.bulk + mbarrier).
#include <algorithm>
#include <cmath>
#include <vector>

#include "sgn_device.cuh"

namespace sgn {

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kTileW = 8, kTileH = 4;
constexpr int kMaxBins = 1024 + 1;

struct RenderParams {
  GridDev grid;
  const MlpPack* pack;
  const MlpF32* f32;
  RaySource src;
  const float* bins;      // [S+1] euclidean edges shared by all rays, or
  const float* ray_bins;  // [V*H*W, S+1] per-ray euclidean edges (cascade)
  int V, H, W, S;
  int tiles_x, tiles_y, num_tiles;
  float* rgb;
  float* depth;
  float* acc;
  // near-tie refinement of the median depth (fp16-MMA path): pixels whose pick sits within `refine_delta` of moving to a
  // neighbouring bin are appended here and re-marched with the fp32 density by k_refine_depth
  uint32_t* refine_list;
  uint32_t* refine_count;
  uint32_t* refine_flags;   // one bit per pixel, set by the march; k_refine_compact turns it into the (locally sorted) list
  float refine_delta;
};

struct TileCoord {
  int v, px, py;  // pixel of this lane inside view v (clamped), and validity
  bool valid;
};

__device__ __forceinline__ TileCoord tile_pixel(const RenderParams& p, int tile, int row) {
  int per_view = p.tiles_x * p.tiles_y;
  TileCoord c;
  c.v = tile / per_view;
  int r = tile - c.v * per_view;
  int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
  int x, y;
  tile_xy(p.src, tx, ty, row, x, y);
  c.valid = (x < p.W) & (y < p.H);
  c.px = min(x, p.W - 1);
  c.py = min(y, p.H - 1);
  return c;
}

__device__ __forceinline__ void store_pixel(const RenderParams& p, const TileCoord& c, Composite& comp, float last_mid) {
  if (!c.valid) return;
  float o[3], d;
  comp.finish(last_mid, o, d);
  size_t pix = ((size_t)c.v * p.H + c.py) * p.W + c.px;
  p.rgb[pix * 3 + 0] = o[0];
  p.rgb[pix * 3 + 1] = o[1];
  p.rgb[pix * 3 + 2] = o[2];
  p.depth[pix] = d;
  if (p.acc) p.acc[pix] = comp.acc;
  if (p.refine_flags && comp.margin < p.refine_delta) atomicOr(p.refine_flags + (pix >> 5), 1u << (pix & 31));
}

// ------------------------------------------------------------------------------------------------
// Tensor-core path.  Dynamic smem: [MlpPack][stage: kWarps x 32 x kStageStride halfs][mbar][bins S+1]
template <bool kPerRayBins>
__global__ void __launch_bounds__(kThreads, 4) k_render_mma(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  MlpPack* sp = reinterpret_cast<MlpPack*>(smem);
  __half* stage_all = reinterpret_cast<__half*>(smem + sizeof(MlpPack));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sizeof(MlpPack) + kWarps * 32 * kStageStride * sizeof(__half));
  float* sbins = reinterpret_cast<float*>(bar + 2);

  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)sizeof(MlpPack));
    bulk_g2s(sp, p.pack, (uint32_t)sizeof(MlpPack), bar);
  }
  if (!kPerRayBins)
    for (int i = threadIdx.x; i <= p.S; i += kThreads) sbins[i] = p.bins[i];
  mbar_wait(bar, 0);
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __half* stage = stage_all + warp * 32 * kStageStride;
  const int rc = (lane >> 2) + 8 * (lane & 3);  // sample row this lane composites
  const float avg = sp->avg_density;

  for (int tile = blockIdx.x * kWarps + warp; tile < p.num_tiles; tile += gridDim.x * kWarps) {
    const TileCoord cf = tile_pixel(p, tile, lane);  // pixel whose features this lane computes
    const TileCoord cc = tile_pixel(p, tile, rc);    // pixel this lane composites
    float ro[3], d[3];
    load_ray(p.src, cf.v, cf.px, cf.py, ro, d);
    uint32_t ash[2][4];
    {
      float sh[16];
      sh16(d[0], d[1], d[2], sh);
      warp_stage_sh(stage, lane, sh, ash);
    }
    const float* fb = nullptr;
    const float* cb = nullptr;
    if (kPerRayBins) {
      fb = p.ray_bins + (((size_t)cf.v * p.H + cf.py) * p.W + cf.px) * (size_t)(p.S + 1);
      cb = p.ray_bins + (((size_t)cc.v * p.H + cc.py) * p.W + cc.px) * (size_t)(p.S + 1);
    }
    Composite comp;
    float last_mid = 0.f;
    float f_t0 = kPerRayBins ? __ldg(fb) : 0.f, c_t0 = kPerRayBins ? __ldg(cb) : 0.f;
    for (int i = 0; i < p.S; ++i) {
      float tf0, tf1, tc0, tc1;
      if (kPerRayBins) {
        tf0 = f_t0; tf1 = __ldg(fb + i + 1); f_t0 = tf1;
        tc0 = c_t0; tc1 = __ldg(cb + i + 1); c_t0 = tc1;
      } else {
        tf0 = tc0 = sbins[i];
        tf1 = tc1 = sbins[i + 1];
      }
      const float fmid = __fmul_rn(__fadd_rn(tf0, tf1), 0.5f);
      float px, py, pz;
      const bool sel = contract_to_unit(__fadd_rn(ro[0], __fmul_rn(d[0], fmid)),
                                        __fadd_rn(ro[1], __fmul_rn(d[1], fmid)),
                                        __fadd_rn(ro[2], __fmul_rn(d[2], fmid)), px, py, pz);
      float logit, cr, cg, cbv;
      warp_field_eval(p.grid, sp, stage, lane, px, py, pz, ash, logit, cr, cg, cbv);
      const bool sel_c = __shfl_sync(0xffffffffu, (int)sel, rc) != 0;
      const float sigma = sel_c ? avg * expf(logit) : 0.f;
      const float cmid = __fmul_rn(__fadd_rn(tc0, tc1), 0.5f);
      comp.step(sigma, __fsub_rn(tc1, tc0), cmid, sigmoidf_(cr), sigmoidf_(cg), sigmoidf_(cbv));
      last_mid = cmid;
    }
    store_pixel(p, cc, comp, last_mid);
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core parity path: one lane = one ray, nothing crosses lanes.
template <bool kPerRayBins>
__global__ void __launch_bounds__(kThreads) k_render_f32(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  MlpF32* sw = reinterpret_cast<MlpF32*>(smem);
  float* sbins = reinterpret_cast<float*>(smem + sizeof(MlpF32));
  for (int i = threadIdx.x; i < (int)(sizeof(MlpF32) / 16); i += kThreads)
    reinterpret_cast<uint4*>(sw)[i] = reinterpret_cast<const uint4*>(p.f32)[i];
  if (!kPerRayBins)
    for (int i = threadIdx.x; i <= p.S; i += kThreads) sbins[i] = p.bins[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int tile = blockIdx.x * kWarps + warp; tile < p.num_tiles; tile += gridDim.x * kWarps) {
    const TileCoord c = tile_pixel(p, tile, lane);
    float ro[3], d[3], sh[16];
    load_ray(p.src, c.v, c.px, c.py, ro, d);
    sh16(d[0], d[1], d[2], sh);
    const float* rb = kPerRayBins ? p.ray_bins + (((size_t)c.v * p.H + c.py) * p.W + c.px) * (size_t)(p.S + 1) : nullptr;
    Composite comp;
    float last_mid = 0.f;
    for (int i = 0; i < p.S; ++i) {
      const float t0 = kPerRayBins ? __ldg(rb + i) : sbins[i];
      const float t1 = kPerRayBins ? __ldg(rb + i + 1) : sbins[i + 1];
      const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
      float px, py, pz;
      const bool sel = contract_to_unit(__fadd_rn(ro[0], __fmul_rn(d[0], mid)),
                                        __fadd_rn(ro[1], __fmul_rn(d[1], mid)),
                                        __fadd_rn(ro[2], __fmul_rn(d[2], mid)), px, py, pz);
      float feat[32];
#pragma unroll
      for (int l = 0; l < 16; ++l) {
        float2 f = encode_level(p.grid.table + (size_t)l * p.grid.size, p.grid.mask, p.grid.res[l], px, py, pz);
        feat[2 * l] = f.x;
        feat[2 * l + 1] = f.y;
      }
      float logit, c3[3];
      field_mlp_f32(sw, feat, sh, logit, c3);
      const float sigma = sel ? sw->avg_density * expf(logit) : 0.f;
      comp.step(sigma, __fsub_rn(t1, t0), mid, sigmoidf_(c3[0]), sigmoidf_(c3[1]), sigmoidf_(c3[2]));
      last_mid = mid;
    }
    store_pixel(p, c, comp, last_mid);
  }
}

// ------------------------------------------------------------------------------------------------
// Median depth is a discrete pick: the mid-point of the first bin whose cumulative weight reaches one half.  The fp16
// tensor-core MLP carries the density logit to ~1e-3, the cumulative weight to a few 1e-4, so a ray whose pick sits that
// close to a bin boundary can land in the neighbouring bin - a whole bin of depth error (4 % at 128 bins), which a few
// rays in a thousand are enough to push the image's relative L2 past the 1e-3 contract.  k_render_mma lists those rays
// (margin < delta); this kernel re-marches ONLY them with the density in fp32 (hash features + base MLP on CUDA cores,
// the same arithmetic as k_render_f32; colour is not needed) and rewrites their depth.  One lane per listed ray.
struct DensityF32 {
  float w0[64 * 32];
  float b0[64];
  float w1[64];   // row 0 of base layer 1: the density logit
  float b1;
  float avg;
};

// Flag bitmap -> list of pixel ids.  Every block turns 1 024 words (32 768 consecutive pixels) into one contiguous,
// ascending run of the list, so the 32 lanes of a k_refine_depth warp re-march pixels that are neighbours in the image:
// their gathers share cache lines.  (Appending from the march with one atomic per ray left the list in random order - every
// gather its own 32-byte sector: 3.7 ms per million rays in cascade mode, where ~10 % of the rays are listed.)
__global__ void __launch_bounds__(1024) k_refine_compact(const uint32_t* __restrict__ flags, uint32_t num_words,
                                                          uint32_t* __restrict__ count, uint32_t* __restrict__ list) {
  __shared__ uint32_t warp_sum[32];
  __shared__ uint32_t base;
  const uint32_t w = blockIdx.x * 1024u + threadIdx.x;
  uint32_t bits = w < num_words ? flags[w] : 0u;
  const uint32_t cnt = __popc(bits);
  uint32_t incl = cnt;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += u;
  }
  if (lane == 31) warp_sum[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = warp_sum[lane], t = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += u;
    }
    warp_sum[lane] = t - v;                       // exclusive prefix of the warps
    if (lane == 31) base = t ? atomicAdd(count, t) : 0u;
  }
  __syncthreads();
  uint32_t at = base + warp_sum[wid] + incl - cnt;
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    list[at++] = (w << 5) + (uint32_t)b;
  }
}

template <bool kPerRayBins>
__global__ void __launch_bounds__(kThreads) k_refine_depth(const __grid_constant__ RenderParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  DensityF32* sw = reinterpret_cast<DensityF32*>(smem);
  float* sbins = reinterpret_cast<float*>(smem + sizeof(DensityF32));
  for (int i = threadIdx.x; i < 64 * 32; i += kThreads) sw->w0[i] = p.f32->w_base0[i];
  for (int i = threadIdx.x; i < 64; i += kThreads) {
    sw->b0[i] = p.f32->b_base0[i];
    sw->w1[i] = p.f32->w_base1[i];
  }
  if (threadIdx.x == 0) {
    sw->b1 = p.f32->b_base1[0];
    sw->avg = p.f32->avg_density;
  }
  if (!kPerRayBins)
    for (int i = threadIdx.x; i <= p.S; i += kThreads) sbins[i] = p.bins[i];
  __syncthreads();
  const uint32_t n = min(*p.refine_count, (uint32_t)((size_t)p.V * p.H * p.W));
  for (uint32_t e = blockIdx.x * kThreads + threadIdx.x; e < n; e += gridDim.x * kThreads) {
    const uint32_t pix = p.refine_list[e];
    const int v = (int)(pix / ((uint32_t)p.H * p.W));
    const uint32_t r = pix - (uint32_t)v * p.H * p.W;
    const int y = (int)(r / p.W), x = (int)(r - (uint32_t)y * p.W);
    float ro[3], d[3];
    load_ray(p.src, v, x, y, ro, d);
    const float* rb = kPerRayBins ? p.ray_bins + (size_t)pix * (size_t)(p.S + 1) : nullptr;
    float cum_dd = 0.f, cumw = 0.f, depth = 0.f, last_mid = 0.f;
    bool found = false;
    for (int i = 0; i < p.S && !found; ++i) {
      const float t0 = kPerRayBins ? __ldg(rb + i) : sbins[i];
      const float t1 = kPerRayBins ? __ldg(rb + i + 1) : sbins[i + 1];
      const float mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
      float px, py, pz;
      const bool sel = contract_to_unit(__fadd_rn(ro[0], __fmul_rn(d[0], mid)), __fadd_rn(ro[1], __fmul_rn(d[1], mid)),
                                        __fadd_rn(ro[2], __fmul_rn(d[2], mid)), px, py, pz);
      float sigma = 0.f;
      if (sel) {
        float feat[32];
#pragma unroll
        for (int l = 0; l < 16; ++l) {
          float2 f = encode_level(p.grid.table + (size_t)l * p.grid.size, p.grid.mask, p.grid.res[l], px, py, pz);
          feat[2 * l] = f.x;
          feat[2 * l + 1] = f.y;
        }
        // four neurons at a time, weights as LDS.128 (one load per four FMAs instead of one per FMA); every neuron's sum and
        // the logit keep k_render_f32's order of operations, so the result is bit for bit the fp32 path's
        float logit = sw->b1;
#pragma unroll 1
        for (int nn = 0; nn < 64; nn += 4) {
          float a0 = sw->b0[nn], a1 = sw->b0[nn + 1], a2 = sw->b0[nn + 2], a3 = sw->b0[nn + 3];
          const float4* r0 = reinterpret_cast<const float4*>(sw->w0 + nn * 32);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float4 u0 = r0[k], u1 = r0[8 + k], u2 = r0[16 + k], u3 = r0[24 + k];
            a0 = fmaf(u0.x, feat[4 * k], a0), a0 = fmaf(u0.y, feat[4 * k + 1], a0), a0 = fmaf(u0.z, feat[4 * k + 2], a0), a0 = fmaf(u0.w, feat[4 * k + 3], a0);
            a1 = fmaf(u1.x, feat[4 * k], a1), a1 = fmaf(u1.y, feat[4 * k + 1], a1), a1 = fmaf(u1.z, feat[4 * k + 2], a1), a1 = fmaf(u1.w, feat[4 * k + 3], a1);
            a2 = fmaf(u2.x, feat[4 * k], a2), a2 = fmaf(u2.y, feat[4 * k + 1], a2), a2 = fmaf(u2.z, feat[4 * k + 2], a2), a2 = fmaf(u2.w, feat[4 * k + 3], a2);
            a3 = fmaf(u3.x, feat[4 * k], a3), a3 = fmaf(u3.y, feat[4 * k + 1], a3), a3 = fmaf(u3.z, feat[4 * k + 2], a3), a3 = fmaf(u3.w, feat[4 * k + 3], a3);
          }
          logit = fmaf(sw->w1[nn], fmaxf(a0, 0.f), logit);
          logit = fmaf(sw->w1[nn + 1], fmaxf(a1, 0.f), logit);
          logit = fmaf(sw->w1[nn + 2], fmaxf(a2, 0.f), logit);
          logit = fmaf(sw->w1[nn + 3], fmaxf(a3, 0.f), logit);
        }
        sigma = sw->avg * expf(logit);
      }
      const float dd = __fsub_rn(t1, t0) * sigma;
      float w = (1.f - expf(-dd)) * expf(-cum_dd);
      cum_dd += dd;
      if (w != w) w = 0.f;
      cumw += w;
      if (cumw >= 0.5f) {
        found = true;
        depth = mid;
      }
      last_mid = mid;
    }
    if (!found) {   // the loop ran to the end: DepthRenderer clamps the index to the last sample
      const float t0 = kPerRayBins ? __ldg(rb + p.S - 1) : sbins[p.S - 1];
      const float t1 = kPerRayBins ? __ldg(rb + p.S) : sbins[p.S];
      last_mid = __fmul_rn(__fadd_rn(t0, t1), 0.5f);
    }
    p.depth[pix] = found ? depth : last_mid;
  }
}

// ------------------------------------------------------------------------------------------------
// Probes used by the parity tests.
__global__ void k_generate_rays(const float* __restrict__ c2w, const float* __restrict__ intr, int V, int H, int W,
                                float* __restrict__ origins, float* __restrict__ dirs, float* __restrict__ area,
                                float* __restrict__ dnorm) {
  size_t n = (size_t)V * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int v = (int)(i / ((size_t)H * W));
    int r = (int)(i - (size_t)v * H * W);
    int y = r / W, x = r - y * W;
    Camera cam = load_camera(c2w, intr, v);
    float d0[3], dx[3], dy[3];
    float xf = (float)x + 0.5f, yf = (float)y + 0.5f;
    float n0 = ray_direction(cam, xf, yf, 0.f, 0.f, d0);
    origins[3 * i + 0] = cam.o[0]; origins[3 * i + 1] = cam.o[1]; origins[3 * i + 2] = cam.o[2];
    dirs[3 * i + 0] = d0[0]; dirs[3 * i + 1] = d0[1]; dirs[3 * i + 2] = d0[2];
    if (dnorm) dnorm[i] = n0;
    if (area) {
      ray_direction(cam, xf, yf, 1.f, 0.f, dx);
      ray_direction(cam, xf, yf, 0.f, 1.f, dy);
      float ax = 0.f, ay = 0.f;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float ex = __fsub_rn(d0[k], dx[k]), ey = __fsub_rn(d0[k], dy[k]);
        ax = __fadd_rn(ax, __fmul_rn(ex, ex));
        ay = __fadd_rn(ay, __fmul_rn(ey, ey));
      }
      area[i] = __fmul_rn(__fsqrt_rn(ax), __fsqrt_rn(ay));
    }
  }
}

__global__ void k_hash_encode(const GridDev grid, const float* __restrict__ pos, int64_t N,
                              long long* __restrict__ indices, float* __restrict__ feats) {
  const int L = grid.num_levels;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N * L; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t n = i / L;
    int l = (int)(i - n * L);
    float px = pos[3 * n], py = pos[3 * n + 1], pz = pos[3 * n + 2];
    if (indices) {
      LevelCoords lc = level_coords(grid.res[l], px, py, pz);
      uint32_t idx[8];
      corner_rows(lc, grid.mask, idx);
#pragma unroll
      for (int c = 0; c < 8; ++c) indices[i * 8 + c] = (long long)idx[c] + (long long)l * grid.size;
    }
    if (feats) {
      float2 f = encode_level(grid.table + (size_t)l * grid.size, grid.mask, grid.res[l], px, py, pz);
      feats[n * 2 * L + 2 * l] = f.x;
      feats[n * 2 * L + 2 * l + 1] = f.y;
    }
  }
}

// One sample per lane; N padded to full warps by the host (invalid lanes clamp to the last sample).
__global__ void __launch_bounds__(kThreads) k_field_eval_mma(const GridDev grid, const MlpPack* __restrict__ pack,
                                                             const float* __restrict__ pos,
                                                             const float* __restrict__ dir, int64_t N,
                                                             float* __restrict__ density, float* __restrict__ rgb) {
  extern __shared__ __align__(128) unsigned char smem[];
  MlpPack* sp = reinterpret_cast<MlpPack*>(smem);
  __half* stage_all = reinterpret_cast<__half*>(smem + sizeof(MlpPack));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + sizeof(MlpPack) + kWarps * 32 * kStageStride * sizeof(__half));
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)sizeof(MlpPack));
    bulk_g2s(sp, pack, (uint32_t)sizeof(MlpPack), bar);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __half* stage = stage_all + warp * 32 * kStageStride;
  const int rc = (lane >> 2) + 8 * (lane & 3);
  const int64_t nwarps = (N + 31) / 32;
  for (int64_t w = blockIdx.x * (int64_t)kWarps + warp; w < nwarps; w += (int64_t)gridDim.x * kWarps) {
    int64_t i = min(w * 32 + lane, N - 1);
    float sh[16];
    sh16(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], sh);
    uint32_t ash[2][4];
    warp_stage_sh(stage, lane, sh, ash);
    float px, py, pz;
    bool sel = contract_to_unit(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], px, py, pz);
    float logit, cr, cg, cb;
    warp_field_eval(grid, sp, stage, lane, px, py, pz, ash, logit, cr, cg, cb);
    bool sel_c = __shfl_sync(0xffffffffu, (int)sel, rc) != 0;
    int64_t o = w * 32 + rc;
    if (o < N) {
      density[o] = sel_c ? sp->avg_density * expf(logit) : 0.f;
      rgb[3 * o + 0] = sigmoidf_(cr);
      rgb[3 * o + 1] = sigmoidf_(cg);
      rgb[3 * o + 2] = sigmoidf_(cb);
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_field_eval_f32(const GridDev grid, const MlpF32* __restrict__ w32,
                                                             const float* __restrict__ pos,
                                                             const float* __restrict__ dir, int64_t N,
                                                             float* __restrict__ density, float* __restrict__ rgb) {
  extern __shared__ __align__(128) unsigned char smem[];
  MlpF32* sw = reinterpret_cast<MlpF32*>(smem);
  for (int i = threadIdx.x; i < (int)(sizeof(MlpF32) / 16); i += kThreads)
    reinterpret_cast<uint4*>(sw)[i] = reinterpret_cast<const uint4*>(w32)[i];
  __syncthreads();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    float sh[16], feat[32], px, py, pz;
    sh16(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], sh);
    bool sel = contract_to_unit(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], px, py, pz);
#pragma unroll
    for (int l = 0; l < 16; ++l) {
      float2 f = encode_level(grid.table + (size_t)l * grid.size, grid.mask, grid.res[l], px, py, pz);
      feat[2 * l] = f.x;
      feat[2 * l + 1] = f.y;
    }
    float logit, c3[3];
    field_mlp_f32(sw, feat, sh, logit, c3);
    density[i] = sel ? sw->avg_density * expf(logit) : 0.f;
    rgb[3 * i + 0] = sigmoidf_(c3[0]);
    rgb[3 * i + 1] = sigmoidf_(c3[1]);
    rgb[3 * i + 2] = sigmoidf_(c3[2]);
  }
}

// ------------------------------------------------------------------------------------------------
// host helpers
int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// torch.linspace(0, 1, n) on CPU/CUDA for float: symmetric evaluation around the midpoint.
static void linspace01(int n, std::vector<float>& out) {
  out.resize(n);
  float step = (1.f - 0.f) / (float)(n - 1);
  int half = n / 2;
  for (int i = 0; i < n; ++i) out[i] = i < half ? 0.f + step * (float)i : 1.f - step * (float)(n - i - 1);
}
static float h_spacing(float x) { return x < 1.f ? x / 2.f : 1.f - 1.f / (2.f * x); }
static float h_spacing_inv(float x) { return x < 0.5f ? 2.f * x : 1.f / (2.f - 2.f * x); }

void host_flat_bins(int S, float near_p, float far_p, std::vector<float>& out) {
  std::vector<float> u;
  linspace01(S + 1, u);
  volatile float s_near = h_spacing(near_p), s_far = h_spacing(far_p);
  out.resize(S + 1);
  for (int i = 0; i <= S; ++i) {
    volatile float a = u[i] * s_far;
    volatile float b = (1.f - u[i]) * s_near;
    out[i] = h_spacing_inv(a + b);
  }
}

int g_render_ctas_per_sm = 4;  // sgn_set_option "render_ctas_per_sm": 1 leaves the SM mostly free for a co-resident kernel
int g_refine_depth = 1;        // sgn_set_option "render_refine_depth": fp32 re-march of near-tie median picks (fp16-MMA path)
float g_refine_delta = 1e-3f;  // sgn_set_option "render_refine_delta_ppm": margin below which a ray is re-marched

size_t mma_smem_bytes(int S, bool per_ray) {
  return sizeof(MlpPack) + kWarps * 32 * kStageStride * sizeof(__half) + 16 + (per_ray ? 0 : (size_t)(S + 1) * 4);
}
size_t f32_smem_bytes(int S, bool per_ray) { return sizeof(MlpF32) + (per_ray ? 0 : (size_t)(S + 1) * 4); }

template <class K>
static int set_smem(K kernel, size_t bytes) {
  SGN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return SGN_OK;
}

// Launch the main-field march.  d_bins: [S+1] shared edges, or d_ray_bins per ray.
int launch_render(const SgnField* f, const RaySource& src, int V, int H, int W, int S,
                  const float* d_bins, const float* d_ray_bins, int mlp_mode, float* d_rgb, float* d_depth,
                  float* d_acc, cudaStream_t st) {
  RenderParams p;
  p.grid = f->grid;
  p.pack = f->d_pack;
  p.f32 = f->d_f32;
  p.src = src;
  p.bins = d_bins;
  p.ray_bins = d_ray_bins;
  p.V = V; p.H = H; p.W = W; p.S = S;
  const int tw = 1 << src.tw_log2, th = 32 >> src.tw_log2;
  p.tiles_x = (W + tw - 1) / tw;
  p.tiles_y = (H + th - 1) / th;
  p.num_tiles = V * p.tiles_x * p.tiles_y;
  p.rgb = d_rgb; p.depth = d_depth; p.acc = d_acc;
  p.refine_list = nullptr; p.refine_count = nullptr; p.refine_flags = nullptr; p.refine_delta = 0.f;
  const bool per_ray = d_ray_bins != nullptr;
  const int blocks_needed = (p.num_tiles + kWarps - 1) / kWarps;
  if (mlp_mode == SGN_MLP_FP16_MMA) {
    uint32_t* d_list = nullptr;
    const size_t npix = (size_t)V * H * W;
    const size_t nwords = (npix + 31) / 32;
    if (g_refine_depth && npix < ((size_t)1 << 32)) {
      SGN_CUDA(scratch_alloc(&d_list, (npix + 1 + nwords) * sizeof(uint32_t), st));   // count | list | flag bitmap
      p.refine_count = d_list;
      p.refine_list = d_list + 1;
      p.refine_flags = d_list + 1 + npix;
      SGN_CUDA(cudaMemsetAsync(d_list, 0, sizeof(uint32_t), st));
      SGN_CUDA(cudaMemsetAsync(p.refine_flags, 0, nwords * sizeof(uint32_t), st));
      p.refine_delta = g_refine_delta;
    }
    struct FreeList {
      uint32_t* q; cudaStream_t s;
      ~FreeList() { if (q) cudaFreeAsync(q, s); }
    } free_list{d_list, st};
    size_t smem = mma_smem_bytes(S, per_ray);
    int grid = std::min(blocks_needed, sm_count() * g_render_ctas_per_sm);
    if (per_ray) {
      int rc = set_smem(k_render_mma<true>, smem);
      if (rc) return rc;
      k_render_mma<true><<<grid, kThreads, smem, st>>>(p);
    } else {
      int rc = set_smem(k_render_mma<false>, smem);
      if (rc) return rc;
      k_render_mma<false><<<grid, kThreads, smem, st>>>(p);
    }
    if (d_list) {
      SGN_LAUNCH_CHECK();
      k_refine_compact<<<(unsigned)((nwords + 1023) / 1024), 1024, 0, st>>>(p.refine_flags, (uint32_t)nwords, p.refine_count, p.refine_list);
      SGN_LAUNCH_CHECK();
      const size_t rsmem = sizeof(DensityF32) + (per_ray ? 0 : (size_t)(S + 1) * 4);
      const int rgrid = std::min((int)((npix + kThreads - 1) / kThreads), sm_count() * 8);
      if (per_ray) {
        int rc = set_smem(k_refine_depth<true>, rsmem);
        if (rc) return rc;
        k_refine_depth<true><<<rgrid, kThreads, rsmem, st>>>(p);
      } else {
        int rc = set_smem(k_refine_depth<false>, rsmem);
        if (rc) return rc;
        k_refine_depth<false><<<rgrid, kThreads, rsmem, st>>>(p);
      }
    }
  } else {
    size_t smem = f32_smem_bytes(S, per_ray);
    int grid = std::min(blocks_needed, sm_count() * 4);
    if (per_ray) {
      int rc = set_smem(k_render_f32<true>, smem);
      if (rc) return rc;
      k_render_f32<true><<<grid, kThreads, smem, st>>>(p);
    } else {
      int rc = set_smem(k_render_f32<false>, smem);
      if (rc) return rc;
      k_render_f32<false><<<grid, kThreads, smem, st>>>(p);
    }
  }
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

int render_cascade(const SgnField* f, const RaySource& src, int V, int H, int W,
                   const SgnRenderOpts* o, float* d_rgb, float* d_depth, float* d_acc, cudaStream_t st);

}  // namespace sgn

using namespace sgn;

static int render_from(const SgnField* f, const RaySource& src, int V, int H, int W, const SgnRenderOpts* o,
                       float* d_rgb, float* d_depth, float* d_acc, void* stream) {
  SGN_CHECK_ARG(o->mlp_mode == SGN_MLP_FP16_MMA || o->mlp_mode == SGN_MLP_FP32, "bad mlp_mode");
  SGN_CHECK_ARG(o->num_samples >= 1 && o->num_samples < kMaxBins, "num_samples must be 1..1024");
  SGN_CHECK_ARG(o->far_plane > o->near_plane && o->near_plane >= 0.f, "need 0 <= near < far");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (o->mode == 1) return render_cascade(f, src, V, H, W, o, d_rgb, d_depth, d_acc, st);
  SGN_CHECK_ARG(o->mode == 0, "mode must be 0 (flat) or 1 (cascade)");
  const int S = o->num_samples;
  std::vector<float> bins;
  const float* hb = o->h_bins;
  if (!hb) {
    host_flat_bins(S, o->near_plane, o->far_plane, bins);
    hb = bins.data();
  }
  float* d_bins = nullptr;
  SGN_CUDA(scratch_alloc(&d_bins, (size_t)(S + 1) * 4, st));
  SGN_CUDA(cudaMemcpyAsync(d_bins, hb, (size_t)(S + 1) * 4, cudaMemcpyHostToDevice, st));
  // the pageable source must stay alive until the copy is consumed: it is staged synchronously by the
  // runtime for pageable memory, so `bins` may go out of scope after this call returns.
  int rc = launch_render(f, src, V, H, W, S, d_bins, nullptr, o->mlp_mode, d_rgb, d_depth, d_acc, st);
  cudaFreeAsync(d_bins, st);
  return rc;
}

extern "C" int sgn_render_views(const SgnField* f, const float* d_c2w, const float* d_intr, int V, int H, int W,
                                const SgnRenderOpts* o, float* d_rgb, float* d_depth, float* d_acc, void* stream) {
  SGN_CHECK_ARG(f && o, "null field/opts");
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(V == 0 || (d_c2w && d_intr && d_rgb && d_depth), "null pointer");
  if (V == 0) return SGN_OK;
  const RaySource src{d_c2w, d_intr, nullptr, nullptr, 3};
  return render_from(f, src, V, H, W, o, d_rgb, d_depth, d_acc, stream);
}

extern "C" int sgn_render_rays(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N,
                               const SgnRenderOpts* o, float* d_rgb, float* d_depth, float* d_acc, void* stream) {
  SGN_CHECK_ARG(f && o, "null field/opts");
  SGN_CHECK_ARG(N >= 0 && N < ((int64_t)1 << 31) - 64, "ray count must be in [0, 2^31)");
  if (N == 0) return SGN_OK;
  SGN_CHECK_ARG(d_origins && d_directions && d_rgb && d_depth, "null pointer");
  // the sampling cascade keeps [rays, 257] scratch rows: bundles go through it 2^20 rays at a time
  const int64_t chunk = o->mode == 1 ? ((int64_t)1 << 20) : N;
  for (int64_t r0 = 0; r0 < N; r0 += chunk) {
    const int64_t n = std::min(chunk, N - r0);
    const RaySource src{nullptr, nullptr, d_origins + 3 * r0, d_directions + 3 * r0, 5};
    int rc = render_from(f, src, 1, 1, (int)n, o, d_rgb + 3 * r0, d_depth + r0, d_acc ? d_acc + r0 : nullptr, stream);
    if (rc) return rc;
  }
  return SGN_OK;
}

extern "C" int sgn_render_views_host(const SgnField* f, const float* h_c2w, const float* h_intr, int V, int H, int W,
                                     const SgnRenderOpts* o, float* h_rgb, float* h_depth, float* h_acc) {
  SGN_CHECK_ARG(f && h_c2w && h_intr && o && h_rgb && h_depth, "null pointer");
  SGN_CHECK_ARG(V > 0 && H > 0 && W > 0, "bad image shape");
  size_t npix = (size_t)V * H * W;
  float *d_c2w = nullptr, *d_intr = nullptr, *d_out = nullptr;
  cudaStream_t st = 0;
  SGN_CUDA(cudaMalloc(&d_c2w, (size_t)V * 12 * 4));
  SGN_CUDA(cudaMalloc(&d_intr, (size_t)V * 4 * 4));
  SGN_CUDA(cudaMalloc(&d_out, npix * 5 * 4));
  float* d_rgb = d_out;
  float* d_depth = d_out + npix * 3;
  float* d_acc = d_out + npix * 4;
  int rc = SGN_OK;
  do {
    if (cudaMemcpyAsync(d_c2w, h_c2w, (size_t)V * 12 * 4, cudaMemcpyHostToDevice, st) != cudaSuccess ||
        cudaMemcpyAsync(d_intr, h_intr, (size_t)V * 4 * 4, cudaMemcpyHostToDevice, st) != cudaSuccess) {
      set_error("H2D of cameras failed");
      rc = SGN_ERR_CUDA;
      break;
    }
    rc = sgn_render_views(f, d_c2w, d_intr, V, H, W, o, d_rgb, d_depth, h_acc ? d_acc : nullptr, st);
    if (rc) break;
    if (cudaMemcpyAsync(h_rgb, d_rgb, npix * 3 * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(h_depth, d_depth, npix * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        (h_acc && cudaMemcpyAsync(h_acc, d_acc, npix * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess)) {
      set_error("D2H of images failed");
      rc = SGN_ERR_CUDA;
      break;
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      set_error(std::string("render failed: ") + cudaGetErrorString(e));
      rc = SGN_ERR_CUDA;
    }
  } while (0);
  cudaFree(d_c2w);
  cudaFree(d_intr);
  cudaFree(d_out);
  return rc;
}

extern "C" int sgn_generate_rays(const float* d_c2w, const float* d_intr, int V, int H, int W, float* d_origins,
                                 float* d_directions, float* d_pixel_area, float* d_dir_norm, void* stream) {
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0, "bad image shape");
  if (V == 0) return SGN_OK;
  SGN_CHECK_ARG(d_c2w && d_intr && d_origins && d_directions, "null pointer");
  size_t n = (size_t)V * H * W;
  int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)sm_count() * 8);
  k_generate_rays<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_c2w, d_intr, V, H, W, d_origins,
                                                                             d_directions, d_pixel_area, d_dir_norm);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_hash_encode(const SgnField* f, int which, const float* d_pos, int64_t N, int64_t* d_indices,
                               float* d_features, void* stream) {
  SGN_CHECK_ARG(f != nullptr, "null field");
  SGN_CHECK_ARG(which >= 0 && which <= f->num_proposals, "grid index out of range");
  SGN_CHECK_ARG(N >= 0, "negative N");
  if (N == 0) return SGN_OK;  // empty input: pointers may be null
  SGN_CHECK_ARG(d_pos != nullptr, "null pointer");
  const GridDev& g = which == 0 ? f->grid : f->h_prop[which - 1].grid;
  int64_t work = N * g.num_levels;
  int blocks = (int)std::min<int64_t>((work + 255) / 256, (int64_t)sm_count() * 8);
  k_hash_encode<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      g, d_pos, N, reinterpret_cast<long long*>(d_indices), d_features);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_field_eval(const SgnField* f, const float* d_pos, const float* d_dir, int64_t N, int mlp_mode,
                              float* d_density, float* d_rgb, void* stream) {
  SGN_CHECK_ARG(f != nullptr, "null field");
  SGN_CHECK_ARG(N >= 0, "negative N");
  SGN_CHECK_ARG(mlp_mode == SGN_MLP_FP16_MMA || mlp_mode == SGN_MLP_FP32, "bad mlp_mode");
  if (N == 0) return SGN_OK;  // empty input: pointers may be null
  SGN_CHECK_ARG(d_pos && d_dir && d_density && d_rgb, "null pointer");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (mlp_mode == SGN_MLP_FP16_MMA) {
    size_t smem = mma_smem_bytes(0, true);
    int rc = set_smem(k_field_eval_mma, smem);
    if (rc) return rc;
    int64_t nw = (N + 31) / 32;
    int blocks = (int)std::min<int64_t>((nw + kWarps - 1) / kWarps, (int64_t)sm_count() * 4);
    k_field_eval_mma<<<blocks, kThreads, smem, st>>>(f->grid, f->d_pack, d_pos, d_dir, N, d_density, d_rgb);
  } else {
    size_t smem = sizeof(MlpF32);
    int rc = set_smem(k_field_eval_f32, smem);
    if (rc) return rc;
    int blocks = (int)std::min<int64_t>((N + kThreads - 1) / kThreads, (int64_t)sm_count() * 4);
    k_field_eval_f32<<<blocks, kThreads, smem, st>>>(f->grid, f->d_f32, d_pos, d_dir, N, d_density, d_rgb);
  }
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

namespace sgn { extern int g_pair_stages; extern int g_attn_variant; extern int g_attn_idle_ns; extern int g_attn_short_kv; extern int g_attn_split; extern int g_attn_shape; }

extern "C" int sgn_set_option(const char* name, int value) {
  SGN_CHECK_ARG(name != nullptr, "null option name");
  const std::string n(name);
  if (n == "render_ctas_per_sm") {
    SGN_CHECK_ARG(value >= 1 && value <= 4, "render_ctas_per_sm must be 1..4");
    sgn::g_render_ctas_per_sm = value;
    return SGN_OK;
  }
  if (n == "render_refine_depth") {
    SGN_CHECK_ARG(value == 0 || value == 1, "render_refine_depth must be 0 or 1");
    sgn::g_refine_depth = value;
    return SGN_OK;
  }
  if (n == "render_refine_delta_ppm") {
    SGN_CHECK_ARG(value >= 0 && value <= 500000, "render_refine_delta_ppm must be 0..500000");
    sgn::g_refine_delta = (float)value * 1e-6f;
    return SGN_OK;
  }
  if (n == "gemm_pair_stages") {
    SGN_CHECK_ARG(value == 5 || value == 6, "gemm_pair_stages must be 5 or 6");
    sgn::g_pair_stages = value;
    return SGN_OK;
  }
  if (n == "attn_variant") {
    SGN_CHECK_ARG(value >= 0 && value <= 3, "attn_variant must be 0..3 (bit flags, see sgn_attn.cu)");
    sgn::g_attn_variant = value;
    return SGN_OK;
  }
  if (n == "attn_short_kv") {
    SGN_CHECK_ARG(value == 0 || value == 1, "attn_short_kv must be 0 or 1");
    sgn::g_attn_short_kv = value;
    return SGN_OK;
  }
  if (n == "attn_shape") {
    SGN_CHECK_ARG(value >= 0 && value <= 2, "attn_shape must be 0 (2 x 128 rows, 128-key tiles), 1 (3 x 128 rows, 96-key tiles) or 2 (128 rows, two CTAs per SM)");
    sgn::g_attn_shape = value;
    return SGN_OK;
  }
  if (n == "attn_split") {
    SGN_CHECK_ARG(value == 0 || value == 1, "attn_split must be 0 or 1");
    sgn::g_attn_split = value;
    return SGN_OK;
  }
  if (n == "attn_idle_ns") {
    SGN_CHECK_ARG(value >= 0 && value <= 1000, "attn_idle_ns must be 0..1000");
    sgn::g_attn_idle_ns = value;
    return SGN_OK;
  }
  sgn::set_error("invalid argument: unknown option " + n);
  return SGN_ERR_INVALID_ARG;
}
