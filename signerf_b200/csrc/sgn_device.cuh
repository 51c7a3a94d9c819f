// Device-side building blocks shared by the render / probe kernels.
// Geometry follows nerfstudio 1.0.x op order in fp32 with explicit round-to-nearest intrinsics so
// that nvcc never contracts mul+add into an FMA where torch rounds twice (hash indices must match
// the oracle bit for bit when fed the same positions).
#pragma once
#include "sgn_common.cuh"

namespace sgn {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier + bulk copy (TMA unit)
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion counted on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- ray generation (A2)
struct Camera {
  float r[3][3];
  float o[3];
  float fx, fy, cx, cy;
};

__device__ __forceinline__ Camera load_camera(const float* __restrict__ c2w, const float* __restrict__ intr, int v) {
  Camera c;
  const float* m = c2w + 12 * v;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c.r[i][0] = __ldg(m + 4 * i + 0);
    c.r[i][1] = __ldg(m + 4 * i + 1);
    c.r[i][2] = __ldg(m + 4 * i + 2);
    c.o[i] = __ldg(m + 4 * i + 3);
  }
  c.fx = __ldg(intr + 4 * v + 0);
  c.fy = __ldg(intr + 4 * v + 1);
  c.cx = __ldg(intr + 4 * v + 2);
  c.cy = __ldg(intr + 4 * v + 3);
  return c;
}

// Cameras._generate_rays_from_coords for a perspective pinhole: direction through the point
// (x + ox, y + oy) in pixel units (pixel centre = integer + 0.5).  Returns the pre-normalisation norm.
__device__ __forceinline__ float ray_direction(const Camera& c, float x, float y, float addx, float addy, float d[3]) {
  float u = __fdiv_rn(addx != 0.f ? __fadd_rn(__fsub_rn(x, c.cx), addx) : __fsub_rn(x, c.cx), c.fx);
  float v = __fdiv_rn(-(addy != 0.f ? __fadd_rn(__fsub_rn(y, c.cy), addy) : __fsub_rn(y, c.cy)), c.fy);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    d[i] = __fadd_rn(__fadd_rn(__fmul_rn(u, c.r[i][0]), __fmul_rn(v, c.r[i][1])), __fmul_rn(-1.f, c.r[i][2]));
  float n = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d[0], d[0]), __fmul_rn(d[1], d[1])), __fmul_rn(d[2], d[2])));
  n = fmaxf(n, 1e-6f);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = __fdiv_rn(d[i], n);
  return n;
}

// Where a kernel's rays come from: pinhole cameras (c2w / intr, the pixel grid of V views, warp tile 8x4 pixels) or an
// explicit bundle of N rays (origins / directions [N,3]: nerfstudio's RayBundle as `get_outputs_for_camera_ray_bundle`
// receives it, datasetgenerator.py:691-694; addressed as ONE H=1, W=N image with 32x1 warp tiles).
struct RaySource {
  const float* c2w;
  const float* intr;
  const float* rays_o;
  const float* rays_d;
  int tw_log2;  // log2 of the warp tile's width: 3 (8x4 pixels) for cameras, 5 (32 consecutive rays) for bundles
};

// pixel (x, y) of lane-row `row` of tile (tx, ty)
__device__ __forceinline__ void tile_xy(const RaySource& rs, int tx, int ty, int row, int& x, int& y) {
  x = (tx << rs.tw_log2) + (row & ((1 << rs.tw_log2) - 1));
  y = ty * (32 >> rs.tw_log2) + (row >> rs.tw_log2);
}

// origin + unit direction of the ray through pixel (x, y) of view v (cameras) or ray x of the bundle
__device__ __forceinline__ void load_ray(const RaySource& rs, int v, int x, int y, float o[3], float d[3]) {
  if (rs.rays_o) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      o[k] = __ldg(rs.rays_o + 3 * (size_t)x + k);
      d[k] = __ldg(rs.rays_d + 3 * (size_t)x + k);
    }
  } else {
    const Camera cam = load_camera(rs.c2w, rs.intr, v);
    o[0] = cam.o[0]; o[1] = cam.o[1]; o[2] = cam.o[2];
    ray_direction(cam, (float)x + 0.5f, (float)y + 0.5f, 0.f, 0.f, d);
  }
}

// ---------------------------------------------------------------- samplers (A3)
__device__ __forceinline__ float spacing_fn(float x) {  // UniformLinDispPiecewiseSampler
  return x < 1.f ? __fmul_rn(x, 0.5f) : __fsub_rn(1.f, __fdiv_rn(1.f, __fmul_rn(2.f, x)));
}
__device__ __forceinline__ float spacing_inv(float x) {
  return x < 0.5f ? __fmul_rn(2.f, x) : __fdiv_rn(1.f, __fsub_rn(2.f, __fmul_rn(2.f, x)));
}
__device__ __forceinline__ float to_euclid(float u, float s_near, float s_far) {
  return spacing_inv(__fadd_rn(__fmul_rn(u, s_far), __fmul_rn(__fsub_rn(1.f, u), s_near)));
}

// ---------------------------------------------------------------- contraction + selector (A4)
// SceneContraction(L-inf) -> (x+2)/4 -> selector -> positions * selector.
__device__ __forceinline__ bool contract_to_unit(float x, float y, float z, float& px, float& py, float& pz) {
  float mag = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  if (!(mag < 1.f)) {
    float s = __fsub_rn(2.f, __fdiv_rn(1.f, mag));
    x = __fmul_rn(s, __fdiv_rn(x, mag));
    y = __fmul_rn(s, __fdiv_rn(y, mag));
    z = __fmul_rn(s, __fdiv_rn(z, mag));
  }
  px = __fmul_rn(__fadd_rn(x, 2.f), 0.25f);
  py = __fmul_rn(__fadd_rn(y, 2.f), 0.25f);
  pz = __fmul_rn(__fadd_rn(z, 2.f), 0.25f);
  bool sel = (px > 0.f) & (px < 1.f) & (py > 0.f) & (py < 1.f) & (pz > 0.f) & (pz < 1.f);
  if (!sel) px = py = pz = 0.f;
  return sel;
}

// ---------------------------------------------------------------- hash grid (A4)
struct LevelCoords {
  uint32_t xf, xc, yf, yc, zf, zc;  // already multiplied by the primes
  float ox, oy, oz;
};

__device__ __forceinline__ LevelCoords level_coords(float res, float px, float py, float pz) {
  LevelCoords L;
  float qx = __fmul_rn(px, res), qy = __fmul_rn(py, res), qz = __fmul_rn(pz, res);
  float fx = floorf(qx), fy = floorf(qy), fz = floorf(qz);
  L.ox = __fsub_rn(qx, fx);
  L.oy = __fsub_rn(qy, fy);
  L.oz = __fsub_rn(qz, fz);
  // ceil(q) = floor(q) + (q > floor(q)) for the non-negative coordinates of the unit cube: an integer add instead of a second
  // round + convert pair per axis (both run on the quarter-rate XU pipe: 96 of K1's 192 conversions per sample)
  const uint32_t ix = (uint32_t)(int)fx, iy = (uint32_t)(int)fy, iz = (uint32_t)(int)fz;
  L.xf = ix;
  L.xc = ix + (L.ox > 0.f ? 1u : 0u);
  L.yf = iy * kPrimeY;
  L.yc = (iy + (L.oy > 0.f ? 1u : 0u)) * kPrimeY;
  L.zf = iz * kPrimeZ;
  L.zc = (iz + (L.oz > 0.f ? 1u : 0u)) * kPrimeZ;
  return L;
}

// Corner order of HashEncoding.pytorch_fwd: 0 ccc, 1 cfc, 2 ffc, 3 fcc, 4 ccf, 5 cff, 6 fff, 7 fcf.
__device__ __forceinline__ void corner_rows(const LevelCoords& L, uint32_t mask, uint32_t idx[8]) {
  idx[0] = (L.xc ^ L.yc ^ L.zc) & mask;
  idx[1] = (L.xc ^ L.yf ^ L.zc) & mask;
  idx[2] = (L.xf ^ L.yf ^ L.zc) & mask;
  idx[3] = (L.xf ^ L.yc ^ L.zc) & mask;
  idx[4] = (L.xc ^ L.yc ^ L.zf) & mask;
  idx[5] = (L.xc ^ L.yf ^ L.zf) & mask;
  idx[6] = (L.xf ^ L.yf ^ L.zf) & mask;
  idx[7] = (L.xf ^ L.yc ^ L.zf) & mask;
}

// Hash-table gradient of one level, scatter-add with the lanes of a warp merged first.  Consecutive lanes hold
// consecutive samples of a ray; at the coarse levels (and wherever the PDF sampler concentrates samples) whole runs of them
// sit in the same cell, i.e. hit the same eight rows.  Runs of equal cells are summed with a segmented shuffle reduction and
// only the first lane of a run issues the eight vector atomics (a coarse level otherwise takes millions of atomics on a
// few thousand addresses).  Must be called by all 32 lanes; lanes without work pass g0 = g1 = 0.
__device__ __forceinline__ void scatter_level_merged(float2* __restrict__ gt, const LevelCoords& L, uint32_t mask, float g0,
                                                     float g1, int lane) {
  const unsigned full = 0xffffffffu;
  bool same = true;   // same cell as the previous lane: the six corner coordinates agree (the primes are odd: bijective)
  same &= __shfl_up_sync(full, L.xf, 1) == L.xf;
  same &= __shfl_up_sync(full, L.yf, 1) == L.yf;
  same &= __shfl_up_sync(full, L.zf, 1) == L.zf;
  same &= __shfl_up_sync(full, L.xc, 1) == L.xc;
  same &= __shfl_up_sync(full, L.yc, 1) == L.yc;
  same &= __shfl_up_sync(full, L.zc, 1) == L.zc;
  const bool head = lane == 0 || !same;
  const unsigned heads = __ballot_sync(full, head);
  const unsigned after = lane == 31 ? 0u : heads & ~((2u << lane) - 1u);
  const int end = after ? __ffs(after) - 1 : 32;   // first lane of the next run
  const float mx = 1.f - L.ox, my = 1.f - L.oy, mz = 1.f - L.oz;
  // corner order of HashEncoding.pytorch_fwd: 0 ccc, 1 cfc, 2 ffc, 3 fcc, 4 ccf, 5 cff, 6 fff, 7 fcf (x, y, z)
  const float wt[8] = {L.ox * L.oy * L.oz, L.ox * my * L.oz, mx * my * L.oz, mx * L.oy * L.oz,
                       L.ox * L.oy * mz,   L.ox * my * mz,   mx * my * mz,   mx * L.oy * mz};
  float vx[8], vy[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) vx[c] = wt[c] * g0, vy[c] = wt[c] * g1;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const bool take = lane + o < end;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float ux = __shfl_down_sync(full, vx[c], o), uy = __shfl_down_sync(full, vy[c], o);
      if (take) vx[c] += ux, vy[c] += uy;
    }
  }
  if (head) {
    uint32_t idx[8];
    corner_rows(L, mask, idx);
#pragma unroll
    for (int c = 0; c < 8; ++c)
      if (vx[c] != 0.f || vy[c] != 0.f) atomicAdd(gt + idx[c], make_float2(vx[c], vy[c]));
  }
}

__device__ __forceinline__ float2 lerp2(float2 a, float wa, float2 b, float wb) {
  return make_float2(fmaf(a.x, wa, b.x * wb), fmaf(a.y, wa, b.y * wb));
}

// Trilinear features of one level. `tab` already points at the level's first row.
__device__ __forceinline__ float2 encode_level(const float2* __restrict__ tab, uint32_t mask, float res, float px,
                                               float py, float pz) {
  LevelCoords L = level_coords(res, px, py, pz);
  uint32_t idx[8];
  corner_rows(L, mask, idx);
  float2 f[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) f[c] = __ldg(tab + idx[c]);
  float mx = 1.f - L.ox, my = 1.f - L.oy, mz = 1.f - L.oz;
  float2 f03 = lerp2(f[0], L.ox, f[3], mx);
  float2 f12 = lerp2(f[1], L.ox, f[2], mx);
  float2 f56 = lerp2(f[5], L.ox, f[6], mx);
  float2 f47 = lerp2(f[4], L.ox, f[7], mx);
  float2 f0312 = lerp2(f03, L.oy, f12, my);
  float2 f4756 = lerp2(f47, L.oy, f56, my);
  return lerp2(f0312, L.oz, f4756, mz);
}

// ---------------------------------------------------------------- SH degree 4 (A5)
// components_from_spherical_harmonics(levels=4) evaluated on (d+1)/2 exactly as the torch
// fallback of SHEncoding is fed by NerfactoField.get_outputs.
__device__ __forceinline__ void sh16(float dx, float dy, float dz, float out[16]) {
  float x = (dx + 1.f) * 0.5f, y = (dy + 1.f) * 0.5f, z = (dz + 1.f) * 0.5f;
  float xx = x * x, yy = y * y, zz = z * z;
  out[0] = 0.28209479177387814f;
  out[1] = 0.4886025119029199f * y;
  out[2] = 0.4886025119029199f * z;
  out[3] = 0.4886025119029199f * x;
  out[4] = 1.0925484305920792f * x * y;
  out[5] = 1.0925484305920792f * y * z;
  out[6] = 0.9461746957575601f * zz - 0.31539156525251999f;
  out[7] = 1.0925484305920792f * x * z;
  out[8] = 0.5462742152960396f * (xx - yy);
  out[9] = 0.5900435899266435f * y * (3.f * xx - yy);
  out[10] = 2.890611442640554f * x * y * z;
  out[11] = 0.4570457994644658f * y * (5.f * zz - 1.f);
  out[12] = 0.3731763325901154f * z * (5.f * zz - 3.f);
  out[13] = 0.4570457994644658f * x * (5.f * zz - 1.f);
  out[14] = 1.445305721320277f * z * (xx - yy);
  out[15] = 0.5900435899266435f * x * (xx - 3.f * yy);
}

// ---------------------------------------------------------------- tensor-core MLP (A5)
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) { return pack_h2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3])
               : "r"(addr));
}

constexpr int kStageStride = 40;  // halfs per staging row: 80 B keeps STS.128 / ldmatrix conflict-free

// 64-wide hidden layer for 16 samples: A (K = 16*KT) -> ReLU(acc*scale + bias) as next-layer A fragments.
template <int KT>
__device__ __forceinline__ void dense64_relu(const uint4* __restrict__ w, const float* __restrict__ bias, float scale,
                                             int lane, const uint32_t (&a)[KT][4], uint32_t (&out)[4][4]) {
  const int t = lane & 3;
#pragma unroll
  for (int jp = 0; jp < 4; ++jp) {
    float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      uint4 b = w[(kk * 4 + jp) * 32 + lane];
      mma16816(c0, a[kk], b.x, b.y);
      mma16816(c1, a[kk], b.z, b.w);
    }
    float2 b0 = *reinterpret_cast<const float2*>(bias + 16 * jp + 2 * t);
    float2 b1 = *reinterpret_cast<const float2*>(bias + 16 * jp + 8 + 2 * t);
    out[jp][0] = pack_relu_h2(fmaf(c0[0], scale, b0.x), fmaf(c0[1], scale, b0.y));
    out[jp][1] = pack_relu_h2(fmaf(c0[2], scale, b0.x), fmaf(c0[3], scale, b0.y));
    out[jp][2] = pack_relu_h2(fmaf(c1[0], scale, b1.x), fmaf(c1[1], scale, b1.y));
    out[jp][3] = pack_relu_h2(fmaf(c1[2], scale, b1.x), fmaf(c1[3], scale, b1.y));
  }
}

// Field MLPs for one m-tile of 16 samples held as mma A-fragments.
//   af[2]  : hash features (k 0..31, pre-scaled by feat_scale)
//   ash    : SH(16) fragment of the same 16 rows
// Results in C-fragment layout (g = lane/4, t = lane%4):
//   lg0 / lg1 : density logit of rows g / g+8          (meaningful on t == 0)
//   rgb[4]    : pre-sigmoid colour, rows g (c0,c1) and g+8 (c2,c3), columns 2t, 2t+1
__device__ __forceinline__ void field_mlp_mtile(const MlpPack* __restrict__ sp, int lane, const uint32_t (&af)[2][4],
                                                const uint32_t (&ash)[4], float& lg0, float& lg1, float (&rgb)[4]) {
  const int t = lane & 3;
  uint32_t hid[4][4];
  dense64_relu<2>(sp->w_base0, sp->b_base0, sp->inv_feat_scale, lane, af, hid);
  // base layer 1: 64 -> 16, no activation
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    uint4 b = sp->w_base1[kk * 32 + lane];
    mma16816(c0, hid[kk], b.x, b.y);
    mma16816(c1, hid[kk], b.z, b.w);
  }
  {
    float2 b0 = *reinterpret_cast<const float2*>(sp->b_base1 + 2 * t);
    float2 b1 = *reinterpret_cast<const float2*>(sp->b_base1 + 8 + 2 * t);
    c0[0] += b0.x; c0[1] += b0.y; c0[2] += b0.x; c0[3] += b0.y;
    c1[0] += b1.x; c1[1] += b1.y; c1[2] += b1.x; c1[3] += b1.y;
  }
  lg0 = c0[0];
  lg1 = c0[2];
  // head input = [SH16 | (slot0 = 0) geo15]; column 0 of the base output is the density logit
  uint32_t hin[2][4];
  hin[0][0] = ash[0]; hin[0][1] = ash[1]; hin[0][2] = ash[2]; hin[0][3] = ash[3];
  hin[1][0] = pack_h2(t == 0 ? 0.f : c0[0], c0[1]);
  hin[1][1] = pack_h2(t == 0 ? 0.f : c0[2], c0[3]);
  hin[1][2] = pack_h2(c1[0], c1[1]);
  hin[1][3] = pack_h2(c1[2], c1[3]);
  uint32_t h1[4][4], h2[4][4];
  dense64_relu<2>(sp->w_head0, sp->b_head0, 1.f, lane, hin, h1);
  dense64_relu<4>(sp->w_head1, sp->b_head1, 1.f, lane, h1, h2);
  rgb[0] = rgb[1] = rgb[2] = rgb[3] = 0.f;
#pragma unroll
  for (int kp = 0; kp < 2; ++kp) {
    uint4 b = sp->w_head2[kp * 32 + lane];
    mma16816(rgb, h2[2 * kp], b.x, b.y);
    mma16816(rgb, h2[2 * kp + 1], b.z, b.w);
  }
  float2 bo = *reinterpret_cast<const float2*>(sp->b_head2 + 2 * t);
  rgb[0] += bo.x; rgb[1] += bo.y; rgb[2] += bo.x; rgb[3] += bo.y;
}

// Evaluate the field for the 32 samples a warp holds (one per lane):
//   lane L supplies contracted position p (already masked by the selector) and has staged its SH row;
//   on return lane q = (g, t) owns sample row rc = g + 8t and receives (logit, r, g, b) pre-activation.
// `stage` is the warp's [32][kStageStride] fp16 staging tile.
__device__ __forceinline__ void warp_field_eval(const GridDev& grid, const MlpPack* __restrict__ sp,
                                                __half* __restrict__ stage, int lane, float px, float py, float pz,
                                                const uint32_t (&ash)[2][4], float& logit, float& cr, float& cg,
                                                float& cb) {
  const float fs = sp->feat_scale;
  // the 4-level group is unrolled, the loop over groups is not: fully unrolled, the 16 levels + MLP exceed the 32 KB
  // instruction cache (ncu: 15 % of warp samples stalled on no_instruction)
#pragma unroll 1
  for (int l4 = 0; l4 < 4; ++l4) {
    uint32_t h[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int l = 4 * l4 + q;
      float2 f = encode_level(grid.table + (size_t)l * grid.size, grid.mask, grid.res[l], px, py, pz);
      h[q] = pack_h2(f.x * fs, f.y * fs);
    }
    *reinterpret_cast<uint4*>(stage + lane * kStageStride + 8 * l4) = make_uint4(h[0], h[1], h[2], h[3]);
  }
  __syncwarp();
  const int t = lane & 3;
  const int base = lane & ~3;
  const uint32_t ld_row = (lane & 7) + 8 * ((lane >> 3) & 1);
  const uint32_t ld_col = 8 * (lane >> 4);
  // one copy of the MLP code for both 16-sample m-tiles (instruction-cache footprint, see above)
#pragma unroll 1
  for (int m = 0; m < 2; ++m) {
    uint32_t af[2][4];
    uint32_t addr = smem_u32(stage + (16 * m + ld_row) * kStageStride + ld_col);
    ldmatrix_x4(af[0], addr);
    ldmatrix_x4(af[1], addr + 32);
    float lg0, lg1, c[4];
    uint32_t am[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) am[i] = m ? ash[1][i] : ash[0][i];
    field_mlp_mtile(sp, lane, af, am, lg0, lg1, c);
    // route rows {g, g+8} of this m-tile to lanes t = 2m, 2m+1 of the quad
    float lgA = __shfl_sync(0xffffffffu, lg0, base);
    float lgB = __shfl_sync(0xffffffffu, lg1, base);
    float rA = __shfl_sync(0xffffffffu, c[0], base);
    float gA = __shfl_sync(0xffffffffu, c[1], base);
    float rB = __shfl_sync(0xffffffffu, c[2], base);
    float gB = __shfl_sync(0xffffffffu, c[3], base);
    float bA = __shfl_sync(0xffffffffu, c[0], base + 1);
    float bB = __shfl_sync(0xffffffffu, c[2], base + 1);
    if (t == 2 * m) { logit = lgA; cr = rA; cg = gA; cb = bA; }
    if (t == 2 * m + 1) { logit = lgB; cr = rB; cg = gB; cb = bB; }
  }
  __syncwarp();
}

// Stage the per-lane SH rows and read them back as A-fragments for both m-tiles.
__device__ __forceinline__ void warp_stage_sh(__half* __restrict__ stage, int lane, const float sh[16],
                                              uint32_t (&ash)[2][4]) {
  uint4 lo = make_uint4(pack_h2(sh[0], sh[1]), pack_h2(sh[2], sh[3]), pack_h2(sh[4], sh[5]), pack_h2(sh[6], sh[7]));
  uint4 hi = make_uint4(pack_h2(sh[8], sh[9]), pack_h2(sh[10], sh[11]), pack_h2(sh[12], sh[13]), pack_h2(sh[14], sh[15]));
  *reinterpret_cast<uint4*>(stage + lane * kStageStride) = lo;
  *reinterpret_cast<uint4*>(stage + lane * kStageStride + 8) = hi;
  __syncwarp();
  const uint32_t ld_row = (lane & 7) + 8 * ((lane >> 3) & 1);
  const uint32_t ld_col = 8 * (lane >> 4);
  ldmatrix_x4(ash[0], smem_u32(stage + ld_row * kStageStride + ld_col));
  ldmatrix_x4(ash[1], smem_u32(stage + (16 + ld_row) * kStageStride + ld_col));
  __syncwarp();
}

// ---------------------------------------------------------------- fp32 CUDA-core MLP (parity path)
// head_bias: 64 per-ray values replacing b_head0 (training with per-image appearance embeddings), or null.
static __device__ __noinline__ void field_mlp_f32(const MlpF32* __restrict__ w, const float feat[32], const float sh[16],
                                           float& logit, float rgb[3], const float* __restrict__ head_bias = nullptr) {
  float h[64];
  for (int n = 0; n < 64; ++n) {
    float a = w->b_base0[n];
#pragma unroll
    for (int k = 0; k < 32; ++k) a = fmaf(w->w_base0[n * 32 + k], feat[k], a);
    h[n] = fmaxf(a, 0.f);
  }
  float hin[32];
#pragma unroll
  for (int k = 0; k < 16; ++k) hin[k] = sh[k];
  for (int n = 0; n < 16; ++n) {
    float a = w->b_base1[n];
    for (int k = 0; k < 64; ++k) a = fmaf(w->w_base1[n * 64 + k], h[k], a);
    hin[16 + n] = a;
  }
  logit = hin[16];
  hin[16] = 0.f;
  for (int n = 0; n < 64; ++n) {
    float a = head_bias ? __ldg(head_bias + n) : w->b_head0[n];
    for (int k = 0; k < 32; ++k) a = fmaf(w->w_head0[n * 32 + k], hin[k], a);
    h[n] = fmaxf(a, 0.f);
  }
  float h2[64];
  for (int n = 0; n < 64; ++n) {
    float a = w->b_head1[n];
    for (int k = 0; k < 64; ++k) a = fmaf(w->w_head1[n * 64 + k], h[k], a);
    h2[n] = fmaxf(a, 0.f);
  }
  for (int n = 0; n < 3; ++n) {
    float a = w->b_head2[n];
    for (int k = 0; k < 64; ++k) a = fmaf(w->w_head2[n * 64 + k], h2[k], a);
    rgb[n] = a;
  }
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

// ---------------------------------------------------------------- compositing (A6)
struct Composite {
  float cum_dd = 0.f, r = 0.f, g = 0.f, b = 0.f, acc = 0.f, cumw = 0.f, depth = 0.f;
  float lr = 0.f, lg = 0.f, lb = 0.f;
  float margin = 0.f;  // distance of 0.5 from the cumulative weights either side of the median pick (see finish)
  bool found = false;
  __device__ __forceinline__ void step(float sigma, float delta, float tmid, float cr, float cg, float cb) {
    float dd = delta * sigma;
    float alpha = 1.f - expf(-dd);
    float T = expf(-cum_dd);
    cum_dd += dd;
    float w = alpha * T;
    if (w != w) w = 0.f;  // nan_to_num
    r = fmaf(w, cr, r); g = fmaf(w, cg, g); b = fmaf(w, cb, b);
    acc += w;
    cumw += w;
    if (!found && cumw >= 0.5f) {
      found = true;
      depth = tmid;
      margin = fminf(cumw - 0.5f, 0.5f - (cumw - w));   // how far the pick is from moving to a neighbouring bin
    }
    lr = cr; lg = cg; lb = cb;
  }
  // RGBRenderer("last_sample") + eval clamp; DepthRenderer("median") with the index clamp to S-1.
  __device__ __forceinline__ void finish(float last_tmid, float out_rgb[3], float& out_depth) {
    float rem = 1.f - acc;
    out_rgb[0] = fminf(fmaxf(fmaf(lr, rem, r), 0.f), 1.f);
    out_rgb[1] = fminf(fmaxf(fmaf(lg, rem, g), 0.f), 1.f);
    out_rgb[2] = fminf(fmaxf(fmaf(lb, rem, b), 0.f), 1.f);
    out_depth = found ? depth : last_tmid;
    if (!found) margin = 0.5f - cumw;   // never reached one half: the pick is the last bin unless the total weight grows
  }
};

// ---------------------------------------------------------------- weight gradients of the training kernels
// dW[n][k] += sum_s D[doff + n][s] * A[aoff + k][s], db[n] += sum_s D[doff + n][s] over [component][sample] slabs
// (k_outer_reduce, sgn_train.cu): up to five layers per launch.
struct OuterLayer {
  int doff, N, aoff, K, ldw;
  float* dW;
  float* db;
};
struct OuterParams {
  OuterLayer layer[5];
};
void launch_outer_reduce(const float* D, const float* A, const OuterParams& op, int layers, int64_t count, int64_t cap,
                         cudaStream_t st);

}  // namespace sgn
