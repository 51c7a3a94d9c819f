// K2-K4 — everything between the NeRF render and the diffuser that the reference does with torch boolean
// indexing, a GPU->CPU->GPU cv2.dilate round trip and F.interpolate:
//   K2  AABB slab test + visibility + per-view min/max depth + condition image
//       (signerf/utils/intersection.py:5-56, signerf/datasetgenerator/datasetgenerator.py:758-818)
//   K3  cv2.dilate with the MORPH_ELLIPSE structuring element (datasetgenerator.py:775-778)
//   K4  bilinear resize + paste into / cut out of the reference sheet (datasetgenerator.py:498-539, :570-589,
//       :633-659), masked blend (:562, :656), tensor_to_image quantisation (utils/image_tensor_converter.py:22-30)
// All kernels are HBM-streaming elementwise/stencil passes: coalesced row-major access, grid-stride loops sized
// to the SM count, no host synchronisation.
#include <algorithm>
#include <cmath>
#include <vector>

#include "sgn_device.cuh"

namespace sgn {

static int nsm() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}
static int grid_for(size_t n, int threads, int per_sm = 8) {
  return (int)std::max<size_t>(1, std::min<size_t>((n + threads - 1) / threads, (size_t)nsm() * per_sm));
}

struct ViewStats {  // per view, device
  unsigned int count;
  unsigned int min_bits;  // positive floats order like unsigned ints
  unsigned int max_bits;
  unsigned int pad;
};

__global__ void k_stats_init(ViewStats* s, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) {
    s[i].count = 0;
    s[i].min_bits = 0x7f800000u;  // +inf
    s[i].max_bits = 0u;
    s[i].pad = 0;
  }
}

// intersect_with_aabb + visible mask + per-view reductions.  One block handles a contiguous pixel range of ONE view.
__global__ void __launch_bounds__(256) k_visibility(const float* __restrict__ c2w, const float* __restrict__ intr, int V,
                                                    int H, int W, const float* __restrict__ depth, SgnMaskOpts o,
                                                    uint8_t* __restrict__ vis, ViewStats* __restrict__ stats) {
  const int blocks_per_view = gridDim.x / V;
  const int v = blockIdx.x / blocks_per_view;
  const int b = blockIdx.x - v * blocks_per_view;
  const Camera cam = load_camera(c2w, intr, v);
  const int npix = H * W;
  unsigned int cnt = 0;
  float dmin = __int_as_float(0x7f800000), dmax = 0.f;
  for (int i = b * blockDim.x + threadIdx.x; i < npix; i += blocks_per_view * blockDim.x) {
    int y = i / W, x = i - y * W;
    float d[3];
    ray_direction(cam, (float)x + 0.5f, (float)y + 0.5f, 0.f, 0.f, d);
    float tn = -INFINITY, tf = INFINITY;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float frac = __fdiv_rn(1.f, __fadd_rn(d[k], 1e-6f));
      float ta = __fmul_rn(__fsub_rn(o.aabb[k], cam.o[k]), frac);
      float tb = __fmul_rn(__fsub_rn(o.aabb[3 + k], cam.o[k]), frac);
      tn = fmaxf(tn, fminf(ta, tb));
      tf = fminf(tf, fmaxf(ta, tb));
    }
    const float dep = depth[(size_t)v * npix + i];
    bool visible = (tn < dep) && (dep < tf) && (tn < tf) && (tn > 0.f);
    if (o.inverse_mask) visible = !visible;
    vis[(size_t)v * npix + i] = visible ? 1 : 0;
    if (visible) {
      ++cnt;
      if (dep > 0.f) {
        dmin = fminf(dmin, dep);
        dmax = fmaxf(dmax, dep);
      }
    }
  }
  for (int s = 16; s; s >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, s);
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, s));
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, s));
  }
  if ((threadIdx.x & 31) == 0 && cnt) {
    atomicAdd(&stats[v].count, cnt);
    atomicMin(&stats[v].min_bits, __float_as_uint(dmin));
    atomicMax(&stats[v].max_bits, __float_as_uint(dmax));
  }
}

// Row-wise inclusive->exclusive prefix counts: pre[v][y][x] = #set pixels in columns [0, x).  One warp per row.
__global__ void k_row_prefix(const uint8_t* __restrict__ in, int rows, int W, int* __restrict__ pre) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const uint8_t* src = in + (size_t)row * W;
    int* dst = pre + (size_t)row * (W + 1);
    int carry = 0;
    if (lane == 0) dst[0] = 0;
    for (int x0 = 0; x0 < W; x0 += 32) {
      int x = x0 + lane;
      int v = x < W ? (src[x] != 0) : 0;
      int s = v;
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += t;
      }
      if (x < W) dst[x + 1] = carry + s;
      carry += __shfl_sync(0xffffffffu, s, 31);
    }
  }
}

constexpr int kMaxSeRows = 255;
struct SeRuns {  // structuring element as one horizontal run [j1, j2) per row, anchor at (ax, ay)
  short j1[kMaxSeRows], j2[kMaxSeRows];
  int kh, ax, ay;
};

__global__ void __launch_bounds__(256) k_dilate_runs(const int* __restrict__ pre, int V, int H, int W,
                                                     const __grid_constant__ SeRuns se, uint8_t* __restrict__ out) {
  const size_t n = (size_t)V * H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W);
    size_t r = i / W;
    int y = (int)(r % H);
    size_t vbase = (r - y);  // first row index of this view
    int hit = 0;
    for (int k = 0; k < se.kh && !hit; ++k) {
      int yy = y + k - se.ay;
      if (yy < 0 || yy >= H || se.j2[k] <= se.j1[k]) continue;
      int lo = max(x + se.j1[k] - se.ax, 0), hi = min(x + se.j2[k] - 1 - se.ax, W - 1);
      if (lo > hi) continue;
      const int* p = pre + (vbase + yy) * (size_t)(W + 1);
      hit = (p[hi + 1] - p[lo]) > 0;
    }
    out[i] = hit ? 1 : 0;
  }
}

// shape_depth / shape_color != nullptr: combine_shape_with_depth (datasetgenerator.py:794-807): where the proxy mesh is
// in front of the NeRF surface the condition is the R channel of the mesh's colour image / 255 instead of the normalised
// NeRF depth.
__global__ void __launch_bounds__(256) k_condition(const float* __restrict__ depth, const ViewStats* __restrict__ stats,
                                                   int V, int npix, SgnMaskOpts o, float* __restrict__ cond,
                                                   uint8_t* __restrict__ mask, float* __restrict__ out_stats,
                                                   const float* __restrict__ shape_depth = nullptr,
                                                   const uint8_t* __restrict__ shape_color = nullptr) {
  const size_t n = (size_t)V * npix;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int v = (int)(i / npix);
    const ViewStats s = stats[v];
    float c = 0.f;
    if (s.count > 0) {
      float mn = o.use_manual_depth ? o.manual_min : __fsub_rn(__uint_as_float(s.min_bits), o.depth_radius);
      float mx = o.use_manual_depth ? o.manual_max : __fadd_rn(__uint_as_float(s.max_bits), o.depth_radius);
      float nrm = __fdiv_rn(__fsub_rn(depth[i], mn), __fsub_rn(mx, mn));
      if (shape_depth) {
        const float sd = shape_depth[i];
        if (sd > 0.f && sd < depth[i]) nrm = __fdiv_rn((float)shape_color[3 * i], 255.f);
      }
      c = __fsub_rn(1.f, fminf(fmaxf(nrm, 0.f), 1.f));
    } else {
      mask[i] = 0;  // not visible: zero mask even when inverse_mask dilated nothing
    }
    cond[i] = c;
    if (out_stats && i - (size_t)v * npix == 0) {
      out_stats[4 * v + 0] = s.count > 0 ? 1.f : 0.f;
      out_stats[4 * v + 1] = __uint_as_float(s.min_bits);
      out_stats[4 * v + 2] = __uint_as_float(s.max_bits);
      out_stats[4 * v + 3] = (float)s.count;
    }
  }
}

// cv2.getStructuringElement(MORPH_ELLIPSE, (kw, kh)) as per-row runs (OpenCV morph.dispatch.cpp).
static int make_ellipse(int kw, int kh, SeRuns* se) {
  if (kw < 1 || kh < 1 || kh > kMaxSeRows || kw > 32767) return -1;
  const int r = kh / 2, c = kw / 2;
  const double inv_r2 = r ? 1.0 / ((double)r * r) : 0.0;
  se->kh = kh;
  se->ax = c;
  se->ay = r;
  for (int i = 0; i < kh; ++i) {
    int j1 = 0, j2 = 0;
    const int dy = i - r;
    if (std::abs(dy) <= r) {
      const int dx = (int)std::nearbyint(c * std::sqrt((r * r - dy * dy) * inv_r2));  // cvRound: half to even
      j1 = std::max(c - dx, 0);
      j2 = std::min(c + dx + 1, kw);
    }
    se->j1[i] = (short)j1;
    se->j2[i] = (short)j2;
  }
  return 0;
}

static int dilate_impl(const uint8_t* d_in, int V, int H, int W, int kw, int kh, uint8_t* d_out, cudaStream_t st) {
  SeRuns se;
  if (make_ellipse(kw, kh, &se)) {
    set_error("structuring element too large (kh <= 255)");
    return SGN_ERR_INVALID_ARG;
  }
  int* pre = nullptr;
  SGN_CUDA(scratch_alloc(&pre, (size_t)V * H * (W + 1) * sizeof(int), st));
  k_row_prefix<<<grid_for((size_t)V * H * 32, 256), 256, 0, st>>>(d_in, V * H, W, pre);
  SGN_LAUNCH_CHECK();
  k_dilate_runs<<<grid_for((size_t)V * H * W, 256), 256, 0, st>>>(pre, V, H, W, se, d_out);
  SGN_LAUNCH_CHECK();
  cudaFreeAsync(pre, st);
  return SGN_OK;
}

// ------------------------------------------------------------------------------------------ K4
struct Lerp1D {
  int i0, i1;
  float l0, l1;
};
// ATen area_pixel_compute_source_index(align_corners=False) + linear weights, fp32.
__device__ __forceinline__ Lerp1D lerp_index(int dst, float scale, int in_size) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  Lerp1D r;
  r.i0 = min((int)src, in_size - 1);
  r.i1 = r.i0 + (r.i0 < in_size - 1 ? 1 : 0);
  r.l1 = src - (float)r.i0;
  r.l0 = 1.f - r.l1;
  return r;
}

template <class T>
__global__ void __launch_bounds__(256) k_sheet_paste(const T* __restrict__ src, int V, int H, int W, int C,
                                                     float* __restrict__ sheet, int sheet_w, int cols, int border,
                                                     int th, int tw, int first_cell, float threshold) {
  const float sh = (float)H / (float)th, sw = (float)W / (float)tw;
  const size_t n = (size_t)V * th * tw;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(i % tw);
    size_t r = i / tw;
    int y = (int)(r % th);
    int v = (int)(r / th);
    int cell = first_cell + v;
    int row = cell / cols, col = cell - row * cols;
    Lerp1D ly = lerp_index(y, sh, H), lx = lerp_index(x, sw, W);
    const T* base = src + (size_t)v * H * W * C;
    float* dst = sheet + ((size_t)(row * (th + border) + y) * sheet_w + (col * (tw + border) + x)) * C;
    for (int c = 0; c < C; ++c) {
      float p00 = (float)base[((size_t)ly.i0 * W + lx.i0) * C + c], p01 = (float)base[((size_t)ly.i0 * W + lx.i1) * C + c];
      float p10 = (float)base[((size_t)ly.i1 * W + lx.i0) * C + c], p11 = (float)base[((size_t)ly.i1 * W + lx.i1) * C + c];
      float val = ly.l0 * (lx.l0 * p00 + lx.l1 * p01) + ly.l1 * (lx.l0 * p10 + lx.l1 * p11);
      dst[c] = threshold >= 0.f ? (val > threshold ? 1.f : 0.f) : val;
    }
  }
}

__global__ void __launch_bounds__(256) k_sheet_cut(const float* __restrict__ sheet, int sheet_w, int C, int cols,
                                                   int border, int th, int tw, int cell, float* __restrict__ out, int H,
                                                   int W) {
  const float sh = (float)th / (float)H, sw = (float)tw / (float)W;
  const int row = cell / cols, col = cell - row * cols;
  const float* base = sheet + ((size_t)(row * (th + border)) * sheet_w + col * (tw + border)) * C;
  const size_t n = (size_t)H * W;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    int x = (int)(i % W), y = (int)(i / W);
    Lerp1D ly = lerp_index(y, sh, th), lx = lerp_index(x, sw, tw);
    for (int c = 0; c < C; ++c) {
      float p00 = base[((size_t)ly.i0 * sheet_w + lx.i0) * C + c], p01 = base[((size_t)ly.i0 * sheet_w + lx.i1) * C + c];
      float p10 = base[((size_t)ly.i1 * sheet_w + lx.i0) * C + c], p11 = base[((size_t)ly.i1 * sheet_w + lx.i1) * C + c];
      out[i * C + c] = ly.l0 * (lx.l0 * p00 + lx.l1 * p01) + ly.l1 * (lx.l0 * p10 + lx.l1 * p11);
    }
  }
}

__global__ void k_blend(const float* __restrict__ edited, const float* __restrict__ base, const float* __restrict__ mask,
                        size_t npix, int C, float* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix * C; i += (size_t)gridDim.x * blockDim.x) {
    float m = mask[i / C];
    out[i] = __fadd_rn(__fmul_rn(edited[i], m), __fmul_rn(base[i], __fsub_rn(1.f, m)));
  }
}

__global__ void k_quantize(const float* __restrict__ in, size_t n, uint8_t* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = (uint8_t)(__float2int_rz(__fmul_rn(in[i], 255.f)) & 0xff);
}

// Exchange step of the multi-GPU path fused with the tile packing: rank r holds views r, r+world, ... of every grid g and
// stores the packed pixels [rgb | depth | cond | mask] straight into grid g's owner (rank g) through NVLink-mapped peer
// pointers - an all-to-all without staging buffers or a collective launch.  peer[g] -> [num_views, H, W, 6] fp32.
__global__ void k_scatter_tiles_peer(const float* __restrict__ rgb, const float* __restrict__ depth,
                                     const float* __restrict__ cond, const uint8_t* __restrict__ mask, int G, int v_loc,
                                     size_t hw, int world, int rank, const long long* __restrict__ peer) {
  const size_t n = (size_t)G * v_loc * hw;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i % hw;
    const size_t t = i / hw;            // local tile index = g * v_loc + j
    const int j = (int)(t % v_loc), g = (int)(t / v_loc);
    const int view = rank + j * world;  // global view id inside grid g
    float2* dst = reinterpret_cast<float2*>(reinterpret_cast<float*>(peer[g]) + ((size_t)view * hw + pix) * 6);
    dst[0] = make_float2(rgb[i * 3 + 0], rgb[i * 3 + 1]);
    dst[1] = make_float2(rgb[i * 3 + 2], depth[i]);
    dst[2] = make_float2(cond[i], (float)mask[i]);
  }
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_dilate_ellipse(const uint8_t* d_in, int V, int H, int W, int kw, int kh, uint8_t* d_out,
                                  void* stream) {
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(kw >= 1 && kh >= 1, "kernel size must be >= 1");
  if (V == 0) return SGN_OK;
  SGN_CHECK_ARG(d_in && d_out, "null pointer");
  return dilate_impl(d_in, V, H, W, kw, kh, d_out, reinterpret_cast<cudaStream_t>(stream));
}

static int mask_condition_impl(const float* d_c2w, const float* d_intr, int V, int H, int W, const float* d_depth,
                               const SgnMaskOpts* o, const float* d_shape_depth, const uint8_t* d_shape_color,
                               uint8_t* d_mask, float* d_cond, float* d_stats, void* stream) {
  SGN_CHECK_ARG(o != nullptr, "null opts");
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(V == 0 || (d_c2w && d_intr && d_depth && d_mask && d_cond), "null pointer");
  SGN_CHECK_ARG((o->dilate_w == 0) == (o->dilate_h == 0), "dilate_w/h must both be zero or both positive");
  if (V == 0) return SGN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t npix = (size_t)H * W;
  ViewStats* stats = nullptr;
  uint8_t* vis = nullptr;
  const bool dil = o->dilate_w > 0;
  SGN_CUDA(scratch_alloc(&stats, sizeof(ViewStats) * V, st));
  if (dil) SGN_CUDA(scratch_alloc(&vis, (size_t)V * npix, st));
  k_stats_init<<<(V + 127) / 128, 128, 0, st>>>(stats, V);
  SGN_LAUNCH_CHECK();
  int bpv = std::max(1, std::min((int)((npix + 255) / 256), std::max(1, nsm() * 8 / V)));
  k_visibility<<<bpv * V, 256, 0, st>>>(d_c2w, d_intr, V, H, W, d_depth, *o, dil ? vis : d_mask, stats);
  SGN_LAUNCH_CHECK();
  int rc = SGN_OK;
  if (dil) rc = dilate_impl(vis, V, H, W, o->dilate_w, o->dilate_h, d_mask, st);
  if (rc == SGN_OK) {
    k_condition<<<grid_for((size_t)V * npix, 256), 256, 0, st>>>(d_depth, stats, V, (int)npix, *o, d_cond, d_mask, d_stats,
                                                                 d_shape_depth, d_shape_color);
    count_launch();
    if (cudaPeekAtLastError() != cudaSuccess) {
      set_error(std::string("k_condition launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      rc = SGN_ERR_CUDA;
    }
  }
  if (vis) cudaFreeAsync(vis, st);
  cudaFreeAsync(stats, st);
  return rc;
}

extern "C" int sgn_mask_condition(const float* d_c2w, const float* d_intr, int V, int H, int W, const float* d_depth,
                                  const SgnMaskOpts* o, uint8_t* d_mask, float* d_cond, float* d_stats, void* stream) {
  return mask_condition_impl(d_c2w, d_intr, V, H, W, d_depth, o, nullptr, nullptr, d_mask, d_cond, d_stats, stream);
}

extern "C" int sgn_mask_condition_combined(const float* d_c2w, const float* d_intr, int V, int H, int W,
                                           const float* d_depth, const float* d_shape_depth,
                                           const uint8_t* d_shape_color, const SgnMaskOpts* o, uint8_t* d_mask,
                                           float* d_cond, float* d_stats, void* stream) {
  SGN_CHECK_ARG(V == 0 || (d_shape_depth && d_shape_color), "null shape depth / colour");
  return mask_condition_impl(d_c2w, d_intr, V, H, W, d_depth, o, d_shape_depth, d_shape_color, d_mask, d_cond, d_stats,
                             stream);
}

extern "C" int sgn_sheet_paste(const void* d_src, int src_u8, int V, int H, int W, int C, float* d_sheet, int sheet_h,
                               int sheet_w, int rows, int cols, int border, int tile_h, int tile_w, int first_cell,
                               float threshold, void* stream) {
  SGN_CHECK_ARG(V == 0 || (d_src && d_sheet), "null pointer");
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0 && C > 0 && tile_h > 0 && tile_w > 0, "bad shape");
  SGN_CHECK_ARG(rows > 0 && cols > 0 && border >= 0 && first_cell >= 0 && first_cell + V <= rows * cols,
                "tiles do not fit the grid");
  SGN_CHECK_ARG(rows * tile_h + (rows - 1) * border <= sheet_h && cols * tile_w + (cols - 1) * border <= sheet_w,
                "sheet smaller than the tile grid");
  if (V == 0) return SGN_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  int g = grid_for((size_t)V * tile_h * tile_w, 256);
  if (src_u8)
    k_sheet_paste<uint8_t><<<g, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(d_src), V, H, W, C, d_sheet, sheet_w, cols,
                                              border, tile_h, tile_w, first_cell, threshold);
  else
    k_sheet_paste<float><<<g, 256, 0, st>>>(reinterpret_cast<const float*>(d_src), V, H, W, C, d_sheet, sheet_w, cols,
                                            border, tile_h, tile_w, first_cell, threshold);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_sheet_cut(const float* d_sheet, int sheet_h, int sheet_w, int C, int rows, int cols, int border,
                             int tile_h, int tile_w, int cell, float* d_out, int H, int W, void* stream) {
  SGN_CHECK_ARG(d_sheet && d_out, "null pointer");
  SGN_CHECK_ARG(H > 0 && W > 0 && C > 0 && tile_h > 0 && tile_w > 0, "bad shape");
  SGN_CHECK_ARG(cell >= 0 && cell < rows * cols, "cell out of range");
  SGN_CHECK_ARG(rows * tile_h + (rows - 1) * border <= sheet_h && cols * tile_w + (cols - 1) * border <= sheet_w,
                "sheet smaller than the tile grid");
  k_sheet_cut<<<grid_for((size_t)H * W, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_sheet, sheet_w, C, cols, border, tile_h, tile_w, cell, d_out, H, W);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_blend_masked(const float* d_edited, const float* d_base, const float* d_mask, int64_t npix, int C,
                                float* d_out, void* stream) {
  SGN_CHECK_ARG(npix >= 0 && C > 0, "bad shape");
  if (npix == 0) return SGN_OK;
  SGN_CHECK_ARG(d_edited && d_base && d_mask && d_out, "null pointer");
  k_blend<<<grid_for((size_t)npix * C, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_edited, d_base, d_mask,
                                                                                              (size_t)npix, C, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_quantize_u8(const float* d_in, int64_t n, uint8_t* d_out, void* stream) {
  SGN_CHECK_ARG(n >= 0, "negative n");
  if (n == 0) return SGN_OK;
  SGN_CHECK_ARG(d_in && d_out, "null pointer");
  k_quantize<<<grid_for((size_t)n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_in, (size_t)n, d_out);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_scatter_tiles_peer(const float* d_rgb, const float* d_depth, const float* d_cond, const uint8_t* d_mask,
                                      int G, int v_loc, int H, int W, int world, int rank, const int64_t* d_peer_ptrs,
                                      void* stream) {
  SGN_CHECK_ARG(G >= 0 && v_loc >= 0 && H > 0 && W > 0 && world >= 1 && rank >= 0 && rank < world, "bad shape / rank");
  SGN_CHECK_ARG(G <= world, "one destination rank per grid: G <= world");
  if (G == 0 || v_loc == 0) return SGN_OK;
  SGN_CHECK_ARG(d_rgb && d_depth && d_cond && d_mask && d_peer_ptrs, "null pointer");
  const size_t n = (size_t)G * v_loc * H * W;
  k_scatter_tiles_peer<<<grid_for(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      d_rgb, d_depth, d_cond, d_mask, G, v_loc, (size_t)H * W, world, rank, reinterpret_cast<const long long*>(d_peer_ptrs));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}
