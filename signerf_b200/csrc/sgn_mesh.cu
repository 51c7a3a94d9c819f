// the proxy-mesh depth of reference signerf/renderer/renderer.py:149-196 (pyrender / EGL
// OffscreenRenderer over a trimesh) and the masking_mode == "shape" branch of render_camera
// (signerf/datasetgenerator/datasetgenerator.py:711-757), without OpenGL and without the GPU -> CPU -> GPU round trip.
//
//   sgn_rasterize_depth       z-buffer rasteriser with OpenGL conventions: pyrender's IntrinsicsCamera projection
//                             (znear 1e-4, zfar 10), the Blender -> OpenGL axis swap, pixel centres at +0.5, 8 sub-pixel
//                             bits, top-left fill rule, back-face culling (pyrender's default single-sided material),
//                             LESS depth test on a 24-bit fixed-point buffer, vertical flip, pyrender's
//                             buffer -> metric-depth formula, 0 = empty.
//   sgn_mask_condition_shape  visible = (proxy < nerf) & (proxy > 0), optional inversion, elliptical dilation,
//                             per-view min / max, composed condition image.
// Vertex transforms run in double so that the numpy oracle (oracle/mesh_ref.py) and the kernel snap every vertex to the
// same sub-pixel; coverage and depth are integer / exactly-rounded from there on.
#include <algorithm>

#include "sgn_common.cuh"

namespace sgn {

static inline int grid_m(size_t n, int block, int per_sm = 8) {
  size_t want = (n + block - 1) / block;
  return (int)std::max<size_t>(1, std::min<size_t>(want, (size_t)sm_count() * per_sm));
}
#define STM(s) reinterpret_cast<cudaStream_t>(s)

constexpr unsigned kDepthMax = (1u << 24) - 1u;   // GL_DEPTH_COMPONENT24, cleared to 1.0

struct MeshXf {
  double model[16];   // row-major 4x4: Blender -> OpenGL swap @ [R S | t]
  double znear, zfar;
};

struct VtxOut {
  long long X, Y;     // window coordinates in 1/256 pixel, y up (OpenGL)
  double zw;          // window depth in [0,1]
  int ok;             // w > 0
  int pad;
};

// one thread per (view, vertex)
__global__ void k_mesh_transform(const float* __restrict__ verts, int Nv, const __grid_constant__ MeshXf xf,
                                 const float* __restrict__ c2w, const float* __restrict__ intr, int V, int H, int W,
                                 VtxOut* __restrict__ out) {
  const long long n = (long long)V * Nv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / Nv), k = (int)(i - (long long)v * Nv);
    const double x = verts[3 * k], y = verts[3 * k + 1], z = verts[3 * k + 2];
    // world position in OpenGL axes
    const double* M = xf.model;
    const double wx = M[0] * x + M[1] * y + M[2] * z + M[3];
    const double wy = M[4] * x + M[5] * y + M[6] * z + M[7];
    const double wz = M[8] * x + M[9] * y + M[10] * z + M[11];
    // camera pose in OpenGL axes: convert @ c2w, i.e. rows (r0, r2, -r1); view = rigid inverse
    const float* c = c2w + (size_t)v * 12;
    const double r0[4] = {c[0], c[1], c[2], c[3]};
    const double r1[4] = {c[8], c[9], c[10], c[11]};
    const double r2[4] = {-(double)c[4], -(double)c[5], -(double)c[6], -(double)c[7]};
    const double dx = wx - r0[3], dy = wy - r1[3], dz = wz - r2[3];
    const double ex = r0[0] * dx + r1[0] * dy + r2[0] * dz;   // R^T (p - t)
    const double ey = r0[1] * dx + r1[1] * dy + r2[1] * dz;
    const double ez = r0[2] * dx + r1[2] * dy + r2[2] * dz;
    const double fx = intr[4 * v], fy = intr[4 * v + 1], cx = intr[4 * v + 2], cy = intr[4 * v + 3];
    const double n_ = xf.znear, f_ = xf.zfar;
    // pyrender IntrinsicsCamera.get_projection_matrix
    const double clip_x = (2.0 * fx / W) * ex + (1.0 - 2.0 * cx / W) * ez;
    const double clip_y = (2.0 * fy / H) * ey + (2.0 * cy / H - 1.0) * ez;
    const double clip_z = ((f_ + n_) / (n_ - f_)) * ez + ((2.0 * f_ * n_) / (n_ - f_));
    const double clip_w = -ez;
    VtxOut o;
    o.ok = clip_w > 0.0;
    o.pad = 0;
    if (o.ok) {
      const double xw = (clip_x / clip_w + 1.0) * 0.5 * W, yw = (clip_y / clip_w + 1.0) * 0.5 * H;
      o.X = llrint(xw * 256.0);
      o.Y = llrint(yw * 256.0);
      o.zw = (clip_z / clip_w + 1.0) * 0.5;
    } else {
      o.X = o.Y = 0;
      o.zw = 0.0;
    }
    out[i] = o;
  }
}

__device__ __forceinline__ long long edge_fn(long long ax, long long ay, long long bx, long long by, long long px,
                                             long long py) {
  return (bx - ax) * (py - ay) - (by - ay) * (px - ax);
}
// top-left rule in OpenGL window coordinates (y up) for a counter-clockwise triangle: an edge owns the pixel centres on
// it when it is a left edge (going down) or a top edge (horizontal, going left)
__device__ __forceinline__ bool owns_edge(long long ax, long long ay, long long bx, long long by) {
  const long long dx = bx - ax, dy = by - ay;
  return dy < 0 || (dy == 0 && dx < 0);
}

// one thread per (view, face)
__global__ void k_mesh_raster(const VtxOut* __restrict__ vt, const int* __restrict__ faces, int Nv, int Nf, int V, int H,
                              int W, int cull_back, unsigned* __restrict__ zbuf) {
  const long long n = (long long)V * Nf;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i / Nf), f = (int)(i - (long long)v * Nf);
    VtxOut a = vt[(size_t)v * Nv + faces[3 * f]], b = vt[(size_t)v * Nv + faces[3 * f + 1]],
           c = vt[(size_t)v * Nv + faces[3 * f + 2]];
    if (!(a.ok && b.ok && c.ok)) continue;        // triangles reaching behind the eye are dropped (no near clipping)
    long long area = edge_fn(a.X, a.Y, b.X, b.Y, c.X, c.Y);
    if (area == 0) continue;
    if (area < 0) {
      if (cull_back) continue;                   // clockwise in window space = back face
      VtxOut t = b;
      b = c, c = t, area = -area;
    }
    long long minx = min(a.X, min(b.X, c.X)), maxx = max(a.X, max(b.X, c.X));
    long long miny = min(a.Y, min(b.Y, c.Y)), maxy = max(a.Y, max(b.Y, c.Y));
    // pixel (px, py_up) has its centre at (256 px + 128, 256 py_up + 128)
    int x0 = (int)max(0ll, (minx - 128 + 255) >> 8), x1 = (int)min((long long)W - 1, (maxx - 128) >> 8);
    int y0 = (int)max(0ll, (miny - 128 + 255) >> 8), y1 = (int)min((long long)H - 1, (maxy - 128) >> 8);
    const long long b0 = owns_edge(b.X, b.Y, c.X, c.Y) ? 0 : 1;   // edge opposite a
    const long long b1 = owns_edge(c.X, c.Y, a.X, a.Y) ? 0 : 1;
    const long long b2 = owns_edge(a.X, a.Y, b.X, b.Y) ? 0 : 1;
    const double inv_area = 1.0 / (double)area;
    for (int py = y0; py <= y1; ++py) {
      const long long cy = 256ll * py + 128;
      for (int px = x0; px <= x1; ++px) {
        const long long cxp = 256ll * px + 128;
        const long long e0 = edge_fn(b.X, b.Y, c.X, c.Y, cxp, cy);
        const long long e1 = edge_fn(c.X, c.Y, a.X, a.Y, cxp, cy);
        const long long e2 = edge_fn(a.X, a.Y, b.X, b.Y, cxp, cy);
        if (e0 < b0 || e1 < b1 || e2 < b2) continue;
        const double zw = ((double)e0 * a.zw + (double)e1 * b.zw + (double)e2 * c.zw) * inv_area;
        if (!(zw >= 0.0 && zw <= 1.0)) continue;  // near / far clip per fragment
        const unsigned d = (unsigned)min((double)kDepthMax, floor(zw * (double)kDepthMax + 0.5));
        const int row = H - 1 - py;               // pyrender flips the read-back buffer
        atomicMin(zbuf + ((size_t)v * H + row) * W + px, d);
      }
    }
  }
}

__global__ void k_mesh_clear(unsigned* z, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) z[i] = kDepthMax;
}

// pyrender renderer.py _read_main_framebuffer: depth_im = 2 d - 1; 2 n f / (f + n - depth_im (f - n)); 1.0 -> 0
__global__ void k_mesh_resolve(const unsigned* __restrict__ z, size_t n, float two_nf, float f_plus_n, float f_minus_n,
                               float* __restrict__ depth) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned d = z[i];
    float out = 0.f;
    if (d != kDepthMax) {
      const float df = __fdiv_rn((float)d, (float)kDepthMax);           // fixed point -> GL_FLOAT
      const float ndc = __fsub_rn(__fmul_rn(2.f, df), 1.f);
      out = __fdiv_rn(two_nf, __fsub_rn(f_plus_n, __fmul_rn(ndc, f_minus_n)));
    }
    depth[i] = out;
  }
}

// ------------------------------------------------------------------ masking_mode == "shape"
struct ShapeStats {
  unsigned count, min_bits, max_bits, pad;
};
__global__ void k_shape_stats_init(ShapeStats* s, int V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < V) s[i].count = 0, s[i].min_bits = 0x7f800000u, s[i].max_bits = 0u, s[i].pad = 0;
}
// visible = (proxy < nerf) & (proxy > 0) (^ inverse); min over visible & proxy > 0, max over the WHOLE proxy image
__global__ void __launch_bounds__(256) k_shape_visibility(const float* __restrict__ proxy, const float* __restrict__ nerf,
                                                          int V, int npix, int inverse, uint8_t* __restrict__ vis,
                                                          ShapeStats* __restrict__ stats) {
  const int blocks_per_view = gridDim.x / V;
  const int v = blockIdx.x / blocks_per_view, b = blockIdx.x - v * blocks_per_view;
  unsigned cnt = 0;
  float dmin = __int_as_float(0x7f800000), dmax = 0.f;
  for (int i = b * blockDim.x + threadIdx.x; i < npix; i += blocks_per_view * blockDim.x) {
    const float p = proxy[(size_t)v * npix + i], d = nerf[(size_t)v * npix + i];
    bool m = (p < d) && (p > 0.f);
    if (inverse) m = !m;
    vis[(size_t)v * npix + i] = m ? 1 : 0;
    cnt += m;
    if (m && p > 0.f) dmin = fminf(dmin, p);
    dmax = fmaxf(dmax, p);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    dmax = fmaxf(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
  }
  if ((threadIdx.x & 31) == 0) {
    if (cnt) atomicAdd(&stats[v].count, cnt);
    atomicMin(&stats[v].min_bits, __float_as_uint(dmin));   // non-negative floats order like unsigned ints
    atomicMax(&stats[v].max_bits, __float_as_uint(dmax));
  }
}

__global__ void __launch_bounds__(256) k_shape_condition(const float* __restrict__ proxy, const float* __restrict__ nerf,
                                                         const uint8_t* __restrict__ vis, const ShapeStats* __restrict__ stats,
                                                         int V, int npix, float radius, int use_manual, float man_min,
                                                         float man_max, float* __restrict__ cond, uint8_t* __restrict__ mask,
                                                         float* __restrict__ out_stats) {
  const size_t n = (size_t)V * npix;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int v = (int)(i / npix);
    const ShapeStats s = stats[v];
    float c = 0.f;
    if (s.count > 0) {
      const float mn = use_manual ? man_min : __fsub_rn(__uint_as_float(s.min_bits), radius);
      const float mx = use_manual ? man_max : __fadd_rn(__uint_as_float(s.max_bits), radius);
      const float span = __fsub_rn(mx, mn);
      const float on = __fdiv_rn(__fsub_rn(proxy[i], mn), span), nn = __fdiv_rn(__fsub_rn(nerf[i], mn), span);
      // visible * object + (~visible) * nerf, as the reference writes it (products with 0 / 1, then a sum)
      const float m = vis[i] ? 1.f : 0.f;
      const float mix = __fadd_rn(__fmul_rn(m, on), __fmul_rn(1.f - m, nn));
      c = __fsub_rn(1.f, fminf(fmaxf(mix, 0.f), 1.f));
    } else {
      mask[i] = 0;
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

// pyrender's colour image of a mesh lit by ambient light only (renderer.py:129-130): the flat material colour on
// covered pixels, the scene background elsewhere.
__global__ void __launch_bounds__(256) k_shape_color(const float* __restrict__ depth, size_t npix, uchar3 fg, uchar3 bg,
                                                     uint8_t* __restrict__ out) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
    const uchar3 c = depth[i] > 0.f ? fg : bg;
    out[3 * i + 0] = c.x;
    out[3 * i + 1] = c.y;
    out[3 * i + 2] = c.z;
  }
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_shape_color_u8(const float* d_depth, int64_t npix, const uint8_t* h_fg_rgb, const uint8_t* h_bg_rgb,
                                  uint8_t* d_color, void* stream) {
  SGN_CHECK_ARG(npix >= 0, "negative pixel count");
  if (npix == 0) return SGN_OK;
  SGN_CHECK_ARG(d_depth && h_fg_rgb && h_bg_rgb && d_color, "null pointer");
  k_shape_color<<<grid_m((size_t)npix, 256), 256, 0, STM(stream)>>>(d_depth, (size_t)npix,
                                                                    make_uchar3(h_fg_rgb[0], h_fg_rgb[1], h_fg_rgb[2]),
                                                                    make_uchar3(h_bg_rgb[0], h_bg_rgb[1], h_bg_rgb[2]), d_color);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int64_t sgn_rasterize_ws_bytes(int Nv, int V, int H, int W) {
  if (Nv <= 0 || V <= 0 || H <= 0 || W <= 0) return 0;
  return (int64_t)V * Nv * (int64_t)sizeof(VtxOut) + (int64_t)V * H * W * 4 + 256;
}

extern "C" int sgn_rasterize_depth(const float* d_vertices, const int32_t* d_faces, int Nv, int Nf, const double* h_model,
                                   const float* d_c2w, const float* d_intr, int V, int H, int W, double znear, double zfar,
                                   int cull_back, void* d_ws, float* d_depth, void* stream) {
  SGN_CHECK_ARG(Nv >= 0 && Nf >= 0 && V >= 0 && H > 0 && W > 0 && znear > 0.0 && zfar > znear, "bad rasteriser arguments");
  if (V == 0) return SGN_OK;
  SGN_CHECK_ARG(d_c2w && d_intr && d_depth && d_ws && h_model, "null pointer");
  SGN_CHECK_ARG((Nv == 0 || d_vertices) && (Nf == 0 || d_faces), "null mesh pointer");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_ws) & 15) == 0, "workspace must be 16-byte aligned");
  cudaStream_t st = STM(stream);
  VtxOut* vt = reinterpret_cast<VtxOut*>(d_ws);
  unsigned* zbuf = reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(d_ws) + ((size_t)V * Nv * sizeof(VtxOut) + 255) / 256 * 256);
  const size_t npix = (size_t)V * H * W;
  k_mesh_clear<<<grid_m(npix, 256), 256, 0, st>>>(zbuf, npix);
  SGN_LAUNCH_CHECK();
  if (Nv > 0 && Nf > 0) {
    MeshXf xf;
    for (int i = 0; i < 16; ++i) xf.model[i] = h_model[i];
    xf.znear = znear, xf.zfar = zfar;
    k_mesh_transform<<<grid_m((size_t)V * Nv, 256), 256, 0, st>>>(d_vertices, Nv, xf, d_c2w, d_intr, V, H, W, vt);
    SGN_LAUNCH_CHECK();
    k_mesh_raster<<<grid_m((size_t)V * Nf, 128), 128, 0, st>>>(vt, d_faces, Nv, Nf, V, H, W, cull_back, zbuf);
    SGN_LAUNCH_CHECK();
  }
  // numpy evaluates pyrender's formula in float32 with the Python-float scalars rounded to float32
  k_mesh_resolve<<<grid_m(npix, 256), 256, 0, st>>>(zbuf, npix, (float)(2.0 * znear * zfar), (float)(zfar + znear),
                                                    (float)(zfar - znear), d_depth);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int sgn_mask_condition_shape(const float* d_proxy_depth, const float* d_depth, int V, int H, int W,
                                        const SgnMaskOpts* o, uint8_t* d_mask, float* d_cond, float* d_stats, void* stream) {
  SGN_CHECK_ARG(o != nullptr, "null opts");
  SGN_CHECK_ARG(V >= 0 && H > 0 && W > 0, "bad image shape");
  SGN_CHECK_ARG(V == 0 || (d_proxy_depth && d_depth && d_mask && d_cond), "null pointer");
  SGN_CHECK_ARG((o->dilate_w == 0) == (o->dilate_h == 0), "dilate_w/h must both be zero or both positive");
  if (V == 0) return SGN_OK;
  cudaStream_t st = STM(stream);
  const size_t npix = (size_t)H * W;
  ShapeStats* stats = nullptr;
  uint8_t* vis = nullptr;
  SGN_CUDA(scratch_alloc(&stats, sizeof(ShapeStats) * V, st));
  SGN_CUDA(scratch_alloc(&vis, (size_t)V * npix, st));
  k_shape_stats_init<<<(V + 127) / 128, 128, 0, st>>>(stats, V);
  SGN_LAUNCH_CHECK();
  const int bpv = std::max(1, std::min((int)((npix + 255) / 256), std::max(1, sm_count() * 8 / V)));
  k_shape_visibility<<<bpv * V, 256, 0, st>>>(d_proxy_depth, d_depth, V, (int)npix, o->inverse_mask, vis, stats);
  SGN_LAUNCH_CHECK();
  int rc = SGN_OK;
  if (o->dilate_w > 0) rc = sgn_dilate_ellipse(vis, V, H, W, o->dilate_w, o->dilate_h, d_mask, stream);
  else if (cudaMemcpyAsync(d_mask, vis, (size_t)V * npix, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = SGN_ERR_CUDA;
  if (rc == SGN_OK) {
    k_shape_condition<<<grid_m((size_t)V * npix, 256), 256, 0, st>>>(d_proxy_depth, d_depth, vis, stats, V, (int)npix,
                                                                     o->depth_radius, o->use_manual_depth, o->manual_min,
                                                                     o->manual_max, d_cond, d_mask, d_stats);
    count_launch();
    if (cudaPeekAtLastError() != cudaSuccess) {
      set_error(std::string("k_shape_condition launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      rc = SGN_ERR_CUDA;
    }
  }
  cudaFreeAsync(vis, st);
  cudaFreeAsync(stats, st);
  return rc;
}
