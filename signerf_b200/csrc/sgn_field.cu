// Field handle: takes nerfacto parameters (hash tables by reference, MLPs by value) and prepares
// the device-side blocks the render kernels consume (fp16 mma fragments, fp32 parity weights,
// folded appearance bias).  Replaces nothing in the reference by itself; it is the parameter
// hand-off for `graph` in DatasetGenerator.render_camera (datasetgenerator.py:691-694).
#include <cmath>
#include <cstring>
#include <vector>

#include "sgn_common.cuh"

namespace sgn {

static thread_local std::string t_error;
std::atomic<uint64_t> g_launches{0};
void set_error(const std::string& msg) { t_error = msg; }

static uint32_t pack_half2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  uint32_t u;
  memcpy(&u, &h, 4);
  return u;
}

// W: accessor (n, k) -> float.  N multiple of 16, K multiple of 16.
template <class F>
static void pack_b_pairs(uint4* out, int N, int K, F W) {
  for (int kk = 0; kk < K / 16; ++kk)
    for (int jp = 0; jp < N / 16; ++jp)
      for (int lane = 0; lane < 32; ++lane) {
        int g = lane >> 2, t = lane & 3;
        int n0 = 16 * jp + g, n1 = n0 + 8, k0 = 16 * kk + 2 * t;
        uint4 v;
        v.x = pack_half2(W(n0, k0), W(n0, k0 + 1));
        v.y = pack_half2(W(n0, k0 + 8), W(n0, k0 + 9));
        v.z = pack_half2(W(n1, k0), W(n1, k0 + 1));
        v.w = pack_half2(W(n1, k0 + 8), W(n1, k0 + 9));
        out[(kk * (N / 16) + jp) * 32 + lane] = v;
      }
}

// N = 8 (one n-tile), packed by k-tile pairs.
template <class F>
static void pack_b_n8(uint4* out, int K, F W) {
  for (int kp = 0; kp < K / 32; ++kp)
    for (int lane = 0; lane < 32; ++lane) {
      int g = lane >> 2, t = lane & 3;
      int k0 = 32 * kp + 2 * t;
      uint4 v;
      v.x = pack_half2(W(g, k0), W(g, k0 + 1));
      v.y = pack_half2(W(g, k0 + 8), W(g, k0 + 9));
      v.z = pack_half2(W(g, k0 + 16), W(g, k0 + 17));
      v.w = pack_half2(W(g, k0 + 24), W(g, k0 + 25));
      out[kp * 32 + lane] = v;
    }
}

__global__ void k_max_abs(const float* __restrict__ x, size_t n, unsigned int* out) {
  float m = 0.f;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float v = fabsf(x[i]);
    if (v == v) m = fmaxf(m, v);
  }
  for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));  // non-negative floats order as uints
}

static int fill_grid(const SgnHashGrid& g, GridDev* out, const char* name) {
  SGN_CHECK_ARG(g.d_table != nullptr && g.h_scalings != nullptr, std::string(name) + ": null table/scalings");
  SGN_CHECK_ARG(g.num_levels >= 1 && g.num_levels <= kMaxLevels, std::string(name) + ": num_levels must be 1..16");
  SGN_CHECK_ARG(g.log2_size >= 1 && g.log2_size <= 24, std::string(name) + ": log2_size must be 1..24");
  out->table = reinterpret_cast<const float2*>(g.d_table);
  out->num_levels = g.num_levels;
  out->size = 1u << g.log2_size;
  out->mask = out->size - 1;
  for (int i = 0; i < kMaxLevels; ++i) out->res[i] = i < g.num_levels ? g.h_scalings[i] : 0.f;
  return SGN_OK;
}

// fp16 tensor-core fragments of the fp32 parameter block (kernel order: w_head0 columns [SH16 | slot | geo15], folded bias)
static void pack_from_f32(const MlpF32& Q, float feat_scale, MlpPack& P) {
  memset(&P, 0, sizeof(P));
  pack_b_pairs(P.w_base0, 64, 32, [&](int n, int k) { return Q.w_base0[n * 32 + k]; });
  pack_b_pairs(P.w_base1, 16, 64, [&](int n, int k) { return Q.w_base1[n * 64 + k]; });
  pack_b_pairs(P.w_head0, 64, 32, [&](int n, int k) { return Q.w_head0[n * 32 + k]; });
  pack_b_pairs(P.w_head1, 64, 64, [&](int n, int k) { return Q.w_head1[n * 64 + k]; });
  pack_b_n8(P.w_head2, 64, [&](int n, int k) { return n < 3 ? Q.w_head2[n * 64 + k] : 0.f; });
  memcpy(P.b_base0, Q.b_base0, sizeof(P.b_base0));
  memcpy(P.b_base1, Q.b_base1, sizeof(P.b_base1));
  memcpy(P.b_head0, Q.b_head0, sizeof(P.b_head0));
  memcpy(P.b_head1, Q.b_head1, sizeof(P.b_head1));
  for (int n = 0; n < 3; ++n) P.b_head2[n] = Q.b_head2[n];
  P.feat_scale = feat_scale;
  P.inv_feat_scale = 1.f / feat_scale;
  P.avg_density = Q.avg_density;
}

// power of two that maps max|table| into fp16's comfortable range; synchronises `st`
static int table_feat_scale(const GridDev& g, cudaStream_t st, float* out) {
  unsigned int* d_max = nullptr;
  const size_t n = (size_t)g.num_levels * g.size * 2;
  SGN_CUDA(cudaMalloc(&d_max, 4));
  cudaMemsetAsync(d_max, 0, 4, st);
  k_max_abs<<<296, 256, 0, st>>>(reinterpret_cast<const float*>(g.table), n, d_max);
  count_launch();
  unsigned int bits = 0;
  cudaError_t e = cudaMemcpyAsync(&bits, d_max, 4, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_max);
  if (e != cudaSuccess) {
    set_error(std::string("hash-table scan failed: ") + cudaGetErrorString(e));
    return SGN_ERR_CUDA;
  }
  float maxabs;
  memcpy(&maxabs, &bits, 4);
  float scale = 1.f;
  if (maxabs > 0.f && std::isfinite(maxabs)) {
    int ex = (int)std::ceil(std::log2((double)maxabs));
    int sh = 14 - ex;
    if (sh > 60) sh = 60;
    if (sh < -60) sh = -60;
    scale = std::ldexp(1.f, sh);
  }
  *out = scale;
  return SGN_OK;
}

static bool lin_is(const SgnLinear& l, int in, int out) {
  return l.h_weight && l.h_bias && l.in_dim == in && l.out_dim == out;
}

}  // namespace sgn

using namespace sgn;

extern "C" const char* sgn_last_error(void) { return t_error.c_str(); }
extern "C" int sgn_abi_version(void) { return 1; }
extern "C" uint64_t sgn_launch_count(void) { return g_launches.load(); }

extern "C" int sgn_field_create(const SgnFieldDesc* d, SgnField** out) {
  SGN_CHECK_ARG(d && out, "null desc/out");
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error("no CUDA device visible: signerf_b200 has no CPU fallback");
    return SGN_ERR_NO_DEVICE;
  }
  // The tensor-core kernel is specialised for the nerfacto-default architecture.
  if (!(d->grid.num_levels == 16 && lin_is(d->base[0], 32, 64) && lin_is(d->base[1], 64, 16) &&
        lin_is(d->head[0], 63, 64) && lin_is(d->head[1], 64, 64) && lin_is(d->head[2], 64, 3))) {
    set_error("unsupported field architecture: expected 16-level F=2 grid, base 32->64->16, head 63->64->64->3");
    return SGN_ERR_UNSUPPORTED;
  }
  SGN_CHECK_ARG(d->h_appearance != nullptr, "null appearance embedding");
  SGN_CHECK_ARG(d->num_proposals == 0 || d->num_proposals == 2, "num_proposals must be 0 or 2");

  SgnField* f = new SgnField();
  cudaGetDevice(&f->device);
  f->avg_density = d->average_init_density;
  f->num_proposals = d->num_proposals;
  int rc = fill_grid(d->grid, &f->grid, "grid");
  if (rc != SGN_OK) { delete f; return rc; }

  float feat_scale = 1.f;
  rc = table_feat_scale(f->grid, 0, &feat_scale);
  if (rc != SGN_OK) { delete f; return rc; }

  // ---- pack MLPs
  std::vector<MlpPack> packv(1);
  std::vector<MlpF32> f32v(1);
  MlpPack& P = packv[0];
  MlpF32& Q = f32v[0];
  memset(&Q, 0, sizeof(Q));
  const float* B0 = d->base[0].h_weight;
  const float* B1 = d->base[1].h_weight;
  const float* H0 = d->head[0].h_weight;
  const float* H1 = d->head[1].h_weight;
  const float* H2 = d->head[2].h_weight;
  // head layer 0 in kernel order: k 0..15 = SH, k 16 = unused logit slot, k 17..31 = geo 0..14
  auto h0 = [&](int n, int k) -> float {
    if (k < 16) return H0[n * 63 + k];
    if (k == 16) return 0.f;
    return H0[n * 63 + 15 + (k - 16)];
  };
  for (int n = 0; n < 64; ++n) {
    double acc = d->head[0].h_bias[n];
    for (int a = 0; a < 32; ++a) acc += (double)H0[n * 63 + 31 + a] * (double)d->h_appearance[a];
    Q.b_head0[n] = (float)acc;
    Q.b_base0[n] = d->base[0].h_bias[n];
    Q.b_head1[n] = d->head[1].h_bias[n];
    for (int k = 0; k < 32; ++k) {
      Q.w_base0[n * 32 + k] = B0[n * 32 + k];
      Q.w_head0[n * 32 + k] = h0(n, k);
    }
    for (int k = 0; k < 64; ++k) Q.w_head1[n * 64 + k] = H1[n * 64 + k];
  }
  for (int n = 0; n < 16; ++n) {
    Q.b_base1[n] = d->base[1].h_bias[n];
    for (int k = 0; k < 64; ++k) Q.w_base1[n * 64 + k] = B1[n * 64 + k];
  }
  for (int n = 0; n < 3; ++n) {
    Q.b_head2[n] = d->head[2].h_bias[n];
    for (int k = 0; k < 64; ++k) Q.w_head2[n * 64 + k] = H2[n * 64 + k];
  }
  Q.avg_density = d->average_init_density;
  pack_from_f32(Q, feat_scale, P);

  auto fail = [&](const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(cudaGetLastError()));
    sgn_field_destroy(f);
    return SGN_ERR_CUDA;
  };
  if (cudaMalloc(&f->d_pack, sizeof(MlpPack)) != cudaSuccess) return fail("cudaMalloc pack");
  if (cudaMalloc(&f->d_f32, sizeof(MlpF32)) != cudaSuccess) return fail("cudaMalloc f32");
  if (cudaMemcpy(f->d_pack, &P, sizeof(P), cudaMemcpyHostToDevice) != cudaSuccess) return fail("upload pack");
  if (cudaMemcpy(f->d_f32, &Q, sizeof(Q), cudaMemcpyHostToDevice) != cudaSuccess) return fail("upload f32");

  for (int i = 0; i < d->num_proposals; ++i) {
    PropDev& pd = f->h_prop[i];
    memset(&pd, 0, sizeof(pd));
    rc = fill_grid(d->prop_grid[i], &pd.grid, "prop_grid");
    if (rc != SGN_OK) { sgn_field_destroy(f); return rc; }
    const int K = 2 * d->prop_grid[i].num_levels;
    if (!(K == 10 && lin_is(d->prop_mlp[i][0], 10, 16) && lin_is(d->prop_mlp[i][1], 16, 1))) {
      set_error("unsupported proposal network: expected 5-level grid + MLP 10->16->1");
      sgn_field_destroy(f);
      return SGN_ERR_UNSUPPORTED;
    }
    memcpy(pd.w0, d->prop_mlp[i][0].h_weight, sizeof(pd.w0));
    memcpy(pd.b0, d->prop_mlp[i][0].h_bias, sizeof(pd.b0));
    memcpy(pd.w1, d->prop_mlp[i][1].h_weight, sizeof(pd.w1));
    pd.b1 = d->prop_mlp[i][1].h_bias[0];
    pd.avg_density = d->average_init_density;
    if (cudaMalloc(&f->d_prop[i], sizeof(PropDev)) != cudaSuccess) return fail("cudaMalloc prop");
    if (cudaMemcpy(f->d_prop[i], &pd, sizeof(pd), cudaMemcpyHostToDevice) != cudaSuccess) return fail("upload prop");
  }
  *out = f;
  return SGN_OK;
}

extern "C" int sgn_field_refresh(SgnField* f, void* stream) {
  SGN_CHECK_ARG(f != nullptr, "null field");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float feat_scale = 1.f;
  int rc = table_feat_scale(f->grid, st, &feat_scale);
  if (rc != SGN_OK) return rc;
  std::vector<MlpF32> q(1);
  std::vector<MlpPack> p(1);
  SGN_CUDA(cudaMemcpyAsync(q.data(), f->d_f32, sizeof(MlpF32), cudaMemcpyDeviceToHost, st));
  SGN_CUDA(cudaStreamSynchronize(st));
  pack_from_f32(q[0], feat_scale, p[0]);
  SGN_CUDA(cudaMemcpyAsync(f->d_pack, p.data(), sizeof(MlpPack), cudaMemcpyHostToDevice, st));
  SGN_CUDA(cudaStreamSynchronize(st));
  return SGN_OK;
}

extern "C" void sgn_field_destroy(SgnField* f) {
  if (!f) return;
  if (f->d_pack) cudaFree(f->d_pack);
  if (f->d_f32) cudaFree(f->d_f32);
  for (int i = 0; i < 2; ++i)
    if (f->d_prop[i]) cudaFree(f->d_prop[i]);
  delete f;
}
