// K5: the dense contractions of the SDXL / ControlNet UNet on the 5th-generation tensor cores.
//
//   C[M,N] = A[M,K] . W[N,K]^T  (+ bias, + per-image row bias, + residual, GEGLU)      fp16 x fp16 -> fp32 (TMEM)
//
// One persistent warp-specialised kernel serves both operand shapes:
//   linear : A is a row-major [M,K] fp16 matrix (tokens x channels)              -> 2-D TMA tiles
//   conv3x3: A is an NHWC fp16 image, M = B*H*W output pixels, K = 9*Cin; the im2col is IMPLICIT: for tap (dy,dx)
//            and channel block c0 the producer issues one 4-D TMA box {64 ch, 16 px, 8 rows, 1 image} at
//            (c0, x0+dx-1, y0+dy-1, b); out-of-bounds rows/columns are zero-filled by TMA = the conv's zero padding.
// Two launch shapes: single CTAs (128 x N tiles), or CTA PAIRS (tcgen05 cta_group::2, 256 x N tiles): the two CTAs of a
// cluster own vertically adjacent M tiles of the same N tile, each loads its 128 rows of A and only HALF of the weight
// (B) tile, and the leader CTA issues one 256-row MMA that reads both halves.  Per-SM ingest drops from A+B to A+B/2
// per k-block (measured: the TMA -> shared-memory ingest rate of an SM, ~70 B/clk, bounds the single-CTA 128x256 tile
// at ~55 % of the tensor pipe) and the stage shrinks to 32 KB, which buys 6 pipeline stages instead of 4.
// Roles: warp 0 = TMA producer (1 lane), warp 1 = tcgen05.mma issuer (1 lane) + TMEM allocator, warps 2-5 = epilogue
// (TMEM -> registers -> global).  4-stage smem ring (full/empty mbarriers), 2 accumulator buffers in TMEM
// (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "sgn_common.cuh"
#include "sgn_tc.cuh"

namespace sgn {

// ------------------------------------------------------------------ tensor maps (host)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, const uint32_t* elem_strides) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available (no CUDA driver?)");
    return SGN_ERR_NO_DEVICE;
  }
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides ? elem_strides[i] : 1;
    if (i + 1 < rank) s[i] = strides_bytes[i];
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return SGN_ERR_CUDA;
  }
  return SGN_OK;
}

// ------------------------------------------------------------------ kernel
constexpr int kGemmThreads = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two warps per TMEM lane quarter)
constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kStageA = kBM * kBK * 2;    // 16 KB
template <int kCluster> struct GemmCfg {
  static constexpr int kStageB = (256 / kCluster) * kBK * 2;  // 32 KB (single CTA) / 16 KB (half of B per pair CTA)
};
// pair shape: 6 stages (192 KB) by default; 5 stages (160 KB) measure the same in isolation and leave room for a small
// co-resident CTA (sgn_set_option "gemm_pair_stages")
constexpr int kEpiWarps = 8;
constexpr int kEpiStage = 32 * 32 * 4;    // one 32-row x 32-column fp32 chunk per epilogue warp (transpose staging)
constexpr size_t gemm_smem_bytes(int cluster, int stages) {
  return 1024 + (size_t)stages * (kStageA + (256 / cluster) * kBK * 2) + kEpiWarps * kEpiStage + 256;
}
int g_pair_stages = 6;
constexpr int kAccStride = 256;           // TMEM columns between the two accumulator buffers
constexpr int kConvTileW = 16, kConvTileH = 8;

struct GemmParams {
  int M, N, K;  // N = weight rows covered by tiles (multiple of block_n); n_valid <= N columns are stored
  int n_valid;
  int block_n, num_m_tiles, num_n_tiles, num_k_blocks;
  // Tail split: work items [0, tail_start) are full tiles; the tiles of the last, under-filled round are cut into
  // tail_split column slices of tail_bn = block_n / tail_split each (work items tail_start .. total_items), so that a
  // 2.16-wave GEMM (160 tiles on 74 CTA pairs) costs ~2.4 tile times instead of 3.  tail_split = 1: off.
  int tail_start, tail_split, tail_bn, total_items;
  int conv, H, W, cin_blocks, tiles_x, tiles_y;
  const float* bias;
  const float* rowbias;
  int rows_per_batch;
  const float* residual;
  void* out;
  long long ldo;
  int out_f16, geglu, nchw, act_silu;
  int res_prefetch;  // 0: fetch the residual inside the epilogue (A/B knob SGN_GEMM_RES_PREFETCH=0)
  int coalesced;     // 1: 32-column chunks go through the shared-memory transpose (row-contiguous global accesses)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

// One 16-column group of one output row.
// Row-per-thread path: ragged / NCHW chunks, 16-column tails, and everything when `coalesced` is off.
__device__ __forceinline__ void epilogue16(const GemmParams& p, const uint32_t* acc, long long m, int b, int n0) {
  float v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(acc[j]);
  if (n0 + 16 <= p.n_valid && !p.nchw) {
    if (p.bias) {
      const float4* bp = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = __ldg(bp + q);
        v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
      }
    }
    if (p.rowbias) {
      const float4* bp = reinterpret_cast<const float4*>(p.rowbias + (size_t)b * p.n_valid + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = __ldg(bp + q);
        v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
      }
    }
    if (p.geglu) {  // weight rows interleaved (value, gate): out[n/2] = value * gelu(gate)
      __half2 h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        h[q] = __floats2half2_rn(v[4 * q] * gelu_erf(v[4 * q + 1]), v[4 * q + 2] * gelu_erf(v[4 * q + 3]));
      *reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + m * p.ldo + (n0 >> 1)) =
          *reinterpret_cast<uint4*>(h);
      return;
    }
    if (p.residual) {
      const float4* rp = reinterpret_cast<const float4*>(p.residual + m * p.ldo + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 t = rp[q];
        v[4 * q] += t.x, v[4 * q + 1] += t.y, v[4 * q + 2] += t.z, v[4 * q + 3] += t.w;
      }
    }
    if (p.act_silu) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = v[j] / (1.f + __expf(-v[j]));
    }
    if (p.out_f16) {
      __half2 h[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) h[q] = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
      uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.out) + m * p.ldo + n0);
      op[0] = reinterpret_cast<uint4*>(h)[0];
      op[1] = reinterpret_cast<uint4*>(h)[1];
    } else {
      float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + m * p.ldo + n0);
#pragma unroll
      for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    return;
  }
  // ragged / NCHW tail (the UNet's 4-channel output conv): scalar, bounds-checked, fp32 only
  for (int j = 0; j < 16; ++j) {
    int n = n0 + j;
    if (n >= p.n_valid) break;
    float x = v[j];
    if (p.bias) x += __ldg(p.bias + n);
    if (p.rowbias) x += __ldg(p.rowbias + (size_t)b * p.n_valid + n);
    if (p.nchw) {
      long long hw = (long long)p.H * p.W;
      long long pix = m - (long long)b * hw;
      long long o = ((long long)b * p.n_valid + n) * hw + pix;
      if (p.residual) x += p.residual[o];
      reinterpret_cast<float*>(p.out)[o] = x;
    } else {
      if (p.residual) x += p.residual[m * p.ldo + n];
      if (p.act_silu) x = x / (1.f + __expf(-x));
      if (p.out_f16) reinterpret_cast<__half*>(p.out)[m * p.ldo + n] = __float2half_rn(x);
      else reinterpret_cast<float*>(p.out)[m * p.ldo + n] = x;
    }
  }
}

// Cold path kept out of line (it would otherwise cost the hot epilogue its registers): re-reads `ncols` (16 or 32)
// accumulator columns of this warp's lane quarter from TMEM and stores them row-per-thread.  Warp-collective.
__device__ __noinline__ void epilogue_rows(const GemmParams& p, uint32_t taddr, int ncols, long long m, int b, int n0,
                                           bool valid) {
  for (int c = 0; c < ncols; c += 16) {
    uint32_t acc[16];
    tc::tmem_ld16(taddr + c, acc);
    tc::tmem_ld_wait();
    if (valid) epilogue16(p, acc, m, b, n0 + c);
  }
}

// kEpi specialises the hot (coalesced) epilogue at compile time; the instruction footprint of the generic one
// (erff-based GEGLU, SiLU, fp16 / fp32 stores all resident) misses the instruction cache with only 8 epilogue warps.
//   0 generic (runtime flags)   1 fp32 output (+bias, +rowbias, +residual)   2 fp16 output (+bias)   3 GEGLU
template <int kCluster, int kStages, int kEpi>
__global__ void __launch_bounds__(kGemmThreads, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
          const __grid_constant__ CUtensorMap tmBt, const __grid_constant__ GemmParams p) {
  constexpr int kStageB = GemmCfg<kCluster>::kStageB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kStageA;
  uint8_t* sEpi = sB + kStages * kStageB;
  uint64_t* full = reinterpret_cast<uint64_t*>(sEpi + kEpiWarps * kEpiStage);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full[s], 1);          // pair: the leader's expect_tx covers the bytes both CTAs load
      tc::mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&tfull[s], 1);
      tc::mbar_init(&tempty[s], 8 * kCluster);  // pair: the epilogue warps of both CTAs release the leader's buffer
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) {
    if (kCluster == 2) tc::tmem_alloc_pair(tmem_slot, 512);
    else tc::tmem_alloc(tmem_slot, 512);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (kCluster > 1) tc::cluster_sync_all();  // peer barriers initialised before any remote arrive / completion
  tc::tc_fence_after();
  // Programmatic dependent launch (launch_gemm_c): everything above - tensor-map prefetch, barrier init, TMEM allocation,
  // the cluster hand-shake - ran while the previous kernel on the stream was still draining its last CTAs; nothing that
  // kernel wrote is read, and nothing is written, before this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t cta_rank = kCluster > 1 ? tc::cluster_ctarank() : 0;
  // work unit = kCluster vertically adjacent M tiles x one N tile; M tiles past the end are computed on zero-filled
  // rows and never stored
  const int m_groups = (p.num_m_tiles + kCluster - 1) / kCluster;
  (void)m_groups;
  const int total_tiles = p.total_items;
  const int first_tile = blockIdx.x / kCluster, tile_step = gridDim.x / kCluster;
  // work item -> (M group, first column, width); tail items are column slices of the last round's tiles
  auto decode = [&](int t, int& m_group, int& n0, int& bn) -> bool {
    int tile = t, sub = 0;
    const bool tail = t >= p.tail_start;
    if (tail) {
      const int u = t - p.tail_start;
      tile = p.tail_start + u / p.tail_split;
      sub = u - (u / p.tail_split) * p.tail_split;
    }
    m_group = tile / p.num_n_tiles;
    const int n_tile = tile - m_group * p.num_n_tiles;
    bn = tail ? p.tail_bn : p.block_n;
    n0 = n_tile * p.block_n + sub * p.tail_bn;
    return tail;
  };
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step) {
        int m_group, n0, bn;
        const bool tail = decode(t, m_group, n0, bn);
        const CUtensorMap* tmBx = tail ? &tmBt : &tmB;
        const int b_rows = bn / kCluster;  // weight rows this CTA loads per k-block
        const uint32_t tx_bytes = kCluster * (kStageA + (uint32_t)b_rows * (kBK * 2));  // both CTAs complete on the leader
        const int m_tile = m_group * kCluster + (int)cta_rank;
        int b = 0, x0 = 0, y0 = 0;
        if (p.conv) {
          b = m_tile / tiles_per_img;
          int r = m_tile - b * tiles_per_img;
          int ty = r / p.tiles_x;
          y0 = ty * kConvTileH;
          x0 = (r - ty * p.tiles_x) * kConvTileW;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          tc::mbar_wait(&empty[stage], phase ^ 1);
          int tap = 0, cb = kb, dy = 0, dx = 0;
          if (p.conv) {
            tap = kb / p.cin_blocks, cb = kb - tap * p.cin_blocks;
            dy = tap / 3, dx = tap - dy * 3;
          }
          if (kCluster == 1) {
            tc::mbar_expect_tx(&full[stage], tx_bytes);
            if (p.conv) tc::tma_load_4d(sA + stage * kStageA, &tmA, &full[stage], cb * kBK, x0 + dx - 1, y0 + dy - 1, b);
            else tc::tma_load_2d(sA + stage * kStageA, &tmA, &full[stage], kb * kBK, m_tile * kBM);
            tc::tma_load_2d(sB + stage * kStageB, tmBx, &full[stage], kb * kBK, n0);
          } else {
            const uint32_t lead_full = tc::mapa_u32(&full[stage], 0);
            if (cta_rank == 0) tc::mbar_expect_tx(&full[stage], tx_bytes);
            if (p.conv) tc::tma_load_4d_pair(sA + stage * kStageA, &tmA, lead_full, cb * kBK, x0 + dx - 1, y0 + dy - 1, b);
            else tc::tma_load_2d_pair(sA + stage * kStageA, &tmA, lead_full, kb * kBK, m_tile * kBM);
            tc::tma_load_2d_pair(sB + stage * kStageB, tmBx, lead_full, kb * kBK, n0 + (int)cta_rank * b_rows);
          }
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && cta_rank == 0) {  // ---------------- MMA issuer (pair: the leader issues for both SMs)
      const uint32_t idesc_full = tc::umma_idesc_f16(kBM * kCluster, p.block_n, false, false);
      const uint32_t idesc_tail = tc::umma_idesc_f16(kBM * kCluster, p.tail_bn, false, false);
      int stage = 0;
      uint32_t phase = 0;
      int i = 0;
      for (int t = first_tile; t < total_tiles; t += tile_step, ++i) {
        const int buf = i & 1;
        const uint32_t ph = (i >> 1) & 1;
        tc::mbar_wait(&tempty[buf], ph ^ 1);
        tc::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kAccStride;
        const uint32_t idesc = t >= p.tail_start ? idesc_tail : idesc_full;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          tc::mbar_wait(&full[stage], phase);
          tc::tc_fence_after();
          const uint64_t da = tc::umma_desc_sw128(tc::smem_u32(sA + stage * kStageA));
          const uint64_t db = tc::umma_desc_sw128(tc::smem_u32(sB + stage * kStageB));
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {  // +32 B per K=16 step inside the 128-B swizzled row
            if (kCluster == 1) tc::umma_f16_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            else tc::umma_f16_ss_pair(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          }
          if (kCluster == 1) tc::umma_commit(&empty[stage]);
          else tc::umma_commit_pair(&empty[stage]);
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
        if (kCluster == 1) tc::umma_commit(&tfull[buf]);
        else tc::umma_commit_pair(&tfull[buf]);
      }
    }
  } else {  // ---------------- epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const int col_half = (warp - 2) >> 2;
    // The in-place fp32 residual is the one operand that is not TMA-staged: its loads are latency x registers bound
    // (8 float4 per thread in flight), so every thread asks L2 for the residual segment of its NEXT tile before it
    // starts on the current one; by the time those loads are issued the lines are L2 hits.
    auto l2_prefetch_tile = [&](int t) {
      if (!p.res_prefetch || p.residual == nullptr || p.nchw || p.geglu || t >= total_tiles) return;
      int m_group, n_base, bn;
      decode(t, m_group, n_base, bn);
      const int m_tile = m_group * kCluster + (int)cta_rank;
      long long m;
      if (p.conv) {
        const int b = m_tile / tiles_per_img;
        const int r = m_tile - b * tiles_per_img;
        const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
        const int y = ty * kConvTileH + row / kConvTileW, x = tx * kConvTileW + row % kConvTileW;
        if (!(y < p.H && x < p.W && m_tile < p.num_m_tiles)) return;
        m = ((long long)b * p.H + y) * p.W + x;
      } else {
        m = (long long)m_tile * kBM + row;
        if (m >= p.M) return;
      }
      for (int c = col_half * 32; c < bn && n_base + c < p.n_valid; c += 64)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.residual + m * p.ldo + n_base + c));
    };
    l2_prefetch_tile(first_tile);
    // Coalesced path: a 32 x 32 chunk (thread = accumulator row out of TMEM) is transposed through a 4 KB swizzled
    // shared-memory patch so that 8 lanes cover 128 contiguous bytes of one output row: every global residual load /
    // output store instruction then touches 4 full lines instead of 32 partial ones (the row-per-thread pattern costs
    // 8x the L1 wavefronts, which bounded the short-K GEMMs of the transformer blocks).
    float4* patch = reinterpret_cast<float4*>(sEpi + (warp - 2) * kEpiStage);   // [32 rows][8 float4], slot ^ (row & 7)
    const int cq = lane & 7, rsub = lane >> 3;
    int i = 0;
    for (int t = first_tile; t < total_tiles; t += tile_step, ++i) {
      l2_prefetch_tile(t + tile_step);
      const int buf = i & 1;
      const uint32_t ph = (i >> 1) & 1;
      int m_group, n_base, bn;
      decode(t, m_group, n_base, bn);
      const int m_tile = m_group * kCluster + (int)cta_rank;
      int tb = 0, ty0 = 0, tx0 = 0;
      if (p.conv) {
        tb = m_tile / tiles_per_img;
        int r = m_tile - tb * tiles_per_img;
        int ty = r / p.tiles_x;
        ty0 = ty * kConvTileH, tx0 = (r - ty * p.tiles_x) * kConvTileW;
      }
      // Output row of patch row (q*4 + rsub) of this warp's lane quarter: m_q = m_base + (q>>2)*step_hi + (q&3)*4
      // (linear: consecutive rows, step_hi = 16; conv: the tile is 8 image rows x 16 pixels, step_hi = W).
      const bool tile_ok = m_tile < p.num_m_tiles;
      const int ybase = ty0 + lane_base / kConvTileW, xbase = tx0 + rsub;
      const long long m_base = p.conv ? ((long long)tb * p.H + ybase) * p.W + xbase : (long long)m_tile * kBM + lane_base + rsub;
      const int step_hi = p.conv ? p.W : 16;
      auto row_q = [&](int q, long long& m_out) -> bool {
        m_out = m_base + (q >> 2) * step_hi + (q & 3) * 4;
        return p.conv ? (tile_ok && ybase + (q >> 2) < p.H && xbase + (q & 3) * 4 < p.W) : (m_out < p.M);
      };
      long long m;
      int b = 0;
      bool valid;
      if (p.conv) {
        const int y = ty0 + row / kConvTileW, x = tx0 + row % kConvTileW;
        valid = y < p.H && x < p.W && tile_ok;
        m = ((long long)tb * p.H + y) * p.W + x;
        b = tb;
      } else {
        m = (long long)m_tile * kBM + row;
        valid = m < p.M;
        b = p.rows_per_batch > 0 ? (int)(m / p.rows_per_batch) : 0;
      }
      // the two warps of a lane quarter split the tile's columns in 32-wide chunks: even chunks / odd chunks
      int c0 = col_half * 32;
      auto chunk_fast = [&](int c) -> bool { return p.coalesced && c + 32 <= bn && n_base + c + 32 <= p.n_valid; };
      // The fp32 residual (the stream the UNet keeps adding into) is the epilogue's only global read: the first chunk's
      // is fetched before the accumulator is even complete, the next chunk's while the current one is processed.
      const bool res_pre = p.res_prefetch && p.residual != nullptr && !p.geglu;
      // one residual buffer: a second one (next chunk in flight during this chunk's stores) spills, and with 230 KB of
      // shared memory there is no L1 left to catch spills (measured: 38.6 vs 34.3 us on 8192x1280x1280)
      float4 rcur[8];
      auto prefetch = [&](int c, float4 (&r)[8]) -> bool {
        if (!res_pre || !chunk_fast(c)) return false;
        const int n0 = n_base + c + cq * 4;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          long long mq;
          r[q] = row_q(q, mq) ? *reinterpret_cast<const float4*>(p.residual + mq * p.ldo + n0) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return true;
      };
      bool cur_ok = prefetch(c0, rcur);   // first chunk: in flight while the tile's last MMAs complete
      tc::mbar_wait(&tfull[buf], ph);
      tc::tc_fence_after();
      const uint32_t taddr = tmem_base + buf * kAccStride + ((uint32_t)lane_base << 16);
      for (; c0 + 32 <= bn; c0 += 64) {
        if (!chunk_fast(c0)) {
          epilogue_rows(p, taddr + c0, 32, m, b, n_base + c0, valid);
          cur_ok = prefetch(c0 + 64, rcur);
          continue;
        }
        uint32_t acc[32];
        tc::tmem_ld32(taddr + c0, acc);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)   // row `lane` of the patch, float4 slot q at swizzled position q ^ (lane & 7)
          patch[lane * 8 + (q ^ (lane & 7))] = make_float4(__uint_as_float(acc[4 * q]), __uint_as_float(acc[4 * q + 1]),
                                                            __uint_as_float(acc[4 * q + 2]), __uint_as_float(acc[4 * q + 3]));
        __syncwarp();
        const bool e_rowbias = kEpi != 3 && p.rowbias != nullptr;
        const bool e_geglu = kEpi == 3 || (kEpi == 0 && p.geglu);
        const bool e_res = (kEpi == 0 || kEpi == 1) && p.residual != nullptr;
        const bool e_silu = kEpi == 0 && p.act_silu;
        const bool e_f16 = kEpi == 2 || (kEpi == 0 && p.out_f16);
        const int n0 = n_base + c0 + cq * 4;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + n0));
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int rr = q * 4 + rsub;
          float4 v = patch[rr * 8 + (cq ^ (rr & 7))];
          long long mq;
          if (!row_q(q, mq)) continue;
          v.x += bias4.x, v.y += bias4.y, v.z += bias4.z, v.w += bias4.w;
          if (e_rowbias) {
            const int bq = p.conv ? tb : (int)(mq / p.rows_per_batch);
            const float4 t4 = __ldg(reinterpret_cast<const float4*>(p.rowbias + (size_t)bq * p.n_valid + n0));
            v.x += t4.x, v.y += t4.y, v.z += t4.z, v.w += t4.w;
          }
          if (e_geglu) {  // (value, gate) pairs -> 2 outputs
            const __half2 h = __floats2half2_rn(v.x * gelu_erf(v.y), v.z * gelu_erf(v.w));
            *reinterpret_cast<__half2*>(reinterpret_cast<__half*>(p.out) + mq * p.ldo + (n0 >> 1)) = h;
            continue;
          }
          if (e_res) {
            const float4 t4 = cur_ok ? rcur[q] : *reinterpret_cast<const float4*>(p.residual + mq * p.ldo + n0);
            v.x += t4.x, v.y += t4.y, v.z += t4.z, v.w += t4.w;
          }
          if (e_silu) {
            v.x = v.x / (1.f + __expf(-v.x)), v.y = v.y / (1.f + __expf(-v.y));
            v.z = v.z / (1.f + __expf(-v.z)), v.w = v.w / (1.f + __expf(-v.w));
          }
          if (e_f16) {
            __half2 h[2] = {__floats2half2_rn(v.x, v.y), __floats2half2_rn(v.z, v.w)};
            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + mq * p.ldo + n0) = *reinterpret_cast<uint2*>(h);
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + mq * p.ldo + n0) = v;
          }
        }
        __syncwarp();   // the patch is rewritten by the next chunk
        cur_ok = prefetch(c0 + 64, rcur);   // L2 hits (l2_prefetch_tile), in flight across the next TMEM load + transpose
      }
      if (c0 < bn && c0 + 16 == bn)  // 16-column tail (width % 32 == 16) belongs to whoever reaches it
        epilogue_rows(p, taddr + c0, 16, m, b, n_base + c0, valid);
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kCluster == 1) tc::mbar_arrive(&tempty[buf]);
        else tc::mbar_arrive_cluster(tc::mapa_u32(&tempty[buf], 0));
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (kCluster > 1) tc::cluster_sync_all();  // no CTA leaves while its peer can still multicast into it / arrive on it
  if (warp == 1) {
    __syncwarp();
    if (kCluster == 2) tc::tmem_dealloc_pair(tmem_base, 512);
    else tc::tmem_dealloc(tmem_base, 512);
  }
}

// block_n (multiple of 16, <= 256; the last N tile may be ragged: TMA zero-fills the missing weight rows and the
// epilogue skips their columns).  Cost model fitted on B200: a k-block costs max(tensor time 2*bn cycles, load time
// (16 KB + 64 B * bn) / 58 B/clk) -- the load side is latency x in-flight-bytes bound (192 KB of stages per SM against
// ~1.7 us of loaded L2 latency), so wide tiles win even when they leave a partially filled last wave.
// `epi_res_kb`: > 0 for an fp32 in-place-residual GEMM with that many k-blocks.  With K <= 640 its tile time is set by the
// epilogue (~90 cycles per output column per CTA: 1 KB of residual read + write at the ~11 B/clk an SM's epilogue warps
// sustain), so the width that minimises rounds x columns wins, not the widest one (32768x640x640: 64.2 -> 54.8 us).
// column slices the tiles of an under-filled last round are cut into (1 = no split): see GemmParams::tail_split
static int tail_split_for(int bn, long long rem, long long units) {
  static const int enabled = [] { const char* e = getenv("SGN_GEMM_TAIL_SPLIT"); return e ? atoi(e) : 1; }();
  if (!enabled || rem == 0 || rem * 2 > units) return 1;
  int split = (int)std::min<long long>(4, units / rem);
  while (split > 1 && (bn % (split * 32) != 0 || bn / split < 64)) --split;
  return std::max(1, split);
}

static int pick_block_n(long long m_tiles, int N, int epi_res_kb = 0) {
  if (const char* e = getenv("SGN_GEMM_BN")) {  // tuning / debugging override
    int bn = atoi(e);
    if (bn >= 16 && bn <= 256 && bn % 16 == 0) return std::min(bn, (N + 15) / 16 * 16);
  }
  const int n16 = (N + 15) / 16 * 16;
  const long long units = std::max(1, sm_count() / 2);        // CTA pairs
  const long long m_groups = (m_tiles + 1) / 2;
  int best = 16;
  double best_cost = 1e30;
  for (int bn = std::min(256, n16); bn >= 16; bn -= 16) {
    long long tiles = m_groups * ((n16 + bn - 1) / bn);
    long long rounds = (tiles + units - 1) / units;
    double t = std::max(2.0 * bn, 282.0 + 1.1 * bn);
    double cost = (double)rounds * (t + 40.0);                 // + per-tile epilogue hand-over
    if (m_tiles >= 2 && tiles >= units) {                      // CTA pairs: an under-filled last round is cut into slices
      const int split = tail_split_for(bn, tiles % units, units);
      if (split > 1) cost = ((double)(tiles / units) + 1.6 / split) * (t + 40.0);   // narrow slices run ~1.6x slower per column
    }
    if (epi_res_kb > 0 && epi_res_kb <= 10) cost = (double)rounds * std::max(epi_res_kb * t, 90.0 * bn);
    if (cost < best_cost - 1e-9) best_cost = cost, best = bn;
  }
  return best;
}

// Cluster of 2 (multicast weight tiles) whenever there are at least two M tiles to pair; SGN_GEMM_CLUSTER=1 forces the
// single-CTA shape (A/B comparison, debugging).
static int pick_cluster(const GemmParams& p) {
  static const int forced = [] {
    const char* e = getenv("SGN_GEMM_CLUSTER");
    return e ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return p.num_m_tiles >= 2 ? forced : 1;
  return p.num_m_tiles >= 2 ? 2 : 1;
}

static const int g_gemm_pdl = [] {
  const char* e = getenv("SGN_GEMM_PDL");
  return e == nullptr || atoi(e) != 0;
}();

template <int kCluster, int kStages, int kEpi>
static int launch_gemm_c(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBt, const GemmParams& p,
                         cudaStream_t st) {
  constexpr size_t smem = gemm_smem_bytes(kCluster, kStages);
  static bool attr_set = false;
  if (!attr_set) {
    SGN_CUDA(cudaFuncSetAttribute(k_gemm_tc<kCluster, kStages, kEpi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int grid = std::min(p.total_items, sm_count() / kCluster) * kCluster;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kGemmThreads), cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCluster, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  // programmatic dependent launch: the GEMM's CTAs may be scheduled (and run their prologue up to griddepcontrol.wait) as
  // soon as the previous kernel's CTAs leave the SMs, instead of after its completion + a launch latency (690 GEMM launches
  // per UNet step).  SGN_GEMM_PDL=0 restores the plain stream order.
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr, cfg.numAttrs = g_gemm_pdl ? 2 : 1;
  SGN_CUDA(cudaLaunchKernelEx(&cfg, k_gemm_tc<kCluster, kStages, kEpi>, tmA, tmB, tmBt, p));
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

// Tail split (see GemmParams): decided on the host once the tile shape is known; `tmBt` is the weight tensor map with the
// narrower box.  Only when the last round is at most half full and the slices stay >= 64 columns wide.
static int plan_tail(GemmParams& p, int cluster, const void* d_w, uint64_t k_elems, uint64_t n_rows_w, uint64_t ldw_bytes,
                     const CUtensorMap& tmB, CUtensorMap* tmBt) {
  const int units = std::max(1, sm_count() / cluster);
  const int tiles = (p.num_m_tiles + cluster - 1) / cluster * p.num_n_tiles;
  p.tail_start = tiles, p.tail_split = 1, p.tail_bn = p.block_n, p.total_items = tiles;
  *tmBt = tmB;
  const int rem = tiles % units;
  if (cluster != 2 || tiles < units) return SGN_OK;
  const int split = tail_split_for(p.block_n, rem, units);
  if (split < 2) return SGN_OK;
  p.tail_start = tiles - rem, p.tail_split = split, p.tail_bn = p.block_n / split;
  p.total_items = p.tail_start + rem * split;
  uint64_t dw[2] = {k_elems, n_rows_w}, sw[1] = {ldw_bytes};
  uint32_t bw[2] = {kBK, (uint32_t)(p.tail_bn / cluster)};
  return encode_tmap(tmBt, d_w, 2, dw, sw, bw, nullptr);
}

static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBt, GemmParams& p, int cluster,
                       cudaStream_t st) {
  if (cluster != 2) return launch_gemm_c<1, 4, 0>(tmA, tmB, tmBt, p, st);
  if (g_pair_stages == 5) return launch_gemm_c<2, 5, 0>(tmA, tmB, tmBt, p, st);
  static const int spec = [] { const char* e = getenv("SGN_GEMM_EPI_SPEC"); return e ? atoi(e) : 1; }();
  if (spec && p.coalesced == 1 && !p.act_silu) {
    if (p.geglu && !p.rowbias) return launch_gemm_c<2, 6, 3>(tmA, tmB, tmBt, p, st);
    if (p.out_f16 && !p.geglu && !p.residual) return launch_gemm_c<2, 6, 2>(tmA, tmB, tmBt, p, st);
    if (!p.out_f16 && !p.geglu) return launch_gemm_c<2, 6, 1>(tmA, tmB, tmBt, p, st);
  }
  return launch_gemm_c<2, 6, 0>(tmA, tmB, tmBt, p, st);
}

static int fill_epilogue(GemmParams& p, const SgnEpilogue* ep, int n_valid, void* d_out) {
  p.bias = ep ? ep->d_bias : nullptr;
  p.rowbias = ep ? ep->d_rowbias : nullptr;
  p.rows_per_batch = ep ? ep->rows_per_batch : 0;
  p.residual = ep ? ep->d_residual : nullptr;
  p.out_f16 = ep ? ep->out_f16 : 0;
  p.geglu = ep ? ep->geglu : 0;
  p.nchw = ep ? ep->nchw : 0;
  p.act_silu = ep ? ep->act_silu : 0;
  static const int res_prefetch = [] { const char* e = getenv("SGN_GEMM_RES_PREFETCH"); return e ? atoi(e) : 1; }();
  p.res_prefetch = res_prefetch;
  static const int coalesced = [] { const char* e = getenv("SGN_GEMM_COALESCED"); return e ? atoi(e) : 1; }();
  p.out = d_out;
  long long ld = ep ? ep->ldo : 0;
  p.ldo = ld > 0 ? ld : (p.geglu ? n_valid / 2 : n_valid);
  p.coalesced = coalesced && !p.nchw && (p.ldo % ((p.out_f16 || p.geglu) ? 8 : 4)) == 0 && (n_valid % 4) == 0;
  SGN_CHECK_ARG(!p.rowbias || p.rows_per_batch > 0, "rowbias needs rows_per_batch");
  SGN_CHECK_ARG(!p.geglu || (n_valid % 16 == 0 && !p.residual && !p.nchw), "geglu needs N % 16 == 0 and no residual");
  SGN_CHECK_ARG(!p.nchw || !p.out_f16, "nchw output is fp32 only");
  SGN_CHECK_ARG(!p.act_silu || (!p.geglu && !p.nchw), "act_silu does not combine with geglu / nchw");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_out) & 15) == 0, "output must be 16-byte aligned");
  SGN_CHECK_ARG(p.nchw || n_valid % 16 != 0 || (p.ldo % (p.out_f16 || p.geglu ? 8 : 4)) == 0, "ldo breaks 16-byte rows");
  return SGN_OK;
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_gemm_f16(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, int M, int N, int K,
                            const SgnEpilogue* ep, void* d_out, void* stream) {
  SGN_CHECK_ARG(M >= 0 && N > 0 && K > 0, "bad GEMM shape");
  if (M == 0) return SGN_OK;
  SGN_CHECK_ARG(d_a && d_w && d_out, "null pointer");
  SGN_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldw % 8 == 0 && lda >= K && ldw >= K, "K and strides must be multiples of 8");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0,
                "operands must be 16-byte aligned");
  GemmParams p{};
  p.M = M, p.K = K, p.n_valid = N;
  int n_rows = (N + 15) / 16 * 16;  // weight rows past N are zero-filled by TMA
  p.N = n_rows;
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.num_k_blocks = (K + kBK - 1) / kBK;
  const bool res32 = ep && ep->d_residual && !ep->out_f16 && !ep->geglu;
  p.block_n = pick_block_n(p.num_m_tiles, n_rows, res32 ? p.num_k_blocks : 0);
  p.num_n_tiles = (n_rows + p.block_n - 1) / p.block_n;
  p.conv = 0, p.tiles_x = p.tiles_y = 1;
  int rc = fill_epilogue(p, ep, N, d_out);
  if (rc) return rc;
  CUtensorMap tmA, tmB;
  uint64_t da[2] = {(uint64_t)K, (uint64_t)M}, sa[1] = {(uint64_t)lda * 2};
  uint32_t ba[2] = {kBK, kBM};
  rc = encode_tmap(&tmA, d_a, 2, da, sa, ba, nullptr);
  if (rc) return rc;
  const int cluster = pick_cluster(p);
  uint64_t dw[2] = {(uint64_t)K, (uint64_t)N}, sw[1] = {(uint64_t)ldw * 2};
  uint32_t bw[2] = {kBK, (uint32_t)(p.block_n / cluster)};
  rc = encode_tmap(&tmB, d_w, 2, dw, sw, bw, nullptr);
  if (rc) return rc;
  CUtensorMap tmBt;
  rc = plan_tail(p, cluster, d_w, (uint64_t)K, (uint64_t)N, (uint64_t)ldw * 2, tmB, &tmBt);
  if (rc) return rc;
  return launch_gemm(tmA, tmB, tmBt, p, cluster, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int sgn_gemm_plan(int M, int N, int K, int residual_f32, int* h_plan) {
  SGN_CHECK_ARG(M > 0 && N > 0 && K > 0 && h_plan, "bad arguments");
  GemmParams p{};
  p.num_m_tiles = (M + kBM - 1) / kBM;
  p.num_k_blocks = (K + kBK - 1) / kBK;
  const int n_rows = (N + 15) / 16 * 16;
  p.block_n = pick_block_n(p.num_m_tiles, n_rows, residual_f32 ? p.num_k_blocks : 0);
  p.num_n_tiles = (n_rows + p.block_n - 1) / p.block_n;
  const int cluster = pick_cluster(p);
  const int units = std::max(1, sm_count() / cluster);
  const int tiles = (p.num_m_tiles + cluster - 1) / cluster * p.num_n_tiles;
  int split = 1;
  if (cluster == 2 && tiles >= units) split = tail_split_for(p.block_n, tiles % units, units);
  const int rem = tiles % units;
  h_plan[0] = p.block_n, h_plan[1] = cluster, h_plan[2] = tiles, h_plan[3] = split;
  h_plan[4] = split > 1 ? tiles - rem + rem * split : tiles;
  h_plan[5] = units;
  return SGN_OK;
}

extern "C" int sgn_conv3x3_f16(const void* d_x, const void* d_w, int B, int H, int W, int C, int N,
                               const SgnEpilogue* ep, void* d_out, void* stream) {
  SGN_CHECK_ARG(B >= 0 && H > 0 && W > 0 && C > 0 && N > 0, "bad conv shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_x && d_w && d_out, "null pointer");
  SGN_CHECK_ARG(C % kBK == 0, "tensor-core conv needs Cin % 64 == 0 (use sgn_conv2d_direct otherwise)");
  SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_w) & 15) == 0,
                "operands must be 16-byte aligned");
  GemmParams p{};
  p.M = B * H * W, p.K = 9 * C, p.n_valid = N;
  int n_rows = (N + 15) / 16 * 16;
  p.N = n_rows;
  p.conv = 1, p.H = H, p.W = W, p.cin_blocks = C / kBK;
  p.tiles_x = (W + kConvTileW - 1) / kConvTileW;
  p.tiles_y = (H + kConvTileH - 1) / kConvTileH;
  p.num_m_tiles = B * p.tiles_x * p.tiles_y;
  p.block_n = pick_block_n(p.num_m_tiles, n_rows);
  p.num_n_tiles = (n_rows + p.block_n - 1) / p.block_n;
  p.num_k_blocks = 9 * p.cin_blocks;
  int rc = fill_epilogue(p, ep, N, d_out);
  if (rc) return rc;
  if (p.rowbias && p.rows_per_batch != H * W) {
    set_error("invalid argument: conv rowbias is per image (rows_per_batch must be H*W)");
    return SGN_ERR_INVALID_ARG;
  }
  CUtensorMap tmA, tmB;
  uint64_t da[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  uint64_t sa[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t ba[4] = {kBK, kConvTileW, kConvTileH, 1};
  rc = encode_tmap(&tmA, d_x, 4, da, sa, ba, nullptr);
  if (rc) return rc;
  const int cluster = pick_cluster(p);
  uint64_t dw[2] = {(uint64_t)9 * C, (uint64_t)N}, sw[1] = {(uint64_t)9 * C * 2};
  uint32_t bw[2] = {kBK, (uint32_t)(p.block_n / cluster)};
  rc = encode_tmap(&tmB, d_w, 2, dw, sw, bw, nullptr);
  if (rc) return rc;
  CUtensorMap tmBt;
  rc = plan_tail(p, cluster, d_w, (uint64_t)9 * C, (uint64_t)N, (uint64_t)9 * C * 2, tmB, &tmBt);
  if (rc) return rc;
  return launch_gemm(tmA, tmB, tmBt, p, cluster, reinterpret_cast<cudaStream_t>(stream));
}
