// K6: softmax(Q K^T / sqrt(d)) V for the SpatialTransformer blocks of the SDXL / ControlNet UNet (SURVEY §8a row A12:
// sgm CrossAttention, head_dim 64), both the self-attention over all sheet tokens (16 384 @640 ch, 4 096 @1280 ch at
// the 2048^2 sheet) and the cross-attention over the 77 prompt tokens, on tcgen05 with the score tile in TMEM.
//
// One CTA = one 128-query tile of one (image, head).  Roles: warp 0 TMA producer, warp 1 tcgen05.mma issuer,
// warps 2-5 softmax (thread i owns query row i = TMEM lane i).  Per 128-key tile j:
//     S_j  = Q K_j^T                 4 x UMMA 128x128x16 into TMEM S[j%2]            (issued one tile ahead)
//     P_j  = exp2(c S_j - c m_j)     TMEM -> registers (two passes: max, exp) -> fp16 -> swizzled smem P[j%2]
//     O_j  = P_j V_j                 8 x UMMA 128x64x16 (V is the MN-major B operand) into TMEM O[j%2]
//     acc  = acc * alpha_j + O_j     in registers of the softmax threads (the online-softmax rescale never touches TMEM)
// Keys beyond T_kv (ragged last tile, 77-token context) are masked to -inf before the max.
#include "sgn_common.cuh"
#include "sgn_tc.cuh"

namespace sgn {

constexpr int kAttnThreads = 192;
constexpr int kHeadDim = 64;
constexpr int kQTile = 128, kKvTile = 128;
constexpr int kKvStages = 3;
constexpr int kTileBytes = 128 * 64 * 2;  // one [128 x 64] fp16 tile, 128-B rows, SWIZZLE_128B
constexpr size_t kAttnSmem = 1024 + kTileBytes * (1 + 2 * kKvStages + 4) + 256;

struct AttnParams {
  int T_q, T_kv, n_kv_tiles;
  float scale_log2e;
  __half* out;
  long long ldo;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + kTileBytes;
  uint8_t* sV = sK + kKvStages * kTileBytes;
  uint8_t* sP = sV + kKvStages * kTileBytes;  // 2 buffers x 2 K-blocks x 16 KB
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sP + 4 * kTileBytes);
  uint64_t* kv_full = bar_q + 1;
  uint64_t* kv_empty = kv_full + kKvStages;
  uint64_t* s_full = kv_empty + kKvStages;
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = p_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, img = blockIdx.z;
  const int n_kv = p.n_kv_tiles;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmK);
    tc::tma_prefetch_desc(&tmV);
    tc::mbar_init(bar_q, 1);
    for (int s = 0; s < kKvStages; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&p_full[s], 128);
      tc::mbar_init(&o_full[s], 1);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;         // S[0] cols 0..127, S[1] cols 128..255
  const uint32_t tO = tmem_base + 256;   // O[0] cols 256..319, O[1] cols 320..383

  if (warp == 0) {
    if (lane == 0) {  // ---------------- TMA producer
      tc::mbar_expect_tx(bar_q, kTileBytes);
      tc::tma_load_2d(sQ, &tmQ, bar_q, head * kHeadDim, img * p.T_q + qt * kQTile);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        const uint32_t ph = (j / kKvStages) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        tc::mbar_expect_tx(&kv_full[s], 2 * kTileBytes);
        const int row = img * p.T_kv + j * kKvTile;
        tc::tma_load_2d(sK + s * kTileBytes, &tmK, &kv_full[s], head * kHeadDim, row);
        tc::tma_load_2d(sV + s * kTileBytes, &tmV, &kv_full[s], head * kHeadDim, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------- MMA issuer
      const uint32_t idesc_qk = tc::umma_idesc_f16(128, kKvTile, false, false);
      const uint32_t idesc_pv = tc::umma_idesc_f16(128, kHeadDim, false, true);  // B = V, MN-major
      const uint64_t dq = tc::umma_desc_sw128(tc::smem_u32(sQ));
      auto issue_pv = [&](int i) {
        const int b = i & 1, s = i % kKvStages;
        tc::mbar_wait(&p_full[b], (i >> 1) & 1);
        tc::tc_fence_after();
        const uint64_t dv = tc::umma_desc_sw128(tc::smem_u32(sV + s * kTileBytes));
#pragma unroll
        for (int kk = 0; kk < kKvTile / 16; ++kk) {
          // A = P[b]: two 64-key K-blocks of 16 KB, +32 B per 16 keys inside a block;  B = V: 16 keys = 16 rows = 2 KB
          const uint64_t dp = tc::umma_desc_sw128(tc::smem_u32(sP + (2 * b + (kk >> 2)) * kTileBytes)) + 2 * (kk & 3);
          tc::umma_f16_ss(tO + b * kHeadDim, dp, dv + kk * (2048 >> 4), idesc_pv, kk != 0);
        }
        tc::umma_commit(&o_full[b]);
        tc::umma_commit(&kv_empty[s]);
      };
      tc::mbar_wait(bar_q, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kKvStages;
        tc::mbar_wait(&kv_full[s], (j / kKvStages) & 1);
        tc::tc_fence_after();
        const uint64_t dk = tc::umma_desc_sw128(tc::smem_u32(sK + s * kTileBytes));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          tc::umma_f16_ss(tS + (j & 1) * kKvTile, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        tc::umma_commit(&s_full[j & 1]);
        if (j > 0) issue_pv(j - 1);
      }
      issue_pv(n_kv - 1);
    }
  } else {  // ---------------- softmax warps: thread = query row = TMEM lane
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const float sc = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float acc[kHeadDim];
#pragma unroll
    for (int i = 0; i < kHeadDim; ++i) acc[i] = 0.f;
    uint8_t* p_row = sP + (row >> 3) * 1024 + (row & 7) * 128;
    const int sw = row & 7;

    auto accumulate_o = [&](int i, float alpha) {
      const int b = i & 1;
      tc::mbar_wait(&o_full[b], (i >> 1) & 1);
      tc::tc_fence_after();
      uint32_t o[32];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        tc::tmem_ld32(tO + b * kHeadDim + c * 32 + lane_addr, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; ++q) acc[c * 32 + q] = fmaf(acc[c * 32 + q], alpha, __uint_as_float(o[q]));
      }
    };

    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const int kv_rem = p.T_kv - j * kKvTile;  // >= 1
      tc::mbar_wait(&s_full[b], (j >> 1) & 1);
      tc::tc_fence_after();
      const uint32_t ts = tS + b * kKvTile + lane_addr;
      uint32_t s[32];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::tmem_ld32(ts + c * 32, s);
        tc::tmem_ld_wait();
        if (kv_rem >= (c + 1) * 32) {
#pragma unroll
          for (int q = 0; q < 32; ++q) mx = fmaxf(mx, __uint_as_float(s[q]));
        } else {
#pragma unroll
          for (int q = 0; q < 32; ++q)
            if (c * 32 + q < kv_rem) mx = fmaxf(mx, __uint_as_float(s[q]));
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2((m_run - m_new) * sc);  // first tile: exp2(-inf) = 0
      const float neg_m = -m_new * sc;
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        tc::tmem_ld32(ts + c * 32, s);
        tc::tmem_ld_wait();
        uint8_t* blk = p_row + (2 * b + (c >> 1)) * kTileBytes;
#pragma unroll
        for (int g = 0; g < 4; ++g) {  // 8 keys = one 16-byte chunk
          float e[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = c * 32 + g * 8 + q;
            e[q] = col < kv_rem ? ex2(fmaf(__uint_as_float(s[g * 8 + q]), sc, neg_m)) : 0.f;
            psum += e[q];
          }
          __half2 h[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) h[q] = __floats2half2_rn(e[2 * q], e[2 * q + 1]);
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(blk + ((chunk ^ sw) << 4)) = *reinterpret_cast<uint4*>(h);
        }
      }
      l_run = fmaf(l_run, alpha, psum);
      m_run = m_new;
      tc::tc_fence_before();          // S[b] reads done before the issuer overwrites it (tile j+2)
      tc::fence_proxy_async_smem();   // P[b] visible to the tensor core
      tc::mbar_arrive(&p_full[b]);
      if (j > 0) accumulate_o(j - 1, alpha_prev);
      alpha_prev = alpha;
    }
    accumulate_o(n_kv - 1, alpha_prev);

    const int q_row = qt * kQTile + row;
    if (q_row < p.T_q) {
      const float inv = 1.f / l_run;
      __half2 h[kHeadDim / 2];
#pragma unroll
      for (int i = 0; i < kHeadDim / 2; ++i) h[i] = __floats2half2_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
      uint4* op = reinterpret_cast<uint4*>(p.out + ((long long)img * p.T_q + q_row) * p.ldo + head * kHeadDim);
#pragma unroll
      for (int i = 0; i < 8; ++i) op[i] = reinterpret_cast<uint4*>(h)[i];
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sgn

using namespace sgn;

extern "C" int sgn_attention_f16(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v,
                                 int64_t ldv, int B, int heads, int T_q, int T_kv, float scale, void* d_out,
                                 int64_t ldo, void* stream) {
  SGN_CHECK_ARG(B >= 0 && heads > 0 && T_q > 0 && T_kv > 0, "bad attention shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_q && d_k && d_v && d_out, "null pointer");
  SGN_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "row strides must be multiples of 8");
  SGN_CHECK_ARG(ldq >= heads * kHeadDim && ldk >= heads * kHeadDim && ldv >= heads * kHeadDim && ldo >= heads * kHeadDim,
                "row stride smaller than heads * 64");
  SGN_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_k) | reinterpret_cast<uintptr_t>(d_v) |
                  reinterpret_cast<uintptr_t>(d_out)) & 15) == 0, "operands must be 16-byte aligned");
  static bool attr_set = false;
  if (!attr_set) {
    SGN_CUDA(cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAttnSmem));
    attr_set = true;
  }
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[2] = {kHeadDim, 128};
  const uint64_t cols = (uint64_t)heads * kHeadDim;
  uint64_t dq[2] = {cols, (uint64_t)B * T_q}, sq[1] = {(uint64_t)ldq * 2};
  uint64_t dk[2] = {cols, (uint64_t)B * T_kv}, sk[1] = {(uint64_t)ldk * 2}, sv[1] = {(uint64_t)ldv * 2};
  int rc = encode_tmap(&tmQ, d_q, 2, dq, sq, box, nullptr);
  if (rc) return rc;
  rc = encode_tmap(&tmK, d_k, 2, dk, sk, box, nullptr);
  if (rc) return rc;
  rc = encode_tmap(&tmV, d_v, 2, dk, sv, box, nullptr);
  if (rc) return rc;
  AttnParams p;
  p.T_q = T_q, p.T_kv = T_kv;
  p.n_kv_tiles = (T_kv + kKvTile - 1) / kKvTile;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(d_out);
  p.ldo = ldo;
  dim3 grid((T_q + kQTile - 1) / kQTile, heads, B);
  k_attention_tc<<<grid, kAttnThreads, kAttnSmem, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}
