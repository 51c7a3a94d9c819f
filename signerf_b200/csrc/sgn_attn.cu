// K6: softmax(Q K^T / sqrt(d)) V for the SpatialTransformer blocks of the SDXL / ControlNet UNet (SURVEY §8a row A12:
// sgm CrossAttention, head_dim 64), both the self-attention over all sheet tokens (16 384 @640 ch, 4 096 @1280 ch at
// the 2048^2 sheet) and the cross-attention over the 77 prompt tokens, on tcgen05 with the score tiles in TMEM.
//
// One CTA = TWO 128-query tiles (A, B) of one (image, head) sharing every K/V tile.  Roles: warp 0 TMA producer,
// warp 1 tcgen05.mma issuer, warps 2-5 softmax of tile A, warps 6-9 softmax of tile B (thread i owns query row i =
// TMEM lane i).  The two softmax warpgroups ping-pong: while one exponentiates, the tensor core runs the other
// tile's P.V and next Q.K^T.  Per 128-key tile j and query tile X:
//     S_X  = Q_X K_j^T               4 x UMMA 128x128x16 into TMEM S_X
//     P_X  = exp2(c S_X - c m)       ONE TMEM read of the row (128 registers) -> max -> exp -> fp16 -> TMEM P_X
//     O_X += P_X V_j                 8 x UMMA 128x64x16, A = P_X from TMEM, B = V (MN-major smem), accumulating in TMEM
// Shared-memory bandwidth (128 B/clk/SM) is what the tensor core competes for at these small N: keeping P in tensor
// memory (tcgen05.st, A-from-TMEM MMA) halves the smem traffic per key tile (256 KB -> 128 KB per CTA).
// The row of S is pulled into registers in one go and S_X is handed straight back to the issuer (s_free), so
// Q_X K_{j+1}^T runs on the tensor core while the row is still being exponentiated.  The MUFU pipe is the next
// limit at head_dim 64 (16 ex2/clk/SM = 1024 cycles per 128x128 tile against 512 tensor cycles); the two warpgroups
// put two warps on every SM sub-partition so that one warp's FFMA / pack work fills the other's MUFU stalls.
// S is read once and O stays in TMEM: the
// online-softmax reference m is only advanced ("lazy rescale") when the row maximum grew by more than 2^8, in which
// case the warp multiplies its O rows by exp2(c (m_old - m_new)) through tcgen05.ld / tcgen05.st; otherwise P is taken
// against the stale m (P <= 2^8, harmless in fp16 / fp32 accumulation).  O is read once, at the end.
// Keys beyond T_kv (ragged last tile, 77-token context) are masked to -inf before the max.
#include <algorithm>

#include "sgn_common.cuh"
#include "sgn_tc.cuh"

namespace sgn {

constexpr int kHeadDim = 64;
constexpr int kQTile = 128;
constexpr int kQTileBytes = 128 * 64 * 2;  // one [128 x 64] fp16 tile, 128-B rows, SWIZZLE_128B

// Shape of one CTA: kGroups query tiles of 128 rows (one softmax warpgroup each) share every K / V tile of kKv keys.
//   <2, 128>: S 2x128 + O 2x64 + P 2x64 = 512 TMEM columns, two softmax warps per SM sub-partition taking MUFU turns;
//             S_x is released as soon as it is in registers, so Q_x K_{j+1}^T runs under the exponentials of tile j
//   <1, 128>: ONE softmax group per CTA and two CTAs per SM (256 TMEM columns, two K / V stages, 81 KB each): the two
//             query tiles of an SM no longer share K / V loads, barriers or the MMA-issuing thread (sgn_set_option
//             "attn_shape" 2; measured against the default in DESIGN.md)
//   <3, 96> : three free-running softmax warps per sub-partition.  S 3x96 + O 3x64 = 480 columns only fit with P_x
//             written over the first 48 columns of S_x (each thread overwrites its own row after reading it), so
//             Q_x K_{j+1}^T follows P_x V_j on the tensor pipe and a group idles through both - while the other two keep
//             the MUFU pipe busy.  (Three 64-key tiles without the aliasing were measured too: every tcgen05.mma costs
//             >= 60 cycles whatever its N - scratch/mma_bench - so 64-key Q K^T tiles make the tensor pipe the limit.)
template <int kGroups, int kKv>
struct AttnShape {
  static constexpr bool kAlias = kKv == 96;
  static constexpr bool kRegMove = kGroups == 3;   // setmaxnreg (see the kernel)
  static constexpr int kThreads = kRegMove ? 512 : (2 + 4 * kGroups) * 32;
  static constexpr int kQPerCta = kGroups * kQTile;
  static constexpr int kKvBytes = kKv * 64 * 2;
  static constexpr int kStages = kGroups == 1 ? 2 : 4;   // <1, 128>: two CTAs per SM, 81 KB each
  static constexpr int kCtasPerSm = kGroups == 1 ? 2 : 1;
  static constexpr size_t kSmem = 1024 + (size_t)kQTileBytes * kGroups + (size_t)2 * kStages * kKvBytes + 256;
  // TMEM column map
  static constexpr int kColS = 0, kColP = kAlias ? 0 : kGroups * kKv, kStrideP = kAlias ? kKv : kKv / 2;
  static constexpr int kColO = kAlias ? kGroups * kKv : kColP + kGroups * kKv / 2;
  static_assert(kColO + kGroups * kHeadDim <= 512, "tensor memory");
  static constexpr int kTmemCols = kColO + kGroups * kHeadDim <= 256 ? 256 : 512;   // allocation (a power of two)
};

struct AttnParams {
  int T_q, T_kv, n_kv_tiles;
  int n_qt, heads;   // work item -> (query tile group, head, image): item = qt + n_qt * (head + heads * image)
  int n_full, split; // CTAs [0, n_full) take a whole item; the others 1/split of the key tiles of one of the tail items
  int idle_ns;       // MMA-issuer back-off when no tile has work (0 = spin; polling issuer only)
  float scale_log2e;
  __half* out;
  long long ldo;
  float* part;       // [tail item][split][66][rows per CTA]: unnormalised O (64 columns), m, l of every query row
  unsigned* count;   // [tail item] arrivals (zero at launch; the combining CTA resets its counter)
};

// named barriers 1 / 2: the exp2 sections of the two softmax warpgroups of the <2, 128> shape take turns on the MUFU
// pipe, which staggers them by half a period: while one exponentiates, the other loads its next S row and finds the row
// maximum.  (ptxas moves about half of a tile's MUFU instructions above the bar.sync; pinning them behind it with data
// dependencies was measured slower - DESIGN.md.)
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2: two lanes of fp32 per issue slot) and the 3-input maximum (FMNMX3):
// the softmax warps are bound by issue slots next to the MUFU pipe, so every per-element instruction that can be halved is
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// make -C signerf_b200/csrc trace: clock64 time stamps of CTA 0 (lane 0 of the first warp of softmax groups 0 / 1 and
// the issuer), read back by scratch/attn_trace.py through sgn_debug_attn_trace
#ifndef SGN_ATTN_PROBE
#define SGN_ATTN_PROBE 0
#endif
#ifdef SGN_ATTN_TRACE
__device__ long long g_trace[32 * 256];   // [event + 16 * query tile][key tile]
#define TRACE(ev, j) do { if (blockIdx.x == 0 && (j) < 256 && (ev) < 32) g_trace[(ev) * 256 + (j)] = clock64(); } while (0)
#else
#define TRACE(ev, j) do { } while (0)
#endif
// sgn_set_option "attn_variant" (bit flags): 1 = the MMA issuer follows the static event order with blocking waits
// instead of polling, 2 = producer / issuer are the two highest warps of the CTA instead of the two lowest
int g_attn_variant = 3;
int g_attn_shape = 0;     // sgn_set_option "attn_shape": 0 = two query tiles x 128-key tiles (default), 1 = three x 96-key tiles with P aliased over S, 2 = one query tile per CTA, two CTAs per SM
int g_attn_idle_ns = 0;   // sgn_set_option "attn_idle_ns"
int g_attn_split = 1;     // sgn_set_option "attn_split": 0 = never split the tail items over the keys

__device__ __forceinline__ void decode_item(const AttnParams& p, int cta, int& item, int& part, int& kv0, int& n_kv) {
  item = cta, part = 0, kv0 = 0, n_kv = p.n_kv_tiles;
  if (cta >= p.n_full) {
    const int r = cta - p.n_full;
    item = p.n_full + r / p.split;
    part = r % p.split;
    kv0 = (int)((long long)part * p.n_kv_tiles / p.split);
    n_kv = (int)((long long)(part + 1) * p.n_kv_tiles / p.split) - kv0;
  }
}

// Registers are handed out per group of four warps: ten warps are charged as twelve (168 registers per thread), fourteen
// as sixteen (128).  The three-group shape therefore launches sixteen warps - softmax warps 0-11, producer 12, issuer 13,
// two idle - and moves registers with setmaxnreg: the last warpgroup keeps 56 per thread, the softmax warpgroups get 152.
template <int kFlags, int kGroups, int kKv>
__global__ void __launch_bounds__(AttnShape<kGroups, kKv>::kThreads, AttnShape<kGroups, kKv>::kCtasPerSm)
k_attention_tc(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  using Sh = AttnShape<kGroups, kKv>;
  constexpr bool kTurns = kGroups == 2;
  constexpr bool kAlias = Sh::kAlias;
  constexpr bool kStatic = (kFlags & 1) != 0;
  constexpr bool kHigh = (kFlags & 2) != 0 || Sh::kRegMove;
  constexpr int kWarpProd = kHigh ? 4 * kGroups : 0, kWarpIssue = kHigh ? 4 * kGroups + 1 : 1;
  constexpr int kStages = Sh::kStages, kKvBytes = Sh::kKvBytes;
  constexpr int kSoftThreads = 128 * kGroups;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                          // Q_0 .. Q_{kGroups-1}
  uint8_t* sK = sQ + kGroups * kQTileBytes;
  uint8_t* sV = sK + kStages * kKvBytes;
  uint64_t* bar_q = reinterpret_cast<uint64_t*>(sV + kStages * kKvBytes);
  uint64_t* kv_full = bar_q + 1;
  uint64_t* kv_empty = kv_full + kStages;
  uint64_t* s_full = kv_empty + kStages;       // [kGroups] per query tile
  uint64_t* p_full = s_full + kGroups;
  uint64_t* o_full = p_full + kGroups;         // P_x(j) V(j) complete (one phase per key tile)
  uint64_t* s_free = o_full + kGroups;         // S_x(j) is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + kGroups);
  uint32_t* last_flag = tmem_slot + 1;         // split items: this CTA is the last of its item to arrive

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work decode.  A launch is n_full whole items (one CTA each, all keys) followed by the tail items that would only
  // fill part of a last wave: those are cut into `split` key ranges (one CTA each) whose partial (O, m, l) are combined
  // by whichever CTA of the item finishes last.
  int item, part, kv0, n_kv;   // n_kv: key tiles of this CTA, counted from 0 by j below; kv0: its first key tile
  decode_item(p, blockIdx.x, item, part, kv0, n_kv);
  const int qt = item % p.n_qt, head = (item / p.n_qt) % p.heads, img = item / (p.n_qt * p.heads);

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmK);
    tc::tma_prefetch_desc(&tmV);
    tc::mbar_init(bar_q, 1);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&kv_full[s], 1);
      tc::mbar_init(&kv_empty[s], 1);
    }
    for (int s = 0; s < kGroups; ++s) {
      tc::mbar_init(&s_full[s], 1);
      tc::mbar_init(&p_full[s], 4);    // one arrival per softmax warp (128 per-thread arrivals serialise on the barrier word)
      tc::mbar_init(&o_full[s], 1);
      tc::mbar_init(&s_free[s], 4);
    }
    tc::mbar_fence_init();
  }
  if (warp == kWarpIssue) tc::tmem_alloc(tmem_slot, Sh::kTmemCols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base + Sh::kColS;   // S_x: kKv fp32 columns each
  const uint32_t tP = tmem_base + Sh::kColP;   // P_x: fp16 pairs, kKv / 2 columns each
  const uint32_t tO = tmem_base + Sh::kColO;   // O_x: 64 fp32 columns each

  auto role_producer = [&]() {
    if (lane == 0) {  // ---------------- TMA producer
      tc::mbar_expect_tx(bar_q, kGroups * kQTileBytes);
#pragma unroll
      for (int x = 0; x < kGroups; ++x)
        tc::tma_load_2d(sQ + x * kQTileBytes, &tmQ, bar_q, head * kHeadDim, img * p.T_q + qt * Sh::kQPerCta + x * kQTile);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j % kStages;
        const uint32_t ph = (j / kStages) & 1;
        tc::mbar_wait(&kv_empty[s], ph ^ 1);
        tc::mbar_expect_tx(&kv_full[s], 2 * kKvBytes);
        const int row = img * p.T_kv + (kv0 + j) * kKv;
        tc::tma_load_2d(sK + s * kKvBytes, &tmK, &kv_full[s], head * kHeadDim, row);
        tc::tma_load_2d(sV + s * kKvBytes, &tmV, &kv_full[s], head * kHeadDim, row);
      }
    }
  };
  auto role_issuer = [&]() {
    if (lane == 0) {  // ---------------- MMA issuer
      const uint32_t idesc_qk = tc::umma_idesc_f16(128, kKv, false, false);
      const uint32_t idesc_pv = tc::umma_idesc_f16(128, kHeadDim, false, true);  // B = V, MN-major
      auto issue_qk = [&](int x, int j) {
        const uint64_t dq = tc::umma_desc_sw128(tc::smem_u32(sQ + x * kQTileBytes));
        const uint64_t dk = tc::umma_desc_sw128(tc::smem_u32(sK + (j % kStages) * kKvBytes));
#pragma unroll
        for (int k = 0; k < kHeadDim / 16; ++k)
          tc::umma_f16_ss(tS + x * kKv, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
        tc::umma_commit(&s_full[x]);
      };
      auto issue_pv = [&](int x, int j) {
        const uint64_t dv = tc::umma_desc_sw128(tc::smem_u32(sV + (j % kStages) * kKvBytes));
#pragma unroll
        for (int kk = 0; kk < kKv / 16; ++kk) {
          // A = P_x in TMEM: 16 keys = 8 columns;  B = V: 16 keys = 16 rows = 2 KB
          tc::umma_f16_ts(tO + x * kHeadDim, tP + x * Sh::kStrideP + kk * 8, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0);
        }
        if (!kAlias) tc::umma_commit(&o_full[x]);
      };
      tc::mbar_wait(bar_q, 0);
      tc::mbar_wait(&kv_full[0], 0);
      tc::tc_fence_after();
#pragma unroll
      for (int x = 0; x < kGroups; ++x) issue_qk(x, 0);
      if constexpr (kAlias) {
        // P_x(j) lives in S_x: one event per group and key tile.  P_x(j) stored -> P_x V_j, then Q_x K_{j+1}^T into the
        // same columns (the tensor pipe runs them in issue order), one commit: S_x(j+1) complete implies O_x up to date.
        int pv_next[kGroups];
#pragma unroll
        for (int x = 0; x < kGroups; ++x) pv_next[x] = 0;
        int pv_done = 0;   // key tiles every group's P.V has been issued for
        while (pv_done < n_kv) {
          bool progress = false;
#pragma unroll
          for (int x = 0; x < kGroups; ++x) {
            const int jp = pv_next[x];
            if (jp < n_kv && tc::mbar_test(&p_full[x], jp & 1) &&
                (jp + 1 >= n_kv || tc::mbar_test(&kv_full[(jp + 1) % kStages], ((jp + 1) / kStages) & 1))) {
              tc::tc_fence_after();
              TRACE(9 + 16 * x, jp);
              issue_pv(x, jp);
              if (jp + 1 < n_kv) issue_qk(x, jp + 1);      // commits s_full[x]
              else tc::umma_commit(&o_full[x]);
              TRACE(10 + 16 * x, jp);
              pv_next[x] = jp + 1;
              int lo = pv_next[0];
#pragma unroll
              for (int y = 1; y < kGroups; ++y) lo = min(lo, pv_next[y]);
              if (lo > pv_done) {   // every group is through K/V(pv_done)
                tc::umma_commit(&kv_empty[pv_done % kStages]);
                pv_done = lo;
              }
              progress = true;
            }
          }
          if (!progress && p.idle_ns > 0) __nanosleep(p.idle_ns);
        }
      } else if constexpr (kStatic && kGroups == 3) {
        // Free-running groups move in lock step (they share every K / V tile and the MUFU pipe), so their events come in
        // bunches: wait for all of S(j) to be read, then issue the next Q K^T of ALL groups with the k-steps interleaved
        // (group 0 k0, group 1 k0, group 2 k0, group 0 k1, ...), likewise the P V.  The tensor pipe runs MMAs in issue
        // order and a k-step accumulating into the same tile waits for its predecessor to drain - at N = 64 (32 cycles of
        // math) that latency is most of the MMA's time; three independent accumulators back to back fill it.
        auto issue_qk_all = [&](int j) {
          const uint64_t dk = tc::umma_desc_sw128(tc::smem_u32(sK + (j % kStages) * kKvBytes));
#pragma unroll
          for (int k = 0; k < kHeadDim / 16; ++k) {
#pragma unroll
            for (int x = 0; x < kGroups; ++x) {
              const uint64_t dq = tc::umma_desc_sw128(tc::smem_u32(sQ + x * kQTileBytes));
              tc::umma_f16_ss(tS + x * kKv, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0);
            }
          }
#pragma unroll
          for (int x = 0; x < kGroups; ++x) tc::umma_commit(&s_full[x]);
        };
        auto issue_pv_all = [&](int j) {
          const uint64_t dv = tc::umma_desc_sw128(tc::smem_u32(sV + (j % kStages) * kKvBytes));
#pragma unroll
          for (int kk = 0; kk < kKv / 16; ++kk) {
#pragma unroll
            for (int x = 0; x < kGroups; ++x)
              tc::umma_f16_ts(tO + x * kHeadDim, tP + x * Sh::kStrideP + kk * 8, dv + kk * (2048 >> 4), idesc_pv, (j | kk) != 0);
          }
#pragma unroll
          for (int x = 0; x < kGroups; ++x) tc::umma_commit(&o_full[x]);
        };
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) {
            tc::mbar_wait(&kv_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
#pragma unroll
            for (int x = 0; x < kGroups; ++x) tc::mbar_wait(&s_free[x], j & 1);
            tc::tc_fence_after();
            TRACE(8, j + 1);
            TRACE(24, j + 1);
            issue_qk_all(j + 1);
          }
#pragma unroll
          for (int x = 0; x < kGroups; ++x) tc::mbar_wait(&p_full[x], j & 1);
          tc::tc_fence_after();
          TRACE(9, j);
          TRACE(25, j);
          issue_pv_all(j);
          tc::umma_commit(&kv_empty[j % kStages]);   // every tile is through K/V(j)
        }
      } else if constexpr (kStatic) {
        // The steady-state order of the events is fixed by the softmax groups' turns (S_1(j) read, P_0(j) stored,
        // S_0(j+1) read, P_1(j) stored, ...): wait for them in that order with blocking (hardware-suspended) waits instead
        // of polling all the barriers from a sub-partition that also hosts softmax warps.
        for (int j = 0; j < n_kv; ++j) {
          if (j + 1 < n_kv) {
            tc::mbar_wait(&kv_full[(j + 1) % kStages], ((j + 1) / kStages) & 1);
#pragma unroll
            for (int x = 0; x < kGroups; ++x) {
              tc::mbar_wait(&s_free[x], j & 1);
              tc::tc_fence_after();
              TRACE(8 + 16 * x, j + 1);
              issue_qk(x, j + 1);
            }
          }
#pragma unroll
          for (int x = 0; x < kGroups; ++x) {
            tc::mbar_wait(&p_full[x], j & 1);
            tc::tc_fence_after();
            TRACE(9 + 16 * x, j);
            issue_pv(x, j);
          }
          tc::umma_commit(&kv_empty[j % kStages]);   // every tile is through K/V(j)
        }
      } else {
        // Event loop: serve whichever of {S_x free -> next Q K^T, P_x ready -> P V} is ready, per query tile.
        int qk_next[kGroups], pv_next[kGroups];
#pragma unroll
        for (int x = 0; x < kGroups; ++x) qk_next[x] = 1, pv_next[x] = 0;
        int pv_done = 0;   // key tiles every group's P.V has been issued for
        while (pv_done < n_kv) {
          bool progress = false;
#pragma unroll
          for (int x = 0; x < kGroups; ++x) {
            const int jq = qk_next[x];
            if (jq < n_kv && tc::mbar_test(&s_free[x], (jq - 1) & 1) &&
                tc::mbar_test(&kv_full[jq % kStages], (jq / kStages) & 1)) {
              tc::tc_fence_after();
              TRACE(8 + 16 * x, jq);
              issue_qk(x, jq);
              qk_next[x] = jq + 1;
              progress = true;
            }
            const int jp = pv_next[x];
            if (jp < n_kv && tc::mbar_test(&p_full[x], jp & 1)) {
              tc::tc_fence_after();
              TRACE(9 + 16 * x, jp);
              issue_pv(x, jp);
              pv_next[x] = jp + 1;
              int lo = pv_next[0];
#pragma unroll
              for (int y = 1; y < kGroups; ++y) lo = min(lo, pv_next[y]);
              if (lo > pv_done) {   // every tile is through K/V(pv_done)
                tc::umma_commit(&kv_empty[pv_done % kStages]);
                pv_done = lo;
              }
              progress = true;
            }
          }
          if (!progress && p.idle_ns > 0) __nanosleep(p.idle_ns);
        }
      }
    }
  };
  // softmax warpgroups (four warps each); thread = query row = TMEM lane
  auto role_softmax = [&]() {
    const int sw = kHigh ? warp : warp - 2;
    const int x = sw >> 2;
    const int lane_base = (warp & 3) * 32;
    const int row = lane_base + lane;
    const uint32_t lane_addr = (uint32_t)lane_base << 16;
    const float sc = p.scale_log2e;
    float m_run = -INFINITY, l_run = 0.f;
    const uint32_t tp = tP + x * Sh::kStrideP + lane_addr;
    const uint32_t ts = tS + x * kKv + lane_addr;
    const uint32_t to = tO + x * kHeadDim + lane_addr;
    const bool tr = (sw & 3) == 0 && lane == 0 && x < 2;
    const int te = 16 * x;

    int kv_left = p.T_kv - kv0 * kKv;   // keys from this tile on
    if (kTurns && x == 1) named_arrive(1, 256);  // group 0 takes the first turn
    for (int j = 0; j < n_kv; ++j) {
      const int kv_rem = kv_left;  // >= 1
      kv_left -= kKv;
      if (tr) TRACE(te + 0, j);
      tc::mbar_wait(&s_full[x], j & 1);
      tc::tc_fence_after();
      if (tr) TRACE(te + 1, j);
      uint32_t s[kKv];
#if SGN_ATTN_PROBE & 32   // probe 32: only a quarter of the S row is read from tensor memory
      tc::tmem_ld32(ts, s);
#pragma unroll
      for (int c = 32; c < kKv; ++c) s[c] = s[c & 31] + c;
#else
#pragma unroll
      for (int c = 0; c < kKv / 32; ++c) tc::tmem_ld32(ts + c * 32, s + c * 32);
#endif
      tc::tmem_ld_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (!kAlias && lane == 0) tc::mbar_arrive(&s_free[x]);     // the issuer may overwrite S_x with Q_x K_{j+1}^T now
      if (tr) TRACE(te + 2, j);
      if (kv_rem < kKv) {
#pragma unroll
        for (int q = 0; q < kKv; ++q)
          if (q >= kv_rem) s[q] = 0xff800000u;  // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]);
      float mx2 = __uint_as_float(s[2]), mx3 = __uint_as_float(s[3]);
#if SGN_ATTN_PROBE & 4   // probe 4: the maximum of every 8th key only
#pragma unroll
      for (int q = 4; q < kKv - 4; q += 64) {
#else
#pragma unroll
      for (int q = 4; q < kKv - 4; q += 8) {
#endif
        mx0 = max3(mx0, __uint_as_float(s[q]), __uint_as_float(s[q + 1]));
        mx1 = max3(mx1, __uint_as_float(s[q + 2]), __uint_as_float(s[q + 3]));
        mx2 = max3(mx2, __uint_as_float(s[q + 4]), __uint_as_float(s[q + 5]));
        mx3 = max3(mx3, __uint_as_float(s[q + 6]), __uint_as_float(s[q + 7]));
      }
      mx0 = max3(mx0, __uint_as_float(s[kKv - 4]), __uint_as_float(s[kKv - 3]));
      mx1 = max3(mx1, __uint_as_float(s[kKv - 2]), __uint_as_float(s[kKv - 1]));
      const float mx = fmaxf(max3(mx0, mx2, mx3), mx1);
      // P_x(j-1) V(j-1) must be complete before P_x is rewritten or O_x rescaled; P_x(j) V(j) is not issued before
      // this thread arrives on p_full, so O_x is quiescent in between.  The wait is only needed before O is rescaled
      // (rare) or P is rewritten (the first tcgen05.st, 64 exponentials later), which takes the P.V latency off the
      // softmax critical path.
      bool o_ready = kAlias || j == 0;   // aliased: S_x(j) complete implies P_x(j-1) V(j-1) complete
      auto wait_o = [&]() {
        if (!o_ready) {   // warp-uniform
          if (tr) TRACE(te + 5, j);
          tc::mbar_wait(&o_full[x], (j - 1) & 1);
          if (tr) TRACE(te + 6, j);
          tc::tc_fence_after();
          o_ready = true;
        }
      };
      const bool grow = (mx - m_run) * sc > 8.f;   // also true on the first tile (m_run = -inf)
      const bool any_grow = __any_sync(0xffffffffu, grow);
      if (any_grow && j > 0) {
        wait_o();
        const float alpha = grow ? ex2((m_run - mx) * sc) : 1.f;
        uint32_t o[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          tc::tmem_ld16(to + c * 16, o);
          tc::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 16; ++q) o[q] = __float_as_uint(__uint_as_float(o[q]) * alpha);
          tc::tmem_st16(to + c * 16, o);
        }
        tc::tmem_st_wait();
        l_run *= alpha;
      }
      if (grow) m_run = mx;
      const float neg_m = -m_run * sc;
      if (tr) TRACE(te + 3, j);
      if (kTurns) named_sync(1 + x, 256);            // my turn on the MUFU pipe
      if (tr) TRACE(te + 4, j);
      // Packed math: scale + shift as FFMA2, row sums as FADD2, 2 packed accumulator chains; software-pipelined by one
      // 16-key block: the exponentials of block b+1 (in place over the S registers) are issued before block b is summed
      // and packed, so no MUFU result is consumed right behind its issue.
      const uint64_t sc2 = pk2(sc, sc), nm2 = pk2(neg_m, neg_m);
      uint64_t acc_a = pk2(0.f, 0.f), acc_b = pk2(0.f, 0.f);
      auto exp_block = [&](int b) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#if SGN_ATTN_PROBE & 1   // timing probe (wrong results): no scale / shift in front of the exponential
          float t0 = __uint_as_float(s[b * 16 + 2 * q]), t1 = __uint_as_float(s[b * 16 + 2 * q + 1]);
#else
          const uint64_t t = fma2(pk2u(s[b * 16 + 2 * q], s[b * 16 + 2 * q + 1]), sc2, nm2);
          float t0, t1;
          upk2(t, t0, t1);
#endif
          s[b * 16 + 2 * q] = __float_as_uint(ex2(t0));
          s[b * 16 + 2 * q + 1] = __float_as_uint(ex2(t1));
        }
      };
      exp_block(0);
      uint32_t pk[32];
#pragma unroll
      for (int b = 0; b < kKv / 16; ++b) {
        if (b + 1 < kKv / 16) exp_block(b + 1);
#if !(SGN_ATTN_PROBE & 2)   // probe 2: no row sum
#pragma unroll
        for (int q = 0; q < 8; q += 2) {
          acc_a = add2(acc_a, pk2u(s[b * 16 + 2 * q], s[b * 16 + 2 * q + 1]));
          acc_b = add2(acc_b, pk2u(s[b * 16 + 2 * q + 2], s[b * 16 + 2 * q + 3]));
        }
#endif
        const float* e = reinterpret_cast<const float*>(s) + b * 16;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
#if SGN_ATTN_PROBE & 16   // probe 16: no fp32 -> fp16 pack
          pk[(b & 3) * 8 + q] = s[b * 16 + 2 * q] ^ s[b * 16 + 2 * q + 1];
#else
          __half2 h = __floats2half2_rn(e[2 * q], e[2 * q + 1]);
          pk[(b & 3) * 8 + q] = *reinterpret_cast<uint32_t*>(&h);
#endif
        }
        if ((b & 3) == 3) {
          wait_o();
#if SGN_ATTN_PROBE & 8    // probe 8: P is not stored (one column keeps the registers alive)
          tc::tmem_st16(tp + (b >> 2) * 32, pk);
          asm volatile("" ::"r"(pk[16] ^ pk[17] ^ pk[18] ^ pk[19] ^ pk[20] ^ pk[21] ^ pk[22] ^ pk[23] ^ pk[24] ^ pk[25] ^ pk[26] ^ pk[27] ^ pk[28] ^ pk[29] ^ pk[30] ^ pk[31]));
#else
          tc::tmem_st32(tp + (b >> 2) * 32, pk);
#endif
        } else if (b == kKv / 16 - 1) {   // 96 keys: the last two blocks are 16 packed columns
          static_assert((kKv / 16) % 4 == 0 || (kKv / 16) % 4 == 2, "P store granularity");
          tc::tmem_st16(tp + (b >> 2) * 32, pk);
        }
      }
      if (kTurns) named_arrive(2 - x, 256);          // hand the turn to the other warpgroup
      float a0, a1, b0, b1;
      upk2(acc_a, a0, a1);
      upk2(acc_b, b0, b1);
      tc::tmem_st_wait();
      l_run += (a0 + a1) + (b0 + b1);
      tc::tc_fence_before();          // P_x written, O_x accesses done before the issuer touches them
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[x]);
      if (tr) TRACE(te + 7, j);
    }
    if (kTurns && x == 0) named_sync(1, 256);      // absorb group 1's last hand-over
    // O_x complete after the last P.V
    tc::mbar_wait(&o_full[x], kAlias ? 0 : (n_kv - 1) & 1);   // aliased: committed once, behind the last P.V
    tc::tc_fence_after();
    // the work decode again, from an opaque copy of the CTA index: nothing of it stays live across the key loop
    int cta;
    asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(cta));
    int item_e, part_e, kv0_e, n_kv_e;
    decode_item(p, cta, item_e, part_e, kv0_e, n_kv_e);
    const bool is_part = cta >= p.n_full;
    const int q_row = (item_e % p.n_qt) * Sh::kQPerCta + x * kQTile + row;
    __half* orow = p.out + ((long long)(item_e / (p.n_qt * p.heads)) * p.T_q + q_row) * p.ldo + ((item_e / p.n_qt) % p.heads) * kHeadDim;
    if (!is_part) {
      const float inv = 1.f / l_run;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(to + c * 32, o);
        tc::tmem_ld_wait();
        if (q_row < p.T_q) {
          __half2 h[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            h[i] = __floats2half2_rn(__uint_as_float(o[2 * i]) * inv, __uint_as_float(o[2 * i + 1]) * inv);
#pragma unroll
          for (int i = 0; i < 4; ++i) reinterpret_cast<uint4*>(orow + c * 32)[i] = reinterpret_cast<uint4*>(h)[i];
        }
      }
    } else {
      // Partial result of this key range: unnormalised O, m, l, column-major over the query rows of the item so that a
      // warp's stores coalesce.  The CTA of the item that arrives last folds the `split` partials together.
      constexpr int kRows = Sh::kQPerCta;
      const int r_cta = x * kQTile + row;
      float* base = p.part + (size_t)(item_e - p.n_full) * p.split * 66 * kRows;
      float* mine = base + (size_t)part_e * 66 * kRows + r_cta;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tc::tmem_ld32(to + c * 32, o);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) mine[(c * 32 + i) * kRows] = __uint_as_float(o[i]);
      }
      mine[64 * kRows] = m_run;
      mine[65 * kRows] = l_run;
      __threadfence();
      named_sync(3, kSoftThreads);
      if (sw == 0 && lane == 0) {
        const unsigned prev = atomicAdd(&p.count[item_e - p.n_full], 1u);
        const bool last = prev == (unsigned)p.split - 1;
        if (last) p.count[item_e - p.n_full] = 0;   // ready for the next launch
        *last_flag = last ? 1u : 0u;
      }
      named_sync(3, kSoftThreads);
      if (*last_flag) {
        __threadfence();
        float m_all = -INFINITY;
        for (int q = 0; q < p.split; ++q) m_all = fmaxf(m_all, __ldcg(base + (size_t)q * 66 * kRows + 64 * kRows + r_cta));
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        float l_all = 0.f;
        for (int q = 0; q < p.split; ++q) {
          const float* src = base + (size_t)q * 66 * kRows + r_cta;
          const float w = ex2((__ldcg(src + 64 * kRows) - m_all) * sc);
          l_all = fmaf(w, __ldcg(src + 65 * kRows), l_all);
#pragma unroll
          for (int i = 0; i < 64; ++i) acc[i] = fmaf(w, __ldcg(src + i * kRows), acc[i]);
        }
        if (q_row < p.T_q) {
          const float inv = 1.f / l_all;
          __half2 h[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) h[i] = __floats2half2_rn(acc[2 * i] * inv, acc[2 * i + 1] * inv);
#pragma unroll
          for (int i = 0; i < 8; ++i) reinterpret_cast<uint4*>(orow)[i] = reinterpret_cast<uint4*>(h)[i];
        }
      }
    }
  };
  if constexpr (Sh::kRegMove) {
    if (warp >= 4 * kGroups) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == kWarpProd) role_producer();
      else if (warp == kWarpIssue) role_issuer();
    } else {
      asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
      role_softmax();
    }
  } else {
    if (warp == kWarpProd) role_producer();
    else if (warp == kWarpIssue) role_issuer();
    else role_softmax();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == kWarpIssue) {
    __syncwarp();
    tc::tmem_dealloc(tmem_base, Sh::kTmemCols);
  }
}


// ------------------------------------------------------------------ short-context path (attn2: 77 prompt tokens)
// With <= 80 keys a (image, head) has 10 KB of K and V: the tcgen05 kernel above spends its time on per-CTA set-up
// (TMEM allocation, barrier init, one TMA round trip per tile) for a single key tile -- 36 us per launch at 8 192 x 1 280
// where the bytes (Q in, O out) are worth 7 us.  This path is plain FlashAttention-2 on mma.sync: K / V of the head in
// shared memory, 4 warps x 16 queries per tile, S and P in registers, several resident CTAs per SM to hide the loads.
constexpr int kXaKv = 80;      // keys (padded; masked beyond T_kv)
constexpr int kXaQ = 64;       // queries per tile (4 warps x 16)
constexpr int kXaPitch = 72;   // halfs per shared-memory row: 144 B keeps ldmatrix conflict-free

__device__ __forceinline__ void xa_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void xa_ldsm_x4(uint32_t (&r)[4], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(tc::smem_u32(p)));
}
__device__ __forceinline__ void xa_ldsm_x2(uint32_t (&r)[2], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(tc::smem_u32(p)));
}
__device__ __forceinline__ void xa_ldsm_x2_trans(uint32_t (&r)[2], const __half* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(tc::smem_u32(p)));
}
__device__ __forceinline__ uint32_t xa_pack(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(128)
k_cross_attention(const __half* __restrict__ q, long long ldq, const __half* __restrict__ k, long long ldk,
                  const __half* __restrict__ v, long long ldv, int T_q, int T_kv, float scale_log2e,
                  __half* __restrict__ out, long long ldo, int q_per_cta, int causal) {
  __shared__ __align__(16) __half sK[kXaKv * kXaPitch];
  __shared__ __align__(16) __half sV[kXaKv * kXaPitch];
  __shared__ __align__(16) __half sQ2[2][kXaQ * kXaPitch];   // double-buffered query tiles (cp.async)
  const int head = blockIdx.y, img = blockIdx.z;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_begin = blockIdx.x * q_per_cta;
  const int q_end = min(T_q, q_begin + q_per_cta);
  const int g = lane >> 2, col0 = (lane & 3) * 2;
  // 16-byte cp.async copies of one 64 x 64 query tile; rows past T_q are zero-filled (src-size 0)
  auto load_q = [&](int q0, __half* dst) {
    for (int c = tid; c < kXaQ * 8; c += 128) {
      const int r = c >> 3, j = c & 7;
      const bool ok = q0 + r < T_q;
      const __half* src = q + ((long long)img * T_q + (ok ? q0 + r : 0)) * ldq + head * kHeadDim + j * 8;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tc::smem_u32(dst + r * kXaPitch + j * 8)), "l"(src),
                   "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  load_q(q_begin, sQ2[0]);   // the first query tile is in flight while K / V of the head are staged
  {
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    for (int c = tid; c < kXaKv * 8; c += 128) {
      const int r = c >> 3, j = c & 7;
      uint4 kv = zero, vv = zero;
      if (r < T_kv) {
        kv = *reinterpret_cast<const uint4*>(k + ((long long)img * T_kv + r) * ldk + head * kHeadDim + j * 8);
        vv = *reinterpret_cast<const uint4*>(v + ((long long)img * T_kv + r) * ldv + head * kHeadDim + j * 8);
      }
      *reinterpret_cast<uint4*>(sK + r * kXaPitch + j * 8) = kv;
      *reinterpret_cast<uint4*>(sV + r * kXaPitch + j * 8) = vv;
    }
  }
  int buf = 0;
  for (int q0 = q_begin; q0 < q_end; q0 += kXaQ, buf ^= 1) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // this tile (and, first pass, K / V) visible to all; the other buffer's readers are done
    if (q0 + kXaQ < q_end) load_q(q0 + kXaQ, sQ2[buf ^ 1]);   // next tile in flight while this one is computed
    const __half* sQ = sQ2[buf];
    float s[kXaKv / 8][4];
#pragma unroll
    for (int nt = 0; nt < kXaKv / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kHeadDim / 16; ++kk) {
      uint32_t a[4];
      xa_ldsm_x4(a, sQ + (warp * 16 + (lane & 15)) * kXaPitch + kk * 16 + (lane >> 4) * 8);
#pragma unroll
      for (int nt = 0; nt < kXaKv / 8; ++nt) {
        uint32_t b[2];
        xa_ldsm_x2(b, sK + (nt * 8 + (lane & 7)) * kXaPitch + kk * 16 + ((lane >> 3) & 1) * 8);
        xa_mma(s[nt], a, b);
      }
    }
    // thread owns rows g and g + 8 of the warp's 16, columns nt*8 + col0 + {0,1}
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < kXaKv / 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = nt * 8 + col0 + (e & 1);
        // causal (CLIP text transformers): query i attends to keys <= i; rows g / g + 8 of the warp's 16
        const int qrow = q0 + warp * 16 + g + ((e & 2) ? 8 : 0);
        if (key >= T_kv || (causal && key > qrow)) s[nt][e] = -INFINITY;
      }
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float nm0 = -m0 * scale_log2e, nm1 = -m1 * scale_log2e;
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < kXaKv / 8; ++nt) {
      s[nt][0] = ex2(fmaf(s[nt][0], scale_log2e, nm0));
      s[nt][1] = ex2(fmaf(s[nt][1], scale_log2e, nm0));
      s[nt][2] = ex2(fmaf(s[nt][2], scale_log2e, nm1));
      s[nt][3] = ex2(fmaf(s[nt][3], scale_log2e, nm1));
      l0 += s[nt][0] + s[nt][1];
      l1 += s[nt][2] + s[nt][3];
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    float o[kHeadDim / 8][4];
#pragma unroll
    for (int nt = 0; nt < kHeadDim / 8; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < kXaKv / 16; ++kk) {   // P (registers) becomes the A operand: the C layout of two key tiles
      uint32_t a[4] = {xa_pack(s[2 * kk][0], s[2 * kk][1]), xa_pack(s[2 * kk][2], s[2 * kk][3]),
                       xa_pack(s[2 * kk + 1][0], s[2 * kk + 1][1]), xa_pack(s[2 * kk + 1][2], s[2 * kk + 1][3])};
#pragma unroll
      for (int nt = 0; nt < kHeadDim / 8; ++nt) {
        uint32_t b[2];
        xa_ldsm_x2_trans(b, sV + (kk * 16 + (lane & 15)) * kXaPitch + nt * 8);
        xa_mma(o[nt], a, b);
      }
    }
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    __half* o0 = out + ((long long)img * T_q + r0) * ldo + head * kHeadDim + col0;
    __half* o1 = out + ((long long)img * T_q + r1) * ldo + head * kHeadDim + col0;
#pragma unroll
    for (int nt = 0; nt < kHeadDim / 8; ++nt) {
      if (r0 < T_q) *reinterpret_cast<uint32_t*>(o0 + nt * 8) = xa_pack(o[nt][0] * inv0, o[nt][1] * inv0);
      if (r1 < T_q) *reinterpret_cast<uint32_t*>(o1 + nt * 8) = xa_pack(o[nt][2] * inv1, o[nt][3] * inv1);
    }
  }
}

int g_attn_short_kv = 1;   // sgn_set_option "attn_short_kv": 0 sends T_kv <= 80 through the tcgen05 kernel as well

}  // namespace sgn

using namespace sgn;

#ifdef SGN_ATTN_TRACE
extern "C" int sgn_debug_attn_trace(long long* h_out) {
  return cudaMemcpyFromSymbol(h_out, g_trace, sizeof(long long) * 32 * 256) == cudaSuccess ? 0 : -1;
}
#endif

// How a launch of `items` (query pair tile, head, image) work items with n_kv key tiles each is laid over n_sm SMs (one
// CTA per SM at a time): whole waves run one CTA per item; the items of an under-filled last wave are cut into `split`
// key ranges when that shortens the tail (cost in key-tile periods, + 2 per CTA for its prologue and the partial
// store / combine).  T = 4 096 with 20 heads x 2 images on 148 SMs: 640 items = 4.32 waves -> 592 whole items + 48 x 3
// parts of 10-11 key tiles, 128 + 13 instead of 160 periods.
static void plan_split(int items, int n_kv, int n_sm, int* n_full, int* split, int min_tiles = 4) {
  *n_full = items;
  *split = 1;
  const int rem = items % n_sm;
  if (rem == 0 || !g_attn_split) return;
  int best = 1;
  double best_cost = n_kv + 2;
  for (int s = 2; s <= 8 && n_kv / s >= min_tiles; ++s) {
    const double cost = (double)((rem * s + n_sm - 1) / n_sm) * ((n_kv + s - 1) / s + 2);
    if (cost < 0.95 * best_cost) best = s, best_cost = cost;
  }
  if (best > 1) *n_full = items - rem, *split = best;
}

static size_t split_workspace_bytes(int tail_items, int split, int rows_per_cta) {
  return (size_t)tail_items * split * 66 * rows_per_cta * sizeof(float) + (size_t)tail_items * sizeof(unsigned);
}
// rows per CTA / keys per tile of the configured shape (sgn_set_option "attn_shape")
static void shape_dims(int* rows_per_cta, int* kv_tile) {
  *rows_per_cta = g_attn_shape == 1 ? AttnShape<3, 96>::kQPerCta : g_attn_shape == 2 ? AttnShape<1, 128>::kQPerCta : AttnShape<2, 128>::kQPerCta;
  *kv_tile = g_attn_shape == 1 ? 96 : 128;
}

static int attention_impl(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v, int64_t ldv, int B,
                          int heads, int T_q, int T_kv, float scale, void* d_out, int64_t ldo, int causal, void* d_ws,
                          int64_t ws_bytes, void* stream) {
  SGN_CHECK_ARG(B >= 0 && heads > 0 && T_q > 0 && T_kv > 0, "bad attention shape");
  if (B == 0) return SGN_OK;
  SGN_CHECK_ARG(d_q && d_k && d_v && d_out, "null pointer");
  SGN_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "row strides must be multiples of 8");
  SGN_CHECK_ARG(ldq >= heads * kHeadDim && ldk >= heads * kHeadDim && ldv >= heads * kHeadDim && ldo >= heads * kHeadDim,
                "row stride smaller than heads * 64");
  SGN_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_q) | reinterpret_cast<uintptr_t>(d_k) | reinterpret_cast<uintptr_t>(d_v) |
                  reinterpret_cast<uintptr_t>(d_out)) & 15) == 0, "operands must be 16-byte aligned");
  SGN_CHECK_ARG(!causal || (T_kv <= kXaKv && T_q == T_kv), "causal attention is built for self-attention over <= 80 tokens");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (T_kv <= kXaKv && (g_attn_short_kv || causal)) {
    // ~4 resident waves: enough query tiles per CTA to amortise its K / V load, enough CTAs to fill the GPU
    const int tiles = (T_q + kXaQ - 1) / kXaQ;
    const long long want = 4ll * sm_count();
    int tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(8, (long long)tiles * heads * B / want));
    dim3 grid((tiles + tiles_per_cta - 1) / tiles_per_cta, heads, B);
    k_cross_attention<<<grid, 128, 0, st>>>(
        reinterpret_cast<const __half*>(d_q), ldq, reinterpret_cast<const __half*>(d_k), ldk,
        reinterpret_cast<const __half*>(d_v), ldv, T_q, T_kv, scale * 1.4426950408889634f,
        reinterpret_cast<__half*>(d_out), ldo, tiles_per_cta * kXaQ, causal);
    SGN_LAUNCH_CHECK();
    return SGN_OK;
  }
  using Kern = void (*)(CUtensorMap, CUtensorMap, CUtensorMap, AttnParams);
  static const Kern kerns[3][4] = {
      {k_attention_tc<0, 2, 128>, k_attention_tc<1, 2, 128>, k_attention_tc<2, 2, 128>, k_attention_tc<3, 2, 128>},
      {k_attention_tc<0, 3, 96>, k_attention_tc<1, 3, 96>, k_attention_tc<2, 3, 96>, k_attention_tc<3, 3, 96>},
      {k_attention_tc<0, 1, 128>, k_attention_tc<1, 1, 128>, k_attention_tc<2, 1, 128>, k_attention_tc<3, 1, 128>}};
  static const size_t smem_of[3] = {AttnShape<2, 128>::kSmem, AttnShape<3, 96>::kSmem, AttnShape<1, 128>::kSmem};
  static const int threads_of[3] = {AttnShape<2, 128>::kThreads, AttnShape<3, 96>::kThreads, AttnShape<1, 128>::kThreads};
  static bool attr_set = false;
  if (!attr_set) {
    for (int sh = 0; sh < 3; ++sh)
      for (Kern k : kerns[sh]) SGN_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_of[sh]));
    attr_set = true;
  }
  const int shape = g_attn_shape == 1 ? 1 : g_attn_shape == 2 ? 2 : 0;
  int rows_per_cta, kv_tile;
  shape_dims(&rows_per_cta, &kv_tile);
  CUtensorMap tmQ, tmK, tmV;
  const uint32_t box[2] = {kHeadDim, 128}, box_kv[2] = {kHeadDim, (uint32_t)kv_tile};
  const uint64_t cols = (uint64_t)heads * kHeadDim;
  uint64_t dq[2] = {cols, (uint64_t)B * T_q}, sq[1] = {(uint64_t)ldq * 2};
  uint64_t dk[2] = {cols, (uint64_t)B * T_kv}, sk[1] = {(uint64_t)ldk * 2}, sv[1] = {(uint64_t)ldv * 2};
  int rc = encode_tmap(&tmQ, d_q, 2, dq, sq, box, nullptr);
  if (rc) return rc;
  rc = encode_tmap(&tmK, d_k, 2, dk, sk, box_kv, nullptr);
  if (rc) return rc;
  rc = encode_tmap(&tmV, d_v, 2, dk, sv, box_kv, nullptr);
  if (rc) return rc;
  AttnParams p;
  p.T_q = T_q, p.T_kv = T_kv;
  p.n_kv_tiles = (T_kv + kv_tile - 1) / kv_tile;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__half*>(d_out);
  p.ldo = ldo;
  p.n_qt = (T_q + rows_per_cta - 1) / rows_per_cta;
  p.heads = heads;
  const int items = p.n_qt * heads * B;
  p.n_full = items, p.split = 1;
  p.part = nullptr, p.count = nullptr;
  if (d_ws) {   // key-range split of the tail items needs the caller's workspace
    plan_split(items, p.n_kv_tiles, sm_count() * (g_attn_shape == 2 ? 2 : 1), &p.n_full, &p.split);
    if (p.split > 1) {
      const int tail = items - p.n_full;
      const size_t need = split_workspace_bytes(tail, p.split, rows_per_cta);
      SGN_CHECK_ARG((reinterpret_cast<uintptr_t>(d_ws) & 15) == 0 && ws_bytes >= (int64_t)need,
                    "attention workspace too small or misaligned (sgn_attention_workspace_bytes)");
      p.part = reinterpret_cast<float*>(d_ws);
      p.count = reinterpret_cast<unsigned*>(p.part + (size_t)tail * p.split * 66 * rows_per_cta);
      SGN_CUDA(cudaMemsetAsync(p.count, 0, (size_t)tail * sizeof(unsigned), st));
    }
  }
  p.idle_ns = g_attn_idle_ns;
  Kern kern = kerns[shape][g_attn_variant & 3];
  kern<<<p.n_full + (items - p.n_full) * p.split, threads_of[shape], smem_of[shape], st>>>(tmQ, tmK, tmV, p);
  SGN_LAUNCH_CHECK();
  return SGN_OK;
}

extern "C" int64_t sgn_attention_workspace_bytes(int B, int heads, int T_q, int T_kv) {
  if (B <= 0 || heads <= 0 || T_q <= 0 || T_kv <= 0 || (T_kv <= kXaKv && g_attn_short_kv)) return 0;
  int dev_ok = 0;
  if (cudaGetDeviceCount(&dev_ok) != cudaSuccess || dev_ok == 0) return 0;
  int rows_per_cta, kv_tile;
  shape_dims(&rows_per_cta, &kv_tile);
  const int n_qt = (T_q + rows_per_cta - 1) / rows_per_cta, items = n_qt * heads * B;
  int n_full, split;
  plan_split(items, (T_kv + kv_tile - 1) / kv_tile, sm_count() * (g_attn_shape == 2 ? 2 : 1), &n_full, &split);
  return split > 1 ? (int64_t)split_workspace_bytes(items - n_full, split, rows_per_cta) : 0;
}

extern "C" int sgn_attention_f16_ws(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v,
                                    int64_t ldv, int B, int heads, int T_q, int T_kv, float scale, void* d_out,
                                    int64_t ldo, void* d_ws, int64_t ws_bytes, void* stream) {
  return attention_impl(d_q, ldq, d_k, ldk, d_v, ldv, B, heads, T_q, T_kv, scale, d_out, ldo, 0, d_ws, ws_bytes, stream);
}

extern "C" int sgn_attention_f16(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v,
                                 int64_t ldv, int B, int heads, int T_q, int T_kv, float scale, void* d_out,
                                 int64_t ldo, void* stream) {
  return attention_impl(d_q, ldq, d_k, ldk, d_v, ldv, B, heads, T_q, T_kv, scale, d_out, ldo, 0, nullptr, 0, stream);
}

extern "C" int sgn_attention_causal_f16(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v,
                                        int64_t ldv, int B, int heads, int T, float scale, void* d_out, int64_t ldo,
                                        void* stream) {
  return attention_impl(d_q, ldq, d_k, ldk, d_v, ldv, B, heads, T, T, scale, d_out, ldo, 1, nullptr, 0, stream);
}
