// Shared host/device helpers for the signerf_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <string>

#include "../../include/signerf_b200.h"

namespace sgn {

void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launches;
int sm_count();  // SMs of the current device (sgn_render.cu)
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define SGN_CHECK_ARG(cond, msg)                                   \
  do {                                                             \
    if (!(cond)) {                                                 \
      sgn::set_error(std::string("invalid argument: ") + (msg));   \
      return SGN_ERR_INVALID_ARG;                                  \
    }                                                              \
  } while (0)

#define SGN_CUDA(call)                                                                        \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      sgn::set_error(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " (" +     \
                     __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
      return SGN_ERR_CUDA;                                                                    \
    }                                                                                         \
  } while (0)

#define SGN_LAUNCH_CHECK()                 \
  do {                                     \
    sgn::count_launch();                   \
    SGN_CUDA(cudaPeekAtLastError());       \
  } while (0)

// Stream-ordered scratch memory.  The device's default memory pool releases freed blocks back to the driver at every
// synchronisation point unless told otherwise (release threshold 0): a caller that synchronises between two entry points
// (any host read-back does) then pays a fresh driver allocation per call - 10-15 ms for the small buffers, hundreds of ms
// for the sampling cascade's gigabytes.  The first allocation on a device lifts the threshold, so scratch is recycled.
inline cudaError_t scratch_alloc(void** p, size_t bytes, cudaStream_t st) {
  static std::atomic<uint32_t> pool_ready{0};   // bit d: device d's pool keeps its memory
  int dev = 0;
  cudaGetDevice(&dev);
  const uint32_t bit = 1u << (dev & 31);
  if (!(pool_ready.load(std::memory_order_relaxed) & bit)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_ready.fetch_or(bit, std::memory_order_relaxed);
  }
  return cudaMallocAsync(p, bytes, st);
}
template <class T>
inline cudaError_t scratch_alloc(T** p, size_t bytes, cudaStream_t st) {
  return scratch_alloc(reinterpret_cast<void**>(p), bytes, st);
}

constexpr int kMaxLevels = 16;
constexpr uint32_t kPrimeY = 2654435761u;  // HashEncoding.hash_fn primes (x prime is 1)
constexpr uint32_t kPrimeZ = 805459861u;

// Device-side view of one hash grid.
struct GridDev {
  const float2* table;  // [L * T]
  float res[kMaxLevels];
  int num_levels;
  uint32_t mask;  // T - 1
  uint32_t size;  // T
};

// Main-field MLP sizes the tensor-core path is specialised for (nerfacto defaults).
constexpr int kFeat = 32;     // 16 levels x 2
constexpr int kHidden = 64;
constexpr int kBaseOut = 16;  // density logit + 15 geo features
constexpr int kHeadIn = 32;   // [SH16 | logit-slot(0) geo15]; appearance folded into the bias
constexpr int kRgbPad = 8;

// Packed parameter block copied to shared memory with one cp.async.bulk per CTA.
// fp16 B-fragments are stored fragment-major: [ktile][ntile-pair][lane] x uint4.
struct alignas(16) MlpPack {
  uint4 w_base0[2 * 4 * 32];  // K=32 (2 ktiles) x N=64 (4 ntile pairs)
  uint4 w_base1[4 * 1 * 32];  // K=64 x N=16
  uint4 w_head0[2 * 4 * 32];  // K=32 x N=64
  uint4 w_head1[4 * 4 * 32];  // K=64 x N=64
  uint4 w_head2[2 * 32];      // K=64 x N=8, packed by ktile pairs
  float b_base0[64];
  float b_base1[16];
  float b_head0[64];  // head bias + W_head0[:, app] . appearance_mean
  float b_head1[64];
  float b_head2[8];
  float feat_scale;      // power of two applied to hash features before fp16 conversion
  float inv_feat_scale;  // applied to the fp32 accumulator of base layer 0
  float avg_density;
  float pad_;
};

// fp32 weights for the CUDA-core parity path ([out][in] row-major like nn.Linear).
struct alignas(16) MlpF32 {
  float w_base0[64 * 32];
  float w_base1[16 * 64];
  float w_head0[64 * 32];  // columns: SH16 | slot0 (zero) geo15
  float w_head1[64 * 64];
  float w_head2[3 * 64];
  float b_base0[64];
  float b_base1[16];
  float b_head0[64];
  float b_head1[64];
  float b_head2[4];
  float avg_density;
  float pad_[3];
};

// Proposal network: 5-level grid + 10->16->1 MLP, fp32 on CUDA cores.
struct PropDev {
  GridDev grid;
  float w0[16 * 10];
  float b0[16];
  float w1[16];
  float b1;
  float avg_density;
};

}  // namespace sgn

struct SgnField {
  sgn::GridDev grid;
  sgn::MlpPack* d_pack = nullptr;
  sgn::MlpF32* d_f32 = nullptr;
  sgn::PropDev* d_prop[2] = {nullptr, nullptr};
  sgn::PropDev h_prop[2];
  int num_proposals = 0;
  float avg_density = 1.f;
  int device = 0;
};
