// Micro-benchmark: cycles per tcgen05.mma (cta_group::1, kind::f16, M = 128, K = 16) as issued by one thread, for the
// shapes the attention kernel uses: A from shared memory (Q K^T) or tensor memory (P V), N = 64 / 128 / 256, the k-steps
// accumulating into one tile (dependent) or round-robin over 2 / 3 tiles.  nvcc -arch=sm_100a -I../../signerf_b200/csrc
#include <cstdio>
#include "sgn_tc.cuh"
using namespace sgn;

template <int N, int TS, int ACCS, int BMN>
__global__ void __launch_bounds__(128, 1) k_bench(int reps, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // [128 x 64] fp16
  uint8_t* sB = smem + 16384;      // [256 x 64] fp16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
  if (threadIdx.x < 32) tc::tmem_alloc(slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = tc::umma_idesc_f16(128, N, false, BMN != 0);
    const uint64_t da = tc::umma_desc_sw128(tc::smem_u32(sA)), db = tc::umma_desc_sw128(tc::smem_u32(sB));
    for (int i = 0; i < 8; ++i) tc::umma_f16_ss(tm, da, db, idesc, 1);
    tc::umma_commit(bar);
    tc::mbar_wait(bar, 0);
    tc::tc_fence_after();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < 12; ++i) {
        const uint32_t d = tm + (i % ACCS) * N;          // accumulators side by side
        const int k = (i / ACCS) & 3;
        if (TS) tc::umma_f16_ts(d, tm + 448 + k * 8, db + (BMN ? k * (2048 >> 4) : 2 * k), idesc, 1);
        else tc::umma_f16_ss(d, da + 2 * k, db + (BMN ? k * (2048 >> 4) : 2 * k), idesc, 1);
      }
    }
    const long long t1 = clock64();
    tc::umma_commit(bar);
    tc::mbar_wait(bar, 1);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(tm, 512);
}

template <int N, int TS, int ACCS, int BMN>
void run(long long* d) {
  cudaFuncSetAttribute(k_bench<N, TS, ACCS, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60000);
  const int reps = 32;
  for (int grid : {1, 148}) {
    k_bench<N, TS, ACCS, BMN><<<grid, 128, 60000>>>(reps, d);
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    printf("N %3d %s B %s accs %d grid %3d: issue %.1f clk/MMA, complete %.1f clk/MMA (ideal math %d)%s\n", N, TS ? "A=TMEM" : "A=smem",
           BMN ? "MN" : "K ", ACCS, grid, (double)h[0] / (reps * 12), (double)h[1] / (reps * 12), N / 2, e ? cudaGetErrorString(e) : "");
  }
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  run<256, 0, 1, 0>(d); run<128, 0, 1, 0>(d); run<128, 0, 2, 0>(d); run<64, 0, 1, 0>(d); run<64, 0, 2, 0>(d); run<64, 0, 3, 0>(d);
  run<64, 1, 1, 1>(d); run<64, 1, 2, 1>(d); run<64, 1, 3, 1>(d); run<128, 1, 1, 1>(d); run<128, 1, 2, 1>(d);
  run<64, 0, 1, 1>(d); run<64, 0, 3, 1>(d); run<32, 0, 1, 0>(d); run<32, 1, 1, 1>(d); run<16, 1, 1, 1>(d);
  return 0;
}
