// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) by operand layout, measured with
// clock64 around `reps` back-to-back MMAs + one commit on one SM (one CTA).  Cases:
//   K-major SWIZZLE_NONE planes [k/8][row][16 B] (slab_tc.cu's band layout): start address aligned to 128 B, shifted by
//   16 B (a dx tap), shifted by 34 * 16 B (a dy tap at pitch 34), shifted by 40 * 16 B (pitch 40: 128-B aligned rows);
//   K-major SWIZZLE_128B atoms (gconv_tc.cu's TMA layout);  N = 16 / 32 / 64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I mmvae_b200/csrc scripts/micro/umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "tc_common.cuh"
using namespace mmvae::tc;

__global__ void __launch_bounds__(128) rate(int layout, int N, int a_shift_chunks, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_s;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_s), 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_s;
  if (warp == 0) {
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t plane = 40 * 1024;                 // bytes per plane of the band layout
    uint64_t da, db;
    if (layout == 0) {                                // band layout: LBO = plane stride, SBO = 128
      da = make_smem_desc(base + 4096 + a_shift_chunks * 16, plane, 128, SWZ_NONE);
      db = make_smem_desc(base + 2 * plane, 1024, 128, SWZ_NONE);
    } else {                                          // SWIZZLE_128B, rows of 128 B, 8-row atoms of 1024 B
      da = make_smem_desc(base + a_shift_chunks * 1024, 16, 1024, SWZ_128);
      db = make_smem_desc(base + 64 * 1024, 16, 1024, SWZ_128);
    }
    const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
    __syncwarp();
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (elect_one()) mma_bf16(tmem, da, db, idesc, 1);
      __syncwarp();
    }
    if (elect_one()) mma_commit(smem_u32(&bar));
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 64); }
}

int main() {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 161 * 1024 + 1024);
  const int reps = 2000;
  struct Case { const char* name; int layout, shift; } cases[] = {
      {"band layout, aligned start", 0, 0}, {"band layout, +16 B (dx tap)", 0, 1}, {"band layout, +34 chunks (dy tap, pitch 34)", 0, 34},
      {"band layout, +35 chunks (dy+dx)", 0, 35}, {"band layout, +40 chunks (pitch 40)", 0, 40}, {"SWIZZLE_128B atoms", 1, 0}};
  for (auto& c : cases)
    for (int N : {16, 32, 64}) {
      rate<<<1, 128, 161 * 1024 + 1024>>>(c.layout, N, c.shift, reps, d);
      rate<<<1, 128, 161 * 1024 + 1024>>>(c.layout, N, c.shift, reps, d);
      long long h = 0; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      cudaError_t e = cudaGetLastError();
      printf("%-46s N=%2d  %.1f cycles per MMA%s\n", c.name, N, (double)h / reps, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  return 0;
}
