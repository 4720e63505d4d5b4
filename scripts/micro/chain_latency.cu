// Microbenchmark: what one link of a dependent-kernel chain costs on B200, by the way the dependency is expressed.
//   mode 0: plain stream order        mode 1: programmatic dependent launch (trigger at entry, griddepcontrol.wait before the data)
//   mode 2: PDL launch, but the data dependency is a release/acquire counter in global memory (no griddepcontrol.wait)
// Each kernel: `ctas` CTAs x 256 threads, every thread reads 16 B of the previous kernel's output, adds 1, writes 16 B.
// Chain of `len` kernels captured in a CUDA graph, 20 replays timed with events.   nvcc -arch=sm_100a -O3 chain_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <int MODE>
__global__ void __launch_bounds__(256) link(const uint4* in, uint4* out, const unsigned* wait_flag, unsigned* done_flag, unsigned target) {
  if (MODE == 1) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  }
  if (MODE == 2) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (wait_flag) {
      if (threadIdx.x == 0) { while (ld_acquire(wait_flag) < target) { } }
      __syncthreads();
    }
  }
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
  uint4 v = in[i];
  v.x += 1; v.y += 1; v.z += 1; v.w += 1;
  out[i] = v;
  if (MODE == 2) {
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence(); atomicAdd(done_flag, 1u); }
  }
}

template <int MODE>
float run(int ctas, int len, uint4* a, uint4* b, unsigned* flags, cudaStream_t st) {
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
  CK(cudaMemsetAsync(flags, 0, sizeof(unsigned) * (len + 1), st));
  for (int k = 0; k < len; ++k) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = (MODE != 0 && k > 0) ? 1 : 0;
    const uint4* in = (k & 1) ? b : a; uint4* out = (k & 1) ? a : b;
    const unsigned* wf = k > 0 ? flags + k - 1 : nullptr;
    CK(cudaLaunchKernelEx(&cfg, link<MODE>, in, out, wf, flags + k, (unsigned)ctas));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(e0, st));
  for (int i = 0; i < 20; ++i) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1000.f / 20.f / len;
}

int main() {
  const int len = 200;
  uint4 *a, *b; unsigned* flags;
  CK(cudaMalloc(&a, 1024 * 256 * 16)); CK(cudaMalloc(&b, 1024 * 256 * 16)); CK(cudaMalloc(&flags, sizeof(unsigned) * (len + 1)));
  CK(cudaMemset(a, 0, 1024 * 256 * 16)); CK(cudaMemset(b, 0, 1024 * 256 * 16));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  for (int ctas : {32, 128, 148, 296, 592}) {
    float t0 = run<0>(ctas, len, a, b, flags, st), t1 = run<1>(ctas, len, a, b, flags, st), t2 = run<2>(ctas, len, a, b, flags, st);
    unsigned h[4]; CK(cudaMemcpy(h, a, 16, cudaMemcpyDeviceToHost));
    printf("ctas %4d  us per link: stream order %.2f   PDL + griddepcontrol.wait %.2f   PDL + global counter %.2f   (check %u)\n", ctas, t0, t1, t2, h[0]);
  }
  return 0;
}
