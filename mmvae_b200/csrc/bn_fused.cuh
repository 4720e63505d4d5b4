// bn_fused.cuh -- training-mode BatchNorm statistics folded into the kernel that produces the tensor
// (bf16 mode): every CTA adds its per-channel (sum, sum of squares) into fp64 accumulators with
// atomicAdd(double); the CTA that finishes last turns them into (mean, rstd) / (scale, shift) and updates
// the running buffers (nn.BatchNorm2d defaults: momentum 0.1, unbiased running variance,
// num_batches_tracked += 1 -- reference model.py:30,34,61,66,95,137,162,173,202).  No separate finalize
// launch, no per-tile partial rows.  fp64 accumulation makes the result independent of the arrival
// order to well below fp32 resolution.
#pragma once
#include "kernels.cuh"

namespace mmvae {

__device__ __forceinline__ double ld_cg_f64(const double* p) {
  double v;
  asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// Accumulator copy of this CTA: same-address atomics serialise in L2 (~30 cycles each), so the CTAs are
// spread over kBnAccCopies copies of the [2][C] accumulators; the last CTA adds the copies up.
__device__ __forceinline__ double* bn_acc_copy(const BnFused& b) {
  return b.acc + (size_t)(blockIdx.x % kBnAccCopies) * 2 * b.C;
}

// Called by ALL threads of the CTA (contains __syncthreads) after their atomicAdd contributions.
__device__ __forceinline__ void bn_fused_finish(const BnFused& b, unsigned int nctas) {
  __shared__ int bn_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) bn_is_last = (atomicAdd(b.counter, 1u) == nctas - 1u) ? 1 : 0;
  __syncthreads();
  if (!bn_is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
    double s = 0.0, q = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {           // fixed order over the copies
      s += ld_cg_f64(b.acc + (size_t)k * 2 * b.C + c);
      q += ld_cg_f64(b.acc + (size_t)k * 2 * b.C + b.C + c);
    }
    const double mean = s * b.inv_m;
    const double var = fmax(q * b.inv_m - mean * mean, 0.0);
    const float meanf = (float)mean, varf = (float)var;
    if (b.running_mean) {
      b.running_mean[c] = 0.9f * b.running_mean[c] + 0.1f * meanf;
      b.running_var[c] = 0.9f * b.running_var[c] + 0.1f * (float)(var * b.unbias);
      if (c == 0 && b.nbt) *b.nbt += 1;
    }
    const float rstd = 1.0f / sqrtf(varf + 1e-5f);
    b.stat[c] = meanf; b.stat[b.C + c] = rstd;
    const float scale = b.gamma[c] * rstd;
    b.coef[c] = scale; b.coef[b.C + c] = b.beta[c] - meanf * scale;
  }
}

__device__ __forceinline__ double* bn_bwd_acc_copy(const BnBwdFused& b) {
  return b.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 * b.C;
}

// Backward counterpart of bn_fused_finish: called by ALL threads of the CTA after their atomicAdd contributions
// to the [copies][3][C] accumulators; the last CTA writes the backward coefficients (and the raw sums, from which
// the BatchNorm-backward apply kernel publishes d gamma / d beta).
__device__ __forceinline__ void bn_bwd_fused_finish(const BnBwdFused& b, unsigned int nctas) {
  __shared__ int bnb_is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) bnb_is_last = (atomicAdd(b.counter, 1u) == nctas - 1u) ? 1 : 0;
  __syncthreads();
  if (!bnb_is_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < b.C; c += blockDim.x) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {           // fixed order over the copies
      const double* ak = b.acc + (size_t)k * 3 * b.C;
      s0 += ld_cg_f64(ak + c); s1 += ld_cg_f64(ak + b.C + c); s2 += ld_cg_f64(ak + 2 * b.C + c);
    }
    b.bcoef[c] = b.gamma[c] * b.stat[b.C + c];
    b.bcoef[b.C + c] = (float)(s0 * b.inv_rows);
    b.bcoef[2 * b.C + c] = (float)(s1 * b.inv_rows);
    b.bcoef[3 * b.C + c] = (float)s0;
    b.bcoef[4 * b.C + c] = (float)s1;
    if (b.y2) {
      b.bcoef2[c] = b.gamma2[c] * b.stat2[b.C + c];
      b.bcoef2[b.C + c] = (float)(s0 * b.inv_rows);
      b.bcoef2[2 * b.C + c] = (float)(s2 * b.inv_rows);
      b.bcoef2[3 * b.C + c] = (float)s0;
      b.bcoef2[4 * b.C + c] = (float)s2;
    }
  }
}

}  // namespace mmvae
