// loss.cu -- VAE.loss (reference model.py:385-406): Gaussian NLL with fixed sigma (model.py:403) or
// weighted cross-entropy (model.py:400-401), plus the closed-form KL (model.py:364-365), as ONE
// vectorised warp-shuffle reduction kernel over recon/target whose last CTA combines the per-block
// partials in fp64 in a fixed order (no second launch); ONE backward kernel writes d recon, d mu and
// d logvar.  Also the MMD diagnostic (model.py:367-383), the Philox normal generator and fused Adam.
//
// HBM-bound: algorithmic bytes = read recon + read target (forward), + write d_recon (backward).
#include "kernels.cuh"

namespace mmvae {

namespace {

constexpr int kLossBlocks = 592;      // 4 x 148 SMs
constexpr int kLossThreads = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

__device__ __forceinline__ float block_sum_f(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (kLossThreads / 32) ? sh[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;   // valid in warp 0
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}

// scratch layout: float partial[kLossBlocks] (reconstruction term), float klpart[kLossBlocks] (KL terms), unsigned counter
struct LossScratch { float partial[kLossBlocks]; float klpart[kLossBlocks]; unsigned int counter; unsigned int pad[3]; };

// The CTA that finishes last combines the per-CTA partials in fp64 in a fixed order (deterministic whatever the
// arrival order) and resets the counter for the next call.
__device__ void loss_combine(const LossArgs& a, LossScratch* sc, int nparts, int has_recon, float klw, float* out) {
  __shared__ double shd[2][kLossThreads / 32];
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&sc->counter, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0, k = 0.0;
  for (int i = tid; i < nparts; i += kLossThreads) {
    s += (double)__ldcg(sc->partial + i);
    k += (double)__ldcg(sc->klpart + i);
  }
  s = warp_sum_d(s); k = warp_sum_d(k);
  if (lane == 0) { shd[0][warp] = s; shd[1][warp] = k; }
  __syncthreads();
  if (warp == 0) {
    s = warp_sum_d(lane < kLossThreads / 32 ? shd[0][lane] : 0.0);
    k = -0.5 * warp_sum_d(lane < kLossThreads / 32 ? shd[1][lane] : 0.0);
    if (lane == 0) {
      double pxz;
      if (!has_recon) {
        pxz = 0.0;
      } else if (a.kind == 0) {
        const double cnt = (double)a.N * a.C * a.H * a.W;
        pxz = (double)a.nll * (s / (2.0 * (double)a.sigma * (double)a.sigma) +
                               cnt * (log((double)a.sigma) + 0.91893853320467274178));
      } else {
        pxz = (double)a.nll * s;
      }
      const double invn = 1.0 / (double)a.N;
      out[0] = (float)((pxz + (double)klw * k) * invn);
      out[1] = (float)(pxz * invn);
      out[2] = (float)(k * invn);
      sc->counter = 0u;
    }
  }
}

// KL terms of this CTA's grid-stride share of the N*z latent entries (model.py:365), fp32 in runs of <= 16
__device__ __forceinline__ float kl_share(const float* __restrict__ mu, const float* __restrict__ lv, long long nz) {
  float k = 0.f;
  if (mu && lv)
    for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < nz; i += (long long)gridDim.x * kLossThreads) {
      const float l = __ldg(lv + i), m = __ldg(mu + i);
      k += l - expf(l) - m * m + 1.0f;
    }
  return k;
}

// Gaussian NLL (model.py:403): partial of sum (t - r)^2; the constant log(sigma) + 0.5*log(2*pi) is added analytically.
// recon == nullptr: KL only.
__global__ void __launch_bounds__(kLossThreads) loss_gauss_fwd_kernel(LossArgs a, const float* __restrict__ recon,
                                                                      const float* __restrict__ target,
                                                                      const float* __restrict__ mu, const float* __restrict__ lv,
                                                                      const float* __restrict__ kl_dev,
                                                                      LossScratch* __restrict__ sc, float* __restrict__ out) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ float sh[kLossThreads / 32];
  float s = 0.f;
  if (recon) {
    const long long n = (long long)a.N * a.C * a.H * a.W;
    const long long n4 = n >> 2;
    const float4* r4 = reinterpret_cast<const float4*>(recon);
    const float4* t4 = reinterpret_cast<const float4*>(target);
    for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kLossThreads) {
      float4 r = __ldg(r4 + i), t = __ldg(t4 + i);
      float d0 = t.x - r.x, d1 = t.y - r.y, d2 = t.z - r.z, d3 = t.w - r.w;
      s += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
      float d0 = target[(n4 << 2) + threadIdx.x] - recon[(n4 << 2) + threadIdx.x];
      s += d0 * d0;
    }
  }
  float k = kl_share(mu, lv, (long long)a.N * a.z);
  s = block_sum_f(s, sh);
  k = block_sum_f(k, sh);
  if (threadIdx.x == 0) { sc->partial[blockIdx.x] = s; sc->klpart[blockIdx.x] = k; }
  loss_combine(a, sc, gridDim.x, recon != nullptr, kl_dev ? __ldg(kl_dev) : a.kl, out);
}

// Cross-entropy over C logits per pixel (NCHW), weighted by w[target] (model.py:400-401).  A target outside [0, C)
// poisons the loss with NaN instead of reading out of bounds (torch raises a device-side assert there).
__global__ void __launch_bounds__(kLossThreads) loss_ce_fwd_kernel(LossArgs a, const float* __restrict__ recon,
                                                                   const long long* __restrict__ target,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ mu, const float* __restrict__ lv,
                                                                   const float* __restrict__ kl_dev,
                                                                   LossScratch* __restrict__ sc, float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sh[kLossThreads / 32];
  float s = 0.f;
  const int C = a.C, HW = a.H * a.W;
  const long long npix = (long long)a.N * HW;
  for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < npix; i += (long long)gridDim.x * kLossThreads) {
    long long n = i / HW; int hw = (int)(i % HW);
    const float* l = recon + (n * C) * HW + hw;
    const long long tl = target[i];
    if (tl < 0 || tl >= C) { s = __int_as_float(0x7fc00000); continue; }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + (size_t)c * HW));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(__ldg(l + (size_t)c * HW) - mx);
    const int t = (int)tl;
    float lp = __ldg(l + (size_t)t * HW) - mx - logf(se);
    s -= (w ? __ldg(w + t) : 1.0f) * lp;
  }
  float k = kl_share(mu, lv, (long long)a.N * a.z);
  s = block_sum_f(s, sh);
  k = block_sum_f(k, sh);
  if (threadIdx.x == 0) { sc->partial[blockIdx.x] = s; sc->klpart[blockIdx.x] = k; }
  loss_combine(a, sc, gridDim.x, 1, kl_dev ? __ldg(kl_dev) : a.kl, out);
}

// d mu = g * kl/N * mu, d logvar = g * kl/N * 0.5 (exp(logvar) - 1): the CTAs past the reconstruction part of a backward grid
__device__ __forceinline__ void kl_bwd_share(const float* __restrict__ mu, const float* __restrict__ lv, long long nz, float g,
                                             float* __restrict__ d_mu, float* __restrict__ d_lv, int first_block) {
  const int nb = (int)gridDim.x - first_block;
  for (long long i = (blockIdx.x - first_block) * 256LL + threadIdx.x; i < nz; i += nb * 256LL) {
    d_mu[i] = g * mu[i];
    d_lv[i] = g * 0.5f * (expf(lv[i]) - 1.0f);
  }
}

// One backward launch: CTAs [0, rblocks) write d recon, CTAs [rblocks, grid) write d mu / d logvar.
__global__ void __launch_bounds__(256) loss_gauss_bwd_kernel(LossArgs a, const float* __restrict__ recon,
                                                             const float* __restrict__ target,
                                                             const float* __restrict__ mu, const float* __restrict__ lv,
                                                             const float* __restrict__ gout, const float* __restrict__ kl_dev,
                                                             float* __restrict__ d_recon, float* __restrict__ d_mu,
                                                             float* __restrict__ d_lv, int rblocks) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  const float up = __ldg(gout);
  if ((int)blockIdx.x >= rblocks) {
    kl_bwd_share(mu, lv, (long long)a.N * a.z, up * (kl_dev ? __ldg(kl_dev) : a.kl) / (float)a.N, d_mu, d_lv, rblocks);
    return;
  }
  const long long n = (long long)a.N * a.C * a.H * a.W;
  const float g = up * a.nll / (a.sigma * a.sigma * (float)a.N);
  const long long n4 = n >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float4* d4 = reinterpret_cast<float4*>(d_recon);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += rblocks * 256LL) {
    float4 r = __ldg(r4 + i), t = __ldg(t4 + i), d;
    d.x = g * (r.x - t.x); d.y = g * (r.y - t.y); d.z = g * (r.z - t.z); d.w = g * (r.w - t.w);
    d4[i] = d;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    long long i = (n4 << 2) + threadIdx.x;
    d_recon[i] = g * (recon[i] - target[i]);
  }
}

__global__ void __launch_bounds__(256) loss_ce_bwd_kernel(LossArgs a, const float* __restrict__ recon,
                                                          const long long* __restrict__ target, const float* __restrict__ w,
                                                          const float* __restrict__ mu, const float* __restrict__ lv,
                                                          const float* __restrict__ gout, const float* __restrict__ kl_dev,
                                                          float* __restrict__ d_recon, float* __restrict__ d_mu,
                                                          float* __restrict__ d_lv, int rblocks) {
  pdl_wait();
  pdl_trigger();
  const float up = __ldg(gout);
  if ((int)blockIdx.x >= rblocks) {
    kl_bwd_share(mu, lv, (long long)a.N * a.z, up * (kl_dev ? __ldg(kl_dev) : a.kl) / (float)a.N, d_mu, d_lv, rblocks);
    return;
  }
  const int C = a.C, HW = a.H * a.W;
  const long long npix = (long long)a.N * HW;
  const float g = up * a.nll / (float)a.N;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < npix; i += rblocks * 256LL) {
    long long n = i / HW; int hw = (int)(i % HW);
    const float* l = recon + (n * C) * HW + hw;
    float* d = d_recon + (n * C) * HW + hw;
    const long long tl = target[i];
    if (tl < 0 || tl >= C) {                      // see loss_ce_fwd_kernel
      for (int c = 0; c < C; ++c) d[(size_t)c * HW] = __int_as_float(0x7fc00000);
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + (size_t)c * HW));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(__ldg(l + (size_t)c * HW) - mx);
    const float inv = 1.0f / se;
    const int t = (int)tl;
    const float wt = g * (w ? __ldg(w + t) : 1.0f);
    for (int c = 0; c < C; ++c) {
      float p = expf(__ldg(l + (size_t)c * HW) - mx) * inv;
      d[(size_t)c * HW] = wt * (p - (c == t ? 1.0f : 0.0f));
    }
  }
}

// ---- MMD diagnostic (model.py:367-383): k(a, b) = exp(-mean_d((a_d - b_d)^2) / dim) = exp(-|a - b|^2 / dim^2),
// mmd = sum k(x,x) + sum k(y,y) - 2 sum k(x,y) over all N x N pairs, x = true_samples, y = encoding.
// grid (ceil(N / kMmdRows), 3): CTA (t, which) holds kMmdRows rows of the left operand in shared memory; thread j walks
// the rows of the right operand (one row per thread and pass, 16-byte loads, each element reused for all kMmdRows left
// rows) and sums k; the per-CTA sums are combined in fp64 in a fixed order by the CTA that finishes last.
constexpr int kMmdRows = 8, kMmdThreads = 256;
struct MmdScratch { unsigned int counter; unsigned int pad[3]; float part[1]; };
__global__ void __launch_bounds__(kMmdThreads) mmd_kernel(const float* __restrict__ x, const float* __restrict__ y, int N, int z,
                                                          MmdScratch* __restrict__ sc, float* __restrict__ out) {
  extern __shared__ float rows[];                      // [kMmdRows][z]
  __shared__ float sh[kMmdThreads / 32];
  __shared__ double shd[kMmdThreads / 32];
  __shared__ int is_last;
  const int which = blockIdx.y, i0 = blockIdx.x * kMmdRows;
  const float* A = which == 1 ? y : x;                 // 0: (x,x)  1: (y,y)  2: (x,y)
  const float* B = which == 0 ? x : y;
  const int nrows = min(kMmdRows, N - i0);
  for (int e = threadIdx.x; e < kMmdRows * z; e += kMmdThreads) rows[e] = e < nrows * z ? A[(size_t)i0 * z + e] : 0.f;
  __syncthreads();
  const float inv = 1.0f / ((float)z * (float)z);
  float s = 0.f;
  for (int j = threadIdx.x; j < N; j += kMmdThreads) {
    const float* b = B + (size_t)j * z;
    float q[kMmdRows];
#pragma unroll
    for (int r = 0; r < kMmdRows; ++r) q[r] = 0.f;
    if ((z & 3) == 0) {
#pragma unroll 4
      for (int d = 0; d < z; d += 4) {
        const float4 bv = __ldg(reinterpret_cast<const float4*>(b + d));
#pragma unroll
        for (int r = 0; r < kMmdRows; ++r) {
          const float4 av = *reinterpret_cast<const float4*>(rows + r * z + d);
          const float t0 = av.x - bv.x, t1 = av.y - bv.y, t2 = av.z - bv.z, t3 = av.w - bv.w;
          q[r] = fmaf(t0, t0, fmaf(t1, t1, fmaf(t2, t2, fmaf(t3, t3, q[r]))));
        }
      }
    } else {
      for (int d = 0; d < z; ++d) {
        const float bv = __ldg(b + d);
#pragma unroll
        for (int r = 0; r < kMmdRows; ++r) { const float t = rows[r * z + d] - bv; q[r] = fmaf(t, t, q[r]); }
      }
    }
#pragma unroll
    for (int r = 0; r < kMmdRows; ++r) if (r < nrows) s += expf(-q[r] * inv);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  const int nparts = gridDim.x;
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kMmdThreads / 32; ++w) t += sh[w];
    sc->part[which * nparts + blockIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(&sc->counter, 1u) == gridDim.x * gridDim.y - 1u) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double t = 0.0;
  for (int e = threadIdx.x; e < 3 * nparts; e += kMmdThreads) t += (e >= 2 * nparts ? -2.0 : 1.0) * (double)__ldcg(sc->part + e);
  t = warp_sum_d(t);
  if ((threadIdx.x & 31) == 0) shd[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < kMmdThreads / 32; ++w) tot += shd[w];
    out[0] = (float)(tot / (double)N);                 // MMD / N, the 4th return value (model.py:406)
    sc->counter = 0u;
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(unsigned long long seed, unsigned long long offset,
                                                            const unsigned long long* __restrict__ rng_dev,
                                                            unsigned long long stream_id, long long n, float* __restrict__ out) {
  if (rng_dev) { seed = rng_dev[0]; offset = rng_dev[1]; }
  seed ^= stream_id;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    out[i] = philox_normal_at(seed, offset, i);
}

// torch.optim.Adam defaults (main.py:468): bias-corrected, eps added outside the sqrt.
__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                                                   float eps, float wd, float bc1, float bc2_sqrt, float gscale,
                                                   const long long* __restrict__ step_dev) {
  if (step_dev) {                                   // step count in device memory: one captured graph serves every step
    const float t = (float)(*step_dev);
    bc1 = 1.f - powf(b1, t);
    bc2_sqrt = sqrtf(1.f - powf(b2, t));
  }
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    float gi = g[i] * gscale;
    float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    float mi = fmaf(b1, m[i], (1.f - b1) * gi);
    float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
    m[i] = mi; v[i] = vi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

// x = (label - mean) / std with an IEEE fp32 subtract and divide: bit-identical to the reference's torch expression
// (main.py:383-388) on the host
__global__ void __launch_bounds__(256) prepare_input_kernel(const unsigned char* __restrict__ labels, long long n, float mean,
                                                            float std, float* __restrict__ x, long long* __restrict__ target) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    unsigned char l = labels[i];
    x[i] = __fdiv_rn(__fsub_rn((float)l, mean), std);
    if (target) target[i] = (long long)l;
  }
}

inline int grid_for(long long items, int per_block = 256, int cap = 148 * 8) {
  long long b = (items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

}  // namespace

size_t loss_scratch_bytes() { return sizeof(LossScratch); }

// scratch must be zero before its first use (the counter); every call leaves it zero again.
void launch_loss_fwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, const float* kl_dev, float* out, void* scratch, cudaStream_t st) {
  LossScratch* sc = reinterpret_cast<LossScratch*>(scratch);
  int blocks;
  if (!recon || a.kind == 0) {
    long long n = recon ? (long long)a.N * a.C * a.H * a.W : (long long)a.N * a.z * 4;
    blocks = (int)((n / 4 + kLossThreads - 1) / kLossThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > kLossBlocks) blocks = kLossBlocks;
    count_launch();
    launch_pdl(loss_gauss_fwd_kernel, blocks, kLossThreads, 0, st, a, recon, reinterpret_cast<const float*>(target), mu, lv,
               kl_dev, sc, out);
  } else {
    long long npix = (long long)a.N * a.H * a.W;
    blocks = (int)((npix + kLossThreads - 1) / kLossThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > kLossBlocks) blocks = kLossBlocks;
    count_launch();
    launch_pdl(loss_ce_fwd_kernel, blocks, kLossThreads, 0, st, a, recon, reinterpret_cast<const long long*>(target), w, mu,
               lv, kl_dev, sc, out);
  }
}

void launch_loss_bwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, const float* gout, const float* kl_dev,
                     float* d_recon, float* d_mu, float* d_lv, cudaStream_t st) {
  const bool kl = d_mu && d_lv && mu && lv;
  const long long nz = (long long)a.N * a.z;
  const int kblocks = kl ? grid_for(nz, 256, 64) : 0;
  int rblocks = 0;
  if (d_recon) rblocks = a.kind == 0 ? grid_for(((long long)a.N * a.C * a.H * a.W) / 4 + 1) : grid_for((long long)a.N * a.H * a.W);
  if (rblocks + kblocks == 0) return;
  count_launch();
  if (a.kind == 0)
    launch_pdl(loss_gauss_bwd_kernel, rblocks + kblocks, 256, 0, st, a, recon, reinterpret_cast<const float*>(target), mu, lv,
               gout, kl_dev, d_recon, d_mu, d_lv, rblocks);
  else
    launch_pdl(loss_ce_bwd_kernel, rblocks + kblocks, 256, 0, st, a, recon, reinterpret_cast<const long long*>(target), w, mu,
               lv, gout, kl_dev, d_recon, d_mu, d_lv, rblocks);
}

size_t mmd_scratch_bytes(int N) { return sizeof(MmdScratch) + sizeof(float) * 3 * (size_t)((N + kMmdRows - 1) / kMmdRows); }
void launch_mmd(const float* x, const float* y, int N, int z, float* out, void* scratch, cudaStream_t st) {
  count_launch();
  mmd_kernel<<<dim3((N + kMmdRows - 1) / kMmdRows, 3), kMmdThreads, sizeof(float) * kMmdRows * z, st>>>(
      x, y, N, z, reinterpret_cast<MmdScratch*>(scratch), out);
}

void launch_prepare_input(const unsigned char* labels, long long n, float mean, float std, float* x,
                          long long* target, cudaStream_t st) {
  count_launch();
  launch_pdl(prepare_input_kernel, grid_for(n), 256, 0, st, labels, n, mean, std, x, target);
}

void launch_philox_normal(unsigned long long seed, unsigned long long offset, const unsigned long long* rng_dev,
                          unsigned long long stream_id, long long n, float* out, cudaStream_t st) {
  count_launch();
  philox_normal_kernel<<<grid_for(n), 256, 0, st>>>(seed, offset, rng_dev, stream_id, n, out);
}

void launch_adam(long long n, float* p, const float* g, float* m, float* v, float lr, float b1, float b2,
                 float eps, float wd, long long step, const long long* step_dev, float gscale, cudaStream_t st) {
  double bc1 = 1.0 - pow((double)b1, (double)step);
  double bc2 = 1.0 - pow((double)b2, (double)step);
  count_launch();
  adam_kernel<<<grid_for(n), 256, 0, st>>>(n, p, g, m, v, lr, b1, b2, eps, wd, (float)bc1, (float)sqrt(bc2), gscale, step_dev);
}

}  // namespace mmvae
