// loss.cu -- VAE.loss (reference model.py:385-406): Gaussian NLL with fixed sigma (model.py:403) or
// weighted cross-entropy (model.py:400-401), plus the closed-form KL (model.py:364-365), as ONE
// vectorised warp-shuffle reduction pass over recon/target; a second tiny kernel combines the
// per-block partials in fp64 in a fixed order.  Also the Philox normal generator and fused Adam.
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

// Gaussian: partial of sum (t - r)^2 ; the constant log(sigma) + 0.5*log(2*pi) is added analytically.
__global__ void __launch_bounds__(kLossThreads) loss_gauss_fwd_kernel(const float* __restrict__ recon,
                                                                      const float* __restrict__ target,
                                                                      long long n, float* __restrict__ partial) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ float sh[kLossThreads / 32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kLossThreads) {
    float4 r = __ldg(r4 + i), t = __ldg(t4 + i);
    float a = t.x - r.x, b = t.y - r.y, c = t.z - r.z, d = t.w - r.w;
    s += (a * a + b * b) + (c * c + d * d);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    float a = target[(n4 << 2) + threadIdx.x] - recon[(n4 << 2) + threadIdx.x];
    s += a * a;
  }
  s = block_sum_f(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Cross-entropy over C logits per pixel (NCHW), weighted by w[target].
__global__ void __launch_bounds__(kLossThreads) loss_ce_fwd_kernel(const float* __restrict__ recon,
                                                                   const long long* __restrict__ target,
                                                                   const float* __restrict__ w, long long npix, int C, int HW,
                                                                   float* __restrict__ partial) {
  __shared__ float sh[kLossThreads / 32];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)kLossThreads + threadIdx.x; i < npix; i += (long long)gridDim.x * kLossThreads) {
    long long n = i / HW; int hw = (int)(i % HW);
    const float* l = recon + (n * C) * HW + hw;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + (size_t)c * HW));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(__ldg(l + (size_t)c * HW) - mx);
    int t = (int)target[i];
    float lp = __ldg(l + (size_t)t * HW) - mx - logf(se);
    s -= (w ? __ldg(w + t) : 1.0f) * lp;
  }
  s = block_sum_f(s, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// One CTA of 1024 threads: fp32 per-thread partials (<= 16 KL terms / <= 1 loss partial each), combined in fp64
// by warp shuffles in a fixed order.
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__global__ void __launch_bounds__(1024) loss_finalize_kernel(LossArgs a, const float* __restrict__ partial, int nparts,
                                                             const float* __restrict__ mu, const float* __restrict__ lv,
                                                             float* __restrict__ out) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ double sh[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double s = 0.0;
  for (int i = tid; i < nparts; i += 1024) s += (double)partial[i];
  double k = 0.0;
  if (mu && lv) {
    const long long nz = (long long)a.N * a.z;
    float kf = 0.f;
    int cnt = 0;
    for (long long i = tid; i < nz; i += 1024) {
      const float l = __ldg(lv + i), m = __ldg(mu + i);
      kf += l - expf(l) - m * m + 1.0f;
      if (++cnt == 16) { k += (double)kf; kf = 0.f; cnt = 0; }      // bounded fp32 run length
    }
    k += (double)kf;
  }
  s = warp_sum_d(s); k = warp_sum_d(k);
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = k; }
  __syncthreads();
  if (warp == 0) {
    s = warp_sum_d(sh[0][lane]); k = -0.5 * warp_sum_d(sh[1][lane]);
    if (lane == 0) {
      double pxz;
      if (nparts == 0) {
        pxz = 0.0;
      } else if (a.kind == 0) {
        const double cnt = (double)a.N * a.C * a.H * a.W;
        pxz = (double)a.nll * (s / (2.0 * (double)a.sigma * (double)a.sigma) +
                               cnt * (log((double)a.sigma) + 0.91893853320467274178));
      } else {
        pxz = (double)a.nll * s;
      }
      const double invn = 1.0 / (double)a.N;
      out[0] = (float)((pxz + (double)a.kl * k) * invn);
      out[1] = (float)(pxz * invn);
      out[2] = (float)(k * invn);
    }
  }
}

__global__ void __launch_bounds__(256) loss_gauss_bwd_kernel(const float* __restrict__ recon, const float* __restrict__ target,
                                                             long long n, const float* __restrict__ gout, float coef,
                                                             float* __restrict__ d_recon) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  const float g = __ldg(gout) * coef;     // nll / (sigma^2 * N) * upstream
  const long long n4 = n >> 2;
  const float4* r4 = reinterpret_cast<const float4*>(recon);
  const float4* t4 = reinterpret_cast<const float4*>(target);
  float4* d4 = reinterpret_cast<float4*>(d_recon);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n4; i += gridDim.x * 256LL) {
    float4 r = __ldg(r4 + i), t = __ldg(t4 + i), d;
    d.x = g * (r.x - t.x); d.y = g * (r.y - t.y); d.z = g * (r.z - t.z); d.w = g * (r.w - t.w);
    d4[i] = d;
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    long long i = (n4 << 2) + threadIdx.x;
    d_recon[i] = g * (recon[i] - target[i]);
  }
}

__global__ void __launch_bounds__(256) loss_ce_bwd_kernel(const float* __restrict__ recon, const long long* __restrict__ target,
                                                          const float* __restrict__ w, long long npix, int C, int HW,
                                                          const float* __restrict__ gout, float coef, float* __restrict__ d_recon) {
  const float g = __ldg(gout) * coef;     // nll / N * upstream
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < npix; i += gridDim.x * 256LL) {
    long long n = i / HW; int hw = (int)(i % HW);
    const float* l = recon + (n * C) * HW + hw;
    float* d = d_recon + (n * C) * HW + hw;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, __ldg(l + (size_t)c * HW));
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(__ldg(l + (size_t)c * HW) - mx);
    const float inv = 1.0f / se;
    int t = (int)target[i];
    const float wt = g * (w ? __ldg(w + t) : 1.0f);
    for (int c = 0; c < C; ++c) {
      float p = expf(__ldg(l + (size_t)c * HW) - mx) * inv;
      d[(size_t)c * HW] = wt * (p - (c == t ? 1.0f : 0.0f));
    }
  }
}

__global__ void __launch_bounds__(256) kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, long long nz,
                                                     const float* __restrict__ gout, float coef,
                                                     float* __restrict__ d_mu, float* __restrict__ d_lv) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  const float g = __ldg(gout) * coef;     // kl / N * upstream
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nz; i += gridDim.x * 256LL) {
    d_mu[i] = g * mu[i];
    d_lv[i] = g * 0.5f * (expf(lv[i]) - 1.0f);
  }
}

__global__ void __launch_bounds__(256) philox_normal_kernel(unsigned long long seed, unsigned long long offset, long long n,
                                                            float* __restrict__ out) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL)
    out[i] = philox_normal_at(seed, offset, i);
}

// torch.optim.Adam defaults (main.py:468): bias-corrected, eps added outside the sqrt.
__global__ void __launch_bounds__(256) adam_kernel(long long n, float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v, float lr, float b1, float b2,
                                                   float eps, float wd, float bc1, float bc2_sqrt, float gscale) {
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

__global__ void __launch_bounds__(256) prepare_input_kernel(const unsigned char* __restrict__ labels, long long n, float mean,
                                                            float inv_std, float* __restrict__ x, long long* __restrict__ target) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    unsigned char l = labels[i];
    x[i] = ((float)l - mean) * inv_std;
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

size_t loss_scratch_bytes() { return sizeof(float) * kLossBlocks; }

void launch_loss_fwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, float* out, void* scratch, cudaStream_t st) {
  float* partial = reinterpret_cast<float*>(scratch);
  int blocks;
  if (!recon) {
    blocks = 0;                       // KL only
  } else if (a.kind == 0) {
    long long n = (long long)a.N * a.C * a.H * a.W;
    blocks = (int)((n / 4 + kLossThreads - 1) / kLossThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > kLossBlocks) blocks = kLossBlocks;
    count_launch();
    launch_pdl(loss_gauss_fwd_kernel, blocks, kLossThreads, 0, st, recon, reinterpret_cast<const float*>(target), n, partial);
  } else {
    long long npix = (long long)a.N * a.H * a.W;
    blocks = (int)((npix + kLossThreads - 1) / kLossThreads);
    if (blocks < 1) blocks = 1;
    if (blocks > kLossBlocks) blocks = kLossBlocks;
    count_launch();
    loss_ce_fwd_kernel<<<blocks, kLossThreads, 0, st>>>(recon, reinterpret_cast<const long long*>(target), w, npix,
                                                        a.C, a.H * a.W, partial);
  }
  count_launch();
  launch_pdl(loss_finalize_kernel, 1, 1024, 0, st, a, partial, blocks, mu, lv, out);
}

void launch_loss_bwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, const float* gout,
                     float* d_recon, float* d_mu, float* d_lv, cudaStream_t st) {
  if (d_recon) {
    if (a.kind == 0) {
      long long n = (long long)a.N * a.C * a.H * a.W;
      float coef = a.nll / (a.sigma * a.sigma * (float)a.N);
      count_launch();
      launch_pdl(loss_gauss_bwd_kernel, grid_for(n / 4 + 1), 256, 0, st, recon, reinterpret_cast<const float*>(target), n, gout,
                                                                 coef, d_recon);
    } else {
      long long npix = (long long)a.N * a.H * a.W;
      float coef = a.nll / (float)a.N;
      count_launch();
      loss_ce_bwd_kernel<<<grid_for(npix), 256, 0, st>>>(recon, reinterpret_cast<const long long*>(target), w, npix, a.C,
                                                         a.H * a.W, gout, coef, d_recon);
    }
  }
  if (d_mu && d_lv && mu && lv) {
    long long nz = (long long)a.N * a.z;
    count_launch();
    launch_pdl(kl_bwd_kernel, grid_for(nz), 256, 0, st, mu, lv, nz, gout, a.kl / (float)a.N, d_mu, d_lv);
  }
}

void launch_prepare_input(const unsigned char* labels, long long n, float mean, float inv_std, float* x,
                          long long* target, cudaStream_t st) {
  count_launch();
  launch_pdl(prepare_input_kernel, grid_for(n), 256, 0, st, labels, n, mean, inv_std, x, target);
}

void launch_philox_normal(unsigned long long seed, unsigned long long offset, long long n, float* out, cudaStream_t st) {
  count_launch();
  philox_normal_kernel<<<grid_for(n), 256, 0, st>>>(seed, offset, n, out);
}

void launch_adam(long long n, float* p, const float* g, float* m, float* v, float lr, float b1, float b2,
                 float eps, float wd, long long step, float gscale, cudaStream_t st) {
  double bc1 = 1.0 - pow((double)b1, (double)step);
  double bc2 = 1.0 - pow((double)b2, (double)step);
  count_launch();
  adam_kernel<<<grid_for(n), 256, 0, st>>>(n, p, g, m, v, lr, b1, b2, eps, wd, (float)bc1, (float)sqrt(bc2), gscale);
}

}  // namespace mmvae
