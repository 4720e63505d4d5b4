// nb.cu -- pointwise / loss kernels of the notebook variant of the VAE (vae-kl.ipynb:119-166, loop body :210-233;
// SURVEY.md 8(a) row 12, BASELINE configs[4]).  That network has no BatchNorm: bias + ReLU / ELU live in the
// convolution epilogues (GConvParams::act / dact), what remains here is HBM-bound streaming work:
//   nearest upsample (vae-kl.ipynb:152-155) and its adjoint fused with the activation derivative,
//   rsample (vae-kl.ipynb:144-146) and its adjoint fused with the closed-form KL gradient,
//   softmax cross-entropy over the 256 grey levels (vae-kl.ipynb:226) fused with d logits and the last bias gradient.
#include <cstdlib>

#include "kernels.cuh"

namespace mmvae {

namespace {

template <typename T> struct Vec;                       // 16-byte vector of storage elements
template <> struct Vec<float> { static constexpr int n = 4; };
template <> struct Vec<__nv_bfloat16> { static constexpr int n = 8; };

template <typename T>
__device__ __forceinline__ void load_vec(const T* p, float* v) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  if constexpr (sizeof(T) == 4) {
    v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
  } else {
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(&r);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = __bfloat162float(b[e]);
  }
}
template <typename T>
__device__ __forceinline__ void store_vec(T* p, const float* v) {
  uint4 r;
  if constexpr (sizeof(T) == 4) {
    r.x = __float_as_uint(v[0]); r.y = __float_as_uint(v[1]); r.z = __float_as_uint(v[2]); r.w = __float_as_uint(v[3]);
  } else {
    __nv_bfloat16* b = reinterpret_cast<__nv_bfloat16*>(&r);
#pragma unroll
    for (int e = 0; e < 8; ++e) b[e] = __float2bfloat16_rn(v[e]);
  }
  *reinterpret_cast<uint4*>(p) = r;
}

// one thread = one 16-byte channel vector of one INPUT pixel: read once, written f x f times (32-bit index arithmetic)
template <typename T>
__global__ void __launch_bounds__(256) nb_upsample_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C,
                                                         int f, long long total) {
  constexpr int V = Vec<T>::n;
  pdl_wait();
  pdl_trigger();
  const unsigned cv = (unsigned)(C / V), uW = (unsigned)W, uH = (unsigned)H;
  const unsigned Wo = uW * (unsigned)f, Ho = uH * (unsigned)f;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < (unsigned)total; i += gridDim.x * 256u) {
    const unsigned c = i % cv;
    unsigned pix = i / cv;
    const unsigned x = pix % uW; pix /= uW;
    const unsigned y = pix % uH;
    const unsigned n = pix / uH;
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(in) + i);
    uint4* o = reinterpret_cast<uint4*>(out) + ((size_t)(n * Ho + y * f) * Wo + x * f) * cv + c;
    for (int a = 0; a < f; ++a)
      for (int b = 0; b < f; ++b) o[((size_t)a * Wo + b) * cv] = r;
  }
}

// one thread = one 16-byte channel vector of one INPUT-resolution pixel: sums its f x f block
template <typename T>
__global__ void __launch_bounds__(256) nb_upsample_bwd_kernel(const T* __restrict__ dup, const T* __restrict__ a, int act_kind,
                                                             T* __restrict__ dy, int H, int W, int C, int f, long long total) {
  constexpr int V = Vec<T>::n;
  pdl_wait();
  pdl_trigger();
  const unsigned cv = (unsigned)(C / V), uW = (unsigned)W, uH = (unsigned)H;
  const unsigned Wo = uW * (unsigned)f, Ho = uH * (unsigned)f;
  for (unsigned i = blockIdx.x * 256u + threadIdx.x; i < (unsigned)total; i += gridDim.x * 256u) {
    const unsigned c = i % cv;
    unsigned pix = i / cv;
    const unsigned x = pix % uW; pix /= uW;
    const unsigned y = pix % uH;
    const unsigned n = pix / uH;
    float s[V];
#pragma unroll
    for (int e = 0; e < V; ++e) s[e] = 0.f;
    const T* src = dup + (((size_t)(n * Ho + y * f) * Wo + x * f) * cv + c) * V;
    for (int dy_ = 0; dy_ < f; ++dy_)
      for (int dx_ = 0; dx_ < f; ++dx_) {
        float v[V];
        load_vec<T>(src + ((size_t)dy_ * Wo + dx_) * cv * V, v);
#pragma unroll
        for (int e = 0; e < V; ++e) s[e] += v[e];
      }
    if (a) {
      float av[V];
      load_vec<T>(a + (size_t)i * V, av);
#pragma unroll
      for (int e = 0; e < V; ++e) s[e] *= act_deriv(act_kind, av[e]);
    }
    store_vec<T>(dy + (size_t)i * V, s);
  }
}

// thread = one latent element in NCHW order (the order of the user-visible tensors and of the Philox stream)
template <typename T>
__global__ void __launch_bounds__(256) nb_rsample_kernel(const NbSampleArgs a) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)a.N * a.z * a.hw;
  unsigned long long seed = a.seed, offset = a.offset;
  if (a.rng_dev) { seed = a.rng_dev[0]; offset = a.rng_dev[1]; }
  const T* mu_y = reinterpret_cast<const T*>(a.mu_y);
  const T* lv_y = reinterpret_cast<const T*>(a.lv_y);
  T* z_act = reinterpret_cast<T*>(a.z_act);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int p = (int)(i % a.hw);
    const long long t = i / a.hw;
    const int c = (int)(t % a.z);
    const long long n = t / a.z;
    const long long j = (n * a.hw + p) * a.z + c;            // NHWC index
    const float mu = to_f(mu_y[j]), lv = to_f(lv_y[j]);
    const float eps = a.eps ? a.eps[i] : philox_normal_at(seed, offset, i);
    const float enc = fmaf(eps, expf(0.5f * lv), mu);
    a.eps_keep[i] = eps;
    z_act[j] = from_f<T>(enc);
    if (a.mu_out) a.mu_out[i] = mu;
    if (a.lv_out) a.lv_out[i] = lv;
    if (a.enc_out) a.enc_out[i] = enc;
    if (a.eps_out) a.eps_out[i] = eps;
  }
}

// d mu = dz + klw/N * mu;  d logvar = dz * 0.5 * eps * exp(0.5 logvar) + klw/N * 0.5 * (exp(logvar) - 1)
// KL = -0.5 * sum(logvar - exp(logvar) - mu^2 + 1)                                     (vae-kl.ipynb:119-120, 227)
template <typename T>
__global__ void __launch_bounds__(256) nb_rsample_bwd_kernel(const NbSampleBwdArgs a) {
  pdl_wait();
  pdl_trigger();
  const long long total = (long long)a.N * a.z * a.hw;
  const T* dz = reinterpret_cast<const T*>(a.dz);
  const T* mu_y = reinterpret_cast<const T*>(a.mu_y);
  const T* lv_y = reinterpret_cast<const T*>(a.lv_y);
  T* dmu = reinterpret_cast<T*>(a.d_mu_y);
  T* dlv = reinterpret_cast<T*>(a.d_lv_y);
  float kl = 0.f;
  for (long long j = blockIdx.x * 256LL + threadIdx.x; j < total; j += gridDim.x * 256LL) {
    // j = NHWC index; the eps copy is NCHW
    const int c = (int)(j % a.z);
    const long long t = j / a.z;
    const int p = (int)(t % a.hw);
    const long long n = t / a.hw;
    const float eps = a.eps_keep[(n * a.z + c) * a.hw + p];
    const float mu = to_f(mu_y[j]), lv = to_f(lv_y[j]);
    const float g = dz ? to_f(dz[j]) : 0.f;
    const float ev = expf(lv);
    dmu[j] = from_f<T>(g + a.klw_over_n * mu);
    dlv[j] = from_f<T>(g * 0.5f * eps * expf(0.5f * lv) + a.klw_over_n * 0.5f * (ev - 1.f));
    kl += -0.5f * (lv - ev - mu * mu + 1.f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) kl += __shfl_xor_sync(0xffffffffu, kl, o);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = kl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += red[w];
    atomicAdd(a.kl_acc, (double)s);
  }
}

// Softmax cross-entropy, one warp per pixel row of C <= 256 classes: lane l owns classes [8l, 8l+8).
// Reads the logits once, writes d logits once; the per-class column sums (the bias gradient of the last conv) stay in
// registers over all rows of the warp and leave through shared memory + one atomic per class and CTA.
template <typename T>
__global__ void __launch_bounds__(256) nb_ce_kernel(const T* __restrict__ logits, const long long* __restrict__ target,
                                                   T* __restrict__ dlogits, long long rows, int C, float scale,
                                                   double* ce_acc, float* dbias) {
  __shared__ float bsum[8][256];
  __shared__ float lred[8];
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = lane * 8;
  const bool on = c0 < C;
  float db[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) db[e] = 0.f;
  float loss = 0.f;
  const long long wstride = (long long)gridDim.x * 8;
  for (long long r = (long long)blockIdx.x * 8 + warp; r < rows; r += wstride) {
    float v[8];
    if (on) {
      if constexpr (sizeof(T) == 2) {
        load_vec<T>(logits + r * C + c0, v);
      } else {
        load_vec<T>(logits + r * C + c0, v);
        load_vec<T>(logits + r * C + c0 + 4, v + 4);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = -INFINITY;
    }
    float mx = v[0];
#pragma unroll
    for (int e = 1; e < 8; ++e) mx = fmaxf(mx, v[e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float ex[8], s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) { ex[e] = on ? expf(v[e] - mx) : 0.f; s += ex[e]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const int tg = (int)__ldg(target + r);
    const float inv = 1.f / s;
    // the target's logit lives in lane tg / 8
    float tv = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) if (c0 + e == tg) tv = v[e];
    tv = __shfl_sync(0xffffffffu, tv, tg >> 3);
    if (lane == 0) loss += (mx + logf(s)) - tv;
    if (on) {
      float g[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[e] = (ex[e] * inv - (c0 + e == tg ? 1.f : 0.f)) * scale;
        if constexpr (sizeof(T) == 2) g[e] = __bfloat162float(__float2bfloat16_rn(g[e]));
        db[e] += g[e];
      }
      if constexpr (sizeof(T) == 2) {
        store_vec<T>(dlogits + r * C + c0, g);
      } else {
        store_vec<T>(dlogits + r * C + c0, g);
        store_vec<T>(dlogits + r * C + c0 + 4, g + 4);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) bsum[warp][c0 + e] = db[e];
  if (lane == 0) lred[warp] = loss;
  __syncthreads();
  if (dbias) {
    const int c = threadIdx.x;
    if (c < C) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += bsum[w][c];
      atomicAdd(dbias + c, s);
    }
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < 8; ++w) s += lred[w];
    atomicAdd(ce_acc, (double)s);
  }
}

// column sums of a [rows][C] tensor, C <= 256 and C % 8 == 0: thread (c/8 slot, row lane) accumulates in registers
template <typename T>
__global__ void __launch_bounds__(256) nb_colsum_kernel(const T* __restrict__ dy, long long rows, int C, float* dbias) {
  __shared__ float part[256][9];
  pdl_wait();
  pdl_trigger();
  const int cv = C / 8;                       // 16-byte (bf16) / 32-byte (fp32) channel groups per row
  const int slot = threadIdx.x % cv, rl = threadIdx.x / cv, rpb = 256 / cv;
  float s[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = 0.f;
  if (rl < rpb) {
    for (long long r = (long long)blockIdx.x * rpb + rl; r < rows; r += (long long)gridDim.x * rpb) {
      float v[8];
      if constexpr (sizeof(T) == 2) {
        load_vec<T>(dy + r * C + slot * 8, v);
      } else {
        load_vec<T>(dy + r * C + slot * 8, v);
        load_vec<T>(dy + r * C + slot * 8 + 4, v + 4);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s[e] += v[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) part[threadIdx.x][e] = s[e];
  __syncthreads();
  if (threadIdx.x < C) {
    const int c = threadIdx.x, sl = c / 8, e = c % 8;
    float t = 0.f;
    for (int q = 0; q < rpb; ++q) t += part[q * cv + sl][e];
    atomicAdd(dbias + c, t);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) nb_export_nchw_kernel(const T* __restrict__ in, float* __restrict__ out, int HW, int C,
                                                            long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const long long n = t / C;
    out[i] = to_f(in[(n * HW + p) * C + c]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) nb_import_nchw_kernel(const float* __restrict__ in, T* __restrict__ out, int HW, int C,
                                                            long long total) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    const int p = (int)(i % HW);
    const long long t = i / HW;
    const int c = (int)(t % C);
    const long long n = t / C;
    out[(n * HW + p) * C + c] = from_f<T>(in[i]);
  }
}

__global__ void nb_loss_finalize_kernel(const double* acc, float inv_n, float klw, float* out) {
  pdl_wait();
  pdl_trigger();
  const double ce = acc[0] * (double)inv_n, kl = acc[1] * (double)inv_n;
  out[0] = (float)(ce + (double)klw * kl);
  out[1] = (float)ce;
  out[2] = (float)kl;
}

// ---------------- encoder.conv1 of the notebook variant: Conv2d(1 -> 32, k5, s2, p2) + bias + ReLU (vae-kl.ipynb:124,134) ----------------
// K = 25 on a single input channel is not tensor-core work: register-blocked SIMT over tiles of 8 output rows x Wo columns
// of one frame; the fp32 input window (19 x (2 Wo + 3), zero halo) is staged in shared memory.
constexpr int kStemRows = 8;
constexpr int kStemMaxWo = 64;
constexpr int kStemPitch = 2 * kStemMaxWo + 4;       // floats per staged input row

template <int CO>
__device__ __forceinline__ void nb_stem_load_x(const float* __restrict__ x, float* xs, int n, int oy0, int S, int Wo) {
  const int rows = 2 * kStemRows + 3, cols = 2 * Wo + 3;
  const float* img = x + (size_t)n * S * S;
  for (int e = threadIdx.x; e < rows * cols; e += 256) {
    const int r = e / cols, c = e - r * cols;
    const int iy = 2 * oy0 - 2 + r, ix = c - 2;
    xs[r * kStemPitch + c] = ((unsigned)iy < (unsigned)S && (unsigned)ix < (unsigned)S) ? __ldg(img + (size_t)iy * S + ix) : 0.f;
  }
}

// the same window through 4-byte cp.async (zero fill outside the image): the next tile loads while the current one computes
template <int CO>
__device__ __forceinline__ void nb_stem_prefetch_x(const float* __restrict__ x, float* xs, int n, int oy0, int S, int Wo) {
  const int rows = 2 * kStemRows + 3, cols = 2 * Wo + 3;
  const float* img = x + (size_t)n * S * S;
  const unsigned base = (unsigned)__cvta_generic_to_shared(xs);
  for (int e = threadIdx.x; e < rows * cols; e += 256) {
    const int r = e / cols, c = e - r * cols;
    const int iy = 2 * oy0 - 2 + r, ix = c - 2;
    const bool ok = (unsigned)iy < (unsigned)S && (unsigned)ix < (unsigned)S;
    const float* src = ok ? img + (size_t)iy * S + ix : x;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(base + (unsigned)(r * kStemPitch + c) * 4u), "l"(src),
                 "r"(ok ? 4u : 0u)
                 : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

// forward: warp = 8 output channels, lanes = 32 consecutive pixels of the tile; weights broadcast from shared memory
__global__ void __launch_bounds__(256) nb_stem_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int N,
                                                         int S) {
  __shared__ float xs2[2][(2 * kStemRows + 3) * kStemPitch];
  __shared__ __align__(16) float ws[25][32];
  __shared__ float bs[32];
  const int Wo = S / 2, Ho = S / 2, tiles_per_frame = Ho / kStemRows;
  for (int e = threadIdx.x; e < 800; e += 256) ws[e % 25][e / 25] = __ldg(w + e);      // [co][tap] -> [tap][co]
  if (threadIdx.x < 32) bs[threadIdx.x] = __ldg(bias + threadIdx.x);
  pdl_wait();
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cg = warp & 3, half = warp >> 2;           // channel group of 8; which half of the tile's pixel chunks
  const int npix = kStemRows * Wo, ntiles = N * tiles_per_frame;
  int buf = 0;
  if ((int)blockIdx.x < ntiles)
    nb_stem_prefetch_x<32>(x, xs2[0], blockIdx.x / tiles_per_frame, (blockIdx.x % tiles_per_frame) * kStemRows, S, Wo);
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
    const int n = tile / tiles_per_frame, oy0 = (tile - n * tiles_per_frame) * kStemRows;
    const int next = tile + gridDim.x;
    if (next < ntiles) {                               // the other buffer was released by the barrier that ended the previous tile
      nb_stem_prefetch_x<32>(x, xs2[buf ^ 1], next / tiles_per_frame, (next % tiles_per_frame) * kStemRows, S, Wo);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* xs = xs2[buf];
    for (int p0 = half * 32; p0 < npix; p0 += 64) {
      const int px = p0 + lane, py = px / Wo, pxx = px - py * Wo;
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = bs[cg * 8 + c];
      const float* xp = xs + (2 * py) * kStemPitch + 2 * pxx;
#pragma unroll
      for (int ky = 0; ky < 5; ++ky)
#pragma unroll
        for (int kx = 0; kx < 5; ++kx) {
          const float xv = xp[ky * kStemPitch + kx];
          const float4 w0 = *reinterpret_cast<const float4*>(&ws[ky * 5 + kx][cg * 8]);
          const float4 w1 = *reinterpret_cast<const float4*>(&ws[ky * 5 + kx][cg * 8 + 4]);
          acc[0] = fmaf(xv, w0.x, acc[0]); acc[1] = fmaf(xv, w0.y, acc[1]); acc[2] = fmaf(xv, w0.z, acc[2]); acc[3] = fmaf(xv, w0.w, acc[3]);
          acc[4] = fmaf(xv, w1.x, acc[4]); acc[5] = fmaf(xv, w1.y, acc[5]); acc[6] = fmaf(xv, w1.z, acc[6]); acc[7] = fmaf(xv, w1.w, acc[7]);
        }
      uint4 pk;
      __nv_bfloat162* pw = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int c = 0; c < 4; ++c) pw[c] = __floats2bfloat162_rn(fmaxf(acc[2 * c], 0.f), fmaxf(acc[2 * c + 1], 0.f));
      *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + oy0 + py) * Wo + pxx) * 32 + cg * 8) = pk;
    }
    __syncthreads();                                   // everyone is done with this buffer before it is refilled
  }
}

// weight + bias gradient: lane = output channel, warp = one of the tile's 8 output rows.  The warp slides a 5 x 5 input window
// along its row in registers (stride 2: ten new values per pixel, broadcast shared-memory loads) and keeps all 25 tap sums
// of its channel in registers over every tile of the CTA: 11 shared-memory loads per 25 FMAs (a first version with warp =
// tap group and one load per FMA was bound by the shared-memory pipe: 0.375 ms).  One cross-warp reduction through shared
// memory and one atomicAdd per (tap, channel) per CTA at the end.
__global__ void __launch_bounds__(256) nb_stem_wgrad_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                           float* __restrict__ dw, float* __restrict__ dbias, int N, int S) {
  __shared__ float xs[(2 * kStemRows + 3) * kStemPitch];
  __shared__ __align__(16) __nv_bfloat16 dys[kStemRows * kStemMaxWo * 32];
  pdl_wait();
  pdl_trigger();
  const int Wo = S / 2, Ho = S / 2, tiles_per_frame = Ho / kStemRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int npix = kStemRows * Wo;
  float acc[25], bsum = 0.f;
#pragma unroll
  for (int t = 0; t < 25; ++t) acc[t] = 0.f;
  for (int tile = blockIdx.x; tile < N * tiles_per_frame; tile += gridDim.x) {
    const int n = tile / tiles_per_frame, oy0 = (tile - n * tiles_per_frame) * kStemRows;
    __syncthreads();
    nb_stem_load_x<32>(x, xs, n, oy0, S, Wo);
    const uint4* src = reinterpret_cast<const uint4*>(dy + ((size_t)n * Ho + oy0) * Wo * 32);
    for (int e = threadIdx.x; e < npix * 4; e += 256) reinterpret_cast<uint4*>(dys)[e] = __ldg(src + e);
    __syncthreads();
    const float* xr = xs + (2 * warp) * kStemPitch;      // input rows 2 py .. 2 py + 4 of output row py = warp
    const __nv_bfloat16* dr = dys + (size_t)warp * Wo * 32 + lane;
    float win[5][5];
#pragma unroll
    for (int r = 0; r < 5; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) win[r][c + 2] = xr[r * kStemPitch + c];      // columns 0..2 sit where the first shift puts them
#pragma unroll 2
    for (int pxx = 0; pxx < Wo; ++pxx) {
#pragma unroll
      for (int r = 0; r < 5; ++r) {
        win[r][0] = win[r][2]; win[r][1] = win[r][3]; win[r][2] = win[r][4];
        win[r][3] = xr[r * kStemPitch + 2 * pxx + 3];
        win[r][4] = xr[r * kStemPitch + 2 * pxx + 4];
      }
      const float g = __bfloat162float(dr[pxx * 32]);
      bsum += g;
#pragma unroll
      for (int r = 0; r < 5; ++r)
#pragma unroll
        for (int c = 0; c < 5; ++c) acc[r * 5 + c] = fmaf(win[r][c], g, acc[r * 5 + c]);
    }
  }
  // cross-warp reduction: the dY buffer is free now
  __syncthreads();
  float* red = reinterpret_cast<float*>(dys);            // [8 warps][26][32]
#pragma unroll
  for (int t = 0; t < 25; ++t) red[(warp * 26 + t) * 32 + lane] = acc[t];
  red[(warp * 26 + 25) * 32 + lane] = bsum;
  __syncthreads();
  for (int e = threadIdx.x; e < 26 * 32; e += 256) {
    const int t = e >> 5, co = e & 31;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += red[(w * 26 + t) * 32 + co];
    if (t < 25) atomicAdd(dw + co * 25 + t, v);
    else if (dbias) atomicAdd(dbias + co, v);
  }
}

inline int grid_for(long long work_items, int cap = 148 * 16) {
  long long b = (work_items + 255) / 256;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

template <typename T>
void launch_nb_upsample(const T* in, T* out, int N, int H, int W, int C, int f, cudaStream_t st) {
  const long long total = (long long)N * H * W * (C / Vec<T>::n);          // input vectors (< 2^32: 4 GB of bf16 input)
  count_launch();
  launch_pdl(nb_upsample_kernel<T>, dim3(grid_for(total, 148 * 8)), dim3(256), 0, st, in, out, H, W, C, f, total);
}
template <typename T>
void launch_nb_upsample_bwd(const T* dup, const T* a, int act_kind, T* dy, int N, int H, int W, int C, int f, cudaStream_t st) {
  const long long total = (long long)N * H * W * (C / Vec<T>::n);
  count_launch();
  launch_pdl(nb_upsample_bwd_kernel<T>, dim3(grid_for(total, 148 * 8)), dim3(256), 0, st, dup, a, act_kind, dy, H, W, C, f, total);
}
template <typename T>
void launch_nb_rsample(const NbSampleArgs& a, cudaStream_t st) {
  count_launch();
  launch_pdl(nb_rsample_kernel<T>, dim3(grid_for((long long)a.N * a.z * a.hw)), dim3(256), 0, st, a);
}
template <typename T>
void launch_nb_rsample_bwd(const NbSampleBwdArgs& a, cudaStream_t st) {
  count_launch();
  launch_pdl(nb_rsample_bwd_kernel<T>, dim3(grid_for((long long)a.N * a.z * a.hw, 296)), dim3(256), 0, st, a);
}
template <typename T>
void launch_nb_ce(const T* logits, const long long* target, T* dlogits, long long rows, int C, float scale, double* ce_acc,
                  float* dbias, cudaStream_t st) {
  long long b = (rows + 7) / 8;
  const int grid = (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
  count_launch();
  launch_pdl(nb_ce_kernel<T>, dim3(grid), dim3(256), 0, st, logits, target, dlogits, rows, C, scale, ce_acc, dbias);
}
template <typename T>
void launch_nb_colsum(const T* dy, long long rows, int C, float* dbias, cudaStream_t st) {
  const int rpb = 256 / (C / 8);
  long long b = (rows + rpb - 1) / rpb;
  const int grid = (int)(b < 1 ? 1 : (b > 148 * 2 ? 148 * 2 : b));
  count_launch();
  launch_pdl(nb_colsum_kernel<T>, dim3(grid), dim3(256), 0, st, dy, rows, C, dbias);
}
template <typename T>
void launch_nb_export_nchw(const T* in, float* out, int N, int HW, int C, cudaStream_t st) {
  const long long total = (long long)N * HW * C;
  count_launch();
  nb_export_nchw_kernel<T><<<grid_for(total, 148 * 32), 256, 0, st>>>(in, out, HW, C, total);
}
template <typename T>
void launch_nb_import_nchw(const float* in, T* out, int N, int HW, int C, cudaStream_t st) {
  const long long total = (long long)N * HW * C;
  count_launch();
  nb_import_nchw_kernel<T><<<grid_for(total, 148 * 32), 256, 0, st>>>(in, out, HW, C, total);
}
bool nb_stem_supported(int Ci, int Co, int S, int k, int s, int pad) {
  static const bool off = getenv("MMVAE_NO_NB_STEM") != nullptr;
  return !off && Ci == 1 && Co == 32 && k == 5 && s == 2 && pad == 2 && S % (2 * kStemRows) == 0 && S / 2 <= kStemMaxWo;
}
void launch_nb_stem_fwd(const float* x, const float* w, const float* bias, void* out, int N, int S, cudaStream_t st) {
  const int tiles = N * (S / 2 / kStemRows);
  count_launch();
  launch_pdl(nb_stem_fwd_kernel, dim3(tiles < 148 * 4 ? tiles : 148 * 4), dim3(256), 0, st, x, w, bias,
             reinterpret_cast<__nv_bfloat16*>(out), N, S);
}
void launch_nb_stem_wgrad(const float* x, const void* dy, float* dw, float* dbias, int N, int S, cudaStream_t st) {
  const int tiles = N * (S / 2 / kStemRows);
  count_launch();
  launch_pdl(nb_stem_wgrad_kernel, dim3(tiles < 148 * 4 ? tiles : 148 * 4), dim3(256), 0, st, x,
             reinterpret_cast<const __nv_bfloat16*>(dy), dw, dbias, N, S);
}
void launch_nb_loss_finalize(const double* acc, float inv_n, float klw, float* out, cudaStream_t st) {
  count_launch();
  launch_pdl(nb_loss_finalize_kernel, dim3(1), dim3(1), 0, st, acc, inv_n, klw, out);
}

#define NB_INST(T)                                                                                                          \
  template void launch_nb_upsample<T>(const T*, T*, int, int, int, int, int, cudaStream_t);                                 \
  template void launch_nb_upsample_bwd<T>(const T*, const T*, int, T*, int, int, int, int, int, cudaStream_t);              \
  template void launch_nb_rsample<T>(const NbSampleArgs&, cudaStream_t);                                                    \
  template void launch_nb_rsample_bwd<T>(const NbSampleBwdArgs&, cudaStream_t);                                             \
  template void launch_nb_ce<T>(const T*, const long long*, T*, long long, int, float, double*, float*, cudaStream_t);      \
  template void launch_nb_colsum<T>(const T*, long long, int, float*, cudaStream_t);                                        \
  template void launch_nb_export_nchw<T>(const T*, float*, int, int, int, cudaStream_t);                                    \
  template void launch_nb_import_nchw<T>(const float*, T*, int, int, int, cudaStream_t);
NB_INST(float)
NB_INST(__nv_bfloat16)

}  // namespace mmvae
