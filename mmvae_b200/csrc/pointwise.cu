// pointwise.cu -- BatchNorm statistics / apply / backward, encoder heads + rsample, layout helpers.
//
// BatchNorm2d training semantics (reference constructs nn.BatchNorm2d with defaults at
// model.py:30,34,61,66,95,137,162,173,202): biased batch variance for normalisation, running
// buffers updated with momentum 0.1 and the unbiased variance, num_batches_tracked += 1.
// The per-channel sums come from the producing conv's epilogue as per-CTA partials and are
// combined here in a fixed order in fp64, so a step is bit-reproducible.
#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "bn_fused.cuh"
#include "kernels.cuh"

namespace mmvae {

namespace {

constexpr float kBnEps = 1e-5f;
constexpr float kBnMomentum = 0.1f;

template <int NT>
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int tid = threadIdx.x;
  sh[tid] = v;
  __syncthreads();
#pragma unroll
  for (int s = NT / 2; s > 0; s >>= 1) {
    if (tid < s) sh[tid] += sh[tid + s];
    __syncthreads();
  }
  double r = sh[0];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(128) bn_finalize_kernel(const BnFinalizeArgs a) {
  __shared__ double sh[128];
  const int c = blockIdx.x;
  float mean, var;
  if (a.training) {
    // Chan et al. pairwise combination of per-CTA (n_i, sum_i, M2_i), in fp64 and in a fixed order.
    double s1 = 0.0;
    for (int i = threadIdx.x; i < a.sl.parts; i += 128) s1 += (double)a.partials[(size_t(i) * a.C + c) * 2 + 0];
    s1 = block_sum<128>(s1, sh);
    const double mu = s1 / (double)a.m;
    double s2 = 0.0;
    for (int i = threadIdx.x; i < a.sl.parts; i += 128) {
      int row0 = (i % a.sl.parts_per_var) * a.sl.tile_rows;
      double ni = a.sl.counts ? (double)a.sl.counts[i] : (double)min(a.sl.tile_rows, a.sl.rows_per_var - row0);
      if (ni <= 0.0) continue;
      const double si = (double)a.partials[(size_t(i) * a.C + c) * 2 + 0];
      double m2i = (double)a.partials[(size_t(i) * a.C + c) * 2 + 1];
      if (a.sl.sumsq) m2i = fmax(m2i - si * si / ni, 0.0);      // row carries sum of squares: M2 = sum x^2 - (sum x)^2 / n
      double mi = si / ni - mu;
      s2 += m2i + ni * mi * mi;
    }
    s2 = block_sum<128>(s2, sh);
    double vv = s2 / (double)a.m;
    mean = (float)mu; var = (float)vv;
    if (threadIdx.x == 0 && a.running_mean) {
      double unb = a.m > 1 ? vv * ((double)a.m / (double)(a.m - 1)) : vv;
      a.running_mean[c] = (1.f - kBnMomentum) * a.running_mean[c] + kBnMomentum * mean;
      a.running_var[c] = (1.f - kBnMomentum) * a.running_var[c] + kBnMomentum * (float)unb;
      if (c == 0 && a.counter) *a.counter += 1;
    }
  } else {
    mean = a.running_mean[c]; var = a.running_var[c];
  }
  if (threadIdx.x == 0) {
    float rstd = 1.0f / sqrtf(var + kBnEps);
    a.stat[c] = mean; a.stat[a.C + c] = rstd;
    float scale = a.gamma[c] * rstd;
    a.coef[c] = scale; a.coef[a.C + c] = a.beta[c] - mean * scale;
  }
}

template <typename T, int VEC> struct Vec;
template <> struct Vec<float, 4> { using type = float4; };
template <> struct Vec<float, 1> { using type = float; };
template <> struct Vec<__nv_bfloat16, 8> { using type = uint4; };
template <> struct Vec<__nv_bfloat16, 1> { using type = __nv_bfloat16; };

template <typename T, int VEC>
__device__ __forceinline__ void load_vec(const T* p, float (&f)[VEC]) {
  typename Vec<T, VEC>::type raw = *reinterpret_cast<const typename Vec<T, VEC>::type*>(p);
  const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = to_f(e[i]);
}
template <typename T, int VEC>
__device__ __forceinline__ void store_vec(T* p, const float (&f)[VEC]) {
  typename Vec<T, VEC>::type raw;
  T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
  for (int i = 0; i < VEC; ++i) e[i] = from_f<T>(f[i]);
  *reinterpret_cast<typename Vec<T, VEC>::type*>(p) = raw;
}

// raw 16-byte vector of the storage type (kept packed in registers until it is used)
template <typename T, int VEC> struct RawVec { typename Vec<T, VEC>::type r; };
template <typename T, int VEC>
__device__ __forceinline__ void unpack_vec(const RawVec<T, VEC>& raw, float (&f)[VEC]) {
  const T* e = reinterpret_cast<const T*>(&raw.r);
#pragma unroll
  for (int i = 0; i < VEC; ++i) f[i] = to_f(e[i]);
}
template <typename T, int VEC>
__device__ __forceinline__ RawVec<T, VEC> load_raw(const T* p) {
  RawVec<T, VEC> v;
  v.r = *reinterpret_cast<const typename Vec<T, VEC>::type*>(p);
  return v;
}

// incoming gradient: storage type, or fp32 (the user-facing d_recon)
template <typename T, int VEC>
__device__ __forceinline__ void load_grad(const void* dA, int is_f32, size_t off, float (&g)[VEC]) {
  if (is_f32) {
    const float* p = reinterpret_cast<const float*>(dA) + off;
#pragma unroll
    for (int k = 0; k < VEC; ++k) g[k] = p[k];
  } else {
    load_vec<T, VEC>(reinterpret_cast<const T*>(dA) + off, g);
  }
}

// VEC consecutive per-channel coefficients (c0 % VEC == 0, arrays 16-byte aligned): 128-bit loads instead of VEC scalar ones
template <int VEC>
__device__ __forceinline__ void load_coef(const float* __restrict__ p, float (&v)[VEC]) {
  if constexpr (VEC % 4 == 0) {
#pragma unroll
    for (int k = 0; k < VEC; k += 4) {
      const float4 t = __ldg(reinterpret_cast<const float4*>(p + k));
      v[k] = t.x; v[k + 1] = t.y; v[k + 2] = t.z; v[k + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < VEC; ++k) v[k] = __ldg(p + k);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T* __restrict__ y, const float* __restrict__ coef,
                                                       const T* __restrict__ y2, const float* __restrict__ coef2,
                                                       T* __restrict__ out, long long nvec, int C, int relu) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
    int c0 = (int)((i * VEC) % C);
    float v[VEC], r[VEC];
    float sc[VEC], sh[VEC];
    load_vec<T, VEC>(y + i * VEC, v);
    load_coef<VEC>(coef + c0, sc); load_coef<VEC>(coef + C + c0, sh);
#pragma unroll
    for (int k = 0; k < VEC; ++k) r[k] = fmaf(v[k], sc[k], sh[k]);
    if (y2) {
      load_vec<T, VEC>(y2 + i * VEC, v);
      load_coef<VEC>(coef2 + c0, sc); load_coef<VEC>(coef2 + C + c0, sh);
#pragma unroll
      for (int k = 0; k < VEC; ++k) r[k] += fmaf(v[k], sc[k], sh[k]);
    }
    if (relu) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) r[k] = fmaxf(r[k], 0.f);
    }
    store_vec<T, VEC>(out + i * VEC, r);
  }
}

// one output channel (NHWC == NCHW): 8 values per thread, 16-byte load, two 16-byte stores
__global__ void __launch_bounds__(256) bn_apply_out_c1_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ coef,
                                                              float* __restrict__ out, long long nvec) {
  pdl_wait();
  pdl_trigger();
  const float sc = __ldg(coef), sh = __ldg(coef + 1);
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
    float v[8];
    load_vec<__nv_bfloat16, 8>(y + i * 8, v);
    float4 a = make_float4(fmaf(v[0], sc, sh), fmaf(v[1], sc, sh), fmaf(v[2], sc, sh), fmaf(v[3], sc, sh));
    float4 b = make_float4(fmaf(v[4], sc, sh), fmaf(v[5], sc, sh), fmaf(v[6], sc, sh), fmaf(v[7], sc, sh));
    reinterpret_cast<float4*>(out)[2 * i] = a;
    reinterpret_cast<float4*>(out)[2 * i + 1] = b;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_apply_out_kernel(const T* __restrict__ y, const float* __restrict__ coef,
                                                           float* __restrict__ out, long long total, int HW, int C) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    int hw = (int)(i % HW); long long t = i / HW; int c = (int)(t % C); long long n = t / C;
    float v = to_f(y[(n * HW + hw) * C + c]);
    out[i] = fmaf(v, __ldg(coef + c), __ldg(coef + C + c));
  }
}

// ---------------- encoder heads: avg-pool + two 1x1 convs + rsample (model.py:123-128,148-150) ----------
// grid (N, kHeadsSplit): CTA (n, q) pools frame n and produces latent channels [q * zq, (q+1) * zq) of mu and logvar
constexpr int kHeadsSplit = 8;
template <typename T>
__global__ void __launch_bounds__(128) heads_fwd_kernel(const HeadsArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  extern __shared__ float sm[];          // pooled[C] then out[2 * zq]
  const int zq = (a.z + kHeadsSplit - 1) / kHeadsSplit;
  const int z0 = blockIdx.y * zq, z1 = min(a.z, z0 + zq);
  float* pooled = sm;
  float* outv = sm + a.C;
  const int n = blockIdx.x, tid = threadIdx.x;
  const T* feat = reinterpret_cast<const T*>(a.feat) + size_t(n) * a.hw * a.C;
  const float inv = 1.0f / (float)a.hw;
  for (int c = tid; c < a.C; c += 128) {
    float s = 0.f;
    for (int p = 0; p < a.hw; ++p) s += to_f(feat[size_t(p) * a.C + c]);
    s *= inv;
    pooled[c] = s;
    if (blockIdx.y == 0) a.pooled[size_t(n) * a.C + c] = s;
  }
  __syncthreads();
  const int nz = z1 - z0;
  const int nout = a.w_lv ? 2 * nz : nz;
  const int warp = tid >> 5, lane = tid & 31;
  for (int o = warp; o < nout; o += 4) {
    const bool is_lv = o >= nz;
    const int zc = z0 + (is_lv ? o - nz : o);
    const float* w = (is_lv ? a.w_lv : a.w_mu) + size_t(zc) * a.C;
    float s = 0.f;
    for (int c = lane; c < a.C; c += 32) s = fmaf(__ldg(w + c), pooled[c], s);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) outv[o] = s;
  }
  __syncthreads();
  const size_t NZ = size_t(a.N) * a.z;
  for (int t = tid; t < nz; t += 128) {
    const int zc = z0 + t;
    size_t idx = size_t(n) * a.z + zc;
    float mu = outv[t];
    float lv = 0.f, eps = 0.f, sd = 0.f, enc = mu;
    if (a.w_lv) {
      lv = outv[nz + t];
      eps = a.eps ? a.eps[idx] : philox_normal_at(a.rng_dev ? a.rng_dev[0] : a.seed, a.rng_dev ? a.rng_dev[1] : a.offset, (long long)idx);
      sd = expf(0.5f * lv);
      enc = fmaf(eps, sd, mu);
    }
    a.heads[idx] = mu; a.heads[NZ + idx] = lv; a.heads[2 * NZ + idx] = eps; a.heads[3 * NZ + idx] = sd;
    a.mu_out[idx] = mu;
    if (a.lv_out) a.lv_out[idx] = lv;
    a.enc_out[idx] = enc;
    if (a.eps_out) a.eps_out[idx] = eps;
    reinterpret_cast<T*>(a.z_act)[idx] = from_f<T>(enc);
  }
  if (a.rng_adv) {
    // device-resident generator state (a captured graph draws fresh noise per replay): every CTA has read the offset
    // above; the one that finishes last advances it for the next forward
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(a.rng_ticket, 1u) == gridDim.x * gridDim.y - 1u) { a.rng_adv[1] += a.rng_inc; *a.rng_ticket = 0u; }
    }
  }
}

template <typename T>
__global__ void cast_latent_kernel(const float* __restrict__ in, T* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) out[i] = from_f<T>(in[i]);
}

// dz -> (dmu, dlogvar) -> dpooled -> d(encoder output)
template <typename T>
__global__ void __launch_bounds__(128) heads_bwd_kernel(const HeadsBwdArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  extern __shared__ float sm[];          // dmu[z], dlv[z]
  float* dmu = sm; float* dlv = sm + a.z;
  const int n = blockIdx.x, tid = threadIdx.x;
  const size_t NZ = size_t(a.N) * a.z;
  for (int zc = tid; zc < a.z; zc += 128) {
    size_t idx = size_t(n) * a.z + zc;
    float dz = 0.f;
    if (a.dz_act) dz += to_f(reinterpret_cast<const T*>(a.dz_act)[idx]);
    if (a.d_enc) dz += a.d_enc[idx];
    float gm = dz + (a.d_mu ? a.d_mu[idx] : 0.f);
    float gl = a.d_lv ? a.d_lv[idx] : 0.f;
    if (a.w_lv) gl += dz * 0.5f * a.heads[2 * NZ + idx] * a.heads[3 * NZ + idx];   // dz * 0.5 * eps * std
    dmu[zc] = gm; dlv[zc] = gl;
    if (blockIdx.y == 0) { a.dheads[idx] = gm; a.dheads[NZ + idx] = gl; }
  }
  __syncthreads();
  T* dfeat = reinterpret_cast<T*>(a.dfeat) + size_t(n) * a.hw * a.C;
  const float inv = 1.0f / (float)a.hw;
  for (int c = blockIdx.y * 128 + tid; c < a.C; c += 128 * gridDim.y) {      // grid.y CTAs share a frame's channels
    float s = 0.f;
    for (int zc = 0; zc < a.z; ++zc) {
      s = fmaf(dmu[zc], __ldg(a.w_mu + size_t(zc) * a.C + c), s);
      if (a.w_lv) s = fmaf(dlv[zc], __ldg(a.w_lv + size_t(zc) * a.C + c), s);
    }
    a.dpool[size_t(n) * a.C + c] = s;
    T v = from_f<T>(s * inv);
    for (int p = 0; p < a.hw; ++p) dfeat[size_t(p) * a.C + c] = v;
  }
}

// ---------------- BatchNorm backward ----------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnBwdArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ float red[256 * VEC * 3];
  const int CV = a.C / VEC;
  const int RPI = 256 / CV;              // rows per iteration (CV <= 256 guaranteed by the launcher)
  const int tid = threadIdx.x;
  const int cv = tid % CV, rsub = tid / CV;
  const bool active = rsub < RPI;
  float s0[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s0[k] = s1[k] = s2[k] = 0.f; }
  const int c0 = cv * VEC;
  float mean[VEC], rstd[VEC], mean2[VEC], rstd2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    mean[k] = a.stat[c0 + k]; rstd[k] = a.stat[a.C + c0 + k];
    mean2[k] = a.y2 ? a.stat2[c0 + k] : 0.f; rstd2[k] = a.y2 ? a.stat2[a.C + c0 + k] : 0.f;
  }
  if (active) {
#pragma unroll 4
    for (long long r = (long long)blockIdx.x * RPI + rsub; r < a.rows; r += (long long)gridDim.x * RPI) {
      size_t off = size_t(r) * a.C + c0;
      float g[VEC], yv[VEC];
      load_grad<T, VEC>(a.dA, a.dA_f32, off, g);
      if (a.dA2) {                                   // the gradient arrives in two parts (main branch, shortcut branch)
        float g2[VEC];
        load_vec<T, VEC>(reinterpret_cast<const T*>(a.dA2) + off, g2);
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = to_f(from_f<T>(g[k] + g2[k]));
      }
      if (a.a) {
        float av[VEC];
        load_vec<T, VEC>(reinterpret_cast<const T*>(a.a) + off, av);
#pragma unroll
        for (int k = 0; k < VEC; ++k) g[k] = av[k] > 0.f ? g[k] : 0.f;
      }
      load_vec<T, VEC>(reinterpret_cast<const T*>(a.y) + off, yv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) { s0[k] += g[k]; s1[k] = fmaf(g[k], (yv[k] - mean[k]) * rstd[k], s1[k]); }
      if (a.y2) {
        load_vec<T, VEC>(reinterpret_cast<const T*>(a.y2) + off, yv);
#pragma unroll
        for (int k = 0; k < VEC; ++k) s2[k] = fmaf(g[k], (yv[k] - mean2[k]) * rstd2[k], s2[k]);
      }
    }
  }
  // red[which][rsub][c]
  if (active) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      red[(0 * RPI + rsub) * a.C + c0 + k] = s0[k];
      red[(1 * RPI + rsub) * a.C + c0 + k] = s1[k];
      red[(2 * RPI + rsub) * a.C + c0 + k] = s2[k];
    }
  }
  __syncthreads();
  for (int e = tid; e < 3 * a.C; e += 256) {
    int which = e / a.C, c = e % a.C;
    float s = 0.f;
    for (int q = 0; q < RPI; ++q) s += red[(which * RPI + q) * a.C + c];
    if (a.acc) atomicAdd(a.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 * a.C + which * a.C + c, (double)s);
    else a.partials[(size_t(blockIdx.x) * a.C + c) * 3 + which] = s;
  }
  if (!a.acc) return;
  // fused finalize: the block that finishes last turns the fp64 sums into the backward coefficients
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double im = 1.0 / (double)a.rows;
  for (int c = tid; c < a.C; c += 256) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {
      const double* ak = a.acc + (size_t)k * 3 * a.C;
      s0 += ld_cg_f64(ak + c); s1 += ld_cg_f64(ak + a.C + c); s2 += ld_cg_f64(ak + 2 * a.C + c);
    }
    a.g_beta[c] = (float)s0; a.g_gamma[c] = (float)s1;
    a.bcoef[c] = a.gamma[c] * a.stat[a.C + c];
    a.bcoef[a.C + c] = (float)(s0 * im);
    a.bcoef[2 * a.C + c] = (float)(s1 * im);
    if (a.y2) {
      a.g_beta2[c] = (float)s0; a.g_gamma2[c] = (float)s2;
      a.bcoef2[c] = a.gamma2[c] * a.stat2[a.C + c];
      a.bcoef2[a.C + c] = (float)(s0 * im);
      a.bcoef2[2 * a.C + c] = (float)(s2 * im);
    }
  }
}

__global__ void __launch_bounds__(128) bn_bwd_finalize_kernel(const BnBwdArgs a, int nblocks, long long m) {
  __shared__ double sh[128];
  const int c = blockIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 128) {
    const float* p = a.partials + (size_t(i) * a.C + c) * 3;
    s0 += (double)p[0]; s1 += (double)p[1]; s2 += (double)p[2];
  }
  s0 = block_sum<128>(s0, sh);
  s1 = block_sum<128>(s1, sh);
  s2 = block_sum<128>(s2, sh);
  if (threadIdx.x == 0) {
    const double im = 1.0 / (double)m;
    a.g_beta[c] = (float)s0; a.g_gamma[c] = (float)s1;
    a.bcoef[c] = a.gamma[c] * a.stat[a.C + c];
    a.bcoef[a.C + c] = (float)(s0 * im);
    a.bcoef[2 * a.C + c] = (float)(s1 * im);
    if (a.y2) {
      a.g_beta2[c] = (float)s0; a.g_gamma2[c] = (float)s2;
      a.bcoef2[c] = a.gamma2[c] * a.stat2[a.C + c];
      a.bcoef2[a.C + c] = (float)(s0 * im);
      a.bcoef2[2 * a.C + c] = (float)(s2 * im);
    }
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnBwdArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  const long long nvec = a.rows * a.C / VEC;
  if (a.reduced && blockIdx.x == 0) {
    // the reduction ran in the producer of dA (BnBwdFused): publish d beta = S0, d gamma = S1 from its sums
    for (int c = threadIdx.x; c < a.C; c += 256) {
      a.g_beta[c] = a.bcoef[3 * a.C + c]; a.g_gamma[c] = a.bcoef[4 * a.C + c];
      if (a.y2) { a.g_beta2[c] = a.bcoef2[3 * a.C + c]; a.g_gamma2[c] = a.bcoef2[4 * a.C + c]; }
    }
  }
#pragma unroll 2
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < nvec; i += gridDim.x * 256LL) {
    const int c0 = (int)((i * VEC) % a.C);
    const size_t off = size_t(i) * VEC;
    float g[VEC], yv[VEC], o[VEC];
    load_grad<T, VEC>(a.dA, a.dA_f32, off, g);
    if (a.dA2) {
      float g2[VEC];
      load_vec<T, VEC>(reinterpret_cast<const T*>(a.dA2) + off, g2);
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[k] = to_f(from_f<T>(g[k] + g2[k]));
    }
    if (a.a) {
      float av[VEC];
      load_vec<T, VEC>(reinterpret_cast<const T*>(a.a) + off, av);
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[k] = av[k] > 0.f ? g[k] : 0.f;
    }
    load_vec<T, VEC>(reinterpret_cast<const T*>(a.y) + off, yv);
    {
      float mean[VEC], rstd[VEC], b0[VEC], b1[VEC], b2[VEC];
      load_coef<VEC>(a.stat + c0, mean); load_coef<VEC>(a.stat + a.C + c0, rstd);
      load_coef<VEC>(a.bcoef + c0, b0); load_coef<VEC>(a.bcoef + a.C + c0, b1); load_coef<VEC>(a.bcoef + 2 * a.C + c0, b2);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float xh = (yv[k] - mean[k]) * rstd[k];
        o[k] = b0[k] * (g[k] - b1[k] - xh * b2[k]);
      }
    }
    store_vec<T, VEC>(reinterpret_cast<T*>(a.dY) + off, o);
    if (a.y2) {
      load_vec<T, VEC>(reinterpret_cast<const T*>(a.y2) + off, yv);
      float mean[VEC], rstd[VEC], b0[VEC], b1[VEC], b2[VEC];
      load_coef<VEC>(a.stat2 + c0, mean); load_coef<VEC>(a.stat2 + a.C + c0, rstd);
      load_coef<VEC>(a.bcoef2 + c0, b0); load_coef<VEC>(a.bcoef2 + a.C + c0, b1); load_coef<VEC>(a.bcoef2 + 2 * a.C + c0, b2);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        float xh = (yv[k] - mean[k]) * rstd[k];
        o[k] = b0[k] * (g[k] - b1[k] - xh * b2[k]);
      }
      store_vec<T, VEC>(reinterpret_cast<T*>(a.dY2) + off, o);
    }
  }
}

// BatchNorm backward of a small / medium tensor in ONE launch: every thread keeps its (at most kCoopE) 16-byte vectors
// of g = dA * [a > 0], y (and y2) packed in registers across a grid-wide barrier -- phase 1 reduces the three per-channel
// sums into the fp64 accumulators, the barrier is an atomic counter every CTA spins on (the grid is at most one CTA per
// SM and needs little of it, so all CTAs are co-resident eventually whatever else runs), phase 2 turns the sums into the
// coefficients and writes dY (dY2) from the registers.  Replaces bn_bwd_reduce + bn_bwd_apply: one launch boundary and
// one pass over the inputs less.
constexpr int kCoopE = 4;
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__global__ void __launch_bounds__(256, 2) bn_bwd_coop_kernel(const BnBwdArgs a, int E) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  typedef __nv_bfloat16 T;
  constexpr int VEC = 8;
  __shared__ float red[256 * VEC * 3];
  __shared__ float coef[2][5][256];                  // per branch: mean, rstd, scale, c1, c2
  const int CV = a.C / VEC, RPI = 256 / CV;
  const int tid = threadIdx.x;
  const int cv = tid % CV, rsub = tid / CV, c0 = cv * VEC;
  const T* dA = reinterpret_cast<const T*>(a.dA);
  const T* am = reinterpret_cast<const T*>(a.a);
  const T* y1 = reinterpret_cast<const T*>(a.y);
  const T* y2 = reinterpret_cast<const T*>(a.y2);
  float mean[VEC], rstd[VEC], mean2[VEC], rstd2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) {
    mean[k] = a.stat[c0 + k]; rstd[k] = a.stat[a.C + c0 + k];
    mean2[k] = y2 ? a.stat2[c0 + k] : 0.f; rstd2[k] = y2 ? a.stat2[a.C + c0 + k] : 0.f;
  }
  RawVec<T, VEC> rg[kCoopE], ry[kCoopE], rz[kCoopE];
  float s0[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s0[k] = s1[k] = s2[k] = 0.f; }
  const long long rstride = (long long)gridDim.x * RPI;
  const long long r00 = (long long)blockIdx.x * RPI + rsub;
  // ---- phase 1: loads (all issued up front), mask, sums ----
#pragma unroll
  for (int e = 0; e < kCoopE; ++e) {
    const long long r = r00 + e * rstride;
    if (e < E && r < a.rows) {
      const size_t off = size_t(r) * a.C + c0;
      rg[e] = load_raw<T, VEC>(dA + off);
      ry[e] = load_raw<T, VEC>(y1 + off);
      if (y2) rz[e] = load_raw<T, VEC>(y2 + off);
      if (a.dA2) {                                   // the gradient arrives in two parts (main branch, shortcut branch)
        const RawVec<T, VEC> r2 = load_raw<T, VEC>(reinterpret_cast<const T*>(a.dA2) + off);
        const T* g2 = reinterpret_cast<const T*>(&r2.r);
        T* ge = reinterpret_cast<T*>(&rg[e].r);
#pragma unroll
        for (int k = 0; k < VEC; ++k) ge[k] = from_f<T>(to_f(ge[k]) + to_f(g2[k]));
      }
      if (am) {
        const RawVec<T, VEC> ra = load_raw<T, VEC>(am + off);
        const T* ae = reinterpret_cast<const T*>(&ra.r);
        T* ge = reinterpret_cast<T*>(&rg[e].r);
#pragma unroll
        for (int k = 0; k < VEC; ++k) if (!(to_f(ae[k]) > 0.f)) ge[k] = from_f<T>(0.f);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < kCoopE; ++e) {
    const long long r = r00 + e * rstride;
    if (e < E && r < a.rows) {
      float g[VEC], yv[VEC];
      unpack_vec<T, VEC>(rg[e], g);
      unpack_vec<T, VEC>(ry[e], yv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) { s0[k] += g[k]; s1[k] = fmaf(g[k], (yv[k] - mean[k]) * rstd[k], s1[k]); }
      if (y2) {
        unpack_vec<T, VEC>(rz[e], yv);
#pragma unroll
        for (int k = 0; k < VEC; ++k) s2[k] = fmaf(g[k], (yv[k] - mean2[k]) * rstd2[k], s2[k]);
      }
    }
  }
  // lanes of a warp with the same channel vector (lane % CV; CV is a power of two <= 32) first add up by shuffles, so that
  // the shared-memory stage sums 8 warps instead of RPI = 256 / CV row groups serially (64 dependent loads at C = 32)
  for (int d = CV; d < 32; d <<= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], d);
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], d);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], d);
    }
  }
  const int wrp = tid >> 5;
  if ((tid & 31) < CV) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      red[(0 * 8 + wrp) * a.C + c0 + k] = s0[k];
      red[(1 * 8 + wrp) * a.C + c0 + k] = s1[k];
      red[(2 * 8 + wrp) * a.C + c0 + k] = s2[k];
    }
  }
  __syncthreads();
  for (int e = tid; e < 3 * a.C; e += 256) {
    const int which = e / a.C, c = e % a.C;
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) sum += red[(which * 8 + q) * a.C + c];
    atomicAdd(a.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 * a.C + which * a.C + c, (double)sum);
  }
  // ---- grid-wide barrier ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(a.counter, 1u);
    while (ld_acquire_u32(a.counter) < gridDim.x) { __nanosleep(32); }
  }
  __syncthreads();
  // ---- phase 2: coefficients (one channel per thread), then dY from the registers ----
  const double im = 1.0 / (double)a.rows;
  for (int c = tid; c < a.C; c += 256) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {
      const double* ak = a.acc + (size_t)k * 3 * a.C;
      t0 += ld_cg_f64(ak + c); t1 += ld_cg_f64(ak + a.C + c); t2 += ld_cg_f64(ak + 2 * a.C + c);
    }
    coef[0][0][c] = a.stat[c]; coef[0][1][c] = a.stat[a.C + c];
    coef[0][2][c] = a.gamma[c] * a.stat[a.C + c]; coef[0][3][c] = (float)(t0 * im); coef[0][4][c] = (float)(t1 * im);
    if (y2) {
      coef[1][0][c] = a.stat2[c]; coef[1][1][c] = a.stat2[a.C + c];
      coef[1][2][c] = a.gamma2[c] * a.stat2[a.C + c]; coef[1][3][c] = (float)(t0 * im); coef[1][4][c] = (float)(t2 * im);
    }
    if (blockIdx.x == 0) {
      a.g_beta[c] = (float)t0; a.g_gamma[c] = (float)t1;
      if (y2) { a.g_beta2[c] = (float)t0; a.g_gamma2[c] = (float)t2; }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < kCoopE; ++e) {
    const long long r = r00 + e * rstride;
    if (e < E && r < a.rows) {
      const size_t off = size_t(r) * a.C + c0;
      float g[VEC], yv[VEC], o[VEC];
      unpack_vec<T, VEC>(rg[e], g);
      unpack_vec<T, VEC>(ry[e], yv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int c = c0 + k;
        const float xh = (yv[k] - coef[0][0][c]) * coef[0][1][c];
        o[k] = coef[0][2][c] * (g[k] - coef[0][3][c] - xh * coef[0][4][c]);
      }
      store_vec<T, VEC>(reinterpret_cast<T*>(a.dY) + off, o);
      if (y2) {
        unpack_vec<T, VEC>(rz[e], yv);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const int c = c0 + k;
          const float xh = (yv[k] - coef[1][0][c]) * coef[1][1][c];
          o[k] = coef[1][2][c] * (g[k] - coef[1][3][c] - xh * coef[1][4][c]);
        }
        store_vec<T, VEC>(reinterpret_cast<T*>(a.dY2) + off, o);
      }
    }
  }
}

// BatchNorm backward in ONE launch with a SMALL footprint (bf16 storage): a persistent grid of at most two CTAs per SM
// walks its rows twice -- pass 1 reduces the three per-channel sums (S1 is accumulated as sum g*y and centred at the end:
// S1 = rstd * (sum g*y - mean * S0)) into the fp64 accumulators, a grid-wide barrier follows, pass 2 re-reads the same
// rows (they are L2-resident: the whole working set of a layer is a few MB of the 126 MB L2; the big decoder tensors
// re-stream, as the two-kernel path did) and writes dY (dY2).  Unlike bn_bwd_coop_kernel nothing is carried in registers
// across the barrier: ~48 registers and 7 KB of shared memory per CTA, so the grid is co-resident BESIDE the weight-gradient
// kernels of the auxiliary stream and the tail of the previous kernel (the register-resident version needs a whole SM's
// register file per CTA pair and, measured, waited 8-15 us for the SMs to drain: profiles/r02_bn_bwd.md).
__global__ void __launch_bounds__(256, 4) bn_bwd_sweep_kernel(const BnBwdArgs a) {
  typedef __nv_bfloat16 T;
  constexpr int VEC = 8;
  extern __shared__ float red[];                       // [3 sums][8 warps][C]: combined in a FIXED order (run-to-run
                                                       // reproducible: a 1e-7 wobble here is amplified to 1e-2 at the stem by
                                                       // the bf16 rounding of the ~25 gradient tensors downstream)
  __shared__ float coef[2][3][256];                    // per branch: scale, c1 (mean of g), c2 (mean of g * xhat)
  __shared__ float mr[2][2][256];                      // per branch: mean, rstd
  const int CV = a.C / VEC, RPI = 256 / CV;
  const int tid = threadIdx.x;
  const int cv = tid % CV, rsub = tid / CV, c0 = cv * VEC;
  const T* dA = reinterpret_cast<const T*>(a.dA);
  const T* dA2 = reinterpret_cast<const T*>(a.dA2);
  const T* am = reinterpret_cast<const T*>(a.a);
  const T* y1 = reinterpret_cast<const T*>(a.y);
  const T* y2 = reinterpret_cast<const T*>(a.y2);
  for (int c = tid; c < a.C; c += 256) {
    mr[0][0][c] = a.stat[c]; mr[0][1][c] = a.stat[a.C + c];
    if (y2) { mr[1][0][c] = a.stat2[c]; mr[1][1][c] = a.stat2[a.C + c]; }
  }
  const long long rstride = (long long)gridDim.x * RPI;
  const long long r0 = (long long)blockIdx.x * RPI + rsub;
  // The ReLU output and the raw conv outputs were written by the FORWARD and are cold by now: their rows are pulled into L2
  // under the predecessor's tail, before the dependency wait; only the incoming gradient is the predecessor's.
  for (long long r = r0; r < a.rows && !a.late_loads; r += rstride) {
    const size_t off = size_t(r) * a.C + c0;
    if (am) asm volatile("prefetch.global.L2 [%0];" ::"l"(am + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(y1 + off));
    if (y2) asm volatile("prefetch.global.L2 [%0];" ::"l"(y2 + off));
  }
  pdl_wait();
  pdl_trigger();
  // g = (dA [+ dA2]) * [a > 0], rounded to the storage type once (what the accumulate-in-place path stored)
  auto load_g = [&](size_t off, float (&g)[VEC]) {
    load_vec<T, VEC>(dA + off, g);
    if (dA2) {
      float g2[VEC];
      load_vec<T, VEC>(dA2 + off, g2);
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[k] = to_f(from_f<T>(g[k] + g2[k]));
    }
    if (am) {
      float av[VEC];
      load_vec<T, VEC>(am + off, av);
#pragma unroll
      for (int k = 0; k < VEC; ++k) g[k] = av[k] > 0.f ? g[k] : 0.f;
    }
  };
  // ---- pass 1 ----
  float s0[VEC], s1[VEC], s2[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) { s0[k] = s1[k] = s2[k] = 0.f; }
#pragma unroll 1
  for (long long r = r0; r < a.rows; r += rstride) {
    const size_t off = size_t(r) * a.C + c0;
    float g[VEC], yv[VEC];
    load_g(off, g);
    load_vec<T, VEC>(y1 + off, yv);
#pragma unroll
    for (int k = 0; k < VEC; ++k) { s0[k] += g[k]; s1[k] = fmaf(g[k], yv[k], s1[k]); }
    if (y2) {
      load_vec<T, VEC>(y2 + off, yv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) s2[k] = fmaf(g[k], yv[k], s2[k]);
    }
  }
  for (int d = CV; d < 32; d <<= 1) {
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], d);
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], d);
      s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], d);
    }
  }
  const int wrp = tid >> 5;
  if ((tid & 31) < CV) {                               // lanes < CV of every warp hold the warp's sums of their channel vector
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      red[(0 * 8 + wrp) * a.C + c0 + k] = s0[k]; red[(1 * 8 + wrp) * a.C + c0 + k] = s1[k];
      red[(2 * 8 + wrp) * a.C + c0 + k] = s2[k];
    }
  }
  __syncthreads();
  for (int e = tid; e < 3 * a.C; e += 256) {
    const int which = e / a.C, c = e - which * a.C;
    float sum = 0.f, t0 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { sum += red[(which * 8 + q) * a.C + c]; t0 += red[q * a.C + c]; }
    // centre the raw moments: sum g * xhat = rstd * (sum g*y - mean * sum g), per CTA in fp32 (|mean * S0| ~ |sum g*y|
    // only when the channel mean dominates its spread; the fp64 accumulation across CTAs keeps the rest exact)
    if (which == 1 || (which == 2 && y2)) sum = mr[which - 1][1][c] * (sum - mr[which - 1][0][c] * t0);
    if (which < 2 || y2) atomicAdd(a.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 * a.C + which * a.C + c, (double)sum);
  }
  // ---- grid-wide barrier (the grid is co-resident: at most coop_max_ctas() CTAs) ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(a.counter, 1u);
    while (ld_acquire_u32(a.counter) < gridDim.x) { __nanosleep(20); }
  }
  __syncthreads();
  // ---- coefficients ----
  const double im = 1.0 / (double)a.rows;
  for (int c = tid; c < a.C; c += 256) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {
      const double* ak = a.acc + (size_t)k * 3 * a.C;
      t0 += ld_cg_f64(ak + c); t1 += ld_cg_f64(ak + a.C + c);
      if (y2) t2 += ld_cg_f64(ak + 2 * a.C + c);
    }
    coef[0][0][c] = a.gamma[c] * mr[0][1][c]; coef[0][1][c] = (float)(t0 * im); coef[0][2][c] = (float)(t1 * im);
    if (y2) { coef[1][0][c] = a.gamma2[c] * mr[1][1][c]; coef[1][1][c] = (float)(t0 * im); coef[1][2][c] = (float)(t2 * im); }
    if (blockIdx.x == 0) {
      a.g_beta[c] = (float)t0; a.g_gamma[c] = (float)t1;
      if (y2) { a.g_beta2[c] = (float)t0; a.g_gamma2[c] = (float)t2; }
    }
  }
  __syncthreads();
  // ---- pass 2 ----
#pragma unroll 1
  for (long long r = r0; r < a.rows; r += rstride) {
    const size_t off = size_t(r) * a.C + c0;
    float g[VEC], yv[VEC], o[VEC];
    load_g(off, g);
    load_vec<T, VEC>(y1 + off, yv);
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      const int c = c0 + k;
      const float xh = (yv[k] - mr[0][0][c]) * mr[0][1][c];
      o[k] = coef[0][0][c] * (g[k] - coef[0][1][c] - xh * coef[0][2][c]);
    }
    store_vec<T, VEC>(reinterpret_cast<T*>(a.dY) + off, o);
    if (y2) {
      load_vec<T, VEC>(y2 + off, yv);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const int c = c0 + k;
        const float xh = (yv[k] - mr[1][0][c]) * mr[1][1][c];
        o[k] = coef[1][0][c] * (g[k] - coef[1][1][c] - xh * coef[1][2][c]);
      }
      store_vec<T, VEC>(reinterpret_cast<T*>(a.dY2) + off, o);
    }
  }
}

// BatchNorm backward of a SMALL tensor (a few MB, C >= 32) without any grid-wide synchronisation: the reduction of
// BatchNorm is per channel, so the work is partitioned BY CHANNEL over thread-block clusters -- cluster = one vector of 8
// channels, its S CTAs (x 256 threads x E rows per thread) cover all rows of those channels.  The three sums never leave
// the cluster: warp shuffles -> shared memory -> every CTA pushes its 24 partial sums into the slot [its rank] of every
// peer through distributed shared memory -> ONE hardware cluster barrier -> every CTA adds the S slots in rank order
// (fixed order: bit-reproducible).  The rows stay in registers across the barrier (the cluster's CTAs are co-scheduled by
// the hardware: no residency assumption, no spinning), dY (dY2) is written from them: every byte is touched once.
// Replaces, for these layers, bn_bwd_sweep_kernel's fp64 global atomics + __threadfence + atomic grid barrier + second pass
// (12-15 us on the backward chain per BatchNorm layer, profiles/r02_bn_cluster.md).  A thread reads 16 B of a C * 2 B row:
// half-used 32-byte sectors for the loads, but these tensors are bound by latency, not by L2 bandwidth.
__device__ __forceinline__ uint32_t bnc_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t bnc_cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void bnc_st_peer_f32(const float* local, uint32_t rank, float v) {
  uint32_t la = (uint32_t)__cvta_generic_to_shared(local), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bnc_unpack8(const uint4& u, float (&v)[8]) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t bnc_pack2(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ uint4 bnc_pack8(const float (&v)[8]) {
  return make_uint4(bnc_pack2(v[0], v[1]), bnc_pack2(v[2], v[3]), bnc_pack2(v[4], v[5]), bnc_pack2(v[6], v[7]));
}

// BatchNorm-backward REDUCTION alone for a large bf16 tensor whose dY is formed by its consumer (the fused stem backward:
// BnBwdArgs::no_apply): raw moments sum g, sum g*y in fp32 per thread (centred per CTA: S1 = rstd * (sum g*y - mean * S0)),
// two rows per trip with all eight loads issued first, the forward-written operands pulled into L2 before the dependency
// wait, fixed-order shared-memory reduction, fp64 atomics across CTAs, and the CTA that finishes last writes the
// coefficients (scale, c1, c2) and d gamma / d beta.  One branch only (no y2).
template <bool kTwo>
__global__ void __launch_bounds__(256, kTwo ? 3 : 4) bn_bwd_reduce_bf16_kernel(const BnBwdArgs a) {
  extern __shared__ float red[];                       // [2 or 3 sums][8 warps][C]
  __shared__ int is_last;
  const int CV = a.C / 8, RPI = 256 / CV;
  const int tid = threadIdx.x;
  const int cv = tid % CV, rsub = tid / CV;
  const uint4* dA = reinterpret_cast<const uint4*>(a.dA);
  const uint4* dA2 = reinterpret_cast<const uint4*>(a.dA2);
  const uint4* am = reinterpret_cast<const uint4*>(a.a);
  const uint4* y1 = reinterpret_cast<const uint4*>(a.y);
  const uint4* y2 = reinterpret_cast<const uint4*>(a.y2);
  const long long rstride = (long long)gridDim.x * RPI;
  const long long r0 = (long long)blockIdx.x * RPI + rsub;
  for (long long r = r0; r < a.rows; r += rstride) {
    const size_t off = size_t(r) * CV + cv;
    if (am) asm volatile("prefetch.global.L2 [%0];" ::"l"(am + off));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(y1 + off));
    if (kTwo) asm volatile("prefetch.global.L2 [%0];" ::"l"(y2 + off));
  }
  pdl_wait();
  pdl_trigger();
  float s0[8], s1[8], s2[kTwo ? 8 : 1];
#pragma unroll
  for (int k = 0; k < 8; ++k) { s0[k] = 0.f; s1[k] = 0.f; if (kTwo) s2[k] = 0.f; }
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
  for (long long r = r0; r < a.rows; r += 2 * rstride) {
    uint4 qg[2], qg2[2], qa[2], qy[2], qz[kTwo ? 2 : 1];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long ru = r + u * rstride;
      const bool ok = ru < a.rows;
      const size_t off = size_t(ok ? ru : 0) * CV + cv;
      qg[u] = ok ? dA[off] : z4;
      qg2[u] = (dA2 && ok) ? dA2[off] : z4;
      qa[u] = (am && ok) ? am[off] : z4;
      qy[u] = ok ? y1[off] : z4;
      if (kTwo) qz[u] = ok ? y2[off] : z4;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float g[8], v[8];
      bnc_unpack8(qg[u], g);
      if (dA2) {
        bnc_unpack8(qg2[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = __bfloat162float(__float2bfloat16_rn(g[k] + v[k]));
      }
      if (am) {
        bnc_unpack8(qa[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
      }
      bnc_unpack8(qy[u], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0[k] += g[k]; s1[k] = fmaf(g[k], v[k], s1[k]); }
      if (kTwo) {
        bnc_unpack8(qz[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) s2[k] = fmaf(g[k], v[k], s2[k]);
      }
    }
  }
  for (int d = CV; d < 32; d <<= 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s0[k] += __shfl_xor_sync(0xffffffffu, s0[k], d);
      s1[k] += __shfl_xor_sync(0xffffffffu, s1[k], d);
      if (kTwo) s2[k] += __shfl_xor_sync(0xffffffffu, s2[k], d);
    }
  }
  const int wrp = tid >> 5, c0 = cv * 8;
  if ((tid & 31) < CV) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[(0 * 8 + wrp) * a.C + c0 + k] = s0[k]; red[(1 * 8 + wrp) * a.C + c0 + k] = s1[k];
      if (kTwo) red[(2 * 8 + wrp) * a.C + c0 + k] = s2[k];
    }
  }
  __syncthreads();
  for (int e = tid; e < (kTwo ? 3 : 2) * a.C; e += 256) {
    const int which = e / a.C, c = e - which * a.C;
    float sum = 0.f, t0 = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) { sum += red[(which * 8 + q) * a.C + c]; t0 += red[q * a.C + c]; }
    if (which == 1) sum = a.stat[a.C + c] * (sum - a.stat[c] * t0);
    if (which == 2) sum = a.stat2[a.C + c] * (sum - a.stat2[c] * t0);
    atomicAdd(a.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 * a.C + which * a.C + c, (double)sum);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) is_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const double im = 1.0 / (double)a.rows;
  for (int c = tid; c < a.C; c += 256) {
    double t0 = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) {
      const double* ak = a.acc + (size_t)k * 3 * a.C;
      t0 += ld_cg_f64(ak + c); t1 += ld_cg_f64(ak + a.C + c);
      if (kTwo) t2 += ld_cg_f64(ak + 2 * a.C + c);
    }
    a.g_beta[c] = (float)t0; a.g_gamma[c] = (float)t1;
    a.bcoef[c] = a.gamma[c] * a.stat[a.C + c];
    a.bcoef[a.C + c] = (float)(t0 * im);
    a.bcoef[2 * a.C + c] = (float)(t1 * im);
    if (kTwo) {
      a.g_beta2[c] = (float)t0; a.g_gamma2[c] = (float)t2;
      a.bcoef2[c] = a.gamma2[c] * a.stat2[a.C + c];
      a.bcoef2[a.C + c] = (float)(t0 * im);
      a.bcoef2[2 * a.C + c] = (float)(t2 * im);
    }
  }
}

// bn_bwd_apply for bf16 storage with the per-channel algebra hoisted out of the streaming loop: the grid stride is a
// multiple of C, so a thread sees ONE vector of 8 channels for its whole life and dY = A * g + B * y + D with three
// coefficients per channel and branch kept in registers (the generic kernel re-loads ten coefficient vectors per
// iteration: more LSU instructions than data loads).  Two vectors per trip, every load issued before the first use.
// `reverse`: walk the tensor from its END: the producer of dA (and of the forward tensors it re-read) walked it from the
// start, so the end is what is still in L2 (126 MB against 134-168 MB touched by the uplayer5 / tail kernels).
__global__ void __launch_bounds__(256) bn_bwd_apply_bf16_kernel(const BnBwdArgs a, int reverse) {
  const long long nvec = a.rows * a.C / 8;
  const long long stride = (long long)gridDim.x * 256;
  const long long i0 = blockIdx.x * 256LL + threadIdx.x;
  const int c0 = (int)((i0 * 8) % a.C);
  const bool two = a.y2 != nullptr;
  float mean[8], rstd[8], mean2[8], rstd2[8];
  load_coef<8>(a.stat + c0, mean); load_coef<8>(a.stat + a.C + c0, rstd);      // written by the forward
  if (two) { load_coef<8>(a.stat2 + c0, mean2); load_coef<8>(a.stat2 + a.C + c0, rstd2); }
  pdl_wait();
  pdl_trigger();
  if (a.reduced && blockIdx.x == 0) {
    for (int c = threadIdx.x; c < a.C; c += 256) {
      a.g_beta[c] = a.bcoef[3 * a.C + c]; a.g_gamma[c] = a.bcoef[4 * a.C + c];
      if (two) { a.g_beta2[c] = a.bcoef2[3 * a.C + c]; a.g_gamma2[c] = a.bcoef2[4 * a.C + c]; }
    }
  }
  float A[8], B[8], D[8], A2[8], B2[8], D2[8];
  {
    float b0[8], b1[8], b2[8];
    load_coef<8>(a.bcoef + c0, b0); load_coef<8>(a.bcoef + a.C + c0, b1); load_coef<8>(a.bcoef + 2 * a.C + c0, b2);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t = b0[k] * b2[k] * rstd[k];
      A[k] = b0[k]; B[k] = -t; D[k] = fmaf(t, mean[k], -b0[k] * b1[k]);
    }
    if (two) {
      load_coef<8>(a.bcoef2 + c0, b0); load_coef<8>(a.bcoef2 + a.C + c0, b1); load_coef<8>(a.bcoef2 + 2 * a.C + c0, b2);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float t = b0[k] * b2[k] * rstd2[k];
        A2[k] = b0[k]; B2[k] = -t; D2[k] = fmaf(t, mean2[k], -b0[k] * b1[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) { A2[k] = 0.f; B2[k] = 0.f; D2[k] = 0.f; }
    }
  }
  const uint4* dA = reinterpret_cast<const uint4*>(a.dA);
  const uint4* dA2 = reinterpret_cast<const uint4*>(a.dA2);
  const uint4* am = reinterpret_cast<const uint4*>(a.a);
  const uint4* y1 = reinterpret_cast<const uint4*>(a.y);
  const uint4* y2 = reinterpret_cast<const uint4*>(a.y2);
  uint4* dY = reinterpret_cast<uint4*>(a.dY);
  uint4* dY2 = reinterpret_cast<uint4*>(a.dY2);
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  const long long rbase = nvec - (a.C >> 3) + 2 * (c0 >> 3);     // mirrored row, same vector within the row
#pragma unroll 1
  for (long long i = i0; i < nvec; i += 2 * stride) {
    uint4 qg[2], qg2[2], qa[2], qy[2], qz[2];
    long long j[2];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long iu = i + u * stride;
      ok[u] = iu < nvec;
      // reverse: the same thread <-> channel-vector assignment, rows from the end (nvec - 1 - i is congruent to a different
      // channel vector, so the vector INDEX within the row is kept and only the row order is mirrored)
      j[u] = ok[u] ? (reverse ? rbase - iu : iu) : 0;
      qg[u] = ok[u] ? dA[j[u]] : z4;
      qg2[u] = (dA2 && ok[u]) ? dA2[j[u]] : z4;
      qa[u] = (am && ok[u]) ? am[j[u]] : z4;
      qy[u] = ok[u] ? y1[j[u]] : z4;
      qz[u] = (two && ok[u]) ? y2[j[u]] : z4;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (!ok[u]) continue;
      float g[8], v[8], o[8];
      bnc_unpack8(qg[u], g);
      if (dA2) {
        bnc_unpack8(qg2[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = __bfloat162float(__float2bfloat16_rn(g[k] + v[k]));
      }
      if (am) {
        bnc_unpack8(qa[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
      }
      bnc_unpack8(qy[u], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(A[k], g[k], fmaf(B[k], v[k], D[k]));
      dY[j[u]] = bnc_pack8(o);
      if (two) {
        bnc_unpack8(qz[u], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = fmaf(A2[k], g[k], fmaf(B2[k], v[k], D2[k]));
        dY2[j[u]] = bnc_pack8(o);
      }
    }
  }
}

constexpr int kBncMaxCluster = 16;

template <int E>
__global__ void __launch_bounds__(256, E <= 2 ? 3 : 2) bn_bwd_cluster_kernel(const BnBwdArgs a) {
  __shared__ float wred[8][24];
  __shared__ float slots[kBncMaxCluster][24];          // [peer rank][3 sums][8 channels]
  __shared__ float tot[24];
  __shared__ __align__(16) float cst[2][3][8];         // [branch][rstd | -mean * rstd | gamma * rstd][channel]
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const uint32_t S = bnc_cluster_nctarank(), rank = bnc_cluster_ctarank();
  const int c0 = (int)(blockIdx.x / S) * 8;
  const bool two = a.y2 != nullptr;
  // xhat = y * rstd + nmr; statistics and affine weights were written long before the predecessor: no need to wait for it
  if (tid < 16) {
    const int br = tid >> 3, k = tid & 7;
    if (br == 0 || two) {
      const float* st = br ? a.stat2 : a.stat;
      const float m = st[c0 + k], r = st[a.C + c0 + k];
      cst[br][0][k] = r; cst[br][1][k] = -m * r; cst[br][2][k] = (br ? a.gamma2 : a.gamma)[c0 + k] * r;
    } else {
      cst[1][0][k] = 0.f; cst[1][1][k] = 0.f; cst[1][2][k] = 0.f;
    }
  }
  __syncthreads();
  const uint4* dA = reinterpret_cast<const uint4*>(a.dA);
  const uint4* dA2 = reinterpret_cast<const uint4*>(a.dA2);
  const uint4* am = reinterpret_cast<const uint4*>(a.a);
  const uint4* y1 = reinterpret_cast<const uint4*>(a.y);
  const uint4* y2 = reinterpret_cast<const uint4*>(a.y2);
  const int CV = a.C >> 3, cv = c0 >> 3;
  // ---- every load of this thread in flight at once.  The ReLU output and the raw conv outputs were written by the
  // FORWARD (cold in L2 by now, ~1 us from HBM): they are fetched BEFORE the dependency wait, under the predecessor's tail;
  // only the incoming gradient is the predecessor's ----
  uint4 rg[E], rg2[E], ra[E], ry[E], rz[E];
  size_t off[E];
  bool ok[E];
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const long long r = ((long long)e * S + rank) * 256 + tid;
    ok[e] = r < a.rows;
    off[e] = (size_t)(ok[e] ? r : 0) * CV + cv;
    const bool early = ok[e] && !a.late_loads;
    ra[e] = (am && early) ? am[off[e]] : z4;
    ry[e] = early ? y1[off[e]] : z4;
    rz[e] = (two && early) ? y2[off[e]] : z4;
  }
  pdl_wait();
  pdl_trigger();
  if (a.late_loads) {
#pragma unroll
    for (int e = 0; e < E; ++e) {
      ra[e] = (am && ok[e]) ? ld_cg_u4(am + off[e]) : z4;
      ry[e] = ok[e] ? ld_cg_u4(y1 + off[e]) : z4;
      rz[e] = (two && ok[e]) ? ld_cg_u4(y2 + off[e]) : z4;
    }
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    rg[e] = ok[e] ? dA[off[e]] : z4;
    rg2[e] = (dA2 && ok[e]) ? dA2[off[e]] : z4;
  }
  // ---- g = (dA [+ dA2]) * [a > 0] rounded to bf16 once (what the accumulate-in-place path stored); the three sums ----
  float s[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) s[k] = 0.f;
#pragma unroll
  for (int e = 0; e < E; ++e) {
    float g[8], v[8];
    bnc_unpack8(rg[e], g);
    if (dA2) {
      bnc_unpack8(rg2[e], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = __bfloat162float(__float2bfloat16_rn(g[k] + v[k]));
    }
    if (am) {
      bnc_unpack8(ra[e], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
    }
    rg[e] = bnc_pack8(g);                              // exact: g is a bf16 value
    bnc_unpack8(ry[e], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += g[k];
      s[8 + k] = fmaf(g[k], fmaf(v[k], cst[0][0][k], cst[0][1][k]), s[8 + k]);
    }
    if (two) {
      bnc_unpack8(rz[e], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[16 + k] = fmaf(g[k], fmaf(v[k], cst[1][0][k], cst[1][1][k]), s[16 + k]);
    }
  }
  const int nsum = two ? 24 : 16;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
    for (int k = 0; k < 24; ++k)
      if (k < nsum) s[k] += __shfl_xor_sync(0xffffffffu, s[k], d);
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 24; ++k) wred[wrp][k] = s[k];
  }
  __syncthreads();
  if (tid < 24) {
    float p = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) p += wred[w][tid];
    for (uint32_t r = 0; r < S; ++r) bnc_st_peer_f32(&slots[rank][tid], r, p);
  }
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (tid < 24) {
    float t = 0.f;
    for (uint32_t r = 0; r < S; ++r) t += slots[r][tid];
    tot[tid] = t;
  }
  __syncthreads();
  // ---- dY = gamma * rstd * (g - mean(g) - xhat * mean(g * xhat)) from the registers ----
  const float im = (float)(1.0 / (double)a.rows);
  if (rank == 0 && tid < 8) {
    a.g_beta[c0 + tid] = tot[tid]; a.g_gamma[c0 + tid] = tot[8 + tid];
    if (two) { a.g_beta2[c0 + tid] = tot[tid]; a.g_gamma2[c0 + tid] = tot[16 + tid]; }
  }
  uint4* dY = reinterpret_cast<uint4*>(a.dY);
  uint4* dY2 = reinterpret_cast<uint4*>(a.dY2);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (!ok[e]) continue;
    float g[8], v[8], o[8];
    bnc_unpack8(rg[e], g);
    bnc_unpack8(ry[e], v);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = cst[0][2][k] * (g[k] - tot[k] * im - fmaf(v[k], cst[0][0][k], cst[0][1][k]) * (tot[8 + k] * im));
    dY[off[e]] = bnc_pack8(o);
    if (two) {
      bnc_unpack8(rz[e], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = cst[1][2][k] * (g[k] - tot[k] * im - fmaf(v[k], cst[1][0][k], cst[1][1][k]) * (tot[16 + k] * im));
      dY2[off[e]] = bnc_pack8(o);
    }
  }
}

// BatchNorm backward of a ONE-channel tensor (the decoder's output BatchNorm, model.py:173,193) whose incoming gradient is
// fp32 (d recon from the loss): same scheme as bn_bwd_coop_kernel -- every thread keeps its (at most kCoopE) groups of 4
// gradients and pre-activations in registers across the grid barrier -- for rows up to grid * 256 * 4 * kCoopE.
__global__ void __launch_bounds__(256, 2) bn_bwd_c1_coop_kernel(const BnBwdArgs a, int E) {
  pdl_wait();
  pdl_trigger();
  typedef __nv_bfloat16 T;
  __shared__ float red[2][8];
  __shared__ float coef[3];
  const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
  const float4* dA = reinterpret_cast<const float4*>(a.dA);
  const uint2* y = reinterpret_cast<const uint2*>(a.y);
  const long long nq = a.rows >> 2;                   // groups of 4 (rows % 4 == 0 guaranteed by the launcher)
  const float mean = a.stat[0], rstd = a.stat[1];
  float4 g[kCoopE];
  float xh[kCoopE][4];
  float s0 = 0.f, s1 = 0.f;
  const long long q00 = (long long)blockIdx.x * 256 + tid, qstride = (long long)gridDim.x * 256;
#pragma unroll
  for (int e = 0; e < kCoopE; ++e) {
    const long long q = q00 + e * qstride;
    if (e < E && q < nq) {
      g[e] = __ldg(dA + q);
      const uint2 yr = __ldg(y + q);
      xh[e][0] = (__uint_as_float(yr.x << 16) - mean) * rstd; xh[e][1] = (__uint_as_float(yr.x & 0xffff0000u) - mean) * rstd;
      xh[e][2] = (__uint_as_float(yr.y << 16) - mean) * rstd; xh[e][3] = (__uint_as_float(yr.y & 0xffff0000u) - mean) * rstd;
      s0 += (g[e].x + g[e].y) + (g[e].z + g[e].w);
      s1 = fmaf(g[e].x, xh[e][0], fmaf(g[e].y, xh[e][1], fmaf(g[e].z, xh[e][2], fmaf(g[e].w, xh[e][3], s1))));
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, d); s1 += __shfl_xor_sync(0xffffffffu, s1, d); }
  if (lane == 0) { red[0][wrp] = s0; red[1][wrp] = s1; }
  __syncthreads();
  if (tid < 2) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[tid][w];
    atomicAdd(a.acc + (size_t)(blockIdx.x % kBnAccCopies) * 3 + tid, (double)t);
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    atomicAdd(a.counter, 1u);
    while (ld_acquire_u32(a.counter) < gridDim.x) { __nanosleep(32); }
  }
  __syncthreads();
  if (tid == 0) {
    double t0 = 0.0, t1 = 0.0;
#pragma unroll
    for (int k = 0; k < kBnAccCopies; ++k) { t0 += ld_cg_f64(a.acc + (size_t)k * 3); t1 += ld_cg_f64(a.acc + (size_t)k * 3 + 1); }
    const double im = 1.0 / (double)a.rows;
    coef[0] = a.gamma[0] * rstd; coef[1] = (float)(t0 * im); coef[2] = (float)(t1 * im);
    if (blockIdx.x == 0) { a.g_beta[0] = (float)t0; a.g_gamma[0] = (float)t1; }
  }
  __syncthreads();
  const float sc = coef[0], c1 = coef[1], c2 = coef[2];
  uint2* dY = reinterpret_cast<uint2*>(a.dY);
#pragma unroll
  for (int e = 0; e < kCoopE; ++e) {
    const long long q = q00 + e * qstride;
    if (e < E && q < nq) {
      const float o0 = sc * (g[e].x - c1 - xh[e][0] * c2), o1 = sc * (g[e].y - c1 - xh[e][1] * c2);
      const float o2 = sc * (g[e].z - c1 - xh[e][2] * c2), o3 = sc * (g[e].w - c1 - xh[e][3] * c2);
      __nv_bfloat162 lo = __floats2bfloat162_rn(o0, o1), hi = __floats2bfloat162_rn(o2, o3);
      dY[q] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                           long long total, int C, int HW) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
    int c = (int)(i % C); long long t = i / C; int hw = (int)(t % HW); long long n = t / HW;
    out[i] = in[(n * C + c) * HW + hw];
  }
}

inline int grid_for(long long work_items, int per_block = 256, int cap = 148 * 8) {
  long long b = (work_items + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

template <typename T> constexpr int vec_of() { return 16 / (int)sizeof(T); }

}  // namespace

void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st) {
  count_launch();
  bn_finalize_kernel<<<a.C, 128, 0, st>>>(a);
}

template <typename T>
void launch_bn_apply(const T* y, const float* coef, const T* y2, const float* coef2, T* out,
                     long long rows, int C, int relu, cudaStream_t st) {
  constexpr int V = vec_of<T>();
  if (C % V == 0) {
    long long nvec = rows * C / V;
    count_launch();
    launch_pdl(bn_apply_kernel<T, V>, grid_for(nvec), 256, 0, st, y, coef, y2, coef2, out, nvec, C, relu);
  } else {
    long long nvec = rows * C;
    count_launch();
    launch_pdl(bn_apply_kernel<T, 1>, grid_for(nvec), 256, 0, st, y, coef, y2, coef2, out, nvec, C, relu);
  }
}

template <typename T>
void launch_bn_apply_out(const T* y, const float* coef, float* out_nchw, int N, int HW, int C, cudaStream_t st) {
  long long total = (long long)N * HW * C;
  count_launch();
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (C == 1 && (total & 7) == 0 && (reinterpret_cast<uintptr_t>(out_nchw) & 15) == 0) {
      launch_pdl(bn_apply_out_c1_kernel, grid_for(total / 8), 256, 0, st, y, coef, out_nchw, total / 8);
      return;
    }
  }
  launch_pdl(bn_apply_out_kernel<T>, grid_for(total), 256, 0, st, y, coef, out_nchw, total, HW, C);
}

template <typename T>
void launch_heads_fwd(const HeadsArgs& a, cudaStream_t st) {
  size_t smem = sizeof(float) * (size_t(a.C) + 2 * size_t((a.z + kHeadsSplit - 1) / kHeadsSplit));
  count_launch();
  launch_pdl(heads_fwd_kernel<T>, dim3(a.N, kHeadsSplit), 128, smem, st, a);
}

template <typename T>
void launch_cast_latent(const float* enc, T* z_act, long long n, cudaStream_t st) {
  count_launch();
  cast_latent_kernel<T><<<grid_for(n), 256, 0, st>>>(enc, z_act, n);
}

template <typename T>
void launch_heads_bwd(const HeadsBwdArgs& a, cudaStream_t st) {
  size_t smem = sizeof(float) * 2 * size_t(a.z);
  count_launch();
  launch_pdl(heads_bwd_kernel<T>, dim3(a.N, std::max(1, std::min(4, a.C / 128))), 128, smem, st, a);
  // the weight gradients of the two heads are a separate launch (launch_heads_wgrad): the caller decides the stream
}

// co-resident CTAs of the grid-barrier BatchNorm-backward kernels on the current device (cached per device; the smaller of
// the two kernels' occupancies)
static int coop_max_ctas() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int v = cached[dev].load();
  if (v > 0) return v;
  int per_sm = 0, per_sm1 = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, bn_bwd_c1_coop_kernel, 256, 0) != cudaSuccess) { cudaGetLastError(); return 0; }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_coop_kernel, 256, 0) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  v = std::min(std::min(per_sm, per_sm1), 2) * sms;
  static const int cap = [] { const char* e = getenv("MMVAE_COOP_CTAS"); return e ? atoi(e) : 0; }();   // A/B runs
  if (cap > 0) v = std::min(v, cap);
  cached[dev].store(v);
  return v;
}

// co-resident CTAs of bn_bwd_sweep_kernel (two per SM at most; cached per device)
static int sweep_max_ctas() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int v = cached[dev].load();
  if (v > 0) return v;
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_sweep_kernel, 256, sizeof(float) * 3 * 8 * 256) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  static const int per = [] { const char* e = getenv("MMVAE_BN_SWEEP_PER_SM"); return e ? atoi(e) : 2; }();
  v = std::min(per_sm, per) * sms;
  cached[dev].store(v);
  return v;
}

// largest cluster bn_bwd_cluster_kernel may be launched with on this device: 16 (non-portable, opted in) when the device can
// hold at least one such cluster, else 8 (portable); cached per device
static int bnc_max_cluster() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  int v = cached[dev].load();
  if (v > 0) return v;
  static const int cap = [] { const char* e = getenv("MMVAE_BN_CLUSTER_MAX"); return e ? atoi(e) : 8; }();   // A/B runs
  v = 8;
  if (cap >= 16) {
    bool ok16 = true;
    auto opt_in = [&](auto kern) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); ok16 = false; return; }
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(16); cfg.blockDim = dim3(256);
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 16; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 1) { cudaGetLastError(); ok16 = false; }
    };
    opt_in(bn_bwd_cluster_kernel<1>); opt_in(bn_bwd_cluster_kernel<2>); opt_in(bn_bwd_cluster_kernel<4>);
    if (ok16) v = 16;
  } else if (cap >= 1) {
    v = std::min(cap, 8);
  }
  cached[dev].store(v);
  return v;
}

// cluster shape of bn_bwd_cluster_kernel for `rows` rows: the fewest rows per thread (E in 1, 2, 4) whose cluster (a power of
// two of CTAs x 256 threads x E rows) stays portable (<= 8 CTAs), else E = 4 with up to `smax` CTAs; false: too many rows
static bool bnc_shape(long long rows, int smax, int& E, int& S) {
  for (int pass = 0; pass < 2; ++pass) {
    const int lim = pass == 0 ? std::min(smax, 8) : smax;
    for (int e = 1; e <= 4; e <<= 1) {
      const long long need = (rows + 256LL * e - 1) / (256LL * e);
      int s = 1;
      while (s < need) s <<= 1;
      if (s <= lim) { E = e; S = s; return true; }
    }
  }
  return false;
}

template <typename T>
void launch_bn_bwd(const BnBwdArgs& a_in, cudaStream_t st) {
  constexpr int V = vec_of<T>();
  BnBwdArgs a = a_in;
  static const bool late = getenv("MMVAE_BN_LATE_LOADS") != nullptr;
  a.late_loads = late ? 1 : 0;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // channel-partitioned clusters (no grid-wide synchronisation) for the small many-channel tensors
    static const bool cluster_off = getenv("MMVAE_NO_BN_CLUSTER") != nullptr;
    if (!cluster_off && !a.no_apply && !a.reduced && !a.dA_f32 && a.C % 8 == 0 && a.C >= 32) {
      int E = 0, S = 0;
      const int smax = bnc_max_cluster();
      if (smax > 0 && bnc_shape(a.rows, smax, E, S)) {
        const int grid = (a.C / 8) * S;
        cudaError_t err;
        if (E == 1) err = launch_pdl_cluster(bn_bwd_cluster_kernel<1>, grid, 256, 0, st, S, a);
        else if (E == 2) err = launch_pdl_cluster(bn_bwd_cluster_kernel<2>, grid, 256, 0, st, S, a);
        else err = launch_pdl_cluster(bn_bwd_cluster_kernel<4>, grid, 256, 0, st, S, a);
        if (err == cudaSuccess) { count_launch(); return; }
        cudaGetLastError();                            // could not be launched in this shape: the grid-barrier kernel below
      }
    }
    // one cooperative launch when every thread's share fits its registers: half a register file per SM for two CTAs,
    // so it stays co-resident with the weight-gradient kernels of the auxiliary stream
    static const bool coop_off = getenv("MMVAE_NO_COOP_BN") != nullptr;
    static const bool sweep_off = getenv("MMVAE_NO_BN_SWEEP") != nullptr;       // A/B: the register-resident kernel below
    static const long long sweep_max = [] { const char* e = getenv("MMVAE_BN_SWEEP_MAX_MB"); return (long long)(e ? atoi(e) : 1 << 20) << 20; }();
    const int CV = a.C / 8;
    if (!coop_off && !sweep_off && !a.no_apply && a.acc && !a.reduced && !a.dA_f32 && a.C % 8 == 0 && CV >= 1 && CV <= 32 && 256 % CV == 0 &&
        a.rows * a.C * 2 <= sweep_max) {
      const int RPI = 256 / CV;
      const long long row_groups = (a.rows + RPI - 1) / RPI;
      const int grid = (int)std::min<long long>(sweep_max_ctas(), row_groups);
      if (grid > 0) {
        count_launch();
        launch_pdl(bn_bwd_sweep_kernel, grid, 256, sizeof(float) * 3 * 8 * a.C, st, a);
        return;
      }
    }
    if (!coop_off && !a.no_apply && a.acc && !a.reduced && !a.dA_f32 && a.C % 8 == 0 && CV >= 1 && CV <= 32 && 256 % CV == 0) {
      const int RPI = 256 / CV;
      const long long row_groups = (a.rows + RPI - 1) / RPI;
      // up to two CTAs per SM (__launch_bounds__(256, 2), 34 KB of shared memory each): tensors up to 4.8 MB.  Whatever
      // else occupies the SMs finishes without waiting for this stream, and a programmatic dependent of this launch is
      // not scheduled before every CTA here has started, so all CTAs reach the grid barrier.  MMVAE_COOP_CTAS: A/B.
      // The grid never exceeds what the device can hold at once (occupancy x SM count, queried per device): on a part with
      // fewer SMs, or where only one CTA fits, a larger grid would leave CTAs unscheduled and the resident ones spinning.
      const int grid = (int)std::min<long long>(coop_max_ctas(), row_groups);
      const int E = grid > 0 ? (int)((row_groups + grid - 1) / grid) : kCoopE + 1;
      if (E <= kCoopE) {
        count_launch();
        launch_pdl(bn_bwd_coop_kernel, grid, 256, 0, st, a, E);
        return;
      }
    }
  }
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // the one-channel output BatchNorm with an fp32 incoming gradient: its own grid-barrier kernel
    static const bool coop_off = getenv("MMVAE_NO_COOP_BN") != nullptr;
    if (!coop_off && a.acc && !a.reduced && a.dA_f32 && a.C == 1 && !a.a && !a.y2 && (a.rows & 3) == 0) {
      const long long groups = ((a.rows >> 2) + 255) / 256;
      const int grid = (int)std::min<long long>(coop_max_ctas(), groups);
      const int E = grid > 0 ? (int)((groups + grid - 1) / grid) : kCoopE + 1;
      if (E <= kCoopE) {
        count_launch();
        launch_pdl(bn_bwd_c1_coop_kernel, grid, 256, 0, st, a, E);
        return;
      }
    }
  }
  bool lean_done = false;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    const int CV = a.C / 8;
    // lean reduction (+ the hoisted-coefficient apply below unless the consumer forms dY itself)
    static const bool lean_off = getenv("MMVAE_NO_BN_LEAN_REDUCE") != nullptr;
    if ((a.no_apply || !lean_off) && a.acc && !a.reduced && !a.dA_f32 && a.C % 8 == 0 && CV >= 1 && CV <= 32 && 256 % CV == 0) {
      const int RPI = 256 / CV;
      const long long row_groups = (a.rows + RPI - 1) / RPI;
      count_launch();
      if (a.y2) launch_pdl(bn_bwd_reduce_bf16_kernel<true>, (int)std::min<long long>(148 * 3, row_groups), 256, sizeof(float) * 3 * 8 * a.C, st, a);
      else launch_pdl(bn_bwd_reduce_bf16_kernel<false>, (int)std::min<long long>(148 * 4, row_groups), 256, sizeof(float) * 2 * 8 * a.C, st, a);
      if (a.no_apply) return;
      lean_done = true;
    }
  }
  const bool vec_ok = (a.C % V == 0) && (a.C / V <= 256);
  int nblocks;
  if (vec_ok) {
    int rpi = 256 / (a.C / V);
    nblocks = (int)((a.rows + rpi - 1) / rpi);
  } else {
    int rpi = a.C <= 256 ? 256 / a.C : 1;
    nblocks = (int)((a.rows + rpi - 1) / rpi);
  }
  if (nblocks > 592) nblocks = 592;
  if (nblocks < 1) nblocks = 1;
  if (a.reduced || lean_done) { /* masked and reduced by the producer of dA / by the lean kernel above */ }
  else if (vec_ok) { count_launch(); launch_pdl(bn_bwd_reduce_kernel<T, V>, nblocks, 256, 0, st, a); }
  else { count_launch(); launch_pdl(bn_bwd_reduce_kernel<T, 1>, nblocks, 256, 0, st, a); }
  if (!a.acc && !a.reduced) {
    count_launch();
    bn_bwd_finalize_kernel<<<a.C, 128, 0, st>>>(a, nblocks, a.rows);
  }
  if (a.no_apply) return;                            // the consumer of dY forms it from bcoef itself (fused stem backward)
  long long total = a.rows * a.C;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    // hoisted-coefficient kernel: needs a thread's channel vector to be loop-invariant (grid stride a multiple of C / 8)
    static const int apply_mode = [] { const char* e = getenv("MMVAE_BN_APPLY"); return e ? atoi(e) : 3; }();   // A/B: bit 0 fast kernel, bit 1 reverse
    const int grid = grid_for(total / V);
    if ((apply_mode & 1) && !a.dA_f32 && a.C % 8 == 0 && ((long long)grid * 256) % (a.C / 8) == 0) {
      count_launch();
      launch_pdl(bn_bwd_apply_bf16_kernel, grid, 256, 0, st, a, (apply_mode >> 1) & 1);
      return;
    }
  }
  if (a.C % V == 0) { count_launch(); launch_pdl(bn_bwd_apply_kernel<T, V>, grid_for(total / V), 256, 0, st, a); }
  else { count_launch(); launch_pdl(bn_bwd_apply_kernel<T, 1>, grid_for(total), 256, 0, st, a); }
}

void launch_nchw_to_nhwc(const float* in, float* out, int N, int C, int HW, cudaStream_t st) {
  long long total = (long long)N * C * HW;
  count_launch();
  nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, st>>>(in, out, total, C, HW);
}

#define INST(T)                                                                                              \
  template void launch_bn_apply<T>(const T*, const float*, const T*, const float*, T*, long long, int, int, cudaStream_t); \
  template void launch_bn_apply_out<T>(const T*, const float*, float*, int, int, int, cudaStream_t);         \
  template void launch_heads_fwd<T>(const HeadsArgs&, cudaStream_t);                                         \
  template void launch_cast_latent<T>(const float*, T*, long long, cudaStream_t);                            \
  template void launch_heads_bwd<T>(const HeadsBwdArgs&, cudaStream_t);                                      \
  template void launch_bn_bwd<T>(const BnBwdArgs&, cudaStream_t);
INST(float)
INST(__nv_bfloat16)
#undef INST

}  // namespace mmvae
