// gconv_simt.cu -- fp32-accumulate SIMT gather-convolution and weight-gradient kernels.
//
// These are the arithmetic of the fp32 validation mode (MMVAE_PREC_FP32), and the bring-up /
// remainder path of the bf16 mode for layer shapes the tcgen05 kernels do not cover (stem conv
// with Ci = in_channels, tail conv with Co = out_channels).  Every Conv2d / ConvTranspose2d of
// the reference (model.py:12-20, 60-65, 94, 159-161, 172, 198-201) and every dgrad is expressed
// as a GConvParams instance by api.cu.
#include "kernels.cuh"

namespace mmvae {

namespace {

constexpr int kThreads = 256;
constexpr int BK = 16;

template <typename T>
__device__ __forceinline__ float load_in(const GConvParams& p, int n, int iy, int ix, int ci) {
  if (p.in_nchw_f32)
    return reinterpret_cast<const float*>(p.in)[((size_t(n) * p.Ci + ci) * p.Hi + iy) * p.Wi + ix];
  return to_f(reinterpret_cast<const T*>(p.in)[((size_t(n) * p.Hi + iy) * p.Wi + ix) * p.Ci + ci]);
}

// BN_ in {16, 32, 64}; the CTA tile is BM x BN_ with BM*BN_ = 4096 and a 4x4 micro-tile per thread.
template <typename T, int BN_>
__global__ void __launch_bounds__(kThreads) gconv_simt_kernel(const __grid_constant__ GConvParams p) {
  constexpr int BM = 4096 / BN_;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN_ + 4];
  __shared__ float red[2][BM / 4][BN_];
  __shared__ int row_n[BM];
  __shared__ short row_i[BM], row_j[BM];

  const GVar& v = p.var[blockIdx.z];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN_;
  for (int r = tid; r < BM; r += kThreads) {
    int m = m0 + r;
    if (m < p.M) {
      int j = m % p.Wg; int t = m / p.Wg; int i = t % p.Hg; int n = t / p.Hg;
      row_n[r] = n; row_i[r] = (short)i; row_j[r] = (short)j;
    } else {
      row_n[r] = -1; row_i[r] = 0; row_j[r] = 0;
    }
  }
  __syncthreads();

  const int K = v.ntaps * p.Ci;
  const int tx = tid % (BN_ / 4), ty = tid / (BN_ / 4);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const bool aligned = (p.Ci % BK) == 0;
  for (int k0 = 0; k0 < K; k0 += BK) {
    int t_c = 0, ci_c = 0;
    if (aligned) { t_c = k0 / p.Ci; ci_c = k0 - t_c * p.Ci; }
    // A tile: BM x BK gathered input values
    for (int e = tid; e < BM * BK; e += kThreads) {
      int r = e / BK, kk = e % BK;
      int k = k0 + kk;
      float val = 0.f;
      int n = row_n[r];
      if (k < K && n >= 0) {
        int t, ci;
        if (aligned) { t = t_c; ci = ci_c + kk; } else { t = k / p.Ci; ci = k - t * p.Ci; }
        int iy = row_i[r] * p.is + v.dy[t], ix = row_j[r] * p.is + v.dx[t];
        if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) val = load_in<T>(p, n, iy, ix, ci);
      }
      As[kk][r] = val;
    }
    // B tile: BK x BN_ weights
    for (int e = tid; e < BK * BN_; e += kThreads) {
      int kk = e / BN_, c = e % BN_;
      int k = k0 + kk, co = n0 + c;
      float val = 0.f;
      if (k < K && co < p.Co) {
        int t, ci;
        if (aligned) { t = t_c; ci = ci_c + kk; } else { t = k / p.Ci; ci = k - t * p.Ci; }
        val = __ldg(p.w + v.wofs[t] + (size_t)ci * p.w_sci + (size_t)co * p.w_sco);
      }
      Bs[kk][c] = val;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue: bias, store, per-channel partial statistics.  The statistics are taken over the values
  // as stored (after rounding to the storage type), as a (sum, M2 = sum of squared deviations from the
  // tile mean) pair per CTA: bn_finalize combines the pairs with Chan's formula, so the batch variance
  // has no E[y^2] - mean^2 cancellation (a 1e-6 relative variance error is amplified by var/eps in the
  // BatchNorm backward when only a few values share a channel).
  T* out = reinterpret_cast<T*>(p.out);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = ty * 4 + i;
    int n = row_n[r];
    int oy = v.oy0 + p.os * row_i[r], ox = v.ox0 + p.os * row_j[r];
    bool valid = n >= 0 && oy < p.Ho && ox < p.Wo;   // ragged parity sub-grid of an odd-sized stride-2 dgrad
    size_t base = valid ? ((size_t(n) * p.Ho + oy) * p.Wo + ox) * p.Co : 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      float val = 0.f;
      if (valid && co < p.Co) {
        val = acc[i][j];
        if (p.bias) val += p.bias[co];
        if (p.act) val = act_apply(p.act, val);
        if (p.dact) val *= act_deriv(p.dact_kind, to_f(reinterpret_cast<const T*>(p.dact)[base + co]));
        if (p.accumulate) val += to_f(out[base + co]);
        T tv = from_f<T>(val);
        out[base + co] = tv;
        val = to_f(tv);
      }
      acc[i][j] = val;
    }
  }
  if (p.partials) {
    float* csum = &Bs[0][0];            // BN_ floats each, the B tile is dead by now
    float* cmean = &Bs[1][0];
    const int n_valid = min(BM, p.M - m0);
#pragma unroll
    for (int j = 0; j < 4; ++j) red[0][ty][tx * 4 + j] = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
    __syncthreads();
    for (int c = tid; c < BN_; c += kThreads) {
      float s = 0.f;
      for (int q = 0; q < BM / 4; ++q) s += red[0][q][c];         // fixed order: deterministic
      csum[c] = s; cmean[c] = s / (float)n_valid;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float mu = cmean[tx * 4 + j], s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int r = ty * 4 + i;
        float dlt = (row_n[r] >= 0) ? acc[i][j] - mu : 0.f;
        s = fmaf(dlt, dlt, s);
      }
      red[1][ty][tx * 4 + j] = s;
    }
    __syncthreads();
    for (int c = tid; c < BN_; c += kThreads) {
      int co = n0 + c;
      if (co >= p.Co) continue;
      float s = 0.f;
      for (int q = 0; q < BM / 4; ++q) s += red[1][q][c];
      size_t prow = size_t(blockIdx.z) * gridDim.x + blockIdx.x;
      p.partials[(prow * p.Co + co) * 2 + 0] = csum[c];
      p.partials[(prow * p.Co + co) * 2 + 1] = s;
    }
  }
}

template <typename T>
__device__ __forceinline__ float load_in_w(const WGradParams& p, int n, int iy, int ix, int ci) {
  if (p.in_nchw_f32)
    return reinterpret_cast<const float*>(p.in)[((size_t(n) * p.Ci + ci) * p.Hi + iy) * p.Wi + ix];
  return to_f(reinterpret_cast<const T*>(p.in)[((size_t(n) * p.Hi + iy) * p.Wi + ix) * p.Ci + ci]);
}

// dW tile: BKR (k = t*Ci+ci) x BN_ (co), reduced over a slice of the M rows; BKR*BN_ = 4096.
template <typename T, int BN_>
__global__ void __launch_bounds__(kThreads) wgrad_simt_kernel(const __grid_constant__ WGradParams p) {
  constexpr int BKR = 4096 / BN_;
  constexpr int BR = 16;
  __shared__ __align__(16) float As[BR][BKR + 4];
  __shared__ __align__(16) float Bs[BR][BN_ + 4];

  const int vi = blockIdx.z / p.nsplit, split = blockIdx.z % p.nsplit;
  const GVar& v = p.var[vi];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * BKR, n0 = blockIdx.y * BN_;
  const int K = v.ntaps * p.Ci;
  const int tx = tid % (BN_ / 4), ty = tid / (BN_ / 4);
  const int m_lo = split * p.rows_per_split;
  const int m_hi = min(p.M, m_lo + p.rows_per_split);

  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

  const T* dout = reinterpret_cast<const T*>(p.dout);
  for (int mb = m_lo; mb < m_hi; mb += BR) {
    for (int e = tid; e < BR * BKR; e += kThreads) {
      int mm = e / BKR, kk = e % BKR;
      int m = mb + mm, k = k0 + kk;
      float val = 0.f;
      if (m < m_hi && k < K) {
        int j = m % p.Wg; int tt = m / p.Wg; int i = tt % p.Hg; int n = tt / p.Hg;
        int t = k / p.Ci, ci = k - t * p.Ci;
        int iy = i * p.is + v.dy[t], ix = j * p.is + v.dx[t];
        if (iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) val = load_in_w<T>(p, n, iy, ix, ci);
      }
      As[mm][kk] = val;
    }
    for (int e = tid; e < BR * BN_; e += kThreads) {
      int mm = e / BN_, c = e % BN_;
      int m = mb + mm, co = n0 + c;
      float val = 0.f;
      if (m < m_hi && co < p.Co) {
        int j = m % p.Wg; int tt = m / p.Wg; int i = tt % p.Hg; int n = tt / p.Hg;
        int oy = v.oy0 + p.os * i, ox = v.ox0 + p.os * j;
        if (oy < p.Ho && ox < p.Wo)                 // gather grids rounded up past an odd-sized output (nb_pad_grid)
          val = to_f(dout[((size_t(n) * p.Ho + oy) * p.Wo + ox) * p.Co + co]);
      }
      Bs[mm][c] = val;
    }
    __syncthreads();
#pragma unroll
    for (int mm = 0; mm < BR; ++mm) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[mm][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[mm][tx * 4]);
      float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int k = k0 + ty * 4 + i;
    if (k >= K) continue;
    int t = k / p.Ci, ci = k - t * p.Ci;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int co = n0 + tx * 4 + j;
      if (co >= p.Co) continue;
      atomicAdd(p.dw + v.wofs[t] + (size_t)ci * p.w_sci + (size_t)co * p.w_sco, acc[i][j]);
    }
  }
}

}  // namespace

// returns the layout of the per-CTA partial statistics it wrote
template <typename T>
StatLayout launch_gconv_simt(const GConvParams& p, cudaStream_t st) {
  StatLayout sl{0, 0, 0, 0};
  if (p.M <= 0) return sl;
  dim3 grid;
  if (p.Co <= 16) {
    grid = dim3((p.M + 255) / 256, 1, p.nvar);
    count_launch();
    gconv_simt_kernel<T, 16><<<grid, kThreads, 0, st>>>(p);
  } else if (p.Co <= 32) {
    grid = dim3((p.M + 127) / 128, 1, p.nvar);
    count_launch();
    gconv_simt_kernel<T, 32><<<grid, kThreads, 0, st>>>(p);
  } else {
    grid = dim3((p.M + 63) / 64, (p.Co + 63) / 64, p.nvar);
    count_launch();
    gconv_simt_kernel<T, 64><<<grid, kThreads, 0, st>>>(p);
  }
  sl.parts = (int)(grid.x * grid.z); sl.parts_per_var = (int)grid.x;
  sl.tile_rows = p.Co <= 16 ? 256 : (p.Co <= 32 ? 128 : 64); sl.rows_per_var = p.M;
  return sl;
}

template <typename T>
void launch_wgrad_simt(const WGradParams& p0, cudaStream_t st) {
  if (p0.M <= 0) return;
  WGradParams p = p0;
  int maxK = 0;
  for (int i = 0; i < p.nvar; ++i) maxK = max(maxK, p.var[i].ntaps * p.Ci);
  int bn = p.Co <= 16 ? 16 : (p.Co <= 32 ? 32 : 64);
  int bkr = 4096 / bn;
  int gx = (maxK + bkr - 1) / bkr, gy = (p.Co + bn - 1) / bn;
  int base = gx * gy * p.nvar;
  int nsplit = max(1, 592 / base);
  int max_split = (p.M + 63) / 64;            // at least 64 rows per split
  nsplit = min(nsplit, max(1, max_split));
  int rps = (p.M + nsplit - 1) / nsplit;
  rps = (rps + 15) / 16 * 16;
  nsplit = (p.M + rps - 1) / rps;
  p.nsplit = nsplit; p.rows_per_split = rps;
  dim3 grid(gx, gy, p.nvar * nsplit);
  if (bn == 16) { count_launch(); wgrad_simt_kernel<T, 16><<<grid, kThreads, 0, st>>>(p); }
  else if (bn == 32) { count_launch(); wgrad_simt_kernel<T, 32><<<grid, kThreads, 0, st>>>(p); }
  else { count_launch(); wgrad_simt_kernel<T, 64><<<grid, kThreads, 0, st>>>(p); }
}

template StatLayout launch_gconv_simt<float>(const GConvParams&, cudaStream_t);
template StatLayout launch_gconv_simt<__nv_bfloat16>(const GConvParams&, cudaStream_t);
template void launch_wgrad_simt<float>(const WGradParams&, cudaStream_t);
template void launch_wgrad_simt<__nv_bfloat16>(const WGradParams&, cudaStream_t);

}  // namespace mmvae
