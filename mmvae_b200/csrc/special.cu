// special.cu -- the two layers that are not tensor-core work: the 1-channel stem
// Conv2d(1 -> 32w, k5 s2 p2) (model.py:94) and the 1-channel tail Conv2d(16w -> 1, k3 s1 p1, bias)
// (model.py:172), forward and backward, as HBM-bound SIMT kernels with shared-memory tiles.
//
//   stem_fwd    x fp32 NCHW -> y bf16 NHWC + per-tile BatchNorm partials        (GEMM K = 25)
//   stem_wgrad  dW[co][5][5] = sum_p x(p + tap) * dY[p][co]                     (no dgrad: x needs none)
//   tail_fwd    a bf16 NHWC -> y bf16 [N,H,W,1] + per-tile BatchNorm partials   (GEMM N = 1: a GEMV)
//   tail_bwd    ONE pass over (a, dY): dX[q][ci] and dW[ci][3][3] share the dY neighbourhood of q
//
// All of them walk tiles of whole output rows of one frame; tiles are numbered so that tile i covers
// GEMM rows [i * tile_rows, (i+1) * tile_rows), which is what bn_finalize's StatLayout expects.
#include "bn_fused.cuh"
#include "kernels.cuh"

namespace mmvae {

namespace {

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// ------------------------------------------------------------------------------------------------
// stem
// ------------------------------------------------------------------------------------------------
constexpr int kStemMaxRows = 21, kStemMaxCols = 67;     // (2R+3) x (2Wo+3) input halo tile, R*Wo <= 128, Wo <= 32

__device__ __forceinline__ void stem_load_x(const StemArgs& a, int n, int oy0, float* xs, int rows, int cols, int nthreads) {
  const float* xn = a.x + (size_t)n * a.S * a.S;
  for (int e = threadIdx.x; e < rows * cols; e += nthreads) {
    int r = e / cols, c = e - r * cols;
    int iy = 2 * oy0 - 2 + r, ix = c - 2;
    xs[e] = ((unsigned)iy < (unsigned)a.S && (unsigned)ix < (unsigned)a.S) ? __ldg(xn + (size_t)iy * a.S + ix) : 0.f;
  }
}

template <int CO>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const StemArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int H = CO / 2;                          // channels per thread
  __shared__ float xs[kStemMaxRows * kStemMaxCols];
  __shared__ __align__(16) float ws[25 * CO];
  const int tid = threadIdx.x, lane = tid & 31;
  float run_s[H], run_q[H];
#pragma unroll
  for (int j = 0; j < H; ++j) { run_s[j] = 0.f; run_q[j] = 0.f; }
  for (int e = tid; e < 25 * CO; e += 256) { int co = e / 25, t = e - co * 25; ws[t * CO + co] = __ldg(a.w + e); }
  const int rows = 2 * a.R + 3, cols = 2 * a.Wo + 3;
  const int p = tid >> 1, h = tid & 1;
  const int oy_l = p / a.Wo, ox = p - oy_l * a.Wo;
  const bool in_tile = p < a.R * a.Wo;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
    __syncthreads();                                 // previous tile's readers are done with xs
    stem_load_x(a, n, oy0, xs, rows, cols, 256);
    __syncthreads();
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = 0.f;
    if (in_tile) {
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const float* xr = xs + (2 * oy_l + kh) * cols + 2 * ox;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const float xv = xr[kw];
          const float4* wr = reinterpret_cast<const float4*>(ws + (kh * 5 + kw) * CO + h * H);
#pragma unroll
          for (int j4 = 0; j4 < H / 4; ++j4) {
            float4 w4 = wr[j4];
            acc[4 * j4 + 0] = fmaf(xv, w4.x, acc[4 * j4 + 0]); acc[4 * j4 + 1] = fmaf(xv, w4.y, acc[4 * j4 + 1]);
            acc[4 * j4 + 2] = fmaf(xv, w4.z, acc[4 * j4 + 2]); acc[4 * j4 + 3] = fmaf(xv, w4.w, acc[4 * j4 + 3]);
          }
        }
      }
      const size_t m = ((size_t)n * a.Ho + oy0 + oy_l) * a.Wo + ox;
      uint4* dst = reinterpret_cast<uint4*>(a.y + m * CO + h * H);
#pragma unroll
      for (int q = 0; q < H / 8; ++q)
        dst[q] = make_uint4(pack2(acc[8 * q], acc[8 * q + 1]), pack2(acc[8 * q + 2], acc[8 * q + 3]),
                            pack2(acc[8 * q + 4], acc[8 * q + 5]), pack2(acc[8 * q + 6], acc[8 * q + 7]));
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = bf16_round(acc[j]);      // statistics over the values as stored
    }
    if (a.bn.acc) {
#pragma unroll
      for (int j = 0; j < H; ++j) { run_s[j] += acc[j]; run_q[j] = fmaf(acc[j], acc[j], run_q[j]); }
    }
  }
  if (a.bn.acc) {
    // per-channel sums over this CTA's pixels: lanes of equal parity hold the same channel half
    __shared__ float wred[8][2][CO];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      float v = run_s[j], w = run_q[j];
#pragma unroll
      for (int d = 2; d < 32; d <<= 1) { v += __shfl_xor_sync(0xffffffffu, v, d); w += __shfl_xor_sync(0xffffffffu, w, d); }
      if (lane < 2) { wred[tid >> 5][0][h * H + j] = v; wred[tid >> 5][1][h * H + j] = w; }
    }
    __syncthreads();
    if (tid < 2 * CO) {
      const int which = tid / CO, c = tid - which * CO;
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += wred[w8][which][c];
      atomicAdd(bn_acc_copy(a.bn) + which * CO + c, (double)t);
    }
    bn_fused_finish(a.bn, gridDim.x);
  }
}

// thread = (pixel slice, kernel row kh, 4 output channels): 5 x 4 accumulators kept across all tiles of the CTA
template <int CO>
__global__ void __launch_bounds__(320) stem_wgrad_kernel(const StemArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int CG = CO / 4;                         // channel groups
  constexpr int PER = 5 * CG;                        // threads per pixel slice
  constexpr int NSL = 320 / PER;                     // pixel slices
  __shared__ float xs[kStemMaxRows * kStemMaxCols];
  __shared__ __align__(16) __nv_bfloat16 dys[128 * CO];
  __shared__ float red[NSL][25 * CO];
  const int tid = threadIdx.x;
  const int sl = tid / PER, rem = tid - sl * PER, kh = rem / CG, cg = rem - kh * CG;
  const bool active = sl < NSL;
  const int rows = 2 * a.R + 3, cols = 2 * a.Wo + 3, npx = a.R * a.Wo;
  float acc[5][4];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
    __syncthreads();
    stem_load_x(a, n, oy0, xs, rows, cols, 320);
    const uint4* src = reinterpret_cast<const uint4*>(a.dy + ((size_t)n * a.Ho + oy0) * a.Wo * CO);
    for (int e = tid; e < npx * CO / 8; e += 320) reinterpret_cast<uint4*>(dys)[e] = __ldg(src + e);
    __syncthreads();
    if (active) {
      for (int p = sl; p < npx; p += NSL) {
        const int oy_l = p / a.Wo, ox = p - oy_l * a.Wo;
        const uint2 g2 = *reinterpret_cast<const uint2*>(dys + p * CO + cg * 4);
        const float2 g01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g2.x));
        const float2 g23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&g2.y));
        const float* xr = xs + (2 * oy_l + kh) * cols + 2 * ox;
#pragma unroll
        for (int kw = 0; kw < 5; ++kw) {
          const float xv = xr[kw];
          acc[kw][0] = fmaf(xv, g01.x, acc[kw][0]); acc[kw][1] = fmaf(xv, g01.y, acc[kw][1]);
          acc[kw][2] = fmaf(xv, g23.x, acc[kw][2]); acc[kw][3] = fmaf(xv, g23.y, acc[kw][3]);
        }
      }
    }
  }
  __syncthreads();
  if (active) {
#pragma unroll
    for (int kw = 0; kw < 5; ++kw)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[sl][(kh * 5 + kw) * CO + cg * 4 + j] = acc[kw][j];
  }
  __syncthreads();
  for (int e = tid; e < 25 * CO; e += 320) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < NSL; ++q) s += red[q][e];
    const int tap = e / CO, co = e - tap * CO;
    atomicAdd(a.dw + co * 25 + tap, s);
  }
}

// ------------------------------------------------------------------------------------------------
// tail
// ------------------------------------------------------------------------------------------------
constexpr int kTailThreads = 512;

__device__ __forceinline__ float block_sum512(float v, float* sh) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kTailThreads / 32; ++w) r += sh[w];
  return r;
}

// thread = (pixel, group of 8 input channels); the CI/8 lanes of a pixel are adjacent
template <int CI>
__global__ void __launch_bounds__(kTailThreads) tail_fwd_kernel(const TailArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int G = CI / 8;
  constexpr int PX = kTailThreads / G;               // pixels per tile
  extern __shared__ __align__(16) unsigned char tail_smem[];
  uint4* tile = reinterpret_cast<uint4*>(tail_smem);                 // [(R+2)][(W+2)][G] 16-byte chunks
  __shared__ float ws[9 * CI];
  __shared__ float sh[kTailThreads / 32];
  const int tid = threadIdx.x;
  for (int e = tid; e < 9 * CI; e += kTailThreads) { int ci = e / 9, t = e - ci * 9; ws[t * CI + ci] = __ldg(a.w + e); }
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  const int W2 = a.W + 2;
  const int p = tid / G, grp = tid - p * G;
  const int oy_l = p / a.W, ox = p - oy_l * a.W;
  const uint4* in = reinterpret_cast<const uint4*>(a.in);
  float run_s = 0.f, run_q = 0.f;
  for (int t_i = blockIdx.x; t_i < a.ntiles; t_i += gridDim.x) {
    const int n = t_i / a.tiles_per_frame, oy0 = (t_i - n * a.tiles_per_frame) * a.R;
    __syncthreads();
    for (int e = tid; e < (a.R + 2) * W2 * G; e += kTailThreads) {
      int pix = e / G, c = e - pix * G;
      int r = pix / W2, cc = pix - r * W2;
      int iy = oy0 - 1 + r, ix = cc - 1;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if ((unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W) v = __ldg(in + (((size_t)n * a.H + iy) * a.W + ix) * G + c);
      tile[e] = v;
    }
    __syncthreads();
    float acc = 0.f;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        float f[8];
        unpack8(tile[((oy_l + kh) * W2 + ox + kw) * G + grp], f);
        const float* wr = ws + (kh * 3 + kw) * CI + grp * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(f[i], wr[i], acc);
      }
#pragma unroll
    for (int d = 1; d < G; d <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    acc += bias;
    const size_t m = ((size_t)n * a.H + oy0 + oy_l) * a.W + ox;
    const __nv_bfloat16 o = __float2bfloat16_rn(acc);
    if (grp == 0) a.y[m] = o;
    if (a.bn.acc && grp == 0) { const float v = __bfloat162float(o); run_s += v; run_q = fmaf(v, v, run_q); }
  }
  if (a.bn.acc) {
    const float s = block_sum512(run_s, sh);
    const float q = block_sum512(run_q, sh);
    if (tid == 0) { atomicAdd(bn_acc_copy(a.bn), (double)s); atomicAdd(bn_acc_copy(a.bn) + 1, (double)q); }
    bn_fused_finish(a.bn, gridDim.x);
  }
}

// One pass over the tail conv's input activation a and its output gradient dY:
//   dX[q][ci]      = sum_t dY[q + (1-kh, 1-kw)] * w[ci][t]
//   dW[ci][t]     += a[q][ci] * dY[q + (1-kh, 1-kw)]
template <int CI>
__global__ void __launch_bounds__(kTailThreads) tail_bwd_kernel(const TailArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int G = CI / 8;
  extern __shared__ __align__(16) unsigned char tail_smem[];
  float* gs = reinterpret_cast<float*>(tail_smem);                   // [(R+2)][(W+2)] dY with zero halo
  float* red = gs + (a.R + 2) * (a.W + 2);                           // [16 warps][G][72]
  __shared__ float ws[9 * CI];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int e = tid; e < 9 * CI; e += kTailThreads) ws[e] = __ldg(a.w + e);        // [ci][t]
  const int W2 = a.W + 2;
  const int p = tid / G, grp = tid - p * G;
  const int oy_l = p / a.W, ox = p - oy_l * a.W;
  const uint4* in = reinterpret_cast<const uint4*>(a.in);
  uint4* dx = reinterpret_cast<uint4*>(a.dx);
  float accw[72];                                    // [8 channels][9 taps]
#pragma unroll
  for (int i = 0; i < 72; ++i) accw[i] = 0.f;
  for (int t_i = blockIdx.x; t_i < a.ntiles; t_i += gridDim.x) {
    const int n = t_i / a.tiles_per_frame, oy0 = (t_i - n * a.tiles_per_frame) * a.R;
    __syncthreads();
    for (int e = tid; e < (a.R + 2) * W2; e += kTailThreads) {
      int r = e / W2, cc = e - r * W2;
      int iy = oy0 - 1 + r, ix = cc - 1;
      gs[e] = ((unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W)
                  ? __bfloat162float(a.dy[((size_t)n * a.H + iy) * a.W + ix]) : 0.f;
    }
    __syncthreads();
    const size_t q = (((size_t)n * a.H + oy0 + oy_l) * a.W + ox) * G + grp;
    float f[8];
    unpack8(__ldg(in + q), f);
    float g[9];
#pragma unroll
    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) g[kh * 3 + kw] = gs[(oy_l + 2 - kh) * W2 + ox + 2 - kw];
    float o[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float* wr = ws + (grp * 8 + c) * 9;
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) { s = fmaf(g[t], wr[t], s); accw[c * 9 + t] = fmaf(f[c], g[t], accw[c * 9 + t]); }
      o[c] = s;
    }
    dx[q] = make_uint4(pack2(o[0], o[1]), pack2(o[2], o[3]), pack2(o[4], o[5]), pack2(o[6], o[7]));
  }
  // reduce the 72 accumulators over the pixels of the warp (lanes of equal grp), halving the live set per step
  constexpr int LB = G == 1 ? 0 : (G == 2 ? 1 : 2);  // lane bits that index grp
  int base = 0;
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 36; ++i) {
      float send = hi ? accw[i] : accw[i + 36], keep = hi ? accw[i + 36] : accw[i];
      accw[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
    base += hi ? 36 : 0;
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      float send = hi ? accw[i] : accw[i + 18], keep = hi ? accw[i + 18] : accw[i];
      accw[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    base += hi ? 18 : 0;
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      float send = hi ? accw[i] : accw[i + 9], keep = hi ? accw[i + 9] : accw[i];
      accw[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    base += hi ? 9 : 0;
  }
#pragma unroll
  for (int d = 2; d >= (1 << LB); d >>= 1) {
#pragma unroll
    for (int i = 0; i < 9; ++i) accw[i] += __shfl_xor_sync(0xffffffffu, accw[i], d);
  }
  __syncthreads();
  if ((lane & 3 & ~((1 << LB) - 1)) == 0) {          // one lane per (grp, base) inside the warp
#pragma unroll
    for (int i = 0; i < 9; ++i) red[(warp * G + grp) * 72 + base + i] = accw[i];
  }
  __syncthreads();
  for (int e = tid; e < G * 72; e += kTailThreads) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kTailThreads / 32; ++w) s += red[w * G * 72 + e];
    atomicAdd(a.dw + e, s);                          // e = (grp*8 + c)*9 + t == weight index [0][ci][kh][kw]
  }
}

// dW_head[zc][c] = sum_n dhead[n][zc] * pooled[n][c].  Block = 8 latent rows x 64 channels, 4 slices of the batch;
// blockIdx.y selects mu / logvar, blockIdx.z the channel block.
__global__ void __launch_bounds__(256) heads_wgrad2_kernel(const float* __restrict__ dh, const float* __restrict__ pooled,
                                                           float* __restrict__ g_mu, float* __restrict__ g_lv,
                                                           int N, int z, int C) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ float red[4][8][64];
  const float* d = dh + (size_t)blockIdx.y * N * z;
  float* gw = blockIdx.y == 0 ? g_mu : g_lv;
  const int zc0 = blockIdx.x * 8;
  const int cl = threadIdx.x & 63, sl = threadIdx.x >> 6;
  const int c = blockIdx.z * 64 + cl;
  float s[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = 0.f;
  if (c < C) {
#pragma unroll 4
    for (int n = sl; n < N; n += 4) {
      const float pv = __ldg(pooled + (size_t)n * C + c);
      const float* dr = d + (size_t)n * z + zc0;
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = fmaf(zc0 + j < z ? __ldg(dr + j) : 0.f, pv, s[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[sl][j][cl] = s[j];
  __syncthreads();
  for (int e = threadIdx.x; e < 8 * 64; e += 256) {
    const int j = e >> 6, cc = e & 63;
    const int co = blockIdx.z * 64 + cc;
    if (zc0 + j < z && co < C)
      gw[(size_t)(zc0 + j) * C + co] = (red[0][j][cc] + red[1][j][cc]) + (red[2][j][cc] + red[3][j][cc]);
  }
}

}  // namespace

// R = the largest divisor of Ho with R * Wo <= cap pixels
static int rows_per_tile(int Ho, int Wo, int cap) {
  int best = 0;
  for (int r = 1; r <= Ho; ++r)
    if (Ho % r == 0 && r * Wo <= cap) best = r;
  return best;
}

bool stem_supported(int Cin, int Co, int S, int k, int s, int p) {
  if (Cin != 1 || k != 5 || s != 2 || p != 2 || (Co != 32 && Co != 64) || (S & 1)) return false;
  const int Ho = S / 2;
  if (Ho > 32) return false;
  return rows_per_tile(Ho, Ho, 128) > 0;
}

static void stem_fill(StemArgs& a) {
  a.Ho = a.S / 2; a.Wo = a.S / 2;
  a.R = rows_per_tile(a.Ho, a.Wo, 128);
  a.tiles_per_frame = a.Ho / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
}

StatLayout launch_stem_fwd(StemArgs a, int Co, cudaStream_t st) {
  stem_fill(a);
  const int grid = min(a.ntiles, 148 * 3);
  count_launch();
  if (Co == 32) launch_pdl(stem_fwd_kernel<32>, grid, 256, 0, st, a);
  else launch_pdl(stem_fwd_kernel<64>, grid, 256, 0, st, a);
  return StatLayout{0, 0, 0, 0};     // statistics are finalised inside the kernel (a.bn)
}

void launch_stem_wgrad(StemArgs a, int Co, cudaStream_t st) {
  stem_fill(a);
  const int grid = min(a.ntiles, 2 * 148);
  count_launch();
  if (Co == 32) launch_pdl(stem_wgrad_kernel<32>, grid, 320, 0, st, a);
  else launch_pdl(stem_wgrad_kernel<64>, grid, 320, 0, st, a);
}

bool tail_supported(int Ci, int Co, int H, int k, int s, int p) {
  if (Co != 1 || k != 3 || s != 1 || p != 1 || (Ci != 16 && Ci != 32)) return false;
  const int px = kTailThreads / (Ci / 8);
  const int R = rows_per_tile(H, H, px);
  return R > 0 && R * H == px;                       // every thread owns exactly one (pixel, channel group)
}

static void tail_fill(TailArgs& a, int Ci) {
  a.R = rows_per_tile(a.H, a.W, kTailThreads / (Ci / 8));
  a.tiles_per_frame = a.H / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
}

StatLayout launch_tail_fwd(TailArgs a, int Ci, cudaStream_t st) {
  tail_fill(a, Ci);
  const int grid = min(a.ntiles, 148 * 3);
  const size_t smem = (size_t)(a.R + 2) * (a.W + 2) * Ci * 2;
  count_launch();
  if (Ci == 16) launch_pdl(tail_fwd_kernel<16>, grid, kTailThreads, smem, st, a);
  else launch_pdl(tail_fwd_kernel<32>, grid, kTailThreads, smem, st, a);
  return StatLayout{0, 0, 0, 0};     // statistics are finalised inside the kernel (a.bn)
}

void launch_tail_bwd(TailArgs a, int Ci, cudaStream_t st) {
  tail_fill(a, Ci);
  const int grid = min(a.ntiles, 148);
  const size_t smem = sizeof(float) * ((size_t)(a.R + 2) * (a.W + 2) + (size_t)(kTailThreads / 32) * (Ci / 8) * 72);
  count_launch();
  if (Ci == 16) launch_pdl(tail_bwd_kernel<16>, grid, kTailThreads, smem, st, a);
  else launch_pdl(tail_bwd_kernel<32>, grid, kTailThreads, smem, st, a);
}

void launch_heads_wgrad(const float* dheads, const float* pooled, float* g_mu, float* g_lv, int N, int z, int C,
                        cudaStream_t st) {
  dim3 grid((z + 7) / 8, g_lv ? 2 : 1, (C + 63) / 64);
  count_launch();
  launch_pdl(heads_wgrad2_kernel, grid, 256, 0, st, dheads, pooled, g_mu, g_lv, N, z, C);
}

}  // namespace mmvae
