// special.cu -- the two layers that are not tensor-core work: the 1-channel stem
// Conv2d(1 -> 32w, k5 s2 p2) (model.py:94) and the 1-channel tail Conv2d(16w -> 1, k3 s1 p1, bias)
// (model.py:172), forward and backward, as register-blocked fp32 SIMT kernels over shared-memory tiles.
//
//   stem_fwd    x fp32 NCHW -> y bf16 NHWC + fused BatchNorm statistics          (GEMM K = 25)
//               thread = 4 consecutive output pixels x 8 channels: a weight vector is reused by 4 pixels
//   stem_wgrad  dW[co][5][5] = sum_p x(p + tap) * dY[p][co]                      (no dgrad: x needs none)
//               warp = (kernel row, 8 channels), lane = pixel slice; cp.async double-buffered tiles
//   tail_fwd    a bf16 NHWC -> y bf16 [N,H,W,1] + fused BatchNorm statistics     (GEMM N = 1: a GEMV)
//               thread = a vertical strip of 8 pixels x 8 channels, weights in registers, sliding 3-row window
//   tail_bwd    ONE pass over (a, dY): dX[q][ci], dW[ci][3][3], and -- the tensor dX is the gradient of the last decoder
//               block's output -- that block's ReLU mask and BatchNorm-backward sums (BnBwdFused)
//
// Arithmetic stays fp32 on fp32 x / fp32 weights (the tensor cores would round x and w to bf16).
#include "bn_fused.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

namespace {

using tc::cp_async16;
using tc::cp_async_commit;
using tc::cp_async_wait;
using tc::smem_u32;

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf_lo(u.x); f[1] = bf_hi(u.x); f[2] = bf_lo(u.y); f[3] = bf_hi(u.y);
  f[4] = bf_lo(u.z); f[5] = bf_hi(u.z); f[6] = bf_lo(u.w); f[7] = bf_hi(u.w);
}
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

// ------------------------------------------------------------------------------------------------
// stem
// ------------------------------------------------------------------------------------------------
constexpr int kStemXs = 1408;          // floats: (2R+3) rows x pitch of the x halo tile (19 x 68 or 35 x 36)

__device__ __forceinline__ void stem_x_prefetch(const StemArgs& a, int tile, float* xs, int nthreads);

template <int CO>
__global__ void __launch_bounds__(256, 3) stem_fwd_kernel(const StemArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int CG = CO / 8;                         // channel groups of 8
  __shared__ __align__(16) float xs2[2][kStemXs];      // double-buffered x halo tile (cp.async; the next tile loads under this one)
  __shared__ __align__(16) float ws[25 * CO];
  __shared__ float wred[8][2][CO];
  const int tid = threadIdx.x, lane = tid & 31;
  const int cg = tid % CG, quad = tid / CG;
  const int r = quad / a.qpr, xq = quad - r * a.qpr;
  const bool active = r < a.R;
  float run_s[8], run_q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { run_s[j] = 0.f; run_q[j] = 0.f; }
  for (int e = tid; e < 25 * CO; e += 256) { int co = e / 25, t = e - co * 25; ws[t * CO + co] = __ldg(a.w + e); }
  for (int e = tid; e < 2 * kStemXs; e += 256) (&xs2[0][0])[e] = 0.f;            // halo columns stay zero
  __syncthreads();
  const int pitch = a.pitch;
  int buf = 0;
  if ((int)blockIdx.x < a.ntiles) stem_x_prefetch(a, blockIdx.x, xs2[0], 256);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, buf ^= 1) {
    const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
    const int next = tile + gridDim.x;
    __syncthreads();                                 // the previous tile's readers are done with the buffer refilled now
    if (next < a.ntiles) stem_x_prefetch(a, next, xs2[buf ^ 1], 256);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const float* xs = xs2[buf];
    if (!active) continue;
    float acc[4][8];
#pragma unroll
    for (int px = 0; px < 4; ++px)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[px][c] = 0.f;
#pragma unroll
    for (int kh = 0; kh < 5; ++kh) {
      const float* xr = xs + (2 * r + kh) * pitch + 8 * xq;           // input column 2*ox - 2 lives at tile column 2*ox
      float xv[12];
      *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(xr);
      *reinterpret_cast<float4*>(xv + 4) = *reinterpret_cast<const float4*>(xr + 4);
      *reinterpret_cast<float4*>(xv + 8) = *reinterpret_cast<const float4*>(xr + 8);
#pragma unroll
      for (int kw = 0; kw < 5; ++kw) {
        const float4 w0 = *reinterpret_cast<const float4*>(ws + (kh * 5 + kw) * CO + cg * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + (kh * 5 + kw) * CO + cg * 8 + 4);
#pragma unroll
        for (int px = 0; px < 4; ++px) {
          const float v = xv[2 * px + kw];
          acc[px][0] = fmaf(v, w0.x, acc[px][0]); acc[px][1] = fmaf(v, w0.y, acc[px][1]);
          acc[px][2] = fmaf(v, w0.z, acc[px][2]); acc[px][3] = fmaf(v, w0.w, acc[px][3]);
          acc[px][4] = fmaf(v, w1.x, acc[px][4]); acc[px][5] = fmaf(v, w1.y, acc[px][5]);
          acc[px][6] = fmaf(v, w1.z, acc[px][6]); acc[px][7] = fmaf(v, w1.w, acc[px][7]);
        }
      }
    }
    const int oy = oy0 + r;
    if (oy < a.Ho) {
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const int ox = 4 * xq + px;
        if (ox < a.Wo) {
          const size_t m = ((size_t)n * a.Ho + oy) * a.Wo + ox;
          *reinterpret_cast<uint4*>(a.y + m * CO + cg * 8) =
              make_uint4(pack2(acc[px][0], acc[px][1]), pack2(acc[px][2], acc[px][3]), pack2(acc[px][4], acc[px][5]),
                         pack2(acc[px][6], acc[px][7]));
          if (a.bn.acc) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {            // statistics over the values as stored
              const float v = bf16_round(acc[px][c]);
              run_s[c] += v; run_q[c] = fmaf(v, v, run_q[c]);
            }
          }
        }
      }
    }
  }
  cp_async_wait<0>();
  if (a.bn.acc) {
    // per-channel sums over this CTA's pixels: lanes with equal (lane % CG) hold the same channel group
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = run_s[c], w = run_q[c];
#pragma unroll
      for (int d = CG; d < 32; d <<= 1) { v += __shfl_xor_sync(0xffffffffu, v, d); w += __shfl_xor_sync(0xffffffffu, w, d); }
      if (lane < CG) { wred[tid >> 5][0][lane * 8 + c] = v; wred[tid >> 5][1][lane * 8 + c] = w; }
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

// Weight gradient of the stem: 20 warps = (kernel row kh, 8 channels), the 32 lanes of a warp are 32 pixel slices of
// the tile; each thread keeps 5 (kw) x 8 (channels) accumulators over all tiles of the CTA.  blockIdx.y selects a block
// of 32 output channels.  Tiles (x halo rows + the bf16 dY rows) are double-buffered with cp.async.
constexpr int kStemWgThreads = 640;
constexpr int kStemDyBytes = 256 * 64;               // 256 pixels x 32 channels x 2 bytes

// x tile of a stem tile: (2R+3) input rows as 8-byte pairs, cp.async; rows outside the image are zero-filled, the halo
// columns (-2, -1, S, S+1, ..) of the tile are zeroed once by the kernel and never written here.
__device__ __forceinline__ void stem_x_prefetch(const StemArgs& a, int tile, float* xs, int nthreads) {
  const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
  const int rows = 2 * a.R + 3, pitch = a.pitch, tid = threadIdx.x;
  const float* xn = a.x + (size_t)n * a.S * a.S;
  const int halfp = a.S >> 1;                        // 8-byte pairs per input row
  if ((halfp & (halfp - 1)) == 0) {                  // power of two (S = 32, 64): shifts only
    const int j = tid & (halfp - 1), r0 = tid / halfp, rstep = nthreads / halfp;
    const uint32_t d0 = smem_u32(xs + 2 + 2 * j);
    for (int rr = r0; rr < rows; rr += rstep) {
      const int iy = 2 * oy0 - 2 + rr;
      const bool ok = (unsigned)iy < (unsigned)a.S;
      cp_async8(d0 + (uint32_t)(rr * pitch) * 4u, ok ? (const void*)(xn + (size_t)iy * a.S + 2 * j) : (const void*)xn, ok ? 8u : 0u);
    }
    return;
  }
  for (int e = tid; e < rows * halfp; e += nthreads) {
    const int rr = e / halfp, j = e - rr * halfp;
    const int iy = 2 * oy0 - 2 + rr;
    const bool ok = (unsigned)iy < (unsigned)a.S;
    cp_async8(smem_u32(xs + rr * pitch + 2 + 2 * j), ok ? (const void*)(xn + (size_t)iy * a.S + 2 * j) : (const void*)xn, ok ? 8u : 0u);
  }
}

__device__ __forceinline__ void stem_wgrad_prefetch(const StemArgs& a, int tile, int co_base, int CO, float* xs, unsigned char* dys) {
  const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
  const int tid = threadIdx.x;
  stem_x_prefetch(a, tile, xs, kStemWgThreads);
  const int rv = min(a.R, a.Ho - oy0);               // valid output rows of this tile
  const int npx = rv * a.Wo;
  const unsigned char* src = reinterpret_cast<const unsigned char*>(a.dy + (((size_t)n * a.Ho + oy0) * a.Wo) * CO + co_base);
  for (int e = tid; e < npx * 4; e += kStemWgThreads) {
    const int p = e >> 2, c = e & 3;
    cp_async16(smem_u32(dys + p * 64 + ((c ^ ((p >> 1) & 3)) << 4)), src + (size_t)p * CO * 2 + c * 16, 16u);
  }
}

__global__ void __launch_bounds__(kStemWgThreads, 1) stem_wgrad_kernel(const StemArgs a, int CO) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  __shared__ __align__(16) float xs[2][kStemXs];
  __shared__ __align__(16) unsigned char dys[2][kStemDyBytes];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kh = warp >> 2, cg = warp & 3;
  const int co_base = blockIdx.y * 32;
  const int pitch = a.pitch;
  for (int e = tid; e < 2 * kStemXs; e += kStemWgThreads) (&xs[0][0])[e] = 0.f;      // halo columns stay zero
  __syncthreads();
  float acc[5][8];
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
  int buf = 0;
  if ((int)blockIdx.x < a.ntiles) stem_wgrad_prefetch(a, blockIdx.x, co_base, CO, xs[0], dys[0]);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, buf ^= 1) {
    const int next = tile + gridDim.x;
    if (next < a.ntiles) stem_wgrad_prefetch(a, next, co_base, CO, xs[buf ^ 1], dys[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
    const int npx = min(a.R, a.Ho - oy0) * a.Wo;
    const float* xb = xs[buf];
    const unsigned char* db = dys[buf];
    if (a.Wo == 32) {
      // one output row per pass of the warp: pixel p = 32 * i + lane, so the row index is i, the column is the lane and the
      // swizzle term of the dY chunk is a per-lane constant -- the loop is two pointer increments, 4 loads, 40 FMAs
      const unsigned char* dp = db + lane * 64 + ((cg ^ ((lane >> 1) & 3)) << 4);
      const float* xr = xb + kh * pitch + 2 * lane;
      const int nrow = npx >> 5;
#pragma unroll 4
      for (int i = 0; i < nrow; ++i, dp += 2048, xr += 2 * pitch) {
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(dp), g);
        const float2 x01 = *reinterpret_cast<const float2*>(xr), x23 = *reinterpret_cast<const float2*>(xr + 2);
        const float xv[5] = {x01.x, x01.y, x23.x, x23.y, xr[4]};
#pragma unroll
        for (int kw = 0; kw < 5; ++kw)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[kw][c] = fmaf(xv[kw], g[c], acc[kw][c]);
      }
    } else {
#pragma unroll 4
      for (int p = lane; p < npx; p += 32) {
        int oy_l, ox;
        a.fd_wo.divmod(p, oy_l, ox);
        float g[8];
        unpack8(*reinterpret_cast<const uint4*>(db + p * 64 + ((cg ^ ((p >> 1) & 3)) << 4)), g);
        const float* xr = xb + (2 * oy_l + kh) * pitch + 2 * ox;
        const float2 x01 = *reinterpret_cast<const float2*>(xr), x23 = *reinterpret_cast<const float2*>(xr + 2);
        const float xv[5] = {x01.x, x01.y, x23.x, x23.y, xr[4]};
#pragma unroll
        for (int kw = 0; kw < 5; ++kw)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[kw][c] = fmaf(xv[kw], g[c], acc[kw][c]);
      }
    }
    __syncthreads();                                 // everyone is done with this buffer before it is refilled
  }
  cp_async_wait<0>();
#pragma unroll
  for (int kw = 0; kw < 5; ++kw)
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v = acc[kw][c];
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
      if (lane == 0) atomicAdd(a.dw + (size_t)(co_base + cg * 8 + c) * 25 + kh * 5 + kw, v);
    }
}

// ------------------------------------------------------------------------------------------------
// tail
// ------------------------------------------------------------------------------------------------
// Forward.  Thread = (image column x, group g of 8 input channels) of one band of 8 output rows; the G = CI/8 lanes of
// a pixel are adjacent, so a quarter warp reads 128 contiguous bytes of the tile.  The thread walks the band's 10 input
// rows once (3 x 16-byte loads per row, converted once) and keeps its 9 x 8 weights in registers.
// Tile loader: the (R+2) x W interior pixels of a tile as 16-byte chunks, cp.async.  A row holds W*G chunks (a power of two
// dividing 256, tail_supported): thread -> (row of this pass, chunk of the row) by shifts only; rows outside the image are
// zero-filled; the two halo COLUMNS of the tile are zeroed once by the kernel and never written here.  (The first version
// derived (row, column, group) of every chunk from a flat index with two runtime divisions: 66 % of the kernel's
// instructions were not FMAs, profiles/r02_pointwise_loss_hbm.md.)
template <int CI>
__device__ __forceinline__ void tail_prefetch(const TailArgs& a, int tile, unsigned char* dst, int nthreads) {
  constexpr int G = CI / 8;
  const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
  const int W2 = a.W + 2;
  const int cpr = a.W * G;                           // chunks per row
  const int idx = threadIdx.x & (cpr - 1), rsub = threadIdx.x / cpr, rstep = nthreads / cpr;
  const uint4* in = reinterpret_cast<const uint4*>(a.in) + (size_t)n * a.H * cpr + idx;
  const uint32_t d0 = smem_u32(dst) + (uint32_t)(G + idx) * 16u;          // column 1 of a tile row
  for (int rr = rsub; rr < a.R + 2; rr += rstep) {
    const int iy = oy0 - 1 + rr;
    const bool ok = (unsigned)iy < (unsigned)a.H;
    cp_async16(d0 + (uint32_t)(rr * W2 * G) * 16u, ok ? (const void*)(in + (size_t)iy * cpr) : (const void*)a.in, ok ? 16u : 0u);
  }
}

template <int CI>
__global__ void __launch_bounds__(256, 2) tail_fwd_kernel(const TailArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int G = CI / 8;
  extern __shared__ __align__(16) unsigned char tail_smem[];         // 2 x [(R+2)][(W+2)][G] 16-byte chunks
  __shared__ float sh[2][8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W2 = a.W + 2;
  const size_t tile_bytes = (size_t)(a.R + 2) * W2 * G * 16;
  const int g = tid % G, x = (tid / G) % a.W, band = tid / (G * a.W);
  float wreg[9][8];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int c = 0; c < 8; ++c) wreg[t][c] = __ldg(a.w + (size_t)(g * 8 + c) * 9 + t);
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  float run_s = 0.f, run_q = 0.f;
  int buf = 0;
  // zero halo columns of both buffers (the loader only writes the interior)
  for (int e = tid; e < 2 * (a.R + 2) * 2 * G; e += 256) {
    const int b = e / ((a.R + 2) * 2 * G), r2 = e - b * ((a.R + 2) * 2 * G);
    const int rr = r2 / (2 * G), side = (r2 / G) & 1, c = r2 % G;
    reinterpret_cast<uint4*>(tail_smem + b * tile_bytes)[((size_t)rr * W2 + (side ? W2 - 1 : 0)) * G + c] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  if ((int)blockIdx.x < a.ntiles) tail_prefetch<CI>(a, blockIdx.x, tail_smem, 256);
  cp_async_commit();
  for (int t_i = blockIdx.x; t_i < a.ntiles; t_i += gridDim.x, buf ^= 1) {
    const int next = t_i + gridDim.x;
    if (next < a.ntiles) tail_prefetch<CI>(a, next, tail_smem + (buf ^ 1) * tile_bytes, 256);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int n = t_i / a.tiles_per_frame, oy0 = (t_i - n * a.tiles_per_frame) * a.R;
    const uint4* tile = reinterpret_cast<const uint4*>(tail_smem + buf * tile_bytes);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
    for (int rr = 0; rr < 10; ++rr) {                // input row y0 - 1 + rr feeds output rows rr - kh
      const uint4* row = tile + ((size_t)(band * 8 + rr) * W2 + x) * G + g;
      float v[3][8];
      unpack8(row[0], v[0]); unpack8(row[G], v[1]); unpack8(row[2 * G], v[2]);
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
        const int o = rr - kh;
        if (o >= 0 && o < 8) {
#pragma unroll
          for (int kw = 0; kw < 3; ++kw)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[o] = fmaf(v[kw][c], wreg[kh * 3 + kw][c], acc[o]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
#pragma unroll
      for (int d = 1; d < G; d <<= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], d);
      const int oy = oy0 + band * 8 + o;
      if (g == 0 && oy < a.H) {
        const __nv_bfloat16 ob = __float2bfloat16_rn(acc[o] + bias);
        a.y[((size_t)n * a.H + oy) * a.W + x] = ob;
        const float v = __bfloat162float(ob);
        run_s += v; run_q = fmaf(v, v, run_q);
      }
    }
    __syncthreads();                                 // everyone is done with this buffer before it is refilled
  }
  cp_async_wait<0>();
  if (a.bn.acc) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { run_s += __shfl_xor_sync(0xffffffffu, run_s, d); run_q += __shfl_xor_sync(0xffffffffu, run_q, d); }
    if (lane == 0) { sh[0][warp] = run_s; sh[1][warp] = run_q; }
    __syncthreads();
    if (tid < 2) {
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < 8; ++w8) t += sh[tid][w8];
      atomicAdd(bn_acc_copy(a.bn) + tid, (double)t);
    }
    bn_fused_finish(a.bn, gridDim.x);
  }
}

// Backward: one pass over the tail conv's input activation a and its output gradient dY.
//   dX[q][ci]      = sum_t dY[q + (1-kh, 1-kw)] * w[ci][t]
//   dW[ci][t]     += a[q][ci] * dY[q + (1-kh, 1-kw)]
// and, when bb.acc is set (dX is the gradient of the last decoder block's output a = relu(bn2(y) + bn_s(y2))):
//   g = bf16(dX) * [a > 0] is what gets stored;  S0 = sum g, S1 = sum g * xhat(y), S2 = sum g * xhat(y2) per channel.
// Thread = (image column x, chunk of 4 channels) of a band of 8 rows; the CI/4 lanes of a pixel are adjacent, so a warp
// reads / writes 256 contiguous bytes of every NHWC tensor.  Weights (4 x 9) and dW accumulators (4 x 9) in registers.
template <int CI, bool kColBlocks>
__global__ void __launch_bounds__(256, 2) tail_bwd_kernel(const TailArgs a) {
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
  constexpr int CH = CI / 4;                         // 4-channel chunks per pixel
  constexpr int NACC = 36 + 12;                      // dW accumulators + (S0, S1', S2') x 4 channels
  extern __shared__ __align__(16) unsigned char tail_smem[];
  float* gs = reinterpret_cast<float*>(tail_smem);                   // [(R+2)][(W+2)] dY with zero halo
  float* red = gs + (a.R + 2) * (a.W + 2);                           // [8 warps][CH][NACC]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int W2 = a.W + 2;
  // a row of W pixels x CH chunks that does not fit the 256 threads (W = 64 with 32 channels: the widened model) is walked in
  // column blocks of 256 / CH pixels, one band of 8 rows per CTA tile
  const int cols = kColBlocks ? 256 / CH : a.W;
  const int ch = tid % CH, xl = (tid / CH) % cols, band = tid / (CH * cols);
  const bool fuse = a.bb.acc != nullptr;
  const bool two = fuse && a.bb.y2 != nullptr;
  float wreg[4][9], accw[4][9], sb[12];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int t = 0; t < 9; ++t) { wreg[c][t] = __ldg(a.w + (size_t)(ch * 4 + c) * 9 + t); accw[c][t] = 0.f; }
#pragma unroll
  for (int i = 0; i < 12; ++i) sb[i] = 0.f;
  const uint2* in = reinterpret_cast<const uint2*>(a.in);
  const uint2* y1 = reinterpret_cast<const uint2*>(a.bb.y);
  const uint2* y2 = reinterpret_cast<const uint2*>(a.bb.y2);
  uint2* dx = reinterpret_cast<uint2*>(a.dx);
  for (int t_i = blockIdx.x; t_i < a.ntiles; t_i += gridDim.x) {
    const int n = t_i / a.tiles_per_frame, oy0 = (t_i - n * a.tiles_per_frame) * a.R;
    __syncthreads();
    for (int e = tid; e < (a.R + 2) * W2; e += 256) {
      const int rr = e / W2, cc = e - rr * W2;
      const int iy = oy0 - 1 + rr, ix = cc - 1;
      gs[e] = ((unsigned)iy < (unsigned)a.H && (unsigned)ix < (unsigned)a.W)
                  ? __bfloat162float(a.dy[((size_t)n * a.H + iy) * a.W + ix]) : 0.f;
    }
    __syncthreads();
    const int r0 = band * 8;
    for (int xb = 0; xb < (kColBlocks ? a.W : 1); xb += cols) {
    const int x = xb + xl;
    float gw[3][3];                                  // dY window rows o-1, o, o+1 (tile rows r0+o .. r0+o+2), columns x-1 .. x+1
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) gw[i + 1][j] = gs[(r0 + i) * W2 + x + j];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // the global loads of four rows first (independent: 4 x 3 requests in flight per thread)
      uint2 av[4], y1v[4], y2v[4];
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int oy = oy0 + r0 + half * 4 + o;
        const size_t q = (((size_t)n * a.H + min(oy, a.H - 1)) * a.W + x) * CH + ch;
        av[o] = __ldg(in + q);
        y1v[o] = fuse ? __ldg(y1 + q) : make_uint2(0u, 0u);
        y2v[o] = two ? __ldg(y2 + q) : make_uint2(0u, 0u);
      }
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int ro = half * 4 + o;
#pragma unroll
        for (int j = 0; j < 3; ++j) { gw[0][j] = gw[1][j]; gw[1][j] = gw[2][j]; gw[2][j] = gs[(r0 + ro + 2) * W2 + x + j]; }
        const int oy = oy0 + r0 + ro;
        if (oy >= a.H) continue;
        const float af[4] = {bf_lo(av[o].x), bf_hi(av[o].x), bf_lo(av[o].y), bf_hi(av[o].y)};
        float o4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float sacc = 0.f;
#pragma unroll
          for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const float gv = gw[2 - kh][2 - kw];     // dY[q + (1-kh, 1-kw)]
              sacc = fmaf(gv, wreg[c][kh * 3 + kw], sacc);
              accw[c][kh * 3 + kw] = fmaf(af[c], gv, accw[c][kh * 3 + kw]);
            }
          o4[c] = sacc;
        }
        if (fuse) {
          const float yf[4] = {bf_lo(y1v[o].x), bf_hi(y1v[o].x), bf_lo(y1v[o].y), bf_hi(y1v[o].y)};
          const float zf[4] = {bf_lo(y2v[o].x), bf_hi(y2v[o].x), bf_lo(y2v[o].y), bf_hi(y2v[o].y)};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float gm = af[c] > 0.f ? bf16_round(o4[c]) : 0.f;
            o4[c] = gm;
            sb[c] += gm; sb[4 + c] = fmaf(gm, yf[c], sb[4 + c]); sb[8 + c] = fmaf(gm, zf[c], sb[8 + c]);
          }
        }
        const size_t q = (((size_t)n * a.H + oy) * a.W + x) * CH + ch;
        dx[q] = make_uint2(pack2(o4[0], o4[1]), pack2(o4[2], o4[3]));
      }
    }
    }
  }
  // reduce the per-thread accumulators over the lanes with the same channel chunk, then over the warps
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NACC; ++i) {
    float v = i < 36 ? accw[i / 9][i % 9] : sb[i - 36];
#pragma unroll
    for (int d = CH; d < 32; d <<= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if (lane < CH) red[(warp * CH + lane) * NACC + i] = v;
  }
  __syncthreads();
  for (int e = tid; e < CH * NACC; e += 256) {
    const int c4 = e / NACC, i = e - c4 * NACC;
    float s = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) s += red[(w8 * CH + c4) * NACC + i];
    red[e] = s;                                      // warp 0's slot now holds the CTA total (read and written by this thread only)
    if (i < 36) atomicAdd(a.dw + (size_t)(c4 * 4 + i / 9) * 9 + i % 9, s);      // weight index [0][ci][kh][kw]
  }
  if (fuse) {
    __syncthreads();
    if (tid < CI) {
      // S1 = rstd * (sum g*y - mean * S0): the subtraction in fp64 on the CTA totals
      const int c4 = tid >> 2, c = tid & 3;
      const double s0 = (double)red[c4 * NACC + 36 + c];
      const double s1 = (double)red[c4 * NACC + 40 + c], s2 = (double)red[c4 * NACC + 44 + c];
      double* acc = bn_bwd_acc_copy(a.bb) + tid;
      atomicAdd(acc, s0);
      atomicAdd(acc + a.bb.C, (double)a.bb.stat[a.bb.C + tid] * (s1 - (double)a.bb.stat[tid] * s0));
      if (two) atomicAdd(acc + 2 * a.bb.C, (double)a.bb.stat2[a.bb.C + tid] * (s2 - (double)a.bb.stat2[tid] * s0));
    }
    bn_bwd_fused_finish(a.bb, gridDim.x);
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

// ---------------- host: tile geometry ----------------
bool stem_supported(int Cin, int Co, int S, int k, int s, int p) {
  if (Cin != 1 || k != 5 || s != 2 || p != 2 || (Co != 32 && Co != 64) || (S & 1)) return false;
  const int Wo = S / 2;
  return Wo >= 4 && Wo <= 32;
}

static void stem_fill(StemArgs& a, int quads_per_tile) {
  a.Ho = a.S / 2; a.Wo = a.S / 2;
  a.fd_wo = FastDiv(a.Wo);
  a.qpr = (a.Wo + 3) / 4;
  a.R = quads_per_tile / a.qpr;
  if (a.R > a.Ho) a.R = a.Ho;
  if (a.R * a.Wo > 256) a.R = 256 / a.Wo;            // the weight-gradient kernel stages at most 256 pixels of dY
  a.tiles_per_frame = (a.Ho + a.R - 1) / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
  a.pitch = (8 * a.qpr + 4 + 3) & ~3;               // columns -2 .. 2*4*qpr + 1 of the input, rounded to 16 bytes
}

StatLayout launch_stem_fwd(StemArgs a, int Co, cudaStream_t st) {
  if (a.bn.acc && stem_tc_supported(Co, a.S)) {      // training-mode forward with fused statistics: the tensor-core kernel
    launch_stem_fwd_tc(a, Co, st);
    return StatLayout{0, 0, 0, 0};
  }
  stem_fill(a, 256 / (Co / 8));
  const int grid = min(a.ntiles, 148 * 3);
  count_launch();
  if (Co == 32) launch_pdl(stem_fwd_kernel<32>, grid, 256, 0, st, a);
  else launch_pdl(stem_fwd_kernel<64>, grid, 256, 0, st, a);
  return StatLayout{0, 0, 0, 0};     // statistics are finalised inside the kernel (a.bn)
}

void launch_stem_wgrad(StemArgs a, int Co, cudaStream_t st) {
  if (stem_wgrad_tc_supported(Co, a.S)) { launch_stem_wgrad_tc(a, st); return; }
  stem_fill(a, 64);
  const int ny = Co / 32;
  dim3 grid(min(a.ntiles, 148 / ny), ny);
  count_launch();
  launch_pdl(stem_wgrad_kernel, grid, kStemWgThreads, 0, st, a, Co);
}

bool tail_supported(int Ci, int Co, int H, int k, int s, int p) {
  if (Co != 1 || k != 3 || s != 1 || p != 1 || (Ci != 16 && Ci != 32)) return false;
  // forward: one band = 8 rows x W columns x Ci/8 threads must fit the 256 threads; backward: W x Ci/4 threads either fit
  // (whole bands per CTA) or are a multiple of 256 (column blocks, tail_bwd_kernel)
  const int wf = H * (Ci / 8), wb = H * (Ci / 4);
  return H % 8 == 0 && wf <= 256 && 256 % wf == 0 && (wb <= 256 ? 256 % wb == 0 : wb % 256 == 0);
}

static void tail_fill(TailArgs& a, int threads_per_pixel) {
  const int bands = max(1, 256 / (a.W * threads_per_pixel));
  a.R = 8 * bands;
  if (a.R > a.H) a.R = a.H;
  a.tiles_per_frame = (a.H + a.R - 1) / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
}

StatLayout launch_tail_fwd(TailArgs a, int Ci, cudaStream_t st) {
  if (a.bn.acc && tail_fwd_tc_supported(Ci, a.H, a.W)) { launch_tail_fwd_tc(a, Ci, st); return StatLayout{0, 0, 0, 0}; }
  tail_fill(a, Ci / 8);
  const int grid = min(a.ntiles, 148 * 2);
  const size_t smem = 2 * (size_t)(a.R + 2) * (a.W + 2) * Ci * 2;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tail_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(tail_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done = true;
  }
  count_launch();
  if (Ci == 16) launch_pdl(tail_fwd_kernel<16>, grid, 256, smem, st, a);
  else launch_pdl(tail_fwd_kernel<32>, grid, 256, smem, st, a);
  return StatLayout{0, 0, 0, 0};     // statistics are finalised inside the kernel (a.bn)
}

void launch_tail_bwd(TailArgs a, int Ci, cudaStream_t st) {
  if (tail_bwd_tc_supported(Ci, a.H, a.W)) { launch_tail_bwd_tc(a, Ci, st); return; }
  tail_fill(a, Ci / 4);
  const int grid = min(a.ntiles, 148 * 2);
  const size_t smem = sizeof(float) * ((size_t)(a.R + 2) * (a.W + 2) + (size_t)8 * (Ci / 4) * 48);
  count_launch();
  if (Ci == 16) launch_pdl(tail_bwd_kernel<16, false>, grid, 256, smem, st, a);
  else if (a.W * (Ci / 4) <= 256) launch_pdl(tail_bwd_kernel<32, false>, grid, 256, smem, st, a);
  else launch_pdl(tail_bwd_kernel<32, true>, grid, 256, smem, st, a);
}

void launch_heads_wgrad(const float* dheads, const float* pooled, float* g_mu, float* g_lv, int N, int z, int C,
                        cudaStream_t st) {
  dim3 grid((z + 7) / 8, g_lv ? 2 : 1, (C + 63) / 64);
  count_launch();
  launch_pdl(heads_wgrad2_kernel, grid, 256, 0, st, dheads, pooled, g_mu, g_lv, N, z, C);
}

}  // namespace mmvae
