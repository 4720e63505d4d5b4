// tail_tc.cu -- the one-channel tail convolution Conv2d(16w -> 1, k3 s1 p1, bias) (reference model.py:172) on the tensor
// cores: its backward pass (tail_bwd_tc) and its forward pass (tail_fwd_tc).
//
// The fp32 SIMT kernels of special.cu spend 55 thread instructions per element on 18 FMAs (profiles/
// r02_pointwise_loss_hbm.md: 58 us for 136 MB of traffic, issue bound).  Here the FMAs go to tcgen05 and the threads keep
// only what is per element anyway (ReLU mask, BatchNorm-backward sums, the stores).
//
// Backward, per tile of 128 consecutive pixels (R = 128 / W image rows), thread = pixel:
//   A  = im2col of the ONE-channel dY: A[q][t] = dY[q + (1-kh, 1-kw)], 9 taps padded to 16, built by the thread from a
//        cp.async-staged halo tile and written as plane[t / 8][pixel][8 taps] (bf16 is exact: dY is bf16)
//   dX[q][ci]  = A[q][:] . w[ci][:]            MMA 1: M = 128 pixels, N = CI, K = 16, twice: B = the bf16 head, then the bf16
//                                              tail of the fp32 weights, accumulated
//   dW[ci][t] += sum_q A[q][t] * a[q][ci]      MMA 2: the pixel axis is K (8 k-steps), both operands MN-major views: A is the
//                                              same shared-memory image, B = the activation tile plane[ci / 8][pixel][8 ch];
//                                              M = 128 with taps in rows 0..8 (planes 2..15 of the A descriptor run into the
//                                              staging ring behind it; those accumulator rows are never read); the accumulator
//                                              lives in TMEM for the CTA's whole life
//   epilogue (of the PREVIOUS tile, under this tile's MMAs): g = bf16(dX) * [a > 0], S0 / S1 / S2 sums per thread and channel
//   (transposed once at the end), 32-byte stores.  a, y, y2 of a tile arrive by cp.async in the plane layout (conflict-free
//   16-byte reads by thread = pixel, and a is the MMA-2 operand as it lies).
#include "bn_fused.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

namespace {

using namespace tc;

constexpr int kPlane = 128 * 16;                     // one operand plane: 128 pixels x 16 bytes
constexpr int kHaloBytes = 1024;                     // (R + 2) rows x (W + 16) bf16 of dY, interior at element 8 of a row

__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(v));
  lo = v - hi;
}
__device__ __forceinline__ void unpack8(const uint4& u, float* v) {
  v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
  v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
  v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
  v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
}

template <int CI>
__global__ void __launch_bounds__(128, CI == 16 ? 4 : 2) tail_bwd_tc_kernel(const TailArgs a) {
  constexpr int G = CI / 8;                          // 16-byte chunks (planes) per pixel
  constexpr int kTens = G * kPlane;                  // one staged tensor tile
  constexpr int kSlot = 3 * kTens + kHaloBytes;      // a | y | y2 | dY halo
  constexpr int kABuf = 2 * kPlane;                  // im2col operand: 2 planes (taps 0..7, 8..15)
  constexpr int kWPlane = 2 * CI * 16;               // weight plane: 2 CI rows (hi | lo) x 8 taps
  constexpr int kTmemCols = 3 * CI <= 64 ? 64 : 128;     // D1 double-buffered (2 x CI) + D2 (CI)
  extern __shared__ __align__(128) unsigned char ttc_smem[];
  __shared__ __align__(8) unsigned long long mma_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_red[4][3][CI];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  // [A buf 0 | A buf 1 | W planes | slot 0 | slot 1 | slot 2]
  const uint32_t a_base = (smem_u32(ttc_smem) + 127u) & ~127u;
  const uint32_t w_base = a_base + 2 * kABuf, s_base = w_base + 2 * kWPlane;
  unsigned char* a_ptr = ttc_smem + (a_base - smem_u32(ttc_smem));
  unsigned char* w_ptr = a_ptr + 2 * kABuf;
  unsigned char* s_ptr = w_ptr + 2 * kWPlane;
  static_assert(kABuf + 16 * kPlane <= 2 * kABuf + 2 * kWPlane + 3 * kSlot, "phantom planes leave the allocation");
  if (tid == 0) {
    mbar_init(smem_u32(&mma_done[0]), 1);
    mbar_init(smem_u32(&mma_done[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), (uint32_t)kTmemCols);
  // halo columns / pad of the dY tiles stay zero (the loader writes the interior only)
  for (int s = 0; s < 3; ++s)
    for (int e = tid; e < kHaloBytes / 4; e += 128) reinterpret_cast<uint32_t*>(s_ptr + s * kSlot + 3 * kTens)[e] = 0u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();
  const bool fuse = a.bb.acc != nullptr;
  const bool two = fuse && a.bb.y2 != nullptr;
  // weights w[ci][9] fp32 -> planes [t / 8][row][8 taps] with row ci = bf16 head, row CI + ci = tail
  for (int e = tid; e < 2 * CI * 2; e += 128) {
    const int row = e >> 1, k8 = e & 1, ci = row % CI;
    const bool tail = row >= CI;
    uint32_t pk[4];
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
      float v[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int t = k8 * 8 + 2 * p2 + j;
        float hi = 0.f, lo = 0.f;
        if (t < 9) split_bf16(__ldg(a.w + (size_t)ci * 9 + t), hi, lo);
        v[j] = tail ? lo : hi;
      }
      pk[p2] = pack_bf16x2(v[0], v[1]);
    }
    *reinterpret_cast<uint4*>(w_ptr + k8 * kWPlane + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  const int w_log2 = 31 - __clz(a.W);
  const int ry = tid >> w_log2, ox = tid & (a.W - 1);
  const int hp = a.W + 16;                           // halo row pitch in elements
  const uint32_t idesc1 = make_idesc_bf16(128, CI, 0, 0);
  const uint32_t idesc2 = make_idesc_bf16(128, CI, 1, 1);
  const uint32_t d2 = tmem + (uint32_t)(2 * CI);

  // stage tile `tile` into ring slot `slot`
  auto prefetch = [&](int tile, int slot) {
    const uint32_t sb = s_base + (uint32_t)(slot * kSlot);
    const size_t base = (size_t)tile * 128 * CI * 2;                     // tiles are contiguous runs of 128 pixels
#pragma unroll
    for (int i = 0; i < G; ++i) {
      const int e = tid + i * 128, p = e / G, c = e % G;
      const uint32_t d = sb + (uint32_t)(c * kPlane + p * 16);
      const size_t off = base + (size_t)e * 16;
      cp_async16(d, reinterpret_cast<const unsigned char*>(a.in) + off, 16u);
      if (fuse) cp_async16(d + kTens, reinterpret_cast<const unsigned char*>(a.bb.y) + off, 16u);
      if (two) cp_async16(d + 2 * kTens, reinterpret_cast<const unsigned char*>(a.bb.y2) + off, 16u);
    }
    // dY rows oy0 - 1 .. oy0 + R, W / 8 chunks each; rows outside the image are zero-filled
    const int cpr = a.W >> 3;
    if (tid < (a.R + 2) * cpr) {
      const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
      const int r = tid / cpr, c = tid - r * cpr, iy = oy0 - 1 + r;
      const bool ok = (unsigned)iy < (unsigned)a.H;
      const __nv_bfloat16* src = a.dy + ((size_t)n * a.H + (ok ? iy : 0)) * a.W + c * 8;
      cp_async16(sb + 3 * kTens + (uint32_t)(r * hp + 8 + c * 8) * 2u, src, ok ? 16u : 0u);
    }
  };

  constexpr bool kPerThread = CI == 16;              // BatchNorm-backward sums per thread and channel over all tiles
  constexpr int NG = CI / 16, NR = kPerThread ? CI : NG;
  float s0[NR], s1[NR], s2[NR];
#pragma unroll
  for (int i = 0; i < NR; ++i) { s0[i] = 0.f; s1[i] = 0.f; s2[i] = 0.f; }
  const int lane_col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);

  // epilogue of the tile in TMEM buffer b / ring slot `slot`
  auto epilogue = [&](int tile, int b, int slot, uint32_t parity) {
    mbar_wait(smem_u32(&mma_done[b]), parity);
    tc_fence_after();
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * CI);
    const unsigned char* sp = s_ptr + slot * kSlot + tid * 16;
    __nv_bfloat16* out = a.dx + ((size_t)tile * 128 + tid) * CI;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float d[16], yv[16], zv[16];
      uint32_t au[8];
      tmem_ld16(tl + (uint32_t)(g * 16), d);
      if (fuse) {
        *reinterpret_cast<uint4*>(au) = *reinterpret_cast<const uint4*>(sp + (2 * g) * kPlane);
        *reinterpret_cast<uint4*>(au + 4) = *reinterpret_cast<const uint4*>(sp + (2 * g + 1) * kPlane);
        unpack8(*reinterpret_cast<const uint4*>(sp + kTens + (2 * g) * kPlane), yv);
        unpack8(*reinterpret_cast<const uint4*>(sp + kTens + (2 * g + 1) * kPlane), yv + 8);
        if (two) {
          unpack8(*reinterpret_cast<const uint4*>(sp + 2 * kTens + (2 * g) * kPlane), zv);
          unpack8(*reinterpret_cast<const uint4*>(sp + 2 * kTens + (2 * g + 1) * kPlane), zv + 8);
        }
      }
      // g = bf16(dX) * [a > 0] on the packed pairs: the ReLU gate is the sign / zero test of a's bf16 bit pattern, the
      // values that enter the sums are the stored ones (a shift / a mask each)
      uint32_t pk[8];
      float gm[16];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        pk[e] = pack_bf16x2(d[2 * e], d[2 * e + 1]);
        if (fuse) {
          const uint32_t keep = ((int)(au[e] << 16) > 0 ? 0x0000ffffu : 0u) | ((int)(au[e] & 0xffff0000u) > 0 ? 0xffff0000u : 0u);
          pk[e] &= keep;
        }
        gm[2 * e] = __uint_as_float(pk[e] << 16); gm[2 * e + 1] = __uint_as_float(pk[e] & 0xffff0000u);
      }
      *reinterpret_cast<uint4*>(out + g * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(out + g * 16 + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      if (fuse) {
        if constexpr (kPerThread) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            s0[e] += gm[e]; s1[e] = fmaf(gm[e], yv[e], s1[e]);
            if (two) s2[e] = fmaf(gm[e], zv[e], s2[e]);
          }
        } else {
          float t1[16], t2[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) { t1[e] = gm[e] * yv[e]; t2[e] = two ? gm[e] * zv[e] : 0.f; }
          warp_colsum16(gm, lane);
          warp_colsum16(t1, lane);
          s0[g] += gm[0]; s1[g] += t1[0];
          if (two) { warp_colsum16(t2, lane); s2[g] += t2[0]; }
        }
      }
    }
    tc_fence_before();
  };

  int it = 0, prev_tile = -1;
  if ((int)blockIdx.x < a.ntiles) prefetch(blockIdx.x, 0);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++it) {
    const int buf = it & 1, slot = it % 3;
    cp_async_wait<0>();
    __syncthreads();                                 // tile `it` is staged; everyone is past the epilogue of tile it-2
    {                                                // its ring slot ((it+1) % 3) takes tile it+1
      const int next = tile + gridDim.x;
      if (next < a.ntiles) prefetch(next, (it + 1) % 3);
      cp_async_commit();
    }
    // ---- this pixel's im2col row of dY: 9 taps, A[t] = halo[ry + 2 - kh][ox + 9 - kw] ----
    {
      const unsigned short* hs = reinterpret_cast<const unsigned short*>(s_ptr + slot * kSlot + 3 * kTens);
      uint32_t t9[9];
#pragma unroll
      for (int kh = 0; kh < 3; ++kh)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) t9[kh * 3 + kw] = hs[(ry + 2 - kh) * hp + ox + 9 - kw];
      unsigned char* ap = a_ptr + buf * kABuf + tid * 16;
      *reinterpret_cast<uint4*>(ap) = make_uint4(t9[0] | (t9[1] << 16), t9[2] | (t9[3] << 16), t9[4] | (t9[5] << 16), t9[6] | (t9[7] << 16));
      *reinterpret_cast<uint4*>(ap + kPlane) = make_uint4(t9[8], 0u, 0u, 0u);
    }
    fence_proxy_async_smem();                        // operand writes (generic proxy, cp.async) -> the tensor core
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ab = a_base + (uint32_t)(buf * kABuf), sb = s_base + (uint32_t)(slot * kSlot);
        const uint64_t d_a = make_smem_desc(ab, kPlane, 128, SWZ_NONE);
        mma_bf16(tmem + (uint32_t)(buf * CI), d_a, make_smem_desc(w_base, kWPlane, 128, SWZ_NONE), idesc1, 0);
        mma_bf16(tmem + (uint32_t)(buf * CI), d_a, make_smem_desc(w_base + CI * 16, kWPlane, 128, SWZ_NONE), idesc1, 1);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)               // 16 pixels per MMA: two core matrices of 8 pixels, 128 B apart
          mma_bf16(d2, make_smem_desc(ab + ks * 256, 128, kPlane, SWZ_NONE), make_smem_desc(sb + ks * 256, 128, kPlane, SWZ_NONE),
                   idesc2, (it | ks) != 0);
        mma_commit(smem_u32(&mma_done[buf]));
      }
      __syncwarp();
    }
    // ---- epilogue of the PREVIOUS tile, under this tile's MMAs ----
    if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, (it + 2) % 3, (uint32_t)(((it - 1) >> 1) & 1));
    prev_tile = tile;
  }
  cp_async_wait<0>();
  if (prev_tile >= 0) epilogue(prev_tile, (it - 1) & 1, (it + 2) % 3, (uint32_t)(((it - 1) >> 1) & 1));
  // dW: accumulator rows = taps 0..8, columns = channels (all MMAs of this CTA are complete: the last commit was waited for)
  if (warp == 0 && it >= 1) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float v[16];
      tmem_ld16(d2 + (uint32_t)(g * 16), v);
      if (lane < 9) {
#pragma unroll
        for (int e = 0; e < 16; ++e) atomicAdd(a.dw + (size_t)(g * 16 + e) * 9 + lane, v[e]);
      }
    }
  }
  if (fuse) {
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if constexpr (kPerThread) {
        warp_colsum16(s0, lane);
        warp_colsum16(s1, lane);
        warp_colsum16(s2, lane);
        if ((lane & 1) == 0) { stat_red[warp][0][lane_col] = s0[0]; stat_red[warp][1][lane_col] = s1[0]; stat_red[warp][2][lane_col] = s2[0]; }
      } else if ((lane & 1) == 0) {
        stat_red[warp][0][g * 16 + lane_col] = s0[g]; stat_red[warp][1][g * 16 + lane_col] = s1[g]; stat_red[warp][2][g * 16 + lane_col] = s2[g];
      }
    }
    __syncthreads();
    if (tid < CI) {
      // S1 = rstd * (sum g*y - mean * S0): the subtraction in fp64 on the CTA totals
      double t[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) t[k] = (double)((stat_red[0][k][tid] + stat_red[1][k][tid]) + (stat_red[2][k][tid] + stat_red[3][k][tid]));
      double* acc = bn_bwd_acc_copy(a.bb) + tid;
      atomicAdd(acc, t[0]);
      atomicAdd(acc + a.bb.C, (double)a.bb.stat[a.bb.C + tid] * (t[1] - (double)a.bb.stat[tid] * t[0]));
      if (two) atomicAdd(acc + 2 * a.bb.C, (double)a.bb.stat2[a.bb.C + tid] * (t[2] - (double)a.bb.stat2[tid] * t[0]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, (uint32_t)kTmemCols);
  }
  if (fuse) bn_bwd_fused_finish(a.bb, gridDim.x);
}


// ------------------------------------------------------------------------------------------------
// forward: y[q] = bias + sum_{t, ci} a[q + (kh-1, kw-1)][ci] * w[ci][t]      (GEMM N = 1: padded to the minimum N = 8)
// ------------------------------------------------------------------------------------------------
// The slab formulation of slab_tc.cu: the frame is addressed as ONE linear run of padded positions h = (row + 1) * (W + 2) +
// (col + 1) with a zero halo, staged as plane[ci / 8][position][8 channels] (uniform 16-byte pitch along the pixel axis), so a
// convolution tap is a shifted start address of the A descriptor: 9 tcgen05.mma (M = 128 positions, N = 8, K = 16 channels)
// per 128 positions and 16 input channels walk the taps over the same resident copy.  B[t] holds the weights of tap t as rows
// 0 / 1 = bf16 head / tail of w[:][t] (rows 2..7 zero): accumulator column 0 + column 1 is the fp32-weight result.  Rows of the
// accumulator that fall on halo columns are computed and dropped.  A CTA works on chunks of kTiles x 128 output positions of
// one frame (staged with one row + one position of halo either side, a ring of three cp.async stages), issues the chunk's MMAs and
// runs the epilogue of the previous chunk under them: bias, bf16 store, BatchNorm statistics of the values as stored.
constexpr int kFwdTiles = 3;

struct TailFwdGeom {
  int Wp;                      // W + 2
  int out_per_frame;           // H * Wp output positions (rows 1..H of the padded frame, all columns)
  int tiles_per_frame;         // ceil(out_per_frame / 128)
  int chunks_per_frame;        // ceil(tiles_per_frame / kFwdTiles)
  int count;                   // staged positions of a chunk: kFwdTiles * 128 + 2 * Wp + 2
  int plane;                   // bytes of one staged plane (count * 16 rounded up to 128)
  int nchunks;
  unsigned int wp_mul;         // ceil(2^32 / Wp): h / Wp == umulhi(h, wp_mul) for h < 2^16
};

template <int CI>
__global__ void __launch_bounds__(128, 4) tail_fwd_tc_kernel(const TailArgs a, const TailFwdGeom g) {
  constexpr int G = CI / 8, KS = CI / 16;
  constexpr int kWTap = G * 128;                     // weight bytes of one tap: G planes x 8 rows x 16 B
  extern __shared__ __align__(128) unsigned char ttc_smem[];
  __shared__ __align__(8) unsigned long long mma_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sh[2][4];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  const uint32_t st_base = (smem_u32(ttc_smem) + 127u) & ~127u;          // [stage 0 | stage 1 | stage 2 | weights]
  const uint32_t stage_bytes = (uint32_t)(G * g.plane);
  const uint32_t w_base = st_base + 3u * stage_bytes;
  unsigned char* w_ptr = ttc_smem + (w_base - smem_u32(ttc_smem));
  if (tid == 0) {
    mbar_init(smem_u32(&mma_done[0]), 1);
    mbar_init(smem_u32(&mma_done[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 64u);               // 2 buffers x kFwdTiles x 8 columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();
  // weights: tap t, plane c: row 0 = head, row 1 = tail of w[8c .. 8c + 7][t], rows 2..7 zero
  for (int e = tid; e < 9 * G * 8; e += 128) {
    const int t = e / (G * 8), c = (e / 8) % G, row = e & 7;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (row < 2) {
#pragma unroll
      for (int p2 = 0; p2 < 4; ++p2) {
        float h0, l0, h1, l1;
        split_bf16(__ldg(a.w + (size_t)(c * 8 + 2 * p2) * 9 + t), h0, l0);
        split_bf16(__ldg(a.w + (size_t)(c * 8 + 2 * p2 + 1) * 9 + t), h1, l1);
        pk[p2] = row == 0 ? pack_bf16x2(h0, h1) : pack_bf16x2(l0, l1);
      }
    }
    *reinterpret_cast<uint4*>(w_ptr + t * kWTap + c * 128 + row * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  const float bias = a.bias ? __ldg(a.bias) : 0.f;
  const uint32_t idesc = make_idesc_bf16(128, 8, 0, 0);
  const bool stats = a.bn.acc != nullptr;
  float run_s = 0.f, run_q = 0.f;

  // stage chunk `chunk`: padded positions [hfirst, hfirst + count) of its frame, hfirst = k * kFwdTiles * 128 - 1 (one padded
  // row + one position before the chunk's first output).  Walked by padded ROW: the interior of a row is W * G consecutive
  // 16-byte chunks of the NHWC tensor (thread -> (column, plane) by shifts only; the first version derived (row, column) of
  // every staged position with a multiply-high division and spent 5.4 M of its 7.3 M instructions there); rows 0 and H + 1
  // are zero-filled, the two halo columns of every row are zeroed by plain stores.
  auto prefetch = [&](int chunk, int sidx) {
    const int n = chunk / g.chunks_per_frame, k = chunk - n * g.chunks_per_frame;
    const int hfirst = k * (kFwdTiles * 128) - 1;
    const uint32_t sb = st_base + (uint32_t)sidx * stage_bytes;
    const unsigned char* frame = reinterpret_cast<const unsigned char*>(a.in + (size_t)n * a.H * a.W * CI);
    const int r0 = (int)__umulhi((unsigned)max(hfirst, 0), g.wp_mul);
    const int r1 = min((int)__umulhi((unsigned)(hfirst + g.count - 1), g.wp_mul), a.H + 1);
    const int row_chunks = a.W * G;
    for (int r = r0; r <= r1; ++r) {
      const int jrow = r * g.Wp - hfirst;                                // stage position of padded column 0 of this row
      const bool ok_row = r >= 1 && r <= a.H;
      const unsigned char* src = frame + (size_t)(ok_row ? r - 1 : 0) * row_chunks * 16;
      for (int idx = tid; idx < row_chunks; idx += 128) {
        const int x = idx / G, c = idx % G, j = jrow + 1 + x;
        if (j >= 0 && j < g.count) cp_async16(sb + (uint32_t)(c * g.plane + j * 16), src + (size_t)idx * 16, ok_row ? 16u : 0u);
      }
    }
    // halo columns 0 and Wp - 1 of the staged rows
    const int nh = (r1 - r0 + 1) * 2 * G;
    if (tid < nh) {
      const int c = tid % G, side = (tid / G) & 1, r = r0 + tid / (2 * G);
      const int j = r * g.Wp - hfirst + (side ? g.Wp - 1 : 0);
      if (j >= 0 && j < g.count) *reinterpret_cast<uint4*>(ttc_smem + (sb - smem_u32(ttc_smem)) + c * g.plane + j * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  };

  auto epilogue = [&](int chunk, int b, uint32_t parity) {
    mbar_wait(smem_u32(&mma_done[b]), parity);
    tc_fence_after();
    const int n = chunk / g.chunks_per_frame, k = chunk - n * g.chunks_per_frame;
    const int ntl = min(kFwdTiles, g.tiles_per_frame - k * kFwdTiles);
    __nv_bfloat16* yn = a.y + (size_t)n * a.H * a.W;
#pragma unroll
    for (int m = 0; m < kFwdTiles; ++m) {
      if (m < ntl) {
        float v[8];
        tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)((b * kFwdTiles + m) * 8), v);
        const int o = (k * kFwdTiles + m) * 128 + tid;                   // output position; padded index h = Wp + o
        const int r = (int)__umulhi((unsigned)o, g.wp_mul), col = o - r * g.Wp;      // image row r, padded column col
        if (r < a.H && col >= 1 && col <= a.W) {
          const __nv_bfloat16 ob = __float2bfloat16_rn((v[0] + v[1]) + bias);
          yn[(size_t)r * a.W + col - 1] = ob;
          const float f = __bfloat162float(ob);
          run_s += f; run_q = fmaf(f, f, run_q);
        }
      }
    }
    tc_fence_before();
  };

  int it = 0, prev_chunk = -1;
  if ((int)blockIdx.x < g.nchunks) prefetch(blockIdx.x, 0);
  cp_async_commit();
  for (int chunk = blockIdx.x; chunk < g.nchunks; chunk += gridDim.x, ++it) {
    const int buf = it & 1, stage = it % 3;
    cp_async_wait<0>();
    fence_proxy_async_smem();
    __syncthreads();                                 // chunk `it` is staged; everyone is past the epilogue of chunk it-2, which
    {                                                // waited for the MMAs that read stage (it + 1) % 3: it takes chunk it+1
      const int next = chunk + gridDim.x;
      if (next < g.nchunks) prefetch(next, (it + 1) % 3);
      cp_async_commit();
    }
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const int k = chunk % g.chunks_per_frame;
        const int ntl = min(kFwdTiles, g.tiles_per_frame - k * kFwdTiles);
        const uint32_t sb = st_base + (uint32_t)stage * stage_bytes;
        for (int m = 0; m < ntl; ++m) {
          const uint32_t dtm = tmem + (uint32_t)((buf * kFwdTiles + m) * 8);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
            const uint32_t shift = (uint32_t)((m * 128 + (t / 3) * g.Wp + (t % 3)) * 16);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              mma_bf16(dtm, make_smem_desc(sb + shift + (uint32_t)(2 * ks * g.plane), (uint32_t)g.plane, 128, SWZ_NONE),
                       make_smem_desc(w_base + (uint32_t)(t * kWTap + 2 * ks * 128), 128, 128, SWZ_NONE), idesc, (t | ks) != 0);
          }
        }
        mma_commit(smem_u32(&mma_done[buf]));
      }
      __syncwarp();
    }
    if (prev_chunk >= 0) epilogue(prev_chunk, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
    prev_chunk = chunk;
  }
  cp_async_wait<0>();
  if (prev_chunk >= 0) epilogue(prev_chunk, (it - 1) & 1, (uint32_t)(((it - 1) >> 1) & 1));
  if (stats) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { run_s += __shfl_xor_sync(0xffffffffu, run_s, d); run_q += __shfl_xor_sync(0xffffffffu, run_q, d); }
    if (lane == 0) { sh[0][warp] = run_s; sh[1][warp] = run_q; }
    __syncthreads();
    if (tid < 2) atomicAdd(bn_acc_copy(a.bn) + tid, (double)((sh[tid][0] + sh[tid][1]) + (sh[tid][2] + sh[tid][3])));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 64u);
  }
  if (stats) bn_fused_finish(a.bn, gridDim.x);
}

}  // namespace

// W a power of two with 8 <= W <= 128 and whole tiles of 128 pixels: R = 128 / W rows, H % R == 0
bool tail_bwd_tc_supported(int Ci, int H, int W) {
  static const bool off = getenv("MMVAE_NO_TAIL_TC") != nullptr;
  if (off || (Ci != 16 && Ci != 32)) return false;
  if (W < 8 || W > 128 || (W & (W - 1))) return false;
  const int R = 128 / W;
  return H % R == 0 && (R + 2) * (W + 16) * 2 <= kHaloBytes && (R + 2) * (W / 8) <= 128;
}

void launch_tail_bwd_tc(TailArgs a, int Ci, cudaStream_t st) {
  a.R = 128 / a.W;
  a.tiles_per_frame = a.H / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
  const int G = Ci / 8;
  const size_t smem = 2 * 2 * kPlane + 2 * (2 * Ci * 16) + 3 * (size_t)(3 * G * kPlane + kHaloBytes) + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tail_bwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(tail_bwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    attr_done = true;
  }
  static const int per_sm_env = getenv("MMVAE_TAIL_CTAS") ? atoi(getenv("MMVAE_TAIL_CTAS")) : 0;
  const int per_sm = per_sm_env ? per_sm_env : (Ci == 16 ? 4 : 2);
  const int grid = a.ntiles < 148 * per_sm ? a.ntiles : 148 * per_sm;
  count_launch();
  if (Ci == 16) launch_pdl(tail_bwd_tc_kernel<16>, grid, 128, smem, st, a);
  else launch_pdl(tail_bwd_tc_kernel<32>, grid, 128, smem, st, a);
}

}  // namespace mmvae

namespace mmvae {

bool tail_fwd_tc_supported(int Ci, int H, int W) {
  static const bool off = getenv("MMVAE_NO_TAIL_TC") != nullptr || getenv("MMVAE_NO_TAIL_FWD_TC") != nullptr;
  if (off || (Ci != 16 && Ci != 32)) return false;
  // positions of a frame below 2^16 (the multiply-high division), two stages within shared memory
  const size_t plane = (size_t)(kFwdTiles * 128 + 2 * (W + 2) + 2) * 16 + 127;
  return (H + 2) * (W + 2) + kFwdTiles * 128 < 65536 && W >= 4 && 3 * (Ci / 8) * plane + 9 * (Ci / 8) * 128 + 128 <= (Ci == 16 ? 100 : 200) * 1024;
}

void launch_tail_fwd_tc(TailArgs a, int Ci, cudaStream_t st) {
  TailFwdGeom g;
  g.Wp = a.W + 2;
  g.out_per_frame = a.H * g.Wp;
  g.tiles_per_frame = (g.out_per_frame + 127) / 128;
  g.chunks_per_frame = (g.tiles_per_frame + kFwdTiles - 1) / kFwdTiles;
  g.count = kFwdTiles * 128 + 2 * g.Wp + 2;
  g.plane = (g.count * 16 + 127) & ~127;
  g.nchunks = a.N * g.chunks_per_frame;
  g.wp_mul = (unsigned int)((0x100000000ull + g.Wp - 1) / g.Wp);
  const int G = Ci / 8;
  const size_t smem = 3 * (size_t)G * g.plane + 9 * G * 128 + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(tail_fwd_tc_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(tail_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  static const int per_sm_env = getenv("MMVAE_TAIL_FWD_CTAS") ? atoi(getenv("MMVAE_TAIL_FWD_CTAS")) : 0;
  const int per_sm = per_sm_env ? per_sm_env : (Ci == 16 ? 4 : 2);
  const int grid = g.nchunks < 148 * per_sm ? g.nchunks : 148 * per_sm;
  count_launch();
  if (Ci == 16) launch_pdl(tail_fwd_tc_kernel<16>, grid, 128, smem, st, a, g);
  else launch_pdl(tail_fwd_tc_kernel<32>, grid, 128, smem, st, a, g);
}

}  // namespace mmvae
