// stem_tc.cu -- the one-channel stem convolution Conv2d(1 -> 32w, k5 s2 p2) (reference model.py:94) on the tensor cores.
//
// As a GEMM the stem is M = N*Ho*Wo pixels, N = CO channels, K = 25 taps: far too thin for TMA im2col (one input channel)
// and, as the fp32 SIMT kernel of special.cu showed (31 us for 21 MB of traffic, 55 % of its instructions FFMA, issue
// bound), too many FMAs per byte for the CUDA cores to keep up with HBM.  Here the CTA BUILDS the im2col operand in shared
// memory itself: thread = output pixel, 25 taps read from a cp.async-staged fp32 halo tile of x, written as the canonical
// un-swizzled K-major layout plane[k / 8][pixel][8 taps] (the layout slab_tc.cu uses: LBO = plane stride, SBO = 128).
// fp32 fidelity on bf16 tensor cores: every operand is split into a bf16 head and a bf16 tail (v = hi + lo, |lo| <= 2^-9 |v|)
// and the product is evaluated as A_hi*W_hi + A_lo*W_hi + A_hi*W_lo (the lo*lo term is 2^-18): 6 tcgen05.mma (M = 128,
// N = CO, K = 16) per 128-pixel tile, fp32 accumulation in TMEM -- the result matches the fp32 SIMT kernel to ~1e-5.
// Epilogue as in gconv_tc.cu: tcgen05.ld -> bf16 NHWC stores, BatchNorm statistics of the values as stored (bn_fused.cuh).
// Pipeline per CTA (128 threads, persistent over tiles): x tile t+1 loads (cp.async) and the operand of tile t+1 is built and
// its MMAs issued BEFORE the epilogue of tile t runs, so the tensor pipe works under the stores of the previous tile.
#include "bn_fused.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

namespace {

using namespace tc;

constexpr int kStemTcXs = 1408;                    // floats of one x halo tile: (2R+3) rows x pitch (11 x 68 or 19 x 36)

__device__ __forceinline__ void cp_async8_zfill(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}

// x halo tile of output rows [oy0, oy0 + R): input rows 2*oy0 - 2 .. 2*oy0 + 2R, as 8-byte pairs; rows outside the image
// are zero-filled, the halo columns of the tile are zeroed once by the kernel and never written here
__device__ __forceinline__ void stem_tc_prefetch(const StemArgs& a, int tile, float* xs) {
  const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
  const int rows = 2 * a.R + 3, halfp = a.S >> 1;          // halfp is a power of two <= 32 (launcher)
  const float* xn = a.x + (size_t)n * a.S * a.S;
  const int j = threadIdx.x & (halfp - 1), rstep = 128 / halfp;
  const uint32_t d0 = smem_u32(xs + 2 + 2 * j);
  for (int rr = threadIdx.x / halfp; rr < rows; rr += rstep) {
    const int iy = 2 * oy0 - 2 + rr;
    const bool ok = (unsigned)iy < (unsigned)a.S;
    cp_async8_zfill(d0 + (uint32_t)(rr * a.pitch) * 4u, ok ? (const void*)(xn + (size_t)iy * a.S + 2 * j) : (const void*)xn, ok ? 8u : 0u);
  }
}

__device__ __forceinline__ void split_bf16(float v, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(v));
  lo = v - hi;
}

template <int CO>
__global__ void __launch_bounds__(128, CO == 32 ? 4 : 3) stem_fwd_tc_kernel(const StemArgs a) {
  extern __shared__ __align__(128) unsigned char stc_smem[];
  // [2 buffers][hi | lo][4 planes][128 pixels][16 B]  then  W [hi | lo][4 planes][CO][16 B]
  constexpr int kPlane = 128 * 16, kOper = 4 * kPlane, kBuf = 2 * kOper;
  constexpr int kWPlane = CO * 16, kWOper = 4 * kWPlane;
  __shared__ __align__(16) float xs2[2][kStemTcXs];
  __shared__ __align__(8) unsigned long long mma_done[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_red[4][2][CO];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  const uint32_t a_base = (smem_u32(stc_smem) + 127u) & ~127u, w_base = a_base + 2u * kBuf;
  unsigned char* a_ptr = stc_smem + (a_base - smem_u32(stc_smem));
  if (tid == 0) {
    mbar_init(smem_u32(&mma_done[0]), 1);
    mbar_init(smem_u32(&mma_done[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 2 * CO <= 64 ? 64u : 128u);
  for (int e = tid; e < 2 * kStemTcXs; e += 128) (&xs2[0][0])[e] = 0.f;      // halo columns stay zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();                                       // the weights (optimizer) and x (input kernel) are final from here on
  pdl_trigger();
  // weights: W[co][25] fp32 -> hi / lo bf16 planes, K-major: plane k8 holds taps 8*k8 .. 8*k8 + 7 of every channel
  for (int e = tid; e < CO * 4; e += 128) {
    const int co = e >> 2, k8 = e & 3;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p2 = 0; p2 < 4; ++p2) {
      float h0 = 0.f, l0 = 0.f, h1 = 0.f, l1 = 0.f;
      const int k0 = k8 * 8 + 2 * p2;
      if (k0 < 25) split_bf16(__ldg(a.w + (size_t)co * 25 + k0), h0, l0);
      if (k0 + 1 < 25) split_bf16(__ldg(a.w + (size_t)co * 25 + k0 + 1), h1, l1);
      hi[p2] = pack_bf16x2(h0, h1); lo[p2] = pack_bf16x2(l0, l1);
    }
    unsigned char* wp = a_ptr + 2 * kBuf + k8 * kWPlane + co * 16;
    *reinterpret_cast<uint4*>(wp) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(wp + kWOper) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  // this thread's pixel inside a tile: row ry of the tile's R output rows, column ox
  const int wo_log2 = 31 - __clz(a.Wo);
  const int ry = tid >> wo_log2, ox = tid & (a.Wo - 1);
  const uint32_t idesc = make_idesc_bf16(128, CO, 0, 0);
  constexpr int NG = CO / 16;
  const int lane_col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
  // BatchNorm sums of this thread's pixels: CO = 32 keeps them per thread and column over all tiles (64 registers) and
  // transposes ONCE at the end; CO = 64 transposes per tile (the per-tile butterflies were ~250 of the ~890 instructions a
  // thread spent on a tile)
  constexpr bool kPerThread = CO == 32;
  constexpr int NR = kPerThread ? CO : NG;
  float run_s[NR], run_q[NR];
#pragma unroll
  for (int g = 0; g < NR; ++g) { run_s[g] = 0.f; run_q[g] = 0.f; }
  const bool stats = a.bn.acc != nullptr;

  // epilogue of the tile whose accumulator sits in TMEM buffer b (its MMAs were committed to mma_done[b])
  auto epilogue = [&](int tile, int b, uint32_t parity) {
    mbar_wait(smem_u32(&mma_done[b]), parity);
    tc_fence_after();
    const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
    const size_t m = ((size_t)n * a.Ho + oy0 + ry) * a.Wo + ox;
    const uint32_t tl = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(b * CO);
    __nv_bfloat16* out = a.y + m * CO;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      float v[16];
      tmem_ld16(tl + (uint32_t)(g * 16), v);
      uint4 p0, p1;
      p0.x = pack_bf16x2(v[0], v[1]); p0.y = pack_bf16x2(v[2], v[3]); p0.z = pack_bf16x2(v[4], v[5]); p0.w = pack_bf16x2(v[6], v[7]);
      p1.x = pack_bf16x2(v[8], v[9]); p1.y = pack_bf16x2(v[10], v[11]); p1.z = pack_bf16x2(v[12], v[13]); p1.w = pack_bf16x2(v[14], v[15]);
      *reinterpret_cast<uint4*>(out + g * 16) = p0;
      *reinterpret_cast<uint4*>(out + g * 16 + 8) = p1;
      if (stats) {
        // the values as stored: unpack the bf16 pairs just written (a shift / a mask each)
        const uint32_t pk[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) { v[2 * e] = __uint_as_float(pk[e] << 16); v[2 * e + 1] = __uint_as_float(pk[e] & 0xffff0000u); }
        if constexpr (kPerThread) {
#pragma unroll
          for (int e = 0; e < 16; ++e) { run_s[g * 16 + e] += v[e]; run_q[g * 16 + e] = fmaf(v[e], v[e], run_q[g * 16 + e]); }
        } else {
          float sq[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) sq[e] = v[e] * v[e];
          warp_colsum16(v, lane);
          warp_colsum16(sq, lane);
          run_s[g] += v[0]; run_q[g] += sq[0];
        }
      }
    }
    tc_fence_before();
  };

  int buf = 0, it = 0, prev_tile = -1;
  if ((int)blockIdx.x < a.ntiles) stem_tc_prefetch(a, blockIdx.x, xs2[0]);
  cp_async_commit();
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, buf ^= 1, ++it) {
    const int next = tile + gridDim.x;
    if (next < a.ntiles) stem_tc_prefetch(a, next, xs2[buf ^ 1]);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();                                 // x tile of this iteration is complete for every thread
    // ---- build this pixel's im2col row (25 taps, hi / lo) into operand buffer `buf` ----
    {
      const float* xr = xs2[buf] + (2 * ry) * a.pitch + 2 * ox;           // input column 2*ox - 2 lives at tile column 2*ox
      float hi[32], lo[32];
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const float2 x01 = *reinterpret_cast<const float2*>(xr + kh * a.pitch);
        const float2 x23 = *reinterpret_cast<const float2*>(xr + kh * a.pitch + 2);
        const float x4 = xr[kh * a.pitch + 4];
        split_bf16(x01.x, hi[kh * 5 + 0], lo[kh * 5 + 0]); split_bf16(x01.y, hi[kh * 5 + 1], lo[kh * 5 + 1]);
        split_bf16(x23.x, hi[kh * 5 + 2], lo[kh * 5 + 2]); split_bf16(x23.y, hi[kh * 5 + 3], lo[kh * 5 + 3]);
        split_bf16(x4, hi[kh * 5 + 4], lo[kh * 5 + 4]);
      }
#pragma unroll
      for (int k = 25; k < 32; ++k) { hi[k] = 0.f; lo[k] = 0.f; }
      unsigned char* ap = a_ptr + buf * kBuf + tid * 16;
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) {
        *reinterpret_cast<uint4*>(ap + k8 * kPlane) =
            make_uint4(pack_bf16x2(hi[k8 * 8], hi[k8 * 8 + 1]), pack_bf16x2(hi[k8 * 8 + 2], hi[k8 * 8 + 3]),
                       pack_bf16x2(hi[k8 * 8 + 4], hi[k8 * 8 + 5]), pack_bf16x2(hi[k8 * 8 + 6], hi[k8 * 8 + 7]));
        *reinterpret_cast<uint4*>(ap + kOper + k8 * kPlane) =
            make_uint4(pack_bf16x2(lo[k8 * 8], lo[k8 * 8 + 1]), pack_bf16x2(lo[k8 * 8 + 2], lo[k8 * 8 + 3]),
                       pack_bf16x2(lo[k8 * 8 + 4], lo[k8 * 8 + 5]), pack_bf16x2(lo[k8 * 8 + 6], lo[k8 * 8 + 7]));
      }
    }
    fence_proxy_async_smem();                        // generic-proxy operand writes -> visible to the tensor core
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ab = a_base + (uint32_t)(buf * kBuf);
        const uint32_t dtm = tmem + (uint32_t)(buf * CO);
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {             // two k-steps of 16 taps = two planes each
          const uint64_t a_hi = make_smem_desc(ab + ks * 2 * kPlane, kPlane, 128, SWZ_NONE);
          const uint64_t a_lo = make_smem_desc(ab + kOper + ks * 2 * kPlane, kPlane, 128, SWZ_NONE);
          const uint64_t w_hi = make_smem_desc(w_base + ks * 2 * kWPlane, kWPlane, 128, SWZ_NONE);
          const uint64_t w_lo = make_smem_desc(w_base + kWOper + ks * 2 * kWPlane, kWPlane, 128, SWZ_NONE);
          mma_bf16(dtm, a_hi, w_hi, idesc, ks != 0);
          mma_bf16(dtm, a_lo, w_hi, idesc, 1);
          mma_bf16(dtm, a_hi, w_lo, idesc, 1);
        }
        mma_commit(smem_u32(&mma_done[buf]));
      }
      __syncwarp();
    }
    // ---- epilogue of the PREVIOUS tile, under this tile's MMAs ----
    if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
    prev_tile = tile;
  }
  cp_async_wait<0>();
  if (prev_tile >= 0) epilogue(prev_tile, buf ^ 1, (uint32_t)(((it - 1) >> 1) & 1));
  if (stats) {
    if constexpr (kPerThread) {
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        float vs[16], vq[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { vs[e] = run_s[g * 16 + e]; vq[e] = run_q[g * 16 + e]; }
        warp_colsum16(vs, lane);
        warp_colsum16(vq, lane);
        if ((lane & 1) == 0) { stat_red[warp][0][g * 16 + lane_col] = vs[0]; stat_red[warp][1][g * 16 + lane_col] = vq[0]; }
      }
    } else if ((lane & 1) == 0) {
#pragma unroll
      for (int g = 0; g < NG; ++g) { stat_red[warp][0][g * 16 + lane_col] = run_s[g]; stat_red[warp][1][g * 16 + lane_col] = run_q[g]; }
    }
    __syncthreads();
    for (int e = tid; e < 2 * CO; e += 128) {
      const int which = e / CO, c = e - which * CO;
      const float t = (stat_red[0][which][c] + stat_red[1][which][c]) + (stat_red[2][which][c] + stat_red[3][which][c]);
      atomicAdd(bn_acc_copy(a.bn) + which * a.bn.C + c, (double)t);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 2 * CO <= 64 ? 64u : 128u);
  }
  if (stats) bn_fused_finish(a.bn, gridDim.x);
}


// ------------------------------------------------------------------------------------------------
// weight gradient: dW[co][tap] = sum_pixels X[pixel][tap] * dY[pixel][co]          (no data gradient: x needs none)
// ------------------------------------------------------------------------------------------------
// The pixel axis is the GEMM K axis (as in slab_wgrad): D[M = tap][N = co] += A^T B with both operands MN-major,
//   A = the im2col operand of the forward kernel, the SAME shared-memory image plane[tap / 8][pixel][8 taps] read as an
//       MN-major matrix (core matrix = 8 pixels x 8 taps; LBO = 128 B to the next 8 pixels, SBO = plane stride),
//   B = dY[pixel][co] staged by cp.async as plane[co / 8][pixel][8 channels].
// M is 128 for cta_group::1: the 25 taps occupy rows 0..24 of the accumulator, planes 4..15 of the A descriptor run on
// into whatever shared memory follows (inside the allocation; those accumulator rows are never read).  x is split into a
// bf16 head and tail (two MMAs per 16 pixels), dY is bf16 already: the result matches the fp32 SIMT kernel to ~1e-6.
// One TMEM accumulator lives for the CTA's whole life; 16 MMAs per 128-pixel tile; the epilogue is 25 x 32 atomics per CTA.
// kFuse: the B operand is not a stored dY but the stem BatchNorm's backward evaluated in the loader: a thread owns one
// 16-byte channel chunk (8 channels: tid & 3) of 4 pixels of the tile, keeps dY = A * g + B * y + D as three coefficients
// per channel in registers, holds the NEXT tile's four raw vectors (g1, g2, mask, y) in registers while the current tile's
// MMAs run, and writes the bf16 dY chunk straight into the operand plane (st.shared): dY never touches HBM, and the
// BatchNorm backward of the 16.8 MB stem output is one reduction pass instead of reduce + grid barrier + apply.
template <bool kFuse>
__global__ void __launch_bounds__(128, 3) stem_wgrad_tc_kernel(const StemArgs a) {
  constexpr int CO = 32;
  extern __shared__ __align__(128) unsigned char stc_smem[];
  constexpr int kPlane = 128 * 16, kOper = 4 * kPlane, kABuf = 2 * kOper;       // A: [hi | lo][4 planes][128 px][16 B]
  constexpr int kBBuf = (CO / 8) * kPlane;                                      // B: [CO/8 planes][128 px][16 B]
  constexpr int kSlots = 3;                                                     // cp.async targets (x tile, B) are 3 deep
  __shared__ __align__(16) float xs3[kSlots][kStemTcXs];
  __shared__ __align__(8) unsigned long long mma_done[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  // A buffers first, the B slots behind them: the phantom planes 4..15 of the last A operand (24 KB + 32 KB) end exactly
  // where the B slots end (32 KB + 3 x 8 KB)
  const uint32_t a_base = (smem_u32(stc_smem) + 127u) & ~127u;
  const uint32_t b_base = a_base + 2 * kABuf;
  unsigned char* a_ptr = stc_smem + (a_base - smem_u32(stc_smem));
  static_assert(kABuf + kOper + 16 * kPlane <= 2 * kABuf + kSlots * kBBuf, "phantom planes leave the allocation");
  if (tid == 0) {
    mbar_init(smem_u32(&mma_done[0]), 1);
    mbar_init(smem_u32(&mma_done[1]), 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(&tmem_base_s), 32u);
  for (int e = tid; e < kSlots * kStemTcXs; e += 128) (&xs3[0][0])[e] = 0.f;    // halo columns stay zero
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const int wo_log2 = 31 - __clz(a.Wo);
  const int ry = tid >> wo_log2, ox = tid & (a.Wo - 1);
  const uint32_t idesc = make_idesc_bf16(128, CO, 1, 1);                        // both operands MN-major

  // stage tile `tile` into slot `slot`: the x halo tile and the 128 x CO dY tile (chunk c of pixel p -> plane c, row p)
  auto prefetch = [&](int tile, int slot) {
    stem_tc_prefetch(a, tile, xs3[slot]);
    if constexpr (!kFuse) {
      const int n = tile / a.tiles_per_frame, oy0 = (tile - n * a.tiles_per_frame) * a.R;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.dy + (((size_t)n * a.Ho + oy0) * a.Wo) * CO);
      const uint32_t dst = b_base + (uint32_t)(slot * kBBuf);
#pragma unroll
      for (int i = 0; i < CO / 8; ++i) {
        const int e = tid + i * 128, p = e >> 2, c = e & 3;                     // CO / 8 == 4 chunks per pixel
        cp_async16(dst + (uint32_t)(c * kPlane + p * 16), src + (size_t)p * (CO * 2) + c * 16, 16u);
      }
    }
  };
  // fused BatchNorm backward: chunk e = tid + 128 i of the tile (pixel e >> 2, channel chunk e & 3 == tid & 3)
  uint4 qg[4], qg2[4], qm[4], qy[4];
  float cA[8], cB[8], cD[8];
  const size_t tile_chunks = (size_t)128 * (CO / 8);
  auto load_fwd = [&](int tile) {                    // what the forward wrote: may be fetched before the dependency wait
    const uint4* pm = reinterpret_cast<const uint4*>(a.mask) + (size_t)tile * tile_chunks + tid;
    const uint4* py = reinterpret_cast<const uint4*>(a.yraw) + (size_t)tile * tile_chunks + tid;
#pragma unroll
    for (int i = 0; i < 4; ++i) { qm[i] = pm[i * 128]; qy[i] = py[i * 128]; }
  };
  auto load_grad = [&](int tile) {
    const uint4* p1 = reinterpret_cast<const uint4*>(a.g1) + (size_t)tile * tile_chunks + tid;
#pragma unroll
    for (int i = 0; i < 4; ++i) qg[i] = p1[i * 128];
    if (a.g2) {
      const uint4* p2 = reinterpret_cast<const uint4*>(a.g2) + (size_t)tile * tile_chunks + tid;
#pragma unroll
      for (int i = 0; i < 4; ++i) qg2[i] = p2[i * 128];
    }
  };
  auto unpack8 = [](const uint4& u, float* v) {
    v[0] = __uint_as_float(u.x << 16); v[1] = __uint_as_float(u.x & 0xffff0000u);
    v[2] = __uint_as_float(u.y << 16); v[3] = __uint_as_float(u.y & 0xffff0000u);
    v[4] = __uint_as_float(u.z << 16); v[5] = __uint_as_float(u.z & 0xffff0000u);
    v[6] = __uint_as_float(u.w << 16); v[7] = __uint_as_float(u.w & 0xffff0000u);
  };
  auto transform = [&](int bslot) {                  // registers -> bf16 dY chunks in the B operand plane (tid & 3)
    unsigned char* bp = stc_smem + (b_base - smem_u32(stc_smem)) + bslot * kBBuf + (tid & 3) * kPlane + (tid >> 2) * 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float g[8], v[8], o[8];
      unpack8(qg[i], g);
      if (a.g2) {
        unpack8(qg2[i], v);
#pragma unroll
        for (int k = 0; k < 8; ++k) g[k] = __bfloat162float(__float2bfloat16_rn(g[k] + v[k]));
      }
      unpack8(qm[i], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] = v[k] > 0.f ? g[k] : 0.f;
      unpack8(qy[i], v);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = fmaf(cA[k], g[k], fmaf(cB[k], v[k], cD[k]));
      *reinterpret_cast<uint4*>(bp + i * 32 * 16) =
          make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
    }
  };

  int it = 0;
  const int t0 = blockIdx.x, tstep = gridDim.x;
  if constexpr (kFuse) {
    if (t0 < a.ntiles) load_fwd(t0);                 // forward-written operands of the first tile: under the predecessor's tail
  }
  pdl_wait();
  pdl_trigger();
  if constexpr (kFuse) {
    const int ch = (tid & 3) * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float sc = a.bcoef[ch + k], c1 = a.bcoef[CO + ch + k], c2 = a.bcoef[2 * CO + ch + k];
      const float mean = a.stat[ch + k], rstd = a.stat[CO + ch + k];
      const float t = sc * c2 * rstd;
      cA[k] = sc; cB[k] = -t; cD[k] = fmaf(t, mean, -sc * c1);
    }
    if (t0 < a.ntiles) load_grad(t0);
  }
  if (t0 < a.ntiles) prefetch(t0, 0);
  cp_async_commit();
  if (t0 + tstep < a.ntiles) prefetch(t0 + tstep, 1);
  cp_async_commit();
  for (int tile = t0; tile < a.ntiles; tile += tstep, ++it) {
    const int slot = it % kSlots, buf = it & 1;
    cp_async_wait<1>();                              // tile `it` has landed (at most the next tile's group is pending)
    __syncthreads();
    // the MMAs of tile it-1 read B[(it-1) % 3] (= the slot refilled below) and A[buf ^ 1]; those of tile it-2 read A[buf]
    if (it >= 2) mbar_wait(smem_u32(&mma_done[buf]), (uint32_t)(((it - 2) >> 1) & 1));
    // ---- build this pixel's im2col row (25 taps, hi / lo) into A[buf] ----
    {
      const float* xr = xs3[slot] + (2 * ry) * a.pitch + 2 * ox;
      float hi[32], lo[32];
#pragma unroll
      for (int kh = 0; kh < 5; ++kh) {
        const float2 x01 = *reinterpret_cast<const float2*>(xr + kh * a.pitch);
        const float2 x23 = *reinterpret_cast<const float2*>(xr + kh * a.pitch + 2);
        const float x4 = xr[kh * a.pitch + 4];
        split_bf16(x01.x, hi[kh * 5 + 0], lo[kh * 5 + 0]); split_bf16(x01.y, hi[kh * 5 + 1], lo[kh * 5 + 1]);
        split_bf16(x23.x, hi[kh * 5 + 2], lo[kh * 5 + 2]); split_bf16(x23.y, hi[kh * 5 + 3], lo[kh * 5 + 3]);
        split_bf16(x4, hi[kh * 5 + 4], lo[kh * 5 + 4]);
      }
#pragma unroll
      for (int k = 25; k < 32; ++k) { hi[k] = 0.f; lo[k] = 0.f; }
      unsigned char* ap = a_ptr + buf * kABuf + tid * 16;
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) {
        *reinterpret_cast<uint4*>(ap + k8 * kPlane) =
            make_uint4(pack_bf16x2(hi[k8 * 8], hi[k8 * 8 + 1]), pack_bf16x2(hi[k8 * 8 + 2], hi[k8 * 8 + 3]),
                       pack_bf16x2(hi[k8 * 8 + 4], hi[k8 * 8 + 5]), pack_bf16x2(hi[k8 * 8 + 6], hi[k8 * 8 + 7]));
        *reinterpret_cast<uint4*>(ap + kOper + k8 * kPlane) =
            make_uint4(pack_bf16x2(lo[k8 * 8], lo[k8 * 8 + 1]), pack_bf16x2(lo[k8 * 8 + 2], lo[k8 * 8 + 3]),
                       pack_bf16x2(lo[k8 * 8 + 4], lo[k8 * 8 + 5]), pack_bf16x2(lo[k8 * 8 + 6], lo[k8 * 8 + 7]));
      }
    }
    if constexpr (kFuse) transform(slot);            // B[slot]: last read by the MMAs of tile it-3, complete (wait above)
    fence_proxy_async_smem();                        // operand writes (generic proxy, and cp.async) -> the tensor core
    tc_fence_before();
    __syncthreads();
    if constexpr (kFuse) {                           // the next tile's raw vectors fly while this tile's MMAs run
      if (tile + tstep < a.ntiles) { load_fwd(tile + tstep); load_grad(tile + tstep); }
    }
    if (warp == 0) {
      tc_fence_after();
      if (elect_one()) {
        const uint32_t ab = a_base + (uint32_t)(buf * kABuf), bb = b_base + (uint32_t)(slot * kBBuf);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {             // 16 pixels per MMA: two core matrices of 8 pixels, 128 B apart
          const uint64_t d_b = make_smem_desc(bb + ks * 256, 128, kPlane, SWZ_NONE);
          mma_bf16(tmem, make_smem_desc(ab + ks * 256, 128, kPlane, SWZ_NONE), d_b, idesc, (it | ks) != 0);
          mma_bf16(tmem, make_smem_desc(ab + kOper + ks * 256, 128, kPlane, SWZ_NONE), d_b, idesc, 1);
        }
        mma_commit(smem_u32(&mma_done[buf]));
      }
      __syncwarp();
    }
    // ---- refill the slot of tile it-1 with tile it+2: its readers (the MMAs of tile it-1) must be done ----
    if (it >= 1) mbar_wait(smem_u32(&mma_done[buf ^ 1]), (uint32_t)(((it - 1) >> 1) & 1));
    const int nxt = tile + 2 * tstep;
    if (nxt < a.ntiles) prefetch(nxt, (it + 2) % kSlots);
    cp_async_commit();
  }
  cp_async_wait<0>();
  // all MMAs of this CTA: the last commit covers every earlier one
  if (it >= 1) mbar_wait(smem_u32(&mma_done[(it - 1) & 1]), (uint32_t)(((it - 1) >> 1) & 1));
  tc_fence_after();
  if (warp == 0 && it >= 1) {
    // TMEM lane = tap (0..24), column = channel
    float v[32];
    tmem_ld32(tmem, v);
    if (lane < 25) {
#pragma unroll
      for (int co = 0; co < CO; ++co) atomicAdd(a.dw + (size_t)co * 25 + lane, v[co]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32u);
  }
}

}  // namespace

// Wo a power of two with 128 % Wo == 0 and whole tiles per frame: S = 64 (4 rows per tile), S = 32 (8 rows), S = 16
bool stem_tc_supported(int Co, int S) {
  static const bool off = getenv("MMVAE_NO_STEM_TC") != nullptr;
  const int Wo = S / 2;
  if (off || (Co != 32 && Co != 64) || (S & 1) || Wo < 8 || Wo > 32 || (Wo & (Wo - 1)) != 0) return false;
  const int R = 128 / Wo;
  return (S / 2) % R == 0 && (2 * R + 3) * (2 * Wo + 4) <= kStemTcXs;
}

bool stem_wgrad_tc_supported(int Co, int S) {
  static const bool off = getenv("MMVAE_NO_STEM_WGRAD_TC") != nullptr;
  return !off && Co == 32 && stem_tc_supported(Co, S);
}

void launch_stem_wgrad_tc(StemArgs a, cudaStream_t st) {
  a.Ho = a.S / 2; a.Wo = a.S / 2;
  a.R = 128 / a.Wo;
  a.tiles_per_frame = a.Ho / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
  a.pitch = 2 * a.Wo + 4;
  // B slots + two A buffers + room for the phantom planes 4..15 of the last (lo) operand of the second A buffer
  const size_t smem = 2 * 16384 + 3 * 4 * 2048 + 128;        // two A buffers, three B slots
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(stem_wgrad_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(stem_wgrad_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_done = true;
  }
  static const int per_sm = getenv("MMVAE_STEM_WG_CTAS") ? atoi(getenv("MMVAE_STEM_WG_CTAS")) : 3;
  const int grid = a.ntiles < 148 * per_sm ? a.ntiles : 148 * per_sm;
  count_launch();
  if (a.yraw) launch_pdl(stem_wgrad_tc_kernel<true>, grid, 128, smem, st, a);
  else launch_pdl(stem_wgrad_tc_kernel<false>, grid, 128, smem, st, a);
}

void launch_stem_fwd_tc(StemArgs a, int Co, cudaStream_t st) {
  a.Ho = a.S / 2; a.Wo = a.S / 2;
  a.R = 128 / a.Wo;
  a.tiles_per_frame = a.Ho / a.R;
  a.ntiles = a.N * a.tiles_per_frame;
  a.pitch = 2 * a.Wo + 4;
  const size_t smem = 2 * (2 * 4 * 128 * 16) + 2 * 4 * (size_t)Co * 16 + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(stem_fwd_tc_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(stem_fwd_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr_done = true;
  }
  static const int per_sm_env = getenv("MMVAE_STEM_CTAS") ? atoi(getenv("MMVAE_STEM_CTAS")) : 0;
  const int per_sm = per_sm_env ? per_sm_env : (Co == 32 ? 4 : 3);
  const int grid = a.ntiles < 148 * per_sm ? a.ntiles : 148 * per_sm;
  count_launch();
  if (Co == 32) launch_pdl(stem_fwd_tc_kernel<32>, grid, 128, smem, st, a);
  else launch_pdl(stem_fwd_tc_kernel<64>, grid, 128, smem, st, a);
}

}  // namespace mmvae
