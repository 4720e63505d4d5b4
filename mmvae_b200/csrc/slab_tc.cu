// slab_tc.cu -- tcgen05 kernels for the narrow (16 / 32 channel), large-image layers of the decoder: the
// ConvTranspose2d k4 s2 p1 of uplayer4 / uplayer5 (model.py:62-65), forward and data gradient.
//
// Why a second kernel family: with 16 channels a pixel is 32 bytes, and an im2col TMA box row of 32 bytes costs the
// TMA unit as much as a 128-byte one -- gconv_tc_kernel is bound by the TMA row rate on these layers,
// and it fetches every input pixel once per tap (16x).  Here a band of image rows is copied into shared memory ONCE
// (cp.async, 16 bytes per thread, coalesced), in the un-swizzled canonical operand layout
//     plane[channel / 8][padded position][8 channels]      position = (row + 1) * (W + 2) + (col + 1), zero halo
// in which the pixel axis has a uniform 16-byte pitch.  A convolution tap is then nothing but a shifted start address
// of the UMMA shared-memory descriptor (LBO = plane stride, SBO = 128): 9 (forward) / 16 (data gradient) tcgen05.mma
// per 128 positions walk the taps over the same resident copy.  Rows of the accumulator that fall on halo positions
// are computed and dropped.
//
//   slab_fwd_kernel    all four output parities of the transposed conv in one GEMM: D[position][(py, px, co)] =
//                      sum over the 3x3 input neighbourhood of x[position + (dy, dx)] * Wf[(dy, dx)][ci][(py, px, co)]
//                      (Wf = the 4x4 kernel scattered by parity, zeros elsewhere), N = 64, + fused BatchNorm statistics
//   slab_dgrad_kernel  dX[i, j] = sum_{ky, kx} dY[2i - 1 + ky, 2j - 1 + kx] W[:, :, ky, kx]^T: dY is staged as its four
//                      parity sub-images, each tap is (sub-image, shift); optional accumulate onto the main branch's
//                      gradient and fused ReLU mask + BatchNorm-backward sums of the consumer (BnBwdFused)
//
//   slab_wgrad_kernel  dW[ci][co][ky][kx] = sum over positions of x[p - shift] * dYq[p]: the pixel axis is the GEMM K axis
//                      (both operands MN-major views of the same plane layout); A = the four dY sub-images stacked along
//                      M, B = three copies of x pre-shifted by 0 / +1 / -1 columns stacked along N, one MMA per row shift;
//                      the three accumulators live in TMEM for the CTA's whole life, one epilogue with atomics at the end
//
// Warp roles: 4 producer warps (cp.async), 1 MMA warp, NSETS x 4 epilogue warps (set s drains TMEM accumulator s).
#include "bn_fused.cuh"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

using namespace tc;

namespace {

constexpr int kProd = 128;                           // producer threads (warps 0-3)

struct SlabGeom {
  int N, H, W;                 // small grid: the conv's input (forward) / dX (data gradient)
  int P;                       // padded pitch W + 2
  int Hb, bands;               // band rows, bands per image
  int T;                       // 128-position tiles per band (positions P .. (Hb+1)*P - 1, rounded up)
  int guard;                   // chunks in front of position 0 of a plane
  int plane_bytes;             // bytes per plane (guard + positions + tile overrun + guard)
  int nslabs;                  // N * bands
  int wlog2;                   // log2(W)
  FastDiv fd_p, fd_bands;
};

struct SlabFwdParams {
  unsigned long long* trace;   // debugging timeline (mmvae_debug_set_trace) or nullptr
  const __nv_bfloat16* x;      // [N][H][W][Cx]
  __nv_bfloat16* y;            // [N][2H][2W][16]
  const float* w;              // fp32 [Cx][16][4][4]
  BnFused bn;
  int Cx;
  SlabGeom g;
};

struct SlabDgradParams {
  const __nv_bfloat16* dy;     // [N][2H][2W][16]
  __nv_bfloat16* dx;           // [N][H][W][Cn]
  const float* w;              // fp32 [Cn][16][4][4]
  int accumulate;
  BnBwdFused bb;
  int Cn;
  SlabGeom g;
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define SLAB_TRACE(p, s) do { if ((p).trace) (p).trace[(size_t)blockIdx.x * 32 + (s)] = gtimer(); } while (0)

__device__ __forceinline__ void zero_smem(unsigned char* base, size_t bytes, int nthreads) {
  uint4* p = reinterpret_cast<uint4*>(base);
  for (size_t e = threadIdx.x; e < bytes / 16; e += nthreads) p[e] = make_uint4(0u, 0u, 0u, 0u);
}

// ------------------------------------------------------------------------------------------------
// forward: ConvTranspose2d k4 s2 p1, Cx -> 16 channels, all four output parities per input position
// ------------------------------------------------------------------------------------------------
constexpr int kFwdSets = 3;                          // epilogue warp sets
constexpr int kFwdBufs = 4;                          // TMEM accumulators (tile g -> buffer g % 4, set g % 3): MMA runs ahead of the drain
constexpr int kFwdThreads = (4 + 1 + 4 * kFwdSets) * 32;

template <int kCx>
__global__ void __launch_bounds__(kFwdThreads, 1) slab_fwd_kernel(const __grid_constant__ SlabFwdParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[2], empty[2], tfull[kFwdBufs], tempty[kFwdBufs];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_red[4 * kFwdSets][2][16];
  const SlabGeom& G = p.g;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  constexpr int planes = kCx >> 3, plog2 = kCx == 16 ? 1 : 2;
  if (tid == 0) SLAB_TRACE(p, 0);
  const uint32_t xbytes = (uint32_t)planes * (uint32_t)G.plane_bytes;      // one buffer of the band
  const uint32_t tapB = (uint32_t)planes * 1024u;                          // weights of one tap: [Cx/8][64 rows][16 B]
  unsigned char* wsm = smem + 2 * (size_t)xbytes;

  // ---- prologue (kernel parameters and constant weights only: overlaps the previous kernel's tail) ----
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&full[b]), kProd); mbar_init(smem_u32(&empty[b]), 1); }
    for (int s = 0; s < kFwdBufs; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 256);
  zero_smem(smem, 2 * (size_t)xbytes + 9 * (size_t)tapB, kFwdThreads);
  __syncthreads();
  // Wf[tap = (dy+1)*3 + (dx+1)][j = ci/8][n = (py*2+px)*16 + co][ci % 8] = W[ci][co][ky][kx] with ky = py + 1 - 2*dy
  // (oy = 2*iy - 1 + ky): every weight lands in exactly one (tap, parity) slot, the rest of Wf stays zero.
  for (int e = tid; e < kCx * 16 * 4; e += kFwdThreads) {          // one float4 = W[ci][co][ky][0..3]
    const int ky = e & 3, co = (e >> 2) & 15, ci = e >> 6;
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w) + e);
    const int py = (ky + 1) & 1, dy = (py + 1 - ky) / 2;
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int kx = 0; kx < 4; ++kx) {
      const int px = (kx + 1) & 1, dx = (px + 1 - kx) / 2;
      const int tap = (dy + 1) * 3 + dx + 1, n = (py * 2 + px) * 16 + co;
      reinterpret_cast<__nv_bfloat16*>(wsm)[(size_t)tap * (tapB >> 1) + (size_t)(ci >> 3) * 512 + n * 8 + (ci & 7)] = __float2bfloat16_rn(wv[kx]);
    }
  }
  fence_proxy_async_smem();                          // zeros and weights -> visible to the tensor core's reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) SLAB_TRACE(p, 1);
  pdl_wait();
  pdl_trigger();
  if (tid == 0) SLAB_TRACE(p, 2);

  const uint32_t x_base = smem_u32(smem), w_base = smem_u32(wsm);
  if (warp < 4) {
    // ---------------- producers: one band (+ halo rows) per slab, 16 bytes per cp.async ----------------
    const int per_row = G.W * planes;
    const int total = (G.Hb + 2) * per_row;
    int it = 0;
    for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(smem_u32(&empty[buf]), (uint32_t)(((it >> 1) & 1) ^ 1));
      int n, b;
      G.fd_bands.divmod(s, n, b);
      const int h0 = b * G.Hb;
      const uint32_t dst0 = x_base + (uint32_t)buf * xbytes + (uint32_t)G.guard * 16u;
      for (int e = tid; e < total; e += kProd) {
        const int pl = e & (planes - 1), t = e >> plog2;
        const int c = t & (G.W - 1), r = (t >> G.wlog2) - 1;
        const int hy = h0 + r;
        const bool ok = (unsigned)hy < (unsigned)G.H;
        const __nv_bfloat16* src = ok ? p.x + (((size_t)n * G.H + hy) * G.W + c) * kCx + pl * 8 : p.x;
        cp_async16(dst0 + (uint32_t)pl * (uint32_t)G.plane_bytes + (uint32_t)((r + 1) * G.P + c + 1) * 16u, src, ok ? 16u : 0u);
      }
      cp_async_commit();
      if (tid == 0 && it == 0) SLAB_TRACE(p, 3);
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&full[buf]));
      if (tid == 0 && it == 0) SLAB_TRACE(p, 4);
      if (tid == 0 && it == 1) SLAB_TRACE(p, 5);
    }
  } else if (warp == 4) {
    {
      // ---------------- MMA issue: 9 taps x (Cx / 16) k-steps per tile ----------------
      // the whole warp walks the loops (uniform control flow: descriptors stay in uniform registers), one elected lane
      // issues -- see elect_one(), tc_common.cuh
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      const uint64_t da0 = make_smem_desc(x_base, (uint32_t)G.plane_bytes, 128, SWZ_NONE);
      const uint64_t db0 = make_smem_desc(w_base, 1024, 128, SWZ_NONE);
      constexpr int ksteps = planes >> 1;
      // descriptor offsets ((address >> 4) field) of every (tap, k-step): constant over tiles, kept in registers
      uint32_t aoff[9 * ksteps], boff[9 * ksteps];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap)
#pragma unroll
        for (int kk = 0; kk < ksteps; ++kk) {
          aoff[tap * ksteps + kk] = (uint32_t)((tap / 3 - 1) * G.P + (tap % 3 - 1)) + (((uint32_t)(2 * kk) * (uint32_t)G.plane_bytes) >> 4);
          boff[tap * ksteps + kk] = ((uint32_t)tap * tapB + (uint32_t)kk * 2048u) >> 4;
        }
      int it = 0, gt = 0;
      for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&full[buf]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        if (it == 0 && lane == 0) SLAB_TRACE(p, 6);
        for (int t = 0; t < G.T; ++t, ++gt) {
          const int ab = gt % kFwdBufs;
          mbar_wait(smem_u32(&tempty[ab]), (uint32_t)(((gt / kFwdBufs) & 1) ^ 1));
          tc_fence_after();
          const uint32_t dtm = tmem + (uint32_t)(ab * 64);
          // taps shift the start address down as well as up: 32-bit arithmetic on the descriptor's low word
          const uint32_t dat = (uint32_t)da0 + (((uint32_t)buf * xbytes) >> 4) + (uint32_t)(G.guard + G.P + 128 * t);
          if (elect_one()) {
#pragma unroll
            for (int m = 0; m < 9 * ksteps; ++m)
              mma_bf16(dtm, (da0 & 0xFFFFFFFF00000000ull) | (uint64_t)(dat + aoff[m]), db0 + (uint64_t)boff[m], idesc, m != 0);
            mma_commit(smem_u32(&tfull[ab]));
          }
          __syncwarp();
          if (gt == 0 && lane == 0) SLAB_TRACE(p, 7);
        }
        if (elect_one()) mma_commit(smem_u32(&empty[buf]));   // the band's shared-memory copy is free once these MMAs are done
        __syncwarp();
        if (it == 0 && lane == 0) SLAB_TRACE(p, 8);
      }
      if (lane == 0) SLAB_TRACE(p, 9);
    }
  } else {
    // ---------------- epilogue: set `es` drains accumulator `es` (tiles es, es + kFwdSets, ...) ----------------
    const int ew = warp - 5, es = ew >> 2, q = warp & 3;
    const bool stats = p.bn.acc != nullptr;
    // BatchNorm sums of this thread's rows, two channels per 64-bit register (FADD2 / FFMA2): the epilogue sets are bound
    // by instruction issue (3 warps per scheduler: profiles/r02_slab_epilogue.md), so the statistics are taken from the
    // fp32 accumulators as they come out of TMEM -- no unpacking of the rounded pairs, 4 packed instructions per 4 values
    // instead of 12 scalar ones.  (The mean / variance of the unrounded values differ from those of the stored bf16 values
    // by the mean of ~1e6 independent rounding errors per channel: ~1e-6 relative.)
    unsigned long long rs2[8], rq2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { rs2[e] = 0ull; rq2[e] = 0ull; }
    const int r = q * 32 + lane;
    for (int gt = es;; gt += kFwdSets) {
      const int it = gt / G.T, t = gt - it * G.T;
      const int s = blockIdx.x + it * gridDim.x;
      if (s >= G.nslabs) break;
      int n, b;
      G.fd_bands.divmod(s, n, b);
      int rr, cc;
      G.fd_p.divmod(G.P + 128 * t + r, rr, cc);
      const bool valid = rr <= G.Hb && cc >= 1 && cc <= G.W;
      const int i = b * G.Hb + rr - 1, j = cc - 1;
      const int ab = gt % kFwdBufs;
      mbar_wait(smem_u32(&tfull[ab]), (uint32_t)((gt / kFwdBufs) & 1));
      tc_fence_after();
      if (gt == 0 && ew == 0 && lane == 0) SLAB_TRACE(p, 10);
      const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * 64);
#pragma unroll
      for (int py = 0; py < 2; ++py) {
        uint32_t r0[16], r1[16];                     // both loads in flight, one wait
        tmem_ld16_issue(tl + (uint32_t)(py * 32), r0);
        tmem_ld16_issue(tl + (uint32_t)(py * 32 + 16), r1);
        tmem_ld_wait();
        if (valid) {
          uint4* dst = reinterpret_cast<uint4*>(p.y + (((size_t)n * 2 * G.H + 2 * i + py) * 2 * G.W + 2 * j) * 16);
          uint32_t pk[16];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            pk[e] = pack_bf16x2(__uint_as_float(r0[2 * e]), __uint_as_float(r0[2 * e + 1]));
            pk[8 + e] = pack_bf16x2(__uint_as_float(r1[2 * e]), __uint_as_float(r1[2 * e + 1]));
          }
          dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
          dst[2] = make_uint4(pk[8], pk[9], pk[10], pk[11]);
          dst[3] = make_uint4(pk[12], pk[13], pk[14], pk[15]);
          if (stats) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {            // channels (2e, 2e + 1) of the two output columns px = 0, 1
              const unsigned long long a2 = f2_pack(__uint_as_float(r0[2 * e]), __uint_as_float(r0[2 * e + 1]));
              const unsigned long long b2 = f2_pack(__uint_as_float(r1[2 * e]), __uint_as_float(r1[2 * e + 1]));
              rs2[e] = f2_add(rs2[e], f2_add(a2, b2));
              rq2[e] = f2_fma(a2, a2, f2_fma(b2, b2, rq2[e]));
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[ab]));
      if (gt == 0 && ew == 0 && lane == 0) SLAB_TRACE(p, 11);
    }
    float run_s[16], run_q[16];
#pragma unroll
    for (int e = 0; e < 8; ++e) { f2_unpack(rs2[e], run_s[2 * e], run_s[2 * e + 1]); f2_unpack(rq2[e], run_q[2 * e], run_q[2 * e + 1]); }
    if (ew == 0 && lane == 0) SLAB_TRACE(p, 12);
    if (stats) {
      const int lane_col = warp_colsum16(run_s, lane);
      warp_colsum16(run_q, lane);
      if ((lane & 1) == 0) { stat_red[ew][0][lane_col] = run_s[0]; stat_red[ew][1][lane_col] = run_q[0]; }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kFwdSets) : "memory");
      const int et = ew * 32 + lane;
      if (et < 32) {
        const int which = et >> 4, c = et & 15;
        float t = 0.f;
#pragma unroll
        for (int w8 = 0; w8 < 4 * kFwdSets; ++w8) t += stat_red[w8][which][c];
        atomicAdd(bn_acc_copy(p.bn) + which * p.bn.C + c, (double)t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
  if (tid == 0) SLAB_TRACE(p, 13);
  if (p.bn.acc) bn_fused_finish(p.bn, gridDim.x);
  if (tid == 0) SLAB_TRACE(p, 14);
}

// ------------------------------------------------------------------------------------------------
// data gradient of ConvTranspose2d k4 s2 p1: 16 -> Cn channels, a stride-2 4x4 convolution over dY
// ------------------------------------------------------------------------------------------------
constexpr int kDgSets = 3;
constexpr int kDgBufs = 2 * kDgSets;
constexpr int kDgThreads = (4 + 1 + 4 * kDgSets) * 32;

template <int kCn, bool kFuse>
__global__ void __launch_bounds__(kDgThreads, 1) slab_dgrad_kernel(const __grid_constant__ SlabDgradParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[2], empty[2], tfull[kDgBufs], tempty[kDgBufs];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_red[4 * kDgSets][3][16];
  const SlabGeom& G = p.g;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  constexpr int kPlanes = 8;                          // 4 parity sub-images x 2 planes (16 channels)
  constexpr uint32_t kTapB = 2u * kCn * 16u;          // weights of one tap: [2][Cn rows][16 B]
  constexpr uint32_t kCols = kDgBufs * kCn <= 128 ? 128u : 256u;
  const uint32_t ybytes = (uint32_t)kPlanes * (uint32_t)G.plane_bytes;
  unsigned char* wsm = smem + 2 * (size_t)ybytes;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&full[b]), kProd); mbar_init(smem_u32(&empty[b]), 1); }
    for (int s = 0; s < kDgBufs; ++s) { mbar_init(smem_u32(&tfull[s]), 1); mbar_init(smem_u32(&tempty[s]), 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), kCols);
  zero_smem(smem, 2 * (size_t)ybytes, kDgThreads);
  // B[tap = ky*4 + kx][j = co/8][n = ci][co % 8] = W[ci][co][ky][kx]; one float4 = W[ci][co][ky][0..3]
  for (int e = tid; e < kCn * 16 * 4; e += kDgThreads) {
    const int ky = e & 3, co = (e >> 2) & 15, ci = e >> 6;
    const float4 w4 = __ldg(reinterpret_cast<const float4*>(p.w) + e);
    const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int kx = 0; kx < 4; ++kx)
      reinterpret_cast<__nv_bfloat16*>(wsm)[(size_t)(ky * 4 + kx) * (kTapB >> 1) + (size_t)(co >> 3) * (kCn * 8) + ci * 8 + (co & 7)] =
          __float2bfloat16_rn(wv[kx]);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const uint32_t y_base = smem_u32(smem), w_base = smem_u32(wsm);
  if (warp < 4) {
    // ---------------- producers: dY rows 2*(h0-1) .. 2*(h0+Hb)+1, split into the four parity sub-images ----------------
    const int W2 = 2 * G.W;
    const int per_row = W2 * 2;                       // 16-byte chunks per dY row
    const int total = (2 * G.Hb + 4) * per_row;
    int it = 0;
    for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(smem_u32(&empty[buf]), (uint32_t)(((it >> 1) & 1) ^ 1));
      int n, b;
      G.fd_bands.divmod(s, n, b);
      const int h0 = b * G.Hb;
      const uint32_t dst0 = y_base + (uint32_t)buf * ybytes + (uint32_t)G.guard * 16u;
      for (int e = tid; e < total; e += kProd) {
        const int pl = e & 1, t = e >> 1;
        const int C = t & (W2 - 1), Rr = t >> (G.wlog2 + 1);             // dY column, row relative to 2*(h0-1)
        const int gy = 2 * (h0 - 1) + Rr;
        const bool ok = (unsigned)gy < (unsigned)(2 * G.H);
        const __nv_bfloat16* src = ok ? p.dy + (((size_t)n * 2 * G.H + gy) * W2 + C) * 16 + pl * 8 : p.dy;
        const int qy = Rr & 1, a = Rr >> 1, qx = C & 1, bb = C >> 1;      // sub-image position (a - 1, bb)
        cp_async16(dst0 + (uint32_t)((qy * 2 + qx) * 2 + pl) * (uint32_t)G.plane_bytes + (uint32_t)(a * G.P + bb + 1) * 16u, src,
                   ok ? 16u : 0u);
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&full[buf]));
    }
  } else if (warp == 4) {
    {
      // ---------------- MMA issue: 16 taps per tile; tap ky reads sub-image row parity qy at row shift da ----------------
      // (whole warp, one elected lane issues: elect_one(), tc_common.cuh)
      const uint32_t idesc = make_idesc_bf16(128, kCn, 0, 0);
      const uint64_t da0 = make_smem_desc(y_base, (uint32_t)G.plane_bytes, 128, SWZ_NONE);
      const uint64_t db0 = make_smem_desc(w_base, kCn * 16, 128, SWZ_NONE);
      // descriptor offsets ((address >> 4) field) of the 16 taps: constant over tiles, kept in registers
      uint32_t aoff[16];
#pragma unroll
      for (int tap = 0; tap < 16; ++tap) {
        const int ky = tap >> 2, kx = tap & 3;
        const int qy = (ky + 1) & 1, da = (ky - 1 - qy) / 2;              // 2i - 1 + ky = 2*(i + da) + qy
        const int qx = (kx + 1) & 1, db_ = (kx - 1 - qx) / 2;
        aoff[tap] = (((uint32_t)((qy * 2 + qx) * 2) * (uint32_t)G.plane_bytes) >> 4) + (uint32_t)(da * G.P + db_);
      }
      int it = 0, gt = 0;
      for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&full[buf]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        for (int t = 0; t < G.T; ++t, ++gt) {
          const int ab = gt % kDgBufs;
          mbar_wait(smem_u32(&tempty[ab]), (uint32_t)(((gt / kDgBufs) & 1) ^ 1));
          tc_fence_after();
          const uint32_t dtm = tmem + (uint32_t)(ab * kCn);
          // taps shift the start address down as well as up: 32-bit arithmetic on the descriptor's low word
          const uint32_t dat = (uint32_t)da0 + (((uint32_t)buf * ybytes) >> 4) + (uint32_t)(G.guard + G.P + 128 * t);
          if (elect_one()) {
#pragma unroll
            for (int tap = 0; tap < 16; ++tap)
              mma_bf16(dtm, (da0 & 0xFFFFFFFF00000000ull) | (uint64_t)(dat + aoff[tap]), db0 + (uint64_t)((uint32_t)tap * (kTapB >> 4)), idesc,
                       tap != 0);
            mma_commit(smem_u32(&tfull[ab]));
          }
          __syncwarp();
        }
        if (elect_one()) mma_commit(smem_u32(&empty[buf]));
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue ----------------
    const int ew = warp - 5, es = ew >> 2, q = warp & 3;
    constexpr int NG = kCn / 16;
    float run0[kFuse ? 16 : 1], run1[kFuse ? 16 : 1], run2[kFuse ? 16 : 1];
#pragma unroll
    for (int e = 0; e < (kFuse ? 16 : 1); ++e) { run0[e] = 0.f; run1[e] = 0.f; run2[e] = 0.f; }
    const bool two = kFuse && p.bb.y2 != nullptr;
    const int r = q * 32 + lane;
    for (int gt = es;; gt += kDgSets) {
      const int it = gt / G.T, t = gt - it * G.T;
      const int s = blockIdx.x + it * gridDim.x;
      if (s >= G.nslabs) break;
      int n, b;
      G.fd_bands.divmod(s, n, b);
      int rr, cc;
      G.fd_p.divmod(G.P + 128 * t + r, rr, cc);
      const bool valid = rr <= G.Hb && cc >= 1 && cc <= G.W;
      const size_t obase = (((size_t)n * G.H + (b * G.Hb + rr - 1)) * G.W + (cc - 1)) * kCn;
      const int ab = gt % kDgBufs;
      mbar_wait(smem_u32(&tfull[ab]), (uint32_t)((gt / kDgBufs) & 1));
      tc_fence_after();
      const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(ab * kCn);
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        float v[16];
        tmem_ld16(tl + (uint32_t)(gq * 16), v);
        if (valid) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const size_t o = obase + gq * 16 + h * 8;
            uint4* dst = reinterpret_cast<uint4*>(p.dx + o);
            if (p.accumulate) {
              const uint4 old = *dst;
              const __nv_bfloat16* ob = reinterpret_cast<const __nv_bfloat16*>(&old);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[h * 8 + e] += __bfloat162float(ob[e]);
            }
            if constexpr (kFuse) {
              // g = bf16(dX) * [a > 0] is what gets stored; S0 = sum g, S1' = sum g*y, S2' = sum g*y2 (centred at the end)
              const uint4 yr = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.y) + o));
              uint4 ar = make_uint4(0u, 0u, 0u, 0u), zr = make_uint4(0u, 0u, 0u, 0u);
              if (p.bb.a) ar = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.a) + o));
              if (two) zr = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.y2) + o));
              const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(&yr);
              const __nv_bfloat16* ab_ = reinterpret_cast<const __nv_bfloat16*>(&ar);
              const __nv_bfloat16* zb = reinterpret_cast<const __nv_bfloat16*>(&zr);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float gm = __bfloat162float(__float2bfloat16_rn(v[h * 8 + e]));
                if (p.bb.a && !(__bfloat162float(ab_[e]) > 0.f)) gm = 0.f;
                v[h * 8 + e] = gm;
                run0[h * 8 + e] += gm;
                run1[h * 8 + e] = fmaf(gm, __bfloat162float(yb[e]), run1[h * 8 + e]);
                run2[h * 8 + e] = fmaf(gm, __bfloat162float(zb[e]), run2[h * 8 + e]);
              }
            }
            *dst = make_uint4(pack_bf16x2(v[h * 8 + 0], v[h * 8 + 1]), pack_bf16x2(v[h * 8 + 2], v[h * 8 + 3]),
                              pack_bf16x2(v[h * 8 + 4], v[h * 8 + 5]), pack_bf16x2(v[h * 8 + 6], v[h * 8 + 7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[ab]));
    }
    if constexpr (kFuse) {
      const int lane_col = warp_colsum16(run0, lane);
      warp_colsum16(run1, lane);
      warp_colsum16(run2, lane);
      if ((lane & 1) == 0) { stat_red[ew][0][lane_col] = run0[0]; stat_red[ew][1][lane_col] = run1[0]; stat_red[ew][2][lane_col] = run2[0]; }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kDgSets) : "memory");
      const int et = ew * 32 + lane;
      if (et < 16) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int w8 = 0; w8 < 4 * kDgSets; ++w8) { s0 += (double)stat_red[w8][0][et]; s1 += (double)stat_red[w8][1][et]; s2 += (double)stat_red[w8][2][et]; }
        // S1 = rstd * (sum g*y - mean * S0): the subtraction in fp64 on the CTA totals
        double* acc = bn_bwd_acc_copy(p.bb) + et;
        atomicAdd(acc, s0);
        atomicAdd(acc + p.bb.C, (double)p.bb.stat[p.bb.C + et] * (s1 - (double)p.bb.stat[et] * s0));
        if (two) atomicAdd(acc + 2 * p.bb.C, (double)p.bb.stat2[p.bb.C + et] * (s2 - (double)p.bb.stat2[et] * s0));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
  if constexpr (kFuse) {
    if (p.bb.finish) bn_bwd_fused_finish(p.bb, gridDim.x);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient of ConvTranspose2d k4 s2 p1 (Cx -> 16 channels)
// ------------------------------------------------------------------------------------------------
struct SlabWgradParams {
  const __nv_bfloat16* x;      // [N][H][W][Cx]   layer input
  const __nv_bfloat16* dy;     // [N][2H][2W][16] output gradient
  float* dw;                   // fp32 [Cx][16][4][4], atomically accumulated (pre-zeroed)
  int Cx;
  SlabGeom g;
};
constexpr int kWgThreads = (4 + 1 + 4) * 32;

template <int kCx>
__global__ void __launch_bounds__(kWgThreads, 1) slab_wgrad_kernel(const __grid_constant__ SlabWgradParams p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long full[2], empty[2], accum;
  __shared__ uint32_t tmem_base_s;
  const SlabGeom& G = p.g;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  constexpr int xplanes = kCx >> 3, xlog2 = kCx == 16 ? 1 : 2;
  constexpr int kN = 3 * kCx;                         // (column shift, ci)
  constexpr int kPlanesBuf = 8 + 3 * xplanes;         // dY: 4 sub-images x 2 planes, then 3 shifted copies of x
  constexpr uint32_t kCols = 3 * kN <= 256 ? 256u : 512u;
  const uint32_t bbytes = (uint32_t)kPlanesBuf * (uint32_t)G.plane_bytes;

  if (tid == 0) {
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&full[b]), kProd); mbar_init(smem_u32(&empty[b]), 1); }
    mbar_init(smem_u32(&accum), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), kCols);
  // padding planes: the M = 128 A operand reads 16 plane-strided atoms from the first dY plane (rows 64..127 are
  // garbage and never leave TMEM), which must stay inside the allocation for the second buffer too
  constexpr int kPad = kPlanesBuf >= 16 ? 0 : 16 - kPlanesBuf;
  zero_smem(smem, (size_t)(2 * kPlanesBuf + kPad) * (size_t)G.plane_bytes, kWgThreads);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const uint32_t base = smem_u32(smem);
  if (warp < 4) {
    // ---------------- producers ----------------
    const int W2 = 2 * G.W;
    const int ny = 2 * G.Hb * W2 * 2;                 // dY chunks of the band (rows 2*h0 .. 2*(h0+Hb)-1)
    const int nx = (G.Hb + 2) * G.W * xplanes;        // x chunks (rows h0-1 .. h0+Hb)
    int it = 0;
    for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
      const int buf = it & 1;
      mbar_wait(smem_u32(&empty[buf]), (uint32_t)(((it >> 1) & 1) ^ 1));
      int n, b;
      G.fd_bands.divmod(s, n, b);
      const int h0 = b * G.Hb;
      const uint32_t dst0 = base + (uint32_t)buf * bbytes + (uint32_t)G.guard * 16u;
      for (int e = tid; e < ny; e += kProd) {
        const int pl = e & 1, t = e >> 1;
        const int C = t & (W2 - 1), Rr = t >> (G.wlog2 + 1);             // dY column, row relative to 2*h0
        const __nv_bfloat16* src = p.dy + (((size_t)n * 2 * G.H + 2 * h0 + Rr) * W2 + C) * 16 + pl * 8;
        const int qy = Rr & 1, a = Rr >> 1, qx = C & 1, bb = C >> 1;      // sub-image position (a, bb)
        cp_async16(dst0 + (uint32_t)((qy * 2 + qx) * 2 + pl) * (uint32_t)G.plane_bytes + (uint32_t)((a + 1) * G.P + bb + 1) * 16u, src, 16u);
      }
      const uint32_t xdst0 = dst0 + 8u * (uint32_t)G.plane_bytes;
      for (int e = tid; e < nx; e += kProd) {
        const int pl = e & (xplanes - 1), t = e >> xlog2;
        const int c = t & (G.W - 1), r = (t >> G.wlog2) - 1;
        const int hy = h0 + r;
        const bool ok = (unsigned)hy < (unsigned)G.H;
        const __nv_bfloat16* src = ok ? p.x + (((size_t)n * G.H + hy) * G.W + c) * kCx + pl * 8 : p.x;
        const uint32_t d = xdst0 + (uint32_t)pl * (uint32_t)G.plane_bytes + (uint32_t)((r + 1) * G.P + c + 1) * 16u;
#pragma unroll
        for (int cp = 0; cp < 3; ++cp) {               // copy cp holds x shifted by db = (0, +1, -1)[cp] columns: copy[p] = x[p - db]
          const int db = cp == 2 ? -1 : cp;
          cp_async16(d + (uint32_t)(cp * xplanes) * (uint32_t)G.plane_bytes + (uint32_t)(db * 16), src, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      cp_async_wait<0>();
      fence_proxy_async_smem();
      mbar_arrive(smem_u32(&full[buf]));
    }
  } else if (warp == 4) {
    {
      // ---------------- MMA issue: per 16 positions, one MMA per row shift da ----------------
      // (whole warp, one elected lane issues: elect_one(), tc_common.cuh)
      // MN-major operands without swizzle: 8 channels contiguous (16 B), positions at a 16-byte pitch; the stride between
      // channel atoms (planes) goes to the SBO field, the stride between groups of 8 positions (128 B) to the LBO field.
      const uint32_t idesc = make_idesc_bf16(128, kN, 1, 1);
      const uint32_t pb = (uint32_t)G.plane_bytes;
      const uint64_t da0 = make_smem_desc(base, 128, pb, SWZ_NONE);
      const uint64_t db0 = make_smem_desc(base + 8u * pb, 128, pb, SWZ_NONE);
      const int ksteps = (G.Hb * G.P + 15) >> 4;      // positions P .. (Hb+1)*P - 1 (band rows incl. their halo columns)
      int it = 0;
      bool first = true;
      for (int s = blockIdx.x; s < G.nslabs; s += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(smem_u32(&full[buf]), (uint32_t)((it >> 1) & 1));
        tc_fence_after();
        const uint32_t p0 = (((uint32_t)buf * bbytes) >> 4) + (uint32_t)(G.guard + G.P);
        const uint32_t alo = (uint32_t)da0 + p0, blo = (uint32_t)db0 + p0;
        if (elect_one()) {
          const uint32_t acc0 = first ? 0u : 1u;
          for (int k = 0; k < ksteps; ++k) {
#pragma unroll
            for (int d3 = 0; d3 < 3; ++d3)              // da = d3 - 1: dYq[p] pairs with x[p - da*P - db]
              mma_bf16(tmem + (uint32_t)(d3 * kN), (da0 & 0xFFFFFFFF00000000ull) | (uint64_t)(alo + (uint32_t)(16 * k)),
                       (db0 & 0xFFFFFFFF00000000ull) | (uint64_t)(blo + (uint32_t)(16 * k) - (uint32_t)((d3 - 1) * G.P)), idesc,
                       k == 0 ? acc0 : 1u);
          }
          mma_commit(smem_u32(&empty[buf]));
        }
        __syncwarp();
        first = false;
      }
      if (elect_one()) mma_commit(smem_u32(&accum));
      __syncwarp();
    }
  } else {
    // ---------------- epilogue (once): rows (qy, qx, co) of the three accumulators -> dW[ci][co][ky][kx] ----------------
    mbar_wait(smem_u32(&accum), 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = q * 32 + lane;                    // TMEM lane = (qy*2 + qx)*16 + co for rows < 64
    if (q < 2) {
      const int sub = row >> 4, co = row & 15, qy = sub >> 1, qx = sub & 1;
#pragma unroll
      for (int d3 = 0; d3 < 3; ++d3) {
        const int da = d3 - 1;
        const int ky = qy ? 2 + 2 * da : 1 + 2 * da;  // 2i - 1 + ky = 2*(i + da) + qy
        const bool rowok = (unsigned)ky < 4u;
#pragma unroll
        for (int cp = 0; cp < 3; ++cp) {
          const int dbe = cp == 2 ? -1 : cp;
          const int kx = qx ? 2 + 2 * dbe : 1 + 2 * dbe;
          const bool ok = rowok && (unsigned)kx < 4u;
#pragma unroll
          for (int c0 = 0; c0 < kCx; c0 += 16) {
            float v[16];
            tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(d3 * kN + cp * kCx + c0), v);
            if (ok) {
#pragma unroll
              for (int e = 0; e < 16; ++e) atomicAdd(p.dw + ((size_t)((c0 + e) * 16 + co) * 4 + ky) * 4 + kx, v[e]);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, kCols);
  }
}

// ---------------- host ----------------
// total_planes: planes of all buffers together; tiles: the planes are read by 128-position MMA tiles (forward / data
// gradient), which may run past the band by up to one tile
bool slab_geom(SlabGeom& G, int N, int H, int W, int total_planes, int weight_bytes, int max_hb, bool tiles = true) {
  if (H != W || (W != 16 && W != 32) || N < 1) return false;
  G.N = N; G.H = H; G.W = W; G.P = W + 2;
  G.wlog2 = W == 16 ? 4 : 5;
  G.guard = G.P + 2;
  // the tallest band (a divisor of H, at most max_hb rows) whose buffers + weights fit 200 KB
  for (int hb = max_hb; hb >= 4; hb >>= 1) {
    if (H % hb) continue;
    const int T = (hb * G.P + 127) / 128;
    const int chunks = G.guard + (tiles ? G.P + 128 * T : (hb + 2) * G.P + 16) + G.guard;
    const int plane_bytes = ((chunks * 16 + 127) & ~127) + 32;       // +32: neighbouring planes land on different banks
    if ((long long)total_planes * plane_bytes + weight_bytes + 1024 > 200 * 1024) continue;
    G.Hb = hb; G.T = T; G.plane_bytes = plane_bytes;
    G.bands = H / hb; G.nslabs = N * G.bands;
    G.fd_p = FastDiv(G.P); G.fd_bands = FastDiv(G.bands);
    return true;
  }
  return false;
}

bool slab_enabled() {
  static bool on = [] { const char* e = getenv("MMVAE_NO_SLAB"); return !(e && e[0] == '1'); }();
  return on;
}

}  // namespace

bool slab_supported_gconv(const GConvParams& p) {
  if (!slab_enabled() || p.in_nchw_f32 || p.bias) return false;
  if (p.conv_class == 1) {                 // ConvTranspose2d k4 s2 p1 forward: Ci -> 16
    SlabGeom G;
    return p.Co == 16 && (p.Ci == 16 || p.Ci == 32) && !p.accumulate && !p.bb.acc &&
           slab_geom(G, p.N, p.Hi, p.Wi, 2 * (p.Ci / 8), 9 * (p.Ci / 8) * 1024, 32);
  }
  if (p.conv_class == 2) {                 // its data gradient: 16 -> Co channels (Co = the conv's Ci)
    SlabGeom G;
    if (p.bb.acc && (p.Co != 16 || p.bb.var_mask != 1 || !p.bb.finish)) return false;
    return p.Ci == 16 && (p.Co == 16 || p.Co == 32) && slab_geom(G, p.N, p.Ho, p.Wo, 16, 16 * 2 * p.Co * 16, 16);
  }
  return false;
}

void launch_slab_gconv(const GConvParams& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(slab_fwd_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaFuncSetAttribute(slab_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaFuncSetAttribute(slab_dgrad_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaFuncSetAttribute(slab_dgrad_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaFuncSetAttribute(slab_dgrad_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    attr_done = true;
  }
  count_launch();
  if (p.conv_class == 1) {
    SlabFwdParams q{};
    q.x = reinterpret_cast<const __nv_bfloat16*>(p.in); q.y = reinterpret_cast<__nv_bfloat16*>(p.out); q.w = p.w;
    q.bn = p.bn; q.Cx = p.Ci; q.trace = debug_trace_buffer();
    const int wbytes = 9 * (p.Ci / 8) * 1024;
    slab_geom(q.g, p.N, p.Hi, p.Wi, 2 * (p.Ci / 8), wbytes, 32);
    const size_t smem = 2 * (size_t)(p.Ci / 8) * q.g.plane_bytes + wbytes;
    if (p.Ci == 16) launch_pdl(slab_fwd_kernel<16>, min(q.g.nslabs, 148), kFwdThreads, smem, st, q);
    else launch_pdl(slab_fwd_kernel<32>, min(q.g.nslabs, 148), kFwdThreads, smem, st, q);
  } else {
    SlabDgradParams q{};
    q.dy = reinterpret_cast<const __nv_bfloat16*>(p.in); q.dx = reinterpret_cast<__nv_bfloat16*>(p.out); q.w = p.w;
    q.accumulate = p.accumulate; q.bb = p.bb; q.Cn = p.Co;
    const int wbytes = 16 * 2 * p.Co * 16;
    slab_geom(q.g, p.N, p.Ho, p.Wo, 16, wbytes, 16);
    const size_t smem = 2 * (size_t)8 * q.g.plane_bytes + wbytes;
    const int grid = min(q.g.nslabs, 148);
    if (p.Co == 32) launch_pdl(slab_dgrad_kernel<32, false>, grid, kDgThreads, smem, st, q);
    else if (p.bb.acc) launch_pdl(slab_dgrad_kernel<16, true>, grid, kDgThreads, smem, st, q);
    else launch_pdl(slab_dgrad_kernel<16, false>, grid, kDgThreads, smem, st, q);
  }
}

// planes of one wgrad buffer, and the padding behind the second buffer that keeps the 16-atom A operand in bounds
static int wgrad_planes(int Ci) { return 8 + 3 * (Ci / 8); }
static int wgrad_pad(int Ci) { return wgrad_planes(Ci) >= 16 ? 0 : 16 - wgrad_planes(Ci); }

bool slab_supported_wgrad(const WGradParams& p) {
  if (!slab_enabled() || p.in_nchw_f32 || p.conv_class != 1) return false;
  SlabGeom G;
  return p.Co == 16 && (p.Ci == 16 || p.Ci == 32) &&
         slab_geom(G, p.N, p.Hi, p.Wi, 2 * wgrad_planes(p.Ci) + wgrad_pad(p.Ci), 0, 8, false);
}

void launch_slab_wgrad(const WGradParams& p, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(slab_wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    cudaFuncSetAttribute(slab_wgrad_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024);
    attr_done = true;
  }
  SlabWgradParams q{};
  q.x = reinterpret_cast<const __nv_bfloat16*>(p.in); q.dy = reinterpret_cast<const __nv_bfloat16*>(p.dout); q.dw = p.dw;
  q.Cx = p.Ci;
  const int total = 2 * wgrad_planes(p.Ci) + wgrad_pad(p.Ci);
  slab_geom(q.g, p.N, p.Hi, p.Wi, total, 0, 8, false);
  const size_t smem = (size_t)total * q.g.plane_bytes;
  count_launch();
  if (p.Ci == 16) launch_pdl(slab_wgrad_kernel<16>, min(q.g.nslabs, 148), kWgThreads, smem, st, q);
  else launch_pdl(slab_wgrad_kernel<32>, min(q.g.nslabs, 148), kWgThreads, smem, st, q);
}

}  // namespace mmvae
