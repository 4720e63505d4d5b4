// gconv_tc.cu -- tcgen05 (5th-generation tensor core) implicit-GEMM kernels of the bf16 mode.
//
//   gconv_tc_kernel   forward conv / transposed conv / data gradient ("gather-convolution", kernels.cuh):
//                     D[128 pixels][BN channels] = A[128][K] * W[K][BN]; A is gathered from the NHWC bf16
//                     activation with cp.async (im2col on the fly, zero fill at the borders) into a
//                     K-major SWIZZLE_128B tile, W arrives as pre-packed, pre-swizzled tiles through the
//                     TMA engine (cp.async.bulk), the fp32 accumulator lives in TMEM, and the epilogue
//                     (tcgen05.ld) rounds to bf16, stores NHWC and emits the per-CTA BatchNorm partials.
//   wgrad_tc_kernel   weight gradient: D[128 (tap,ci)][BN co] = sum over pixels of A^T * dY; both operands
//                     are MN-major (the pixel axis is the GEMM K axis), split over pixel ranges, reduced
//                     into the fp32 gradient arena with red.global.add.
//   pack_weights_kernel  fp32 reference-layout weights -> bf16 tiles in the exact shared-memory image
//                     (8-row x 128-byte swizzle atoms) that gconv_tc_kernel bulk-copies.
//
// Warp roles (160 threads): warps 0-3 gather operands and run the epilogue (thread t owns TMEM lane t),
// warp 4 allocates TMEM and its lane 0 issues every tcgen05.mma.
#include "geom.hpp"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

using namespace tc;

namespace {

constexpr int kGatherThreads = 128;
constexpr int kTcThreads = 160;
constexpr int kStageA = 128 * 128;      // bytes: 128 rows x 64 bf16 (fprop) or 2 x (64 pixels x 64 bf16) (wgrad)
constexpr int kMaxStages = 8;

struct __align__(8) TcShared {
  unsigned long long full[kMaxStages];
  unsigned long long empty[kMaxStages];
  unsigned long long accum;
  uint32_t tmem_base;
  int tap_dy[kMaxTaps], tap_dx[kMaxTaps];
  float part[4][256];
  float mean[256];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Sum over the 32 rows held by the lanes of a warp of 16 per-thread column values: a transposing
// butterfly (16 shuffles).  On return lanes (2c, 2c+1) ... hold column perm(lane): returns the column
// index this lane ended up owning; the sum is in v[0].
__device__ __forceinline__ int warp_colsum16(float (&v)[16], int lane) {
  // step 1: lanes with bit4 = 0 keep columns [0,8), bit4 = 1 keep [8,16)
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float send = hi ? v[i] : v[i + 8];
      float keep = hi ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float send = hi ? v[i] : v[i + 4];
      float keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      float send = hi ? v[i] : v[i + 2];
      float keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    float send = hi ? v[0] : v[1];
    float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// ------------------------------------------------------------------------------------------------
// forward / data-gradient gather-convolution: persistent, warp-specialised
//   warps 0-3  producers: A gather (cp.async) + B tile (bulk copy), operand ring of S stages
//   warp  4    TMEM allocation; lane 0 issues tcgen05.mma into one of two accumulator buffers
//   warps 5-8  epilogue of tile i while the producers / MMA already work on tile i+1
// ------------------------------------------------------------------------------------------------
constexpr int kGconvThreads = 288;

__device__ __forceinline__ void decode_tile(const GConvParams& p, int t, int& mt, int& v, int& nt) {
  nt = t % p.n_tiles;
  int r = t / p.n_tiles;
  v = r % p.nvar;
  mt = r / p.nvar;
}

__global__ void __launch_bounds__(kGconvThreads) gconv_tc_kernel(const __grid_constant__ GConvParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float part[2][4][256];
  __shared__ float mean_s[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int BN = p.tc_bn, S = p.tc_stages;
  const int stageB = BN * 128;

  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 1024 bytes)
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = ring, b_base = ring + (uint32_t)S * kStageA;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full[s]), kGatherThreads + 1);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull[b]), 1);
      mbar_init(smem_u32(&tempty[b]), 128);
    }
    fence_barrier_init();
  }
  const uint32_t ncols = 2 * BN <= 32 ? 32u : (2 * BN <= 64 ? 64u : (2 * BN <= 128 ? 128u : (2 * BN <= 256 ? 256u : 512u)));
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp < 4) {
    // ---------------- producers: thread -> 16-byte chunk j of rows rg + 16*i ----------------
    const int j = tid & 7, rg = tid >> 3;
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(p.in);
    const uint32_t row_off = (uint32_t)(rg >> 3) * 1024u + (uint32_t)(rg & 7) * 128u + (uint32_t)((j ^ (rg & 7)) << 4);
    const int D = S - 1;
    int g = 0, stage = 0, sig_stage = 0;          // chunks issued; ring slot of chunk g; slot of the next chunk to signal
    uint32_t ephase = 1;                          // parity to wait on empty[stage]: the first lap passes immediately
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      int mt, vi, nt;
      decode_tile(p, t, mt, vi, nt);
      const GVar& var = p.var[vi];
      const int K = var.ntaps * p.Ci;
      const int nchunks = (K + 63) >> 6;
      const int m0 = mt * 128;
      int pix_base[8];      // element offset of input pixel (n, 0, 0), or -1 when the row is beyond M
      int iyx[8];           // (iy0 << 16) | ix0 (already multiplied by the input stride)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        int m = m0 + rg + 16 * i;
        if (m < p.M) {
          int tq, jj, n, ii;
          p.fd_wg.divmod(m, tq, jj);
          p.fd_hg.divmod(tq, n, ii);
          pix_base[i] = n * p.Hi * p.Wi;
          iyx[i] = ((ii * p.is) << 16) | (jj * p.is);
        } else {
          pix_base[i] = -1; iyx[i] = 0;
        }
      }
      const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wpack) + (size_t)vi * p.wpack_var_stride +
                                  (size_t)((nt * BN) >> 3) * 1024;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait(smem_u32(&empty[stage]), ephase);
        if (tid == 0) {
          const uint32_t bar = smem_u32(&full[stage]);
          mbar_arrive_expect_tx(bar, (uint32_t)stageB);
          bulk_g2s(b_base + (uint32_t)stage * stageB, wsrc + (size_t)kc * p.co_pad * 128, (uint32_t)stageB, bar);
        }
        const int k = kc * 64 + j * 8;
        const bool kin = k < K;
        int ci = 0, dy = 0, dx = 0;
        if (kin) { int tap; p.fd_ci.divmod(k, tap, ci); dy = var.dy[tap]; dx = var.dx[tap]; }
        const uint32_t dst = a_base + (uint32_t)stage * kStageA + row_off;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int iy = (iyx[i] >> 16) + dy, ix = (iyx[i] & 0xffff) + dx;
          bool ok = kin && pix_base[i] >= 0 && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
          const __nv_bfloat16* src = ok ? in + ((size_t)(pix_base[i] + iy * p.Wi + ix) * p.Ci + ci) : in;
          cp_async16(dst + (uint32_t)i * 2048u, src, ok ? 16u : 0u);
        }
        cp_async_commit();
        if (++stage == S) { stage = 0; ephase ^= 1u; }
        if (g >= D) {
          cp_async_wait_dyn(D);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full[sig_stage]));
          if (++sig_stage == S) sig_stage = 0;
        }
        ++g;
      }
    }
    // drain: signal the last min(g, D) chunks
    cp_async_wait_dyn(0);
    fence_proxy_async_smem();
    for (int e = (g < D ? g : D); e > 0; --e) {
      mbar_arrive(smem_u32(&full[sig_stage]));
      if (++sig_stage == S) sig_stage = 0;
    }
  } else if (warp == 4) {
    if (lane == 0) {
      // ---------------- MMA issue ----------------
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      int stage = 0;
      uint32_t fphase = 0;
      int i = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
        int mt, vi, nt;
        decode_tile(p, t, mt, vi, nt);
        const int K = p.var[vi].ntaps * p.Ci;
        const int nchunks = (K + 63) >> 6;
        const int buf = i & 1;
        mbar_wait(smem_u32(&tempty[buf]), (uint32_t)(((i >> 1) & 1) ^ 1));   // epilogue drained this buffer
        tc_fence_after();
        const uint32_t dtm = tmem + (uint32_t)(buf * BN);
        for (int kc = 0; kc < nchunks; ++kc) {
          mbar_wait(smem_u32(&full[stage]), fphase);
          tc_fence_after();
          const int kleft = K - kc * 64;
          const int nk = kleft >= 64 ? 4 : (kleft + 15) >> 4;
          const uint32_t sa = a_base + (uint32_t)stage * kStageA, sb = b_base + (uint32_t)stage * stageB;
          for (int q = 0; q < nk; ++q) {
            uint64_t da = make_smem_desc(sa + q * 32, 16, 1024, SWZ_128);
            uint64_t db = make_smem_desc(sb + q * 32, 16, 1024, SWZ_128);
            mma_bf16(dtm, da, db, idesc, (kc | q) != 0);
          }
          mma_commit(smem_u32(&empty[stage]));
          if (++stage == S) { stage = 0; fphase ^= 1u; }
        }
        mma_commit(smem_u32(&tfull[buf]));
      }
    }
  } else {
    // ---------------- epilogue: TMEM -> bf16 NHWC (+ BatchNorm partial statistics) ----------------
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;                         // tile row == TMEM lane
    const int et = (warp - 5) * 32 + lane;               // index among the 128 epilogue threads
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    float run_n = 0.f, run_mean = 0.f, run_m2 = 0.f;     // merged statistics of column `et` over this CTA's tiles
    int i = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x, ++i) {
      int mt, vi, nt;
      decode_tile(p, t, mt, vi, nt);
      const GVar& var = p.var[vi];
      const int m0 = mt * 128, n0 = nt * BN;
      const int buf = i & 1;
      const int m = m0 + r;
      bool valid = m < p.M;
      size_t obase = 0;
      if (valid) {
        int tq, jj, n, ii;
        p.fd_wg.divmod(m, tq, jj);
        p.fd_hg.divmod(tq, n, ii);
        int oy = var.oy0 + p.os * ii, ox = var.ox0 + p.os * jj;
        valid = oy < p.Ho && ox < p.Wo;                  // ragged parity sub-grid of an odd-sized stride-2 dgrad
        obase = ((size_t)(n * p.Ho + oy) * p.Wo + ox) * p.Co;
      }
      mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld16(tlane + (uint32_t)c0, v);
        const int co0 = n0 + c0;
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e) if (co0 + e < p.Co) v[e] += __ldg(p.bias + co0 + e);
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int co = co0 + h * 8;
          if (valid && co < p.Co) {
            uint4* dst = reinterpret_cast<uint4*>(out + obase + co);
            if (p.accumulate) {
              uint4 old = *dst;
              const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(&old);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[h * 8 + e] += __bfloat162float(o[e]);
            }
            uint4 pk;
            pk.x = pack_bf16x2(v[h * 8 + 0], v[h * 8 + 1]); pk.y = pack_bf16x2(v[h * 8 + 2], v[h * 8 + 3]);
            pk.z = pack_bf16x2(v[h * 8 + 4], v[h * 8 + 5]); pk.w = pack_bf16x2(v[h * 8 + 6], v[h * 8 + 7]);
            *dst = pk;
          }
        }
        if (p.partials) {
          // statistics over the values as stored (after rounding), zero for rows / channels outside the tensor
#pragma unroll
          for (int e = 0; e < 16; ++e)
            v[e] = (valid && co0 + e < p.Co) ? __bfloat162float(__float2bfloat16_rn(v[e])) : 0.f;
          int c = warp_colsum16(v, lane);
          if ((lane & 1) == 0) part[0][q][c0 + c] = v[0];
        }
      }
      if (!p.partials) {
        tc_fence_before();
        mbar_arrive(smem_u32(&tempty[buf]));
        continue;
      }
      const int n_valid = min(128, p.M - m0);
      asm volatile("bar.sync 1, 128;" ::: "memory");
      float tile_sum = 0.f;
      if (et < BN) {
        tile_sum = (part[0][0][et] + part[0][1][et]) + (part[0][2][et] + part[0][3][et]);
        mean_s[et] = tile_sum / (float)n_valid;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      // second pass: M2 about the tile mean (no E[y^2] - mean^2 cancellation)
      for (int c0 = 0; c0 < BN; c0 += 16) {
        float v[16];
        tmem_ld16(tlane + (uint32_t)c0, v);
        const int co0 = n0 + c0;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x = v[e];
          if (p.bias && co0 + e < p.Co) x += __ldg(p.bias + co0 + e);
          float d = (valid && co0 + e < p.Co) ? __bfloat162float(__float2bfloat16_rn(x)) - mean_s[c0 + e] : 0.f;
          v[e] = d * d;
        }
        int c = warp_colsum16(v, lane);
        if ((lane & 1) == 0) part[1][q][c0 + c] = v[0];
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[buf]));               // accumulator buffer is free for tile i+2
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < BN) {
        const float tile_m2 = (part[1][0][et] + part[1][1][et]) + (part[1][2][et] + part[1][3][et]);
        if (p.tc_merge) {
          // Chan et al.: merge (n_valid, tile_sum, tile_m2) into the running (n, mean, M2)
          const float nb = (float)n_valid, mb = tile_sum / nb;
          const float nn = run_n + nb, delta = mb - run_mean;
          run_mean += delta * (nb / nn);
          run_m2 += tile_m2 + delta * delta * (run_n * nb / nn);
          run_n = nn;
        } else if (n0 + et < p.Co) {
          const size_t prow = (size_t)vi * p.tiles_m + mt;
          p.partials[(prow * p.Co + n0 + et) * 2 + 0] = tile_sum;
          p.partials[(prow * p.Co + n0 + et) * 2 + 1] = tile_m2;
        }
      }
    }
    if (p.partials && p.tc_merge) {
      if (et < BN && et < p.Co) {
        p.partials[((size_t)blockIdx.x * p.Co + et) * 2 + 0] = run_mean * run_n;
        p.partials[((size_t)blockIdx.x * p.Co + et) * 2 + 1] = run_m2;
      }
      if (et == 0) p.part_counts[blockIdx.x] = run_n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, ncols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads) wgrad_tc_kernel(const __grid_constant__ WGradParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ TcShared sh;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int vi = blockIdx.z / p.nsplit, split = blockIdx.z % p.nsplit;
  const GVar& var = p.var[vi];
  const int BN = p.tc_bn, S = p.tc_stages;
  const int rowB = BN * 2;                     // bytes per pixel row of the dY tile: 32 / 64 / 128
  const int stageB = 64 * rowB;
  const int K = var.ntaps * p.Ci;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int m_lo = split * p.rows_per_split;
  const int m_hi = min(p.M, m_lo + p.rows_per_split);
  const int nchunks = (m_hi - m_lo + 63) >> 6;

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = ring, b_base = ring + (uint32_t)S * kStageA;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&sh.full[s]), kGatherThreads);
      mbar_init(smem_u32(&sh.empty[s]), 1);
    }
    mbar_init(smem_u32(&sh.accum), 1);
    fence_barrier_init();
  }
  const uint32_t ncols = BN <= 32 ? 32u : 64u;
  if (warp == 4) tmem_alloc(smem_u32(&sh.tmem_base), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sh.tmem_base;

  if (k0 >= K || nchunks <= 0) {
    // nothing to do for this (variant, k-tile, split): variants of a stride-2 dgrad-style geometry differ in K
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
    return;
  }

  if (warp < 4) {
    // A^T: thread -> 16-byte chunk jk (8 consecutive (tap,ci) indices) of pixels pg + 8*i
    const int jk = tid & 15, pg = tid >> 4;
    const int k = k0 + jk * 8;
    const bool kin = k < K;
    int ci = 0, dy = 0, dx = 0;
    if (kin) { int tap; p.fd_ci.divmod(k, tap, ci); dy = var.dy[tap]; dx = var.dx[tap]; }
    const uint32_t a_off = (uint32_t)(jk >> 3) * 8192u + (uint32_t)pg * 128u + (uint32_t)(((jk & 7) ^ pg) << 4);
    // dY: 64 pixels x (BN/8) chunks, thread -> transfers e = tid + 128*q
    const int cpr = BN >> 3;                   // chunks per pixel row: 2 / 4 / 8
    const int cpr_log2 = cpr == 2 ? 1 : (cpr == 4 ? 2 : 3);
    const int nb = cpr >> 1;                   // transfers per thread: 1 / 2 / 4
    const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(p.in);
    const __nv_bfloat16* dout = reinterpret_cast<const __nv_bfloat16*>(p.dout);
    const int D = S - 1;
    for (int it = 0; it < nchunks + D; ++it) {
      if (it < nchunks) {
        const int s = it % S;
        if (it >= S) mbar_wait(smem_u32(&sh.empty[s]), (uint32_t)((it / S) - 1) & 1u);
        const int mb = m_lo + it * 64;
        const uint32_t adst = a_base + (uint32_t)s * kStageA + a_off;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = mb + pg + 8 * i;
          if (!kin) continue;                    // rows k >= K of the accumulator are never read: leave the tile as is
          bool ok = m < m_hi;
          const __nv_bfloat16* src = in;
          if (ok) {
            int t, jj, n, ii;
            p.fd_wg.divmod(m, t, jj);
            p.fd_hg.divmod(t, n, ii);
            int iy = ii * p.is + dy, ix = jj * p.is + dx;
            ok = (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
            if (ok) src = in + ((size_t)((n * p.Hi + iy) * p.Wi + ix) * p.Ci + ci);
          }
          cp_async16(adst + (uint32_t)i * 1024u, src, ok ? 16u : 0u);
        }
        const uint32_t bdst = b_base + (uint32_t)s * stageB;
        for (int q = 0; q < nb; ++q) {
          const int e = tid + 128 * q;
          const int px = e >> cpr_log2, c = e & (cpr - 1);
          const int m = mb + px;
          const int co = n0 + c * 8;
          bool ok = m < m_hi && co < p.Co;
          const __nv_bfloat16* src = dout;
          if (ok) {
            int t, jj, n, ii;
            p.fd_wg.divmod(m, t, jj);
            p.fd_hg.divmod(t, n, ii);
            int oy = var.oy0 + p.os * ii, ox = var.ox0 + p.os * jj;
            ok = oy < p.Ho && ox < p.Wo;
            if (ok) src = dout + ((size_t)((n * p.Ho + oy) * p.Wo + ox) * p.Co + co);
          }
          uint32_t a = (uint32_t)px * rowB + (uint32_t)c * 16u;
          a ^= ((a >> 7) & (uint32_t)(cpr - 1)) << 4;
          cp_async16(bdst + a, src, ok ? 16u : 0u);
        }
      }
      cp_async_commit();
      if (it >= D) {
        cp_async_wait_dyn(D);
        fence_proxy_async_smem();
        mbar_arrive(smem_u32(&sh.full[(it - D) % S]));
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
    const uint32_t bswz = BN == 64 ? SWZ_128 : (BN == 32 ? SWZ_64 : SWZ_32);
    for (int kc = 0; kc < nchunks; ++kc) {
      const int s = kc % S;
      mbar_wait(smem_u32(&sh.full[s]), (uint32_t)(kc / S) & 1u);
      tc_fence_after();
      const uint32_t sa = a_base + (uint32_t)s * kStageA, sb = b_base + (uint32_t)s * stageB;
#pragma unroll
      for (int q = 0; q < 4; ++q) {            // 16 pixels per MMA
        uint64_t da = make_smem_desc(sa + q * 2048, 8192, 1024, SWZ_128);
        uint64_t db = make_smem_desc(sb + q * 16 * rowB, 8 * rowB, 8 * rowB, bswz);
        mma_bf16(tmem, da, db, idesc, (kc | q) != 0);
      }
      mma_commit(smem_u32(&sh.empty[s]));
    }
    mma_commit(smem_u32(&sh.accum));
  }

  if (warp < 4) {
    mbar_wait(smem_u32(&sh.accum), 0);
    tc_fence_after();
    const int k = k0 + tid;
    const bool kin = k < K;
    size_t wbase = 0;
    if (kin) { int tap = k / p.Ci; int ci = k - tap * p.Ci; wbase = (size_t)var.wofs[tap] + (size_t)ci * p.w_sci; }
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < BN; c0 += 16) {
      float v[16];
      tmem_ld16(tlane + (uint32_t)c0, v);
      if (kin) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int co = n0 + c0 + i;
          if (co < p.Co) atomicAdd(p.dw + wbase + (size_t)co * p.w_sco, v[i]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, ncols);
  }
}

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const __grid_constant__ PackTable tab, const float* __restrict__ params,
                                                           unsigned char* __restrict__ ws) {
  const PackOp& op = tab.ops[blockIdx.y];
  ConvGeom g{op.kind, op.k, op.s, op.p, op.Ci, op.Co};
  int op_ci, op_co, w_sci, w_sco;
  geom_strides(g, op.dir, op_ci, op_co, w_sci, w_sco);
  const int nvar = geom_nvar(g, op.dir);
  const int co_pad = (op_co + 15) & ~15;
  const int maxchunks = op.maxchunks;
  // one thread per 16-byte destination chunk: (variant, k-chunk, row n, 8 consecutive k)
  const long long per_var = (long long)maxchunks * co_pad * 8;
  const long long total = per_var * nvar;
  const float* w = params + op.w_off;
  unsigned char* dst = ws + (size_t)op.dst_off16 * 16;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += gridDim.x * 256LL) {
    const int v = (int)(e / per_var);
    long long r = e - v * per_var;
    const int kc = (int)(r / (co_pad * 8));
    int rr = (int)(r - (long long)kc * co_pad * 8);
    const int n = rr >> 3, c = rr & 7;
    const int K = geom_ntaps(g, op.dir, v) * op_ci;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    const int kbase = kc * 64 + c * 8;
    if (n < op_co && kbase < K) {
      int tap = kbase / op_ci, ci = kbase - tap * op_ci;
      int dy, dx, wofs;
      geom_tap(g, op.dir, v, tap, dy, dx, wofs);
      const float* src = w + wofs + (size_t)ci * w_sci + (size_t)n * w_sco;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __ldg(src + (size_t)i * w_sci);
      pk[0] = pack_bf16x2(f[0], f[1]); pk[1] = pack_bf16x2(f[2], f[3]);
      pk[2] = pack_bf16x2(f[4], f[5]); pk[3] = pack_bf16x2(f[6], f[7]);
    }
    const size_t off = ((size_t)v * maxchunks + kc) * co_pad * 128 + (size_t)(n >> 3) * 1024 + (size_t)(n & 7) * 128 +
                       (size_t)((c ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(dst + off) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

inline int pow2_floor(int x) { int r = 1; while (r * 2 <= x) r *= 2; return r; }

}  // namespace

bool tc_supported_gconv(const GConvParams& p) {
  return !p.in_nchw_f32 && p.Ci % 8 == 0 && p.Co % 8 == 0 && p.wpack != nullptr && p.Hi < 32768 && p.Wi < 32768;
}
bool tc_supported_wgrad(const WGradParams& p) {
  return !p.in_nchw_f32 && p.Ci % 8 == 0 && p.Co % 8 == 0;
}

StatLayout launch_gconv_tc(const GConvParams& p0, cudaStream_t st) {
  StatLayout sl{0, 0, 0, 0};
  if (p0.M <= 0) return sl;
  GConvParams p = p0;
  const int tiles_m = (p.M + 127) / 128;
  const int co_pad = (p.Co + 15) & ~15;
  // widest channel tile that still leaves >= ~1 CTA per SM; never narrower than 32 unless the layer is
  int bn = co_pad > 128 ? 128 : pow2_floor(co_pad);   // <= 128: one epilogue thread per column for the statistics
  if (co_pad % bn != 0) bn = 16;
  while (bn > 32 && (long long)tiles_m * p.nvar * ((co_pad + bn - 1) / bn) < 128) bn >>= 1;
  int maxchunks = 1;
  for (int v = 0; v < p.nvar; ++v) maxchunks = max(maxchunks, (p.var[v].ntaps * p.Ci + 63) / 64);
  const int stage_bytes = kStageA + bn * 128;
  int stages = min(kMaxStages, max(2, (100 * 1024) / stage_bytes));
  p.tc_bn = bn; p.tc_stages = stages; p.co_pad = co_pad;
  p.tiles_m = tiles_m; p.n_tiles = (co_pad + bn - 1) / bn; p.total_tiles = tiles_m * p.nvar * p.n_tiles;
  p.tc_merge = (p.partials != nullptr && p.n_tiles == 1 && p.part_counts != nullptr) ? 1 : 0;
  p.fd_wg = FastDiv(p.Wg); p.fd_hg = FastDiv(p.Hg); p.fd_ci = FastDiv(p.Ci);
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  const int grid = min(p.total_tiles, 2 * 148);
  count_launch();
  gconv_tc_kernel<<<grid, kGconvThreads, smem, st>>>(p);
  if (p.tc_merge) {
    sl.parts = grid; sl.parts_per_var = grid; sl.tile_rows = 128; sl.rows_per_var = p.M; sl.counts = p.part_counts;
  } else {
    sl.parts = tiles_m * p.nvar; sl.parts_per_var = tiles_m; sl.tile_rows = 128; sl.rows_per_var = p.M;
  }
  return sl;
}

void launch_wgrad_tc(const WGradParams& p0, cudaStream_t st) {
  if (p0.M <= 0) return;
  WGradParams p = p0;
  int maxK = 0;
  for (int i = 0; i < p.nvar; ++i) maxK = max(maxK, p.var[i].ntaps * p.Ci);
  const int co_pad = (p.Co + 15) & ~15;
  const int bn = co_pad >= 64 ? 64 : (co_pad >= 32 ? 32 : 16);
  const int gx = (maxK + 127) / 128, gy = (co_pad + bn - 1) / bn;
  const int base = gx * gy * p.nvar;
  int nsplit = max(1, (2 * 148) / base);
  nsplit = min(nsplit, max(1, (p.M + 255) / 256));      // at least 256 pixels per split
  int rps = (p.M + nsplit - 1) / nsplit;
  rps = (rps + 63) / 64 * 64;
  nsplit = (p.M + rps - 1) / rps;
  p.nsplit = nsplit; p.rows_per_split = rps; p.tc_bn = bn;
  p.fd_wg = FastDiv(p.Wg); p.fd_hg = FastDiv(p.Hg); p.fd_ci = FastDiv(p.Ci);
  const int nchunks = rps / 64;
  const int stage_bytes = kStageA + 64 * bn * 2;
  int stages = min(min(nchunks, kMaxStages), max(2, (100 * 1024) / stage_bytes));
  if (nchunks == 1) stages = 1;
  p.tc_stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  dim3 grid(gx, gy, p.nvar * nsplit);
  count_launch();
  wgrad_tc_kernel<<<grid, kTcThreads, smem, st>>>(p);
}

void launch_pack_weights(const PackTable& tab, const float* params, void* ws, cudaStream_t st) {
  if (tab.n <= 0) return;
  dim3 grid(48, tab.n);
  count_launch();
  pack_weights_kernel<<<grid, 256, 0, st>>>(tab, params, reinterpret_cast<unsigned char*>(ws));
}

}  // namespace mmvae
