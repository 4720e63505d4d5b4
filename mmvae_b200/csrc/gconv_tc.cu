// gconv_tc.cu -- tcgen05 (5th-generation tensor core) implicit-GEMM kernels of the bf16 mode.
//
//   gconv_tc_kernel   forward conv / transposed conv / data gradient ("gather-convolution", kernels.cuh):
//                     D[128 pixels][BN channels] = A[128][K] * W[K][BN].  Persistent and warp-specialised:
//                     one producer thread stages, per (tap, channel block), a TMA box of the NHWC bf16
//                     activation (im2col on the fly: the tap is a coordinate offset, the conv padding is the
//                     TMA out-of-bounds zero fill, a stride-2 conv is a traversal stride) next to the
//                     pre-packed weight tile (bulk copy); one thread issues tcgen05.mma into one of two
//                     TMEM accumulators; four epilogue warps drain the other one (tcgen05.ld -> bf16 NHWC +
//                     per-warp BatchNorm partial sums kept in registers across the CTA's tiles).
//                     Layer shapes whose 128-pixel tile is not a box (odd sizes of the cropped models) use
//                     the same kernel with a cp.async gather by four producer warps instead.
//   wgrad_tc_kernel   weight gradient: D[128 (tap,ci)][BN co] = sum over pixels of A^T * dY; both operands
//                     are MN-major (the pixel axis is the GEMM K axis), TMA-staged the same way, split over
//                     pixel ranges, reduced into the fp32 gradient arena with red.global.add.
//   pack_weights_kernel  fp32 reference-layout weights -> bf16 tiles in the exact shared-memory image
//                     (8-row x 128-byte swizzle atoms) that gconv_tc_kernel bulk-copies.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "bn_fused.cuh"
#include "geom.hpp"
#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

using namespace tc;

namespace {

constexpr int kGatherThreads = 128;
constexpr int kTcThreads = 288;         // 4 producer warps, 1 MMA warp, 4 epilogue warps
constexpr int kStageA = 128 * 128;      // bytes: 128 rows x 64 bf16 (fprop) or 128 (tap,ci) x 64 pixels (wgrad)
constexpr int kMaxStages = 12;
constexpr int kTabCap = 320;            // TMA coordinate table entries (variants x k-chunks x sub-tiles)

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// debugging timeline: slot s of this CTA's row (p.trace == nullptr in production)
#define MMVAE_TRACE(p, s) do { if ((p).trace) (p).trace[(size_t)blockIdx.x * 32 + (s)] = globaltimer_ns(); } while (0)

__device__ __forceinline__ uint32_t swz_code(int row_bytes) { return row_bytes == 128 ? SWZ_128 : (row_bytes == 64 ? SWZ_64 : SWZ_32); }

// ------------------------------------------------------------------------------------------------
// forward / data-gradient gather-convolution
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode_tile(const GConvParams& p, int t, int& mt, int& v, int& nt) {
  int r;
  p.fd_ntiles.divmod(t, r, nt);
  p.fd_nvar.divmod(r, mt, v);
}


// kBwd: data-gradient instantiation whose epilogue also masks the gradient with the consumer's ReLU and reduces the
// consumer's BatchNorm-backward sums (BnBwdFused) -- more live registers, so two CTAs per SM instead of three
// kAct: BatchNorm-free networks (notebook variant): out = act(acc + bias) / out = acc * act'(dact) in the epilogue
template <int kBN, bool kBwd, bool kAct = false>
__global__ void __launch_bounds__(kTcThreads, kBwd ? 2 : 3) gconv_tc_kernel(const __grid_constant__ GConvParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float stat_red[4][kBwd ? 3 : 2][kBN];
  __shared__ uint32_t tma_tab[kTabCap];
  // the warp index through a shuffle: ptxas then knows the role dispatch below is warp-uniform control flow
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;
  if (tid == 0) MMVAE_TRACE(p, 0);
  constexpr int BN = kBN;
  const int S = p.tc_stages;
  const int stageB = BN * 128;
  const int kb = p.tc_kb;                       // channels per A sub-tile: 64 (gather) or min(Ci, 64) (TMA)
  const int kbB = kb * 2;                       // bytes per sub-tile row
  const int sub_bytes = 128 * kbB;

  // 1024-byte aligned operand ring (SWIZZLE_128B atoms are 1024 bytes)
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = ring, b_base = ring + (uint32_t)S * kStageA;

  // TMA coordinate table: entry (variant, k-chunk, sub-tile) = channel offset | dx << 16 | dy << 24 of the box; built by
  // all threads here (kernel parameters only: legal before the PDL wait) so that the producer thread's loop is lean
  const int kb_log2 = p.tc_kb_log2, sub_shift = 6 - kb_log2;
  if (p.use_tma) {
    const int total = (p.nvar * p.tc_maxchunks) << sub_shift;
    for (int e = tid; e < total; e += kTcThreads) {
      const int g = e & ((1 << sub_shift) - 1), r = e >> sub_shift;
      const int vi = r / p.tc_maxchunks, kc = r - vi * p.tc_maxchunks;
      const int k = kc * 64 + (g << kb_log2);
      uint32_t ent = 0;
      if (k < p.var[vi].ntaps * p.Ci) {
        int tap, ci;
        p.fd_ci.divmod(k, tap, ci);
        ent = (uint32_t)ci | ((uint32_t)(unsigned char)p.var[vi].dx[tap] << 16) | ((uint32_t)(unsigned char)p.var[vi].dy[tap] << 24);
      }
      tma_tab[e] = ent;
    }
  }
  // csz > 1: the CTAs of a cluster share one pixel tile (one channel tile each) and multicast their slice of every A box
  // to all of them, so a ring slot is free only when the MMA threads of ALL csz CTAs have released it
  const int csz = p.tc_csz > 1 ? p.tc_csz : 1;
  // Split-K over a thread-block cluster (deep-K layers on few pixel tiles: launch_gconv_tc): the ksplit CTAs of a cluster
  // own ONE tile and a contiguous range of its k-chunks each; the non-leaders hand their fp32 partial accumulators to the
  // leader's shared memory (DSMEM stores), one cluster barrier, and the leader's epilogue adds them to its own.
  const int ksplit = p.tc_ksplit > 1 ? p.tc_ksplit : 1;
  const uint32_t krank = ksplit > 1 ? cluster_ctarank() : 0u;
  const int t_begin = ksplit > 1 ? (int)(blockIdx.x / (unsigned)ksplit) : (int)blockIdx.x;
  const int t_step = ksplit > 1 ? p.total_tiles : (int)gridDim.x;          // one tile per cluster
  if (tid == 0) {
    if (p.use_tma) prefetch_tensormap(&p.tmap_a);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full[s]), p.use_tma ? 1 : kGatherThreads + 1);
      mbar_init(smem_u32(&empty[s]), (uint32_t)csz);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tfull[b]), 1);
      mbar_init(smem_u32(&tempty[b]), 128);
    }
    fence_barrier_init();
  }
  // (k-steps rotating over 2 / 4 accumulators per tile were tried for the deep-K layers: neutral, profiles/r01_issue_loops.md)
  const int accw = 2 * BN;
  const uint32_t ncols = accw <= 32 ? 32u : (accw <= 64 ? 64u : (accw <= 128 ? 128u : (accw <= 256 ? 256u : 512u)));
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  tc_fence_before();
  if (csz > 1) cluster_sync_all();              // peers' barriers are initialised before anything is multicast at them
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint16_t cmask = (uint16_t)((1u << csz) - 1u);
  if (tid == 0) MMVAE_TRACE(p, 1);
  pdl_wait();                                   // everything above overlapped the previous kernel's tail
  pdl_trigger();
  if (tid == 0) MMVAE_TRACE(p, 2);

  if (warp < 4) {
    if (p.use_tma) {
      // ---------------- producers: the four warps, TMA boxes ----------------
      // Every ring slot has ONE owner warp (slot & 3): all four walk the same chunk sequence and each issues the chunks
      // that land in its slots (an owner per slot keeps a waiter at most one phase ahead of its barrier; dealing chunks
      // round-robin regardless of the slot does not).  Whole warps walk the loop (uniform control flow, see elect_one),
      // one elected lane issues.  tc_flags bit 1: single producer (A/B).
      const int nprod = (p.tc_flags & 2) ? 1 : 4;
      if (warp < nprod) {                       // whole warps walk the loop, one elected lane issues (see elect_one)
        const uint64_t tmap = reinterpret_cast<uint64_t>(&p.tmap_a);
        const int pmask = nprod - 1, pw = warp;
        const uint32_t crank = csz > 1 ? cluster_ctarank() : 0u;
        const uint32_t mc_off = crank * (uint32_t)p.tc_mc_bytes;      // this CTA's slice of every A sub-tile
        const int mc_n = (int)crank * p.tc_mc_imgs;
        int stage = 0;
        uint32_t ephase = 1;                    // the first lap over the ring passes immediately
        for (int t = t_begin; t < p.total_tiles; t += t_step) {
          int mt, vi, nt;
          decode_tile(p, t, mt, vi, nt);
          const GVar& var = p.var[vi];
          const int K = var.ntaps * p.Ci;
          const int nchunks = (K + 63) >> 6;
          const int kc0 = ksplit > 1 ? ((int)krank * nchunks) / ksplit : 0;
          const int kc1 = ksplit > 1 ? (((int)krank + 1) * nchunks) / ksplit : nchunks;
          // origin of the tile's pixel box in the gather grid: m0 -> (n0, i0, 0)
          int n0, i0, j0, rem;
          p.fd_hw.divmod(mt * 128, n0, rem);
          p.fd_wg.divmod(rem, i0, j0);
          const size_t wstep = (size_t)p.co_pad * 128;
          const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.wpack) + (size_t)vi * p.wpack_var_stride +
                                      (size_t)((nt * BN) >> 3) * 1024 + (size_t)kc0 * wstep;
          if (pw == 0 && lane == 0 && t == (int)blockIdx.x) MMVAE_TRACE(p, 16);
          const int xb = j0 * p.is, yb = i0 * p.is;
          const uint32_t* tab = tma_tab + ((vi * p.tc_maxchunks) << sub_shift);
          const int nsub_full = 1 << sub_shift, nsub_last = (K - (nchunks - 1) * 64 + kb - 1) >> kb_log2;
          for (int kc = kc0; kc < kc1; ++kc, wsrc += wstep) {
            if ((stage & pmask) == pw) {
              mbar_wait(smem_u32(&empty[stage]), ephase);
              const uint32_t bar = smem_u32(&full[stage]);
              const int nsub = kc + 1 < nchunks ? nsub_full : nsub_last;
              if (elect_one()) {
                mbar_arrive_expect_tx(bar, (uint32_t)(stageB + nsub * sub_bytes));
                bulk_g2s(b_base + (uint32_t)stage * stageB, wsrc, (uint32_t)stageB, bar);
                uint32_t dst = a_base + (uint32_t)stage * kStageA;
                for (int g = 0; g < nsub; ++g, dst += (uint32_t)sub_bytes) {
                  const uint32_t ent = tab[(kc << sub_shift) + g];
                  if (csz > 1)
                    tma_load_4d_mc(dst + mc_off, tmap, bar, (int)(ent & 0xffffu), xb + (int)(signed char)(ent >> 16),
                                   yb + (int)(signed char)(ent >> 24), n0 + mc_n, cmask);
                  else
                    tma_load_4d(dst, tmap, bar, (int)(ent & 0xffffu), xb + (int)(signed char)(ent >> 16), yb + (int)(signed char)(ent >> 24), n0);
                }
              }
              __syncwarp();
            }
            if (++stage == S) { stage = 0; ephase ^= 1u; }
          }
        }
        if (csz > 1) {
          // a peer's MMA thread still arrives on this CTA's `empty` barriers for the last lap: wait for those arrivals
          // (one more lap of waits, nothing issued) so that no CTA of the cluster exits under a remote arrive
          for (int e = 0; e < S; ++e) {
            if ((stage & pmask) == pw) mbar_wait(smem_u32(&empty[stage]), ephase);
            if (++stage == S) { stage = 0; ephase ^= 1u; }
          }
        }
        if (pw == 0 && lane == 0) MMVAE_TRACE(p, 4);
      }
    } else {
      // ---------------- producers: cp.async gather, thread -> 16-byte chunk j of rows rg + 16*i ----------------
      const int j = tid & 7, rg = tid >> 3;
      const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(p.in);
      const uint32_t row_off = (uint32_t)(rg >> 3) * 1024u + (uint32_t)(rg & 7) * 128u + (uint32_t)((j ^ (rg & 7)) << 4);
      const int D = S - 1;
      int g = 0, stage = 0, sig_stage = 0;      // chunks issued; ring slot of chunk g; slot of the next chunk to signal
      uint32_t ephase = 1;
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
    }
  } else if (warp == 4) {
    {
      // ---------------- MMA issue: the whole warp walks the loop, one elected lane issues ----------------
      const uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
      const uint32_t aswz = swz_code(kbB);
      const uint32_t asbo = 8u * (uint32_t)kbB;
      // descriptors of stage 0 / k-step 0; a stage or k-step only moves the (address >> 4) field (no carry: smem < 256 KB)
      const uint64_t da0 = make_smem_desc(a_base, 16, asbo, aswz);
      const uint64_t db0 = make_smem_desc(b_base, 16, 1024, SWZ_128);
      uint32_t aoff[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        aoff[q] = ((uint32_t)((q * 16) >> kb_log2) * (uint32_t)sub_bytes + (uint32_t)((q * 16) & (kb - 1)) * 2u) >> 4;
      // The loop below is ONE warp's serial instruction stream: at ~6 cycles per dependent instruction every instruction
      // in it is ~3 ns per k-chunk, and 100 of them (per-chunk trace stamps, experiment flags, descriptor arithmetic from
      // the stage index) were 0.3 us per chunk -- more than the TMA loads and the MMAs of the chunk together (measured with
      // loads and MMAs switched off).  So: running descriptors / barrier addresses that advance by a constant per stage,
      // nothing per chunk that is not needed to issue it.
      const uint32_t a_step = (uint32_t)kStageA >> 4, b_step = (uint32_t)stageB >> 4;
      const uint32_t full0 = smem_u32(&full[0]), empty0 = smem_u32(&empty[0]);
      uint32_t fbar = full0, ebar = empty0;
      uint64_t da = da0, db = db0;
      int stage = 0;
      uint32_t fphase = 0;
      int i = 0;
      for (int t = t_begin; t < p.total_tiles; t += t_step, ++i) {
        int mt, vi, nt;
        decode_tile(p, t, mt, vi, nt);
        const int K = p.var[vi].ntaps * p.Ci;
        const int nchunks = (K + 63) >> 6;
        const int kc0 = ksplit > 1 ? ((int)krank * nchunks) / ksplit : 0;
        const int kc1 = ksplit > 1 ? (((int)krank + 1) * nchunks) / ksplit : nchunks;
        const int nk_last = (K - (nchunks - 1) * 64 + 15) >> 4;           // MMAs of the last k-chunk (1..4)
        const int buf = i & 1;
        mbar_wait(smem_u32(&tempty[buf]), (uint32_t)(((i >> 1) & 1) ^ 1));   // epilogue drained this buffer
        tc_fence_after();
        const uint32_t dtm = tmem + (uint32_t)(buf * BN);
        for (int kc = kc0; kc < kc1; ++kc) {
          mbar_wait(fbar, fphase);
          tc_fence_after();
          if (elect_one()) {
            if (kc + 1 < nchunks) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                mma_bf16(dtm, da + aoff[q], db + (uint64_t)(q * 2), idesc, ((kc - kc0) | q) != 0);
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (q < nk_last) mma_bf16(dtm, da + aoff[q], db + (uint64_t)(q * 2), idesc, ((kc - kc0) | q) != 0);
            }
            if (csz > 1) mma_commit_mc(ebar, cmask);
            else mma_commit(ebar);
          }
          __syncwarp();
          da += a_step; db += b_step; fbar += 8; ebar += 8;
          if (++stage == S) { stage = 0; fphase ^= 1u; da = da0; db = db0; fbar = full0; ebar = empty0; }
        }
        if (elect_one()) mma_commit(smem_u32(&tfull[buf]));
        __syncwarp();
        if (i == 0 && lane == 0) MMVAE_TRACE(p, 6);
      }
      if (lane == 0) MMVAE_TRACE(p, 7);
    }
  } else {
    // ---------------- epilogue: TMEM -> bf16 NHWC (+ fused BatchNorm statistics) ----------------
    constexpr int NG = kBN / 16;                         // 16-column groups
    constexpr bool kPerThread = kBN == 16;               // statistics kept per thread (row) across tiles: no shuffles per tile
    constexpr int NR = kPerThread ? 16 : NG;
    const int q = warp & 3;                              // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;                         // tile row == TMEM lane
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
    const bool stats = p.bn.acc != nullptr && krank == 0;  // split-K: the leader's epilogue sees the complete sums
    const bool merge = p.tc_merge != 0;                  // one channel tile: sums can live in registers over all tiles
    // split-K partials in the LEADER's shared memory, behind the operand ring: [peer - 1][column][row] fp32
    const uint32_t part_base = ring + (uint32_t)S * (uint32_t)(kStageA + stageB);
    const int lane_col = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
    float run_s[NR], run_q[NR];
#pragma unroll
    for (int e = 0; e < NR; ++e) { run_s[e] = 0.f; run_q[e] = 0.f; }
    float run_b[kBwd ? 3 * NG : 1];                      // kBwd: S0, S1, S2 of this lane's column per 16-column group
#pragma unroll
    for (int e = 0; e < (kBwd ? 3 * NG : 1); ++e) run_b[e] = 0.f;
    const bool bwd2 = kBwd && p.bb.y2 != nullptr;
    int i = 0;
    for (int t = t_begin; t < p.total_tiles; t += t_step, ++i) {
      int mt, vi, nt;
      decode_tile(p, t, mt, vi, nt);
      const GVar& var = p.var[vi];
      const int m0 = mt * 128, n0 = nt * BN;
      const int buf = i & 1;
      const int m = m0 + r;
      if (ksplit > 1 && krank != 0) {
        // non-leader: this CTA's partial accumulator -> the leader's shared memory, then the cluster barrier
        mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((i >> 1) & 1));
        tc_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
        const uint32_t dst0 = dsmem_addr(part_base + (((krank - 1u) * (uint32_t)BN) * 128u + (uint32_t)r) * 4u, 0u);
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
          uint32_t rv[16];
          tmem_ld16_issue(tl + (uint32_t)(gq * 16), rv);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 16; ++e) st_dsmem_u32(dst0 + (uint32_t)((gq * 16 + e) * 128 * 4), rv[e]);
        }
        tc_fence_before();
        cluster_sync_all();
        break;
      }
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
      const bool fuse_v = kBwd && ((p.bb.var_mask >> vi) & 1);       // this tile's pixels are final here: mask + reduce
      mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((i >> 1) & 1));
      tc_fence_after();
      if (ksplit > 1) cluster_sync_all();                // the peers' partials have landed in this CTA's shared memory
      if (i == 0 && tid == 160) MMVAE_TRACE(p, 8);
      const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN);
      uint32_t rnext[16];                                // the next 16-column group is in flight while this one is worked on
      tmem_ld16_issue(tlane, rnext);
#pragma unroll
      for (int gq = 0; gq < NG; ++gq) {
        const int c0 = gq * 16;
        float v[16];
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 16; ++e) v[e] = __uint_as_float(rnext[e]);
        if (gq + 1 < NG) tmem_ld16_issue(tlane + (uint32_t)(c0 + 16), rnext);
        if (ksplit > 1) {
          for (int pe = 0; pe < ksplit - 1; ++pe) {
            const uint32_t src = part_base + (uint32_t)(((pe * BN + c0) * 128 + r) * 4);
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] += ld_smem_f32(src + (uint32_t)(e * 128 * 4));
          }
        }
        const int co0 = n0 + c0;
        if (p.bias) {
#pragma unroll
          for (int e = 0; e < 16; ++e) if (co0 + e < p.Co) v[e] += __ldg(p.bias + co0 + e);
        }
        if constexpr (kAct) {
          if (p.act == ACT_RELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = fmaxf(v[e], 0.f);
          } else if (p.act == ACT_ELU) {
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = v[e] > 0.f ? v[e] : expm1f(v[e]);
          }
        }
        float t1[kBwd ? 16 : 1], t2[kBwd ? 16 : 1];      // kBwd: g * xhat(y), g * xhat(y2)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int co = co0 + h * 8;
          if constexpr (kBwd) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { t1[h * 8 + e] = 0.f; t2[h * 8 + e] = 0.f; }
          }
          if (valid && co < p.Co) {
            uint4* dst = reinterpret_cast<uint4*>(out + obase + co);
            if constexpr (kAct) {
              if (p.dact) {
                const uint4 ar = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dact) + obase + co));
                const __nv_bfloat16* ab = reinterpret_cast<const __nv_bfloat16*>(&ar);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[h * 8 + e] *= act_deriv(p.dact_kind, __bfloat162float(ab[e]));
              }
            }
            if (p.accumulate) {
              uint4 old = *dst;
              const __nv_bfloat16* o = reinterpret_cast<const __nv_bfloat16*>(&old);
#pragma unroll
              for (int e = 0; e < 8; ++e) v[h * 8 + e] += __bfloat162float(o[e]);
            }
            if (kBwd && fuse_v) {
              // g = bf16(dA) * [a > 0]; the masked value is what gets stored and what the sums see
              const uint4 yr = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.y) + obase + co));
              const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(&yr);
              uint4 ar = make_uint4(0u, 0u, 0u, 0u), y2r = make_uint4(0u, 0u, 0u, 0u);
              if (p.bb.a) ar = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.a) + obase + co));
              if (bwd2) y2r = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.bb.y2) + obase + co));
              const __nv_bfloat16* ab = reinterpret_cast<const __nv_bfloat16*>(&ar);
              const __nv_bfloat16* y2b = reinterpret_cast<const __nv_bfloat16*>(&y2r);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                float g = __bfloat162float(__float2bfloat16_rn(v[h * 8 + e]));
                if (p.bb.a && !(__bfloat162float(ab[e]) > 0.f)) g = 0.f;
                v[h * 8 + e] = g;
                const int c = co + e;
                t1[h * 8 + e] = g * ((__bfloat162float(yb[e]) - __ldg(p.bb.stat + c)) * __ldg(p.bb.stat + p.bb.C + c));
                if (bwd2) t2[h * 8 + e] = g * ((__bfloat162float(y2b[e]) - __ldg(p.bb.stat2 + c)) * __ldg(p.bb.stat2 + p.bb.C + c));
              }
            }
            uint4 pk;
            pk.x = pack_bf16x2(v[h * 8 + 0], v[h * 8 + 1]); pk.y = pack_bf16x2(v[h * 8 + 2], v[h * 8 + 3]);
            pk.z = pack_bf16x2(v[h * 8 + 4], v[h * 8 + 5]); pk.w = pack_bf16x2(v[h * 8 + 6], v[h * 8 + 7]);
            *dst = pk;
          }
        }
        if (stats) {
          // statistics over the values as stored (after rounding), zero for rows / channels outside the tensor
#pragma unroll
          for (int e = 0; e < 16; ++e)
            v[e] = (valid && co0 + e < p.Co) ? __bfloat162float(__float2bfloat16_rn(v[e])) : 0.f;
          if (kPerThread && merge) {
#pragma unroll
            for (int e = 0; e < 16; ++e) { run_s[e % NR] += v[e]; run_q[e % NR] = fmaf(v[e], v[e], run_q[e % NR]); }
          } else {
            float sq[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) sq[e] = v[e] * v[e];
            warp_colsum16(v, lane);
            warp_colsum16(sq, lane);
            if (merge) {
              run_s[gq % NR] += v[0]; run_q[gq % NR] += sq[0];
            } else if ((lane & 1) == 0 && co0 + lane_col < p.Co) {
              atomicAdd(bn_acc_copy(p.bn) + co0 + lane_col, (double)v[0]);
              atomicAdd(bn_acc_copy(p.bn) + p.bn.C + co0 + lane_col, (double)sq[0]);
            }
          }
        }
        if constexpr (kBwd) {
          if (fuse_v) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (!(valid && co0 + e < p.Co)) v[e] = 0.f;
            warp_colsum16(v, lane);
            warp_colsum16(t1, lane);
            if (bwd2) warp_colsum16(t2, lane);
            if (merge) {
              run_b[gq * 3 + 0] += v[0]; run_b[gq * 3 + 1] += t1[0]; run_b[gq * 3 + 2] += t2[0];
            } else if ((lane & 1) == 0 && co0 + lane_col < p.Co) {
              double* acc = bn_bwd_acc_copy(p.bb) + co0 + lane_col;
              atomicAdd(acc, (double)v[0]);
              atomicAdd(acc + p.bb.C, (double)t1[0]);
              if (bwd2) atomicAdd(acc + 2 * p.bb.C, (double)t2[0]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tempty[buf]));               // accumulator buffer is free for tile i+2
      if (i == 0 && tid == 160) MMVAE_TRACE(p, 9);
    }
    if (tid == 160) MMVAE_TRACE(p, 10);
    if (stats && merge) {
      // combine the four epilogue warps in shared memory, then one set of atomics per CTA
      if constexpr (kPerThread) {
        warp_colsum16(run_s, lane);                      // NR == 16 here
        warp_colsum16(run_q, lane);
        if ((lane & 1) == 0) { stat_red[q][0][lane_col] = run_s[0]; stat_red[q][1][lane_col] = run_q[0]; }
      } else if ((lane & 1) == 0) {
#pragma unroll
        for (int gq = 0; gq < NG; ++gq) {
          stat_red[q][0][gq * 16 + lane_col] = run_s[gq % NR];
          stat_red[q][1][gq * 16 + lane_col] = run_q[gq % NR];
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int et = (warp - 5) * 32 + lane;
      for (int e = et; e < 2 * kBN; e += 128) {
        const int which = e / kBN, c = e - which * kBN;
        if (c < p.Co) {
          const float t = (stat_red[0][which][c] + stat_red[1][which][c]) + (stat_red[2][which][c] + stat_red[3][which][c]);
          atomicAdd(bn_acc_copy(p.bn) + which * p.bn.C + c, (double)t);
        }
      }
    }
    if constexpr (kBwd) {
      if (p.bb.acc && merge) {
        if ((lane & 1) == 0) {
#pragma unroll
          for (int gq = 0; gq < NG; ++gq)
#pragma unroll
            for (int w3 = 0; w3 < 3; ++w3) stat_red[q][w3][gq * 16 + lane_col] = run_b[gq * 3 + w3];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int et = (warp - 5) * 32 + lane;
        for (int e = et; e < (bwd2 ? 3 : 2) * kBN; e += 128) {
          const int which = e / kBN, c = e - which * kBN;
          if (c < p.Co) {
            const float t = (stat_red[0][which][c] + stat_red[1][which][c]) + (stat_red[2][which][c] + stat_red[3][which][c]);
            atomicAdd(bn_bwd_acc_copy(p.bb) + which * p.bb.C + c, (double)t);
          }
        }
      }
    }
  }
  if (tid == 160) MMVAE_TRACE(p, 11);
  if (ksplit > 1 && warp <= 4) cluster_sync_all();       // producers / MMA warp: their share of the split-K cluster barrier
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, ncols);
  }
  if (tid == 0) MMVAE_TRACE(p, 12);
  if (p.bn.acc) bn_fused_finish(p.bn, gridDim.x);
  if constexpr (kBwd) {
    if (p.bb.acc && p.bb.finish) bn_bwd_fused_finish(p.bb, gridDim.x);
  }
  if (tid == 0) MMVAE_TRACE(p, 13);
}

// ------------------------------------------------------------------------------------------------
// weight gradient
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads) wgrad_tc_kernel(const __grid_constant__ WGradParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const int vi = blockIdx.z / p.nsplit, split = blockIdx.z % p.nsplit;
  const GVar& var = p.var[vi];
  const int BN = p.tc_bn, S = p.tc_stages;
  const int rowB = BN * 2;                     // bytes per pixel row of the dY tile: 32 / 64 / 128
  const int stageB = 64 * rowB;
  const int kb = p.tc_kb, kbB = kb * 2;        // A^T sub-tile: 64 pixels x kb (tap,ci) values
  const int sub_bytes = 64 * kbB;
  const int K = var.ntaps * p.Ci;
  const int k0 = blockIdx.x * 128, n0 = blockIdx.y * BN;
  const int m_lo = split * p.rows_per_split;
  const int m_hi = min(p.M, m_lo + p.rows_per_split);
  const int nchunks = (m_hi - m_lo + 63) >> 6;
  const bool need_gather = !(p.tma_a && p.tma_b);

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = ring, b_base = ring + (uint32_t)S * kStageA;

  // TMA coordinates of this CTA's (tap, channel block) sub-tiles: channel offset | dx << 16 | dy << 24
  __shared__ uint32_t tma_tab[8];
  if (tid < 8) {
    const int k = k0 + tid * kb;
    uint32_t ent = 0;
    if (k < K) {
      int tap, ci;
      p.fd_ci.divmod(k, tap, ci);
      ent = (uint32_t)ci | ((uint32_t)(unsigned char)var.dx[tap] << 16) | ((uint32_t)(unsigned char)var.dy[tap] << 24);
    }
    tma_tab[tid] = ent;
  }
  if (tid == 0) {
    if (p.tma_a) prefetch_tensormap(&p.tmap_a);
    if (p.tma_b) prefetch_tensormap(&p.tmap_b);
    const int arrivals = (need_gather ? kGatherThreads : 0) + ((p.tma_a || p.tma_b) ? 1 : 0);
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&full[s]), arrivals);
      mbar_init(smem_u32(&empty[s]), 1);
    }
    mbar_init(smem_u32(&accum), 1);
    fence_barrier_init();
  }
  const uint32_t ncols = BN <= 32 ? 32u : 64u;
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), ncols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  if (k0 >= K || nchunks <= 0) {
    // nothing to do for this (variant, k-tile, split)
    tc_fence_before();
    __syncthreads();
    if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, ncols); }
    return;
  }

  if (warp < 4) {
    // ---- TMA part: warp 0 walks the loop and one elected lane issues (uniform control flow, see elect_one); when a
    // cp.async gather shares the stage (mixed mode) the lanes of warp 0 are needed there, and thread 0 issues alone ----
    auto tma_loop = [&](auto whole_tag) {
      constexpr bool kWhole = decltype(whole_tag)::value;
      const uint64_t tmap_a = reinterpret_cast<uint64_t>(&p.tmap_a);
      const uint64_t tmap_b = reinterpret_cast<uint64_t>(&p.tmap_b);
      const int nsub = (min(K, k0 + 128) - k0 + kb - 1) / kb;
      const uint32_t bytes = (p.tma_a ? (uint32_t)(nsub * sub_bytes) : 0u) + (p.tma_b ? (uint32_t)stageB : 0u);
      int stage = 0;
      uint32_t ephase = 1;
      for (int it = 0; it < nchunks; ++it) {
        mbar_wait(smem_u32(&empty[stage]), ephase);
        const uint32_t bar = smem_u32(&full[stage]);
        int n0p, i0, j0, rem;
        p.fd_hw.divmod(m_lo + it * 64, n0p, rem);
        p.fd_wg.divmod(rem, i0, j0);
        if (!kWhole || elect_one()) {
          mbar_arrive_expect_tx(bar, bytes);
          if (p.tma_a) {
            const int xb = j0 * p.is, yb = i0 * p.is;
            uint32_t dst = a_base + (uint32_t)stage * kStageA;
            for (int g = 0; g < nsub; ++g, dst += (uint32_t)sub_bytes) {
              const uint32_t ent = tma_tab[g];
              tma_load_4d(dst, tmap_a, bar, (int)(ent & 0xffffu), xb + (int)(signed char)(ent >> 16), yb + (int)(signed char)(ent >> 24), n0p);
            }
          }
          if (p.tma_b)
            tma_load_4d(b_base + (uint32_t)stage * stageB, tmap_b, bar, n0, var.ox0 + p.os * j0, var.oy0 + p.os * i0, n0p);
        }
        if (kWhole) __syncwarp();
        if (++stage == S) { stage = 0; ephase ^= 1u; }
      }
    };
    if (p.tma_a || p.tma_b) {
      if (!need_gather) { if (warp == 0) tma_loop(std::true_type{}); }
      else if (tid == 0) tma_loop(std::false_type{});
    }
    if (need_gather) {
      // ---- cp.async part ----
      // A^T: thread -> 16-byte chunk jk (8 consecutive (tap,ci) indices) of pixels pg + 8*i
      const int jk = tid & 15, pg = tid >> 4;
      const int k = k0 + jk * 8;
      const bool kin = k < K && !p.tma_a;        // rows k >= K of the accumulator are never read: leave them as they are
      int ci = 0, dy = 0, dx = 0;
      if (kin) { int tap; p.fd_ci.divmod(k, tap, ci); dy = var.dy[tap]; dx = var.dx[tap]; }
      const uint32_t a_off = (uint32_t)(jk >> 3) * 8192u + (uint32_t)pg * 128u + (uint32_t)(((jk & 7) ^ pg) << 4);
      // dY: 64 pixels x (BN/8) chunks, thread -> transfers e = tid + 128*q
      const int cpr = BN >> 3;                   // chunks per pixel row: 2 / 4 / 8
      const int cpr_log2 = cpr == 2 ? 1 : (cpr == 4 ? 2 : 3);
      const int nb = p.tma_b ? 0 : (cpr >> 1);   // transfers per thread: 1 / 2 / 4
      const __nv_bfloat16* in = reinterpret_cast<const __nv_bfloat16*>(p.in);
      const __nv_bfloat16* dout = reinterpret_cast<const __nv_bfloat16*>(p.dout);
      const int D = S - 1;
      int stage = 0, sig_stage = 0;
      uint32_t ephase = 1;
      for (int it = 0; it < nchunks; ++it) {
        mbar_wait(smem_u32(&empty[stage]), ephase);
        const int mb = m_lo + it * 64;
        if (kin) {
          const uint32_t adst = a_base + (uint32_t)stage * kStageA + a_off;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int m = mb + pg + 8 * i;
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
        }
        const uint32_t bdst = b_base + (uint32_t)stage * stageB;
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
        cp_async_commit();
        if (++stage == S) { stage = 0; ephase ^= 1u; }
        if (it >= D) {
          cp_async_wait_dyn(D);
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full[sig_stage]));
          if (++sig_stage == S) sig_stage = 0;
        }
      }
      cp_async_wait_dyn(0);
      fence_proxy_async_smem();
      for (int e = (nchunks < D ? nchunks : D); e > 0; --e) {
        mbar_arrive(smem_u32(&full[sig_stage]));
        if (++sig_stage == S) sig_stage = 0;
      }
    }
  } else if (warp == 4) {
    {                                             // whole warp, one elected lane issues (elect_one)
      const uint32_t idesc = make_idesc_bf16(128, BN, 1, 1);
      const uint32_t bswz = swz_code(rowB), aswz = swz_code(kbB);
      // descriptors of stage 0 / pixel step 0; a stage or a 16-pixel step only moves the (address >> 4) field
      const uint64_t da0 = make_smem_desc(a_base, (uint32_t)sub_bytes, 8 * kbB, aswz);
      const uint64_t db0 = make_smem_desc(b_base, 8 * rowB, 8 * rowB, bswz);
      const uint64_t astep = (uint64_t)((16 * kbB) >> 4), bstep = (uint64_t)((16 * rowB) >> 4);
      int stage = 0;
      uint32_t fphase = 0;
      for (int kc = 0; kc < nchunks; ++kc) {
        mbar_wait(smem_u32(&full[stage]), fphase);
        tc_fence_after();
        const uint64_t da = da0 + (uint64_t)(((uint32_t)stage * kStageA) >> 4);
        const uint64_t db = db0 + (uint64_t)(((uint32_t)stage * (uint32_t)stageB) >> 4);
        if (elect_one()) {
#pragma unroll
          for (int q = 0; q < 4; ++q)            // 16 pixels per MMA
            mma_bf16(tmem, da + q * astep, db + q * bstep, idesc, (kc | q) != 0);
          mma_commit(smem_u32(&empty[stage]));
        }
        __syncwarp();
        if (++stage == S) { stage = 0; fphase ^= 1u; }
      }
      if (elect_one()) mma_commit(smem_u32(&accum));
      __syncwarp();
    }
  } else {
    mbar_wait(smem_u32(&accum), 0);
    tc_fence_after();
    const int q = warp & 3;
    const int k = k0 + q * 32 + lane;
    const bool kin = k < K;
    size_t wbase = 0;
    if (kin) { int tap, ci; p.fd_ci.divmod(k, tap, ci); wbase = (size_t)var.wofs[tap] + (size_t)ci * p.w_sci; }
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
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
  pdl_wait();                                       // PDL: may start while the previous kernel drains
  pdl_trigger();
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

// ---------------- host: TMA descriptors ----------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// NHWC bf16 tensor [N][H][W][C]; box = cb channels x (bw x bh x bn) pixels visited with traversal stride `st`
bool make_tmap(TmaDesc& out, const void* base, int N, int H, int W, int C, int cb, int bw, int bh, int bn, int st) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "TmaDesc must mirror CUtensorMap");
  if (cb * 2 != 32 && cb * 2 != 64 && cb * 2 != 128) return false;
  if (bw * st > 256 || bh * st > 256 || bn > 256) return false;
  cuuint64_t dim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)(bw * st), (cuuint32_t)(bh * st), (cuuint32_t)bn};
  cuuint32_t est[4] = {1u, (cuuint32_t)st, (cuuint32_t)st, 1u};
  CUtensorMapSwizzle sw = cb * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (cb * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(&out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dim, str,
                  box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// Is a run of `rows` consecutive gather-grid positions (n, i, j), starting at a multiple of `rows`, a box?
bool pixel_box(int rows, int Hg, int Wg, int& bw, int& bh, int& bn) {
  if (Wg > rows) { if (Wg % rows) return false; bw = rows; bh = 1; bn = 1; return true; }
  if (rows % Wg) return false;
  bw = Wg;
  const int r = rows / Wg;
  if (r <= Hg) { if (Hg % r) return false; bh = r; bn = 1; return true; }
  if (r % Hg) return false;
  bh = Hg; bn = r / Hg;
  return true;
}

// CTA budget of a weight-gradient launch (MMVAE_WGRAD_CTAS env, tuning).  Weight gradients run on the auxiliary stream
// beside the critical path: leaving ~1/5 of the SMs free for the main stream's kernels beats one CTA per SM
// (r01: base step 1.2435 ms at 148, 1.232 +- 0.003 ms at 88 .. 132; the widened and notebook steps do not care.  r02 final,
// with the shorter BatchNorm-backward links: 1.1072 ms at 140, 1.0980 at 120, 1.0915-1.0922 at 64 / 80 / 100: 96).
int wgrad_ctas() {
  static int m = [] { const char* e = getenv("MMVAE_WGRAD_CTAS"); return e ? atoi(e) : 96; }();
  return m;
}
// the notebook variant asks for 120 (WGradParams::cta_budget: 5.67 ms / step against 5.74 at 96) unless the switch is set
bool wgrad_ctas_forced() {
  static bool f = getenv("MMVAE_WGRAD_CTAS") != nullptr;
  return f;
}
int gconv_per_sm() {
  static int m = [] { const char* e = getenv("MMVAE_GCONV_PER_SM"); return e ? atoi(e) : 3; }();
  return m;
}

// Measured neutral to slightly negative on every deep-K layer once the issue loops were lean
// (profiles/r01_issue_loops.md): OFF by default, kept for experiments.
int mc_min_chunks() {                   // TMA multicast over clusters for layers at least this deep (MMVAE_MC_MIN_CHUNKS env; 0 = off)
  static int m = [] { const char* e = getenv("MMVAE_MC_MIN_CHUNKS"); return e ? atoi(e) : 0; }();
  return m;
}

int tma_mask() {                        // bit 0: gconv A, bit 1: wgrad A, bit 2: wgrad dY   (MMVAE_TMA env, debugging)
  static int m = [] { const char* e = getenv("MMVAE_TMA"); return e ? atoi(e) : 7; }();
  return m;
}

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
  if (slab_supported_gconv(p0)) { launch_slab_gconv(p0, st); return sl; }
  GConvParams p = p0;
  const int tiles_m = (p.M + 127) / 128;
  const int co_pad = (p.Co + 15) & ~15;
  // widest channel tile (<= 128) that still leaves >= ~1 CTA per SM; never narrower than 32 unless the layer is
  int bn = co_pad > 128 ? 128 : pow2_floor(co_pad);
  if (co_pad % bn != 0) bn = 16;
  while (bn > 32 && (long long)tiles_m * p.nvar * ((co_pad + bn - 1) / bn) < 128) bn >>= 1;
  { static int fb = [] { const char* e = getenv("MMVAE_FORCE_BN"); return e ? atoi(e) : 0; }(); if (fb && co_pad % fb == 0) bn = fb; }
  const int stage_bytes = kStageA + bn * 128;
  int maxchunks = 1;
  for (int v = 0; v < p.nvar; ++v) maxchunks = max(maxchunks, (p.var[v].ntaps * p.Ci + 63) / 64);
  // shallow-K layers live on many small CTAs per SM (latency hiding by occupancy), deep-K layers on a deep ring
  int stages = maxchunks <= 2 ? 3 : min(8, max(2, (100 * 1024) / stage_bytes));
  // launches with at most one CTA per SM own the whole shared memory: a deep ring hides the L2 latency of deep-K layers
  {
    const int tiles = tiles_m * p.nvar * ((co_pad + bn - 1) / bn);
    if (tiles <= 148 && maxchunks > 2) stages = min(min(kMaxStages, maxchunks * ((tiles + 147) / 148)), max(2, (198 * 1024) / stage_bytes));
  }
  p.tc_bn = bn; p.tc_stages = stages; p.co_pad = co_pad;
  p.tc_maxchunks = maxchunks;
  p.tiles_m = tiles_m; p.n_tiles = (co_pad + bn - 1) / bn; p.total_tiles = tiles_m * p.nvar * p.n_tiles;
  p.tc_merge = p.n_tiles == 1 ? 1 : 0;
  p.fd_wg = FastDiv(p.Wg); p.fd_hg = FastDiv(p.Hg); p.fd_ci = FastDiv(p.Ci);
  p.fd_hw = FastDiv(p.Hg * p.Wg); p.fd_ntiles = FastDiv(p.n_tiles); p.fd_nvar = FastDiv(p.nvar);
  // TMA staging of A: the 128-pixel tile must be a box of the gather grid and the channel block a swizzle width
  p.use_tma = 0; p.tc_kb = 64;
  int bw, bh, bnb;
  const int kb = p.Ci >= 64 ? 64 : p.Ci;
  if ((tma_mask() & 1) && (kb == 16 || kb == 32 || kb == 64) && p.Ci % kb == 0 && pixel_box(128, p.Hg, p.Wg, bw, bh, bnb) &&
      make_tmap(p.tmap_a, p.in, p.N, p.Hi, p.Wi, p.Ci, kb, bw, bh, bnb, p.is)) {
    p.use_tma = 1; p.tc_kb = kb;
  }
  if (p.use_tma && p.nvar * maxchunks * (64 / p.tc_kb) > kTabCap) { p.use_tma = 0; p.tc_kb = 64; }   // coordinate table too small
  p.tc_kb_log2 = p.tc_kb == 64 ? 6 : (p.tc_kb == 32 ? 5 : 4);
  // Deep-K layers on few pixel tiles (the 2x2 .. 4x4 maps in the middle of the network) are bound by the rate at which
  // one SM's TMA unit turns box rows into requests (a 128-row A box per k-chunk, the same box in every channel tile).
  // Their channel tiles form a thread-block cluster: each CTA fetches 1/csz of the box (a run of images) and multicasts
  // it to all of them, so an SM generates 128/csz rows per k-chunk instead of 128.
  p.tc_csz = 1; p.tc_mc_imgs = 0; p.tc_mc_bytes = 0;
  if (p.use_tma && mc_min_chunks() > 0 && maxchunks >= mc_min_chunks() && (p.n_tiles == 2 || p.n_tiles == 4 || p.n_tiles == 8) &&
      bnb % p.n_tiles == 0) {
    const int csz = p.n_tiles;
    if (make_tmap(p.tmap_a, p.in, p.N, p.Hi, p.Wi, p.Ci, p.tc_kb, bw, bh, bnb / csz, p.is)) {
      p.tc_csz = csz; p.tc_mc_imgs = bnb / csz; p.tc_mc_bytes = (128 / csz) * p.tc_kb * 2;
    } else {
      make_tmap(p.tmap_a, p.in, p.N, p.Hi, p.Wi, p.Ci, p.tc_kb, bw, bh, bnb, p.is);
    }
  }
  const bool bwd = p.bb.acc != nullptr;
  // Split-K over a cluster: layers whose tiles do not fill the machine and whose K is deep (layer3/4, uplayer1/2 and their
  // data gradients).  The k-chunks of a tile are dealt to 2 / 4 / 8 CTAs of a cluster; every variant keeps >= 4 chunks per CTA.
  p.tc_ksplit = 1;
  {
    // OFF by default (MMVAE_KSPLIT_MAX=4 turns it on; profiles/r02_split_k.md): the k-loop of encoder.layer4.0.conv2
    // shrinks from 7.4 to 3 us, but the DSMEM hand-over + cluster barrier cost 2.5 us and the fixed parts of the launch
    // (prologue, epilogue, statistics) do not shrink: 10.3 vs 11.3 us on that layer, 0.6-1.6 us SLOWER on the 18-chunk layers.
    static const int ks_max = [] { const char* e = getenv("MMVAE_KSPLIT_MAX"); return e ? atoi(e) : 1; }();
    int minchunks = 1 << 30;
    for (int v = 0; v < p.nvar; ++v) minchunks = min(minchunks, (p.var[v].ntaps * p.Ci + 63) / 64);
    if (p.use_tma && p.tc_csz == 1 && !bwd && !(p.act || p.dact) && p.total_tiles <= 148) {
      int ks = 1;
      while (ks * 2 <= ks_max && p.total_tiles * ks * 2 <= 2 * 148 && minchunks / (ks * 2) >= 4) ks *= 2;
      if (ks > 1) {
        // ring + partials of one CTA must leave room for a second CTA on the SM when the grid exceeds one per SM
        const int per_sm_k = p.total_tiles * ks > 148 ? 2 : 1;
        const int budget = (per_sm_k == 2 ? 110 : 198) * 1024 - (ks - 1) * bn * 128 * 4;
        const int chunks_per_cta = (maxchunks + ks - 1) / ks;
        const int st2 = min(min(kMaxStages, chunks_per_cta), budget / stage_bytes);
        if (st2 >= 2) { p.tc_ksplit = ks; stages = st2; p.tc_stages = stages; }
      }
    }
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024 + (size_t)(p.tc_ksplit - 1) * bn * 128 * 4;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(gconv_tc_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<32, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(gconv_tc_kernel<128, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  const int per_sm = min(gconv_per_sm(), bwd ? 2 : (smem <= 72 * 1024 ? 3 : 2));
  int grid = min(p.total_tiles, per_sm * 148);
  if (p.tc_csz > 1) grid -= grid % p.tc_csz;     // whole clusters; total_tiles is a multiple of n_tiles == csz
  if (p.tc_ksplit > 1) grid = p.total_tiles * p.tc_ksplit;      // one tile per cluster
  p.trace = debug_trace_buffer();
  { static int fl = [] { const char* e = getenv("MMVAE_TC_FLAGS"); return e ? atoi(e) : 0; }(); p.tc_flags = fl; }
  count_launch();
#define LAUNCH_GCONV(...) \
  (p.tc_ksplit > 1 ? launch_pdl_cluster(__VA_ARGS__, grid, kTcThreads, smem, st, p.tc_ksplit, p) : \
   p.tc_csz > 1 ? launch_pdl_cluster(__VA_ARGS__, grid, kTcThreads, smem, st, p.tc_csz, p) : launch_pdl(__VA_ARGS__, grid, kTcThreads, smem, st, p))
  if (p.act || p.dact) {
    switch (bn) {
      case 16: LAUNCH_GCONV(gconv_tc_kernel<16, false, true>); break;
      case 32: LAUNCH_GCONV(gconv_tc_kernel<32, false, true>); break;
      case 64: LAUNCH_GCONV(gconv_tc_kernel<64, false, true>); break;
      default: LAUNCH_GCONV(gconv_tc_kernel<128, false, true>); break;
    }
  } else if (bwd) {
    switch (bn) {
      case 16: LAUNCH_GCONV(gconv_tc_kernel<16, true>); break;
      case 32: LAUNCH_GCONV(gconv_tc_kernel<32, true>); break;
      case 64: LAUNCH_GCONV(gconv_tc_kernel<64, true>); break;
      default: LAUNCH_GCONV(gconv_tc_kernel<128, true>); break;
    }
  } else {
    switch (bn) {
      case 16: LAUNCH_GCONV(gconv_tc_kernel<16, false>); break;
      case 32: LAUNCH_GCONV(gconv_tc_kernel<32, false>); break;
      case 64: LAUNCH_GCONV(gconv_tc_kernel<64, false>); break;
      default: LAUNCH_GCONV(gconv_tc_kernel<128, false>); break;
    }
  }
#undef LAUNCH_GCONV
  return sl;       // statistics are finalised inside the kernel (p.bn); no partial rows
}

void launch_wgrad_tc(const WGradParams& p0, cudaStream_t st) {
  if (p0.M <= 0) return;
  if (slab_supported_wgrad(p0)) { launch_slab_wgrad(p0, st); return; }
  WGradParams p = p0;
  int maxK = 0;
  for (int i = 0; i < p.nvar; ++i) maxK = max(maxK, p.var[i].ntaps * p.Ci);
  const int co_pad = (p.Co + 15) & ~15;
  const int bn = co_pad >= 64 ? 64 : (co_pad >= 32 ? 32 : 16);
  const int gx = (maxK + 127) / 128, gy = (co_pad + bn - 1) / bn;
  const int base = gx * gy * p.nvar;
  int nsplit = max(1, (p.cta_budget > 0 && !wgrad_ctas_forced() ? p.cta_budget : wgrad_ctas()) / base);
  nsplit = min(nsplit, max(1, (p.M + 255) / 256));      // at least 256 pixels per split
  int rps = (p.M + nsplit - 1) / nsplit;
  rps = (rps + 63) / 64 * 64;
  nsplit = (p.M + rps - 1) / rps;
  p.nsplit = nsplit; p.rows_per_split = rps; p.tc_bn = bn;
  p.fd_wg = FastDiv(p.Wg); p.fd_hg = FastDiv(p.Hg); p.fd_ci = FastDiv(p.Ci); p.fd_hw = FastDiv(p.Hg * p.Wg);
  const int nchunks = rps / 64;
  const int stage_bytes = kStageA + 64 * bn * 2;
  int stages = min(min(nchunks, kMaxStages), max(2, (100 * 1024) / stage_bytes));
  if (nchunks == 1) stages = 1;
  p.tc_stages = stages;
  // TMA staging: a 64-pixel chunk must be a box of the gather grid
  p.tma_a = p.tma_b = 0; p.tc_kb = 64;
  int bw, bh, bnb;
  const int kb = p.Ci >= 64 ? 64 : p.Ci;
  if (pixel_box(64, p.Hg, p.Wg, bw, bh, bnb)) {
    if ((tma_mask() & 2) && (kb == 16 || kb == 32 || kb == 64) && p.Ci % kb == 0 &&
        make_tmap(p.tmap_a, p.in, p.N, p.Hi, p.Wi, p.Ci, kb, bw, bh, bnb, p.is)) {
      p.tma_a = 1; p.tc_kb = kb;
    }
    if ((tma_mask() & 4) && p.Co % bn == 0 && make_tmap(p.tmap_b, p.dout, p.N, p.Ho, p.Wo, p.Co, bn, bw, bh, bnb, p.os))
      p.tma_b = 1;
  }
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    attr_done = true;
  }
  dim3 grid(gx, gy, p.nvar * nsplit);
  count_launch();
  launch_pdl(wgrad_tc_kernel, grid, kTcThreads, smem, st, p);
}

void launch_pack_weights(const PackTable& tab, const float* params, void* ws, cudaStream_t st) {
  if (tab.n <= 0) return;
  dim3 grid(48, tab.n);
  count_launch();
  launch_pdl(pack_weights_kernel, grid, 256, 0, st, tab, params, reinterpret_cast<unsigned char*>(ws));
}

}  // namespace mmvae
