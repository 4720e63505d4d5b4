// nb_tail.cu -- the tensor-core-bound end of the notebook variant (vae-kl.ipynb:160,166,226): decoder.conv4, a 3x3 conv
// from 32 channels to the 256 grey-level logits on 128x128 frames (94 % of the network's FLOPs),
// as dedicated tcgen05 kernels instead of the generic gather-convolution of gconv_tc.cu, which refetches the packed
// weights (80 KB) and an im2col copy of the input (72 KB) from L2 for every 128-pixel tile.
//
//   nb_tail_fwd_kernel   out[128 pixels of one image row][256 classes] = sum over the 9 taps of A_tap[128][32] W_tap[32][256]
//     * the whole weight tensor lives in shared memory for the life of the CTA (144 KB bf16, converted from the fp32
//       masters by the CTA itself: no pack launch) as nine K-major SWIZZLE_64B tiles [256 co][32 ci];
//     * an input image row is staged once per band as three dx-shifted copies [128 pixels][32 ci] (TMA boxes starting at
//       x = -1, 0, +1; the halo pixels and rows arrive as TMA out-of-bounds zeros) and serves the three output rows that
//       see it as dy = +1, 0, -1 (un-swizzled planes with the dx as a descriptor shift were tried first: the tensor core
//       fetches un-swizzled operands 16 bytes per row, 2.6 ms; profiles/r01_nb_tail.md).  Three row slots: the oldest
//       row is released after the first six MMAs of a tile, so its successor loads under the other twelve;
//     * a CTA walks bands of 32 consecutive rows of one image; 18 MMAs (M=128, N=256, K=16) per row into one of two
//       256-column TMEM accumulators;
//     * epilogue, logits mode: + bias -> bf16 NHWC.  Cross-entropy mode (training): the 256 logits of a pixel sit in one
//       TMEM lane = one thread, so max / sum-exp / log-likelihood need no shuffles; the thread writes
//       d logits = (softmax - onehot) / N directly: the logits themselves never reach HBM.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "kernels.cuh"
#include "tc_common.cuh"

namespace mmvae {

using namespace tc;

namespace {

constexpr int kW = 128;                     // pixels per image row = GEMM M
constexpr int kCi = 32, kCo = 256;
constexpr int kCopyA = kW * 64;             // bytes of one dx-shifted copy of an input row: 128 pixels x 32 channels, SWIZZLE_64B
constexpr int kSlotA = 3 * kCopyA;          // one input image row: the copies for dx = -1, 0, +1
constexpr int kSlots = 3;
constexpr int kTapW = kCo * 64;             // bytes per tap of the weights: 256 rows (co) x 32 ci, SWIZZLE_64B, K-major
constexpr int kWBytes = 9 * kTapW;          // 147,456
constexpr int kFwdThreads = 320;            // warp 0: TMA producer, warp 1: MMA issue + TMEM owner, warps 2-9: two epilogue groups
constexpr size_t kFwdSmem = (size_t)kWBytes + (size_t)kSlots * kSlotA + 1024;

// MMVAE_NB_TAIL_DBG (tuning experiments, results in profiles/r01_nb_tail.md): 1 = epilogues only follow the barrier protocol,
// 2 = no TMA loads (the MMAs run on whatever is in shared memory)
int nb_tail_dbg() {
  static int v = [] { const char* e = getenv("MMVAE_NB_TAIL_DBG"); return e ? atoi(e) : 0; }();
  return v;
}

// 32-byte store: one full sector per thread and instruction (STG.256)
__device__ __forceinline__ void st_global_v8(void* ptr, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
               "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

// CTA budget of the persistent kernels of this file (148 = one per SM); MMVAE_NB_MAX_CTAS lowers it so that a test with a
// handful of frames still makes every CTA walk several bands.  Read on every launch (tests set and clear it).
int nb_max_ctas() {
  const char* e = getenv("MMVAE_NB_MAX_CTAS");
  const int v = e ? atoi(e) : 148;
  return v >= 1 && v <= 148 ? v : 148;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct NbTailFwd {
  TmaDesc tmap_x;              // upsampled decoder.conv3 activation [N][H][128][32] bf16; box = 32 channels x 128 pixels
  const float* w;              // decoder.conv4.weight [256][32][3][3] fp32
  const float* bias;           // [256]
  __nv_bfloat16* out;          // logits, or d logits in cross-entropy mode: [N][H][128][256]
  const long long* target;     // [N][H][128] class indices: cross-entropy mode; nullptr: logits mode
  double* ce_acc;              // += sum over pixels of (logsumexp - logit[target])
  float scale, inv_scale;      // d logits scale 1 / N, and N
  int H, rows_per_band, bands_per_image, total_bands;
  int dbg;
};

__global__ void __launch_bounds__(kFwdThreads, 1) nb_tail_fwd_kernel(const __grid_constant__ NbTailFwd p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kSlots], empty[kSlots], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[kCo], ebias_s[kCo];   // bias; exp(bias - max bias)
  __shared__ float loss_red[8];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base, a_base = base + kWBytes;
  unsigned char* w_gen = smem_raw + (base - smem_u32(smem_raw));

  // ---- prologue (kernel parameters and weights only: legal before the PDL wait) ----
  if (tid == 0) {
    prefetch_tensormap(&p.tmap_x);
    for (int s = 0; s < kSlots; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  for (int e = tid; e < kCo * kCi * 9; e += kFwdThreads) {
    const int co = e / (kCi * 9), r = e - co * (kCi * 9);
    const int ci = r / 9, t = r - ci * 9;
    *reinterpret_cast<__nv_bfloat16*>(w_gen + (size_t)t * kTapW + swz_off<64>((uint32_t)co, (uint32_t)(ci >> 3)) + (ci & 7) * 2) =
        __float2bfloat16_rn(__ldg(p.w + e));
  }
  if (warp == 2) {
    float bm = -INFINITY;
    for (int c = lane; c < kCo; c += 32) bm = fmaxf(bm, __ldg(p.bias + c));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, o));
    for (int c = lane; c < kCo; c += 32) { const float b = __ldg(p.bias + c); bias_s[c] = b; ebias_s[c] = __expf(b - bm); }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const int R = p.rows_per_band;
  // The single-thread roles below run their loops on the WHOLE warp and only the issuing instructions sit under
  // elect_one() (tc_common.cuh): uniform control flow keeps descriptors and addresses in uniform registers; inside an
  // `if (lane == 0)` every UTCHMMA / UTMALDG operand went through R2UR and an ELECT .. BRA.U.ANY wrapper.
  if (warp == 0) {
    {
      // ---------------- producer: one input row = 4 TMA boxes ----------------
      const uint64_t tmap = reinterpret_cast<uint64_t>(&p.tmap_x);
      int g = 0;                                       // running input-row index of this CTA
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
        const int n = band / p.bands_per_image, y0 = (band - n * p.bands_per_image) * R;
        for (int i = 0; i < R + 2; ++i, ++g) {
          const int slot = g % kSlots;
          mbar_wait(smem_u32(&empty[slot]), (uint32_t)(((g / kSlots) & 1) ^ 1));
          const uint32_t bar = smem_u32(&full[slot]);
          if (elect_one()) {
            if (p.dbg & 2) {
              mbar_arrive(bar);
            } else {
              mbar_arrive_expect_tx(bar, (uint32_t)kSlotA);
              const uint32_t dst = a_base + (uint32_t)slot * kSlotA;
#pragma unroll
              for (int c = 0; c < 3; ++c) tma_load_4d(dst + (uint32_t)c * kCopyA, tmap, bar, 0, c - 1, y0 - 1 + i, n);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    {
      // ---------------- MMA issue ----------------
      const uint32_t idesc = make_idesc_bf16(128, kCo, 0, 0);
      const uint64_t da0 = make_smem_desc(a_base, 16, 512, SWZ_64);
      const uint64_t db0 = make_smem_desc(w_base, 16, 512, SWZ_64);
      int g0 = 0, q = 0;                               // running index of the band's first input row; running output row
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x, g0 += R + 2) {
        for (int o = 0; o < R; ++o, ++q) {
          const int buf = q & 1;
          mbar_wait(smem_u32(&tempty[buf]), (uint32_t)(((q >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t dtm = tmem + (uint32_t)(buf * kCo);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int g = g0 + o + ky;
            const uint32_t slot = (uint32_t)(g % kSlots);
            if (o == 0 || ky == 2) {                   // rows o, o+1 were waited for by the previous output row
              mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / kSlots) & 1));
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t da = da0 + (uint64_t)((slot * kSlotA + (uint32_t)kx * kCopyA + (uint32_t)ks * 32u) >> 4);
                  const uint64_t db = db0 + (uint64_t)(((uint32_t)(ky * 3 + kx) * kTapW + (uint32_t)ks * 32u) >> 4);
                  mma_bf16(dtm, da, db, idesc, (ky | kx | ks) != 0);
                }
              // input row o is dead after its dy = -1 use by output row o: release it now so that row o + 3 loads under
              // the remaining twelve MMAs of this tile and the first twelve of the next
              if (ky == 0 && o < R - 1) mma_commit(smem_u32(&empty[slot]));
            }
            __syncwarp();
          }
          if (elect_one()) {
            mma_commit(smem_u32(&tfull[buf]));
            if (o == R - 1) {                          // end of the band: its last three input rows
              mma_commit(smem_u32(&empty[(g0 + R - 1) % kSlots]));
              mma_commit(smem_u32(&empty[(g0 + R) % kSlots]));
              mma_commit(smem_u32(&empty[(g0 + R + 1) % kSlots]));
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------- epilogue: thread = pixel = TMEM lane; two warp groups take alternate output rows ----------------
    // (one group = one warp per scheduler: TMEM / MUFU latency would be exposed; two groups overlap each other)
    const int lq = warp & 3;                           // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;                   // handles output rows with q % 2 == grp, i.e. TMEM buffer grp
    const int x = lq * 32 + lane;
    const bool ce = p.target != nullptr;
    constexpr float kLog2e = 1.4426950408889634f;
    float loss = 0.f;
    __nv_bfloat16* pend = nullptr;                     // target element of this thread's previous row, still to be fixed up
    int q = 0;
    for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
      const int n = band / p.bands_per_image, y0 = (band - n * p.bands_per_image) * R;
      for (int o = 0; o < R; ++o, ++q) {
        if ((q & 1) != grp) continue;
        const int buf = grp;
        const size_t pix = ((size_t)n * p.H + (y0 + o)) * kW + x;
        __nv_bfloat16* orow = p.out + pix * kCo;
        int tg = 0;
        if (ce) tg = (int)__ldg(p.target + pix);
        mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((q >> 1) & 1));
        tc_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * kCo);
        if (p.dbg & 1) { tc_fence_before(); mbar_arrive(smem_u32(&tempty[buf])); continue; }
        if (!ce) {
#pragma unroll 2
          for (int g = 0; g < 8; ++g) {
            float v[32];
            tmem_ld32(tl + (uint32_t)(g * 32), v);
            uint4 pk[4];
            uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
            for (int e = 0; e < 32; e += 2) pw[e >> 1] = pack_bf16x2(v[e] + bias_s[g * 32 + e], v[e + 1] + bias_s[g * 32 + e + 1]);
#pragma unroll
            for (int h = 0; h < 4; ++h) reinterpret_cast<uint4*>(orow + g * 32)[h] = pk[h];
          }
          tc_fence_before();
          mbar_arrive(smem_u32(&tempty[buf]));
          continue;
        }
        // pass 1: maximum of the raw accumulators; + max bias is an upper bound of the row maximum, which is all the
        // stabilisation needs (the biases span a fraction of a unit)
        float mx = -INFINITY;
#pragma unroll 2
        for (int g = 0; g < 8; ++g) {
          float v[32];
          tmem_ld32(tl + (uint32_t)(g * 32), v);
#pragma unroll
          for (int e = 0; e < 32; ++e) mx = fmaxf(mx, v[e]);
        }
        // pass 2: e_c = exp(v_c - mx) * exp(b_c - bmax) written back over the accumulator (a register copy of the 256
        // values would not fit beside the other epilogue group), and their sum
        float sum0 = 0.f, sum1 = 0.f;
        const float mxl = mx * kLog2e;
#pragma unroll 2
        for (int g = 0; g < 16; ++g) {
          float v[16];
          tmem_ld16(tl + (uint32_t)(g * 16), v);
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            v[e] = ex2_approx(fmaf(v[e], kLog2e, -mxl)) * ebias_s[g * 16 + e];
            v[e + 1] = ex2_approx(fmaf(v[e + 1], kLog2e, -mxl)) * ebias_s[g * 16 + e + 1];
            sum0 += v[e]; sum1 += v[e + 1];
          }
          tmem_st16(tl + (uint32_t)(g * 16), v);
        }
        tmem_st_wait();
        // pass 3: d logits (without the one-hot) = softmax * scale
        const float inv = p.scale / (sum0 + sum1);
        // A thread owns a pixel (= 512-byte row of the output), so a plain store instruction of the warp touches 32 rows,
        // one 32-byte sector each: the store path, not HBM, then bounds the kernel.  Each 64-class block is therefore
        // transposed inside the lane quad first (4 x 4 units of 16 classes = 32 bytes, two shuffle stages): afterwards lane
        // j of the quad holds unit j of all four pixels, and store instruction m writes pixel m of every quad with four
        // adjacent 32-byte pieces: 8 full 128-byte lines per instruction instead of 32 quarter lines.
        const int qj = lane & 3;
        __nv_bfloat16* qrow = orow - (size_t)qj * kCo + qj * 16;          // pixel 0 of my quad, my unit's column offset
#pragma unroll 1
        for (int b = 0; b < 4; ++b) {
          uint32_t u[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float v[16];
            tmem_ld16(tl + (uint32_t)(b * 64 + j * 16), v);
#pragma unroll
            for (int e = 0; e < 16; e += 2) u[j][e >> 1] = pack_bf16x2(v[e] * inv, v[e + 1] * inv);
          }
#pragma unroll
          for (int bit = 0; bit < 2; ++bit) {
            const bool hi = (lane >> bit) & 1;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              if ((r >> bit) & 1) continue;
              const int r2 = r | (1 << bit);
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const uint32_t send = hi ? u[r][e] : u[r2][e];
                const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1 << bit);
                if (hi) u[r][e] = got; else u[r2][e] = got;
              }
            }
          }
#pragma unroll
          for (int m = 0; m < 4; ++m)
            st_global_v8(qrow + (size_t)m * kCo + b * 64, make_uint4(u[m][0], u[m][1], u[m][2], u[m][3]),
                         make_uint4(u[m][4], u[m][5], u[m][6], u[m][7]));
        }
        __syncwarp();                                  // my row was written by the four lanes of my quad: order it before my read-back
        tc_fence_before();
        mbar_arrive(smem_u32(&tempty[buf]));           // the accumulator is free for output row q + 2
        // the target class: read back this thread's own softmax[target] * scale (the per-class select over 256 register
        // values would cost more than the whole pass), take the log-likelihood from it and subtract the one-hot
        // (deferred by one row: by then the row's stores have drained and the load does not wait for them)
        if (pend) {
          const float pt = __bfloat162float(*reinterpret_cast<volatile __nv_bfloat16*>(pend));
          loss -= __logf(fmaxf(pt, 1e-30f) * p.inv_scale);
          *pend = __float2bfloat16_rn(pt - p.scale);
        }
        pend = orow + tg;
      }
    }
    if (pend) {
      const float pt = __bfloat162float(*reinterpret_cast<volatile __nv_bfloat16*>(pend));
      loss -= __logf(fmaxf(pt, 1e-30f) * p.inv_scale);
      *pend = __float2bfloat16_rn(pt - p.scale);
    }
    if (ce) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) loss += __shfl_xor_sync(0xffffffffu, loss, o);
      if (lane == 0) loss_red[warp - 2] = loss;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += loss_red[w];
        atomicAdd(p.ce_acc, (double)t);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradient of decoder.conv4 (+ its bias gradient):
//   dW[co][ci][ky][kx] = sum over pixels of Xu[y + ky - 1][x + kx - 1][ci] * G[y][x][co],   G = d logits
// GEMM with the pixel axis as K, both operands MN-major exactly as TMA delivers the NHWC rows: per image row y and
// per ky one chain of 8 MMAs (M = 128 = the three dx copies of input row y + ky - 1 stacked along M + 32 don't-care rows,
// N = 128 classes, K = 16 pixels) into the ky-th of three TMEM accumulators that live for the CTA's whole life.  A CTA owns
// one half of the classes and every 74th band of rows: G is read from HBM exactly once, the input rows once per class
// half.  The 32 don't-care rows of M read a constant all-ones slice, so rows 96..127 of an accumulator are the sum over
// pixels of G = the bias gradient, at no extra tensor-core work.  Four warps drain the accumulators with red.global.add at
// the end.  (A first version column-summed the G tiles from shared memory in the idle warps: +0.25 ms of bank conflicts.)
// ------------------------------------------------------------------------------------------------
constexpr int kWgSlotsA = 4, kWgSlotsB = 3;
constexpr int kWgSlotA = 4 * kCopyA;        // three dx copies of an input row + a constant all-ones slice: the fourth 32-row block of M
constexpr int kWgSlotB = 2 * kW * 128;      // one row of G for 128 classes: two [128 pixels][64 classes] SWIZZLE_128B tiles
constexpr int kWgThreads = 192;
constexpr size_t kWgSmem = (size_t)kWgSlotsA * kWgSlotA + (size_t)kWgSlotsB * kWgSlotB + 1024;

struct NbTailWgrad {
  TmaDesc tmap_x;              // upsampled input [N][H][128][32] bf16, box = 32 channels x 128 pixels, SWIZZLE_64B
  TmaDesc tmap_g;              // d logits [N][H][128][256] bf16, box = 64 classes x 128 pixels, SWIZZLE_128B
  float* dw;                   // [256][32][3][3] fp32, accumulated with red.global.add (pre-zeroed)
  float* dbias;                // [256] or nullptr
  int H, rows_per_band, bands_per_image, total_bands;
  int dbg;
};

__global__ void __launch_bounds__(kWgThreads, 1) nb_tail_wgrad_kernel(const __grid_constant__ NbTailWgrad p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_a[kWgSlotsA], empty_a[kWgSlotsA], full_b[kWgSlotsB], empty_b[kWgSlotsB], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + kWgSlotsA * kWgSlotA;
  unsigned char* a_gen = smem_raw + (base - smem_u32(smem_raw));
  const int half = blockIdx.x & 1, worker = blockIdx.x >> 1, nworkers = gridDim.x >> 1;

  if (tid == 0) {
    prefetch_tensormap(&p.tmap_x);
    prefetch_tensormap(&p.tmap_g);
    for (int s = 0; s < kWgSlotsA; ++s) { mbar_init(smem_u32(&full_a[s]), 1); mbar_init(smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < kWgSlotsB; ++s) { mbar_init(smem_u32(&full_b[s]), 1); mbar_init(smem_u32(&empty_b[s]), 1); }
    mbar_init(smem_u32(&accum), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  // the fourth slice of every input-row slot is constant 1.0: rows 96..127 of an accumulator become sum over pixels of G, the
  // bias gradient, at no extra tensor-core work (M = 128 either way); TMA only ever writes slices 0..2
  for (int e = tid; e < kWgSlotsA * (kCopyA / 4); e += kWgThreads) {
    const int slot = e / (kCopyA / 4), wd = e - slot * (kCopyA / 4);
    reinterpret_cast<uint32_t*>(a_gen + (size_t)slot * kWgSlotA + 3 * kCopyA)[wd] = 0x3F803F80u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const int R = p.rows_per_band;
  if (warp == 0) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      // ---------------- producer: input rows (three dx copies) and G rows, in the order the MMA thread consumes them ----------------
      const uint64_t tmx = reinterpret_cast<uint64_t>(&p.tmap_x), tmg = reinterpret_cast<uint64_t>(&p.tmap_g);
      int ga = 0, gb = 0;
      for (int band = worker; band < p.total_bands; band += nworkers) {
        const int n = band / p.bands_per_image, y0 = (band - n * p.bands_per_image) * R;
        for (int o = 0; o < R; ++o) {
          for (int i = (o == 0 ? 0 : o + 2); i <= o + 2; ++i, ++ga) {
            const int slot = ga % kWgSlotsA;
            mbar_wait(smem_u32(&empty_a[slot]), (uint32_t)(((ga / kWgSlotsA) & 1) ^ 1));
            const uint32_t bar = smem_u32(&full_a[slot]);
            if (elect_one()) {
              if (p.dbg & 2) {
                mbar_arrive(bar);
              } else {
                mbar_arrive_expect_tx(bar, (uint32_t)(3 * kCopyA));
                const uint32_t dst = a_base + (uint32_t)slot * kWgSlotA;
#pragma unroll
                for (int c = 0; c < 3; ++c) tma_load_4d(dst + (uint32_t)c * kCopyA, tmx, bar, 0, c - 1, y0 - 1 + i, n);
              }
            }
            __syncwarp();
          }
          const int slot = gb % kWgSlotsB;
          mbar_wait(smem_u32(&empty_b[slot]), (uint32_t)(((gb / kWgSlotsB) & 1) ^ 1));
          const uint32_t bar = smem_u32(&full_b[slot]);
          if (elect_one()) {
            if (p.dbg & 2) {
              mbar_arrive(bar);
            } else {
              mbar_arrive_expect_tx(bar, (uint32_t)kWgSlotB);
              const uint32_t dst = b_base + (uint32_t)slot * kWgSlotB;
              tma_load_4d(dst, tmg, bar, half * 128, 0, y0 + o, n);
              tma_load_4d(dst + 16384u, tmg, bar, half * 128 + 64, 0, y0 + o, n);
            }
          }
          __syncwarp();
          ++gb;
        }
      }
    }
  } else if (warp == 1) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      // ---------------- MMA issue ----------------
      const uint32_t idesc = make_idesc_bf16(128, 128, 1, 1);
      const uint64_t da0 = make_smem_desc(a_base, kCopyA, 512, SWZ_64);        // M: 4 x 32 channels, one dx copy apart
      const uint64_t db0 = make_smem_desc(b_base, 16384, 1024, SWZ_128);       // N: 2 x 64 classes
      int g0 = 0, gb = 0;
      bool first = true;
      for (int band = worker; band < p.total_bands; band += nworkers, g0 += R + 2) {
        for (int o = 0; o < R; ++o, ++gb) {
          const uint32_t bslot = (uint32_t)(gb % kWgSlotsB);
          mbar_wait(smem_u32(&full_b[bslot]), (uint32_t)((gb / kWgSlotsB) & 1));
          const uint64_t db = db0 + (uint64_t)((bslot * kWgSlotB) >> 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int g = g0 + o + ky;
            const uint32_t slot = (uint32_t)(g % kWgSlotsA);
            if (o == 0 || ky == 2) mbar_wait(smem_u32(&full_a[slot]), (uint32_t)((g / kWgSlotsA) & 1));
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)((slot * kWgSlotA) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)     // 16 pixels per MMA
                mma_bf16(tmem + (uint32_t)(ky * 128), da + (uint64_t)(ks * 64), db + (uint64_t)(ks * 128), idesc, !(first && ks == 0));
              if (ky == 0 && o < R - 1) mma_commit(smem_u32(&empty_a[slot]));
            }
            __syncwarp();
          }
          first = false;
          if (elect_one()) {
            mma_commit(smem_u32(&empty_b[bslot]));
            if (o == R - 1) {
              mma_commit(smem_u32(&empty_a[(g0 + R - 1) % kWgSlotsA]));
              mma_commit(smem_u32(&empty_a[(g0 + R) % kWgSlotsA]));
              mma_commit(smem_u32(&empty_a[(g0 + R + 1) % kWgSlotsA]));
            }
          }
          __syncwarp();
        }
      }
      if (elect_one()) mma_commit(smem_u32(&accum));
      __syncwarp();
    }
  } else {
    const int gb = (worker < p.total_bands) ? 1 : 0;
    // ---------------- drain the three accumulators: lane m = (kx, ci), column = class ----------------
    if (gb > 0) {
      mbar_wait(smem_u32(&accum), 0);
      tc_fence_after();
      const int lq = warp & 3, m = lq * 32 + lane;
      const int kx = m >> 5, ci = m & 31;
      for (int ky = 0; ky < 3; ++ky) {
        for (int c0 = 0; c0 < 128; c0 += 16) {
          float v[16];
          tmem_ld16(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ky * 128 + c0), v);
          if (m < 96) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
              atomicAdd(p.dw + ((size_t)(half * 128 + c0 + e) * kCi + ci) * 9 + ky * 3 + kx, v[e]);
          } else if (m == 96 && ky == 0 && p.dbias) {          // the all-ones rows: sum over pixels of G = bias gradient
#pragma unroll
            for (int e = 0; e < 16; ++e) atomicAdd(p.dbias + half * 128 + c0 + e, v[e]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// data gradient of decoder.conv4:  dXu[y'][x'][ci] = sum_{ky,kx,co} G[y' - ky + 1][x' - kx + 1][co] * W[co][ci][ky][kx]
// Computed in the transposed ("col2im") form so that the 512-byte G pixel rows are fetched once and never shifted:
//   T_r[x][(ky, kx, ci)] = sum_co G[r][x][co] * W[co][ci][ky][kx]            one GEMM row block per image row r of G:
//   M = 128 pixels, K = 256 classes (four 64-wide SWIZZLE_128B chunks through a 4-stage TMA ring), N = 3 x 96;
// the ky part of the col2im happens inside the tensor core: the ky-th 96-column block of row r accumulates into the TMEM
// accumulator of OUTPUT row r + ky - 1 (four rotating 96-column accumulators), so an output row is complete once G rows
// y'-1, y', y'+1 have passed.  The kx part is a neighbour exchange in the epilogue (thread = pixel: warp shuffles, plus a
// 64-float shared-memory hand-over at the three warp boundaries).  All 288 x 256 transposed weights stay in shared memory
// (144 KB) for the life of the CTA.
// ------------------------------------------------------------------------------------------------
constexpr int kDgChunkW = 96 * 128;         // bytes of one (ky, 64-class chunk) weight tile: 96 rows (kx, ci) x 64 classes
constexpr int kDgWBytes = 3 * 4 * kDgChunkW;   // 147,456
constexpr int kDgStage = kW * 128;          // one 64-class chunk of a G row: 128 pixels x 128 B
constexpr int kDgStages = 4;
constexpr int kDgThreads = 320;            // warp 0: TMA producer, warp 1: MMA issue, warps 2-9: two epilogue groups
constexpr size_t kDgSmem = (size_t)kDgWBytes + (size_t)kDgStages * kDgStage + 1024;

struct NbTailDgrad {
  TmaDesc tmap_g;              // d logits [N][H][128][256] bf16, box = 64 classes x 128 pixels, SWIZZLE_128B
  const float* w;              // [256][32][3][3] fp32
  __nv_bfloat16* dx;           // gradient wrt the upsampled input [N][H][128][32]
  int H, rows_per_band, bands_per_image, total_bands;
  int dbg;
};

__global__ void __launch_bounds__(kDgThreads, 1) nb_tail_dgrad_kernel(const __grid_constant__ NbTailDgrad p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kDgStages], empty[kDgStages], tfull[4], tempty[4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float edge[4][4][2][kCi];                 // [row & 3][pixel quarter][kx = 0 of lane 0 | kx = 2 of lane 31][ci]
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base, a_base = base + kDgWBytes;
  unsigned char* w_gen = smem_raw + (base - smem_u32(smem_raw));

  if (tid == 0) {
    prefetch_tensormap(&p.tmap_g);
    for (int s = 0; s < kDgStages; ++s) { mbar_init(smem_u32(&full[s]), 1); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 4; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), 128); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), 512);
  for (int e = tid; e < kCo * kCi * 9; e += kDgThreads) {
    const int co = e / (kCi * 9), r = e - co * (kCi * 9);
    const int ci = r / 9, t = r - ci * 9;
    const int ky = t / 3, kx = t - ky * 3;
    *reinterpret_cast<__nv_bfloat16*>(w_gen + (size_t)(ky * 4 + (co >> 6)) * kDgChunkW +
                                      swz_off<128>((uint32_t)(kx * 32 + ci), (uint32_t)((co & 63) >> 3)) + (co & 7) * 2) =
        __float2bfloat16_rn(__ldg(p.w + e));
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const int R = p.rows_per_band;
  if (warp == 0) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      // ---------------- producer: G rows y0-1 .. y0+R, four 64-class chunks each ----------------
      const uint64_t tmg = reinterpret_cast<uint64_t>(&p.tmap_g);
      int g = 0;
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
        const int n = band / p.bands_per_image, y0 = (band - n * p.bands_per_image) * R;
        for (int rr = -1; rr <= R; ++rr)
          for (int kc = 0; kc < 4; ++kc, ++g) {
            const int stage = g % kDgStages;
            mbar_wait(smem_u32(&empty[stage]), (uint32_t)(((g / kDgStages) & 1) ^ 1));
            const uint32_t bar = smem_u32(&full[stage]);
            if (elect_one()) {
              if (p.dbg & 2) {
                mbar_arrive(bar);
              } else {
                mbar_arrive_expect_tx(bar, (uint32_t)kDgStage);
                tma_load_4d(a_base + (uint32_t)stage * kDgStage, tmg, bar, kc * 64, 0, y0 + rr, n);
              }
            }
            __syncwarp();
          }
      }
    }
  } else if (warp == 1) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      // ---------------- MMA issue ----------------
      const uint32_t idesc = make_idesc_bf16(128, 96, 0, 0);
      const uint64_t da0 = make_smem_desc(a_base, 16, 1024, SWZ_128);
      const uint64_t db0 = make_smem_desc(w_base, 16, 1024, SWZ_128);
      int g = 0, qb = 0;                               // running chunk index; running index of the band's first output row
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x, qb += R) {
        for (int rr = -1; rr <= R; ++rr) {
          // a fresh accumulator (output row rr + 1, first touched by ky = 2) must have been drained by the epilogue
          if (rr + 1 < R) {
            const int q = qb + rr + 1;
            mbar_wait(smem_u32(&tempty[q & 3]), (uint32_t)(((q >> 2) & 1) ^ 1));
          }
          for (int kc = 0; kc < 4; ++kc, ++g) {
            const int stage = g % kDgStages;
            mbar_wait(smem_u32(&full[stage]), (uint32_t)((g / kDgStages) & 1));
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)(((uint32_t)stage * kDgStage) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const int oo = rr + ky - 1;            // band-local output row fed by the ky-th block
                if (oo < 0 || oo >= R) continue;
                const uint32_t dtm = tmem + (uint32_t)(((qb + oo) & 3) * 128);
                const uint64_t db = db0 + (uint64_t)(((uint32_t)(ky * 4 + kc) * kDgChunkW) >> 4);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                  mma_bf16(dtm, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, !(ky == 2 && kc == 0 && ks == 0));
              }
              mma_commit(smem_u32(&empty[stage]));
            }
            __syncwarp();
          }
          if (rr >= 1) {
            if (elect_one()) mma_commit(smem_u32(&tfull[(qb + rr - 1) & 3]));      // output row rr - 1 is complete
            __syncwarp();
          }
        }
      }
    }
  } else {
    // ---------------- epilogue: thread = pixel; kx neighbour exchange, bf16 NHWC store.  Two warp groups take alternate output rows ----------------
    const int lq = warp & 3, egrp = (warp - 2) >> 2;
    const int x = lq * 32 + lane;
    int q = 0;
    for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
      const int n = band / p.bands_per_image, y0 = (band - n * p.bands_per_image) * R;
      for (int oo = 0; oo < R; ++oo, ++q) {
        if ((q & 1) != egrp) continue;
        const int slot = q & 3;
        mbar_wait(smem_u32(&tfull[slot]), (uint32_t)((q >> 2) & 1));
        tc_fence_after();
        const uint32_t tl = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(slot * 128);
        if (p.dbg & 1) { tc_fence_before(); mbar_arrive(smem_u32(&tempty[slot])); continue; }
        float t0[32], t1[32], t2[32];                  // kx = 0, 1, 2 blocks of this pixel
        {
          uint32_t r0[32], r1[32], r2[32];
          tmem_ld32_issue(tl, r0);
          tmem_ld32_issue(tl + 32u, r1);
          tmem_ld32_issue(tl + 64u, r2);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e) { t0[e] = __uint_as_float(r0[e]); t1[e] = __uint_as_float(r1[e]); t2[e] = __uint_as_float(r2[e]); }
        }
        tc_fence_before();
        mbar_arrive(smem_u32(&tempty[slot]));
        float (*eb)[2][kCi] = edge[q & 3];         // a group's consecutive rows (q, q + 2) use different buffers: one barrier per row
        if (lane == 0) {
#pragma unroll
          for (int e = 0; e < 32; ++e) eb[lq][0][e] = t0[e];
        }
        if (lane == 31) {
#pragma unroll
          for (int e = 0; e < 32; ++e) eb[lq][1][e] = t2[e];
        }
        if (egrp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
        // out[x] = T[x+1][kx=0] + T[x][kx=1] + T[x-1][kx=2]
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float up = __shfl_down_sync(0xffffffffu, t0[e], 1);      // from pixel x + 1
          float dn = __shfl_up_sync(0xffffffffu, t2[e], 1);        // from pixel x - 1
          if (lane == 31) up = lq < 3 ? eb[lq + 1][0][e] : 0.f;
          if (lane == 0) dn = lq > 0 ? eb[lq - 1][1][e] : 0.f;
          t1[e] += up + dn;
        }
        __nv_bfloat16* orow = p.dx + (((size_t)n * p.H + (y0 + oo)) * kW + x) * kCi;
#pragma unroll
        for (int h = 0; h < 4; h += 2) {
          uint4 pk[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            pk[u].x = pack_bf16x2(t1[(h + u) * 8 + 0], t1[(h + u) * 8 + 1]); pk[u].y = pack_bf16x2(t1[(h + u) * 8 + 2], t1[(h + u) * 8 + 3]);
            pk[u].z = pack_bf16x2(t1[(h + u) * 8 + 4], t1[(h + u) * 8 + 5]); pk[u].w = pack_bf16x2(t1[(h + u) * 8 + 6], t1[(h + u) * 8 + 7]);
          }
          st_global_v8(orow + h * 8, pk[0], pk[1]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// The 3x3 / stride 1 / 32 -> 32 channel convolutions of the decoder (decoder.conv2 on 32x32, decoder.conv3 on 64x64 maps)
// and their data gradients: same row-band scheme as nb_tail_fwd_kernel (resident weights, every input row staged once per
// band as three dx-shifted copies, 18 MMAs per 128-pixel tile), N = 32.  Through the generic im2col kernel these layers
// refetch 72 KB of input + 20 KB of weights per tile from L2 and run at the L2 rate (0.3 ms each at 512 frames of 64x64:
// 8x their HBM bound).  Image rows are 64 or 32 pixels wide, so one 128-pixel M tile is the SAME row of 2 or 4 consecutive
// images: one TMA box {32 ch, W px, 1 row, 128/W images} lands contiguously, no pairing of different rows needed.
//   forward:        out = act(conv(in) + bias)                 Wk[(ky,kx)][ci][co]       = w[co][ci][ky][kx]
//   data gradient:  out = conv with the flipped, transposed w   Wk[(2-ky,2-kx)][co][ci]  = w[co][ci][ky][kx]   (no bias, no act)
// ------------------------------------------------------------------------------------------------
constexpr int kMidSlots = 6;
constexpr int kMidWBytes = 9 * 32 * 64;     // 18,432
constexpr int kMidThreads = 288;            // warps 0-3: cp.async producers, warp 4: MMA issue + TMEM owner, warps 5-8: epilogue
constexpr size_t kMidSmem = (size_t)kMidWBytes + (size_t)kMidSlots * kSlotA + 1024 + 1024;

struct NbMid {
  const __nv_bfloat16* in;     // input [N][H][W][32] bf16
  const float* w;              // [32][32][3][3] fp32 (the conv's own weight tensor)
  const float* bias;           // [32] or nullptr
  __nv_bfloat16* out;          // [N][H][W][32]
  int act;                     // ACT_*
  int dgrad;                   // weights flipped + transposed
  int H, W, imgs, rows_per_band, bands_per_group, total_bands;
};

__global__ void __launch_bounds__(kMidThreads, 1) nb_mid_kernel(const __grid_constant__ NbMid p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full[kMidSlots], empty[kMidSlots], tfull[2], tempty[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[32];
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t w_base = base, a_base = base + kMidWBytes + 1024 - (kMidWBytes & 1023);
  unsigned char* w_gen = smem_raw + (base - smem_u32(smem_raw));

  if (tid == 0) {
    for (int s = 0; s < kMidSlots; ++s) { mbar_init(smem_u32(&full[s]), 128); mbar_init(smem_u32(&empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&tfull[b]), 1); mbar_init(smem_u32(&tempty[b]), 128); }
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 64);
  for (int e = tid; e < 32 * 32 * 9; e += kMidThreads) {
    const int co = e / 288, r = e - co * 288;
    const int ci = r / 9, t = r - ci * 9;
    // B tile of tap t': rows = output channel o, 32 K values = input channel k
    const int tp = p.dgrad ? 8 - t : t, o = p.dgrad ? ci : co, k = p.dgrad ? co : ci;
    *reinterpret_cast<__nv_bfloat16*>(w_gen + (size_t)tp * 2048 + swz_off<64>((uint32_t)o, (uint32_t)(k >> 3)) + (k & 7) * 2) =
        __float2bfloat16_rn(__ldg(p.w + e));
  }
  if (tid < 32) bias_s[tid] = p.bias ? __ldg(p.bias + tid) : 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const int R = p.rows_per_band;
  if (warp < 4) {
    // ---------------- producers: 128 threads stage one input row as three dx-shifted SWIZZLE_64B copies with cp.async ----------------
    // (three TMA boxes of 128 x 64-byte rows per tile keep the TMA unit busy ~1.5 us: the row rate, not the bytes, limits
    // it, and with N = 32 the tensor core needs 0.3 us per tile; the 1536 16-byte cp.async of a row are 12 per thread, the
    // second and third copy hit L1)
    const int chunk = tid & 3, rbase = tid >> 2;       // 16-byte chunk of a pixel; tile rows rbase + 32 r
    constexpr int D = 2;                               // rows in flight before the oldest is published
    int g = 0, gs = 0;                                 // rows issued / rows signalled
    for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
      const int grp = band / p.bands_per_group, y0 = (band - grp * p.bands_per_group) * R;
      for (int i = 0; i < R + 2; ++i, ++g) {
        const int slot = g % kMidSlots;
        mbar_wait(smem_u32(&empty[slot]), (uint32_t)(((g / kMidSlots) & 1) ^ 1));
        const int y = y0 - 1 + i;
        const bool yok = (unsigned)y < (unsigned)p.H;
        const uint32_t dst = a_base + (uint32_t)slot * kSlotA;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = rbase + 32 * r;
          const int img = row / p.W, px = row - img * p.W;
          const __nv_bfloat16* src = p.in + ((((size_t)grp * p.imgs + img) * p.H + (yok ? y : 0)) * p.W) * 32 + chunk * 8;
          const uint32_t d = dst + swz_off<64>((uint32_t)row, (uint32_t)chunk);
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int xs = px + c - 1;
            const bool ok = yok && (unsigned)xs < (unsigned)p.W;
            cp_async16(d + (uint32_t)c * kCopyA, ok ? (const void*)(src + (size_t)xs * 32) : (const void*)p.in, ok ? 16u : 0u);
          }
        }
        cp_async_commit();
        if (g - gs >= D) {
          cp_async_wait<D>();
          fence_proxy_async_smem();
          mbar_arrive(smem_u32(&full[gs % kMidSlots]));
          ++gs;
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async_smem();
    for (; gs < g; ++gs) mbar_arrive(smem_u32(&full[gs % kMidSlots]));
  } else if (warp == 4) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      const uint32_t idesc = make_idesc_bf16(128, 32, 0, 0);
      const uint64_t da0 = make_smem_desc(a_base, 16, 512, SWZ_64);
      const uint64_t db0 = make_smem_desc(w_base, 16, 512, SWZ_64);
      int g0 = 0, q = 0;
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x, g0 += R + 2) {
        for (int o = 0; o < R; ++o, ++q) {
          const int buf = q & 1;
          mbar_wait(smem_u32(&tempty[buf]), (uint32_t)(((q >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t dtm = tmem + (uint32_t)(buf * 32);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int g = g0 + o + ky;
            const uint32_t slot = (uint32_t)(g % kMidSlots);
            if (o == 0 || ky == 2) {
              mbar_wait(smem_u32(&full[slot]), (uint32_t)((g / kMidSlots) & 1));
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t da = da0 + (uint64_t)((slot * kSlotA + (uint32_t)kx * kCopyA + (uint32_t)ks * 32u) >> 4);
                  const uint64_t db = db0 + (uint64_t)(((uint32_t)(ky * 3 + kx) * 2048u + (uint32_t)ks * 32u) >> 4);
                  mma_bf16(dtm, da, db, idesc, (ky | kx | ks) != 0);
                }
              if (ky == 0 && o < R - 1) mma_commit(smem_u32(&empty[slot]));
            }
            __syncwarp();
          }
          if (elect_one()) {
            mma_commit(smem_u32(&tfull[buf]));
            if (o == R - 1) {
              mma_commit(smem_u32(&empty[(g0 + R - 1) % kMidSlots]));
              mma_commit(smem_u32(&empty[(g0 + R) % kMidSlots]));
              mma_commit(smem_u32(&empty[(g0 + R + 1) % kMidSlots]));
            }
          }
          __syncwarp();
        }
      }
    }
  } else {
    // epilogue: thread = pixel of the tile = (image i, column px) of one row
    const int lq = warp & 3;
    const int x = lq * 32 + lane;
    const int img = x / p.W, px = x - img * p.W;
    const bool elu = p.act == ACT_ELU, relu = p.act == ACT_RELU;
    int q = 0;
    for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
      const int grp = band / p.bands_per_group, y0 = (band - grp * p.bands_per_group) * R;
      for (int o = 0; o < R; ++o, ++q) {
        const int buf = q & 1;
        mbar_wait(smem_u32(&tfull[buf]), (uint32_t)((q >> 1) & 1));
        tc_fence_after();
        float v[32];
        tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(buf * 32), v);
        tc_fence_before();
        mbar_arrive(smem_u32(&tempty[buf]));
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          float t = v[e] + bias_s[e];
          if (elu) t = t > 0.f ? t : ex2_approx(t * 1.4426950408889634f) - 1.0f;   // bf16 output: the approximate exp2 is exact enough
          else if (relu) t = fmaxf(t, 0.f);
          v[e] = t;
        }
        uint4 pk[4];
        uint32_t* pw = reinterpret_cast<uint32_t*>(pk);
#pragma unroll
        for (int e = 0; e < 32; e += 2) pw[e >> 1] = pack_bf16x2(v[e], v[e + 1]);
        __nv_bfloat16* orow = p.out + ((((size_t)grp * p.imgs + img) * p.H + (y0 + o)) * p.W + px) * 32;
        st_global_v8(orow, pk[0], pk[1]);
        st_global_v8(orow + 16, pk[2], pk[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

// ------------------------------------------------------------------------------------------------
// Weight + bias gradient of the same 32 -> 32 3x3 convs (decoder.conv2 / conv3): nb_tail_wgrad_kernel's scheme with N = 32.
// Pixel axis = K; per tile (one row of 2 / 4 images = 128 pixels) and ky one chain of 8 MMAs (M = 128 = three dx copies of
// input row y + ky - 1 + the all-ones slice, N = 32, K = 16 pixels) into three 32-column TMEM accumulators that live for the
// CTA's whole life; input copies and the dY row are staged with cp.async by 128 threads.  Through the generic
// wgrad_tc_kernel + column-sum kernel these two layers took 0.73 ms of the auxiliary stream at 512 frames.
// ------------------------------------------------------------------------------------------------
constexpr int kMwSlotsA = 4, kMwSlotsB = 4;
constexpr int kMwSlotB = kW * 64;           // one dY row: 128 pixels x 32 channels, SWIZZLE_64B
constexpr int kMwThreads = 288;             // warps 0-3: cp.async producers, warp 4: MMA issue + TMEM owner, warps 5-8: final drain
constexpr size_t kMwSmem = (size_t)kMwSlotsA * kWgSlotA + (size_t)kMwSlotsB * kMwSlotB + 1024;

struct NbMidWgrad {
  const __nv_bfloat16* in;     // layer input [N][H][W][32]
  const __nv_bfloat16* dy;     // dY [N][H][W][32]
  float* dw;                   // [32][32][3][3], accumulated with red.global.add (pre-zeroed)
  float* dbias;                // [32] or nullptr
  int H, W, imgs, rows_per_band, bands_per_group, total_bands;
};

__global__ void __launch_bounds__(kMwThreads, 1) nb_mid_wgrad_kernel(const __grid_constant__ NbMidWgrad p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  __shared__ __align__(8) unsigned long long full_a[kMwSlotsA], empty_a[kMwSlotsA], full_b[kMwSlotsB], empty_b[kMwSlotsB], accum;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = tid & 31;   // warp-uniform role dispatch for ptxas
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t a_base = base, b_base = base + kMwSlotsA * kWgSlotA;
  unsigned char* a_gen = smem_raw + (base - smem_u32(smem_raw));

  if (tid == 0) {
    for (int s = 0; s < kMwSlotsA; ++s) { mbar_init(smem_u32(&full_a[s]), 128); mbar_init(smem_u32(&empty_a[s]), 1); }
    for (int s = 0; s < kMwSlotsB; ++s) { mbar_init(smem_u32(&full_b[s]), 128); mbar_init(smem_u32(&empty_b[s]), 1); }
    mbar_init(smem_u32(&accum), 1);
    fence_barrier_init();
  }
  if (warp == 4) tmem_alloc(smem_u32(&tmem_base_s), 128);
  for (int e = tid; e < kMwSlotsA * (kCopyA / 4); e += kMwThreads) {          // the constant all-ones fourth slice of every slot
    const int slot = e / (kCopyA / 4), wd = e - slot * (kCopyA / 4);
    reinterpret_cast<uint32_t*>(a_gen + (size_t)slot * kWgSlotA + 3 * kCopyA)[wd] = 0x3F803F80u;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  pdl_wait();
  pdl_trigger();

  const int R = p.rows_per_band;
  if (warp < 4) {
    // ---------------- producers: per output row o the input row o + 2 (rows 0, 1, 2 at o = 0) and dY row o; one cp.async
    // group per staged row, published two groups later ----------------
    const int chunk = tid & 3, rbase = tid >> 2;
    constexpr int D = 2;
    int ga = 0, gb = 0;                                // input rows / dY rows issued
    uint32_t pend[D + 1];                              // barriers of the groups in flight, oldest first
    int npend = 0;
    auto publish_oldest = [&]() {
      fence_proxy_async_smem();
      mbar_arrive(pend[0]);
#pragma unroll
      for (int i = 0; i < D; ++i) pend[i] = pend[i + 1];
      --npend;
    };
    auto commit = [&](uint32_t bar) {
      cp_async_commit();
      pend[npend++] = bar;
      if (npend > D) { cp_async_wait<D>(); publish_oldest(); }
    };
    // Before blocking on a slot that is still in use, publish everything staged so far: at a band boundary the MMA thread
    // needs the last dY row of the old band (still among the unpublished groups) to release the slots the new band waits for.
    auto wait_free = [&](uint32_t bar, uint32_t parity) {
      if (!mbar_try_wait(bar, parity)) {
        cp_async_wait<0>();
        while (npend > 0) publish_oldest();
        mbar_wait(bar, parity);
      }
    };
    for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x) {
      const int grp = band / p.bands_per_group, y0 = (band - grp * p.bands_per_group) * R;
      for (int o = 0; o < R; ++o) {
        for (int i = (o == 0 ? 0 : o + 2); i <= o + 2; ++i, ++ga) {
          const int slot = ga % kMwSlotsA;
          wait_free(smem_u32(&empty_a[slot]), (uint32_t)(((ga / kMwSlotsA) & 1) ^ 1));
          const int y = y0 - 1 + i;
          const bool yok = (unsigned)y < (unsigned)p.H;
          const uint32_t dst = a_base + (uint32_t)slot * kWgSlotA;
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int row = rbase + 32 * r;
            const int img = row / p.W, px = row - img * p.W;
            const __nv_bfloat16* src = p.in + ((((size_t)grp * p.imgs + img) * p.H + (yok ? y : 0)) * p.W) * 32 + chunk * 8;
            const uint32_t d = dst + swz_off<64>((uint32_t)row, (uint32_t)chunk);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int xs = px + c - 1;
              const bool ok = yok && (unsigned)xs < (unsigned)p.W;
              cp_async16(d + (uint32_t)c * kCopyA, ok ? (const void*)(src + (size_t)xs * 32) : (const void*)p.in, ok ? 16u : 0u);
            }
          }
          commit(smem_u32(&full_a[slot]));
        }
        const int slot = gb % kMwSlotsB;
        wait_free(smem_u32(&empty_b[slot]), (uint32_t)(((gb / kMwSlotsB) & 1) ^ 1));
        const uint32_t dst = b_base + (uint32_t)slot * kMwSlotB;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int row = rbase + 32 * r;
          const int img = row / p.W, px = row - img * p.W;
          const __nv_bfloat16* src = p.dy + ((((size_t)grp * p.imgs + img) * p.H + (y0 + o)) * p.W + px) * 32 + chunk * 8;
          cp_async16(dst + swz_off<64>((uint32_t)row, (uint32_t)chunk), src, 16u);
        }
        commit(smem_u32(&full_b[slot]));
        ++gb;
      }
    }
    cp_async_wait<0>();
    while (npend > 0) publish_oldest();
  } else if (warp == 4) {
    {                                                  // whole warp, one elected lane issues (elect_one)
      const uint32_t idesc = make_idesc_bf16(128, 32, 1, 1);
      const uint64_t da0 = make_smem_desc(a_base, kCopyA, 512, SWZ_64);        // M: 4 x 32 channels, one slice apart
      const uint64_t db0 = make_smem_desc(b_base, 512, 512, SWZ_64);           // N: 32 channels = one atom
      int g0 = 0, gb = 0;
      bool first = true;
      for (int band = blockIdx.x; band < p.total_bands; band += gridDim.x, g0 += R + 2) {
        for (int o = 0; o < R; ++o, ++gb) {
          const uint32_t bslot = (uint32_t)(gb % kMwSlotsB);
          const uint64_t db = db0 + (uint64_t)((bslot * kMwSlotB) >> 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int g = g0 + o + ky;
            const uint32_t slot = (uint32_t)(g % kMwSlotsA);
            if (o == 0 || ky == 2) mbar_wait(smem_u32(&full_a[slot]), (uint32_t)((g / kMwSlotsA) & 1));
            // the dY row is staged after input row o + 2: wait for it before the first chain that can run (ky = 0 needs rows
            // that arrived earlier, but the chain also reads dY)
            if (ky == 0) mbar_wait(smem_u32(&full_b[bslot]), (uint32_t)((gb / kMwSlotsB) & 1));
            tc_fence_after();
            const uint64_t da = da0 + (uint64_t)((slot * kWgSlotA) >> 4);
            if (elect_one()) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                mma_bf16(tmem + (uint32_t)(ky * 32), da + (uint64_t)(ks * 64), db + (uint64_t)(ks * 64), idesc, !(first && ks == 0));
              if (ky == 0 && o < R - 1) mma_commit(smem_u32(&empty_a[slot]));
            }
            __syncwarp();
          }
          first = false;
          if (elect_one()) {
            mma_commit(smem_u32(&empty_b[bslot]));
            if (o == R - 1) {
              mma_commit(smem_u32(&empty_a[(g0 + R - 1) % kMwSlotsA]));
              mma_commit(smem_u32(&empty_a[(g0 + R) % kMwSlotsA]));
              mma_commit(smem_u32(&empty_a[(g0 + R + 1) % kMwSlotsA]));
            }
          }
          __syncwarp();
        }
      }
      if (elect_one()) mma_commit(smem_u32(&accum));
      __syncwarp();
    }
  } else if ((int)blockIdx.x < p.total_bands) {
    // ---------------- drain: lane m = (kx, ci) or the all-ones rows, column = co ----------------
    mbar_wait(smem_u32(&accum), 0);
    tc_fence_after();
    const int lq = warp & 3, m = lq * 32 + lane;
    const int kx = m >> 5, ci = m & 31;
    for (int ky = 0; ky < 3; ++ky) {
      float v[32];
      tmem_ld32(tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ky * 32), v);
      if (m < 96) {
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(p.dw + ((size_t)e * 32 + ci) * 9 + ky * 3 + kx, v[e]);
      } else if (m == 96 && ky == 0 && p.dbias) {
#pragma unroll
        for (int e = 0; e < 32; ++e) atomicAdd(p.dbias + e, v[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem, 128);
  }
}

// ---------------- host: TMA descriptor without swizzle ----------------
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
// NHWC bf16 [N][H][W][C]; box = cb channels x bw pixels of one row
bool make_tmap_rows(TmaDesc& out, const void* base, int N, int H, int W, int C, int cb, int bw, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  static_assert(sizeof(CUtensorMap) == sizeof(TmaDesc), "TmaDesc must mirror CUtensorMap");
  cuuint64_t dim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t str[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)cb, (cuuint32_t)bw, 1u, 1u};
  cuuint32_t est[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(reinterpret_cast<CUtensorMap*>(&out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dim, str,
                  box, est, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

bool nb_tail_supported(int Ci, int Co, int H, int W, int k, int s, int pad) {
  static const bool off = getenv("MMVAE_NO_NB_TAIL") != nullptr;
  return !off && Ci == kCi && Co == kCo && W == kW && H % 32 == 0 && k == 3 && s == 1 && pad == 1;
}

bool launch_nb_tail_fwd(const NbTailArgs& a, cudaStream_t st) {
  NbTailFwd p;
  memset(&p, 0, sizeof(p));
  if (!make_tmap_rows(p.tmap_x, a.x, a.N, a.H, kW, kCi, kCi, kW, CU_TENSOR_MAP_SWIZZLE_64B)) return false;
  p.w = a.w; p.bias = a.bias; p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.target = a.target; p.ce_acc = a.ce_acc; p.scale = a.scale; p.inv_scale = a.scale > 0.f ? 1.0f / a.scale : 0.f;
  p.H = a.H; p.rows_per_band = 32; p.bands_per_image = a.H / 32; p.total_bands = a.N * p.bands_per_image;
  p.dbg = nb_tail_dbg();
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nb_tail_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFwdSmem);
    attr_done = true;
  }
  const int grid = p.total_bands < nb_max_ctas() ? p.total_bands : nb_max_ctas();
  count_launch();
  launch_pdl(nb_tail_fwd_kernel, dim3(grid), dim3(kFwdThreads), kFwdSmem, st, p);
  return true;
}

bool nb_mid_supported(int Ci, int Co, int H, int W, int N, int k, int s, int pad) {
  static const bool off = getenv("MMVAE_NO_NB_MID") != nullptr;
  if (off || Ci != 32 || Co != 32 || k != 3 || s != 1 || pad != 1 || H != W) return false;
  if (W != 64 && W != 32) return false;
  return N % (128 / W) == 0 && H % 32 == 0;
}

bool launch_nb_mid(const NbMidArgs& a, cudaStream_t st) {
  NbMid p;
  memset(&p, 0, sizeof(p));
  const int imgs = 128 / a.W;
  p.in = reinterpret_cast<const __nv_bfloat16*>(a.in);
  p.w = a.w; p.bias = a.bias; p.out = reinterpret_cast<__nv_bfloat16*>(a.out); p.act = a.act; p.dgrad = a.dgrad;
  p.H = a.H; p.W = a.W; p.imgs = imgs; p.rows_per_band = 32; p.bands_per_group = a.H / 32;
  p.total_bands = (a.N / imgs) * p.bands_per_group;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nb_mid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMidSmem);
    attr_done = true;
  }
  const int grid = p.total_bands < nb_max_ctas() ? p.total_bands : nb_max_ctas();
  count_launch();
  launch_pdl(nb_mid_kernel, dim3(grid), dim3(kMidThreads), kMidSmem, st, p);
  return true;
}

bool launch_nb_mid_wgrad(const NbMidArgs& a, const void* dy, float* dw, float* dbias, cudaStream_t st) {
  NbMidWgrad p;
  memset(&p, 0, sizeof(p));
  const int imgs = 128 / a.W;
  p.in = reinterpret_cast<const __nv_bfloat16*>(a.in); p.dy = reinterpret_cast<const __nv_bfloat16*>(dy);
  p.dw = dw; p.dbias = dbias;
  p.H = a.H; p.W = a.W; p.imgs = imgs; p.rows_per_band = 32; p.bands_per_group = a.H / 32;
  p.total_bands = (a.N / imgs) * p.bands_per_group;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nb_mid_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMwSmem);
    attr_done = true;
  }
  const int grid = p.total_bands < nb_max_ctas() ? p.total_bands : nb_max_ctas();
  count_launch();
  launch_pdl(nb_mid_wgrad_kernel, dim3(grid), dim3(kMwThreads), kMwSmem, st, p);
  return true;
}

bool launch_nb_tail_dgrad(const NbTailArgs& a, void* dx, cudaStream_t st) {
  NbTailDgrad p;
  memset(&p, 0, sizeof(p));
  if (!make_tmap_rows(p.tmap_g, a.out, a.N, a.H, kW, kCo, 64, kW, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
  p.w = a.w; p.dx = reinterpret_cast<__nv_bfloat16*>(dx);
  p.H = a.H; p.rows_per_band = 32; p.bands_per_image = a.H / 32; p.total_bands = a.N * p.bands_per_image;
  p.dbg = nb_tail_dbg();
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nb_tail_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kDgSmem);
    attr_done = true;
  }
  const int grid = p.total_bands < nb_max_ctas() ? p.total_bands : nb_max_ctas();
  count_launch();
  launch_pdl(nb_tail_dgrad_kernel, dim3(grid), dim3(kDgThreads), kDgSmem, st, p);
  return true;
}

bool launch_nb_tail_wgrad(const NbTailArgs& a, float* dw, float* dbias, cudaStream_t st) {
  NbTailWgrad p;
  memset(&p, 0, sizeof(p));
  if (!make_tmap_rows(p.tmap_x, a.x, a.N, a.H, kW, kCi, kCi, kW, CU_TENSOR_MAP_SWIZZLE_64B)) return false;
  if (!make_tmap_rows(p.tmap_g, a.out, a.N, a.H, kW, kCo, 64, kW, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
  p.dw = dw; p.dbias = dbias;
  p.H = a.H; p.rows_per_band = 32; p.bands_per_image = a.H / 32; p.total_bands = a.N * p.bands_per_image;
  p.dbg = nb_tail_dbg();
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(nb_tail_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWgSmem);
    attr_done = true;
  }
  const int wcap = nb_max_ctas() >= 2 ? nb_max_ctas() / 2 : 1;
  const int workers = p.total_bands < wcap ? p.total_bands : wcap;
  count_launch();
  launch_pdl(nb_tail_wgrad_kernel, dim3(2 * workers), dim3(kWgThreads), kWgSmem, st, p);
  return true;
}

}  // namespace mmvae
