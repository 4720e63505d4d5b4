// kernels.cuh -- launch descriptors and launcher prototypes shared by the .cu files.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>

namespace mmvae {

// kernels launched by this library in this process (bench.py reports the delta over its timed region)
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- programmatic dependent launch (PDL) ----
// A kernel launched through launch_pdl() may be scheduled while its stream predecessor is still running; it must
// execute pdl_wait() before it touches global memory, and calls pdl_trigger() once its own dependents may start
// their prologue (barrier init, TMEM allocation, descriptor prefetch).  Under stream capture the edge becomes a
// programmatic dependency of the CUDA graph.  MMVAE_NO_PDL=1 turns the attribute off (plain stream order).
bool pdl_enabled();
unsigned long long* debug_trace_buffer();   // device buffer set by mmvae_debug_set_trace, or nullptr
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// launch_pdl over thread-block clusters of `csz` consecutive CTAs (grid.x % csz == 0)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int csz,
                                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csz; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Division by a runtime constant without the integer-divide sequence (valid for numerators < 2^31).
struct FastDiv {
  unsigned int mul = 0, shr = 0, div = 1;
  FastDiv() = default;
  explicit FastDiv(int d) {
    div = (unsigned int)d;
    if (d > 1) {
      int lg = 0;
      while ((1u << lg) < (unsigned int)d) ++lg;           // ceil(log2(d))
      const unsigned long long pw = 1ull << (31 + lg);
      mul = (unsigned int)((pw + (unsigned long long)d - 1) / (unsigned long long)d);
      shr = (unsigned int)(lg - 1);
    }
  }
#ifdef __CUDACC__
  __device__ __forceinline__ int quo(int x) const { return div == 1 ? x : (int)(__umulhi((unsigned int)x, mul) >> shr); }
  __device__ __forceinline__ void divmod(int x, int& q, int& r) const { q = quo(x); r = x - q * (int)div; }
#endif
};

// Training-mode BatchNorm statistics folded into the producing kernel (bn_fused.cuh).  acc == nullptr: off.
constexpr int kBnAccCopies = 8;
struct BnFused {
  double* acc;                 // [kBnAccCopies][2][C] sum, sum of squares -- zeroed before the kernel
  unsigned int* counter;       // CTAs finished -- zeroed before the kernel
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* nbt;   // nullptr: do not update
  float* stat; float* coef;    // outputs [2][C] each: (mean, rstd), (scale, shift)
  int C;
  double inv_m, unbias;        // 1/m and m/(m-1) for the m = N*H*W values per channel
};

// BatchNorm-backward reduction folded into the kernel that produces the incoming gradient dA (the data-gradient
// epilogue of the tcgen05 conv kernel): g = bf16(dA) * [a > 0] is what gets stored, and per channel
//   S0 = sum g, S1 = sum g * xhat(y), S2 = sum g * xhat(y2)
// go to fp64 accumulators; the CTA that finishes last writes the backward coefficients.  acc == nullptr: off.
struct BnBwdFused {
  const void* a;               // activation for the ReLU mask (same shape as dA), bf16
  const void* y;  const float* stat;  const float* gamma;     // main branch: raw conv output, (mean, rstd)
  const void* y2; const float* stat2; const float* gamma2;    // second branch or nullptr
  double* acc;                 // [kBnAccCopies][3][C], zeroed
  unsigned int* counter;       // CTAs finished, zeroed
  float* bcoef; float* bcoef2; // [5][C]: scale, S0/m, S1/m, S0, S1 (the last two become d beta, d gamma)
  int C;
  int var_mask;                // parity variants of this launch whose pixels get masked + reduced here (the others are
                               // completed, masked and reduced by a later launch that accumulates onto them)
  int finish;                  // this launch is the last contributor: its last CTA writes bcoef
  double inv_rows;
};

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_ELU = 2 };
#ifdef __CUDACC__
__device__ __forceinline__ float act_apply(int kind, float v) {
  if (kind == ACT_RELU) return fmaxf(v, 0.f);
  if (kind == ACT_ELU) return v > 0.f ? v : expm1f(v);
  return v;
}
// derivative of the activation expressed through its OUTPUT a: relu' = [a > 0]; elu' = a > 0 ? 1 : a + 1 (= e^x)
__device__ __forceinline__ float act_deriv(int kind, float a) {
  if (kind == ACT_RELU) return a > 0.f ? 1.f : 0.f;
  if (kind == ACT_ELU) return a > 0.f ? 1.f : a + 1.f;
  return 1.f;
}
#endif

// Opaque copy of a CUtensorMap (TMA descriptor); lives inside the __grid_constant__ kernel parameter.
struct alignas(64) TmaDesc { unsigned long long v[16]; };

constexpr int kMaxTaps = 25;   // 5x5 stem conv
constexpr int kMaxVar = 4;     // output-parity variants of a stride-2 transposed conv / dgrad

// One output-parity variant of a gather-convolution.
struct GVar {
  int oy0, ox0;                // out y = oy0 + os * i
  int ntaps;
  int wofs[kMaxTaps];          // weight offset of the tap inside the reference-layout weight tensor
  signed char dy[kMaxTaps];    // in y = is * i + dy
  signed char dx[kMaxTaps];
};

// Gather-convolution: every conv / transposed conv / dgrad of the model is an instance.
//   out[n, oy0+os*i, ox0+os*j, co] (+)= bias[co] + sum_t sum_ci in[n, is*i+dy_t, is*j+dx_t, ci] * W[wofs_t + ci*w_sci + co*w_sco]
// for i < Hg, j < Wg; out-of-range input coordinates read as zero.
// GEMM view: M = N*Hg*Wg rows, Co columns, K = ntaps*Ci.
struct GConvParams {
  const void* in;              // NHWC [N,Hi,Wi,Ci], storage type -- or fp32 NCHW when in_nchw_f32
  void* out;                   // NHWC [N,Ho,Wo,Co], storage type
  const float* w;              // fp32 weight tensor in the reference layout
  const float* bias;           // fp32 [Co] or nullptr
  float* partials;             // fp32 [nvar*gridM][Co][2] per-CTA (sum, M2 about the tile mean) or nullptr
  int N, Hi, Wi, Ci, Ho, Wo, Co;
  int Hg, Wg, M;
  int os, is;
  int w_sci, w_sco;
  int nvar;
  int accumulate;              // out += result
  int in_nchw_f32;             // network input x: fp32 NCHW
  int conv_class;              // 1: ConvTranspose2d k4 s2 p1 forward, 2: its data gradient (slab_tc.cu candidates), else 0
  // BatchNorm-free networks (the notebook variant, nb.cu): activation fused into the epilogue, out = act(acc + bias),
  // and, for a data gradient, the derivative of the PRODUCER's activation: out = acc * act'(dact[same index])
  int act;                     // ACT_NONE / ACT_RELU / ACT_ELU
  int dact_kind;               // activation whose output `dact` is
  const void* dact;            // stored activation output (same shape / type as out) or nullptr
  // tcgen05 path: pre-packed bf16 weight tiles [variant][k-chunk of 64][co_pad rows][128 B, swizzled]
  const void* wpack;           // nullptr: no packed weights (SIMT only)
  int wpack_var_stride;        // bytes between variants
  int co_pad;                  // Co rounded up to 16
  float* part_counts;          // rows behind each `partials` row when the kernel merges its tiles (or nullptr)
  BnFused bn;                  // tcgen05 / stem / tail kernels: fused statistics (bn.acc != nullptr)
  BnBwdFused bb;               // tcgen05 data gradient: fused BatchNorm-backward reduction of the consumer (bb.acc != nullptr)
  // set by the launcher
  int tc_bn, tc_stages, tc_merge;
  int tc_kb, use_tma;          // channels per A sub-tile; A staged by TMA boxes (else cp.async gather)
  int tc_kb_log2, tc_maxchunks; // log2(tc_kb); k-chunks (of 64) of the deepest variant
  int tc_flags;                // experiments (MMVAE_TC_FLAGS env)
  // TMA multicast over a thread-block cluster: the tc_csz CTAs of a cluster hold the tc_csz channel tiles of ONE pixel
  // tile; each fetches 1/tc_csz of the A box (tc_mc_imgs images = tc_mc_bytes bytes of every sub-tile) for all of them
  int tc_csz, tc_mc_imgs, tc_mc_bytes;
  int tc_ksplit;               // split-K: CTAs of a cluster sharing one tile (1 = off)
  unsigned long long* trace;   // debugging: [CTA][16] globaltimer stamps (mmvae_debug_set_trace), nullptr = off
  int tiles_m, n_tiles, total_tiles;
  FastDiv fd_wg, fd_hg, fd_ci, fd_hw, fd_ntiles, fd_nvar;
  GVar var[kMaxVar];
  TmaDesc tmap_a;              // NHWC activation, box = tc_kb channels x 128 pixels
};

// Weight gradient of a gather-convolution: dW[wofs_t + ci*w_sci + co*w_sco] += sum_m in(m,t,ci) * dout(m,co)
struct WGradParams {
  const void* in;              // layer input activation (or fp32 NCHW x)
  const void* dout;            // dY, NHWC [N,Ho,Wo,Co], storage type
  float* dw;                   // gradient arena + weight offset (fp32, atomically accumulated; pre-zeroed)
  int N, Hi, Wi, Ci, Ho, Wo, Co;
  int Hg, Wg, M;
  int os, is;
  int w_sci, w_sco;
  int nvar, nsplit, rows_per_split;
  int in_nchw_f32;
  int conv_class;              // 1: ConvTranspose2d k4 s2 p1 (slab_tc.cu candidate), else 0
  int tc_bn, tc_stages;        // set by the launcher (tcgen05 path)
  int tc_kb, tma_a, tma_b;
  int cta_budget;              // CTAs of this weight-gradient launch (0: the library default, wgrad_ctas())
  FastDiv fd_wg, fd_hg, fd_ci, fd_hw;
  GVar var[kMaxVar];
  TmaDesc tmap_a, tmap_b;      // layer input (box = tc_kb channels x 64 pixels) / dY (box = tc_bn channels x 64 pixels)
};

// Layout of the per-CTA partial statistics a conv kernel wrote: `parts` rows of [Co][2] = (sum, M2);
// row i covers GEMM rows [(i % parts_per_var) * tile_rows, +tile_rows) of its variant, clipped to rows_per_var.
// counts != nullptr: row i covers counts[i] rows instead (a persistent kernel merged its tiles).
// sumsq: column 1 of a row is the plain sum of squares instead of M2.
struct StatLayout { int parts, parts_per_var, tile_rows, rows_per_var; const float* counts = nullptr; int sumsq = 0; };

template <typename T> StatLayout launch_gconv_simt(const GConvParams& p, cudaStream_t st);
template <typename T> void launch_wgrad_simt(const WGradParams& p, cudaStream_t st);

// ---- tcgen05 kernels (gconv_tc.cu), bf16 storage only ----
bool tc_supported_gconv(const GConvParams& p);
bool tc_supported_wgrad(const WGradParams& p);
StatLayout launch_gconv_tc(const GConvParams& p, cudaStream_t st);
// shared-memory-resident band kernels for the narrow, large-image transposed convolutions (slab_tc.cu)
bool slab_supported_gconv(const GConvParams& p);
bool slab_supported_wgrad(const WGradParams& p);
void launch_slab_wgrad(const WGradParams& p, cudaStream_t st);
void launch_slab_gconv(const GConvParams& p, cudaStream_t st);
void launch_wgrad_tc(const WGradParams& p, cudaStream_t st);

// One entry per (conv, direction) whose weights are packed for the tcgen05 path.
struct PackOp {
  long long w_off;             // floats into the parameter arena
  unsigned int dst_off16;      // 16-byte units into the workspace
  unsigned short Ci, Co;       // the conv's own channels
  unsigned short maxchunks;    // k-chunks (of 64) per variant
  unsigned char kind, k, s, p, dir;
};
constexpr int kMaxPackOps = 72;
struct PackTable { int n; PackOp ops[kMaxPackOps]; };
void launch_pack_weights(const PackTable& tab, const float* params, void* ws, cudaStream_t st);

// ---- 1-channel stem / tail convolutions (special.cu), bf16 storage only ----
struct StemArgs {
  const float* x;              // [N,1,S,S] fp32 NCHW network input
  const float* w;              // [Co][1][5][5] fp32
  __nv_bfloat16* y;            // forward: [N,Ho,Wo,Co]
  float* partials;             // forward: [tiles][Co][2] or nullptr
  BnFused bn;                  // forward: fused statistics when bn.acc != nullptr (then `partials` is unused)
  const __nv_bfloat16* dy;     // wgrad: dY [N,Ho,Wo,Co]
  // wgrad with the stem BatchNorm's backward folded into its loader (yraw != nullptr; `dy` unused): dY is formed on the fly
  // as scale * (g - c1 - xhat * c2), g = bf16(g1 [+ g2]) * [mask > 0], and never stored (launch_stem_wgrad_tc only)
  const __nv_bfloat16* g1;     // gradient of the stem's ReLU output (main part)
  const __nv_bfloat16* g2;     // its second part (shortcut branch) or nullptr
  const __nv_bfloat16* mask;   // the ReLU output
  const __nv_bfloat16* yraw;   // the raw conv output
  const float* stat;           // [2][Co]: mean, rstd (forward)
  const float* bcoef;          // [3][Co]: scale, c1, c2 (bn_bwd_reduce's finalize)
  float* dw;                   // wgrad: gradient arena slot (atomically accumulated; pre-zeroed)
  int N, S;
  int Ho, Wo, R, tiles_per_frame, ntiles;   // filled by the launcher
  int qpr, pitch;                           // quads (4 output pixels) per output row; floats per row of the x tile
  FastDiv fd_wo;                            // division by Wo (filled by the launcher)
};
struct TailArgs {
  const __nv_bfloat16* in;     // [N,H,W,Ci] activation feeding the tail conv
  const float* w;              // [1][Ci][3][3] fp32
  const float* bias;           // [1] or nullptr
  __nv_bfloat16* y;            // forward: [N,H,W,1]
  float* partials;             // forward: [tiles][1][2] or nullptr
  BnFused bn;                  // forward: fused statistics when bn.acc != nullptr (then `partials` is unused)
  const __nv_bfloat16* dy;     // backward: dY [N,H,W,1]
  __nv_bfloat16* dx;           // backward: dX [N,H,W,Ci]
  float* dw;                   // backward: gradient arena slot (atomically accumulated; pre-zeroed)
  BnBwdFused bb;               // backward: fused BatchNorm-backward reduction of the block that produced `in`
  int N, H, W;
  int R, tiles_per_frame, ntiles;           // filled by the launcher
};
bool stem_supported(int Cin, int Co, int S, int k, int s, int p);
bool tail_supported(int Ci, int Co, int H, int k, int s, int p);
StatLayout launch_stem_fwd(StemArgs a, int Co, cudaStream_t st);
// the same convolution on the tensor cores (stem_tc.cu): im2col built in shared memory, bf16 head + tail split of both operands
bool stem_tc_supported(int Co, int S);
void launch_stem_fwd_tc(StemArgs a, int Co, cudaStream_t st);
bool stem_wgrad_tc_supported(int Co, int S);
void launch_stem_wgrad_tc(StemArgs a, cudaStream_t st);
void launch_stem_wgrad(StemArgs a, int Co, cudaStream_t st);
// the tail convolution's backward pass on the tensor cores (tail_tc.cu)
bool tail_bwd_tc_supported(int Ci, int H, int W);
void launch_tail_bwd_tc(TailArgs a, int Ci, cudaStream_t st);
bool tail_fwd_tc_supported(int Ci, int H, int W);
void launch_tail_fwd_tc(TailArgs a, int Ci, cudaStream_t st);
StatLayout launch_tail_fwd(TailArgs a, int Ci, cudaStream_t st);
void launch_tail_bwd(TailArgs a, int Ci, cudaStream_t st);
void launch_heads_wgrad(const float* dheads, const float* pooled, float* g_mu, float* g_lv, int N, int z, int C,
                        cudaStream_t st);

// ---- pointwise / reduction kernels (pointwise.cu) ----
struct BnFinalizeArgs {
  const float* partials; StatLayout sl; int C; long long m;
  const float* gamma; const float* beta;
  float* running_mean; float* running_var; long long* counter;   // nullptr when not updating
  float* stat;    // [2][C] mean, rstd
  float* coef;    // [2][C] scale, shift
  int training;   // 0: use running stats
};
void launch_bn_finalize(const BnFinalizeArgs& a, cudaStream_t st);

// a = act( y*scale+shift [+ y2*scale2+shift2] ), NHWC, rows x C
template <typename T>
void launch_bn_apply(const T* y, const float* coef, const T* y2, const float* coef2, T* out,
                     long long rows, int C, int relu, cudaStream_t st);
// recon (fp32 NCHW) = y*scale+shift from NHWC storage
template <typename T>
void launch_bn_apply_out(const T* y, const float* coef, float* out_nchw, int N, int HW, int C, cudaStream_t st);

struct HeadsArgs {
  const void* feat;        // encoder output activation NHWC [N,hw,C]
  const float* w_mu; const float* w_lv;   // [z][C] fp32 (w_lv may be nullptr)
  const float* eps;        // [N][z] or nullptr -> Philox
  unsigned long long seed, offset;
  const unsigned long long* rng_dev;   // device {seed, offset} overriding the pair above, or nullptr
  unsigned long long* rng_adv;         // == rng_dev when the kernel is to advance the device offset itself (the CTA that
  unsigned int* rng_ticket;            // finishes last adds rng_inc; ticket: zeroed counter), else nullptr
  unsigned long long rng_inc;
  float* pooled;           // [N][C]
  float* heads;            // [4][N][z]: mu, logvar, eps, std
  float* mu_out; float* lv_out; float* enc_out; float* eps_out;   // user-visible fp32 outputs
  void* z_act;             // decoder input activation [N][z] storage type
  int N, hw, C, z;
};
template <typename T> void launch_heads_fwd(const HeadsArgs& a, cudaStream_t st);
template <typename T> void launch_cast_latent(const float* enc, T* z_act, long long n, cudaStream_t st);

struct HeadsBwdArgs {
  const void* dz_act;      // gradient wrt decoder input [N][z] storage type (or nullptr)
  const float* d_mu; const float* d_lv; const float* d_enc;   // incoming user grads or nullptr
  const float* heads;      // [4][N][z]
  const float* pooled;     // [N][C]
  const float* w_mu; const float* w_lv;
  float* dheads;           // [3][N][z] scratch: dmu, dlogvar
  float* dpool;            // [N][C]
  float* g_wmu; float* g_wlv;   // gradient arena slots
  void* dfeat;             // gradient wrt encoder output activation NHWC [N,hw,C] storage type
  int N, hw, C, z;
};
template <typename T> void launch_heads_bwd(const HeadsBwdArgs& a, cudaStream_t st);

// BatchNorm backward.  g = dA * [a > 0] (mask from the block output when relu), per channel:
//   S0 = sum g, S1 = sum g*xhat(y), S2 = sum g*xhat(y2)
struct BnBwdArgs {
  const void* dA;          // incoming gradient NHWC storage type (or fp32 when dA_f32)
  const void* dA2;         // second part of the incoming gradient (storage type; dA + dA2 is the gradient) or nullptr
  const void* a;           // activation output for the ReLU mask, or nullptr (no ReLU)
  const void* y;  const float* stat;  const float* gamma;           // main branch
  const void* y2; const float* stat2; const float* gamma2;          // second branch or nullptr
  float* partials;         // [blocks][C][3]
  double* acc;             // bf16 mode: [kBnAccCopies][3][C] fp64 accumulators (zeroed) + last-block finalize; nullptr: partials path
  unsigned int* counter;   // blocks finished (zeroed)
  float* bcoef; float* bcoef2;       // [5][C]: scale, c1, c2, S0, S1
  int reduced;             // the producer of dA already masked it and did the reduction (BnBwdFused): apply only, and
                           // publish d gamma / d beta from bcoef
  float* g_gamma; float* g_beta; float* g_gamma2; float* g_beta2;   // gradient arena slots
  void* dY; void* dY2;     // outputs, storage type
  long long rows; int C;
  int dA_f32;
  int no_apply;            // reduce + finalize only (bcoef, d gamma, d beta): the consumer of dY forms it itself
  int late_loads;          // A/B (MMVAE_BN_LATE_LOADS): fetch the forward-written operands only after the dependency wait
};
template <typename T> void launch_bn_bwd(const BnBwdArgs& a, cudaStream_t st);

// NCHW fp32 -> NHWC fp32 (d_recon with more than one channel)
void launch_nchw_to_nhwc(const float* in, float* out, int N, int C, int HW, cudaStream_t st);

// ---- loss (loss.cu) ----
struct LossArgs {
  int kind; float nll, kl, sigma; int N, C, H, W, z;
};
size_t loss_scratch_bytes();
// kl_dev: device scalar overriding a.kl (KL annealing inside a captured graph) or nullptr
void launch_loss_fwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, const float* kl_dev, float* out, void* scratch, cudaStream_t st);
void launch_loss_bwd(const LossArgs& a, const float* recon, const void* target, const float* w,
                     const float* mu, const float* lv, const float* gout, const float* kl_dev,
                     float* d_recon, float* d_mu, float* d_lv, cudaStream_t st);
// MMD diagnostic of model.py:367-383 (x = true_samples, y = encoding, both [N][z] fp32): out[0] = mmd / N
size_t mmd_scratch_bytes(int N);
void launch_mmd(const float* x, const float* y, int N, int z, float* out, void* scratch, cudaStream_t st);
// k-means label map (uint8) -> normalised fp32 network input (+ int64 CE target)   main.py:381-388
void launch_prepare_input(const unsigned char* labels, long long n, float mean, float std, float* x,
                          long long* target, cudaStream_t st);
void launch_philox_normal(unsigned long long seed, unsigned long long offset, const unsigned long long* rng_dev,
                          unsigned long long stream_id, long long n, float* out, cudaStream_t st);
void launch_adam(long long n, float* p, const float* g, float* m, float* v, float lr, float b1, float b2,
                 float eps, float wd, long long step, const long long* step_dev, float gscale, cudaStream_t st);

// ---- notebook-variant VAE: BatchNorm-free pointwise / loss kernels (nb.cu) ----
// nearest-neighbour upsample by f: out[n, f*h+a, f*w+b, c] = in[n,h,w,c]            (vae-kl.ipynb:152-155)
template <typename T> void launch_nb_upsample(const T* in, T* out, int N, int H, int W, int C, int f, cudaStream_t st);
// its adjoint fused with the activation derivative of the tensor that was upsampled:
// dy[n,h,w,c] = act'(a[n,h,w,c]) * sum_{a,b<f} dup[n, f*h+a, f*w+b, c]                 (a == nullptr: no activation)
template <typename T> void launch_nb_upsample_bwd(const T* dup, const T* a, int act_kind, T* dy, int N, int H, int W, int C, int f,
                                                  cudaStream_t st);
struct NbSampleArgs {
  const void* mu_y; const void* lv_y;   // head conv outputs NHWC [N,h,w,z], storage type
  const float* eps;                     // [N,z,h,w] fp32 or nullptr -> Philox
  unsigned long long seed, offset; const unsigned long long* rng_dev;
  float* eps_keep;                      // [N,z,h,w] fp32 workspace copy for the backward
  float* mu_out; float* lv_out; float* enc_out; float* eps_out;   // fp32 NCHW [N,z,h,w] user outputs (any may be nullptr)
  void* z_act;                          // NHWC [N,h,w,z] storage type: decoder input
  int N, hw, z;
};
template <typename T> void launch_nb_rsample(const NbSampleArgs& a, cudaStream_t st);
struct NbSampleBwdArgs {
  const void* dz;                       // NHWC [N,h,w,z] storage type: gradient wrt the decoder input
  const void* mu_y; const void* lv_y; const float* eps_keep;
  void* d_mu_y; void* d_lv_y;           // NHWC storage type
  double* kl_acc;                       // += sum KL terms (unscaled)
  float klw_over_n;                     // kl_weight / N
  int N, hw, z;
};
template <typename T> void launch_nb_rsample_bwd(const NbSampleBwdArgs& a, cudaStream_t st);
// softmax cross-entropy over NHWC logits [rows][C] with int64 targets: dlogits = (softmax - onehot) * scale,
// ce_acc += sum_rows (logsumexp - logit[target]), dbias[c] += sum_rows dlogits[row][c]
template <typename T> void launch_nb_ce(const T* logits, const long long* target, T* dlogits, long long rows, int C, float scale,
                                        double* ce_acc, float* dbias, cudaStream_t st);
// dbias[c] += sum_rows dy[row][c]
template <typename T> void launch_nb_colsum(const T* dy, long long rows, int C, float* dbias, cudaStream_t st);
// NHWC storage -> fp32 NCHW
template <typename T> void launch_nb_export_nchw(const T* in, float* out, int N, int HW, int C, cudaStream_t st);
template <typename T> void launch_nb_import_nchw(const float* in, T* out, int N, int HW, int C, cudaStream_t st);
// encoder.conv1 (1 -> 32, k5 s2 p2) + bias + ReLU and its weight / bias gradient, bf16 storage (dedicated SIMT kernels)
bool nb_stem_supported(int Ci, int Co, int S, int k, int s, int pad);
void launch_nb_stem_fwd(const float* x, const float* w, const float* bias, void* out, int N, int S, cudaStream_t st);
void launch_nb_stem_wgrad(const float* x, const void* dy, float* dw, float* dbias, int N, int S, cudaStream_t st);
// out[0] = ce/N + klw*kl/N, out[1] = ce/N, out[2] = kl/N from the fp64 accumulators acc[0] (ce), acc[1] (kl)
void launch_nb_loss_finalize(const double* acc, float inv_n, float klw, float* out, cudaStream_t st);

// ---- notebook variant: dedicated tcgen05 kernels of decoder.conv4 (32 -> 256 classes, 3x3, 128-pixel rows; nb_tail.cu) ----
struct NbTailArgs {
  const void* x;               // upsampled input activation [N][H][128][32] bf16
  const float* w;              // [256][32][3][3] fp32
  const float* bias;           // [256]
  void* out;                   // forward: logits (target == nullptr) or d logits (cross-entropy mode), [N][H][128][256] bf16
  const long long* target;     // [N][H][128] or nullptr
  double* ce_acc;              // cross-entropy mode: += sum (logsumexp - logit[target])
  float scale;                 // cross-entropy mode: d logits scale (1 / N)
  int N, H;
};
bool nb_tail_supported(int Ci, int Co, int H, int W, int k, int s, int pad);
bool launch_nb_tail_fwd(const NbTailArgs& a, cudaStream_t st);     // false: TMA descriptor could not be encoded
// the 3x3 / stride 1 / 32 -> 32 decoder convolutions on 64- or 32-pixel rows, forward (bias + act) or data gradient
struct NbMidArgs { const void* in; const float* w; const float* bias; void* out; int act, dgrad, N, H, W; };
bool nb_mid_supported(int Ci, int Co, int H, int W, int N, int k, int s, int pad);
bool launch_nb_mid(const NbMidArgs& a, cudaStream_t st);
// their weight + bias gradient: a.in = layer input, dy = gradient of the conv output; dw / dbias pre-zeroed
bool launch_nb_mid_wgrad(const NbMidArgs& a, const void* dy, float* dw, float* dbias, cudaStream_t st);
// data gradient: a.out = d logits, a.w = weights -> dx [N][H][128][32] bf16
bool launch_nb_tail_dgrad(const NbTailArgs& a, void* dx, cudaStream_t st);
// weight + bias gradient: a.x = upsampled input, a.out = d logits; dw / dbias pre-zeroed, accumulated atomically
bool launch_nb_tail_wgrad(const NbTailArgs& a, float* dw, float* dbias, cudaStream_t st);

// ---- device helpers ----
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Philox4x32-10 (Salmon et al. 2011), counter = (c0,c1,0,0), key = (k0,k1)
__device__ __forceinline__ void philox4x32_10(unsigned long long ctr, unsigned long long key, unsigned int out[4]) {
  unsigned int c0 = (unsigned int)ctr, c1 = (unsigned int)(ctr >> 32), c2 = 0u, c3 = 0u;
  unsigned int k0 = (unsigned int)key, k1 = (unsigned int)(key >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    unsigned int hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    unsigned int hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    unsigned int n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// element i of the stream: Box-Muller on the uniform pair (2*(i%2)... ) of counter i/4... see philox_normal_at
__device__ __forceinline__ float philox_normal_at(unsigned long long seed, unsigned long long offset, long long i) {
  unsigned int r[4];
  philox4x32_10(offset + (unsigned long long)(i >> 2), seed, r);
  int pair = (int)((i >> 1) & 1);
  // uniforms in (0,1]: (u + 1) * 2^-32 keeps log() finite
  float u1 = ((float)r[2 * pair] + 1.0f) * 2.3283064365386963e-10f;
  float u2 = ((float)r[2 * pair + 1] + 1.0f) * 2.3283064365386963e-10f;
  u1 = fminf(u1, 1.0f);
  float rad = sqrtf(-2.0f * logf(u1));
  float ang = 6.283185307179586f * u2;
  return (i & 1) ? rad * sinf(ang) : rad * cosf(ang);
}

}  // namespace mmvae
