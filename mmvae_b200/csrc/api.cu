// api.cu -- extern "C" entry points of libmmvae_b200.so (see include/mmvae.h) and the host-side
// orchestration of one VAE step: which kernel runs on which tensors, in which order.
//
// Order of operations follows the reference: VAE.forward model.py:316-342 -> VAE_Encoder.forward
// model.py:114-130 (BasicBlock.forward model.py:39-55) -> rsample model.py:148-150 ->
// VAE_Decoder.forward model.py:181-194 (DeconvBottleneck.forward model.py:70-85); the backward is
// the hand-derived adjoint of the same graph (formulas in SURVEY.md Appendix A / DESIGN.md).
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "kernels.cuh"
#include "plan.hpp"
#include "tc.cuh"

namespace mmvae {

thread_local char g_err[512] = "";
// experiment switch, read once per process: no auxiliary streams (everything on the caller's stream)
static bool aux_disabled() {
  static const bool off = getenv("MMVAE_NO_AUX") != nullptr;
  return off;
}
bool pdl_enabled() {
  static bool on = [] { const char* e = getenv("MMVAE_NO_PDL"); return !(e && e[0] == '1'); }();
  return on;
}
std::atomic<long long> g_launches{0};
static unsigned long long* g_trace_buf = nullptr;
unsigned long long* debug_trace_buffer() { return g_trace_buf; }

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// One check per device per process: the library only carries sm_100a code.
static int check_device() {
  static std::atomic<int> cached[64];
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MMVAE_ERR_NO_DEVICE, "no CUDA device: %s (libmmvae_b200 has no CPU fallback)", cudaGetErrorString(e));
  }
  if (dev >= 0 && dev < 64 && cached[dev].load() == 1) return 0;
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return fail(MMVAE_ERR_ARCH, "device %d is sm_%d%d; libmmvae_b200 is built for sm_100a (B200) only", dev, major, minor);
  if (dev >= 0 && dev < 64) cached[dev].store(1);
  return 0;
}

static int check_launches(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(MMVAE_ERR_CUDA, "%s: CUDA error: %s", what, cudaGetErrorString(e));
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// gather-convolution descriptors
// ------------------------------------------------------------------------------------------------
static void add_tap(GVar& v, int dy, int dx, int wofs) {
  v.dy[v.ntaps] = (signed char)dy; v.dx[v.ntaps] = (signed char)dx; v.wofs[v.ntaps] = wofs; ++v.ntaps;
}

// Gather-convolution descriptor of conv `c` in direction `dir` (geom.hpp enumerates variants and taps).
static void geom_build(const ConvT_& c, int dir, int N, GConvParams& g) {
  memset(&g, 0, sizeof(g));
  const ConvGeom cg = c.geom();
  int op_ci, op_co;
  geom_strides(cg, dir, op_ci, op_co, g.w_sci, g.w_sco);
  g.N = N; g.Ci = op_ci; g.Co = op_co;
  if (dir == DIR_FPROP) {
    g.Hi = c.Hi; g.Wi = c.Wi; g.Ho = c.Ho; g.Wo = c.Wo;
    if (c.kind == CONV) { g.Hg = c.Ho; g.Wg = c.Wo; g.os = 1; g.is = c.s; }
    else if (c.k == 4)  { g.Hg = c.Hi; g.Wg = c.Wi; g.os = 2; g.is = 1; }
    else                { g.Hg = 1; g.Wg = 1; g.os = 2; g.is = 1; }
  } else {
    // data gradient: input = dY [N,Ho,Wo,Co], output = dX [N,Hi,Wi,Ci]
    g.Hi = c.Ho; g.Wi = c.Wo; g.Ho = c.Hi; g.Wo = c.Wi;
    if (c.kind == CONV && c.s == 1)      { g.Hg = c.Hi; g.Wg = c.Wi; g.os = 1; g.is = 1; }
    else if (c.kind == CONV)             { g.Hg = (c.Hi + 1) / 2; g.Wg = (c.Wi + 1) / 2; g.os = 2; g.is = 1; }
    else if (c.k == 4)                   { g.Hg = c.Hi; g.Wg = c.Wi; g.os = 1; g.is = 2; }
    else                                 { g.Hg = 1; g.Wg = 1; g.os = 1; g.is = 1; }
  }
  g.nvar = geom_nvar(cg, dir);
  for (int v = 0; v < g.nvar; ++v) {
    GVar& gv = g.var[v];
    geom_origin(cg, dir, v, gv.oy0, gv.ox0);
    int dy, dx, wofs;
    while (geom_tap(cg, dir, v, gv.ntaps, dy, dx, wofs)) add_tap(gv, dy, dx, wofs);
  }
  g.M = N * g.Hg * g.Wg;
  if (c.kind == CONVT && c.k == 4 && c.s == 2 && c.p == 1) g.conv_class = dir == DIR_FPROP ? 1 : 2;
  if (c.wp_chunks[dir] > 0) {
    g.co_pad = (op_co + 15) & ~15;
    g.wpack_var_stride = c.wp_chunks[dir] * g.co_pad * 128;
  }
}
static void geom_fprop(const ConvT_& c, int N, GConvParams& g) { geom_build(c, DIR_FPROP, N, g); }
static void geom_dgrad(const ConvT_& c, int N, GConvParams& g) { geom_build(c, DIR_DGRAD, N, g); }

// ------------------------------------------------------------------------------------------------
// auxiliary stream: work that is off the critical path of the step (weight gradients, the shortcut
// branch of a block) is enqueued on a second, per-device stream that forks from and joins back into
// the caller's stream with events -- eagerly this is plain concurrency, under stream capture it becomes
// a parallel branch of the CUDA graph.  The pair (stream, events) is the only state the library keeps.
// ------------------------------------------------------------------------------------------------
struct AuxPool {
  cudaStream_t s = nullptr;      // weight gradients, shortcut convs, weight packing
  cudaStream_t s2 = nullptr;     // the shortcut branch's data gradient of a block (backward), beside the main branch's chain
  cudaEvent_t ev[32];
  int next = 0;
  bool ok = false;
};
static AuxPool* aux_pool() {
  static AuxPool pools[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  AuxPool& p = pools[dev];
  if (!p.ok) {
    // lowest priority: when both streams have CTAs waiting for an SM, the critical path (the caller's stream) goes first
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    static const bool flat = getenv("MMVAE_AUX_FLAT_PRIORITY") != nullptr;       // A/B: default priority (measured +0.4 % step time)
    if (cudaStreamCreateWithPriority(&p.s, cudaStreamNonBlocking, flat ? 0 : prio_lo) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (cudaStreamCreateWithFlags(&p.s2, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    for (auto& e : p.ev)
      if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    p.ok = true;
  }
  return &p;
}

// ------------------------------------------------------------------------------------------------
// step executor
// ------------------------------------------------------------------------------------------------
// The stem BatchNorm's backward folded into the stem weight gradient's loader (stem_tc.cu, kFuse): bf16 tensor-core path
// with the tcgen05 stem weight gradient available.  Then the gradient of the raw stem output is never materialised
// (mmvae_workspace_tensor reports it as absent).  MMVAE_NO_STEM_BWD_FUSE: A/B and the stored-dY validation path.
static bool stem_bwd_fused(const Plan& P) {
  static const bool off = getenv("MMVAE_NO_STEM_BWD_FUSE") != nullptr;
  if (off || P.d.arch != MMVAE_ARCH_RESNET || P.d.precision != MMVAE_PREC_BF16 || (P.d.flags & MMVAE_FLAG_FORCE_SIMT) || P.stem < 0)
    return false;
  const ConvT_& c = P.convs[P.stem];
  return c.in < 0 && stem_supported(c.Ci, c.Co, c.Hi, c.k, c.s, c.p) && stem_wgrad_tc_supported(c.Co, c.Hi);
}

template <typename T>
struct Exec {
  const Plan& P;
  char* ws;
  const float* params;
  float* grads;
  float* bnbuf;
  long long* counters;
  cudaStream_t st;
  const float* x;       // fp32 NCHW network input

  AuxPool* aux = nullptr;       // set by use_aux(): second stream for off-critical-path work

  template <typename U> U* at(size_t off) const { return reinterpret_cast<U*>(ws + off); }
  const ActT& act(int i) const { return P.acts[i]; }

  void use_aux() { if (special_ok() && !aux_disabled()) aux = aux_pool(); }
  cudaEvent_t next_event() { cudaEvent_t e = aux->ev[aux->next]; aux->next = (aux->next + 1) & 31; return e; }
  // run f() on the auxiliary stream, ordered after everything enqueued on the main stream so far
  template <typename F> void side(F f) {
    if (!aux) { f(); return; }
    cudaEvent_t e = next_event();
    cudaEventRecord(e, st);
    cudaStreamWaitEvent(aux->s, e, 0);
    cudaStream_t keep = st;
    st = aux->s;
    f();
    st = keep;
    forked = true;
  }
  // the main stream waits for everything enqueued on the auxiliary stream
  void join() {
    if (!aux || !forked) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, aux->s);
    cudaStreamWaitEvent(st, e, 0);
    forked = false;
  }
  bool forked = false;
  // the same pair for the second auxiliary stream
  template <typename F> void side2(F f) {
    if (!aux) { f(); return; }
    cudaEvent_t e = next_event();
    cudaEventRecord(e, st);
    cudaStreamWaitEvent(aux->s2, e, 0);
    cudaStream_t keep = st;
    st = aux->s2;
    f();
    st = keep;
    forked2 = true;
  }
  void join2() {
    if (!aux || !forked2) return;
    cudaEvent_t e = next_event();
    cudaEventRecord(e, aux->s2);
    cudaStreamWaitEvent(st, e, 0);
    forked2 = false;
  }
  bool forked2 = false;

  // dedicated kernels of the 1-channel stem / tail convolutions (bf16 mode)
  bool special_ok() const { return std::is_same<T, __nv_bfloat16>::value && !(P.d.flags & MMVAE_FLAG_FORCE_SIMT); }
  bool use_stem(const ConvT_& c) const {
    return special_ok() && P.stem >= 0 && c.in < 0 && &c == &P.convs[P.stem] && stem_supported(c.Ci, c.Co, c.Hi, c.k, c.s, c.p);
  }
  bool use_tail(const ConvT_& c) const {
    return special_ok() && P.tail >= 0 && &c == &P.convs[P.tail] && c.kind == CONV && tail_supported(c.Ci, c.Co, c.Hi, c.k, c.s, c.p);
  }

  // fp32 master weights -> bf16 tiles of the tcgen05 kernels (both directions), one launch
  void pack_weights() {
    PackTable tab;
    tab.n = 0;
    for (const ConvT_& c : P.convs)
      for (int dir = 0; dir < 2; ++dir) {
        if (c.wp_chunks[dir] <= 0 || tab.n >= kMaxPackOps) continue;
        PackOp& o = tab.ops[tab.n++];
        o.w_off = c.w; o.dst_off16 = (unsigned int)(c.wp_off[dir] / 16);
        o.Ci = (unsigned short)c.Ci; o.Co = (unsigned short)c.Co; o.maxchunks = (unsigned short)c.wp_chunks[dir];
        o.kind = (unsigned char)(c.kind == CONV ? GEOM_CONV : GEOM_CONVT);
        o.k = (unsigned char)c.k; o.s = (unsigned char)c.s; o.p = (unsigned char)c.p; o.dir = (unsigned char)dir;
      }
    launch_pack_weights(tab, params, ws, st);
  }

  // fused training-mode statistics of BatchNorm b (bn_fused.cuh): accumulators were zeroed by clear_bn_acc()
  BnFused bn_fused(const BnT& b) const {
    BnFused f{};
    f.acc = at<double>(b.acc_off); f.counter = at<unsigned int>(b.cnt_off);
    f.gamma = params + b.gamma; f.beta = params + b.beta;
    f.running_mean = bnbuf ? bnbuf + b.rm : nullptr;
    f.running_var = bnbuf ? bnbuf + b.rm + b.C : nullptr;
    f.nbt = counters ? counters + b.idx : nullptr;
    f.stat = at<float>(b.stat_off); f.coef = at<float>(b.coef_off);
    f.C = b.C; f.inv_m = 1.0 / (double)b.m; f.unbias = b.m > 1 ? (double)b.m / (double)(b.m - 1) : 1.0;
    return f;
  }
  void clear_bn_acc() { cudaMemsetAsync(ws + P.bnacc_off, 0, P.bnacc_bytes, st); }

  // y = conv(in) + BatchNorm statistics (fused into the producing kernel on the bf16 path; per-CTA partial rows
  // and a finalize kernel on the SIMT path and in eval mode)
  void conv_bn_fwd(const ConvT_& c) {
    GConvParams g;
    geom_fprop(c, P.d.batch, g);
    if (c.in < 0) { g.in = x; g.in_nchw_f32 = 1; }
    else g.in = at<T>(act(c.in).off);
    g.out = at<T>(act(c.out).off);
    g.w = params + c.w;
    g.wpack = c.wp_chunks[DIR_FPROP] > 0 ? ws + c.wp_off[DIR_FPROP] : nullptr;
    g.bias = c.bias >= 0 ? params + c.bias : nullptr;
    const BnT& b = P.bns[c.bn];
    const bool training = P.d.training != 0;
    bool fused = false;
    StatLayout sl{0, 0, 0, 0};
    if (use_stem(c)) {
      StemArgs a{};
      a.x = x; a.w = g.w; a.y = at<__nv_bfloat16>(act(c.out).off); a.N = P.d.batch; a.S = c.Hi;
      if (training) { a.bn = bn_fused(b); fused = true; }
      launch_stem_fwd(a, c.Co, st);
    } else if (use_tail(c)) {
      TailArgs a{};
      a.in = at<__nv_bfloat16>(act(c.in).off); a.w = g.w; a.bias = g.bias; a.y = at<__nv_bfloat16>(act(c.out).off);
      a.N = P.d.batch; a.H = c.Hi; a.W = c.Wi;
      if (training) { a.bn = bn_fused(b); fused = true; }
      launch_tail_fwd(a, c.Ci, st);
    } else if (std::is_same<T, __nv_bfloat16>::value && tc_supported_gconv(g)) {
      if (training) { g.bn = bn_fused(b); fused = true; }
      launch_gconv_tc(g, st);
    } else {
      g.partials = training ? at<float>(b.part_off) : nullptr;
      sl = launch_gconv_simt<T>(g, st);
    }
    if (fused) return;
    BnFinalizeArgs f;
    f.partials = training ? at<float>(b.part_off) : nullptr; f.sl = sl; f.C = b.C; f.m = b.m;
    f.gamma = params + b.gamma; f.beta = params + b.beta;
    f.running_mean = bnbuf ? bnbuf + b.rm : nullptr;
    f.running_var = bnbuf ? bnbuf + b.rm + b.C : nullptr;
    f.counter = counters ? counters + b.idx : nullptr;
    f.stat = at<float>(b.stat_off); f.coef = at<float>(b.coef_off);
    f.training = P.d.training;
    launch_bn_finalize(f, st);
  }

  void apply(const ConvT_& c, const ConvT_* c2, int out_act, int relu) {
    const ActT& o = act(out_act);
    long long rows = (long long)P.d.batch * o.H * o.W;
    launch_bn_apply<T>(at<T>(act(c.out).off), at<float>(P.bns[c.bn].coef_off),
                       c2 ? at<T>(act(c2->out).off) : nullptr, c2 ? at<float>(P.bns[c2->bn].coef_off) : nullptr,
                       at<T>(o.off), rows, o.C, relu, st);
  }

  void block_fwd(const BlockT& b) {
    const ConvT_& c1 = P.convs[b.c1]; const ConvT_& c2 = P.convs[b.c2]; const ConvT_& cs = P.convs[b.cs];
    side([&] { conv_bn_fwd(cs); });                   // shortcut branch: only needs the block input
    conv_bn_fwd(c1);
    apply(c1, nullptr, b.a1, 1);
    conv_bn_fwd(c2);
    join();
    apply(c2, &cs, b.out, 1);
  }

  void encode(const float* eps, unsigned long long seed, unsigned long long offset, const uint64_t* rng_state,
              float* eps_out, float* mu, float* logvar, float* enc) {
    const ConvT_& s = P.convs[P.stem];
    conv_bn_fwd(s);
    apply(s, nullptr, P.a_stem, 1);
    join();                                           // weight packing (forked by mmvae_forward) ran beside the stem
    for (const BlockT& b : P.enc) block_fwd(b);
    HeadsArgs h;
    const ActT& f = act(P.enc.back().out);
    h.feat = at<T>(f.off);
    h.w_mu = params + P.w_mu; h.w_lv = P.w_lv >= 0 ? params + P.w_lv : nullptr;
    h.eps = eps; h.seed = seed; h.offset = offset; h.rng_dev = reinterpret_cast<const unsigned long long*>(rng_state);
    h.rng_adv = nullptr; h.rng_ticket = nullptr; h.rng_inc = 0;
    if (rng_state && !eps && P.d.require_rsample) {
      // device-resident generator state: the kernel advances the offset itself (its last CTA, through a ticket that the
      // bf16 path zeroes with its per-forward memset, clear_bn_acc)
      h.rng_adv = const_cast<unsigned long long*>(h.rng_dev);
      h.rng_ticket = at<unsigned int>(P.ticket_off);
      h.rng_inc = (unsigned long long)(((long long)P.d.batch * P.d.z_dim + 3) / 4);
      if (!std::is_same<T, __nv_bfloat16>::value) cudaMemsetAsync(h.rng_ticket, 0, sizeof(unsigned int), st);
    }
    h.pooled = at<float>(P.pooled_off); h.heads = at<float>(P.heads_off);
    h.mu_out = mu; h.lv_out = logvar; h.enc_out = enc; h.eps_out = eps_out;
    h.z_act = at<T>(act(P.a_z).off);
    h.N = P.d.batch; h.hw = P.feat_hw; h.C = P.feat_c; h.z = P.d.z_dim;
    launch_heads_fwd<T>(h, st);
  }

  void decode(float* recon) {
    const ConvT_& s = P.convs[P.dstem];
    conv_bn_fwd(s);
    apply(s, nullptr, P.a_dstem, 1);
    for (const BlockT& b : P.dec) block_fwd(b);
    const ConvT_& t = P.convs[P.tail];
    conv_bn_fwd(t);
    launch_bn_apply_out<T>(at<T>(act(t.out).off), at<float>(P.bns[t.bn].coef_off), recon, P.d.batch,
                           t.Ho * t.Wo, t.Co, st);
  }

  // ---------------- backward ----------------
  void wgrad(const ConvT_& c) {
    if (use_stem(c)) {
      StemArgs a{};
      a.x = x; a.dy = at<__nv_bfloat16>(act(c.out).goff); a.dw = grads + c.w; a.N = P.d.batch; a.S = c.Hi;
      launch_stem_wgrad(a, c.Co, st);
      return;
    }
    GConvParams g;
    geom_fprop(c, P.d.batch, g);
    if (P.d.arch == MMVAE_ARCH_NOTEBOOK) nb_pad_grid(c, g);
    WGradParams w;
    memset(&w, 0, sizeof(w));
    if (P.d.arch == MMVAE_ARCH_NOTEBOOK) w.cta_budget = 120;      // no BatchNorm chain beside it to leave SMs for
    if (c.in < 0) { w.in = x; w.in_nchw_f32 = 1; }
    else w.in = at<T>(act(c.in).off);
    w.dout = at<T>(act(c.out).goff);
    w.dw = grads + c.w;
    w.N = g.N; w.Hi = g.Hi; w.Wi = g.Wi; w.Ci = g.Ci; w.Ho = g.Ho; w.Wo = g.Wo; w.Co = g.Co;
    w.Hg = g.Hg; w.Wg = g.Wg; w.M = g.M; w.os = g.os; w.is = g.is; w.w_sci = g.w_sci; w.w_sco = g.w_sco;
    w.nvar = g.nvar; w.conv_class = g.conv_class;
    for (int i = 0; i < g.nvar; ++i) w.var[i] = g.var[i];
    conv_wgrad<T>(w, c, !(P.d.flags & MMVAE_FLAG_FORCE_SIMT), st);
  }

  // ---- BatchNorm-backward reduction fused into the producer of the incoming gradient (BnBwdFused) ----
  // Which BatchNorm(s) consume d(activation `a`): conv c (+ the shortcut conv c2 at a block output); mask = a itself.
  bool consumer_of(int a, const ConvT_*& c, const ConvT_*& c2) const {
    c = c2 = nullptr;
    if (a == P.a_stem) { c = &P.convs[P.stem]; return true; }
    if (a == P.a_dstem) { c = &P.convs[P.dstem]; return true; }
    for (const std::vector<BlockT>* bl : {&P.enc, &P.dec})
      for (const BlockT& b : *bl) {
        if (a == b.a1) { c = &P.convs[b.c1]; return true; }
        if (a == b.out) { c = &P.convs[b.c2]; c2 = &P.convs[b.cs]; return true; }
      }
    return false;
  }
  // The conv whose data gradient completes d(activation `a`): the shortcut of the block that reads `a` (it accumulates
  // onto the main branch's gradient), or conv2 for a block's inner activation.  -1: produced by another kernel.
  int last_dgrad_conv(int a) const {
    for (const std::vector<BlockT>* bl : {&P.enc, &P.dec})
      for (const BlockT& b : *bl) {
        if (a == b.in) return b.cs;
        if (a == b.a1) return b.c2;
      }
    return -1;
  }
  void fill_dgrad(const ConvT_& c, GConvParams& g) const {
    geom_dgrad(c, P.d.batch, g);
    g.in = at<T>(act(c.out).goff);
    g.out = at<T>(act(c.in).goff);
    g.w = params + c.w;
    g.wpack = c.wp_chunks[DIR_DGRAD] > 0 ? ws + c.wp_off[DIR_DGRAD] : nullptr;
  }
  // Does the data gradient of conv c reach every pixel of its input?  (A strided 1x1 shortcut only reaches the even ones.)
  static bool dgrad_covers_all(const GConvParams& g) { return g.os == 1 || g.nvar == g.os * g.os; }
  // The block whose shortcut is conv index ci (or nullptr)
  const BlockT* block_of_shortcut(int ci) const {
    for (const std::vector<BlockT>* bl : {&P.enc, &P.dec})
      for (const BlockT& b : *bl) if (b.cs == ci) return &b;
    return nullptr;
  }
  // A pure function of the plan (backward phases run in separate calls): do the producers of d(a) mask and reduce?
  bool fused_reduce(int a) const {
    static const bool off = getenv("MMVAE_NO_BWD_FUSE") != nullptr;
    if (off || a < 0 || !special_ok()) return false;
    // the last decoder block's output gradient comes out of the fused tail backward kernel (special.cu)
    if (!P.dec.empty() && a == P.dec.back().out && a == P.convs[P.tail].in && use_tail(P.convs[P.tail])) return true;
    const int ci = last_dgrad_conv(a);
    const ConvT_ *c, *c2;
    if (ci < 0 || !consumer_of(a, c, c2)) return false;
    GConvParams g;
    fill_dgrad(P.convs[ci], g);
    if (!tc_supported_gconv(g)) return false;
    // measured (profiles/r01_fused_bn_bwd.md): with one epilogue warp per scheduler the extra loads and column sums
    // only pay off for 16-channel tensors (one 16-column group per tile); wider ones keep the separate reduce kernel
    static const bool all = getenv("MMVAE_BWD_FUSE_ALL") != nullptr;
    if (g.Co > 16 && !all) return false;
    if (dgrad_covers_all(g)) return true;
    // partial coverage (even pixels only): the main branch's data gradient handles the other parity classes
    const BlockT* b = block_of_shortcut(ci);
    if (!b || g.nvar != 1 || g.os != 2 || g.var[0].oy0 != 0 || g.var[0].ox0 != 0) return false;
    GConvParams g1;
    fill_dgrad(P.convs[b->c1], g1);
    return tc_supported_gconv(g1) && g1.os == 2 && g1.nvar == 4;
  }

  // descriptor of the BatchNorm-backward reduction that consumes d(activation `a`) (every pixel, last contributor)
  void fill_bwd_fused(int a, BnBwdFused& f) const {
    const ConvT_ *bc, *bc2;
    consumer_of(a, bc, bc2);
    const BnT& b = P.bns[bc->bn];
    f.a = at<T>(act(a).off);
    f.y = at<T>(act(bc->out).off); f.stat = at<float>(b.stat_off); f.gamma = params + b.gamma;
    f.acc = at<double>(b.acc_off) + kBnAccCopies * 2 * b.C; f.counter = at<unsigned int>(b.cnt_off) + 1;
    f.bcoef = at<float>(b.bcoef_off);
    if (bc2) {
      const BnT& b2 = P.bns[bc2->bn];
      f.y2 = at<T>(act(bc2->out).off); f.stat2 = at<float>(b2.stat_off); f.gamma2 = params + b2.gamma;
      f.bcoef2 = at<float>(b2.bcoef_off);
    }
    f.C = b.C; f.var_mask = 0xF; f.finish = 1;
    f.inv_rows = 1.0 / ((double)P.d.batch * bc->Ho * bc->Wo);
  }

  // Block inputs whose gradient is kept in two parts (Plan: goff2): the shortcut branch's data gradient goes to the
  // second buffer on the second auxiliary stream, beside the main branch's bn_bwd -> dgrad -> bn_bwd -> dgrad chain, and the
  // BatchNorm backward that consumes the gradient adds the parts (BnBwdArgs::dA2): one launch less on the critical path
  // of every block.
  bool split_grad(int a) const {
    static const bool off = getenv("MMVAE_NO_SPLIT_GRAD") != nullptr;
    return !off && aux && a >= 0 && act(a).goff2 != 0 && !fused_reduce(a);
  }

  void dgrad(const ConvT_& c, int accumulate, bool to_alt = false) {
    GConvParams g;
    fill_dgrad(c, g);
    g.accumulate = accumulate;
    if (to_alt) g.out = at<T>(act(c.in).goff2);
    if (c.in >= 0 && fused_reduce(c.in)) {
      const int last = last_dgrad_conv(c.in);
      GConvParams gl;
      fill_dgrad(P.convs[last], gl);
      int var_mask = 0, finish = 0;
      if (&P.convs[last] == &c) { var_mask = (1 << g.nvar) - 1; finish = 1; }
      else if (!dgrad_covers_all(gl)) {              // c is the main-branch conv1 of an encoder block: all but (even, even)
        for (int v = 0; v < g.nvar; ++v) if (g.var[v].oy0 != 0 || g.var[v].ox0 != 0) var_mask |= 1 << v;
      }
      if (var_mask) {
        fill_bwd_fused(c.in, g.bb);
        g.bb.var_mask = var_mask; g.bb.finish = finish;
      }
    }
    conv_dgrad<T>(g, c, st);
  }

  // BatchNorm (+ReLU mask from `mask_act`) backward of conv c (and of the parallel branch c2)
  void bn_bwd(const void* dA, int dA_f32, int mask_act, const ConvT_& c, const ConvT_* c2, bool no_apply = false) {
    const BnT& b = P.bns[c.bn];
    BnBwdArgs a;
    memset(&a, 0, sizeof(a));
    a.dA = dA; a.dA_f32 = dA_f32;
    a.a = mask_act >= 0 ? at<T>(act(mask_act).off) : nullptr;
    a.y = at<T>(act(c.out).off); a.stat = at<float>(b.stat_off); a.gamma = params + b.gamma;
    a.partials = at<float>(b.bpart_off);
    if (special_ok()) { a.acc = at<double>(b.acc_off) + kBnAccCopies * 2 * b.C; a.counter = at<unsigned int>(b.cnt_off) + 1; }
    a.bcoef = at<float>(b.bcoef_off);
    a.g_gamma = grads + b.gamma; a.g_beta = grads + b.beta;
    a.dY = at<T>(act(c.out).goff);
    if (c2) {
      const BnT& b2 = P.bns[c2->bn];
      a.y2 = at<T>(act(c2->out).off); a.stat2 = at<float>(b2.stat_off); a.gamma2 = params + b2.gamma;
      a.bcoef2 = at<float>(b2.bcoef_off);
      a.g_gamma2 = grads + b2.gamma; a.g_beta2 = grads + b2.beta;
      a.dY2 = at<T>(act(c2->out).goff);
    }
    a.rows = (long long)P.d.batch * c.Ho * c.Wo; a.C = c.Co;
    if (mask_act >= 0 && fused_reduce(mask_act)) { a.reduced = 1; a.a = nullptr; }   // dA arrives masked and reduced
    if (mask_act >= 0 && !dA_f32 && dA == (const void*)at<T>(act(mask_act).goff) && split_grad(mask_act))
      a.dA2 = at<T>(act(mask_act).goff2);
    a.no_apply = no_apply ? 1 : 0;
    launch_bn_bwd<T>(a, st);
  }

  void block_bwd(const BlockT& b) {
    const ConvT_& c1 = P.convs[b.c1]; const ConvT_& c2 = P.convs[b.c2]; const ConvT_& cs = P.convs[b.cs];
    bn_bwd(at<T>(act(b.out).goff), 0, b.out, c2, &cs);
    bool split = split_grad(b.in);
    { static const int dbg = [] { const char* e = getenv("MMVAE_SPLIT_DBG"); return e ? atoi(e) : 0; }();   // 1: decoder blocks only, 2: encoder only
      GConvParams gq; fill_dgrad(cs, gq);
      if (dbg == 1 && !dgrad_covers_all(gq)) split = false;
      if (dbg == 2 && dgrad_covers_all(gq)) split = false; }
    if (split) {
      side2([&] {                                     // shortcut branch -> second part of d in, off the critical path
        GConvParams g;
        fill_dgrad(cs, g);
        if (!dgrad_covers_all(g))                     // a strided 1x1 shortcut only reaches the even pixels
          cudaMemsetAsync(at<T>(act(b.in).goff2), 0, sizeof(T) * size_t(P.d.batch) * act(b.in).H * act(b.in).W * act(b.in).C, st);
        dgrad(cs, 0, true);
      });
    }
    side([&] { wgrad(c2); wgrad(cs); });              // weight gradients are off the critical path
    dgrad(c2, 0);                                     // -> d a1
    bn_bwd(at<T>(act(b.a1).goff), 0, b.a1, c1, nullptr);
    side([&] { wgrad(c1); });
    dgrad(c1, 0);                                     // -> d in
    if (split) join2();
    else dgrad(cs, 1);                                // += shortcut
  }

  // ---------------- notebook variant (MMVAE_ARCH_NOTEBOOK; vae-kl.ipynb:119-166, loop body :210-233) ----------------
  // No BatchNorm: bias + activation live in the conv epilogue, the activation derivative in the data-gradient epilogue
  // of the consumer (or in the upsample adjoint when an upsample sits between producer and consumer).
  // A conv whose output is odd-sized (encoder.conv2: 64 -> 31) runs on a gather grid rounded up to the next even size: the
  // extra row / column of outputs is masked in the epilogues (oy < Ho, ox < Wo) and reads as zero through TMA's
  // out-of-bounds fill on the dY side, and a 128- / 64-pixel tile of a 32 x 32 grid is a TMA box where one of a 31 x 31 grid
  // is not (cp.async gather path: 0.2 ms forward, 1.1 ms weight gradient at 512 frames).
  template <typename G> void nb_pad_grid(const ConvT_& c, G& g) const {
    if (c.kind == CONV && (c.Ho & 1) && c.Ho > 1) { g.Hg = c.Ho + 1; g.Wg = c.Wo + 1; g.M = P.d.batch * g.Hg * g.Wg; }
  }
  void nb_conv_fwd(const ConvT_& c, int act_kind) {
    GConvParams g;
    geom_fprop(c, P.d.batch, g);
    nb_pad_grid(c, g);
    if (c.in < 0) { g.in = x; g.in_nchw_f32 = 1; }
    else g.in = at<T>(act(c.in).off);
    g.out = at<T>(act(c.out).off);
    g.w = params + c.w; g.bias = params + c.bias; g.act = act_kind;
    g.wpack = c.wp_chunks[DIR_FPROP] > 0 ? ws + c.wp_off[DIR_FPROP] : nullptr;
    conv_forward<T>(g, c, st);
  }
  int nb_dec_src(int k) const { return k > 0 ? P.convs[P.nb.dc[k - 1]].out : P.nb.a_z; }   // tensor upsampled into decoder.conv{k+1}

  void nb_forward(const float* eps, unsigned long long seed, unsigned long long offset, const uint64_t* rng_state,
                  float* eps_out, float* mu, float* logvar, float* enc, float* recon) {
    const NbT& nb = P.nb;
    const int N = P.d.batch;
    for (int i = 0; i < 4; ++i) {                                                 // vae-kl.ipynb:134-137
      const ConvT_& c = P.convs[nb.e[i]];
      if (i == 0 && nb_stem_ok()) launch_nb_stem_fwd(x, params + c.w, params + c.bias, at<T>(act(c.out).off), N, c.Hi, st);
      else nb_conv_fwd(c, ACT_RELU);
    }
    const ConvT_& cmu = P.convs[nb.cmu]; const ConvT_& clv = P.convs[nb.clv];
    nb_conv_fwd(cmu, ACT_NONE);                                                   // vae-kl.ipynb:139-140
    nb_conv_fwd(clv, ACT_NONE);
    NbSampleArgs a{};
    a.mu_y = at<T>(act(cmu.out).off); a.lv_y = at<T>(act(clv.out).off);
    a.eps = eps; a.seed = seed; a.offset = offset; a.rng_dev = reinterpret_cast<const unsigned long long*>(rng_state);
    a.eps_keep = at<float>(nb.eps_off);
    a.mu_out = mu; a.lv_out = logvar; a.enc_out = enc; a.eps_out = eps_out;
    a.z_act = at<T>(act(nb.a_z).off);
    a.N = N; a.hw = nb.latent_hw * nb.latent_hw; a.z = P.d.z_dim;
    launch_nb_rsample<T>(a, st);                                                  // vae-kl.ipynb:144-146
    nb_decode(recon);
  }
  void nb_decode(float* recon) {
    const NbT& nb = P.nb;
    const int N = P.d.batch;
    for (int k = 0; k < 4; ++k) {                                                 // vae-kl.ipynb:162-166
      const ActT& src = act(nb_dec_src(k));
      launch_nb_upsample<T>(at<T>(src.off), at<T>(act(nb.a_up[k]).off), N, src.H, src.W, src.C, nb.up[k], st);
      if (k == 3 && nb_tail_ok()) {
        if (nb_deferred() && !recon) break;           // fused with the cross-entropy in nb_loss_backward
        if (nb_tail_fwd(nullptr, nullptr, 0.f, at<T>(act(P.convs[nb.dc[3]].out).off))) continue;
      }
      if (nb_mid(P.convs[nb.dc[k]], false, k < 3 ? ACT_ELU : ACT_NONE)) continue;
      nb_conv_fwd(P.convs[nb.dc[k]], k < 3 ? ACT_ELU : ACT_NONE);
    }
    if (recon) {
      const ConvT_& c = P.convs[nb.dc[3]];
      launch_nb_export_nchw<T>(at<T>(act(c.out).off), recon, N, c.Ho * c.Wo, c.Co, st);
    }
  }
  // dedicated row-band kernel of the 32 -> 32 3x3 decoder convs (nb_mid_kernel, nb_tail.cu): forward or data gradient
  bool nb_mid(const ConvT_& c, bool dgrad, int act_kind) {
    if (!special_ok() || !nb_mid_supported(c.Ci, c.Co, c.Hi, c.Wi, P.d.batch, c.k, c.s, c.p)) return false;
    NbMidArgs a{};
    a.in = dgrad ? at<T>(act(c.out).goff) : at<T>(act(c.in).off);
    a.out = dgrad ? at<T>(act(c.in).goff) : at<T>(act(c.out).off);
    a.w = params + c.w; a.bias = dgrad ? nullptr : params + c.bias; a.act = act_kind; a.dgrad = dgrad ? 1 : 0;
    a.N = P.d.batch; a.H = c.Hi; a.W = c.Wi;
    return launch_nb_mid(a, st);
  }
  bool nb_stem_ok() const {
    const ConvT_& c = P.convs[P.nb.e[0]];
    return special_ok() && nb_stem_supported(c.Ci, c.Co, c.Hi, c.k, c.s, c.p);
  }
  // dedicated tcgen05 kernels of decoder.conv4 (nb_tail.cu)
  bool nb_tail_ok() const {
    const ConvT_& c = P.convs[P.nb.dc[3]];
    return special_ok() && nb_tail_supported(c.Ci, c.Co, c.Hi, c.Wi, c.k, c.s, c.p);
  }
  bool nb_deferred() const { return (P.d.flags & MMVAE_FLAG_DEFER_LOGITS) && nb_tail_ok(); }
  bool nb_tail_fwd(const long long* target, double* ce_acc, float scale, void* out) {
    const ConvT_& c = P.convs[P.nb.dc[3]];
    NbTailArgs a{};
    a.x = at<T>(act(c.in).off); a.w = params + c.w; a.bias = params + c.bias; a.out = out;
    a.target = target; a.ce_acc = ce_acc; a.scale = scale; a.N = P.d.batch; a.H = c.Hi;
    return launch_nb_tail_fwd(a, st);
  }
  void nb_dgrad(const ConvT_& c, int accumulate, int dact_kind, int dact_act) {
    GConvParams g;
    fill_dgrad(c, g);
    g.accumulate = accumulate;
    if (dact_act >= 0) { g.dact = at<T>(act(dact_act).off); g.dact_kind = dact_kind; }
    conv_dgrad<T>(g, c, st);
  }
  void nb_param_grads(const ConvT_& c, bool with_bias = true) {
    wgrad(c);
    if (with_bias)
      launch_nb_colsum<T>(at<T>(act(c.out).goff), (long long)P.d.batch * c.Ho * c.Wo, c.Co, grads + c.bias, st);
  }
  // loss = sum CE / N + klw * sum KL / N (vae-kl.ipynb:225-228) and its gradient wrt every parameter
  void nb_loss_backward(const long long* target, float klw, float* out) {
    const NbT& nb = P.nb;
    const int N = P.d.batch;
    const float inv_n = 1.0f / (float)N;
    cudaMemsetAsync(grads, 0, sizeof(float) * size_t(P.n_params), st);
    cudaMemsetAsync(ws + nb.acc_off, 0, sizeof(double) * 2, st);
    double* acc = at<double>(nb.acc_off);
    const ConvT_& c4 = P.convs[nb.dc[3]];
    const long long rows4 = (long long)N * c4.Ho * c4.Wo;
    bool ce_bias = false;                               // decoder.conv4's bias gradient already produced
    if (!(nb_deferred() && nb_tail_fwd(target, acc, inv_n, at<T>(act(c4.out).goff)))) {   // fused: d logits written directly
      launch_nb_ce<T>(at<T>(act(c4.out).off), target, at<T>(act(c4.out).goff), rows4, c4.Co, inv_n, acc, grads + c4.bias, st);
      ce_bias = true;
    }
    for (int k = 3; k >= 0; --k) {
      const ConvT_& c = P.convs[nb.dc[k]];
      if (k == 3 && nb_tail_ok()) {
        side([&] {
          NbTailArgs a{};
          a.x = at<T>(act(c.in).off); a.out = at<T>(act(c.out).goff); a.N = N; a.H = c.Hi;
          if (!launch_nb_tail_wgrad(a, grads + c.w, ce_bias ? nullptr : grads + c.bias, st)) nb_param_grads(c, !ce_bias);
        });
      } else if (special_ok() && nb_mid_supported(c.Ci, c.Co, c.Hi, c.Wi, N, c.k, c.s, c.p)) {
        side([&] {
          NbMidArgs a{};
          a.in = at<T>(act(c.in).off); a.N = N; a.H = c.Hi; a.W = c.Wi;
          if (!launch_nb_mid_wgrad(a, at<T>(act(c.out).goff), grads + c.w, grads + c.bias, st)) nb_param_grads(c);
        });
      } else {
        side([&] { nb_param_grads(c, !(k == 3 && ce_bias)); });
      }
      bool dg_done = false;
      if (k == 3 && nb_tail_ok()) {
        NbTailArgs a{};
        a.out = at<T>(act(c.out).goff); a.w = params + c.w; a.N = N; a.H = c.Hi;
        dg_done = launch_nb_tail_dgrad(a, at<T>(act(c.in).goff), st);
      }
      if (!dg_done && k < 3) dg_done = nb_mid(c, true, ACT_NONE);
      if (!dg_done) nb_dgrad(c, 0, ACT_NONE, -1);       // -> d(upsampled input)
      const int srci = nb_dec_src(k);
      const ActT& src = act(srci);
      launch_nb_upsample_bwd<T>(at<T>(act(nb.a_up[k]).goff), k > 0 ? at<T>(src.off) : nullptr, ACT_ELU, at<T>(src.goff), N,
                                src.H, src.W, src.C, nb.up[k], st);
    }
    const ConvT_& cmu = P.convs[nb.cmu]; const ConvT_& clv = P.convs[nb.clv];
    NbSampleBwdArgs b{};
    b.dz = at<T>(act(nb.a_z).goff);
    b.mu_y = at<T>(act(cmu.out).off); b.lv_y = at<T>(act(clv.out).off); b.eps_keep = at<float>(nb.eps_off);
    b.d_mu_y = at<T>(act(cmu.out).goff); b.d_lv_y = at<T>(act(clv.out).goff);
    b.kl_acc = acc + 1; b.klw_over_n = klw * inv_n;
    b.N = N; b.hw = nb.latent_hw * nb.latent_hw; b.z = P.d.z_dim;
    launch_nb_rsample_bwd<T>(b, st);
    launch_nb_loss_finalize(acc, inv_n, klw, out, st);
    side([&] { nb_param_grads(cmu); nb_param_grads(clv); });
    const int a4 = P.convs[nb.e[3]].out;
    nb_dgrad(cmu, 0, ACT_RELU, a4);
    nb_dgrad(clv, 1, ACT_RELU, a4);
    for (int i = 3; i >= 1; --i) {
      const ConvT_& c = P.convs[nb.e[i]];
      side([&] { nb_param_grads(c); });
      nb_dgrad(c, 0, ACT_RELU, P.convs[nb.e[i - 1]].out);
    }
    {
      const ConvT_& c = P.convs[nb.e[0]];
      if (nb_stem_ok()) launch_nb_stem_wgrad(x, at<T>(act(c.out).goff), grads + c.w, grads + c.bias, N, c.Hi, st);
      else nb_param_grads(c);
    }
    join();
  }

  // gradient-arena range [begin, end) owned by a backward phase
  static void phase_range(const Plan& P, int phase, int64_t& b, int64_t& e) {
    const int64_t deep = P.convs[P.enc[2].c1].w;             // encoder.layer3.0.conv1.weight
    const int64_t dec = P.convs[P.dstem].w;                  // decoder.conv1.weight
    if (phase == MMVAE_BWD_DECODER) { b = dec; e = P.n_params; }
    else if (phase == MMVAE_BWD_ENC_DEEP) { b = deep; e = dec; }
    else { b = 0; e = deep; }
  }
  void clear(int phase) {
    int64_t b, e;
    phase_range(P, phase, b, e);
    cudaMemsetAsync(grads + b, 0, sizeof(float) * size_t(e - b), st);
  }

  void backward(const float* d_mu, const float* d_lv, const float* d_enc, const float* d_recon, int phases_in) {
    const bool defer_join = (phases_in & MMVAE_BWD_DEFER_JOIN) != 0;
    const int phases = phases_in & MMVAE_BWD_ALL;
    if (phases & MMVAE_BWD_DECODER) {
      clear(MMVAE_BWD_DECODER);
      if (d_recon) {
        const ConvT_& t = P.convs[P.tail];
        const float* dr = d_recon;
        if (t.Co > 1) {      // NCHW -> NHWC
          launch_nchw_to_nhwc(d_recon, at<float>(P.drecon_off), P.d.batch, t.Co, t.Ho * t.Wo, st);
          dr = at<float>(P.drecon_off);
        }
        bn_bwd(dr, 1, -1, t, nullptr);
        // decoder.conv2.bias feeds a BatchNorm: its gradient is identically zero;
        // the arena range was cleared above.
        if (use_tail(t)) {
          TailArgs a{};
          a.in = at<__nv_bfloat16>(act(t.in).off); a.w = params + t.w; a.dy = at<__nv_bfloat16>(act(t.out).goff);
          a.dx = at<__nv_bfloat16>(act(t.in).goff); a.dw = grads + t.w; a.N = P.d.batch; a.H = t.Hi; a.W = t.Wi;
          if (fused_reduce(t.in)) fill_bwd_fused(t.in, a.bb);
          launch_tail_bwd(a, t.Ci, st);
        } else {
          wgrad(t);
          dgrad(t, 0);
        }
        for (int i = (int)P.dec.size() - 1; i >= 0; --i) block_bwd(P.dec[i]);
        const ConvT_& s = P.convs[P.dstem];
        bn_bwd(at<T>(act(P.a_dstem).goff), 0, P.a_dstem, s, nullptr);
        side([&] { wgrad(s); });
        dgrad(s, 0);
      }
      if (phases != MMVAE_BWD_ALL && !defer_join) join();   // a phase on its own hands complete gradients to the caller (all-reduce)
                                               // unless the caller fences its consumer itself (MMVAE_BWD_DEFER_JOIN); the
                                               // whole sweep in one call joins the auxiliary stream once, at the end
    }
    if (phases & MMVAE_BWD_ENC_DEEP) {
      clear(MMVAE_BWD_ENC_DEEP);
      HeadsBwdArgs h;
      memset(&h, 0, sizeof(h));
      h.dz_act = d_recon ? at<T>(act(P.a_z).goff) : nullptr;
      h.d_mu = d_mu; h.d_lv = d_lv; h.d_enc = d_enc;
      h.heads = at<float>(P.heads_off); h.pooled = at<float>(P.pooled_off);
      h.w_mu = params + P.w_mu; h.w_lv = P.w_lv >= 0 ? params + P.w_lv : nullptr;
      h.dheads = at<float>(P.dheads_off); h.dpool = at<float>(P.dpool_off);
      h.g_wmu = grads + P.w_mu; h.g_wlv = P.w_lv >= 0 ? grads + P.w_lv : nullptr;
      h.dfeat = at<T>(act(P.enc.back().out).goff);
      h.N = P.d.batch; h.hw = P.feat_hw; h.C = P.feat_c; h.z = P.d.z_dim;
      launch_heads_bwd<T>(h, st);
      side([&] { launch_heads_wgrad(h.dheads, h.pooled, h.g_wmu, h.w_lv ? h.g_wlv : nullptr, h.N, h.z, h.C, st); });
      block_bwd(P.enc[3]);
      block_bwd(P.enc[2]);
      if (phases != MMVAE_BWD_ALL && !defer_join) join();
    }
    if (phases & MMVAE_BWD_ENC_SHALLOW) {
      clear(MMVAE_BWD_ENC_SHALLOW);
      block_bwd(P.enc[1]);
      block_bwd(P.enc[0]);
      const ConvT_& s = P.convs[P.stem];
      if (stem_bwd_fused(P) && use_stem(s) && !fused_reduce(P.a_stem)) {
        // the stem BatchNorm's backward is one reduction pass; the weight-gradient kernel forms dY in its loader from
        // (gradient parts, ReLU output, raw conv output) and the reduction's coefficients: dY is never stored
        bn_bwd(at<T>(act(P.a_stem).goff), 0, P.a_stem, s, nullptr, true);
        const BnT& b = P.bns[s.bn];
        StemArgs a{};
        a.x = x; a.dw = grads + s.w; a.N = P.d.batch; a.S = s.Hi;
        a.g1 = at<__nv_bfloat16>(act(P.a_stem).goff);
        a.g2 = split_grad(P.a_stem) ? at<__nv_bfloat16>(act(P.a_stem).goff2) : nullptr;
        a.mask = at<__nv_bfloat16>(act(P.a_stem).off);
        a.yraw = at<__nv_bfloat16>(act(s.out).off);
        a.stat = at<float>(b.stat_off); a.bcoef = at<float>(b.bcoef_off);
        launch_stem_wgrad(a, s.Co, st);
      } else {
        bn_bwd(at<T>(act(P.a_stem).goff), 0, P.a_stem, s, nullptr);
        wgrad(s);
      }
      forked = forked || (aux != nullptr);     // earlier phases may have left un-joined work on the auxiliary stream
      join();
    }
  }
};

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace mmvae

namespace mmvae {
__global__ void selftest_fill_f32_kernel(float* out, long long n, unsigned int seed) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    unsigned int h = (unsigned int)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    out[i] = ((float)(h & 0xffff) - 32768.0f) * (1.0f / 32768.0f);
  }
}
__global__ void selftest_fill_kernel(__nv_bfloat16* out, long long n, unsigned int seed) {
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    unsigned int h = (unsigned int)i * 2654435761u ^ seed;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
    out[i] = __float2bfloat16_rn(((float)(h & 0xffff) - 32768.0f) * (1.0f / 32768.0f));
  }
}
template <typename T>
__global__ void selftest_cmp_kernel(const T* a, const T* ref, long long n, float* rep) {
  float d2 = 0.f, r2 = 0.f, mx = 0.f;
  for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
    float x = to_f(a[i]), y = to_f(ref[i]);
    float d = x - y;
    if (!(d == d)) d = 1e30f;                      // NaN counts as a huge error
    d2 += d * d; r2 += y * y; mx = fmaxf(mx, fabsf(d));
  }
  atomicAdd(rep + 0, d2); atomicAdd(rep + 1, r2);
  atomicMax(reinterpret_cast<int*>(rep + 2), __float_as_int(mx));
}
}  // namespace mmvae

using namespace mmvae;

extern "C" {

int mmvae_abi_version(void) { return MMVAE_ABI_VERSION; }
const char* mmvae_last_error(void) { return g_err; }

int mmvae_layout(const mmvae_desc* d, mmvae_layout_info* out) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (!out) return fail(MMVAE_ERR_BAD_ARG, "out is NULL");
  out->n_params = P.n_params; out->n_bn_buffers = P.n_bn_buffers;
  out->n_param_tensors = (int32_t)P.params.size(); out->n_bn = (int32_t)P.bns.size();
  out->workspace_bytes = (int64_t)P.ws_bytes; out->decoder_size = P.dec_size; out->crop = P.crop;
  out->train_flops = P.train_flops;
  return 0;
}

int mmvae_param_entry(const mmvae_desc* d, int32_t i, char* name, size_t name_cap, int64_t* offset,
                      int32_t* ndim, int32_t shape[4]) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (i < 0 || i >= (int)P.params.size()) return fail(MMVAE_ERR_BAD_ARG, "parameter index %d out of range", i);
  const ParamT& p = P.params[i];
  if (name && name_cap) { strncpy(name, p.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (offset) *offset = p.off;
  if (ndim) *ndim = p.ndim;
  if (shape) for (int k = 0; k < 4; ++k) shape[k] = p.shape[k];
  return 0;
}

int mmvae_bn_entry(const mmvae_desc* d, int32_t i, char* prefix, size_t prefix_cap, int32_t* channels,
                   int64_t* buffer_offset) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (i < 0 || i >= (int)P.bns.size()) return fail(MMVAE_ERR_BAD_ARG, "BatchNorm index %d out of range", i);
  const BnT& b = P.bns[i];
  if (prefix && prefix_cap) { strncpy(prefix, b.prefix.c_str(), prefix_cap - 1); prefix[prefix_cap - 1] = 0; }
  if (channels) *channels = b.C;
  if (buffer_offset) *buffer_offset = b.rm;
  return 0;
}

int mmvae_workspace_tensor(const mmvae_desc* d, const char* name, int64_t* byte_offset, int32_t dims[4]) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (!name) return fail(MMVAE_ERR_BAD_ARG, "name is NULL");
  std::string nm(name);
  bool grad = false, grad2 = false;
  if (nm.size() > 6 && nm.compare(nm.size() - 6, 6, ".grad2") == 0) { grad2 = true; nm.resize(nm.size() - 6); }
  else if (nm.size() > 5 && nm.compare(nm.size() - 5, 5, ".grad") == 0) { grad = true; nm.resize(nm.size() - 5); }
  const ActT* a = P.find_act(nm.c_str());
  if (!a) return fail(MMVAE_ERR_BAD_ARG, "no workspace tensor named '%s'", name);
  if (grad2 && a->goff2 == 0) return fail(MMVAE_ERR_BAD_ARG, "'%s' has no second gradient buffer", nm.c_str());
  if (grad && P.stem >= 0 && a == &P.acts[P.convs[P.stem].out] && stem_bwd_fused(P))
    return fail(MMVAE_ERR_BAD_ARG, "'%s' is not materialised: the stem weight gradient forms it in its loader", name);
  if (byte_offset) *byte_offset = (int64_t)(grad2 ? a->goff2 : (grad ? a->goff : a->off));
  if (dims) { dims[0] = P.d.batch; dims[1] = a->H; dims[2] = a->W; dims[3] = a->C; }
  return 0;
}

#define MMVAE_COMMON_CHECKS()                                                                   \
  Plan P;                                                                                       \
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());                        \
  if (int rc = check_device()) return rc;                                                       \
  if (!workspace || workspace_bytes < P.ws_bytes)                                               \
    return fail(MMVAE_ERR_BAD_ARG, "workspace too small: %zu < %zu bytes", workspace_bytes, P.ws_bytes); \
  if (!aligned16(workspace)) return fail(MMVAE_ERR_BAD_ARG, "workspace must be 16-byte aligned");

int mmvae_forward(const mmvae_desc* d, const float* x, const float* params, float* bn_buffers,
                  int64_t* bn_counters, const float* eps, uint64_t seed, uint64_t offset, float* eps_out,
                  uint64_t* rng_state, void* workspace, size_t workspace_bytes, float* mu, float* logvar,
                  float* encoding, float* recon, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.arch == MMVAE_ARCH_NOTEBOOK) {
    if (!x || !params) return fail(MMVAE_ERR_BAD_ARG, "x/params must be non-NULL");
    if (!aligned16(params) || !aligned16(x)) return fail(MMVAE_ERR_BAD_ARG, "x and params must be 16-byte aligned");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (P.d.precision == MMVAE_PREC_FP32) {
      Exec<float> E{P, (char*)workspace, params, nullptr, nullptr, nullptr, st, x};
      E.nb_forward(eps, seed, offset, rng_state, eps_out, mu, logvar, encoding, recon);
    } else {
      Exec<__nv_bfloat16> E{P, (char*)workspace, params, nullptr, nullptr, nullptr, st, x};
      E.pack_weights();
      E.nb_forward(eps, seed, offset, rng_state, eps_out, mu, logvar, encoding, recon);
    }
    return check_launches("mmvae_forward");
  }
  if (!x || !params || !mu || !encoding || !recon) return fail(MMVAE_ERR_BAD_ARG, "x/params/mu/encoding/recon must be non-NULL");
  if (d->require_rsample && !logvar) return fail(MMVAE_ERR_BAD_ARG, "logvar must be non-NULL when require_rsample");
  if (!bn_buffers) return fail(MMVAE_ERR_BAD_ARG, "bn_buffers must be non-NULL");
  if (!aligned16(params) || !aligned16(x)) return fail(MMVAE_ERR_BAD_ARG, "x and params must be 16-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.d.precision == MMVAE_PREC_FP32) {
    Exec<float> E{P, (char*)workspace, params, nullptr, bn_buffers, (long long*)bn_counters, st, x};
    if (!P.d.training) { E.counters = nullptr; }
    E.encode(eps, seed, offset, rng_state, eps_out, mu, logvar, encoding);
    E.decode(recon);
  } else {
    Exec<__nv_bfloat16> E{P, (char*)workspace, params, nullptr, bn_buffers, (long long*)bn_counters, st, x};
    if (!P.d.training) { E.counters = nullptr; }
    E.use_aux();
    E.clear_bn_acc();
    if (E.use_stem(P.convs[P.stem])) E.side([&] { E.pack_weights(); });   // the stem kernel does not read packed weights
    else E.pack_weights();
    E.encode(eps, seed, offset, rng_state, eps_out, mu, logvar, encoding);
    E.decode(recon);
  }
  return check_launches("mmvae_forward");
}

int mmvae_decode(const mmvae_desc* d, const float* encoding, const float* params, float* bn_buffers,
                 int64_t* bn_counters, void* workspace, size_t workspace_bytes, float* recon, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.arch == MMVAE_ARCH_NOTEBOOK) {
    if (!encoding || !params || !recon) return fail(MMVAE_ERR_BAD_ARG, "encoding/params/recon must be non-NULL");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // encoding arrives as fp32 NCHW [N,z,h,h]; the decoder reads NHWC storage: reuse rsample with eps = 0 semantics is
    // not possible without mu, so transpose through the export kernel's inverse -- a 1-pixel-per-thread copy
    const int hw = P.nb.latent_hw * P.nb.latent_hw;
    if (P.d.precision == MMVAE_PREC_FP32) {
      Exec<float> E{P, (char*)workspace, params, nullptr, nullptr, nullptr, st, nullptr};
      launch_nb_import_nchw<float>(encoding, E.at<float>(P.acts[P.nb.a_z].off), P.d.batch, hw, P.d.z_dim, st);
      E.nb_decode(recon);
    } else {
      Exec<__nv_bfloat16> E{P, (char*)workspace, params, nullptr, nullptr, nullptr, st, nullptr};
      E.pack_weights();
      launch_nb_import_nchw<__nv_bfloat16>(encoding, E.at<__nv_bfloat16>(P.acts[P.nb.a_z].off), P.d.batch, hw, P.d.z_dim, st);
      E.nb_decode(recon);
    }
    return check_launches("mmvae_decode");
  }
  if (!encoding || !params || !recon || !bn_buffers) return fail(MMVAE_ERR_BAD_ARG, "encoding/params/bn_buffers/recon must be non-NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long nz = (long long)P.d.batch * P.d.z_dim;
  if (P.d.precision == MMVAE_PREC_FP32) {
    Exec<float> E{P, (char*)workspace, params, nullptr, bn_buffers, (long long*)bn_counters, st, nullptr};
    if (!P.d.training) E.counters = nullptr;
    launch_cast_latent<float>(encoding, E.at<float>(P.acts[P.a_z].off), nz, st);
    E.decode(recon);
  } else {
    Exec<__nv_bfloat16> E{P, (char*)workspace, params, nullptr, bn_buffers, (long long*)bn_counters, st, nullptr};
    if (!P.d.training) E.counters = nullptr;
    E.use_aux();
    E.clear_bn_acc();
    E.pack_weights();
    launch_cast_latent<__nv_bfloat16>(encoding, E.at<__nv_bfloat16>(P.acts[P.a_z].off), nz, st);
    E.decode(recon);
  }
  return check_launches("mmvae_decode");
}

int mmvae_backward(const mmvae_desc* d, const float* x, const float* params, void* workspace,
                   size_t workspace_bytes, const float* d_mu, const float* d_logvar, const float* d_encoding,
                   const float* d_recon, float* grads, int32_t phases, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.arch != MMVAE_ARCH_RESNET) return fail(MMVAE_ERR_BAD_DESC, "mmvae_backward: use mmvae_nb_loss_backward for the notebook variant");
  if (!x || !params || !grads) return fail(MMVAE_ERR_BAD_ARG, "x/params/grads must be non-NULL");
  if ((phases & MMVAE_BWD_ALL) == 0 || phases > (MMVAE_BWD_ALL | MMVAE_BWD_DEFER_JOIN))
    return fail(MMVAE_ERR_BAD_ARG, "phases must be a non-empty MMVAE_BWD_* mask");
  if ((phases & MMVAE_BWD_DEFER_JOIN) && (phases & MMVAE_BWD_ALL) != MMVAE_BWD_DECODER && (phases & MMVAE_BWD_ALL) != MMVAE_BWD_ENC_DEEP)
    return fail(MMVAE_ERR_BAD_ARG, "MMVAE_BWD_DEFER_JOIN goes with a single DECODER or ENC_DEEP phase");
  if (!d->training) return fail(MMVAE_ERR_BAD_DESC, "mmvae_backward needs a training-mode forward (batch statistics)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.d.precision == MMVAE_PREC_FP32) {
    Exec<float> E{P, (char*)workspace, params, grads, nullptr, nullptr, st, x};
    E.backward(d_mu, d_logvar, d_encoding, d_recon, phases);
  } else {
    Exec<__nv_bfloat16> E{P, (char*)workspace, params, grads, nullptr, nullptr, st, x};
    E.use_aux();
    E.backward(d_mu, d_logvar, d_encoding, d_recon, phases);
  }
  return check_launches("mmvae_backward");
}

int mmvae_nb_loss_backward(const mmvae_desc* d, const float* x, const int64_t* target, const float* params, void* workspace,
                           size_t workspace_bytes, float kl_weight, float* out, float* grads, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.arch != MMVAE_ARCH_NOTEBOOK) return fail(MMVAE_ERR_BAD_DESC, "mmvae_nb_loss_backward needs arch = MMVAE_ARCH_NOTEBOOK");
  if (!x || !target || !params || !out || !grads) return fail(MMVAE_ERR_BAD_ARG, "x/target/params/out/grads must be non-NULL");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (P.d.precision == MMVAE_PREC_FP32) {
    Exec<float> E{P, (char*)workspace, params, grads, nullptr, nullptr, st, x};
    E.nb_loss_backward(reinterpret_cast<const long long*>(target), kl_weight, out);
  } else {
    Exec<__nv_bfloat16> E{P, (char*)workspace, params, grads, nullptr, nullptr, st, x};
    if (!(P.d.flags & MMVAE_FLAG_FORCE_SIMT) && !aux_disabled()) E.aux = aux_pool();
    E.nb_loss_backward(reinterpret_cast<const long long*>(target), kl_weight, out);
  }
  return check_launches("mmvae_nb_loss_backward");
}

int mmvae_nb_bench_tail(const mmvae_desc* d, int32_t which, const float* params, const int64_t* target, void* workspace,
                        size_t workspace_bytes, float* grads_scratch, int64_t* algo_bytes, int64_t* algo_flops, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.arch != MMVAE_ARCH_NOTEBOOK || P.d.precision != MMVAE_PREC_BF16) return fail(MMVAE_ERR_BAD_DESC, "notebook variant, bf16 only");
  if (!params || !grads_scratch || which < 0 || which > 2 || (which == 0 && !target)) return fail(MMVAE_ERR_BAD_ARG, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Exec<__nv_bfloat16> E{P, (char*)workspace, params, grads_scratch, nullptr, nullptr, st, nullptr};
  if (!E.nb_tail_ok()) return fail(MMVAE_ERR_BAD_DESC, "decoder.conv4 is not covered by the dedicated kernels at this shape");
  const ConvT_& c = P.convs[P.nb.dc[3]];
  const int64_t px = (int64_t)P.d.batch * c.Ho * c.Wo;
  if (algo_flops) *algo_flops = 2 * px * c.Co * c.Ci * 9;
  const int64_t b_in = px * c.Ci * 2, b_g = px * c.Co * 2, b_w = (int64_t)c.Co * c.Ci * 9 * 4;
  NbTailArgs a{};
  a.x = E.at<__nv_bfloat16>(P.acts[c.in].off); a.w = params + c.w; a.bias = params + c.bias;
  a.out = E.at<__nv_bfloat16>(P.acts[c.out].goff); a.N = P.d.batch; a.H = c.Hi;
  bool ok = false;
  if (which == 0) {
    a.target = reinterpret_cast<const long long*>(target); a.ce_acc = E.at<double>(P.nb.acc_off); a.scale = 1.0f / (float)P.d.batch;
    ok = launch_nb_tail_fwd(a, st);
    if (algo_bytes) *algo_bytes = b_in + b_g + px * 8 + b_w;
  } else if (which == 1) {
    ok = launch_nb_tail_dgrad(a, E.at<__nv_bfloat16>(P.acts[c.in].goff), st);
    if (algo_bytes) *algo_bytes = b_g + b_in + b_w;
  } else {
    ok = launch_nb_tail_wgrad(a, grads_scratch + c.w, grads_scratch + c.bias, st);
    if (algo_bytes) *algo_bytes = b_g + b_in + b_w;
  }
  if (!ok) return fail(MMVAE_ERR_CUDA, "TMA descriptor could not be encoded");
  return check_launches("mmvae_nb_bench_tail");
}

int mmvae_aux_fence(void* stream) {
  if (int rc = check_device()) return rc;
  AuxPool* p = aux_pool();
  if (!p) return fail(MMVAE_ERR_CUDA, "auxiliary streams unavailable");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (cudaStream_t a : {p->s, p->s2}) {
    cudaEvent_t e = p->ev[p->next]; p->next = (p->next + 1) & 31;
    cudaEventRecord(e, a);
    cudaStreamWaitEvent(st, e, 0);
  }
  return check_launches("mmvae_aux_fence");
}

int mmvae_backward_range(const mmvae_desc* d, int32_t phase, int64_t* begin, int64_t* end) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (phase != MMVAE_BWD_DECODER && phase != MMVAE_BWD_ENC_DEEP && phase != MMVAE_BWD_ENC_SHALLOW)
    return fail(MMVAE_ERR_BAD_ARG, "phase must be a single MMVAE_BWD_* value");
  if (P.d.arch != MMVAE_ARCH_RESNET) return fail(MMVAE_ERR_BAD_DESC, "mmvae_backward_range: model.py VAE only");
  int64_t b, e;
  Exec<float>::phase_range(P, phase, b, e);
  if (begin) *begin = b;
  if (end) *end = e;
  return 0;
}

size_t mmvae_loss_scratch_bytes(void) { return loss_scratch_bytes(); }

static int loss_args_check(const mmvae_loss_args* a, LossArgs& L) {
  if (!a || a->struct_size != (int32_t)sizeof(mmvae_loss_args)) return fail(MMVAE_ERR_BAD_ARG, "mmvae_loss_args.struct_size mismatch");
  if (a->kind != MMVAE_LOSS_GAUSSIAN && a->kind != MMVAE_LOSS_CATEGORICAL) return fail(MMVAE_ERR_BAD_ARG, "unknown loss kind");
  if (a->batch < 1 || a->channels < 1 || a->height < 1 || a->width < 1) return fail(MMVAE_ERR_BAD_ARG, "bad loss shape");
  if (a->kind == MMVAE_LOSS_GAUSSIAN && !(a->sigma > 0.f)) return fail(MMVAE_ERR_BAD_ARG, "sigma_decoder must be > 0 (main.py:91-93)");
  L.kind = a->kind; L.nll = a->nll; L.kl = a->kl; L.sigma = a->sigma;
  L.N = a->batch; L.C = a->channels; L.H = a->height; L.W = a->width; L.z = a->z_dim;
  return 0;
}

int mmvae_loss_forward(const mmvae_loss_args* a, const float* recon, const void* target, const float* ce_weight,
                       const float* mu, const float* logvar, float* out, void* scratch, void* stream) {
  LossArgs L;
  if (int rc = loss_args_check(a, L)) return rc;
  if (int rc = check_device()) return rc;
  if (!out || !scratch) return fail(MMVAE_ERR_BAD_ARG, "out/scratch must be non-NULL");
  if ((recon == nullptr) != (target == nullptr)) return fail(MMVAE_ERR_BAD_ARG, "recon and target must both be given or both be NULL (KL only)");
  if (!aligned16(recon) || !aligned16(target)) return fail(MMVAE_ERR_BAD_ARG, "recon and target must be 16-byte aligned");
  launch_loss_fwd(L, recon, target, ce_weight, mu, logvar, a->kl_dev, out, scratch, reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_loss_forward");
}

int mmvae_loss_backward(const mmvae_loss_args* a, const float* recon, const void* target, const float* ce_weight,
                        const float* mu, const float* logvar, const float* grad_out, float* d_recon, float* d_mu,
                        float* d_logvar, void* stream) {
  LossArgs L;
  if (int rc = loss_args_check(a, L)) return rc;
  if (int rc = check_device()) return rc;
  if (!grad_out) return fail(MMVAE_ERR_BAD_ARG, "grad_out must be non-NULL");
  if (d_recon && (!recon || !target)) return fail(MMVAE_ERR_BAD_ARG, "d_recon needs recon and target");
  if (!aligned16(recon) || !aligned16(target) || (d_recon && !aligned16(d_recon)))
    return fail(MMVAE_ERR_BAD_ARG, "recon, target and d_recon must be 16-byte aligned");
  launch_loss_bwd(L, recon, target, ce_weight, mu, logvar, grad_out, a->kl_dev, d_recon, d_mu, d_logvar,
                  reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_loss_backward");
}

int64_t mmvae_launch_count(void) { return (int64_t)g_launches.load(); }
void mmvae_debug_set_trace(void* device_buffer) { g_trace_buf = reinterpret_cast<unsigned long long*>(device_buffer); }

int mmvae_prepare_input(const uint8_t* labels, int64_t n, float data_mean, float data_std, float* x, int64_t* target,
                        void* stream) {
  if (int rc = check_device()) return rc;
  if (!labels || !x || n < 0 || !(data_std > 0.f)) return fail(MMVAE_ERR_BAD_ARG, "bad arguments");
  launch_prepare_input(labels, n, data_mean, data_std, x, (long long*)target, reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_prepare_input");
}

int mmvae_philox_normal(uint64_t seed, uint64_t offset, const uint64_t* rng_state, uint64_t stream_id, int64_t n,
                        float* out, void* stream) {
  if (int rc = check_device()) return rc;
  if (!out || n < 0) return fail(MMVAE_ERR_BAD_ARG, "bad arguments");
  launch_philox_normal(seed, offset, reinterpret_cast<const unsigned long long*>(rng_state), stream_id, n, out,
                       reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_philox_normal");
}

size_t mmvae_mmd_scratch_bytes(int32_t n) { return mmd_scratch_bytes(n > 0 ? n : 1); }

int mmvae_mmd(const float* true_samples, const float* encoding, int32_t n, int32_t z_dim, float* out, void* scratch,
              void* stream) {
  if (int rc = check_device()) return rc;
  if (!true_samples || !encoding || !out || !scratch || n < 1 || z_dim < 1 || z_dim > 1024)
    return fail(MMVAE_ERR_BAD_ARG, "mmvae_mmd: bad arguments (z_dim <= 1024)");
  if (!aligned16(true_samples) || !aligned16(encoding)) return fail(MMVAE_ERR_BAD_ARG, "mmvae_mmd: inputs must be 16-byte aligned");
  launch_mmd(true_samples, encoding, n, z_dim, out, scratch, reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_mmd");
}

int mmvae_adam_step(int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float lr,
                    float beta1, float beta2, float eps, float weight_decay, int64_t step, const int64_t* step_dev,
                    float grad_scale, void* stream) {
  if (int rc = check_device()) return rc;
  if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || (!step_dev && step < 1))
    return fail(MMVAE_ERR_BAD_ARG, "bad arguments");
  launch_adam(n, params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step > 0 ? step : 1,
              reinterpret_cast<const long long*>(step_dev), grad_scale, reinterpret_cast<cudaStream_t>(stream));
  return check_launches("mmvae_adam_step");
}


// ------------------------------------------------------------------------------------------------
// self-test of the tcgen05 kernels against the SIMT kernels on identical inputs
// ------------------------------------------------------------------------------------------------

int mmvae_selftest_tc(const mmvae_desc* d, const float* params, void* workspace, size_t workspace_bytes,
                      float* grads_a, float* grads_b, float* report, int32_t report_cap, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.precision != MMVAE_PREC_BF16 || (P.d.flags & MMVAE_FLAG_FORCE_SIMT) || !P.d.training)
    return fail(MMVAE_ERR_BAD_DESC, "mmvae_selftest_tc needs a training-mode bf16 desc without MMVAE_FLAG_FORCE_SIMT");
  if (!params || !grads_a || !grads_b || !report) return fail(MMVAE_ERR_BAD_ARG, "NULL argument");
  const int nconv = (int)P.convs.size();
  if (report_cap < nconv * 4 * 4) return fail(MMVAE_ERR_BAD_ARG, "report needs %d floats", nconv * 16);
  typedef __nv_bfloat16 T;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Exec<T> E{P, (char*)workspace, params, grads_a, nullptr, nullptr, st, nullptr};
  cudaMemsetAsync(report, 0, sizeof(float) * nconv * 16, st);
  cudaMemsetAsync(grads_a, 0, sizeof(float) * P.n_params, st);
  cudaMemsetAsync(grads_b, 0, sizeof(float) * P.n_params, st);
  E.pack_weights();
  const int N = P.d.batch;
  for (int ci = 0; ci < nconv; ++ci) {
    const ConvT_& c = P.convs[ci];
    float* rep = report + ci * 16;
    if (E.use_stem(c)) {
      // stem: dedicated kernels vs the generic SIMT gather-convolution, on a random fp32 input
      const ActT& ao = P.acts[c.out];
      const long long n_out = (long long)N * ao.H * ao.W * ao.C;
      float* xr = E.at<float>(P.drecon_off);
      E.x = xr;
      selftest_fill_f32_kernel<<<296, 256, 0, st>>>(xr, (long long)N * c.Hi * c.Wi, 0x77u);
      GConvParams g;
      geom_fprop(c, N, g);
      g.in = xr; g.in_nchw_f32 = 1; g.w = params + c.w;
      const BnT& b = P.bns[c.bn];
      g.partials = E.at<float>(b.part_off);
      BnFinalizeArgs f;
      memset(&f, 0, sizeof(f));
      f.partials = g.partials; f.C = b.C; f.m = b.m; f.gamma = params + b.gamma; f.beta = params + b.beta;
      f.stat = E.at<float>(b.stat_off); f.coef = E.at<float>(b.coef_off); f.training = 1;
      g.out = E.at<T>(ao.off);
      f.sl = launch_gconv_simt<T>(g, st);
      launch_bn_finalize(f, st);
      cudaMemcpyAsync(E.at<float>(b.bcoef_off), E.at<float>(b.stat_off), sizeof(float) * 2 * b.C, cudaMemcpyDeviceToDevice, st);
      StemArgs sa{};
      sa.x = xr; sa.w = params + c.w; sa.y = E.at<T>(ao.goff); sa.N = N; sa.S = c.Hi;
      E.clear_bn_acc();
      sa.bn = E.bn_fused(b);
      launch_stem_fwd(sa, c.Co, st);
      selftest_cmp_kernel<T><<<148, 256, 0, st>>>(E.at<T>(ao.goff), E.at<T>(ao.off), n_out, rep + 0);
      selftest_cmp_kernel<float><<<1, 256, 0, st>>>(E.at<float>(b.stat_off), E.at<float>(b.bcoef_off), 2 * b.C, rep + 4);
      selftest_fill_kernel<<<296, 256, 0, st>>>(E.at<T>(ao.goff), n_out, 0x9876u + ci);
      WGradParams w; memset(&w, 0, sizeof(w));
      w.in = xr; w.in_nchw_f32 = 1; w.dout = E.at<T>(ao.goff);
      w.N = g.N; w.Hi = g.Hi; w.Wi = g.Wi; w.Ci = g.Ci; w.Ho = g.Ho; w.Wo = g.Wo; w.Co = g.Co;
      w.Hg = g.Hg; w.Wg = g.Wg; w.M = g.M; w.os = g.os; w.is = g.is; w.w_sci = g.w_sci; w.w_sco = g.w_sco;
      w.nvar = g.nvar; w.var[0] = g.var[0];
      w.dw = grads_a + c.w; launch_wgrad_simt<T>(w, st);
      sa.dy = E.at<T>(ao.goff); sa.dw = grads_b + c.w;
      launch_stem_wgrad(sa, c.Co, st);
      selftest_cmp_kernel<float><<<148, 256, 0, st>>>(grads_b + c.w, grads_a + c.w, (long long)c.Ci * c.Co * c.k * c.k, rep + 8);
      continue;
    }
    if (E.use_tail(c)) {
      const ActT& ai = P.acts[c.in]; const ActT& ao = P.acts[c.out];
      const long long n_in = (long long)N * ai.H * ai.W * ai.C, n_out = (long long)N * ao.H * ao.W * ao.C;
      selftest_fill_kernel<<<296, 256, 0, st>>>(E.at<T>(ai.off), n_in, 0x1234u + ci);
      GConvParams g;
      geom_fprop(c, N, g);
      g.in = E.at<T>(ai.off); g.w = params + c.w; g.bias = c.bias >= 0 ? params + c.bias : nullptr;
      const BnT& b = P.bns[c.bn];
      g.partials = E.at<float>(b.part_off);
      BnFinalizeArgs f;
      memset(&f, 0, sizeof(f));
      f.partials = g.partials; f.C = b.C; f.m = b.m; f.gamma = params + b.gamma; f.beta = params + b.beta;
      f.stat = E.at<float>(b.stat_off); f.coef = E.at<float>(b.coef_off); f.training = 1;
      g.out = E.at<T>(ao.off);
      f.sl = launch_gconv_simt<T>(g, st);
      launch_bn_finalize(f, st);
      cudaMemcpyAsync(E.at<float>(b.bcoef_off), E.at<float>(b.stat_off), sizeof(float) * 2 * b.C, cudaMemcpyDeviceToDevice, st);
      TailArgs ta{};
      ta.in = E.at<T>(ai.off); ta.w = params + c.w; ta.bias = g.bias; ta.y = E.at<T>(ao.goff);
      ta.N = N; ta.H = c.Hi; ta.W = c.Wi;
      E.clear_bn_acc();
      ta.bn = E.bn_fused(b);
      launch_tail_fwd(ta, c.Ci, st);
      selftest_cmp_kernel<T><<<148, 256, 0, st>>>(E.at<T>(ao.goff), E.at<T>(ao.off), n_out, rep + 0);
      selftest_cmp_kernel<float><<<1, 256, 0, st>>>(E.at<float>(b.stat_off), E.at<float>(b.bcoef_off), 2 * b.C, rep + 4);
      // backward: dY random; SIMT wgrad + dgrad vs the fused tail_bwd
      selftest_fill_kernel<<<296, 256, 0, st>>>(E.at<T>(ao.goff), n_out, 0x9876u + ci);
      WGradParams w; memset(&w, 0, sizeof(w));
      w.in = E.at<T>(ai.off); w.dout = E.at<T>(ao.goff);
      w.N = g.N; w.Hi = g.Hi; w.Wi = g.Wi; w.Ci = g.Ci; w.Ho = g.Ho; w.Wo = g.Wo; w.Co = g.Co;
      w.Hg = g.Hg; w.Wg = g.Wg; w.M = g.M; w.os = g.os; w.is = g.is; w.w_sci = g.w_sci; w.w_sco = g.w_sco;
      w.nvar = g.nvar; w.var[0] = g.var[0];
      w.dw = grads_a + c.w; launch_wgrad_simt<T>(w, st);
      GConvParams gd;
      geom_dgrad(c, N, gd);
      gd.in = E.at<T>(ao.goff); gd.w = params + c.w; gd.out = E.at<T>(ai.goff);
      launch_gconv_simt<T>(gd, st);
      // the fused kernel's dX goes to a scratch tensor of the same shape (the last block's raw conv2 output)
      T* dx_scratch = E.at<T>(P.acts[P.convs[P.dec.back().c2].out].off);
      ta.dy = E.at<T>(ao.goff); ta.dw = grads_b + c.w; ta.dx = dx_scratch;
      launch_tail_bwd(ta, c.Ci, st);
      selftest_cmp_kernel<float><<<148, 256, 0, st>>>(grads_b + c.w, grads_a + c.w, (long long)c.Ci * c.Co * c.k * c.k, rep + 8);
      selftest_cmp_kernel<T><<<148, 256, 0, st>>>(dx_scratch, E.at<T>(ai.goff), n_in, rep + 12);
      continue;
    }
    if (c.in < 0 || c.wp_chunks[DIR_FPROP] <= 0) continue;        // covered by neither path
    const ActT& ai = P.acts[c.in]; const ActT& ao = P.acts[c.out];
    const long long n_in = (long long)N * ai.H * ai.W * ai.C, n_out = (long long)N * ao.H * ao.W * ao.C;
    // ---- fprop (+ BatchNorm statistics) ----
    selftest_fill_kernel<<<296, 256, 0, st>>>(E.at<T>(ai.off), n_in, 0x1234u + ci);
    GConvParams g;
    geom_fprop(c, N, g);
    g.in = E.at<T>(ai.off); g.w = params + c.w;
    g.wpack = (char*)workspace + c.wp_off[DIR_FPROP];
    const BnT& b = P.bns[c.bn];
    g.partials = E.at<float>(b.part_off);
    g.part_counts = E.at<float>(b.pcnt_off);
    BnFinalizeArgs f;
    memset(&f, 0, sizeof(f));
    f.partials = g.partials; f.C = b.C; f.m = b.m; f.gamma = params + b.gamma; f.beta = params + b.beta;
    f.stat = E.at<float>(b.stat_off); f.coef = E.at<float>(b.coef_off); f.training = 1;
    g.out = E.at<T>(ao.off);
    f.sl = launch_gconv_simt<T>(g, st);
    launch_bn_finalize(f, st);
    cudaMemcpyAsync(E.at<float>(b.bcoef_off), E.at<float>(b.stat_off), sizeof(float) * 2 * b.C, cudaMemcpyDeviceToDevice, st);
    g.out = E.at<T>(ao.goff);
    g.partials = nullptr;
    E.clear_bn_acc();
    g.bn = E.bn_fused(b);
    launch_gconv_tc(g, st);
    selftest_cmp_kernel<T><<<148, 256, 0, st>>>(E.at<T>(ao.goff), E.at<T>(ao.off), n_out, rep + 0);
    selftest_cmp_kernel<float><<<1, 256, 0, st>>>(E.at<float>(b.stat_off), E.at<float>(b.bcoef_off), 2 * b.C, rep + 4);
    // ---- wgrad: in = act(in), dY = random ----
    selftest_fill_kernel<<<296, 256, 0, st>>>(E.at<T>(ao.goff), n_out, 0x9876u + ci);
    E.grads = grads_a;
    {
      GConvParams gg; geom_fprop(c, N, gg);
      WGradParams w; memset(&w, 0, sizeof(w));
      w.in = E.at<T>(ai.off); w.dout = E.at<T>(ao.goff);
      w.N = gg.N; w.Hi = gg.Hi; w.Wi = gg.Wi; w.Ci = gg.Ci; w.Ho = gg.Ho; w.Wo = gg.Wo; w.Co = gg.Co;
      w.Hg = gg.Hg; w.Wg = gg.Wg; w.M = gg.M; w.os = gg.os; w.is = gg.is; w.w_sci = gg.w_sci; w.w_sco = gg.w_sco;
      w.nvar = gg.nvar; w.conv_class = gg.conv_class;
      for (int i = 0; i < gg.nvar; ++i) w.var[i] = gg.var[i];
      w.dw = grads_a + c.w; launch_wgrad_simt<T>(w, st);
      w.dw = grads_b + c.w; launch_wgrad_tc(w, st);
      long long nw = (long long)c.Ci * c.Co * c.k * c.k;
      selftest_cmp_kernel<float><<<148, 256, 0, st>>>(grads_b + c.w, grads_a + c.w, nw, rep + 8);
    }
    // ---- dgrad: dY = act(out).grad (random) -> dX ----
    if (c.wp_chunks[DIR_DGRAD] > 0) {
      GConvParams gd;
      geom_dgrad(c, N, gd);
      gd.in = E.at<T>(ao.goff); gd.w = params + c.w;
      gd.wpack = (char*)workspace + c.wp_off[DIR_DGRAD];
      cudaMemsetAsync(E.at<T>(ai.off), 0, n_in * sizeof(T), st);
      cudaMemsetAsync(E.at<T>(ai.goff), 0, n_in * sizeof(T), st);
      gd.out = E.at<T>(ai.off); launch_gconv_simt<T>(gd, st);
      gd.out = E.at<T>(ai.goff); launch_gconv_tc(gd, st);
      selftest_cmp_kernel<T><<<148, 256, 0, st>>>(E.at<T>(ai.goff), E.at<T>(ai.off), n_in, rep + 12);
    }
  }
  return check_launches("mmvae_selftest_tc");
}

int mmvae_bench_conv(const mmvae_desc* d, int32_t conv_index, int32_t dir, const float* params, void* workspace,
                     size_t workspace_bytes, float* grads_scratch, int64_t* algo_bytes, int64_t* algo_flops, void* stream) {
  MMVAE_COMMON_CHECKS();
  if (P.d.precision != MMVAE_PREC_BF16 || (P.d.flags & MMVAE_FLAG_FORCE_SIMT) || !P.d.training)
    return fail(MMVAE_ERR_BAD_DESC, "mmvae_bench_conv needs a training-mode bf16 desc without MMVAE_FLAG_FORCE_SIMT");
  if (conv_index < 0 || conv_index >= (int)P.convs.size()) return fail(MMVAE_ERR_BAD_ARG, "conv index out of range");
  if (dir < 0 || dir > 2 || !params) return fail(MMVAE_ERR_BAD_ARG, "bad arguments");
  typedef __nv_bfloat16 T;
  const ConvT_& c = P.convs[conv_index];
  if (c.in < 0 || c.wp_chunks[DIR_FPROP] <= 0) return fail(MMVAE_ERR_BAD_ARG, "conv %s is not on the tcgen05 path", c.name.c_str());
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  Exec<T> E{P, (char*)workspace, params, grads_scratch, nullptr, nullptr, st, nullptr};
  const ActT& ai = P.acts[c.in]; const ActT& ao = P.acts[c.out];
  const long long n_in = (long long)P.d.batch * ai.H * ai.W * ai.C, n_out = (long long)P.d.batch * ao.H * ao.W * ao.C;
  const long long nw = (long long)c.Ci * c.Co * c.k * c.k;
  // every conv / transposed conv of the model does batch * (input pixels or output pixels) * Ci*Co*k*k / s^2 MACs
  const long long macs = (c.kind == CONV) ? n_out / ao.C * c.Co * (long long)c.Ci * c.k * c.k
                                          : n_in / ai.C * c.Ci * (long long)c.Co * c.k * c.k;
  if (algo_flops) *algo_flops = 2 * macs;
  if (dir == 0) {
    GConvParams g;
    geom_fprop(c, P.d.batch, g);
    g.in = E.at<T>(ai.off); g.out = E.at<T>(ao.off); g.w = params + c.w;
    g.wpack = (char*)workspace + c.wp_off[DIR_FPROP];
    g.bn = E.bn_fused(P.bns[c.bn]);          // statistics accumulate as in the step; running buffers are not touched
    launch_gconv_tc(g, st);
    if (algo_bytes) *algo_bytes = 2 * (n_in + n_out + nw);
  } else if (dir == 1) {
    if (c.wp_chunks[DIR_DGRAD] <= 0) return fail(MMVAE_ERR_BAD_ARG, "conv %s has no data gradient", c.name.c_str());
    GConvParams g;
    geom_dgrad(c, P.d.batch, g);
    g.in = E.at<T>(ao.goff); g.out = E.at<T>(ai.goff); g.w = params + c.w;
    g.wpack = (char*)workspace + c.wp_off[DIR_DGRAD];
    launch_gconv_tc(g, st);
    if (algo_bytes) *algo_bytes = 2 * (n_in + n_out + nw);
  } else {
    if (!grads_scratch) return fail(MMVAE_ERR_BAD_ARG, "wgrad needs grads_scratch");
    E.wgrad(c);
    if (algo_bytes) *algo_bytes = 2 * (n_in + n_out) + 4 * nw;
  }
  return check_launches("mmvae_bench_conv");
}

int mmvae_conv_entry(const mmvae_desc* d, int32_t i, char* name, size_t name_cap, int32_t shape[8]) {
  Plan P;
  if (!P.build(d)) return fail(MMVAE_ERR_BAD_DESC, "%s", P.err.c_str());
  if (i < 0 || i >= (int)P.convs.size()) return fail(MMVAE_ERR_BAD_ARG, "conv index %d out of range", i);
  const ConvT_& c = P.convs[i];
  if (name && name_cap) { strncpy(name, c.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (shape) { shape[0] = c.kind; shape[1] = c.k; shape[2] = c.s; shape[3] = c.p; shape[4] = c.Ci; shape[5] = c.Co; shape[6] = c.Hi; shape[7] = c.Ho; }
  return 0;
}

}  // extern "C"
