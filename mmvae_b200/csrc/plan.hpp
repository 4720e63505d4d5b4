// plan.hpp -- host-side description of the VAE step: parameter arena, BatchNorm table,
// activation workspace.  Pure C++ (no CUDA), rebuilt from the mmvae_desc on every call, so
// the library stays stateless.
//
// Architecture follows the reference: VAE_Encoder model.py:88-146, BasicBlock model.py:23-55,
// VAE_Decoder model.py:153-209, DeconvBottleneck model.py:57-85.
#pragma once
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mmvae.h"
#include "geom.hpp"

namespace mmvae {

struct ParamT {            // one entry of named_parameters()
  std::string name;
  int64_t off;             // floats into the parameter / gradient arena
  int ndim;
  int shape[4];
  int64_t numel() const { int64_t n = 1; for (int i = 0; i < ndim; ++i) n *= shape[i]; return n; }
};

struct ActT {              // NHWC tensor in the workspace (storage type of the precision mode)
  std::string name;
  size_t off = 0;          // bytes: the tensor itself
  size_t goff = 0;         // bytes: its gradient (same shape / type)
  size_t goff2 = 0;        // bytes: second gradient buffer of a block input (the shortcut branch's data gradient lands
                           // here, beside the main branch's in goff; their sum is the gradient), or 0
  int H = 0, W = 0, C = 0;
};

struct BnT {               // one BatchNorm2d
  std::string prefix;
  int C = 0;
  int64_t gamma = 0, beta = 0;   // parameter arena offsets
  int64_t rm = 0;                // buffer arena offset of running_mean (running_var at rm + C)
  int idx = 0;                   // index into bn_counters
  int64_t m = 0;                 // N*H*W elements per channel
  size_t stat_off = 0;           // fp32 [2][C]: batch mean, rstd        (saved for backward)
  size_t coef_off = 0;           // fp32 [2][C]: scale = gamma*rstd, shift = beta - mean*scale
  size_t part_off = 0;           // fp32 forward partial sums [P][C][2]
  size_t part_cap = 0;           // P capacity
  size_t pcnt_off = 0;           // fp32 [max(P, 1024)]: rows behind each partial row (persistent kernels)
  size_t bpart_off = 0;          // fp32 backward partial sums [PB][C][3]
  size_t bcoef_off = 0;          // fp32 [5][C]: backward coefficients (scale, c1, c2) and the sums S0, S1
  size_t acc_off = 0;            // fp64 [8 copies][2][C] forward (sum, sumsq) then [8 copies][3][C] backward (s0, s1, s2)
  size_t cnt_off = 0;            // uint32 [2]: CTA-done counters (fwd, bwd)
};

enum ConvKind { CONV = 0, CONVT = 1 };

struct ConvT_ {            // one Conv2d / ConvTranspose2d
  std::string name;
  int kind = CONV;
  int k = 1, s = 1, p = 0;
  int Ci = 0, Co = 0;
  int Hi = 0, Wi = 0, Ho = 0, Wo = 0;
  int64_t w = 0;           // parameter arena offset of the weight
  int64_t bias = -1;       // ... of the bias, or -1
  int in = -1, out = -1;   // activation indices (in == -1: the fp32 NCHW network input / latent)
  int bn = -1;
  // tcgen05 path: packed bf16 weight tiles in the workspace (bytes), per direction; 0 chunks = not packed
  size_t wp_off[2] = {0, 0};
  int wp_chunks[2] = {0, 0};       // k-chunks of 64 per variant
  ConvGeom geom() const { return ConvGeom{kind == CONV ? GEOM_CONV : GEOM_CONVT, k, s, p, Ci, Co}; }
};

struct BlockT {            // BasicBlock or DeconvBottleneck: main = c1 -> bn1 -> relu -> c2 -> bn2, shortcut = cs -> bns
  int c1, c2, cs;          // conv indices
  int a1;                  // activation index of relu(bn1(y1))
  int out;                 // activation index of the block output
  int in;                  // activation index of the block input
};

// The notebook variant (MMVAE_ARCH_NOTEBOOK, vae-kl.ipynb:119-166): indices into Plan::convs / Plan::acts
struct NbT {
  int e[4] = {-1, -1, -1, -1};       // encoder.conv1..conv4 (ReLU)
  int cmu = -1, clv = -1;            // encoder.conv_mu / conv_logvar
  int dc[4] = {-1, -1, -1, -1};      // decoder.conv1..conv4 (ELU, ELU, ELU, none)
  int a_z = -1;                      // sampled latent, NHWC [N,h,h,z]
  int a_up[4] = {-1, -1, -1, -1};    // nearest-upsampled input of decoder.conv{k+1}
  int up[4] = {2, 4, 2, 2};          // vae-kl.ipynb:152-155
  size_t eps_off = 0;                // fp32 [N,z,h,h]: the rsample draw, kept for the backward
  size_t acc_off = 0;                // fp64 [2]: sum CE, sum KL
  int latent_hw = 0;                 // h
};

constexpr int kBwdBlocks = 592;   // grid of the BatchNorm-backward reduction (4 x 148 SMs)

struct Plan {
  mmvae_desc d{};
  int esz = 4;             // bytes per activation element
  std::vector<ParamT> params;
  std::vector<BnT> bns;
  std::vector<ActT> acts;
  std::vector<ConvT_> convs;
  int stem = -1, dstem = -1, tail = -1;
  int a_stem = -1, a_dstem = -1, a_z = -1;
  std::vector<BlockT> enc, dec;
  int64_t w_mu = -1, w_lv = -1;
  int feat_c = 0, feat_hw = 0;       // encoder output channels / spatial positions (pool)
  int dec_size = 0, crop = 0;
  int64_t n_params = 0, n_bn_buffers = 0;
  size_t pooled_off = 0;             // fp32 [N][feat_c]
  size_t heads_off = 0;              // fp32 [N][z] x 4: mu, logvar, eps, std
  size_t dheads_off = 0;             // fp32 [N][z] x 3: dmu, dlogvar, dz
  size_t dpool_off = 0;              // fp32 [N][feat_c]
  size_t drecon_off = 0;             // fp32 NHWC copy of d_recon when out_channels > 1
  size_t bnacc_off = 0, bnacc_bytes = 0;   // all BatchNorm accumulators + counters: one memset per forward
  size_t ticket_off = 0;                   // uint32 inside that region: CTA ticket of the heads kernel (rng advance)
  size_t ws_bytes = 0;
  int64_t train_flops = 0;
  NbT nb;
  std::string err;

  // ---- helpers ----
  size_t bump(size_t bytes) {
    size_t o = ws_bytes;
    ws_bytes += (bytes + 255) & ~size_t(255);
    return o;
  }
  int64_t add_param(const std::string& name, int ndim, int a, int b = 1, int c = 1, int e = 1) {
    ParamT p; p.name = name; p.off = n_params; p.ndim = ndim;
    p.shape[0] = a; p.shape[1] = b; p.shape[2] = c; p.shape[3] = e;
    n_params += p.numel();
    params.push_back(p);
    return p.off;
  }
  int add_act(const std::string& name, int H, int W, int C) {
    ActT a; a.name = name; a.H = H; a.W = W; a.C = C;
    size_t bytes = size_t(d.batch) * H * W * C * esz;
    a.off = bump(bytes);
    a.goff = bump(bytes);
    acts.push_back(a);
    return int(acts.size()) - 1;
  }
  int add_bn(const std::string& prefix, int C, int64_t m, size_t part_rows) {
    BnT b; b.prefix = prefix; b.C = C; b.m = m;
    b.gamma = add_param(prefix + ".weight", 1, C);
    b.beta = add_param(prefix + ".bias", 1, C);
    b.rm = n_bn_buffers; n_bn_buffers += 2 * int64_t(C);
    b.idx = int(bns.size());
    b.stat_off = bump(sizeof(float) * 2 * C);
    b.coef_off = bump(sizeof(float) * 2 * C);
    b.part_cap = part_rows;
    b.part_off = bump(sizeof(float) * 2 * C * std::max<size_t>(part_rows, 1024));
    b.pcnt_off = bump(sizeof(float) * std::max<size_t>(part_rows, 1024));
    b.bpart_off = bump(sizeof(float) * 3 * C * kBwdBlocks);
    b.bcoef_off = bump(sizeof(float) * 5 * C);
    bns.push_back(b);
    return b.idx;
  }
  // conv + its BatchNorm; registers parameters in the reference's order (weight, [bias], bn.weight, bn.bias)
  int add_conv(const std::string& name, const std::string& bn_prefix, int kind, int k, int s, int p,
               int Ci, int Co, int Hi, int in_act, bool bias = false) {
    ConvT_ c; c.name = name; c.kind = kind; c.k = k; c.s = s; c.p = p; c.Ci = Ci; c.Co = Co;
    c.Hi = c.Wi = Hi;
    c.Ho = c.Wo = (kind == CONV) ? (Hi + 2 * p - k) / s + 1 : (Hi - 1) * s - 2 * p + k;
    c.in = in_act;
    if (kind == CONV) c.w = add_param(name + ".weight", 4, Co, Ci, k, k);
    else              c.w = add_param(name + ".weight", 4, Ci, Co, k, k);
    if (bias) c.bias = add_param(name + ".bias", 1, Co);
    c.out = add_act(name, c.Ho, c.Wo, Co);
    int64_t m = int64_t(d.batch) * c.Ho * c.Wo;
    // forward partial-sum rows: one per CTA of the producing kernel; the smallest CTA tile covers 64
    // GEMM rows, transposed convs launch up to 4 parity variants
    int nvar = (kind == CONVT) ? 4 : 1;
    int64_t rows_per_var = (kind == CONVT) ? int64_t(d.batch) * Hi * Hi : m;
    if (kind == CONVT && k == 2) rows_per_var = d.batch;
    // ... the tcgen05 kernels write one row per epilogue warp: 4 per 128-row tile, or 4 per CTA (<= 296 CTAs)
    size_t part_rows = std::max<size_t>(size_t(nvar) * size_t((rows_per_var + 63) / 64) * 2 + 8, 1280);
    c.bn = add_bn(bn_prefix, Co, m, part_rows);
    // 2*MAC: forward + wgrad always, dgrad unless it is the first layer (input needs no gradient)
    int64_t macs = (kind == CONV) ? m * Co * int64_t(Ci) * k * k
                                  : int64_t(d.batch) * Hi * Hi * int64_t(Ci) * Co * k * k;
    train_flops += 2 * macs * ((in_act == -1 && name == "encoder.conv1") ? 2 : 3);
    convs.push_back(c);
    return int(convs.size()) - 1;
  }

  // conv + bias without BatchNorm (notebook variant); parameters in the order (weight, bias)
  int add_conv_nb(const std::string& name, int k, int s, int p, int Ci, int Co, int Hi, int in_act) {
    ConvT_ c; c.name = name; c.kind = CONV; c.k = k; c.s = s; c.p = p; c.Ci = Ci; c.Co = Co;
    c.Hi = c.Wi = Hi;
    c.Ho = c.Wo = (Hi + 2 * p - k) / s + 1;
    c.in = in_act;
    c.w = add_param(name + ".weight", 4, Co, Ci, k, k);
    c.bias = add_param(name + ".bias", 1, Co);
    c.out = add_act(name, c.Ho, c.Wo, Co);
    const int64_t macs = int64_t(d.batch) * c.Ho * c.Wo * Co * int64_t(Ci) * k * k;
    train_flops += 2 * macs * (in_act == -1 ? 2 : 3);
    convs.push_back(c);
    return int(convs.size()) - 1;
  }

  // packed bf16 weight tiles of the tcgen05 path, both directions
  void plan_packing() {
    if (d.precision != MMVAE_PREC_BF16 || (d.flags & MMVAE_FLAG_FORCE_SIMT)) return;
    for (auto& c : convs) {
      if (c.Ci % 8 != 0 || c.Co % 8 != 0) continue;          // stem (Ci = in_channels) / tail (Co = out_channels): SIMT
      ConvGeom g = c.geom();
      for (int dir = 0; dir < 2; ++dir) {
        if (dir == DIR_DGRAD && (c.in < 0 || !d.training)) continue;
        int op_ci, op_co, sci, sco;
        geom_strides(g, dir, op_ci, op_co, sci, sco);
        const int nv = geom_nvar(g, dir);
        int mc = 1;
        for (int v = 0; v < nv; ++v) mc = std::max(mc, (geom_ntaps(g, dir, v) * op_ci + 63) / 64);
        const int co_pad = (op_co + 15) & ~15;
        c.wp_chunks[dir] = mc;
        c.wp_off[dir] = bump(size_t(nv) * mc * co_pad * 128);
      }
    }
  }

  // vae-kl.ipynb:122-166
  bool build_nb() {
    if (d.in_channels != 1) { err = "notebook variant: in_channels must be 1"; return false; }
    if (d.image_size != 64 && d.image_size != 128) { err = "notebook variant: image_size must be 64 or 128"; return false; }
    if (d.out_channels % 8 != 0 || d.out_channels > 256) { err = "notebook variant: classes must be a multiple of 8, <= 256"; return false; }
    if (d.z_dim % 8 != 0) { err = "notebook variant: z_dim must be a multiple of 8"; return false; }
    d.training = 1;                                              // no BatchNorm: one mode
    const int C = 32 * d.width, S = d.image_size, z = d.z_dim;
    nb.e[0] = add_conv_nb("encoder.conv1", 5, 2, 2, 1, C, S, -1);
    nb.e[1] = add_conv_nb("encoder.conv2", 5, 2, 1, C, C, convs[nb.e[0]].Ho, convs[nb.e[0]].out);
    nb.e[2] = add_conv_nb("encoder.conv3", 3, 2, 1, C, C, convs[nb.e[1]].Ho, convs[nb.e[1]].out);
    nb.e[3] = add_conv_nb("encoder.conv4", 3, 2, 1, C, C, convs[nb.e[2]].Ho, convs[nb.e[2]].out);
    nb.cmu = add_conv_nb("encoder.conv_mu", 3, 2, 1, C, z, convs[nb.e[3]].Ho, convs[nb.e[3]].out);
    nb.clv = add_conv_nb("encoder.conv_logvar", 3, 2, 1, C, z, convs[nb.e[3]].Ho, convs[nb.e[3]].out);
    const int h = convs[nb.cmu].Ho;
    nb.latent_hw = h;
    nb.a_z = add_act("decoder.input", h, h, z);
    int curC = z, curH = h;
    const int couts[4] = {C, C, C, d.out_channels};
    for (int i = 0; i < 4; ++i) {
      const std::string nm = "decoder.conv" + std::to_string(i + 1);
      curH *= nb.up[i];
      nb.a_up[i] = add_act(nm + ".input", curH, curH, curC);
      nb.dc[i] = add_conv_nb(nm, 3, 1, 1, curC, couts[i], curH, nb.a_up[i]);
      curC = couts[i];
    }
    if (curH != S) { err = "notebook variant: decoder size mismatch"; return false; }
    dec_size = S; crop = 0;
    nb.eps_off = bump(sizeof(float) * size_t(d.batch) * z * h * h);
    nb.acc_off = bump(sizeof(double) * 2);
    plan_packing();
    return true;
  }

  bool build(const mmvae_desc* dd) {
    if (!dd) { err = "desc is NULL"; return false; }
    if (dd->struct_size != (int32_t)sizeof(mmvae_desc)) { err = "mmvae_desc.struct_size mismatch"; return false; }
    d = *dd;
    if (d.batch < 1) { err = "batch must be >= 1"; return false; }
    if (d.arch != MMVAE_ARCH_RESNET && d.arch != MMVAE_ARCH_NOTEBOOK) { err = "unknown arch"; return false; }
    if (d.arch == MMVAE_ARCH_NOTEBOOK) {
      if (d.width < 1 || d.width > 8) { err = "width must be in [1,8]"; return false; }
      if (d.z_dim < 1 || d.z_dim > 1024) { err = "z_dim must be in [1,1024]"; return false; }
      if (d.precision != MMVAE_PREC_FP32 && d.precision != MMVAE_PREC_BF16) { err = "unknown precision"; return false; }
      esz = (d.precision == MMVAE_PREC_BF16) ? 2 : 4;
      return build_nb();
    }
    if (d.in_channels < 1 || d.in_channels > 16) { err = "in_channels must be in [1,16]"; return false; }
    if (d.out_channels < 1 || d.out_channels > 256) { err = "out_channels must be in [1,256]"; return false; }
    if (d.z_dim < 1 || d.z_dim > 1024) { err = "z_dim must be in [1,1024]"; return false; }
    if (d.width < 1 || d.width > 8) { err = "width must be in [1,8]"; return false; }
    if (d.image_size < 16 || d.image_size > 64) {
      err = "input_image_size must be in [16,64] (the reference decoder emits at most 64x64, model.py:169-170,307-310)";
      return false;
    }
    if (d.precision != MMVAE_PREC_FP32 && d.precision != MMVAE_PREC_BF16) { err = "unknown precision"; return false; }
    esz = (d.precision == MMVAE_PREC_BF16) ? 2 : 4;
    const int w = d.width, S = d.image_size;
    dec_size = S > 32 ? 64 : 32;
    crop = (dec_size - S) / 2;                                   // model.py:307-310
    if ((dec_size - S) % 2 != 0) { err = "input_image_size must be even (adjust = (64-S)//2 crop)"; return false; }

    // ---------------- encoder (model.py:92-107) ----------------
    stem = add_conv("encoder.conv1", "encoder.bn1", CONV, 5, 2, 2, d.in_channels, 32 * w, S, -1);
    a_stem = add_act("encoder.relu", convs[stem].Ho, convs[stem].Wo, 32 * w);
    int cur = a_stem, curC = 32 * w, curH = convs[stem].Ho;
    const int planes[4] = {32 * w, 64 * w, 128 * w, 256 * w};
    for (int i = 0; i < 4; ++i) {
      std::string p = "encoder.layer" + std::to_string(i + 1) + ".0";
      BlockT b; b.in = cur;
      b.c1 = add_conv(p + ".conv1", p + ".bn1", CONV, 3, 2, 1, curC, planes[i], curH, cur);
      int H2 = convs[b.c1].Ho;
      b.a1 = add_act(p + ".relu1", H2, H2, planes[i]);
      b.c2 = add_conv(p + ".conv2", p + ".bn2", CONV, 3, 1, 1, planes[i], planes[i], H2, b.a1);
      b.cs = add_conv(p + ".downsample.0", p + ".downsample.1", CONV, 1, 2, 0, curC, planes[i], curH, cur);
      if (convs[b.cs].Ho != H2) { err = "shortcut / main spatial mismatch"; return false; }
      b.out = add_act(p, H2, H2, planes[i]);
      enc.push_back(b);
      cur = b.out; curC = planes[i]; curH = H2;
    }
    feat_c = curC; feat_hw = curH * curH;
    w_mu = add_param("encoder.conv_mu.weight", 4, d.z_dim, curC, 1, 1);
    if (d.require_rsample) w_lv = add_param("encoder.conv_logvar.weight", 4, d.z_dim, curC, 1, 1);
    train_flops += 2LL * 3 * d.batch * curC * d.z_dim * (d.require_rsample ? 2 : 1);
    pooled_off = bump(sizeof(float) * size_t(d.batch) * curC);
    dpool_off = bump(sizeof(float) * size_t(d.batch) * curC);
    heads_off = bump(sizeof(float) * size_t(d.batch) * d.z_dim * 4);
    dheads_off = bump(sizeof(float) * size_t(d.batch) * d.z_dim * 3);

    // ---------------- decoder (model.py:157-173) ----------------
    a_z = add_act("decoder.input", 1, 1, d.z_dim);
    dstem = add_conv("decoder.conv1", "decoder.bn1", CONVT, 2, 1, 0, d.z_dim, 128 * w, 1, a_z);
    a_dstem = add_act("decoder.relu", 2, 2, 128 * w);
    cur = a_dstem; curC = 128 * w; curH = 2;
    const int dplanes[5] = {128 * w, 64 * w, 32 * w, 16 * w, 16 * w};
    const int ndec = S > 32 ? 5 : 4;
    for (int i = 0; i < ndec; ++i) {
      std::string p = "decoder.uplayer" + std::to_string(i + 1) + ".0";
      BlockT b; b.in = cur;
      b.c1 = add_conv(p + ".conv1", p + ".bn1", CONV, 1, 1, 0, curC, dplanes[i], curH, cur);
      b.a1 = add_act(p + ".relu1", curH, curH, dplanes[i]);
      b.c2 = add_conv(p + ".conv2", p + ".bn2", CONVT, 4, 2, 1, dplanes[i], dplanes[i], curH, b.a1);
      b.cs = add_conv(p + ".upsample.0", p + ".upsample.1", CONVT, 4, 2, 1, curC, dplanes[i], curH, cur);
      b.out = add_act(p, 2 * curH, 2 * curH, dplanes[i]);
      dec.push_back(b);
      cur = b.out; curC = dplanes[i]; curH *= 2;
    }
    if (curH != dec_size) { err = "decoder size mismatch"; return false; }
    tail = add_conv("decoder.conv2", "decoder.bn2", CONV, 3, 1, 1, curC, d.out_channels, curH, cur, true);
    drecon_off = bump(sizeof(float) * size_t(d.batch) * dec_size * dec_size * d.out_channels);

    // ---------------- fused BatchNorm statistics: accumulators, contiguous ----------------
    {
      size_t total = 0;
      for (auto& b : bns) { b.acc_off = total; total += sizeof(double) * 8 * 5 * b.C; }
      for (auto& b : bns) { b.cnt_off = total; total += 16; }
      ticket_off = total; total += 16;
      bnacc_bytes = total;
      bnacc_off = bump(total);
      for (auto& b : bns) { b.acc_off += bnacc_off; b.cnt_off += bnacc_off; }
      ticket_off += bnacc_off;
    }

    // split gradients of the block inputs (api.cu block_bwd): bf16 product path, tensors wider than the 16 channels whose
    // BatchNorm-backward reduction is fused into the data-gradient epilogue instead
    if (d.precision == MMVAE_PREC_BF16 && !(d.flags & MMVAE_FLAG_FORCE_SIMT) && d.training) {
      for (const std::vector<BlockT>* bl : {&enc, &dec})
        for (const BlockT& b : *bl) {
          ActT& a = acts[b.in];
          if (a.C > 16) a.goff2 = bump(size_t(d.batch) * a.H * a.W * a.C * esz);
        }
    }
    plan_packing();
    return true;
  }

  const ActT* find_act(const char* name) const {
    for (auto& a : acts) if (a.name == name) return &a;
    return nullptr;
  }
};

}  // namespace mmvae
