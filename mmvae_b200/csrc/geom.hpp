// geom.hpp -- tap enumeration of every Conv2d / ConvTranspose2d of the model and of its data gradient,
// shared by the host (gather descriptors, api.cu) and the device (weight packing, gconv_tc.cu) so both
// always agree on the (variant, tap) -> (input offset, weight offset) mapping.
//
// A "variant" is one output-parity class of a stride-2 transposed convolution (or of the data gradient
// of a stride-2 convolution): within a variant every output pixel sees the same set of taps.
//   ConvTranspose2d k4 s2 p1 (model.py:62-65,198-201): oy = 2*iy - 1 + ky  ->  parity py only sees
//   ky == (py+1) mod 2: four 2x2-tap sub-convolutions, no scatter, no atomics.
#pragma once

#ifdef __CUDACC__
#define MMVAE_HD __host__ __device__
#else
#define MMVAE_HD
#endif

namespace mmvae {

enum { GEOM_CONV = 0, GEOM_CONVT = 1 };
enum { DIR_FPROP = 0, DIR_DGRAD = 1 };

struct ConvGeom {
  int kind, k, s, p;   // as constructed in the reference
  int Ci, Co;          // the conv's own in / out channels (weight [Co][Ci][k][k], transposed: [Ci][Co][k][k])
};

// per-axis tap count of parity `par` for the data gradient of a stride-2 convolution
MMVAE_HD inline int s2_axis_taps(const ConvGeom& g, int par) {
  int first = (par + g.p) & 1;
  return first < g.k ? (g.k - first + 1) / 2 : 0;
}

// (py, px) of the v-th non-empty variant; returns false when v is out of range
MMVAE_HD inline bool dgrad_s2_variant(const ConvGeom& g, int v, int& py, int& px) {
  int idx = 0;
  for (int y = 0; y < 2; ++y)
    for (int x = 0; x < 2; ++x) {
      if (s2_axis_taps(g, y) * s2_axis_taps(g, x) == 0) continue;
      if (idx == v) { py = y; px = x; return true; }
      ++idx;
    }
  return false;
}

MMVAE_HD inline int geom_nvar(const ConvGeom& g, int dir) {
  if (dir == DIR_FPROP) return g.kind == GEOM_CONV ? 1 : 4;
  if (g.kind == GEOM_CONV && g.s == 2) {
    int n = 0, py, px;
    while (dgrad_s2_variant(g, n, py, px)) ++n;
    return n;
  }
  return 1;
}

// output-parity origin of variant v: out y = oy0 + os * i
MMVAE_HD inline void geom_origin(const ConvGeom& g, int dir, int v, int& oy0, int& ox0) {
  oy0 = 0; ox0 = 0;
  if (dir == DIR_FPROP) {
    if (g.kind == GEOM_CONVT) { oy0 = v >> 1; ox0 = v & 1; }
  } else if (g.kind == GEOM_CONV && g.s == 2) {
    dgrad_s2_variant(g, v, oy0, ox0);
  }
}

// Tap t of variant v: input offset (dy, dx) relative to is*i, is*j and the offset of the tap inside the
// reference-layout weight tensor (kh*k + kw).  Returns false when t >= number of taps of the variant.
MMVAE_HD inline bool geom_tap(const ConvGeom& g, int dir, int v, int t, int& dy, int& dx, int& wofs) {
  const int k = g.k;
  if (dir == DIR_FPROP) {
    if (g.kind == GEOM_CONV) {
      if (t >= k * k) return false;
      int kh = t / k, kw = t - kh * k;
      dy = kh - g.p; dx = kw - g.p; wofs = t;
      return true;
    }
    int py = v >> 1, px = v & 1;
    if (k == 4) {
      if (t >= 4) return false;
      int ky = ((py + 1) & 1) + 2 * (t >> 1), kx = ((px + 1) & 1) + 2 * (t & 1);
      dy = (py + 1 - ky) / 2; dx = (px + 1 - kx) / 2; wofs = ky * 4 + kx;
      return true;
    }
    // ConvTranspose2d k2 s1 p0 on a 1x1 input (model.py:159-161): y[ky,kx] = W[:, :, ky, kx]^T z
    if (t >= 1) return false;
    dy = 0; dx = 0; wofs = py * 2 + px;
    return true;
  }
  // ---- data gradient: input = dY, output = dX ----
  if (g.kind == GEOM_CONV) {
    if (g.s == 1) {
      if (t >= k * k) return false;
      int kh = t / k, kw = t - kh * k;
      dy = g.p - kh; dx = g.p - kw; wofs = t;
      return true;
    }
    // stride 2: dX[2i+py] = sum over kh with (py + p - kh) even of dY[i + (py+p-kh)/2]
    int py = 0, px = 0;
    if (!dgrad_s2_variant(g, v, py, px)) return false;
    int ny = s2_axis_taps(g, py), nx = s2_axis_taps(g, px);
    if (t >= ny * nx) return false;
    int a = t / nx, b = t - a * nx;
    int kh = ((py + g.p) & 1) + 2 * a, kw = ((px + g.p) & 1) + 2 * b;
    dy = (py + g.p - kh) / 2; dx = (px + g.p - kw) / 2; wofs = kh * k + kw;
    return true;
  }
  if (k == 4) {
    // dX[iy] = sum_ky dY[2*iy - 1 + ky] W[ci][co][ky][kx]: a stride-2 4x4 convolution over dY
    if (t >= 16) return false;
    dy = (t >> 2) - 1; dx = (t & 3) - 1; wofs = t;
    return true;
  }
  if (t >= 4) return false;
  dy = t >> 1; dx = t & 1; wofs = t;
  return true;
}

MMVAE_HD inline int geom_ntaps(const ConvGeom& g, int dir, int v) {
  int n = 0, a, b, c;
  while (geom_tap(g, dir, v, n, a, b, c)) ++n;
  return n;
}

// channel counts / weight strides of the gather-convolution that implements (g, dir):
// op_ci = channels gathered (the GEMM K axis per tap), op_co = channels produced (GEMM N axis)
MMVAE_HD inline void geom_strides(const ConvGeom& g, int dir, int& op_ci, int& op_co, int& w_sci, int& w_sco) {
  const int kk = g.k * g.k;
  if (dir == DIR_FPROP) {
    op_ci = g.Ci; op_co = g.Co;
    if (g.kind == GEOM_CONV) { w_sci = kk; w_sco = g.Ci * kk; }      // weight [Co][Ci][k][k]
    else                     { w_sci = g.Co * kk; w_sco = kk; }      // weight [Ci][Co][k][k]
  } else {
    op_ci = g.Co; op_co = g.Ci;
    if (g.kind == GEOM_CONV) { w_sci = g.Ci * kk; w_sco = kk; }
    else                     { w_sci = kk; w_sco = g.Co * kk; }
  }
}

}  // namespace mmvae
