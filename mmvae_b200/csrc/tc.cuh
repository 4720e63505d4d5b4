// tc.cuh -- per-layer kernel selection: tcgen05 tensor-core kernels (bf16 mode, shapes they cover)
// or the SIMT gather-convolution (fp32 validation mode and the remaining shapes: the stem conv with
// Ci = in_channels and the tail conv with Co = out_channels).
#pragma once
#include <type_traits>

#include "kernels.cuh"
#include "plan.hpp"

namespace mmvae {

// forward conv: returns the layout of the per-CTA partial statistics it wrote
template <typename T>
StatLayout conv_forward(const GConvParams& g, const ConvT_& c, cudaStream_t st) {
  (void)c;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (tc_supported_gconv(g)) return launch_gconv_tc(g, st);
  }
  return launch_gconv_simt<T>(g, st);
}

template <typename T>
void conv_dgrad(const GConvParams& g, const ConvT_& c, cudaStream_t st) {
  (void)c;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (tc_supported_gconv(g)) { launch_gconv_tc(g, st); return; }
  }
  launch_gconv_simt<T>(g, st);
}

template <typename T>
void conv_wgrad(const WGradParams& w, const ConvT_& c, bool use_tc, cudaStream_t st) {
  (void)c;
  if constexpr (std::is_same<T, __nv_bfloat16>::value) {
    if (use_tc && tc_supported_wgrad(w)) { launch_wgrad_tc(w, st); return; }
  }
  launch_wgrad_simt<T>(w, st);
}

}  // namespace mmvae
