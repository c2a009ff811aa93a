// tcgen05 / TMA implicit-GEMM kernels (placeholder until the tensor-core path lands: every entry
// point reports BDBNN_ERR_UNSUPPORTED and bdbnn_tc_supported() answers 0, so callers take the
// CUDA-core kernels in binconv.cu).
#include "common.cuh"

using namespace bdbnn;

extern "C" int bdbnn_tc_supported(const bdbnn_conv_shape*) { return 0; }

extern "C" int bdbnn_binconv_fwd_tc(const uint16_t*, const uint16_t*, const float*, float*,
                                    const bdbnn_conv_shape*, void*) {
  set_error("binconv_fwd_tc: not built");
  return BDBNN_ERR_UNSUPPORTED;
}
extern "C" int bdbnn_binconv_dgrad_tc(const uint16_t*, const uint16_t*, const uint32_t*, float*,
                                      const bdbnn_conv_shape*, void*) {
  set_error("binconv_dgrad_tc: not built");
  return BDBNN_ERR_UNSUPPORTED;
}
extern "C" size_t bdbnn_wgrad_tc_workspace_bytes(const bdbnn_conv_shape*) { return 0; }
extern "C" int bdbnn_binconv_wgrad_tc(const uint16_t*, const uint16_t*, const uint32_t*, const float*,
                                      float*, const bdbnn_conv_shape*, void*, size_t, void*) {
  set_error("binconv_wgrad_tc: not built");
  return BDBNN_ERR_UNSUPPORTED;
}
