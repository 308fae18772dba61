// tcgen05 implicit-GEMM convolution backend -- placeholder until the kernels land.
#include "conv_internal.h"

bool nvae_conv_tc_supported(const NvaeConvDesc*, int) { return false; }
size_t nvae_conv_tc_ws_bytes(const NvaeConvDesc*, int) { return 0; }
int nvae_conv2d_fwd_tc(const NvaeConvDesc*, const float*, const float*, const float*, const float*, const float*,
                       float*, void*, size_t, cudaStream_t) { return NVAE_E_UNSUPPORTED; }
int nvae_conv2d_dgrad_tc(const NvaeConvDesc*, const float*, const float*, float*, float*, int, void*, size_t,
                         cudaStream_t) { return NVAE_E_UNSUPPORTED; }
int nvae_conv2d_wgrad_tc(const NvaeConvDesc*, const float*, const float*, const float*, float*, void*, size_t,
                         cudaStream_t) { return NVAE_E_UNSUPPORTED; }
