// Internal (non-ABI) entry points shared between the conv dispatch and its backends.
#pragma once
#include "common.cuh"

int nvae_conv_check(const NvaeConvDesc* d);
size_t nvae_colsum_ws_bytes(int C);
int nvae_colsum(const float* a, int64_t rows, int C, int ld, float* out, void* ws, size_t ws_bytes,
                cudaStream_t stream);

// fp32 CUDA-core backend (conv_simt.cu)
int nvae_conv2d_fwd_simt(const NvaeConvDesc* d, const float* x, const float* x2, const float* w, const float* bias,
                         const float* residual, float* y, cudaStream_t stream);
int nvae_conv2d_dgrad_simt(const NvaeConvDesc* d, const float* dy, const float* w, float* dx, float* dx2,
                           int accumulate, cudaStream_t stream);
size_t nvae_conv2d_wgrad_simt_ws_bytes(const NvaeConvDesc* d);
int nvae_conv2d_wgrad_simt(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw,
                           void* ws, size_t ws_bytes, cudaStream_t stream);

// tcgen05 / TMA backend (conv_tc.cu).  which: 0 fwd, 1 dgrad, 2 wgrad.
bool nvae_conv_tc_supported(const NvaeConvDesc* d, int which);
size_t nvae_conv_tc_ws_bytes(const NvaeConvDesc* d, int which);
bool nvae_conv_tc_plan_info(const NvaeConvDesc* d, int which, int32_t* out);
int nvae_conv2d_fwd_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* w_tr, const float* bias,
                       const float* residual, float* y, void* ws, size_t ws_bytes, cudaStream_t stream,
                       const float* pro_stat = nullptr, int pro_act = 0);
int nvae_conv2d_dgrad_tc(const NvaeConvDesc* d, const float* dy, const float* w_rnd, float* dx, float* dx2,
                         int accumulate, void* ws, size_t ws_bytes, cudaStream_t stream);
int nvae_conv2d_wgrad_tc(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw, void* ws,
                         size_t ws_bytes, cudaStream_t stream, const float* pro_stat = nullptr, int pro_act = 0);
// operand prolog  A = act(x * scale + shift)  (rows 2, 3 of a BatchNorm stat block) inside forward / backward-filter
bool nvae_conv_tc_prolog_supported(const NvaeConvDesc* d);
int nvae_round_tf32_inplace(float* p, int64_t n, cudaStream_t stream);
