// Conv2D(padding="same") C-ABI entry points: validate, pick the backend, launch.
//   NVAE_PREC_FP32            -> fp32 CUDA-core implicit GEMM (conv_simt.cu), any shape
//   NVAE_PREC_TF32 / TF32X3   -> tcgen05 implicit GEMM (conv_tc.cu) for the shapes it takes,
//                                 otherwise the fp32 backend (strictly more accurate)
// There is no CPU fallback anywhere: an unsupported descriptor is an error code.
#include "conv_internal.h"

#include <stdlib.h>
#include <string.h>

using namespace nvae;

// NVAE_PDL=1 turns programmatic dependent launch on (read once; see common.cuh)
bool nvae::pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("NVAE_PDL");
    return e != nullptr && strcmp(e, "1") == 0;
  }();
  return on;
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

extern "C" size_t nvae_conv2d_ws_bytes(const NvaeConvDesc* d, int which) {
  if (nvae_conv_check(d) != NVAE_OK) return 0;
  const size_t tc = align256(nvae_conv_tc_ws_bytes(d, which));
  if (which == 2) {
    const size_t simt = align256(nvae_conv2d_wgrad_simt_ws_bytes(d));
    return (tc > simt ? tc : simt) + nvae_colsum_ws_bytes(d->Cout) + 256;
  }
  return tc + 256;
}

extern "C" int nvae_conv2d_fwd(const NvaeConvDesc* d, const float* x, const float* x2, const float* w,
                               const float* w_tr, const float* bias, const float* residual, float* y, void* ws,
                               size_t ws_bytes, nvae_stream_t stream) {
  int rc = nvae_conv_check(d);
  if (rc) return rc;
  if (!x || !w || !y || (d->Cin2 > 0 && !x2)) return NVAE_E_NULLPTR;
  if (d->precision != NVAE_PREC_FP32 && w_tr != nullptr && nvae_conv_tc_supported(d, 0))
    return nvae_conv2d_fwd_tc(d, x, x2, w_tr, bias, residual, y, ws, ws_bytes, stream);
  return nvae_conv2d_fwd_simt(d, x, x2, w, bias, residual, y, stream);
}

extern "C" int nvae_conv2d_dgrad(const NvaeConvDesc* d, const float* dy, const float* w, const float* w_rnd, float* dx,
                                 float* dx2, int accumulate, void* ws, size_t ws_bytes, nvae_stream_t stream) {
  int rc = nvae_conv_check(d);
  if (rc) return rc;
  if (!dy || !w || (!dx && !dx2)) return NVAE_E_NULLPTR;
  if (d->precision != NVAE_PREC_FP32 && w_rnd != nullptr && nvae_conv_tc_supported(d, 1))
    return nvae_conv2d_dgrad_tc(d, dy, w_rnd, dx, dx2, accumulate, ws, ws_bytes, stream);
  return nvae_conv2d_dgrad_simt(d, dy, w, dx, dx2, accumulate, stream);
}

extern "C" int nvae_conv2d_wgrad(const NvaeConvDesc* d, const float* x, const float* x2, const float* dy, float* dw,
                                 float* dbias, void* ws, size_t ws_bytes, nvae_stream_t stream) {
  int rc = nvae_conv_check(d);
  if (rc) return rc;
  if (!x || !dy || !dw || (d->Cin2 > 0 && !x2)) return NVAE_E_NULLPTR;
  const int64_t rows = (int64_t)d->N * d->Ho * d->Wo;
  const int ld = d->y_ld > 0 ? d->y_ld : d->Cout;
  if (d->precision != NVAE_PREC_FP32 && nvae_conv_tc_supported(d, 2)) {
    rc = nvae_conv2d_wgrad_tc(d, x, x2, dy, dw, ws, ws_bytes, stream);
    if (rc) return rc;
    if (dbias != nullptr) {
      const size_t off = align256(nvae_conv_tc_ws_bytes(d, 2));
      if (ws == nullptr || ws_bytes < off + nvae_colsum_ws_bytes(d->Cout)) return NVAE_E_WORKSPACE;
      return nvae_colsum(dy + d->y_off, rows, d->Cout, ld, dbias, (char*)ws + off, ws_bytes - off, stream);
    }
    return NVAE_OK;
  }
  const size_t wneed = align256(nvae_conv2d_wgrad_simt_ws_bytes(d));
  rc = nvae_conv2d_wgrad_simt(d, x, x2, dy, dw, ws, ws_bytes, stream);
  if (rc) return rc;
  if (dbias != nullptr) {
    if (ws == nullptr || ws_bytes < wneed + nvae_colsum_ws_bytes(d->Cout)) return NVAE_E_WORKSPACE;
    return nvae_colsum(dy + d->y_off, rows, d->Cout, ld, dbias, (char*)ws + wneed, ws_bytes - wneed, stream);
  }
  return NVAE_OK;
}

// ---- convolution whose input is act(BN(x)): BN-apply + activation in the operand path ----------------------------
extern "C" int nvae_conv2d_bnact_supported(const NvaeConvDesc* d) {
  if (nvae_conv_check(d) != NVAE_OK || d->precision == NVAE_PREC_FP32) return 0;
  return nvae_conv_tc_supported(d, 0) && nvae_conv_tc_supported(d, 2) && nvae_conv_tc_prolog_supported(d) ? 1 : 0;
}

extern "C" int nvae_conv2d_fwd_bnact(const NvaeConvDesc* d, const float* x, const float* stat, int act, const float* w_tr,
                                     const float* bias, const float* residual, float* y, void* ws, size_t ws_bytes,
                                     nvae_stream_t stream) {
  int rc = nvae_conv_check(d);
  if (rc) return rc;
  if (!x || !stat || !w_tr || !y) return NVAE_E_NULLPTR;
  if (act != NVAE_ACT_NONE && act != NVAE_ACT_SWISH && act != NVAE_ACT_ELU) return NVAE_E_BADSHAPE;
  if (!nvae_conv2d_bnact_supported(d)) return NVAE_E_UNSUPPORTED;  // the caller applies the BN itself: no silent second path
  return nvae_conv2d_fwd_tc(d, x, nullptr, w_tr, bias, residual, y, ws, ws_bytes, stream, stat, act);
}

extern "C" int nvae_conv2d_wgrad_bnact(const NvaeConvDesc* d, const float* x, const float* stat, int act,
                                       const float* dy, float* dw, float* dbias, void* ws, size_t ws_bytes,
                                       nvae_stream_t stream) {
  int rc = nvae_conv_check(d);
  if (rc) return rc;
  if (!x || !stat || !dy || !dw) return NVAE_E_NULLPTR;
  if (act != NVAE_ACT_NONE && act != NVAE_ACT_SWISH && act != NVAE_ACT_ELU) return NVAE_E_BADSHAPE;
  if (!nvae_conv2d_bnact_supported(d)) return NVAE_E_UNSUPPORTED;
  rc = nvae_conv2d_wgrad_tc(d, x, nullptr, dy, dw, ws, ws_bytes, stream, stat, act);
  if (rc) return rc;
  if (dbias != nullptr) {
    const size_t off = align256(nvae_conv_tc_ws_bytes(d, 2));
    if (ws == nullptr || ws_bytes < off + nvae_colsum_ws_bytes(d->Cout)) return NVAE_E_WORKSPACE;
    return nvae_colsum(dy + d->y_off, (int64_t)d->N * d->Ho * d->Wo, d->Cout, d->y_ld > 0 ? d->y_ld : d->Cout, dbias,
                       (char*)ws + off, ws_bytes - off, stream);
  }
  return NVAE_OK;
}

extern "C" int nvae_round_tf32(float* p, int64_t n, nvae_stream_t stream) {
  if (n > 0 && p == nullptr) return NVAE_E_NULLPTR;
  return nvae_round_tf32_inplace(p, n, stream);
}

extern "C" int nvae_conv2d_plan_info(const NvaeConvDesc* d, int which, int32_t* out) {
  if (out == nullptr) return 0;
  for (int i = 0; i < 16; ++i) out[i] = 0;
  if (nvae_conv_check(d) != NVAE_OK || d->precision == NVAE_PREC_FP32 || which < 0 || which > 2) return 0;
  return nvae_conv_tc_plan_info(d, which, out) ? 1 : 0;
}

extern "C" int nvae_conv2d_uses_tensor_cores(const NvaeConvDesc* d, int which) {
  if (nvae_conv_check(d) != NVAE_OK || d->precision == NVAE_PREC_FP32) return 0;
  return nvae_conv_tc_supported(d, which) ? 1 : 0;
}
