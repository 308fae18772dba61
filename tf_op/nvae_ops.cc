// TensorFlow custom-op library over the C ABI of libnvae_b200.so (include/nvae_b200.h): one OpKernel per launcher of
// the residual-cell / latent hot path.  Each Compute() only extracts device pointers and the stream and forwards them;
// outputs and workspaces come from TensorFlow's allocator (the kernels never allocate), a non-zero status becomes
// errors::Internal -- there is no CPU fallback and no CPU kernel is registered.
//
// This is what a maintainer of stevensdavid/nvae-tf builds at their site (tf_op/Makefile, needs TensorFlow's headers):
//   make -C tf_op            ->  tf_op/libnvae_tf_ops.so,  loaded by tf_op/nvae_tf_layers.py with tf.load_op_library
// The build image of this repository has no TensorFlow; `make -C tf_op syntax` (run by tests/test_tf_op_source.py)
// compiles this file against the minimal declarations in tf_op/mock_tf/, so the calls below are at least type-checked
// against the real include/nvae_b200.h prototypes.
#define EIGEN_USE_GPU
#include <algorithm>
#include <cstdint>

#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "nvae_b200.h"

using namespace tensorflow;

namespace {

inline const float* fptr(const Tensor& t) { return t.NumElements() ? t.flat<float>().data() : nullptr; }
inline float* fptr(Tensor* t) { return t->NumElements() ? t->flat<float>().data() : nullptr; }
inline nvae_stream_t stream_of(OpKernelContext* ctx) {
  return reinterpret_cast<nvae_stream_t>(ctx->eigen_device<Eigen::GpuDevice>().stream());
}
// caller-owned workspace of `bytes` bytes (at least one element so the pointer is valid)
inline Status alloc_ws(OpKernelContext* ctx, size_t bytes, Tensor* ws) {
  return ctx->allocate_temp(DT_UINT8, TensorShape({static_cast<int64_t>(std::max<size_t>(bytes, 16))}), ws);
}
inline int64_t rows_of(const Tensor& t) { return t.NumElements() / t.dim_size(t.dims() - 1); }

// TF SAME geometry of common.py / encoder.py / decoder.py convolutions (the odd padding element goes after)
NvaeConvDesc conv_desc(int N, int H, int W, int Cin, int Cin2, int Cout, int R, int S, int stride, int precision) {
  NvaeConvDesc d{};
  d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.Cin2 = Cin2; d.Cout = Cout; d.R = R; d.S = S;
  d.stride = stride; d.precision = precision;
  d.Ho = (H + stride - 1) / stride; d.Wo = (W + stride - 1) / stride;
  d.pad_t = std::max((d.Ho - 1) * stride + R - H, 0) / 2;
  d.pad_l = std::max((d.Wo - 1) * stride + S - W, 0) / 2;
  return d;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Conv2D(padding="same") under SpectralNormalization: common.py:41-47,57-61,152-163; encoder.py:12,61-63,92-98;
// decoder.py:110-112,126-134 (+ tf.concat decoder.py:115 as the second source, + the residual add encoder.py:16)
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeConv2dFwd")
    .Input("x: float").Input("x2: float").Input("w: float").Input("w_tr: float").Input("bias: float")
    .Input("residual: float")
    .Attr("stride: int = 1").Attr("precision: int = 2")
    .Output("y: float");

class NvaeConv2dFwdOp : public OpKernel {
 public:
  explicit NvaeConv2dFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("stride", &stride_));
    OP_REQUIRES_OK(c, c->GetAttr("precision", &precision_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &x2 = ctx->input(1), &w = ctx->input(2), &w_tr = ctx->input(3),
                 &bias = ctx->input(4), &res = ctx->input(5);
    const NvaeConvDesc d = conv_desc(x.dim_size(0), x.dim_size(1), x.dim_size(2), x.dim_size(3),
                                     x2.NumElements() ? x2.dim_size(3) : 0, w.dim_size(3), w.dim_size(0), w.dim_size(1),
                                     stride_, precision_);
    Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({d.N, d.Ho, d.Wo, d.Cout}), &y));
    Tensor ws;
    const size_t ws_bytes = nvae_conv2d_ws_bytes(&d, 0);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_conv2d_fwd(&d, fptr(x), fptr(x2), fptr(w), fptr(w_tr), fptr(bias), fptr(res), fptr(y),
                                   ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_conv2d_fwd failed: ", rc));
  }
 private:
  int stride_, precision_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeConv2dFwd").Device(DEVICE_GPU), NvaeConv2dFwdOp);

REGISTER_OP("NvaeConv2dDgrad")
    .Input("dy: float").Input("w: float").Input("w_rnd: float")
    .Attr("in_h: int").Attr("in_w: int").Attr("cin: int").Attr("cin2: int = 0")
    .Attr("stride: int = 1").Attr("precision: int = 2")
    .Output("dx: float").Output("dx2: float");

class NvaeConv2dDgradOp : public OpKernel {
 public:
  explicit NvaeConv2dDgradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("in_h", &h_)); OP_REQUIRES_OK(c, c->GetAttr("in_w", &w_));
    OP_REQUIRES_OK(c, c->GetAttr("cin", &cin_)); OP_REQUIRES_OK(c, c->GetAttr("cin2", &cin2_));
    OP_REQUIRES_OK(c, c->GetAttr("stride", &stride_)); OP_REQUIRES_OK(c, c->GetAttr("precision", &precision_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &dy = ctx->input(0), &w = ctx->input(1), &w_rnd = ctx->input(2);
    const NvaeConvDesc d = conv_desc(dy.dim_size(0), h_, w_, cin_, cin2_, w.dim_size(3), w.dim_size(0), w.dim_size(1),
                                     stride_, precision_);
    Tensor *dx = nullptr, *dx2 = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({d.N, d.H, d.W, d.Cin}), &dx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({d.N, d.H, d.W, d.Cin2}), &dx2));
    Tensor ws;
    const size_t ws_bytes = nvae_conv2d_ws_bytes(&d, 1);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_conv2d_dgrad(&d, fptr(dy), fptr(w), fptr(w_rnd), fptr(dx), fptr(dx2), /*accumulate=*/0,
                                     ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_conv2d_dgrad failed: ", rc));
  }
 private:
  int h_, w_, cin_, cin2_, stride_, precision_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeConv2dDgrad").Device(DEVICE_GPU), NvaeConv2dDgradOp);

REGISTER_OP("NvaeConv2dWgrad")
    .Input("x: float").Input("x2: float").Input("dy: float")
    .Attr("r: int").Attr("s: int").Attr("stride: int = 1").Attr("precision: int = 2")
    .Output("dw: float").Output("dbias: float");

class NvaeConv2dWgradOp : public OpKernel {
 public:
  explicit NvaeConv2dWgradOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("r", &r_)); OP_REQUIRES_OK(c, c->GetAttr("s", &s_));
    OP_REQUIRES_OK(c, c->GetAttr("stride", &stride_)); OP_REQUIRES_OK(c, c->GetAttr("precision", &precision_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &x2 = ctx->input(1), &dy = ctx->input(2);
    const int cin2 = x2.NumElements() ? x2.dim_size(3) : 0;
    const NvaeConvDesc d = conv_desc(x.dim_size(0), x.dim_size(1), x.dim_size(2), x.dim_size(3), cin2, dy.dim_size(3),
                                     r_, s_, stride_, precision_);
    Tensor *dw = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({d.R, d.S, d.Cin + d.Cin2, d.Cout}), &dw));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({d.Cout}), &db));
    Tensor ws;
    const size_t ws_bytes = nvae_conv2d_ws_bytes(&d, 2);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_conv2d_wgrad(&d, fptr(x), fptr(x2), fptr(dy), fptr(dw), fptr(db), ws.flat<uint8>().data(),
                                     ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_conv2d_wgrad failed: ", rc));
  }
 private:
  int r_, s_, stride_, precision_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeConv2dWgrad").Device(DEVICE_GPU), NvaeConv2dWgradOp);

// ------------------------------------------------------------------------------------------------
// BN -> [swish] -> 1x1 Conv2D with the BN-apply + activation inside the convolution (decoder.py:125-127, 143-144;
// postprocess.py:71-73, 84-96).  stat is the [4, Cin] block of NvaeBnStats / NvaeBnFwd.  The launcher refuses shapes
// the fused operand path does not take (rc = NVAE_E_UNSUPPORTED): nvae_tf_layers.bn_conv2d checks
// nvae_conv2d_bnact_supported first and composes bn_act + conv2d otherwise.
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeConv2dFwdBnact")
    .Input("x: float").Input("stat: float").Input("w: float").Input("w_tr: float").Input("bias: float")
    .Input("residual: float")
    .Attr("act: int = 1").Attr("precision: int = 2")
    .Output("y: float");

class NvaeConv2dFwdBnactOp : public OpKernel {
 public:
  explicit NvaeConv2dFwdBnactOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("precision", &precision_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &stat = ctx->input(1), &w = ctx->input(2), &w_tr = ctx->input(3),
                 &bias = ctx->input(4), &res = ctx->input(5);
    const NvaeConvDesc d = conv_desc(x.dim_size(0), x.dim_size(1), x.dim_size(2), x.dim_size(3), 0, w.dim_size(3),
                                     w.dim_size(0), w.dim_size(1), 1, precision_);
    OP_REQUIRES(ctx, nvae_conv2d_bnact_supported(&d) == 1,
                errors::InvalidArgument("NvaeConv2dFwdBnact: shape not on the fused path; use NvaeBnFwd + NvaeConv2dFwd"));
    Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({d.N, d.Ho, d.Wo, d.Cout}), &y));
    Tensor ws;
    const size_t ws_bytes = nvae_conv2d_ws_bytes(&d, 0);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_conv2d_fwd_bnact(&d, fptr(x), fptr(stat), act_, fptr(w_tr), fptr(bias), fptr(res), fptr(y),
                                         ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_conv2d_fwd_bnact failed: ", rc));
  }
 private:
  int act_, precision_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeConv2dFwdBnact").Device(DEVICE_GPU), NvaeConv2dFwdBnactOp);

REGISTER_OP("NvaeConv2dWgradBnact")
    .Input("x: float").Input("stat: float").Input("dy: float")
    .Attr("act: int = 1").Attr("precision: int = 2")
    .Output("dw: float").Output("dbias: float");

class NvaeConv2dWgradBnactOp : public OpKernel {
 public:
  explicit NvaeConv2dWgradBnactOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("precision", &precision_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &stat = ctx->input(1), &dy = ctx->input(2);
    const NvaeConvDesc d = conv_desc(x.dim_size(0), x.dim_size(1), x.dim_size(2), x.dim_size(3), 0, dy.dim_size(3), 1, 1,
                                     1, precision_);
    Tensor *dw = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({1, 1, d.Cin, d.Cout}), &dw));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({d.Cout}), &db));
    Tensor ws;
    const size_t ws_bytes = nvae_conv2d_ws_bytes(&d, 2);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_conv2d_wgrad_bnact(&d, fptr(x), fptr(stat), act_, fptr(dy), fptr(dw), fptr(db),
                                           ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_conv2d_wgrad_bnact failed: ", rc));
  }
 private:
  int act_, precision_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeConv2dWgradBnact").Device(DEVICE_GPU), NvaeConv2dWgradBnactOp);

// ------------------------------------------------------------------------------------------------
// BatchNormalization(momentum=0.05, epsilon=1e-5) + swish / ELU (+ nearest x2): common.py:148,165-172;
// encoder.py:91-104; decoder.py:125-145.  moving_mean / moving_var are resource-style in/out buffers: the op updates
// the tensors it is given in place (they are passed as ref-like inputs by the Python wrapper).
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeBnFwd")
    .Input("x: float").Input("gamma: float").Input("beta: float").Input("moving_mean: float").Input("moving_var: float")
    .Attr("training: bool = true").Attr("momentum: float = 0.05").Attr("epsilon: float = 1e-5")
    .Attr("act: int = 1").Attr("upsample: bool = false")
    .Output("out: float").Output("stat: float");

class NvaeBnFwdOp : public OpKernel {
 public:
  explicit NvaeBnFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_)); OP_REQUIRES_OK(c, c->GetAttr("momentum", &momentum_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_)); OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("upsample", &up_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &gamma = ctx->input(1), &beta = ctx->input(2);
    Tensor mm = ctx->input(3), mv = ctx->input(4);  // shares the buffers: updated in place
    const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
    Tensor *out = nullptr, *stat = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({N, up_ ? 2 * H : H, up_ ? 2 * W : W, C}), &out));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({4, C}), &stat));
    Tensor ws;
    const size_t ws_bytes = nvae_bn_ws_bytes(rows_of(x), C);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_bn_fwd(fptr(x), rows_of(x), C, fptr(gamma), fptr(beta), fptr(&mm), fptr(&mv), training_ ? 1 : 0,
                               momentum_, eps_, fptr(stat), act_, up_ ? H : 0, up_ ? W : 0, /*round_tf32=*/0, fptr(out),
                               ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bn_fwd failed: ", rc));
  }
 private:
  bool training_, up_;
  float momentum_, eps_;
  int act_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBnFwd").Device(DEVICE_GPU), NvaeBnFwdOp);

REGISTER_OP("NvaeBnActBwd")
    .Input("dout: float").Input("x: float").Input("stat: float")
    .Attr("training: bool = true").Attr("act: int = 1").Attr("upsample: bool = false")
    .Output("dx: float").Output("dgamma: float").Output("dbeta: float");

class NvaeBnActBwdOp : public OpKernel {
 public:
  explicit NvaeBnActBwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_)); OP_REQUIRES_OK(c, c->GetAttr("act", &act_));
    OP_REQUIRES_OK(c, c->GetAttr("upsample", &up_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &dout = ctx->input(0), &x = ctx->input(1), &stat = ctx->input(2);
    const int H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
    Tensor *dx = nullptr, *dg = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &dx));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({C}), &dg));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({C}), &db));
    Tensor ws;
    const size_t ws_bytes = nvae_bn_ws_bytes(rows_of(x), C);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_bn_act_bwd(fptr(dout), fptr(x), rows_of(x), C, fptr(stat), act_, up_ ? H : 0, up_ ? W : 0,
                                   training_ ? 1 : 0, /*dres=*/nullptr, 0.f, /*accumulate=*/0, fptr(dx), fptr(dg), fptr(db),
                                   ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bn_act_bwd failed: ", rc));
  }
 private:
  bool training_, up_;
  int act_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBnActBwd").Device(DEVICE_GPU), NvaeBnActBwdOp);

// ------------------------------------------------------------------------------------------------
// DepthwiseConv2D((5,5)) with the BN-apply + swish of its input fused into the load: decoder.py:130,141-142
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeDwconv5x5Fwd")
    .Input("x: float").Input("stat: float").Input("w: float").Input("bias: float").Attr("act: int = 1")
    .Output("y: float");

class NvaeDwconv5x5FwdOp : public OpKernel {
 public:
  explicit NvaeDwconv5x5FwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("act", &act_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &stat = ctx->input(1), &w = ctx->input(2), &bias = ctx->input(3);
    Tensor* y = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &y));
    const int rc = nvae_dwconv5x5_fwd(fptr(x), fptr(stat), act_, x.dim_size(0), x.dim_size(1), x.dim_size(2),
                                      x.dim_size(3), fptr(w), fptr(bias), fptr(y), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_dwconv5x5_fwd failed: ", rc));
  }
 private:
  int act_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeDwconv5x5Fwd").Device(DEVICE_GPU), NvaeDwconv5x5FwdOp);

REGISTER_OP("NvaeDwconv5x5Bwd")
    .Input("x: float").Input("stat: float").Input("w: float").Input("dy: float").Attr("act: int = 1")
    .Output("da: float").Output("dw: float").Output("dbias: float");

class NvaeDwconv5x5BwdOp : public OpKernel {
 public:
  explicit NvaeDwconv5x5BwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("act", &act_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &stat = ctx->input(1), &w = ctx->input(2), &dy = ctx->input(3);
    const int N = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3);
    Tensor *da = nullptr, *dw = nullptr, *db = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, x.shape(), &da));  // gradient w.r.t. the ACTIVATED input
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, w.shape(), &dw));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({C}), &db));
    Tensor ws;
    const size_t ws_bytes = nvae_dwconv5x5_bwd_filter_ws_bytes(N, H, W, C);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    int rc = nvae_dwconv5x5_bwd_data(fptr(dy), N, H, W, C, fptr(w), fptr(da), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_dwconv5x5_bwd_data failed: ", rc));
    rc = nvae_dwconv5x5_bwd_filter(fptr(x), fptr(stat), act_, fptr(dy), N, H, W, C, fptr(dw), fptr(db),
                                   ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_dwconv5x5_bwd_filter failed: ", rc));
  }
 private:
  int act_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeDwconv5x5Bwd").Device(DEVICE_GPU), NvaeDwconv5x5BwdOp);

// ------------------------------------------------------------------------------------------------
// SqueezeExcitation + residual merge: common.py:129-142 fused with encoder.py:107 / decoder.py:147
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeSeFwd")
    .Input("t: float").Input("stat: float").Input("xres: float")
    .Input("w1: float").Input("b1: float").Input("w2: float").Input("b2: float")
    .Attr("alpha: float = 0.1").Attr("beta: float = 1.0")
    .Output("y: float").Output("pooled: float").Output("hidden: float").Output("gate: float");

class NvaeSeFwdOp : public OpKernel {
 public:
  explicit NvaeSeFwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("alpha", &alpha_)); OP_REQUIRES_OK(c, c->GetAttr("beta", &beta_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &t = ctx->input(0), &stat = ctx->input(1), &xres = ctx->input(2), &w1 = ctx->input(3),
                 &b1 = ctx->input(4), &w2 = ctx->input(5), &b2 = ctx->input(6);
    const int B = t.dim_size(0), HW = t.dim_size(1) * t.dim_size(2), C = t.dim_size(3), hid = w1.dim_size(1);
    Tensor *y = nullptr, *pooled = nullptr, *hidden = nullptr, *gate = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, t.shape(), &y));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B, C}), &pooled));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({B, hid}), &hidden));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({B, C}), &gate));
    const int rc = nvae_se_fwd(fptr(t), fptr(stat), fptr(xres), B, HW, C, hid, fptr(w1), fptr(b1), fptr(w2), fptr(b2),
                               alpha_, beta_, fptr(pooled), fptr(hidden), fptr(gate), fptr(y), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_se_fwd failed: ", rc));
  }
 private:
  float alpha_, beta_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeSeFwd").Device(DEVICE_GPU), NvaeSeFwdOp);

REGISTER_OP("NvaeSeBwd")
    .Input("dy: float").Input("t: float").Input("stat: float").Input("w1: float").Input("w2: float")
    .Input("pooled: float").Input("hidden: float").Input("gate: float")
    .Attr("alpha: float = 0.1").Attr("beta: float = 1.0")
    .Output("dt: float").Output("dxres: float").Output("dw1: float").Output("db1: float").Output("dw2: float")
    .Output("db2: float");

class NvaeSeBwdOp : public OpKernel {
 public:
  explicit NvaeSeBwdOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("alpha", &alpha_)); OP_REQUIRES_OK(c, c->GetAttr("beta", &beta_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &dy = ctx->input(0), &t = ctx->input(1), &stat = ctx->input(2), &w1 = ctx->input(3), &w2 = ctx->input(4),
                 &pooled = ctx->input(5), &hidden = ctx->input(6), &gate = ctx->input(7);
    const int B = t.dim_size(0), HW = t.dim_size(1) * t.dim_size(2), C = t.dim_size(3), hid = w1.dim_size(1);
    Tensor *dt = nullptr, *dxr = nullptr, *dw1 = nullptr, *db1 = nullptr, *dw2 = nullptr, *db2 = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, t.shape(), &dt));   // gradient w.r.t. t' (post-affine)
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, t.shape(), &dxr));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, w1.shape(), &dw1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(3, TensorShape({hid}), &db1));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(4, w2.shape(), &dw2));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(5, TensorShape({C}), &db2));
    Tensor ws;
    const size_t ws_bytes = nvae_se_bwd_ws_bytes(B, C, hid);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_se_bwd(fptr(dy), fptr(t), fptr(stat), B, HW, C, hid, fptr(w1), fptr(w2), fptr(pooled),
                               fptr(hidden), fptr(gate), alpha_, beta_, fptr(dt), fptr(dxr), /*dxres_accumulate=*/0,
                               fptr(dw1), fptr(db1), fptr(dw2), fptr(db2), ws.flat<uint8>().data(), ws_bytes,
                               stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_se_bwd failed: ", rc));
  }
 private:
  float alpha_, beta_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeSeBwd").Device(DEVICE_GPU), NvaeSeBwdOp);

// ------------------------------------------------------------------------------------------------
// Per-group latent math: Sampler.call common.py:76-102, softclamp5 util.py:49-50, KL row sums models.py:197-201
// ------------------------------------------------------------------------------------------------
REGISTER_OP("NvaeLatentFwd")
    .Input("enc_p: float").Input("dec_p: float").Input("eps: float")
    .Output("z: float").Output("kl: float").Output("dist: float");

class NvaeLatentFwdOp : public OpKernel {
 public:
  explicit NvaeLatentFwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &enc = ctx->input(0), &dec = ctx->input(1), &eps = ctx->input(2);  // dec empty: z_idx == 0 branch
    const int B = enc.dim_size(0), HW = enc.dim_size(1) * enc.dim_size(2), L = enc.dim_size(3) / 2;
    Tensor *z = nullptr, *kl = nullptr, *dist = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, eps.shape(), &z));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B}), &kl));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({4, B, HW, L}), &dist));
    const int rc = nvae_latent_fwd(fptr(enc), fptr(dec), fptr(eps), B, HW, L, fptr(z), fptr(kl), nullptr, nullptr,
                                   fptr(dist), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_latent_fwd failed: ", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("NvaeLatentFwd").Device(DEVICE_GPU), NvaeLatentFwdOp);

REGISTER_OP("NvaeLatentBwd")
    .Input("enc_p: float").Input("dec_p: float").Input("eps: float").Input("dz: float").Input("kl_weight: float")
    .Output("d_enc_p: float").Output("d_dec_p: float");

class NvaeLatentBwdOp : public OpKernel {
 public:
  explicit NvaeLatentBwdOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &enc = ctx->input(0), &dec = ctx->input(1), &eps = ctx->input(2), &dz = ctx->input(3),
                 &klw = ctx->input(4);
    const int B = enc.dim_size(0), HW = enc.dim_size(1) * enc.dim_size(2), L = enc.dim_size(3) / 2;
    Tensor *de = nullptr, *dd = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, enc.shape(), &de));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, dec.shape(), &dd));
    const int rc = nvae_latent_bwd(fptr(enc), fptr(dec), fptr(eps), fptr(dz), fptr(klw), B, HW, L, fptr(de), fptr(dd),
                                   stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_latent_bwd failed: ", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("NvaeLatentBwd").Device(DEVICE_GPU), NvaeLatentBwdOp);

// ------------------------------------------------------------------------------------------------
// The once-per-step launchers around the cells: spectral normalisation of every wrapped conv, losses, optimizer, noise.
// ------------------------------------------------------------------------------------------------
// tfa.layers.SpectralNormalization(power_iterations=1) for ALL layers of the model in one op (4 launches) + the operand
// repack the tensor-core convolutions consume.  `params` / `state` are the flat fp32 arenas the variables are views of,
// `layers` the NvaeSnLayer table (bytes), `chunk_layer` its row-chunk index (see Runtime._build_sn_tables).  Updates
// params (W /= sigma), state (u) and pack IN PLACE, like the reference's kernel.assign / u.assign; the op's
// output `sigma` gives the step something to take a control dependency on before the first conv.
REGISTER_OP("NvaeSpectralNorm")
    .Input("params: Ref(float)").Input("state: Ref(float)").Input("pack: Ref(float)").Input("layers: uint8")
    .Input("chunk_layer: int32").Input("ws: Ref(float)")
    .Attr("power_iter: bool = true").Attr("pack_exact: bool = true")
    .Output("sigma: float");

class NvaeSpectralNormOp : public OpKernel {
 public:
  explicit NvaeSpectralNormOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("power_iter", &power_iter_));
    OP_REQUIRES_OK(c, c->GetAttr("pack_exact", &pack_exact_));
  }
  void Compute(OpKernelContext* ctx) override {
    Tensor params = ctx->mutable_input(0, true), state = ctx->mutable_input(1, true), pack = ctx->mutable_input(2, true),
           ws = ctx->mutable_input(5, true);
    const Tensor &layers = ctx->input(3), &chunk_layer = ctx->input(4);
    const int n_layers = static_cast<int>(layers.NumElements() / sizeof(NvaeSnLayer));
    Tensor* sigma = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({n_layers}), &sigma));
    const int rc = nvae_spectral_norm(fptr(&params), fptr(&state), fptr(&pack),
                                      reinterpret_cast<const NvaeSnLayer*>(layers.flat<uint8>().data()), n_layers,
                                      chunk_layer.flat<int32>().data(), static_cast<int>(chunk_layer.NumElements()),
                                      power_iter_ ? 1 : 0, pack_exact_ ? 1 : 0, fptr(sigma), fptr(&ws), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_spectral_norm failed: ", rc));
  }
 private:
  bool power_iter_, pack_exact_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeSpectralNorm").Device(DEVICE_GPU), NvaeSpectralNormOp);

// Batch statistics only ([4,C] = mean, invstd, scale, shift; moving statistics updated): consumers that fuse the apply
REGISTER_OP("NvaeBnStats")
    .Input("x: float").Input("gamma: float").Input("beta: float").Input("moving_mean: Ref(float)")
    .Input("moving_var: Ref(float)")
    .Attr("training: bool = true").Attr("momentum: float = 0.05").Attr("epsilon: float = 1e-5")
    .Output("stat: float");

class NvaeBnStatsOp : public OpKernel {
 public:
  explicit NvaeBnStatsOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("training", &training_));
    OP_REQUIRES_OK(c, c->GetAttr("momentum", &momentum_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &x = ctx->input(0), &gamma = ctx->input(1), &beta = ctx->input(2);
    Tensor mm = ctx->mutable_input(3, true), mv = ctx->mutable_input(4, true);
    const int C = static_cast<int>(x.dim_size(x.dims() - 1));
    Tensor* stat = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({4, C}), &stat));
    Tensor ws;
    const size_t ws_bytes = nvae_bn_ws_bytes(rows_of(x), C);
    OP_REQUIRES_OK(ctx, alloc_ws(ctx, ws_bytes, &ws));
    const int rc = nvae_bn_stats(fptr(x), rows_of(x), C, fptr(gamma), fptr(beta), fptr(&mm), fptr(&mv), training_ ? 1 : 0,
                                 momentum_, eps_, fptr(stat), ws.flat<uint8>().data(), ws_bytes, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bn_stats failed: ", rc));
  }
 private:
  bool training_;
  float momentum_, eps_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBnStats").Device(DEVICE_GPU), NvaeBnStatsOp);

// Bernoulli(logits).log_prob(x) summed over H,W,C (models.py:242-250) and its gradient w.r.t. the logits
REGISTER_OP("NvaeBernoulliLlFwd")
    .Input("logits: float").Input("x: float").Attr("crop: int = 0")
    .Output("recon: float");

class NvaeBernoulliLlFwdOp : public OpKernel {
 public:
  explicit NvaeBernoulliLlFwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("crop", &crop_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &l = ctx->input(0), &x = ctx->input(1);
    const int B = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3), Cl = l.dim_size(3);
    Tensor* recon = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B}), &recon));
    const int rc = nvae_bernoulli_ll_fwd(fptr(l), fptr(x), B, H, W, C, Cl, crop_, fptr(recon), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bernoulli_ll_fwd failed: ", rc));
  }
 private:
  int crop_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBernoulliLlFwd").Device(DEVICE_GPU), NvaeBernoulliLlFwdOp);

REGISTER_OP("NvaeBernoulliLlBwd")
    .Input("logits: float").Input("x: float").Attr("scale: float = 1.0")
    .Output("dlogits: float");

class NvaeBernoulliLlBwdOp : public OpKernel {
 public:
  explicit NvaeBernoulliLlBwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("scale", &scale_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &l = ctx->input(0), &x = ctx->input(1);
    const int B = x.dim_size(0), H = x.dim_size(1), W = x.dim_size(2), C = x.dim_size(3), Cl = l.dim_size(3);
    Tensor* dl = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, l.shape(), &dl));
    const int rc = nvae_bernoulli_ll_bwd(fptr(l), fptr(x), B, H, W, C, Cl, scale_, fptr(dl), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bernoulli_ll_bwd failed: ", rc));
  }
 private:
  float scale_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBernoulliLlBwd").Device(DEVICE_GPU), NvaeBernoulliLlBwdOp);

// KL balancing + beta warm-up + batch mean + total (models.py:121-126, 204-222) in one launch
REGISTER_OP("NvaeLossAssemble")
    .Input("kl_all: float").Input("recon: float").Input("bn_loss: float").Input("alphas: float").Input("hyper: float")
    .Attr("balancing: int = -1")
    .Output("kl_weight: float").Output("kl_loss: float").Output("scalars: float");

class NvaeLossAssembleOp : public OpKernel {
 public:
  explicit NvaeLossAssembleOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("balancing", &bal_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &kl_all = ctx->input(0), &recon = ctx->input(1), &bn_loss = ctx->input(2), &alphas = ctx->input(3),
                 &hyper = ctx->input(4);
    const int G = kl_all.dim_size(0), B = kl_all.dim_size(1);
    Tensor *klw = nullptr, *kl = nullptr, *sc = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({G}), &klw));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({B}), &kl));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(2, TensorShape({2}), &sc));
    const int rc = nvae_loss_assemble(fptr(kl_all), fptr(recon), fptr(bn_loss), fptr(alphas), fptr(hyper), bal_, G, B,
                                      fptr(klw), fptr(kl), fptr(sc), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_loss_assemble failed: ", rc));
  }
 private:
  int bal_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeLossAssemble").Device(DEVICE_GPU), NvaeLossAssembleOp);

// sr_lambda * sum_k max|gamma_k| over the 88 encoder/decoder-group BN layers (models.py:252-267) and its sub-gradient
REGISTER_OP("NvaeBnLossFwd")
    .Input("params: float").Input("offsets: int64").Input("sizes: int32").Attr("sr_lambda: float")
    .Output("loss: float");

class NvaeBnLossFwdOp : public OpKernel {
 public:
  explicit NvaeBnLossFwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("sr_lambda", &lam_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &params = ctx->input(0), &off = ctx->input(1), &sz = ctx->input(2);
    Tensor* loss = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({1}), &loss));
    const int rc = nvae_bn_loss_fwd(fptr(params), reinterpret_cast<const int64_t*>(off.flat<int64>().data()),
                                    sz.flat<int32>().data(), static_cast<int>(off.NumElements()), lam_, fptr(loss),
                                    stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bn_loss_fwd failed: ", rc));
  }
 private:
  float lam_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBnLossFwd").Device(DEVICE_GPU), NvaeBnLossFwdOp);

REGISTER_OP("NvaeBnLossBwd")
    .Input("params: float").Input("grads: Ref(float)").Input("offsets: int64").Input("sizes: int32")
    .Attr("sr_lambda: float")
    .Output("done: float");

class NvaeBnLossBwdOp : public OpKernel {
 public:
  explicit NvaeBnLossBwdOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("sr_lambda", &lam_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &params = ctx->input(0), &off = ctx->input(2), &sz = ctx->input(3);
    Tensor grads = ctx->mutable_input(1, true);
    Tensor* done = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({0}), &done));
    const int rc = nvae_bn_loss_bwd(fptr(params), fptr(&grads), reinterpret_cast<const int64_t*>(off.flat<int64>().data()),
                                    sz.flat<int32>().data(), static_cast<int>(off.NumElements()), lam_, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_bn_loss_bwd failed: ", rc));
  }
 private:
  float lam_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeBnLossBwd").Device(DEVICE_GPU), NvaeBnLossBwdOp);

// optimizers.Adamax + CosineDecay + the beta warm-up (train.py:128-131, models.py:121-122,128-129): the schedule launch
// derives {beta, lr_t, ...} from device counters, the multi-tensor Adamax updates the whole parameter arena in one launch
REGISTER_OP("NvaeScheduleStep")
    .Input("counters: Ref(int64)")
    .Attr("warmup_iters: float").Attr("lr0: float").Attr("decay_steps: float").Attr("beta_1: float = 0.9")
    .Attr("advance: int = 3")
    .Output("hyper: float");

class NvaeScheduleStepOp : public OpKernel {
 public:
  explicit NvaeScheduleStepOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("warmup_iters", &warm_)); OP_REQUIRES_OK(c, c->GetAttr("lr0", &lr0_));
    OP_REQUIRES_OK(c, c->GetAttr("decay_steps", &decay_)); OP_REQUIRES_OK(c, c->GetAttr("beta_1", &b1_));
    OP_REQUIRES_OK(c, c->GetAttr("advance", &adv_));
  }
  void Compute(OpKernelContext* ctx) override {
    Tensor counters = ctx->mutable_input(0, true);
    Tensor* hyper = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({8}), &hyper));
    const int rc = nvae_schedule_step(reinterpret_cast<int64_t*>(counters.flat<int64>().data()), fptr(hyper), warm_, lr0_,
                                      decay_, b1_, adv_, stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_schedule_step failed: ", rc));
  }
 private:
  float warm_, lr0_, decay_, b1_;
  int adv_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeScheduleStep").Device(DEVICE_GPU), NvaeScheduleStepOp);

REGISTER_OP("NvaeAdamax")
    .Input("params: Ref(float)").Input("grads: float").Input("m: Ref(float)").Input("v: Ref(float)").Input("hyper: float")
    .Attr("beta_1: float = 0.9").Attr("beta_2: float = 0.999").Attr("epsilon: float = 1e-7").Attr("grad_scale: float = 1.0")
    .Output("done: float");

class NvaeAdamaxOp : public OpKernel {
 public:
  explicit NvaeAdamaxOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("beta_1", &b1_)); OP_REQUIRES_OK(c, c->GetAttr("beta_2", &b2_));
    OP_REQUIRES_OK(c, c->GetAttr("epsilon", &eps_)); OP_REQUIRES_OK(c, c->GetAttr("grad_scale", &gs_));
  }
  void Compute(OpKernelContext* ctx) override {
    Tensor p = ctx->mutable_input(0, true), m = ctx->mutable_input(2, true), v = ctx->mutable_input(3, true);
    const Tensor &g = ctx->input(1), &hyper = ctx->input(4);
    Tensor* done = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({0}), &done));
    const int rc = nvae_adamax(fptr(&p), fptr(g), fptr(&m), fptr(&v), p.NumElements(), fptr(hyper), b1_, b2_, eps_, gs_,
                               stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_adamax failed: ", rc));
  }
 private:
  float b1_, b2_, eps_, gs_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeAdamax").Device(DEVICE_GPU), NvaeAdamaxOp);

// epsilon ~ N(0,1) (common.py:67) from Philox keyed by (seed, device iteration counter, stream id); z = mu + eps * sigma
REGISTER_OP("NvaePhiloxNormal")
    .Input("counters: int64").Attr("n: int").Attr("seed: int").Attr("stream_id: int = 0")
    .Output("eps: float");

class NvaePhiloxNormalOp : public OpKernel {
 public:
  explicit NvaePhiloxNormalOp(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("n", &n_)); OP_REQUIRES_OK(c, c->GetAttr("seed", &seed_));
    OP_REQUIRES_OK(c, c->GetAttr("stream_id", &sid_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& counters = ctx->input(0);
    Tensor* eps = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({static_cast<int64_t>(n_)}), &eps));
    const int rc = nvae_philox_normal(fptr(eps), n_, static_cast<uint64_t>(seed_),
                                      reinterpret_cast<const int64_t*>(counters.flat<int64>().data()),
                                      static_cast<uint64_t>(sid_), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_philox_normal failed: ", rc));
  }
 private:
  int n_, seed_, sid_;
};
REGISTER_KERNEL_BUILDER(Name("NvaePhiloxNormal").Device(DEVICE_GPU), NvaePhiloxNormalOp);

REGISTER_OP("NvaeReparam")
    .Input("mu: float").Input("sigma: float").Input("eps: float").Attr("sigma_scale: float = 1.0")
    .Output("z: float");

class NvaeReparamOp : public OpKernel {
 public:
  explicit NvaeReparamOp(OpKernelConstruction* c) : OpKernel(c) { OP_REQUIRES_OK(c, c->GetAttr("sigma_scale", &ss_)); }
  void Compute(OpKernelContext* ctx) override {
    const Tensor &mu = ctx->input(0), &sigma = ctx->input(1), &eps = ctx->input(2);
    Tensor* z = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, mu.shape(), &z));
    const int rc = nvae_reparam(fptr(mu), fptr(sigma), fptr(eps), ss_, fptr(z), mu.NumElements(), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_reparam failed: ", rc));
  }
 private:
  float ss_;
};
REGISTER_KERNEL_BUILDER(Name("NvaeReparam").Device(DEVICE_GPU), NvaeReparamOp);

// Importance-weighted NLL bound over K attempts (evaluate.py:111-123): rows k of recon / log_q / log_p -> nll[1]
REGISTER_OP("NvaeIwaeNll")
    .Input("recon: float").Input("log_q: float").Input("log_p: float")
    .Output("per_sample: float").Output("nll: float");

class NvaeIwaeNllOp : public OpKernel {
 public:
  explicit NvaeIwaeNllOp(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor &recon = ctx->input(0), &lq = ctx->input(1), &lp = ctx->input(2);
    const int K = recon.dim_size(0), B = recon.dim_size(1);
    Tensor *ps = nullptr, *nll = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({B}), &ps));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({1}), &nll));
    const int rc = nvae_iwae_nll(fptr(recon), fptr(lq), fptr(lp), K, B, fptr(ps), fptr(nll), stream_of(ctx));
    OP_REQUIRES(ctx, rc == 0, errors::Internal("nvae_iwae_nll failed: ", rc));
  }
};
REGISTER_KERNEL_BUILDER(Name("NvaeIwaeNll").Device(DEVICE_GPU), NvaeIwaeNllOp);
