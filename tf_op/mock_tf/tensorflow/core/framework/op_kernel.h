// MINIMAL stand-in for the TensorFlow C++ API, only for `make -C tf_op syntax` in an image without TensorFlow: just enough
// declarations (no behaviour) for tf_op/nvae_ops.cc to be parsed and type-checked against include/nvae_b200.h.
// The real build (tf_op/Makefile, target `all`) uses TensorFlow's own headers and never sees this directory.
#pragma once
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <string>

namespace Eigen {
struct GpuDevice {
  void* stream() const { return nullptr; }
};
}  // namespace Eigen

namespace tensorflow {
typedef std::uint8_t uint8;
typedef std::int64_t int64;
typedef std::int32_t int32;
enum DataType { DT_FLOAT = 1, DT_UINT8 = 4 };
constexpr const char* DEVICE_GPU = "GPU";

class Status {
 public:
  bool ok() const { return true; }
};
namespace errors {
template <typename... A>
Status Internal(A...) { return Status(); }
template <typename... A>
Status InvalidArgument(A...) { return Status(); }
}  // namespace errors

class TensorShape {
 public:
  TensorShape() {}
  TensorShape(std::initializer_list<std::int64_t>) {}
};

template <typename T>
struct Flat {
  T* data() const { return nullptr; }
};

class Tensor {
 public:
  std::int64_t dim_size(int) const { return 0; }
  int dims() const { return 0; }
  std::int64_t NumElements() const { return 0; }
  const TensorShape& shape() const { return shape_; }
  template <typename T>
  Flat<T> flat() { return Flat<T>(); }
  template <typename T>
  Flat<const T> flat() const { return Flat<const T>(); }
 private:
  TensorShape shape_;
};

class OpKernelConstruction {
 public:
  template <typename T>
  Status GetAttr(const char*, T*) const { return Status(); }
  void SetStatus(const Status&) {}
};

class OpKernelContext {
 public:
  const Tensor& input(int) { return t_; }
  Tensor mutable_input(int, bool) { return t_; }
  Status allocate_output(int, const TensorShape&, Tensor** out) { *out = &t_; return Status(); }
  Status allocate_temp(DataType, const TensorShape&, Tensor*) { return Status(); }
  template <typename D>
  const D& eigen_device() const { static D d; return d; }
  void SetStatus(const Status&) {}
  void CtxFailure(const Status&) {}
 private:
  Tensor t_;
};

class OpKernel {
 public:
  explicit OpKernel(OpKernelConstruction*) {}
  virtual ~OpKernel() {}
  virtual void Compute(OpKernelContext* ctx) = 0;
};

struct KernelDefBuilder {
  KernelDefBuilder& Device(const char*) { return *this; }
};
inline KernelDefBuilder Name(const char*) { return KernelDefBuilder(); }

#define OP_REQUIRES_OK(CTX, ...)          \
  do {                                    \
    ::tensorflow::Status s_(__VA_ARGS__); \
    if (!s_.ok()) {                       \
      (CTX)->SetStatus(s_);               \
      return;                             \
    }                                     \
  } while (0)
#define OP_REQUIRES(CTX, EXP, STATUS) \
  do {                                \
    if (!(EXP)) {                     \
      (CTX)->SetStatus(STATUS);       \
      return;                         \
    }                                 \
  } while (0)
#define NVAE_MOCK_CAT2(a, b) a##b
#define NVAE_MOCK_CAT(a, b) NVAE_MOCK_CAT2(a, b)
#define REGISTER_KERNEL_BUILDER(BUILDER, CLS) \
  static ::tensorflow::KernelDefBuilder NVAE_MOCK_CAT(kdef_, __LINE__) = (BUILDER); \
  static_assert(sizeof(CLS) > 0, "kernel class must be complete")
}  // namespace tensorflow
