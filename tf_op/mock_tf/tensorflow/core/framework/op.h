// MINIMAL stand-in (see op_kernel.h in this directory): REGISTER_OP("...").Input(...).Attr(...).Output(...);
#pragma once
namespace tensorflow {
struct OpDefBuilderMock {
  OpDefBuilderMock& Input(const char*) { return *this; }
  OpDefBuilderMock& Output(const char*) { return *this; }
  OpDefBuilderMock& Attr(const char*) { return *this; }
};
}  // namespace tensorflow
#define NVAE_MOCK_OP_CAT2(a, b) a##b
#define NVAE_MOCK_OP_CAT(a, b) NVAE_MOCK_OP_CAT2(a, b)
#define REGISTER_OP(NAME) static ::tensorflow::OpDefBuilderMock NVAE_MOCK_OP_CAT(opdef_, __LINE__) = ::tensorflow::OpDefBuilderMock()
