// TEST STUB -- not XLA.  Mirrors the SHAPE of the part of xla/ffi/api/ffi.h (jaxlib) that csrc/a2m_xla_ffi.cc uses, so that the
// handler source is at least type-checked in an image without jaxlib (tests/test_host.py::test_xla_ffi_source_type_checks).
// It binds nothing and decodes nothing; the real header is required to build the shim (build.py:build_xla_ffi).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <utility>

#include "xla/ffi/api/c_api.h"

namespace xla::ffi {

enum class DataType { F32, U8 };
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType U8 = DataType::U8;

enum class ErrorCode { kOk, kInvalidArgument, kInternal };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <class T>
struct Span {
  const T* ptr = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return ptr[i]; }
};

template <DataType dtype>
struct NativeOf;
template <>
struct NativeOf<DataType::F32> { using type = float; };
template <>
struct NativeOf<DataType::U8> { using type = uint8_t; };

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeOf<dtype>::type;
  T* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
  size_t size_bytes() const { return count_ * sizeof(T); }

 private:
  T* data_ = nullptr;
  Span<int64_t> dims_;
  size_t count_ = 0;
};

template <class T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }

 private:
  T value_;
};
template <DataType dtype>
using ResultBuffer = Result<Buffer<dtype>>;

template <class T>
struct PlatformStream {};

template <class... Ts>
struct Binding {
  template <class T>
  Binding<Ts..., T> Ctx() const { return {}; }
  template <class T>
  Binding<Ts..., T> Arg() const { return {}; }
  template <class T>
  Binding<Ts..., Result<T>> Ret() const { return {}; }
  template <class T>
  Binding<Ts..., T> Attr(const char*) const { return {}; }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

// the real macro instantiates a handler that decodes the call frame into the Impl's parameter list; the stub only checks that the
// Impl is callable with the types the binding declares (PlatformStream<S> decodes to S)
template <class T>
struct Decoded { using type = T; };
template <class S>
struct Decoded<PlatformStream<S>> { using type = S; };

template <class Fn, class... Ts>
constexpr bool CheckCallable(Fn fn, Binding<Ts...>) {
  using R = decltype(fn(std::declval<typename Decoded<Ts>::type>()...));
  static_assert(sizeof(R) > 0, "handler not callable with the bound types");
  return true;
}

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                      \
  static const bool name##_checked = ::xla::ffi::CheckCallable(impl, binding);  \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
