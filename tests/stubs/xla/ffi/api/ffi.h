// Minimal stand-in for the XLA FFI C++ header (xla/ffi/api/ffi.h of a modern jaxlib), TEST INFRASTRUCTURE ONLY.
//
// jax / jaxlib are not installed in this image, so difflexmm_b200/csrc/jax_ffi_shim.cc cannot be built against the real
// header.  This stub declares just the surface the shim uses -- Buffer / ResultBuffer, Error, PlatformStream, the
// Ffi::Bind() builder and XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the same names and shapes, and makes the handler macro
// CHECK that every handler implementation is invocable with exactly the context / attribute / argument / result types
// its binding declares (the thing that silently rots in unbuilt glue code).  It does not execute anything.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace xla {
namespace ffi {

enum DataType { U8, S32, S64, F32, F64 };
template <DataType T> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };

struct Span {
  const int64_t* p = nullptr;
  size_t n = 0;
  size_t size() const { return n; }
  int64_t operator[](size_t i) const { return p[i]; }
  int64_t back() const { return p[n - 1]; }
};

template <DataType T>
struct Buffer {
  using native = typename NativeOf<T>::type;
  native* data = nullptr;
  Span dims;
  native* typed_data() const { return data; }
  Span dimensions() const { return dims; }
  size_t element_count() const { size_t c = 1; for (size_t i = 0; i < dims.n; ++i) c *= (size_t)dims.p[i]; return c; }
};

template <typename T>
struct Result {
  T value;
  T* operator->() { return &value; }
  const T* operator->() const { return &value; }
  T& operator*() { return value; }
};
template <DataType T> using ResultBuffer = Result<Buffer<T>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename T> struct PlatformStream {};

template <typename... Ts>
struct Binding {
  template <typename C> struct CtxOf;
  template <typename S> struct CtxOf<PlatformStream<S>> { using type = S; };
  template <typename C> Binding<Ts..., typename CtxOf<C>::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename F> static constexpr bool matches = std::is_invocable_r<Error, F, Ts...>::value;
};
struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an `extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*)`; the stand-in defines a symbol of that
// name too and refuses to compile unless the implementation matches the binding
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, impl, binding)                                                        \
  static_assert(decltype(binding)::template matches<decltype(&impl)>,                                               \
                #impl " is not invocable with the context / attribute / argument / result types of its binding");  \
  extern "C" void* symbol(void*) { return nullptr; }
