// MOCK of the XLA FFI C++ API (xla/ffi/api/ffi.h), just enough surface for bindings/pdeopt_jax_ffi.cc to be
// type-checked with g++ -fsyntax-only in an image without jaxlib: buffers, results, spans, errors and a binder whose
// chained calls are accepted but not checked against the handler signature.  TEST INFRASTRUCTURE ONLY — the real
// header ships with jaxlib (jax.ffi.include_dir()).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

namespace xla {
namespace ffi {

enum DataType { F32, F64, U8 };
template <DataType>
struct NativeOf;
template <>
struct NativeOf<F32> { using type = float; };
template <>
struct NativeOf<F64> { using type = double; };
template <>
struct NativeOf<U8> { using type = uint8_t; };

template <typename T>
struct Span {
  const T* b = nullptr;
  size_t n = 0;
  const T* begin() const { return b; }
  const T* end() const { return b + n; }
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return b[i]; }
};

template <DataType D>
struct Buffer {
  using T = typename NativeOf<D>::type;
  T* p = nullptr;
  Span<const int64_t> dims;
  T* typed_data() const { return p; }
  Span<const int64_t> dimensions() const { return dims; }
};

template <typename B>
struct Result {
  B b;
  B* operator->() { return &b; }
  const B* operator->() const { return &b; }
};

enum class ErrorCode { kInvalidArgument, kUnimplemented, kInternal };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

template <typename T>
struct PlatformStream {};

struct Binding {
  template <typename T> Binding& Ctx() { return *this; }
  template <typename T> Binding& Arg() { return *this; }
  template <typename T> Binding& Ret() { return *this; }
  template <typename T> Binding& Attr(const char*) { return *this; }
};
struct Ffi {
  static Binding Bind() { return Binding(); }
};

}  // namespace ffi
}  // namespace xla

// the real macro defines an extern "C" XLA_FFI_Error* symbol(XLA_FFI_CallFrame*); here: reference the handler so that it
// is instantiated and checked, and evaluate the binding expression
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(sym, impl, binding) \
  extern "C" void* sym() {                                  \
    (void)(binding);                                        \
    return reinterpret_cast<void*>(&impl);                  \
  }
