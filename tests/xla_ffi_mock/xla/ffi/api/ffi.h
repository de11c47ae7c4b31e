// STRUCTURAL MOCK of xla/ffi/api/ffi.h -- TEST INFRASTRUCTURE ONLY (tests/test_host.py compiles
// dis_project_b200/csrc/lfm_xla_ffi.cc against it with -fsyntax-only).  The real header ships with jaxlib and is absent
// from the build image.  The mock models just enough of the typed-FFI binding DSL to check what can be checked without
// XLA: that every handler is invocable with exactly the context / argument / attribute / result types its binding
// declares, in that order (a static_assert inside XLA_FFI_DEFINE_HANDLER_SYMBOL).  It executes nothing.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla {
namespace ffi {

enum DataType { U8, S32, S64, F32, F64 };
template <DataType> struct NativeOf;
template <> struct NativeOf<U8> { using type = uint8_t; };
template <> struct NativeOf<S32> { using type = int32_t; };
template <> struct NativeOf<S64> { using type = int64_t; };
template <> struct NativeOf<F32> { using type = float; };
template <> struct NativeOf<F64> { using type = double; };

template <typename T> struct Span {
  const T* p = nullptr; size_t n = 0;
  size_t size() const { return n; }
  const T& operator[](size_t i) const { return p[i]; }
};

template <DataType dtype> class Buffer {
 public:
  using T = typename NativeOf<dtype>::type;
  T* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
 private:
  T* data_ = nullptr; Span<int64_t> dims_; size_t count_ = 0;
};

template <typename T> class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return Error(); }
};

// Ctx<PlatformStream<T>> contributes a T (the stream) to the handler's parameter list
template <typename T> struct PlatformStream { using context_type = T; };

template <typename... Ts> struct Binding {
  template <typename C> auto Ctx() const { return Binding<Ts..., typename C::context_type>(); }
  template <typename A> auto Arg() const { return Binding<Ts..., A>(); }
  template <typename R> auto Ret() const { return Binding<Ts..., Result<R>>(); }
  template <typename A> auto Attr(const char*) const { return Binding<Ts..., A>(); }
  template <typename Fn> static constexpr bool Accepts() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};

struct Ffi {
  static Binding<> Bind() { return Binding<>(); }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                    \
  static_assert(decltype(binding)::template Accepts<decltype(&impl)>(),                                         \
                #impl " is not invocable with the context / argument / attribute / result types of its binding"); \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
