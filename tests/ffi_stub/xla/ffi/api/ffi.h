// Stand-in for XLA's `xla/ffi/api/ffi.h` (jax.ffi.include_dir()), which does not exist in this image: ONLY the API
// surface ffi/pst_xla_ffi.cc uses, with the real header's names and call shapes, so that tests/test_abi.py can
// compile-check the handlers (types, argument order against the binder, the C ABI calls).  It is test
// infrastructure: nothing links against it and it proves nothing about XLA's runtime behaviour.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { F32, S32, U8 };
template <DataType> struct NativeType;
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<U8> { using type = uint8_t; };

struct Span {
  const int64_t* p; size_t n;
  int64_t operator[](size_t i) const { return p[i]; }
  size_t size() const { return n; }
};

template <DataType T>
class Buffer {
 public:
  using native = typename NativeType<T>::type;
  native* typed_data() const { return data_; }
  Span dimensions() const { return Span{dims_, rank_}; }
  size_t element_count() const { size_t n = 1; for (size_t i = 0; i < rank_; ++i) n *= (size_t)dims_[i]; return n; }
  size_t size_bytes() const { return element_count() * sizeof(native); }
 private:
  native* data_ = nullptr; const int64_t* dims_ = nullptr; size_t rank_ = 0;
};
template <DataType T>
class ResultBuffer {
 public:
  Buffer<T>* operator->() { return &b_; }
 private:
  Buffer<T> b_;
};

enum class ErrorCode { kOk, kInvalidArgument, kInternal };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
 private:
  ErrorCode code_ = ErrorCode::kOk; std::string msg_;
};

template <typename S> struct PlatformStream { using type = S; };
template <typename B> struct AsResult;
template <DataType T> struct AsResult<Buffer<T>> { using type = ResultBuffer<T>; };

// binder: records the handler's parameter types in declaration order; To() checks them against the function
template <typename... Ts>
struct Binder {
  template <typename C> constexpr Binder<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename A> constexpr Binder<Ts..., A> Attr(const char*) const { return {}; }
  template <typename A> constexpr Binder<Ts..., A> Arg() const { return {}; }
  template <typename B> constexpr Binder<Ts..., typename AsResult<B>::type> Ret() const { return {}; }  // Ret<Buffer<T>> -> ResultBuffer<T>
  // the real binder's To(): the handler must be callable with exactly the bound parameter list and return Error
  template <typename F> constexpr bool Check(F*) const {
    static_assert(std::is_invocable_r<Error, F*, Ts...>::value, "handler signature does not match the binding");
    return true;
  }
};
struct Ffi { static constexpr Binder<> Bind() { return {}; } };

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binder) \
  extern "C" void* name() { static_assert((binder).Check(&impl), "binding"); return reinterpret_cast<void*>(&impl); }
