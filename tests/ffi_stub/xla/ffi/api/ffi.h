// Minimal stand-in for XLA's header-only FFI API ("xla/ffi/api/ffi.h", shipped with jaxlib under
// jax.ffi.include_dir()) -- TEST INFRASTRUCTURE ONLY.  JAX is not installable in this image, so the real header is
// absent; this stub declares exactly the surface integration/jc_xla_ffi.cc uses (names and signatures as in the public
// XLA FFI documentation) so that CI compiles and links the handler file against libjc_b200.so and drives its *Impl
// functions through fake buffers (tests/test_ffi_shim.py).  It implements no binding machinery: the
// XLA_FFI_DEFINE_HANDLER_SYMBOL symbols it produces only type-check the Bind() chain against the Impl signature.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

struct XLA_FFI_Error;
struct XLA_FFI_CallFrame;

namespace xla::ffi {

enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kUnimplemented = 12, kInternal = 13 };

class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  static Error Internal(std::string m) { return Error(ErrorCode::kInternal, std::move(m)); }
  bool success() const { return code_ == ErrorCode::kOk; }
  ErrorCode code() const { return code_; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

enum DataType { U8 = 0, F64 = 1 };
template <DataType> struct NativeType;
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<F64> { using type = double; };

template <typename T>
class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  const T& operator[](size_t i) const { return p_[i]; }
  size_t size() const { return n_; }
  const T* begin() const { return p_; }
  const T* end() const { return p_ + n_; }

 private:
  const T* p_;
  size_t n_;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeType<dtype>::type;
  Buffer(T* data, std::vector<int64_t> dims) : data_(data), dims_(std::move(dims)) {}
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
  T* typed_data() const { return data_; }
  void* untyped_data() const { return data_; }
  size_t element_count() const {
    size_t n = 1;
    for (int64_t d : dims_) n *= static_cast<size_t>(d);
    return n;
  }
  size_t size_bytes() const { return element_count() * sizeof(T); }

 private:
  T* data_;
  std::vector<int64_t> dims_;
};

template <typename B>
class Result {
 public:
  explicit Result(B b) : b_(std::move(b)) {}
  B* operator->() { return &b_; }
  B& operator*() { return b_; }

 private:
  B b_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

template <typename T> struct PlatformStream {};

// Bind() chain: accumulates the handler's parameter list as a type so that To<Impl>() can type-check it
template <typename... Ts> struct TypeList {};
template <typename T> struct CtxArg { using type = T; };
template <typename T> struct CtxArg<PlatformStream<T>> { using type = T; };

template <typename... Params>
class Binding {
 public:
  template <typename T> Binding<Params..., typename CtxArg<T>::type> Ctx() const { return {}; }
  template <typename T> Binding<Params..., T> Arg() const { return {}; }
  template <typename T> Binding<Params..., Result<T>> Ret() const { return {}; }
  template <typename T> Binding<Params..., T> Attr(std::string_view) const { return {}; }
  template <typename Fn>
  bool To(Fn fn) const {
    static_assert(std::is_invocable_r_v<Error, Fn, Params...>, "Bind() chain does not match the handler's signature");
    (void)fn;
    return true;
  }
};

struct Ffi {
  static Binding<> Bind() { return {}; }
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                       \
  static const bool name##_bound = (binding).To(impl);                           \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { return nullptr; }
