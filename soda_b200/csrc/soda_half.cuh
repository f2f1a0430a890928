// The DSL's `half` element type (IEEE binary16) as a value type for the stencil
// templates and the generated functors.
//
// Every operation rounds once, to nearest-even, exactly like a C++ `_Float16`
// evaluated without excess precision:
//   * + - * are the sm_100a half instructions with an explicit `.rn`
//     (__hadd_rn & co.; the un-suffixed forms may be contracted into HFMA2);
//   * / and the conversions from wider types go through fp32 / fp64
//     instructions that round correctly, and binary32 has more than 2*11+2
//     significand bits, so rounding the fp32 quotient to half is the correctly
//     rounded half quotient.
// Mixed expressions follow the C++ usual arithmetic conversions with `half`
// ranked below float and above every integer type (half + int -> half,
// half + float -> float, half + double -> double), which is what
// soda_b200.ir.common_type computes; the conversions themselves are explicit
// so that no expression compiles by accident through an ambiguous route.
#pragma once

#include <type_traits>

#ifndef SODA_EMU
#include <cuda_fp16.h>
#endif

namespace soda {

struct half_t {
#ifdef SODA_EMU
  using Raw = _Float16;
#else
  using Raw = __half;
#endif
  Raw v;

  half_t() = default;

#ifdef SODA_EMU
  template <typename U, typename std::enable_if<
                            std::is_arithmetic<U>::value, int>::type = 0>
  explicit half_t(U x) : v(static_cast<Raw>(x)) {}
  template <typename U, typename std::enable_if<
                            std::is_arithmetic<U>::value, int>::type = 0>
  explicit operator U() const { return static_cast<U>(v); }
#else
  __device__ __forceinline__ explicit half_t(float x) : v(__float2half_rn(x)) {}
  __device__ __forceinline__ explicit half_t(double x) : v(__double2half(x)) {}
  __device__ __forceinline__ explicit half_t(long long x) : v(__ll2half_rn(x)) {}
  __device__ __forceinline__ explicit half_t(unsigned long long x)
      : v(__ull2half_rn(x)) {}
  __device__ __forceinline__ explicit half_t(long x)
      : v(__ll2half_rn(static_cast<long long>(x))) {}
  __device__ __forceinline__ explicit half_t(unsigned long x)
      : v(__ull2half_rn(static_cast<unsigned long long>(x))) {}
  __device__ __forceinline__ explicit half_t(unsigned x) : v(__uint2half_rn(x)) {}
  // bool, char, short and int
  template <typename U,
            typename std::enable_if<std::is_integral<U>::value &&
                                        (sizeof(U) < 4 ||
                                         std::is_same<U, int>::value),
                                    int>::type = 0>
  __device__ __forceinline__ explicit half_t(U x)
      : v(__int2half_rn(static_cast<int>(x))) {}

  __device__ __forceinline__ explicit operator float() const {
    return __half2float(v);
  }
  __device__ __forceinline__ explicit operator double() const {
    return static_cast<double>(__half2float(v));
  }
  __device__ __forceinline__ explicit operator bool() const {
    return __half2float(v) != 0.0f;
  }
  // truncation toward zero, like a C++ floating -> integer conversion
  template <typename U,
            typename std::enable_if<std::is_integral<U>::value &&
                                        !std::is_same<U, bool>::value,
                                    int>::type = 0>
  __device__ __forceinline__ explicit operator U() const {
    if constexpr (sizeof(U) == 8) {
      if constexpr (std::is_signed<U>::value) {
        return static_cast<U>(__half2ll_rz(v));
      } else {
        return static_cast<U>(__half2ull_rz(v));
      }
    } else if constexpr (sizeof(U) == 4 && !std::is_signed<U>::value) {
      return static_cast<U>(__half2uint_rz(v));
    } else {
      return static_cast<U>(__half2int_rz(v));
    }
  }
#endif
};

static_assert(sizeof(half_t) == 2 && alignof(half_t) == 2, "binary16");
static_assert(std::is_trivially_copyable<half_t>::value, "moved as bytes");

namespace half_detail {
__device__ __forceinline__ half_t wrap(half_t::Raw raw) {
  half_t r;
  r.v = raw;
  return r;
}
}  // namespace half_detail

#ifdef SODA_EMU
#define SODA_HALF_BINARY(op, cuda_expr)                               \
  __device__ __forceinline__ half_t operator op(half_t a, half_t b) { \
    return half_detail::wrap(static_cast<_Float16>(                   \
        static_cast<float>(a.v) op static_cast<float>(b.v)));         \
  }
#else
#define SODA_HALF_BINARY(op, cuda_expr)                               \
  __device__ __forceinline__ half_t operator op(half_t a, half_t b) { \
    return half_detail::wrap(cuda_expr);                              \
  }
#endif
SODA_HALF_BINARY(+, __hadd_rn(a.v, b.v))
SODA_HALF_BINARY(-, __hsub_rn(a.v, b.v))
SODA_HALF_BINARY(*, __hmul_rn(a.v, b.v))
SODA_HALF_BINARY(/, __float2half_rn(__fdiv_rn(__half2float(a.v), __half2float(b.v))))
#undef SODA_HALF_BINARY

__device__ __forceinline__ half_t operator+(half_t a) { return a; }
__device__ __forceinline__ half_t operator-(half_t a) {
#ifdef SODA_EMU
  return half_detail::wrap(-a.v);
#else
  return half_detail::wrap(__hneg(a.v));
#endif
}
__device__ __forceinline__ bool operator!(half_t a) {
  return !static_cast<bool>(a);
}

// comparisons are exact in fp32
#define SODA_HALF_COMPARE(op)                                       \
  __device__ __forceinline__ bool operator op(half_t a, half_t b) { \
    return static_cast<float>(a) op static_cast<float>(b);          \
  }
SODA_HALF_COMPARE(<)
SODA_HALF_COMPARE(>)
SODA_HALF_COMPARE(<=)
SODA_HALF_COMPARE(>=)
SODA_HALF_COMPARE(==)
SODA_HALF_COMPARE(!=)
#undef SODA_HALF_COMPARE

// half (op) arithmetic type: the wider floating type wins, integers convert to
// half
#define SODA_HALF_MIXED(op)                                                    \
  template <typename U, typename std::enable_if<                               \
                            std::is_arithmetic<U>::value, int>::type = 0>      \
  __device__ __forceinline__ auto operator op(half_t a, U b) {                 \
    if constexpr (std::is_floating_point<U>::value) {                          \
      return static_cast<U>(a) op b;                                           \
    } else {                                                                   \
      return a op half_t(b);                                                   \
    }                                                                          \
  }                                                                            \
  template <typename U, typename std::enable_if<                               \
                            std::is_arithmetic<U>::value, int>::type = 0>      \
  __device__ __forceinline__ auto operator op(U a, half_t b) {                 \
    if constexpr (std::is_floating_point<U>::value) {                          \
      return a op static_cast<U>(b);                                           \
    } else {                                                                   \
      return half_t(a) op b;                                                   \
    }                                                                          \
  }
SODA_HALF_MIXED(+)
SODA_HALF_MIXED(-)
SODA_HALF_MIXED(*)
SODA_HALF_MIXED(/)
SODA_HALF_MIXED(<)
SODA_HALF_MIXED(>)
SODA_HALF_MIXED(<=)
SODA_HALF_MIXED(>=)
SODA_HALF_MIXED(==)
SODA_HALF_MIXED(!=)
#undef SODA_HALF_MIXED

}  // namespace soda

// ---- packed binary16 pairs (HADD2 / HMUL2) -------------------------------------
// Two half cells in one 32-bit register, the binary16 counterpart of the fp32
// pairs (F2): programs that only add, subtract and multiply half values are
// evaluated two cells per instruction.  Each half of add / sub / mul rounds to
// nearest-even exactly like the scalar operation (explicit `.rn`, never
// contracted), so results do not change.
namespace soda {

struct H2 {
#ifdef SODA_EMU
  _Float16 lo, hi;
#else
  __half2 v;
#endif
};
static_assert(sizeof(H2) == 4, "two binary16 cells");

#ifdef SODA_EMU
__device__ __forceinline__ H2 h2_pack(half_t lo, half_t hi) { return H2{lo.v, hi.v}; }
__device__ __forceinline__ half_t h2_lo(H2 p) { return half_detail::wrap(p.lo); }
__device__ __forceinline__ half_t h2_hi(H2 p) { return half_detail::wrap(p.hi); }
#define SODA_H2_BINARY(name, op, intrinsic)                                   \
  __device__ __forceinline__ H2 name(H2 a, H2 b) {                            \
    return H2{static_cast<_Float16>(static_cast<float>(a.lo)                  \
                                        op static_cast<float>(b.lo)),         \
              static_cast<_Float16>(static_cast<float>(a.hi)                  \
                                        op static_cast<float>(b.hi))};        \
  }
__device__ __forceinline__ H2 h2_neg(H2 a) { return H2{-a.lo, -a.hi}; }
#else
__device__ __forceinline__ H2 h2_pack(half_t lo, half_t hi) {
  H2 r;
  r.v = __halves2half2(lo.v, hi.v);
  return r;
}
__device__ __forceinline__ half_t h2_lo(H2 p) {
  return half_detail::wrap(__low2half(p.v));
}
__device__ __forceinline__ half_t h2_hi(H2 p) {
  return half_detail::wrap(__high2half(p.v));
}
#define SODA_H2_BINARY(name, op, intrinsic)            \
  __device__ __forceinline__ H2 name(H2 a, H2 b) {     \
    H2 r;                                              \
    r.v = intrinsic(a.v, b.v);                         \
    return r;                                          \
  }
__device__ __forceinline__ H2 h2_neg(H2 a) {
  H2 r;
  r.v = __hneg2(a.v);
  return r;
}
#endif
SODA_H2_BINARY(h2_add, +, __hadd2_rn)
SODA_H2_BINARY(h2_sub, -, __hsub2_rn)
SODA_H2_BINARY(h2_mul, *, __hmul2_rn)
#undef SODA_H2_BINARY

__device__ __forceinline__ H2 h2_splat(half_t v) { return h2_pack(v, v); }
__device__ __forceinline__ H2 operator+(H2 a, H2 b) { return h2_add(a, b); }
__device__ __forceinline__ H2 operator-(H2 a, H2 b) { return h2_sub(a, b); }
__device__ __forceinline__ H2 operator*(H2 a, H2 b) { return h2_mul(a, b); }
__device__ __forceinline__ H2 operator-(H2 a) { return h2_neg(a); }
__device__ __forceinline__ H2 operator+(H2 a) { return a; }
// scalar operands are half values or integer literals (which convert to half in
// a half expression): broadcast to both cells.  Packed programs have no float
// or double operands (soda_b200/codegen/cuda/plan.py, packable).
template <typename S>
__device__ __forceinline__ H2 operator+(H2 a, S b) { return h2_add(a, h2_splat(half_t(b))); }
template <typename S>
__device__ __forceinline__ H2 operator+(S a, H2 b) { return h2_add(h2_splat(half_t(a)), b); }
template <typename S>
__device__ __forceinline__ H2 operator-(H2 a, S b) { return h2_sub(a, h2_splat(half_t(b))); }
template <typename S>
__device__ __forceinline__ H2 operator-(S a, H2 b) { return h2_sub(h2_splat(half_t(a)), b); }
template <typename S>
__device__ __forceinline__ H2 operator*(H2 a, S b) { return h2_mul(a, h2_splat(half_t(b))); }
template <typename S>
__device__ __forceinline__ H2 operator*(S a, H2 b) { return h2_mul(h2_splat(half_t(a)), b); }

}  // namespace soda
