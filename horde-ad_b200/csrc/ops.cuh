// ops.cuh -- scalar semantics of the elementwise op codes (Ast.hs:707-726,
// AstInterpret.hs:396-435; instances at CarriersConcrete.hs:333-343,40-92).
#pragma once
#include <type_traits>

#include "common.cuh"

#ifdef __CUDACC__

template <typename T>
struct hb_is_fp : std::integral_constant<bool, std::is_same<T, double>::value || std::is_same<T, float>::value> {};

// ------------------------------------------------------------------ unary
template <int OP, typename T>
__device__ __forceinline__ T hb_un(T x) {
  if constexpr (OP == HB_NEGATE) return (T)(-x);
  else if constexpr (OP == HB_ABS) {
    if constexpr (hb_is_fp<T>::value) return fabs(x);
    else return (T)(x < 0 ? -x : x);
  } else if constexpr (OP == HB_SIGNUM) {
    // Haskell: signum x | x > 0 = 1 | x < 0 = -1 | otherwise = x
    return x > (T)0 ? (T)1 : (x < (T)0 ? (T)-1 : x);
  } else if constexpr (hb_is_fp<T>::value) {
    if constexpr (OP == HB_RECIP) return (T)1 / x;
    else if constexpr (OP == HB_EXP) return exp(x);
    else if constexpr (OP == HB_LOG) return log(x);
    else if constexpr (OP == HB_SQRT) return sqrt(x);
    else if constexpr (OP == HB_SIN) return sin(x);
    else if constexpr (OP == HB_COS) return cos(x);
    else if constexpr (OP == HB_TAN) return tan(x);
    else if constexpr (OP == HB_ASIN) return asin(x);
    else if constexpr (OP == HB_ACOS) return acos(x);
    else if constexpr (OP == HB_ATAN) return atan(x);
    else if constexpr (OP == HB_SINH) return sinh(x);
    else if constexpr (OP == HB_COSH) return cosh(x);
    else if constexpr (OP == HB_TANH) return tanh(x);
    else if constexpr (OP == HB_ASINH) return asinh(x);
    else if constexpr (OP == HB_ACOSH) return acosh(x);
    else if constexpr (OP == HB_ATANH) return atanh(x);
    else return x;
  } else {
    return x;
  }
}

template <typename T>
__device__ __forceinline__ T hb_un_rt(int op, T x) {
  switch (op) {
    case HB_NEGATE: return hb_un<HB_NEGATE>(x);
    case HB_ABS: return hb_un<HB_ABS>(x);
    case HB_SIGNUM: return hb_un<HB_SIGNUM>(x);
    case HB_RECIP: return hb_un<HB_RECIP>(x);
    case HB_EXP: return hb_un<HB_EXP>(x);
    case HB_LOG: return hb_un<HB_LOG>(x);
    case HB_SQRT: return hb_un<HB_SQRT>(x);
    case HB_SIN: return hb_un<HB_SIN>(x);
    case HB_COS: return hb_un<HB_COS>(x);
    case HB_TAN: return hb_un<HB_TAN>(x);
    case HB_ASIN: return hb_un<HB_ASIN>(x);
    case HB_ACOS: return hb_un<HB_ACOS>(x);
    case HB_ATAN: return hb_un<HB_ATAN>(x);
    case HB_SINH: return hb_un<HB_SINH>(x);
    case HB_COSH: return hb_un<HB_COSH>(x);
    case HB_TANH: return hb_un<HB_TANH>(x);
    case HB_ASINH: return hb_un<HB_ASINH>(x);
    case HB_ACOSH: return hb_un<HB_ACOSH>(x);
    case HB_ATANH: return hb_un<HB_ATANH>(x);
    default: return x;
  }
}

// ----------------------------------------------------------------- binary
template <int OP, typename T>
__device__ __forceinline__ T hb_bin(T a, T b) {
  if constexpr (OP == HB_ADD) return (T)(a + b);
  else if constexpr (OP == HB_SUB) return (T)(a - b);
  else if constexpr (OP == HB_MUL) return (T)(a * b);
  else if constexpr (OP == HB_MAX) return a < b ? b : a;
  else if constexpr (OP == HB_MIN) return b < a ? b : a;
  else if constexpr (hb_is_fp<T>::value) {
    if constexpr (OP == HB_DIVIDE) return a / b;
    else if constexpr (OP == HB_POWER) return pow(a, b);
    else if constexpr (OP == HB_LOGBASE) return log(b) / log(a);  // logBase x y = log y / log x
    else if constexpr (OP == HB_ATAN2) return atan2(a, b);
    else return a;
  } else {
    // array quotH/remH: zero divisors are replaced by 1 first
    // (CarriersConcrete.hs:40-56,140-144); C++ / and % truncate like quot/rem.
    if constexpr (OP == HB_QUOT) { T d = b == 0 ? (T)1 : b; return (T)(a / d); }
    else if constexpr (OP == HB_REM) { T d = b == 0 ? (T)1 : b; return (T)(a % d); }
    else return a;
  }
}

#endif
