// exactmath.cuh -- the reference's f64 arithmetic, op for op, with explicit round-to-nearest
// intrinsics so nvcc can never contract a*b+c into an FMA (Rust/LLVM does not fuse either).
//   cosine  : fnc/util/math/vector.rs:65-71, 279-281, 301-314
//   euclid  : fnc/util/math/vector.rs:288-299
//   manhattan / chebyshev / hamming / pearson: cited at each step function below
// NaN sign convention = x86-64 hardware (where the reference runs): a GENERATED NaN (0/0, inf-inf,
// inf*0) is the negative "real indefinite" 0xFFF8000000000000 and therefore sorts FIRST under
// Number::cmp's total_cmp; a NaN that came in through the data (Rust f64::NAN, positive) propagates
// as a positive NaN and sorts LAST.  Mixed cases are unpinned.
#pragma once
#include <cstdint>

namespace sdb {

struct ExactAcc {
  double acc = 0.0;
  double acc2 = 0.0;  // pearson: sum (x-m1)^2
  bool nan_in = false;
  __device__ __forceinline__ void cosine_step(double x, double q) {
    nan_in |= (x != x);
    acc = __dadd_rn(acc, __dmul_rn(x, q));
  }
  __device__ __forceinline__ void euclid_step(double x, double q) {
    nan_in |= (x != x);
    const double d = __dsub_rn(x, q);
    acc = __dadd_rn(acc, __dmul_rn(d, d));
  }
  //   manhattan: vector.rs:152-157   acc = acc + |x - q|        (Number add, Int(0) start == 0.0 + ...)
  __device__ __forceinline__ void manhattan_step(double x, double q) {
    nan_in |= (x != x);
    acc = __dadd_rn(acc, fabs(__dsub_rn(x, q)));
  }
  //   chebyshev: vector.rs:215-225   fold(f64::MIN, f64::max)  -- f64::max returns the non-NaN operand, as fmax does
  __device__ __forceinline__ void chebyshev_step(double x, double q) { acc = fmax(acc, fabs(__dsub_rn(x, q))); }
  //   hamming  : vector.rs:111-116   count of a != b under Number's PartialEq (0.0 == -0.0, NaN == NaN bitwise)
  __device__ __forceinline__ void hamming_step(double x, double q) {
    const bool eq = (__double_as_longlong(x) == __double_as_longlong(q)) || (x == 0.0 && q == 0.0);
    acc = __dadd_rn(acc, eq ? 0.0 : 1.0);  // exact: an integer count below 2^53
  }
  //   minkowski: vector.rs:163-174   acc = acc + |x - q|^p ; finish: acc^(1/p).  pow() is CUDA's libm here and the
  //              platform libm in the reference: each call agrees to within an ulp or two, not bit for bit.
  __device__ __forceinline__ void minkowski_step(double x, double q, double p) {
    nan_in |= (x != x);
    acc = __dadd_rn(acc, pow(fabs(__dsub_rn(x, q)), p));
  }
  //   pearson  : vector.rs:133-146   pass A: sum x ; pass B: covar += (x-m1)*(q-m2), dev += (x-m1)^2
  __device__ __forceinline__ void sum_step(double x) {
    nan_in |= (x != x);
    acc = __dadd_rn(acc, x);
  }
  __device__ __forceinline__ void pearson_step(double x, double q, double m1, double m2) {
    const double dx = __dsub_rn(x, m1);
    acc = __dadd_rn(acc, __dmul_rn(dx, __dsub_rn(q, m2)));
    acc2 = __dadd_rn(acc2, __dmul_rn(dx, dx));
  }
};

__device__ __forceinline__ double canon_nan(double r, bool nan_in) {
  if (r != r) return __longlong_as_double(nan_in ? 0x7FF8000000000000ll : (long long)0xFFF8000000000000ull);
  return r;
}
__device__ __forceinline__ double cosine_finish(const ExactAcc& a, double row_mag, double q_mag, bool q_nan) {
  const double r = __dsub_rn(1.0, __ddiv_rn(a.acc, __dmul_rn(row_mag, q_mag)));
  return canon_nan(r, a.nan_in || q_nan);
}
__device__ __forceinline__ double euclid_finish(const ExactAcc& a, bool q_nan) {
  return canon_nan(__dsqrt_rn(a.acc), a.nan_in || q_nan);
}

}  // namespace sdb
