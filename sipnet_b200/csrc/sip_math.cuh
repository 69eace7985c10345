// sip_math.cuh -- numeric leaves of the integrator.
//
// Two arithmetic builds of the same kernels (selected by -DSIP_FAST_MATH and the
// matching -fmad switch in the Makefile):
//   validation (default): -fmad=false, IEEE division; every expression keeps the
//     reference's shape so the only difference from the reference binary
//     (gcc -O0, x86-64 SSE2, glibc libm) is the last-ulp behaviour of pow/exp.
//   fast: -fmad=true, same expressions (contraction allowed) with
//     transcendental calls restructured (see sip_pow_q10 etc.).
#pragma once
#include <cmath>

#include "sip_libm.cuh"

namespace sip {

__device__ __forceinline__ double clip01(double x) {  // unitClip, reference common/util.h:38
  return fmin(fmax(x, 0.0), 1.0);
}

__device__ __forceinline__ double safe_ratio(double num, double den) {  // calcRatio, common/util.c:72-75
  const double d = den < 0.000001 ? 0.000001 : den;
  return num / d;
}

// exp / pow: glibc-2.39-exact restatements (sip_libm.cuh) in BOTH builds -- they are
// bit-identical to the reference's libm and ~3x fewer FP64 instructions than CUDA's pow().
__device__ __forceinline__ double sip_exp(double x) { return libm::exp(x); }
__device__ __forceinline__ double sip_pow(double x, double y) { return libm::pow(x, y); }
// pow(x, y) with log_inline(x) already known as (lhi, llo)
__device__ __forceinline__ double sip_pow_cached(double x, double lhi, double llo, double y) {
  return libm::pow_cached(x, libm::LogHL{lhi, llo}, y);
}

}  // namespace sip
