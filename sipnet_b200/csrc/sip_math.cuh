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

namespace sip {

__device__ __forceinline__ double clip01(double x) {  // unitClip, reference common/util.h:38
  return fmin(fmax(x, 0.0), 1.0);
}

__device__ __forceinline__ double safe_ratio(double num, double den) {  // calcRatio, common/util.c:72-75
  const double d = den < 0.000001 ? 0.000001 : den;
  return num / d;
}

__device__ __forceinline__ double sip_exp(double x) { return exp(x); }

__device__ __forceinline__ double sip_pow(double x, double y) { return pow(x, y); }

// pow(2, y) of calcLightEff (sipnet.c:551)
__device__ __forceinline__ double sip_pow2(double y) {
#ifdef SIP_FAST_MATH
  return exp2(y);
#else
  return pow(2.0, y);
#endif
}

}  // namespace sip
