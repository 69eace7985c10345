// sip_math.cuh -- numeric leaves of the integrator.
//
// Every kernel on the model path is compiled with -fmad=false and keeps the reference's expression shapes; exp and
// pow are the glibc-exact restatements of sip_libm.cuh.  (The throughput policy's translation units, sip_run_thr_*.cu,
// are the one exception: -fmad=true, see sip_num.cuh ThroughNum.)
#pragma once
#include <cmath>

#include "sip_libm.cuh"

namespace sip {

// fmax(x, 0.0) and fmin(x, 1.0) against a CONSTANT: one compare and one select give exactly what CUDA's fmax / fmin
// return for every input (x = NaN -> the constant; x = -0 -> +0 for max0, kept for min1), where the two-variable
// library forms cost eight instructions each on sm_100a (no DMNMX).
// (written on the two 32-bit halves: nvcc turns `x > 0.0 ? x : 0.0` back into its nine-instruction fmax sequence)
__device__ __forceinline__ double max0(double x) {
  const bool keep = x > 0.0;
  return __hiloint2double(keep ? __double2hiint(x) : 0, keep ? __double2loint(x) : 0);
}
__device__ __forceinline__ double min1(double x) {
  const bool keep = x < 1.0;
  return __hiloint2double(keep ? __double2hiint(x) : 0x3ff00000, keep ? __double2loint(x) : 0);
}

__device__ __forceinline__ double clip01(double x) {  // unitClip, reference common/util.h:38: fmin(fmax(x, 0.0), 1.0)
  return min1(max0(x));
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
