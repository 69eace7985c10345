// sip_libm.cuh -- exp() and pow() that reproduce glibc 2.39's results bit for bit.
//
// Why: the reference binary's only numeric dependency outside its own source
// tree is glibc's libm.  CUDA's exp/pow differ from glibc's in the
// last 1-2 ulp, and the model amplifies that to ~1e-10..1e-9 on ill-conditioned
// outputs.  Restating glibc's algorithm removes the difference entirely: with
// these two functions and -fmad=false the device performs the same IEEE-754
// operations as the reference's gcc -O0 x86-64 build.
//
// What is restated (no code is copied; glibc is not part of /root/reference):
//   glibc 2.39 sysdeps/ieee754/dbl-64/e_exp.c and e_pow.c -- the table-driven
//   algorithms of ARM Optimized Routines (S. Nagy):
//     exp(x)  = 2^(k/N) * exp(r),  N = 128, k = round(x N/ln2), degree-5 polynomial in r,
//               2^(i/N) from a 128-entry {tail, scale} table;
//     pow(x,y)= exp_inline(y * log_inline(x)) with log(x) = k ln2 + log(c) + log1p(z/c - 1),
//               128 sub-intervals, {invc, logc, logctail} table, degree-7 polynomial,
//               result carried as hi + lo and handed to exp with the tail.
//   The placement of fused multiply-adds follows the *_fma variants that glibc's
//   x86-64 ifunc selects on any FMA/AVX2 CPU (__exp_fma, __pow_fma): every
//   FMA/MUL/ADD below corresponds to one vfmadd/vmulsd/vaddsd of that code path
//   (read from the disassembly of libm.so.6; see DESIGN.md "libm parity").
//   Tables: sip_libm_tables.h (tools/gen_libm_tables.py extracts them from the
//   system libm; tests/test_libm_exact.py checks ~10^7 inputs against the live libm).
//
// The same header compiles for the host (plain C++, used only by the libm
// parity test) and for the device.
#pragma once
#include <cstdint>
#include <cstring>

#include "sip_libm_tables.h"

#if defined(__CUDACC__)
#define SIP_HD __host__ __device__ __forceinline__
#else
#define SIP_HD inline
#endif

namespace sip {
namespace libm {

#if defined(__CUDACC__)
static __device__ const uint64_t d_exp_tab[2 * 128] = SIP_EXP_TAB_INIT;
static __device__ const uint64_t d_powlog_tab[4 * 128] = SIP_POWLOG_TAB_INIT;
#endif
#if !defined(__CUDA_ARCH__)
static const uint64_t h_exp_tab[2 * 128] = SIP_EXP_TAB_INIT;
static const uint64_t h_powlog_tab[4 * 128] = SIP_POWLOG_TAB_INIT;
#endif

SIP_HD uint64_t exp_tab(unsigned i) {
#if defined(__CUDA_ARCH__)
  return __ldg(&d_exp_tab[i]);
#else
  return h_exp_tab[i];
#endif
}
SIP_HD uint64_t powlog_tab(unsigned i) {
#if defined(__CUDA_ARCH__)
  return __ldg(&d_powlog_tab[i]);
#else
  return h_powlog_tab[i];
#endif
}

SIP_HD uint64_t asu64(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}
SIP_HD double asf64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}
// single IEEE operations, immune to the compiler's contraction setting
SIP_HD double FMA(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}
SIP_HD double MUL(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;  // host build uses -ffp-contract=off
#endif
}
SIP_HD double ADD(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
SIP_HD double SUB(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}

constexpr uint64_t kInf = 0x7ff0000000000000ull;
constexpr uint64_t kOne = 0x3ff0000000000000ull;
constexpr uint64_t kSignBit = 0x8000000000000000ull;

SIP_HD double c_(uint64_t bits) { return asf64(bits); }

// __math_oflow / __math_uflow / __math_divzero / __math_invalid: value only (no errno, no flags)
SIP_HD double m_oflow(uint32_t sign) { return asf64((sign ? kSignBit : 0) | kInf); }
SIP_HD double m_uflow(uint32_t sign) { return asf64(sign ? kSignBit : 0); }
SIP_HD double m_nan() { return asf64(0x7ff8000000000000ull); }

// e_exp.c specialcase(): |x| in [512, 1024): the scale 2^(k/N) is outside the double range
SIP_HD double exp_specialcase(double tmp, uint64_t sbits, uint64_t ki) {
  if ((ki & 0x80000000ull) == 0) {
    // k > 0: the exponent of scale might have overflowed by <= 460
    sbits -= 1009ull << 52;
    const double scale = asf64(sbits);
    return MUL(c_(0x7f00000000000000ull) /*0x1p1009*/, FMA(scale, tmp, scale));
  }
  // k < 0: take care in the subnormal range
  sbits += 1022ull << 52;
  const double scale = asf64(sbits);
  const double st = MUL(scale, tmp);
  double y = ADD(scale, st);
  if (y < 1.0) {
    double lo = ADD(SUB(scale, y), st);
    const double hi = ADD(1.0, y);
    lo = ADD(ADD(SUB(1.0, hi), y), lo);
    y = SUB(ADD(lo, hi), 1.0);
    if (y == 0.0) y = 0.0;
  }
  return MUL(c_(0x0010000000000000ull) /*0x1p-1022*/, y);
}

// glibc 2.39 __exp (e_exp.c), FMA variant
SIP_HD double exp(double x) {
  const uint64_t ix = asu64(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {              // not in [2^-54, 512)
    if (abstop - 0x3c9u >= 0x80000000u) return ADD(1.0, x);  // |x| < 2^-54 (0 is a common input)
    if (abstop >= 0x409u) {                    // |x| >= 1024, inf, nan
      if (ix == (kSignBit | kInf)) return 0.0;
      if (abstop >= 0x7ffu) return ADD(1.0, x);
      return (ix >> 63) ? m_uflow(0) : m_oflow(0);
    }
    abstop = 0;  // 512 <= |x| < 1024: handled by exp_specialcase
  }
  // x = ln2/N * k + r, k integer, |r| <= ln2/2N
  const double kdb = FMA(x, c_(SIP_EXP_InvLn2N), c_(SIP_EXP_Shift));
  const uint64_t ki = asu64(kdb);
  const double kd = SUB(kdb, c_(SIP_EXP_Shift));
  double r = FMA(kd, c_(SIP_EXP_NegLn2hiN), x);
  r = FMA(kd, c_(SIP_EXP_NegLn2loN), r);
  const unsigned idx = 2u * (unsigned)(ki & 127u);
  const uint64_t top = ki << 45;
  const double tail = asf64(exp_tab(idx));
  const uint64_t sbits = exp_tab(idx + 1) + top;
  // exp(x) = 2^(k/N) exp(r) ~= scale + scale * (tail + exp(r) - 1)
  const double r2 = MUL(r, r);
  const double p1 = FMA(r, c_(SIP_EXP_C3), c_(SIP_EXP_C2));
  const double t = ADD(r, tail);
  const double p2 = FMA(r, c_(SIP_EXP_C5), c_(SIP_EXP_C4));
  double tmp = FMA(p1, r2, t);
  tmp = FMA(MUL(r2, r2), p2, tmp);
  if (abstop == 0) return exp_specialcase(tmp, sbits, ki);
  const double scale = asf64(sbits);
  return FMA(scale, tmp, scale);
}

// e_pow.c specialcase(): like exp's, but keeps the sign carried in sbits
SIP_HD double pow_specialcase(double tmp, uint64_t sbits, uint64_t ki) {
  if ((ki & 0x80000000ull) == 0) {
    sbits -= 1009ull << 52;
    const double scale = asf64(sbits);
    return MUL(c_(0x7f00000000000000ull), FMA(scale, tmp, scale));
  }
  sbits += 1022ull << 52;
  const double scale = asf64(sbits);
  const double st = MUL(scale, tmp);
  double y = ADD(scale, st);
  if ((y < 0 ? -y : y) < 1.0) {
    const double one = (y < 0.0) ? -1.0 : 1.0;
    double lo = ADD(SUB(scale, y), st);
    const double hi = ADD(y, one);
    lo = ADD(ADD(SUB(one, hi), y), lo);
    y = SUB(ADD(lo, hi), one);
    if (y == 0.0) y = asf64(sbits & kSignBit);
  }
  return MUL(c_(0x0010000000000000ull), y);
}

// e_pow.c exp_inline(): exp(x + xtail) * (-1)^(sign_bias != 0)
SIP_HD double pow_exp_inline(double x, double xtail, uint32_t sign_bias) {
  const uint64_t ix = asu64(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ff;
  if (abstop - 0x3c9u >= 0x3fu) {
    if (abstop - 0x3c9u >= 0x80000000u) {
      const double one = ADD(1.0, x);
      return sign_bias ? -one : one;
    }
    if (abstop >= 0x409u) return (ix >> 63) ? m_uflow(sign_bias) : m_oflow(sign_bias);
    abstop = 0;
  }
  const double kdb = FMA(x, c_(SIP_EXP_InvLn2N), c_(SIP_EXP_Shift));
  const uint64_t ki = asu64(kdb);
  const double kd = SUB(kdb, c_(SIP_EXP_Shift));
  double r = FMA(kd, c_(SIP_EXP_NegLn2hiN), x);
  r = FMA(kd, c_(SIP_EXP_NegLn2loN), r);
  r = ADD(xtail, r);
  const unsigned idx = 2u * (unsigned)(ki & 127u);
  const uint64_t top = (ki + sign_bias) << 45;
  const double tail = asf64(exp_tab(idx));
  const uint64_t sbits = exp_tab(idx + 1) + top;
  const double r2 = MUL(r, r);
  const double p1 = FMA(r, c_(SIP_EXP_C3), c_(SIP_EXP_C2));
  const double t = ADD(r, tail);
  const double p2 = FMA(r, c_(SIP_EXP_C5), c_(SIP_EXP_C4));
  double tmp = FMA(p1, r2, t);
  tmp = FMA(p2, MUL(r2, r2), tmp);
  if (abstop == 0) return pow_specialcase(tmp, sbits, ki);
  const double scale = asf64(sbits);
  return FMA(tmp, scale, scale);
}

// e_pow.c checkint(): 0 = not an integer, 1 = odd, 2 = even
SIP_HD int pow_checkint(uint64_t iy) {
  const int e = (int)(iy >> 52) & 0x7ff;
  if (e < 0x3ff) return 0;
  if (e > 0x3ff + 52) return 2;
  if (iy & ((1ull << (0x3ff + 52 - e)) - 1)) return 0;
  if (iy & (1ull << (0x3ff + 52 - e))) return 1;
  return 2;
}
SIP_HD bool pow_zeroinfnan(uint64_t i) { return 2 * i - 1 >= 2 * kInf - 1; }

struct LogHL {
  double hi, lo;
};

// e_pow.c log_inline(ix, &lo): x = 2^k z, z in [OFF, 2 OFF), 128 sub-intervals
SIP_HD LogHL pow_log_bits(uint64_t ix) {
  const uint64_t tmp = ix - 0x3fe6955500000000ull;
  const unsigned i = (unsigned)(tmp >> 45) & 127u;
  const int k = (int)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & (0xfffull << 52));
  const double z = asf64(iz);
  const double kd = (double)k;
  const double invc = asf64(powlog_tab(4 * i));
  const double logc = asf64(powlog_tab(4 * i + 2));
  const double logctail = asf64(powlog_tab(4 * i + 3));
  const double r = FMA(z, invc, -1.0);  // exact: 1/c has few bits and |z/c - 1| < 1/N
  // k ln2 + log(c) + r
  const double t1 = FMA(kd, c_(SIP_POWLOG_Ln2hi), logc);
  const double t2 = ADD(r, t1);
  const double lo1 = FMA(kd, c_(SIP_POWLOG_Ln2lo), logctail);
  const double lo2 = ADD(SUB(t1, t2), r);
  const double ar = MUL(r, c_(SIP_POWLOG_A0));  // A0 = -0.5
  const double ar2 = MUL(r, ar);
  const double ar3 = MUL(r, ar2);
  // k ln2 + log(c) + r + A0 r r
  const double hi = ADD(t2, ar2);
  const double lo3 = FMA(ar, r, -ar2);
  const double lo4 = ADD(SUB(t2, hi), ar2);
  // p = log1p(r) - r - A0 r r
  const double q56 = FMA(r, c_(SIP_POWLOG_A6), c_(SIP_POWLOG_A5));
  const double q34 = FMA(r, c_(SIP_POWLOG_A4), c_(SIP_POWLOG_A3));
  const double q12 = FMA(r, c_(SIP_POWLOG_A2), c_(SIP_POWLOG_A1));
  const double qa = FMA(q56, ar2, q34);
  const double q = FMA(ar2, qa, q12);
  const double s4 = ADD(ADD(ADD(lo1, lo2), lo3), lo4);
  const double lo = FMA(ar3, q, s4);
  LogHL out;
  out.hi = ADD(hi, lo);
  out.lo = ADD(SUB(hi, out.hi), lo);
  return out;
}

// y * log(x) as ehi + elo, then exp (the tail of e_pow.c pow())
SIP_HD double pow_exp_from_log(const LogHL lx, double y, uint32_t sign_bias) {
  const double ehi = MUL(y, lx.hi);
  const double elo = FMA(y, lx.lo, FMA(lx.hi, y, -ehi));
  return pow_exp_inline(ehi, elo, sign_bias);
}

// glibc 2.39 __pow (e_pow.c), FMA variant
SIP_HD double pow(double x, double y) {
  uint32_t sign_bias = 0;
  uint64_t ix = asu64(x);
  const uint64_t iy = asu64(y);
  uint32_t topx = (uint32_t)(ix >> 52);
  const uint32_t topy = (uint32_t)(iy >> 52);
  if (topx - 0x001u >= 0x7feu || (topy & 0x7ffu) - 0x3beu >= 0x80u) {
    // x < 0x1p-1022 or inf or nan or negative, or |y| < 0x1p-65 or |y| >= 0x1p63 or nan
    if (pow_zeroinfnan(iy)) {
      if (2 * iy == 0) return 1.0;
      if (ix == kOne) return 1.0;
      if (2 * ix > 2 * kInf || 2 * iy > 2 * kInf) return ADD(x, y);
      if (2 * ix == 2 * kOne) return 1.0;
      if ((2 * ix < 2 * kOne) == !(iy >> 63)) return 0.0;  // |x|<1 && y==inf or |x|>1 && y==-inf
      return MUL(y, y);
    }
    if (pow_zeroinfnan(ix)) {
      double x2 = MUL(x, x);
      if ((ix >> 63) && pow_checkint(iy) == 1) {
        x2 = -x2;
        sign_bias = 1;
      }
      if (2 * ix == 0 && (iy >> 63)) return m_oflow(sign_bias);  // __math_divzero
      return (iy >> 63) ? 1 / x2 : x2;
    }
    // here x and y are non-zero finite
    if (ix >> 63) {  // finite x < 0
      const int yint = pow_checkint(iy);
      if (yint == 0) return m_nan();  // __math_invalid
      if (yint == 1) sign_bias = 0x800u << SIP_EXP_TABLE_BITS;
      ix &= 0x7fffffffffffffffull;
      topx &= 0x7ff;
    }
    if ((topy & 0x7ffu) - 0x3beu >= 0x80u) {
      // sign_bias == 0 here because y is not odd
      if (ix == kOne) return 1.0;
      if ((topy & 0x7ffu) < 0x3beu) return ix > kOne ? ADD(1.0, y) : SUB(1.0, y);  // |y| < 2^-65
      return (ix > kOne) == (topy < 0x800u) ? m_oflow(0) : m_uflow(0);
    }
    if (topx == 0) {  // normalise subnormal x so the exponent becomes negative
      ix = asu64(MUL(x, c_(0x4330000000000000ull) /*0x1p52*/));
      ix &= 0x7fffffffffffffffull;
      ix -= 52ull << 52;
    }
  }
  const LogHL lx = pow_log_bits(ix);
  return pow_exp_from_log(lx, y, sign_bias);
}

// A base is "regular" when pow() takes none of its special branches for it:
// positive, finite, normal.  Its log_inline() result can then be computed once
// and reused for every exponent (same operations => same bits as pow(x, y)).
SIP_HD bool pow_base_regular(double x) { return (uint32_t)(asu64(x) >> 52) - 0x001u < 0x7feu; }
SIP_HD bool pow_exponent_regular(double y) { return ((uint32_t)(asu64(y) >> 52) & 0x7ffu) - 0x3beu < 0x80u; }

// log_inline of a regular base; {NaN, NaN} marks "not regular, use pow()"
SIP_HD LogHL pow_log(double x) {
  if (!pow_base_regular(x)) return LogHL{m_nan(), m_nan()};
  return pow_log_bits(asu64(x));
}

// pow(x, y) given lx = pow_log(x).  Bit-identical to pow(x, y).
SIP_HD double pow_cached(double x, const LogHL lx, double y) {
  if (lx.hi == lx.hi && pow_exponent_regular(y)) return pow_exp_from_log(lx, y, 0);
  return pow(x, y);
}

}  // namespace libm
}  // namespace sip
