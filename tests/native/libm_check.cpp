// libm_check.cpp -- host harness for tests/test_libm_exact.py.
// Compiles the PRODUCT header sipnet_b200/csrc/sip_libm.cuh in host mode and
// compares sip::libm::exp / sip::libm::pow with the live glibc libm bit for bit.
// Build: g++ -O2 -ffp-contract=off -mfma -I sipnet_b200/csrc tests/native/libm_check.cpp -lm
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sip_libm.cuh"

static uint64_t s_state = 0x9E3779B97F4A7C15ull;
static uint64_t next_u64() {
  uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double uni(double lo, double hi) { return lo + (hi - lo) * ((next_u64() >> 11) * (1.0 / 9007199254740992.0)); }
static double from_bits(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static uint64_t bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
static bool same(double a, double b) { return bits(a) == bits(b) || (a != a && b != b); }

static long bad_exp = 0, bad_pow = 0, n_exp = 0, n_pow = 0;
static void check_exp(double x) {
  ++n_exp;
  const double a = sip::libm::exp(x), b = ::exp(x);
  if (!same(a, b)) {
    if (bad_exp++ < 10) printf("EXP MISMATCH x=%a ours=%a libm=%a\n", x, a, b);
  }
}
static void check_pow(double x, double y) {
  ++n_pow;
  const double a = sip::libm::pow(x, y), b = ::pow(x, y);
  if (!same(a, b)) {
    if (bad_pow++ < 10) printf("POW MISMATCH x=%a y=%a ours=%a libm=%a\n", x, y, a, b);
  }
  const double c = sip::libm::pow_cached(x, sip::libm::pow_log(x), y);  // reused-log form
  if (!same(c, b)) {
    if (bad_pow++ < 10) printf("POW_CACHED MISMATCH x=%a y=%a ours=%a libm=%a\n", x, y, c, b);
  }
}

int main(int argc, char **argv) {
  const long n = argc > 1 ? atol(argv[1]) : 2000000;
  const double specials[] = {0.0, -0.0, 1.0, -1.0, 2.0, -2.0, 0.5, -0.5, 3.0, -3.0, 1e-300, -1e-300, 1e300, -1e300,
                             5e-324, -5e-324, 2.2250738585072014e-308, INFINITY, -INFINITY, NAN, 709.78, 709.79, -745.13,
                             -745.14, -708.4, 512.0, -512.0, 1023.9, -1023.9, 1024.0, -1024.0, 1e-17, -1e-17, 5.5e-17,
                             0x1p-54, 0x1p-55, 0x1p63, 0x1p-65, 0x1p-66, 0x1p62, 1.0000000000000002, 0.9999999999999999,
                             1075.0, -1075.0, 2047.0, 4503599627370497.0, 9007199254740993.0, 0.1, 10.0, 1e6, 1e-6};
  const int ns = (int)(sizeof specials / sizeof specials[0]);
  for (int i = 0; i < ns; ++i) {
    check_exp(specials[i]);
    for (int j = 0; j < ns; ++j) check_pow(specials[i], specials[j]);
  }
  for (long i = 0; i < n; ++i) {
    // model-shaped arguments
    check_exp(uni(-60.0, 20.0));                  // canopy attenuation, soil resistance, tillage decay
    check_exp(uni(-750.0, 710.0));
    check_exp(uni(-1100.0, 1100.0));
    check_exp(from_bits(next_u64()));
    check_pow(2.0, uni(-300.0, 10.0));            // light effect, sipnet.c:551
    check_pow(uni(1.0, 6.0), uni(-6.0, 6.0));     // Q10 terms
    check_pow(uni(1e-6, 6.0), uni(0.5, 4.0));     // vpd ^ dVpdExp
    check_pow(uni(0.0, 1.0), uni(0.0, 4.0));      // moisture / anaerobic index powers
    check_pow(uni(0.0, 50.0), 2.0);               // (psnTMax - psnTMin)/2 squared
    check_pow(uni(1e-300, 1e300), uni(-3.0, 3.0));
    check_pow(uni(0.0, 2.0), uni(-1100.0, 1100.0));  // overflow / underflow / subnormal results
    check_pow(-uni(0.0, 10.0), (double)(long)uni(-40.0, 40.0));  // negative base, integer exponent
    check_pow(-uni(0.0, 10.0), uni(-4.0, 4.0));                  // negative base, non-integer -> NaN
    check_pow(from_bits(next_u64()), from_bits(next_u64()));
    check_pow(from_bits(next_u64() & 0x000fffffffffffffull), uni(-2.0, 2.0));  // subnormal base
  }
  printf("exp: %ld inputs, %ld mismatches; pow: %ld inputs, %ld mismatches\n", n_exp, bad_exp, n_pow, bad_pow);
  return (bad_exp || bad_pow) ? 1 : 0;
}
