/* libm_sincosf_check.c -- pins the device's restatement of libm's sinf/cosf (csrc/rt_device.cuh: libm_sincosf)
 * against the C library of the machine the tests run on.
 *
 * TEST INFRASTRUCTURE.  The same algorithm (glibc >= 2.28: sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, s_sincosf.h,
 * s_sincosf_data.c -- the ARM optimized-routines sincosf) in plain C with explicit fused multiply-adds, compared
 * bit for bit with sinf()/cosf() over every binary32 in [first, last] with the given stride.
 * usage: check <first bits hex> <last bits hex> <stride>     prints: n sin_mismatches cos_mismatches */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static float poly(double x, double x2, int second_table, int n) {
  const double sg = second_table ? -1.0 : 1.0;
  if ((n & 1) == 0) {
    const double s1c = -0x1.555545995a603p-3, s2c = 0x1.1107605230bc4p-7, s3c = -0x1.994eb3774cf24p-13;
    const double x3 = x * x2, s1 = fma(x2, s3c, s2c), x7 = x3 * x2, s = fma(x3, s1c, x);
    return (float)fma(x7, s1, s);
  }
  const double c0 = sg * 0x1p0, c1 = sg * -0x1.ffffffd0c621cp-2, c2 = sg * 0x1.55553e1068f19p-5,
               c3 = sg * -0x1.6c087e89a359dp-10, c4 = sg * 0x1.99343027bf8c3p-16;
  const double x4 = x2 * x2, q2 = fma(x2, c4, c3), q1 = fma(x2, c1, c0), x6 = x4 * x2, c = fma(x4, c2, q1);
  return (float)fma(x6, q2, c);
}
static float restated(float y, int is_cos) {
  uint32_t u;
  memcpy(&u, &y, 4);
  const unsigned top = (u >> 20) & 0x7ffu;
  const double x = (double)y;
  if (top < 0x3f4u) {
    if (top < 0x398u) return is_cos ? 1.0f : y;
    return poly(x, x * x, 0, is_cos ? 1 : 0);
  }
  if (top >= 0x42fu) return is_cos ? cosf(y) : sinf(y);
  const double r = x * 0x1.45F306DC9C883p+23;
  const int n = ((int32_t)r + 0x800000) >> 24;
  const double xr = fma(-(double)n, 0x1.921FB54442D18p0, x);
  const double sign = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return poly(xr * sign, xr * xr, (n & 2) != 0, is_cos ? (n ^ 1) : n);
}
int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const uint32_t first = (uint32_t)strtoul(argv[1], 0, 16), last = (uint32_t)strtoul(argv[2], 0, 16);
  const uint32_t stride = (uint32_t)strtoul(argv[3], 0, 10);
  long n = 0, bad_s = 0, bad_c = 0;
  for (uint64_t u = first; u <= last; u += stride) {
    const uint32_t b = (uint32_t)u;
    float x;
    memcpy(&x, &b, 4);
    const float a = sinf(x), c = cosf(x), ra = restated(x, 0), rc = restated(x, 1);
    bad_s += memcmp(&a, &ra, 4) != 0;
    bad_c += memcmp(&c, &rc, 4) != 0;
    n++;
  }
  printf("%ld %ld %ld\n", n, bad_s, bad_c);
  return 0;
}
