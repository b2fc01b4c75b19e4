// exhaustive check of reimplemented glibc logf / sinf / cosf against the system libm on the normal_fill domain
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#ifdef USE_FMA
#define MA(a, b, c) fma((a), (b), (c))
#else
#define MA(a, b, c) ((a) * (b) + (c))
#endif
static const double T[16][2] = {
  {0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2}, {0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2},
  {0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2}, {0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3},
  {0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3}, {0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3},
  {0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4}, {0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4},
  {0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
  {0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5}, {0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4},
  {0x1.b2036576afce6p-1, 0x1.526e57720db08p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3},
  {0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2}, {0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2}};
static const double Ln2 = 0x1.62e42fefa39efp-1;
static const double A[3] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2};
static inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
float logf_g(float x) {
  uint32_t ix = asuint(x);
  if (ix == 0x3f800000) return 0;
  uint32_t tmp = ix - 0x3f330000;
  int i = (tmp >> (23 - 4)) % 16;
  int k = (int32_t)tmp >> 23;
  uint32_t iz = ix - (tmp & 0xff800000);
  double invc = T[i][0], logc = T[i][1];
  double z = (double)asfloat(iz);
  double r = MA(z, invc, -1.0);
  double y0 = MA((double)k, Ln2, logc);
  double r2 = r * r;
  double y = MA(A[1], r, A[2]);
  y = MA(A[0], r2, y);
  y = MA(y, r2, (y0 + r));
  return (float)y;
}
static const double HPI_INV = 0x1.45f306dc9c883p+23, HPI = 0x1.921fb54442d18p+0;
static const double C0 = 1.0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
                    C4 = 0x1.99343027bf8c3p-16, S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
// neg: second table entry (c coefficients negated)
static inline float sinf_poly(double x, double x2, int neg, int n) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = MA(x2, S3, S2);
    double x7 = x3 * x2;
    double s = MA(x3, S1, x);
    return (float)MA(x7, s1, s);
  } else {
    double sg = neg ? -1.0 : 1.0;
    double x4 = x2 * x2;
    double c2 = MA(x2, sg * C4, sg * C3);
    double c1 = MA(x2, sg * C1, sg * C0);
    double x6 = x4 * x2;
    double c = MA(x4, sg * C2, c1);
    return (float)MA(x6, c2, c);
  }
}
static inline uint32_t abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }
static const double SIGN[4] = {1.0, -1.0, -1.0, 1.0};
float sinf_g(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double s = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) return y;
    return sinf_poly(x, s, 0, 0);
  }
  double r = x * HPI_INV;
  int n = ((int32_t)r + 0x800000) >> 24;
  x = MA(-(double)n, HPI, x);
  double s = SIGN[n & 3];
  return sinf_poly(x * s, x * x, (n & 2) != 0, n);
}
float cosf_g(float y) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) return 1.0f;
    return sinf_poly(x, x2, 0, 1);
  }
  double r = x * HPI_INV;
  int n = ((int32_t)r + 0x800000) >> 24;
  x = MA(-(double)n, HPI, x);
  double s = SIGN[(n + 1) & 3];
  return sinf_poly(x * s, x * x, (n & 2) != 0, n ^ 1);
}
int main() {
  long bad_l = 0, bad_s = 0, bad_c = 0;
  for (uint32_t k = 0; k < (1u << 24); ++k) {
    float u = (float)((double)k * (1.0 / 16777216.0));
    float u1 = 1.0f - u;
    float a = logf(u1), b = logf_g(u1);
    if (asuint(a) != asuint(b)) { if (bad_l < 5) printf("logf(%a): libm %a mine %a\n", u1, a, b); ++bad_l; }
    float theta = (float)(6.283185307179586 * (double)u);
    float s0 = sinf(theta), s1 = sinf_g(theta);
    if (asuint(s0) != asuint(s1)) { if (bad_s < 5) printf("sinf(%a): libm %a mine %a\n", theta, s0, s1); ++bad_s; }
    float c0 = cosf(theta), c1 = cosf_g(theta);
    if (asuint(c0) != asuint(c1)) { if (bad_c < 5) printf("cosf(%a): libm %a mine %a\n", theta, c0, c1); ++bad_c; }
  }
  printf("mismatches over 2^24 inputs: logf %ld sinf %ld cosf %ld\n", bad_l, bad_s, bad_c);
  return 0;
}
