#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
// FM: bitmask of which mul-add sites are contracted (bit set = fma)
static int FM_ALL = 1;
#define MAD(a, b, c) (FM_ALL ? fmaf((a), (b), (c)) : ((a) * (b) + (c)))
static inline uint32_t asu(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float asf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
static float log_avx(float x) {
  if (x < 1.17549435e-38f) x = 1.17549435e-38f;
  int32_t imm0 = (int32_t)(asu(x) >> 23);
  x = asf((asu(x) & ~0x7f800000u) | asu(0.5f));
  imm0 -= 0x7f;
  float e = (float)imm0;
  e = e + 1.0f;
  int mask = x < 0.707106781186547524f;
  float tmp = mask ? x : 0.0f;
  x = x - 1.0f;
  e = e - (mask ? 1.0f : 0.0f);
  x = x + tmp;
  float z = x * x;
  float y = 7.0376836292E-2f;
  y = MAD(y, x, -1.1514610310E-1f);
  y = MAD(y, x, 1.1676998740E-1f);
  y = MAD(y, x, -1.2420140846E-1f);
  y = MAD(y, x, 1.4249322787E-1f);
  y = MAD(y, x, -1.6668057665E-1f);
  y = MAD(y, x, 2.0000714765E-1f);
  y = MAD(y, x, -2.4999993993E-1f);
  y = MAD(y, x, 3.3333331174E-1f);
  y = y * x;
  if (FM_ALL == 2) { float t = e * -2.12194440e-4f; y = fmaf(y, z, t); }
  else { y = y * z; y = MAD(e, -2.12194440e-4f, y); }
  y = MAD(-z, 0.5f, y);           // y - z*0.5
  x = x + y;
  x = MAD(e, 0.693359375f, x);
  return x;
}
static void sincos_avx(float x, float* s, float* c) {
  uint32_t sign_bit_sin = asu(x) & 0x80000000u;
  x = fabsf(x);
  float y = x * 1.27323954473516f;
  int32_t imm2 = (int32_t)y;
  imm2 = (imm2 + 1) & ~1;
  y = (float)imm2;
  int32_t imm4 = imm2;
  uint32_t swap_sign_bit_sin = ((uint32_t)(imm2 & 4)) << 29;
  int poly_mask = (imm2 & 2) == 0;
  x = MAD(y, -0.78515625f, x);
  x = MAD(y, -2.4187564849853515625e-4f, x);
  x = MAD(y, -3.77489497744594108e-8f, x);
  imm4 = imm4 - 2;
  uint32_t sign_bit_cos = ((uint32_t)(~imm4 & 4)) << 29;
  sign_bit_sin ^= swap_sign_bit_sin;
  float z = x * x;
  float yc = 2.443315711809948E-005f;
  yc = MAD(yc, z, -1.388731625493765E-003f);
  yc = MAD(yc, z, 4.166664568298827E-002f);
  yc = yc * z;
  if (FM_ALL == 2) { float t = z * 0.5f; yc = fmaf(yc, z, -t); }
  else { yc = yc * z; yc = MAD(-z, 0.5f, yc); }
  yc = yc + 1.0f;
  float y2 = -1.9515295891E-4f;
  y2 = MAD(y2, z, 8.3321608736E-3f);
  y2 = MAD(y2, z, -1.6666654611E-1f);
  y2 = y2 * z;
  y2 = MAD(y2, x, x);
  float ysin = poly_mask ? y2 : yc;
  float ycos = poly_mask ? yc : y2;
  *s = asf(asu(ysin) ^ sign_bit_sin);
  *c = asf(asu(ycos) ^ sign_bit_cos);
}
int main(int argc, char** argv) {
  FM_ALL = argc > 1 ? atoi(argv[1]) : 1;
  FILE* f = fopen("raw.bin", "rb"); fseek(f, 0, SEEK_END); long n = ftell(f) / 4; fseek(f, 0, SEEK_SET);
  uint32_t* raw = malloc(n * 4); float* ref = malloc(n * 4);
  fread(raw, 4, n, f); fclose(f);
  f = fopen("ref.bin", "rb"); fread(ref, 4, n, f); fclose(f);
  long bad = 0, badc = 0, bads = 0;
  for (long g = 0; g + 16 <= n; g += 16)
    for (int j = 0; j < 8; ++j) {
      float u1 = 1.0f - (float)((double)(raw[g + j] & 0xFFFFFF) * (1.0 / 16777216.0));
      float u2 = (float)((double)(raw[g + j + 8] & 0xFFFFFF) * (1.0 / 16777216.0));
      float radius = sqrtf(-2.0f * log_avx(u1));
      float theta = 6.283185307179586f * u2;   // two_pi as float times u2
      float s, c;
      sincos_avx(theta, &s, &c);
      float n1 = fmaf(radius * c, 1.0f, 0.0f), n2 = fmaf(radius * s, 1.0f, 0.0f);
      if (asu(n1) != asu(ref[g + j])) { ++badc; if (badc < 4) printf("cos g=%ld j=%d mine %a ref %a\n", g, j, n1, ref[g + j]); }
      if (asu(n2) != asu(ref[g + j + 8])) { ++bads; if (bads < 4) printf("sin g=%ld j=%d mine %a ref %a\n", g, j, n2, ref[g + j + 8]); }
    }
  printf("fma=%d: n=%ld mismatches cos %ld sin %ld\n", FM_ALL, n, badc, bads);
  return 0;
}
