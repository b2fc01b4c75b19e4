#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <random>
#include "inflate_fast.h"
using namespace ccsm;
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> src(n); if (fread(src.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
  std::mt19937 rng(7);
  FastInflate* fi = new FastInflate();
  long p = 0, ok = 0, bad = 0, blocks = 0;
  while (p + 18 <= n && blocks < 60) {
    int xlen = src[p + 10] | (src[p + 11] << 8);
    int bsize = (src[p + 16] | (src[p + 17] << 8)) + 1;
    long off = p + 12 + xlen; int clen = bsize - xlen - 20;
    const uint8_t* t = &src[p + bsize - 4]; int isize = t[0] | (t[1] << 8) | (t[2] << 16) | (t[3] << 24);
    p += bsize; blocks++;
    if (!isize) continue;
    for (int trial = 0; trial < 200; ++trial) {
      uint8_t* in = (uint8_t*)malloc(clen + 8);      // payload + footer, exact size: ASan sees any overread
      memcpy(in, &src[off], clen + 8);
      uint8_t* out = (uint8_t*)malloc(isize);        // exact size: ASan sees any overwrite
      int k = 1 + rng() % 4;
      for (int j = 0; j < k; ++j) in[rng() % clen] = (uint8_t)rng();
      if (trial % 5 == 0) { int cut = rng() % clen; if (fi->run(in, cut, out, isize)) ok++; else bad++; }  // truncated too
      else if (fi->run(in, clen, out, isize)) ok++; else bad++;
      free(in); free(out);
    }
  }
  printf("blocks %ld ok %ld rejected %ld\n", blocks, ok, bad);
}
