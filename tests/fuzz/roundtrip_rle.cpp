#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <random>
#include <chrono>
#include <zlib.h>
#include "deflate_rle.h"
#include "inflate_fast.h"
using namespace ccsm;
static bool check(RleDeflate& enc, FastInflate& dec, const uint8_t* in, int n, size_t* outn) {
  std::vector<uint8_t> out(n + 64 + 8), back(n + 8), back2(n + 8);
  size_t c = enc.run(in, n, out.data(), n + 64);
  if (c == 0 && n >= 0) { printf("enc failed n=%d\n", n); return false; }
  *outn = c;
  z_stream zs = {}; inflateInit2(&zs, -15); zs.next_in = out.data(); zs.avail_in = c; zs.next_out = back.data(); zs.avail_out = n + 8;
  int rc = inflate(&zs, Z_FINISH); size_t got = zs.total_out; size_t used = zs.total_in; inflateEnd(&zs);
  if (rc != Z_STREAM_END || got != (size_t)n || used != c || memcmp(back.data(), in, n)) { printf("zlib mismatch n=%d rc=%d got=%zu used=%zu c=%zu\n", n, rc, got, used, c); return false; }
  if (n > 0 && (!dec.run(out.data(), c, back2.data(), n) || memcmp(back2.data(), in, n))) { printf("fast decoder mismatch n=%d\n", n); return false; }
  return true;
}
int main(int argc, char** argv) {
  RleDeflate* enc = new RleDeflate(); FastInflate* dec = new FastInflate();
  std::mt19937 rng(1);
  long tests = 0; size_t c;
  // synthetic families
  for (int trial = 0; trial < (argc > 2 ? atoi(argv[2]) : 3000); ++trial) {
    int n = trial < 300 ? trial : (int)(rng() % 65536);
    std::vector<uint8_t> d(n + 1);
    int fam = trial % 8;
    for (int i = 0; i < n; ++i) {
      switch (fam) {
        case 0: d[i] = (uint8_t)rng(); break;                                   // random: stored blocks
        case 1: d[i] = 0; break;                                                // one long run
        case 2: d[i] = (uint8_t)((i / (1 + trial % 300)) & 255); break;          // runs of every length
        case 3: d[i] = (uint8_t)(rng() % 4 == 0 ? rng() : 7); break;            // skewed
        case 4: d[i] = (uint8_t)(rng() % 3); break;                             // tiny alphabet
        case 5: { double u = (rng() % 100000) / 100000.0; int v = 0; while (u < 0.5 && v < 60) { u *= 2; ++v; } d[i] = (uint8_t)v; break; }  // geometric: deep trees
        case 6: d[i] = (uint8_t)(i % 2 ? 'A' : 'C'); break;                     // no runs, two symbols
        default: d[i] = (uint8_t)(i < n / 2 ? rng() : 0x55); break;             // half random, half run
      }
    }
    if (!check(*enc, *dec, d.data(), n, &c)) return 1;
    tests++;
  }
  // Fibonacci frequencies: forces the length limiter
  { std::vector<uint8_t> d; uint64_t a = 1, b = 1; for (int s = 0; s < 24 && d.size() < 60000; ++s) { for (uint64_t k = 0; k < a && d.size() < 60000; ++k) d.push_back((uint8_t)s); uint64_t t = a + b; a = b; b = t; }
    std::shuffle(d.begin(), d.end(), rng); if (!check(*enc, *dec, d.data(), (int)d.size(), &c)) return 1; tests++; }
  printf("synthetic round trips ok: %ld\n", tests);
  if (argc > 1) {
    FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> raw(n); if (fread(raw.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
    size_t tot = 0, ztot = 0; double tbest = 1e9, zbest = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      auto t0 = std::chrono::steady_clock::now(); tot = 0;
      std::vector<uint8_t> out(65280 + 128);
      for (long p = 0; p < n; p += 65280) { int m = (int)std::min<long>(65280, n - p); tot += enc->run(&raw[p], m, out.data(), m + 64); }
      tbest = std::min(tbest, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
      t0 = std::chrono::steady_clock::now(); ztot = 0;
      z_stream zs = {}; deflateInit2(&zs, 6, Z_DEFLATED, -15, 9, Z_RLE);
      for (long p = 0; p < n; p += 65280) { int m = (int)std::min<long>(65280, n - p); deflateReset(&zs); zs.next_in = &raw[p]; zs.avail_in = m; std::vector<uint8_t> o2(70000); zs.next_out = o2.data(); zs.avail_out = 70000; deflate(&zs, Z_FINISH); ztot += zs.total_out; }
      deflateEnd(&zs);
      zbest = std::min(zbest, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count());
    }
    for (long p = 0; p < n; p += 65280) { int m = (int)std::min<long>(65280, n - p); if (!check(*enc, *dec, &raw[p], m, &c)) return 1; }
    printf("file: raw %ld  rle-enc %zu (%.4f) %.1f MB/s | zlib Z_RLE %zu (%.4f) %.1f MB/s\n", n, tot, (double)tot / n, n / tbest / 1e6, ztot, (double)ztot / n, n / zbest / 1e6);
  }
}
