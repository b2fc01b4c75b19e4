// GF(2) polynomial machinery for MT19937 jump-ahead (host).
//   phi  = characteristic polynomial of the one-word transition (degree 19937), by Berlekamp-Massey on an output bit stream
//   g_J  = x^J mod phi: the state J words ahead is the XOR of the states i words ahead over the set bits i of g_J
#pragma once
#include <stdint.h>
#include <string.h>
#include <vector>

namespace mtjump {

constexpr int DEG = 19937;
constexpr int NW = (DEG + 63) / 64;  // 312 words

struct MT {  // plain generator (init_genrand + genrand_int32), used only to produce the bit stream for Berlekamp-Massey
  uint32_t s[624];
  int p;
  explicit MT(uint32_t seed) {
    s[0] = seed;
    for (int j = 1; j < 624; ++j) s[j] = 1812433253u * (s[j - 1] ^ (s[j - 1] >> 30)) + (uint32_t)j;
    p = 624;
  }
  uint32_t next() {
    if (p == 624) {
      for (int i = 0; i < 624; ++i) {
        const uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      p = 0;
    }
    return s[p++];  // untempered: any fixed linear functional of the state works
  }
};

typedef std::vector<uint64_t> Poly;  // little-endian bit vector

inline int get(const Poly& a, int i) { return (int)((a[i >> 6] >> (i & 63)) & 1u); }
inline void flip(Poly& a, int i) { a[i >> 6] ^= 1ull << (i & 63); }

// minimal polynomial of the bit sequence (length 2 * DEG + 2), Berlekamp-Massey over GF(2) with word-parallel updates
inline Poly min_poly(const std::vector<uint8_t>& seq) {
  const int n = (int)seq.size();
  const int W = (n + 64) / 64 + 1;
  Poly C(W, 0), B(W, 0), T(W, 0);
  C[0] = B[0] = 1;
  int L = 0, m = 1;
  // the sequence reversed as a bit vector: bit (n - 1 - idx) of R = s_idx, so that the 64 values s_{base}, s_{base-1}, ...
  // are 64 consecutive bits of R starting at bit n - 1 - base
  std::vector<uint64_t> R((n + 63) / 64 + 2, 0);
  for (int idx = 0; idx < n; ++idx)
    if (seq[idx]) R[(n - 1 - idx) >> 6] |= 1ull << ((n - 1 - idx) & 63);
  for (int k = 0; k < n; ++k) {
    int d = 0;
    const int wl = (L >> 6) + 1;
    uint64_t acc = 0;
    for (int w = 0; w < wl; ++w) {
      const int base = k - 64 * w;   // bits b = 0..63 of this word of C pair with s_{base - b}
      if (base < 0) break;
      const int p0 = n - 1 - base;   // first bit of the window in R; bits beyond the sequence are zero
      const int ws = p0 >> 6, bs = p0 & 63;
      uint64_t rs = R[ws] >> bs;
      if (bs) rs |= R[ws + 1] << (64 - bs);
      // indices base - b < 0 would lie past bit n - 1: those words of R are zero
      acc ^= C[w] & rs;
    }
    d = __builtin_parityll(acc);
    if (d) {
      if (2 * L <= k) {
        T = C;
        // C ^= B << m
        const int ws = m >> 6, bs = m & 63;
        for (int w = W - 1; w >= ws; --w) {
          uint64_t v = B[w - ws] << bs;
          if (bs && w - ws - 1 >= 0) v |= B[w - ws - 1] >> (64 - bs);
          C[w] ^= v;
        }
        L = k + 1 - L;
        B = T;
        m = 1;
      } else {
        const int ws = m >> 6, bs = m & 63;
        for (int w = W - 1; w >= ws; --w) {
          uint64_t v = B[w - ws] << bs;
          if (bs && w - ws - 1 >= 0) v |= B[w - ws - 1] >> (64 - bs);
          C[w] ^= v;
        }
        ++m;
      }
    } else {
      ++m;
    }
  }
  // C is the connection polynomial (reversed characteristic): phi(x) = x^L C(1/x)
  Poly phi((L + 64) / 64 + 1, 0);
  for (int i = 0; i <= L; ++i)
    if (get(C, i)) flip(phi, L - i);
  phi.resize((L + 64) / 64 + 1);
  return phi;
}

struct Field {
  Poly phi;  // degree DEG
  // a * b mod phi; a, b of degree < DEG (NW words).  64 pre-shifted copies of b turn every set bit of a into one
  // word-aligned XOR of NW + 1 words.
  Poly mulmod(const Poly& a, const Poly& b) const {
    std::vector<uint64_t> sh(64 * (NW + 1), 0);
    for (int s = 0; s < 64; ++s) {
      uint64_t* d = sh.data() + (size_t)s * (NW + 1);
      for (int w = 0; w < NW; ++w) {
        d[w] ^= b[w] << s;
        if (s) d[w + 1] ^= b[w] >> (64 - s);
      }
    }
    std::vector<uint64_t> r(2 * NW + 2, 0);
    for (int wa = 0; wa < NW; ++wa) {
      uint64_t bits = a[wa];
      while (bits) {
        const int s = __builtin_ctzll(bits);
        bits &= bits - 1;
        const uint64_t* src = sh.data() + (size_t)s * (NW + 1);
        uint64_t* dst = r.data() + wa;
        for (int w = 0; w <= NW; ++w) dst[w] ^= src[w];
      }
    }
    return reduce(r);
  }
  std::vector<int> terms;  // exponents of phi's nonzero terms below DEG (the polynomial is sparse: 135 terms)
  Poly reduce(std::vector<uint64_t>& r) const {
    for (int i = 2 * DEG - 2; i >= DEG; --i) {
      if (!((r[i >> 6] >> (i & 63)) & 1u)) continue;
      const int sh = i - DEG;
      r[i >> 6] ^= 1ull << (i & 63);
      for (int t : terms) {
        const int j = t + sh;
        r[j >> 6] ^= 1ull << (j & 63);
      }
    }
    Poly out(r.begin(), r.begin() + NW);
    out[NW - 1] &= (1ull << (DEG & 63)) - 1;
    return out;
  }
  // x^e mod phi
  Poly xpow(uint64_t e) const {
    Poly res(NW, 0), base(NW, 0);
    res[0] = 1;
    base[0] = 2;  // x
    while (e) {
      if (e & 1) res = mulmod(res, base);
      e >>= 1;
      if (e) base = mulmod(base, base);
    }
    return res;
  }
};

inline Field make_field() {
  MT g(5489u);
  std::vector<uint8_t> seq(2 * DEG + 64);
  for (auto& b : seq) b = (uint8_t)(g.next() & 1u);
  Field f;
  f.phi = min_poly(seq);
  for (int i = 0; i < DEG; ++i)
    if (get(f.phi, i)) f.terms.push_back(i);
  return f;
}

}  // namespace mtjump
