// Raw-DEFLATE encoder for BGZF blocks (RFC 1951): run-length matches (distance 1 only) + dynamic Huffman codes -- the
// compressor behind `CCSM_BGZF_RLE`.  Host code.
//
// Same token stream as zlib's Z_RLE strategy (a match is a run of the previous byte, length 3..258), which on HiFi
// records (packed bases, qualities, kinetics bytes) compresses as well as zlib's default strategy because LZ77 finds next
// to nothing there.  What this encoder drops is zlib's generality: no hash chains, no lazy matching, one pass to tokenise
// and count, one canonical-Huffman build per sub-block of 32 K tokens, one pass to emit -- 2-3x zlib's Z_RLE throughput.
// The reference writes BAM through htslib/zlib (call_modifications.py:410-462); any inflater reads this output.
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>

namespace ccsm {

class RleDeflate {
 public:
  // Compresses in[0..n) (n <= 65535) into one complete raw-DEFLATE stream (last block marked final).  Returns the number
  // of bytes written, or 0 if `cap` is too small (cap >= n + 64 always suffices: incompressible input is stored).
  size_t run(const uint8_t* in, int n, uint8_t* out, size_t cap) {
    if (n < 0 || n > 65535 || cap < (size_t)n + 64) return 0;
    out_ = out;
    bitbuf_ = 0;
    bitcnt_ = 0;
    int pos = 0;
    if (n == 0) {  // an empty stream: one final fixed-Huffman block holding only the end-of-block code
      put_bits(1, 1);
      put_bits(1, 2);
      put_bits(0, 7);
      flush_bits();
      return (size_t)(out_ - out);
    }
    while (pos < n) {
      const int start = pos;
      const int ntok = tokenise(in, n, &pos);
      const bool final_block = pos >= n;
      // the coded size is known before a bit is written: Huffman coding that does not pay (already compressed or
      // random bytes) is replaced by a stored block, so the output never exceeds the input by more than 5 bytes a block
      const size_t coded_bits = plan(ntok);
      const size_t raw_len = (size_t)(pos - start);
      if (coded_bits <= (raw_len + 4) * 8) {
        write_dynamic(ntok, final_block);
      } else {
        put_bits(final_block ? 1 : 0, 1);
        put_bits(0, 2);
        flush_bits();
        const uint16_t len = (uint16_t)raw_len, nlen = (uint16_t)~len;
        memcpy(out_, &len, 2);
        memcpy(out_ + 2, &nlen, 2);
        memcpy(out_ + 4, in + start, raw_len);
        out_ += 4 + raw_len;
      }
    }
    flush_bits();
    return (size_t)(out_ - out);
  }

 private:
  static constexpr int MAX_TOKENS = 32768;  // tokens per sub-block: one Huffman table each
  static constexpr int NLIT = 286, NDIST = 2, NPRE = 19;
  static constexpr uint8_t kOrder[NPRE] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  static constexpr uint16_t kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83,
                                            99, 115, 131, 163, 195, 227, 258};
  static constexpr uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};

  // ---- pass 1: tokens (literal byte, or 256 + run length - 3) and symbol frequencies
  int tokenise(const uint8_t* in, int n, int* ppos) {
    memset(freq_, 0, sizeof(freq_));
    int pos = *ppos, nt = 0;
    while (pos < n && nt < MAX_TOKENS) {
      if (pos > 0 && pos + 2 < n) {
        const uint8_t prev = in[pos - 1];
        if (in[pos] == prev && in[pos + 1] == prev && in[pos + 2] == prev) {
          int len = 3;
          const int lim = std::min(258, n - pos);
          while (len < lim && in[pos + len] == prev) ++len;
          tok_[nt++] = (uint16_t)(256 + len - 3);
          freq_[257 + len_sym(len)]++;
          pos += len;
          continue;
        }
      }
      tok_[nt++] = in[pos];
      freq_[in[pos]]++;
      ++pos;
    }
    freq_[256] = 1;  // end of block
    *ppos = pos;
    return nt;
  }

  static int len_sym(int len) {  // length 3..258 -> symbol index 0..28 (RFC 1951 3.2.5)
    static const LenTable t;
    return t.sym[len - 3];
  }
  struct LenTable {
    uint8_t sym[256];
    LenTable() {
      static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99,
                                        115, 131, 163, 195, 227, 258};
      int s = 0;
      for (int len = 3; len <= 258; ++len) {
        while (s + 1 < 29 && base[s + 1] <= len) ++s;
        sym[len - 3] = (uint8_t)s;
      }
    }
  };

  // ---- canonical Huffman code lengths, limited to `limit` bits
  // Two-queue construction over the symbols sorted by frequency; if the tree comes out deeper than the limit the
  // frequencies are flattened (halved, floor 1) and the tree rebuilt -- rare, and always terminates with a valid code.
  static void code_lengths(const uint32_t* freq_in, int nsym, int limit, uint8_t* lens) {
    uint32_t freq[NLIT];
    for (int i = 0; i < nsym; ++i) freq[i] = freq_in[i];
    for (;;) {
      int order[NLIT], m = 0;
      for (int i = 0; i < nsym; ++i) {
        lens[i] = 0;
        if (freq[i]) order[m++] = i;
      }
      if (m == 0) return;
      if (m == 1) {
        lens[order[0]] = 1;
        return;
      }
      std::sort(order, order + m, [&](int a, int b) { return freq[a] != freq[b] ? freq[a] < freq[b] : a < b; });
      // nodes 0..m-1 leaves (sorted), m..2m-2 internal in creation order (weights non-decreasing)
      uint64_t weight[2 * NLIT];
      int parent[2 * NLIT];
      for (int i = 0; i < m; ++i) weight[i] = freq[order[i]];
      int leaf = 0, inner = m, next = m;
      auto take = [&]() {
        if (leaf < m && (inner >= next || weight[leaf] <= weight[inner])) return leaf++;
        return inner++;
      };
      while (next < 2 * m - 1) {
        const int a = take(), b = take();
        weight[next] = weight[a] + weight[b];
        parent[a] = parent[b] = next;
        ++next;
      }
      int depth[2 * NLIT];
      depth[2 * m - 2] = 0;
      int maxd = 0;
      for (int i = 2 * m - 3; i >= 0; --i) {
        depth[i] = depth[parent[i]] + 1;
        if (i < m && depth[i] > maxd) maxd = depth[i];
      }
      if (maxd <= limit) {
        for (int i = 0; i < m; ++i) lens[order[i]] = (uint8_t)depth[i];
        return;
      }
      for (int i = 0; i < nsym; ++i)
        if (freq[i]) freq[i] = (freq[i] + 1) >> 1;
    }
  }

  // codes for canonical lengths, bit-reversed (DEFLATE sends Huffman codes most significant bit first)
  static void assign_codes(const uint8_t* lens, int nsym, uint16_t* codes) {
    int count[16] = {0};
    for (int i = 0; i < nsym; ++i) count[lens[i]]++;
    count[0] = 0;
    uint32_t next[16], code = 0;
    for (int len = 1; len <= 15; ++len) {
      code = (code + (uint32_t)count[len - 1]) << 1;
      next[len] = code;
    }
    for (int i = 0; i < nsym; ++i) {
      const int len = lens[i];
      if (!len) {
        codes[i] = 0;
        continue;
      }
      uint32_t c = next[len]++, r = 0;
      for (int k = 0; k < len; ++k) {
        r = (r << 1) | (c & 1);
        c >>= 1;
      }
      codes[i] = (uint16_t)r;
    }
  }

  // ---- code lengths + header layout of the sub-block just tokenised; returns its coded size in bits
  size_t plan(int ntok) {
    (void)ntok;
    code_lengths(freq_, NLIT, 15, llen_);
    hlit_ = NLIT;
    while (hlit_ > 257 && llen_[hlit_ - 1] == 0) --hlit_;
    // two distance codes of one bit each: a complete code (every inflater accepts it), symbol 0 = distance 1
    llen_[hlit_] = llen_[hlit_ + 1] = 1;
    const int total = hlit_ + NDIST;
    // code-length alphabet with the run-length symbols 16 / 17 / 18 (RFC 1951 3.2.7)
    np_ = 0;
    uint32_t pfreq[NPRE] = {0};
    for (int i = 0; i < total;) {
      const int v = llen_[i];
      int run = 1;
      while (i + run < total && llen_[i + run] == v) ++run;
      int left = run;
      if (v == 0) {
        while (left >= 11) {
          const int r = std::min(left, 138);
          pre_sym_[np_] = 18; pre_ext_[np_++] = (uint8_t)(r - 11);
          left -= r;
        }
        if (left >= 3) {
          pre_sym_[np_] = 17; pre_ext_[np_++] = (uint8_t)(left - 3);
          left = 0;
        }
      } else {
        pre_sym_[np_] = (uint8_t)v; pre_ext_[np_++] = 0;  // the value itself, then repeats of it
        --left;
        while (left >= 3) {
          const int r = std::min(left, 6);
          pre_sym_[np_] = 16; pre_ext_[np_++] = (uint8_t)(r - 3);
          left -= r;
        }
      }
      for (; left > 0; --left) {
        pre_sym_[np_] = (uint8_t)v; pre_ext_[np_++] = 0;
      }
      i += run;
    }
    for (int i = 0; i < np_; ++i) pfreq[pre_sym_[i]]++;
    code_lengths(pfreq, NPRE, 7, plen_);
    assign_codes(plen_, NPRE, pcode_);
    hclen_ = NPRE;
    while (hclen_ > 4 && plen_[kOrder[hclen_ - 1]] == 0) --hclen_;
    size_t bits = 3 + 14 + 3 * (size_t)hclen_;
    for (int i = 0; i < np_; ++i) {
      const int sy = pre_sym_[i];
      bits += plen_[sy] + (sy == 16 ? 2 : sy == 17 ? 3 : sy == 18 ? 7 : 0);
    }
    for (int sy = 0; sy < 257; ++sy) bits += (size_t)freq_[sy] * llen_[sy];
    for (int k = 0; k < 29; ++k) bits += (size_t)freq_[257 + k] * (llen_[257 + k] + kLenExtra[k] + 1);  // + 1 distance bit
    return bits;
  }

  // ---- pass 2: header + symbols
  void write_dynamic(int ntok, bool final_block) {
    put_bits(final_block ? 1 : 0, 1);
    put_bits(2, 2);
    put_bits((uint32_t)(hlit_ - 257), 5);
    put_bits((uint32_t)(NDIST - 1), 5);
    put_bits((uint32_t)(hclen_ - 4), 4);
    for (int i = 0; i < hclen_; ++i) put_bits(plen_[kOrder[i]], 3);
    for (int i = 0; i < np_; ++i) {
      const int sy = pre_sym_[i];
      put_bits(pcode_[sy], plen_[sy]);
      if (sy == 16) put_bits(pre_ext_[i], 2);
      else if (sy == 17) put_bits(pre_ext_[i], 3);
      else if (sy == 18) put_bits(pre_ext_[i], 7);
    }
    uint16_t lcode[NLIT];
    assign_codes(llen_, hlit_, lcode);
    for (int i = 0; i < ntok; ++i) {
      const int t = tok_[i];
      if (t < 256) {
        put_bits(lcode[t], llen_[t]);
      } else {
        const int len = t - 256 + 3, sy = len_sym(len);
        // length code, its extra bits, then distance symbol 0 (one bit, value 0: the canonical code of the first symbol)
        put_bits(lcode[257 + sy], llen_[257 + sy]);
        put_bits((uint32_t)(len - kLenBase[sy]), kLenExtra[sy]);
        put_bits(0, 1);
      }
    }
    put_bits(lcode[256], llen_[256]);
  }

  // ---- bit writer: up to 16 bits per call, flushed to memory 32 bits at a time
  void put_bits(uint32_t v, int n) {
    bitbuf_ |= (uint64_t)v << bitcnt_;
    bitcnt_ += n;
    if (bitcnt_ >= 32) {
      const uint32_t w = (uint32_t)bitbuf_;
      memcpy(out_, &w, 4);
      out_ += 4;
      bitbuf_ >>= 32;
      bitcnt_ -= 32;
    }
  }
  void flush_bits() {
    while (bitcnt_ > 0) {
      *out_++ = (uint8_t)bitbuf_;
      bitbuf_ >>= 8;
      bitcnt_ -= 8;
    }
    bitbuf_ = 0;
    bitcnt_ = 0;
  }

  uint8_t* out_ = nullptr;
  uint64_t bitbuf_ = 0;
  int bitcnt_ = 0;
  uint32_t freq_[NLIT];
  uint16_t tok_[MAX_TOKENS];
  uint8_t llen_[NLIT + NDIST], pre_sym_[NLIT + NDIST], pre_ext_[NLIT + NDIST], plen_[NPRE];
  uint16_t pcode_[NPRE];
  int hlit_ = 0, hclen_ = 0, np_ = 0;
};

}  // namespace ccsm
