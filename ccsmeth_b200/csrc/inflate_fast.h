// Raw-DEFLATE decoder for BGZF blocks (RFC 1951), host code.
//
// The reference reads BAM through htslib (pysam.AlignmentFile(..., threads=), extract_features.py:60-73); in the
// call_mods pipeline the inflate of the input BAM is the longest host stage once the output is written with the
// run-length strategy (profiles/r01_demo_pipeline_sweep.json).  HiFi records are literal-heavy (packed bases,
// qualities, kinetics bytes), where zlib's inflate decodes one symbol per loop iteration from a 32-bit bit buffer.
// This decoder keeps a 64-bit bit buffer refilled with one unaligned load, uses an 11-bit primary table for the
// literal/length code (8-bit for distances) with second-level tables for longer codes, and a literal fast table that
// yields two literals per lookup when both codes fit the primary index (up to eight literals per refill).  Every BGZF block carries a CRC32 and its inflated size: the caller checks both and falls
// back to zlib for any block this decoder rejects, so a decoder bug cannot produce silently wrong records.
#pragma once
#include <stdint.h>
#include <string.h>

namespace ccsm {

class FastInflate {
 public:
  // Inflates one complete raw-DEFLATE stream of `in_n` bytes into exactly `out_n` bytes.  At least 8 readable bytes
  // must follow in + in_n (the BGZF footer).  Returns true on success (stream ended with its final block and produced
  // exactly out_n bytes).
  bool run(const uint8_t* in, int64_t in_n, uint8_t* out, int64_t out_n) {
    in_next_ = in;
    in_end_ = in + in_n;
    out_begin_ = out;
    out_next_ = out;
    out_end_ = out + out_n;
    bitbuf_ = 0;
    bitsleft_ = 0;
    for (;;) {
      refill_careful();
      if (bitsleft_ < 3) return false;
      const int final_block = (int)(bitbuf_ & 1);
      const int type = (int)((bitbuf_ >> 1) & 3);
      consume(3);
      if (type == 0) {
        if (!stored_block()) return false;
      } else {
        if (type == 1) {
          if (!fixed_ready_) {
            build_fixed();
          }
          lt_ = fixed_lt_;
          dt_ = fixed_dt_;
          pt_ = fixed_pt_;
        } else if (type == 2) {
          if (!dynamic_header()) return false;
          lt_ = dyn_lt_;
          dt_ = dyn_dt_;
          pt_ = dyn_pt_;
        } else {
          return false;
        }
        if (!huffman_block()) return false;
      }
      if (final_block) break;
    }
    return out_next_ == out_end_;
  }

 private:
  static constexpr int LBITS = 11, DBITS = 8, PBITS = 7;
  static constexpr int LT_SIZE = (1 << LBITS) + 288 * 16, DT_SIZE = (1 << DBITS) + 32 * 128;
  // table entry: bits 0-3 code length, 4-7 extra bits (or second-level table bits), 8-10 kind, 15 literal flag,
  // 16-31 value
  enum { K_INVALID = 0, K_LITERAL = 1, K_BASE = 2, K_EOB = 3, K_SUB = 4 };
  static constexpr uint32_t LIT_FLAG = 0x8000u;  // set on literal entries: one test in the hot loop
  static uint32_t entry(int kind, int len, int extra, int value) {
    return (uint32_t)len | ((uint32_t)extra << 4) | ((uint32_t)kind << 8) | ((uint32_t)value << 16) |
           (kind == K_LITERAL ? LIT_FLAG : 0u);
  }
  static int e_len(uint32_t e) { return (int)(e & 15); }
  static int e_extra(uint32_t e) { return (int)((e >> 4) & 15); }
  static int e_kind(uint32_t e) { return (int)((e >> 8) & 7); }
  static int e_value(uint32_t e) { return (int)(e >> 16); }

  static uint64_t load64(const uint8_t* p) {
    uint64_t v;
    memcpy(&v, p, 8);
    return v;  // little-endian hosts only (x86-64 / aarch64)
  }
  static void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
  static void store16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }

  void consume(int n) {
    bitbuf_ >>= n;
    bitsleft_ -= n;
  }
  // byte-wise refill that never reads past in_end_
  void refill_careful() {
    while (bitsleft_ <= 56 && in_next_ < in_end_) {
      bitbuf_ |= (uint64_t)*in_next_++ << bitsleft_;
      bitsleft_ += 8;
    }
  }

  static uint32_t reverse_bits(uint32_t code, int len) {
    uint32_t r = 0;
    for (int i = 0; i < len; ++i) {
      r = (r << 1) | (code & 1);
      code >>= 1;
    }
    return r;
  }

  // Builds a two-level decode table for canonical Huffman code lengths lens[0..n).  value_of(sym, &kind, &extra, &value)
  // describes the symbol.  Returns false for over-subscribed codes; unused slots of incomplete codes stay K_INVALID.
  template <class Describe>
  bool build_table(const uint8_t* lens, int n, int tbits, uint32_t* table, int table_size, Describe describe) {
    int count[16] = {0};
    for (int i = 0; i < n; ++i) count[lens[i]]++;
    count[0] = 0;
    int64_t left = 1;
    for (int len = 1; len <= 15; ++len) {
      left <<= 1;
      left -= count[len];
      if (left < 0) return false;  // over-subscribed
    }
    uint32_t next_code[16];
    uint32_t code = 0;
    for (int len = 1; len <= 15; ++len) {
      code = (code + (uint32_t)count[len - 1]) << 1;
      next_code[len] = code;
    }
    const int tsize = 1 << tbits;
    memset(table, 0, sizeof(uint32_t) * (size_t)tsize);
    // longest code under every primary slot that needs a second-level table
    uint8_t sub_max[1 << LBITS];
    bool any_long = false;
    uint32_t codes[320];
    for (int sym = 0; sym < n; ++sym) {
      const int len = lens[sym];
      if (len == 0) continue;
      const uint32_t rc = reverse_bits(next_code[len]++, len);
      codes[sym] = rc;
      if (len > tbits) {
        if (!any_long) {
          memset(sub_max, 0, (size_t)tsize);
          any_long = true;
        }
        uint8_t& m = sub_max[rc & (uint32_t)(tsize - 1)];
        if (len > m) m = (uint8_t)len;
      }
    }
    int next_free = tsize;
    if (any_long) {
      for (int p = 0; p < tsize; ++p) {
        if (!sub_max[p]) continue;
        const int sbits = sub_max[p] - tbits;
        if (next_free + (1 << sbits) > table_size) return false;
        table[p] = entry(K_SUB, 0, sbits, next_free);
        memset(table + next_free, 0, sizeof(uint32_t) << sbits);
        next_free += 1 << sbits;
      }
    }
    for (int sym = 0; sym < n; ++sym) {
      const int len = lens[sym];
      if (len == 0) continue;
      int kind, extra, value;
      describe(sym, &kind, &extra, &value);
      const uint32_t e = entry(kind, len, extra, value);
      const uint32_t rc = codes[sym];
      if (len <= tbits) {
        for (uint32_t i = rc; i < (uint32_t)tsize; i += 1u << len) table[i] = e;
      } else {
        const uint32_t pe = table[rc & (uint32_t)(tsize - 1)];
        const int sbits = e_extra(pe);
        uint32_t* sub = table + e_value(pe);
        for (uint32_t i = rc >> tbits; i < (1u << sbits); i += 1u << (len - tbits)) sub[i] = e;
      }
    }
    return true;
  }

  static void describe_litlen(int sym, int* kind, int* extra, int* value) {
    static const uint16_t base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99,
                                      115, 131, 163, 195, 227, 258};
    static const uint8_t ext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    if (sym < 256) {
      *kind = K_LITERAL; *extra = 0; *value = sym;
    } else if (sym == 256) {
      *kind = K_EOB; *extra = 0; *value = 0;
    } else if (sym <= 285) {
      *kind = K_BASE; *extra = ext[sym - 257]; *value = base[sym - 257];
    } else {
      *kind = K_INVALID; *extra = 0; *value = 0;  // 286, 287 never appear in valid data
    }
  }
  static void describe_dist(int sym, int* kind, int* extra, int* value) {
    static const uint16_t base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537,
                                      2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint8_t ext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    if (sym < 30) {
      *kind = K_BASE; *extra = ext[sym]; *value = base[sym];
    } else {
      *kind = K_INVALID; *extra = 0; *value = 0;
    }
  }
  static void describe_precode(int sym, int* kind, int* extra, int* value) {
    *kind = K_LITERAL; *extra = 0; *value = sym;
  }

  // Literal fast table for the hot loop: pt[i] = total code bits | count << 4 | lit1 << 8 | lit2 << 16, where count is
  // 2 when the LBITS bits i hold two complete literal codes, 1 for one, 0 when the first symbol is not a primary-table
  // literal (the loop then goes through lt).
  static void build_pairs(const uint32_t* lt, uint32_t* pt) {
    for (uint32_t i = 0; i < (1u << LBITS); ++i) {
      const uint32_t e = lt[i];
      if (!(e & LIT_FLAG)) {
        pt[i] = 0;
        continue;
      }
      const uint32_t l1 = e & 15;
      const uint32_t e2 = lt[i >> l1];  // the bits after the first code, zero-extended
      if ((e2 & LIT_FLAG) && l1 + (e2 & 15) <= (uint32_t)LBITS)
        pt[i] = (l1 + (e2 & 15)) | (2u << 4) | ((e >> 16) << 8) | ((e2 >> 16) << 16);
      else
        pt[i] = l1 | (1u << 4) | ((e >> 16) << 8);
    }
  }

  void build_fixed() {
    uint8_t lens[288 + 32];
    for (int i = 0; i < 144; ++i) lens[i] = 8;
    for (int i = 144; i < 256; ++i) lens[i] = 9;
    for (int i = 256; i < 280; ++i) lens[i] = 7;
    for (int i = 280; i < 288; ++i) lens[i] = 8;
    for (int i = 0; i < 32; ++i) lens[288 + i] = 5;
    build_table(lens, 288, LBITS, fixed_lt_, LT_SIZE, describe_litlen);
    build_table(lens + 288, 32, DBITS, fixed_dt_, DT_SIZE, describe_dist);
    build_pairs(fixed_lt_, fixed_pt_);
    fixed_ready_ = true;
  }

  bool stored_block() {
    // drop the rest of the current byte, hand whole prefetched bytes back to the input
    consume(bitsleft_ & 7);
    in_next_ -= bitsleft_ >> 3;
    bitbuf_ = 0;
    bitsleft_ = 0;
    if (in_end_ - in_next_ < 4) return false;
    const uint32_t len = (uint32_t)in_next_[0] | ((uint32_t)in_next_[1] << 8);
    const uint32_t nlen = (uint32_t)in_next_[2] | ((uint32_t)in_next_[3] << 8);
    in_next_ += 4;
    if ((len ^ 0xffffu) != nlen) return false;
    if ((int64_t)len > in_end_ - in_next_ || (int64_t)len > out_end_ - out_next_) return false;
    memcpy(out_next_, in_next_, len);
    in_next_ += len;
    out_next_ += len;
    return true;
  }

  bool dynamic_header() {
    refill_careful();
    if (bitsleft_ < 14) return false;
    const int hlit = (int)(bitbuf_ & 31) + 257;
    const int hdist = (int)((bitbuf_ >> 5) & 31) + 1;
    const int hclen = (int)((bitbuf_ >> 10) & 15) + 4;
    consume(14);
    if (hlit > 286 || hdist > 30) return false;
    static const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t plens[19] = {0};
    for (int i = 0; i < hclen; ++i) {
      refill_careful();
      if (bitsleft_ < 3) return false;
      plens[order[i]] = (uint8_t)(bitbuf_ & 7);
      consume(3);
    }
    uint32_t ptable[1 << PBITS];
    if (!build_table(plens, 19, PBITS, ptable, 1 << PBITS, describe_precode)) return false;
    uint8_t lens[288 + 32 + 140];
    const int total = hlit + hdist;
    int i = 0;
    while (i < total) {
      refill_careful();
      const uint32_t e = ptable[bitbuf_ & ((1u << PBITS) - 1)];
      if (e_kind(e) != K_LITERAL || e_len(e) > bitsleft_) return false;
      consume(e_len(e));
      const int sym = e_value(e);
      if (sym < 16) {
        lens[i++] = (uint8_t)sym;
        continue;
      }
      int rep, val = 0;
      if (sym == 16) {
        if (i == 0 || bitsleft_ < 2) return false;
        val = lens[i - 1];
        rep = 3 + (int)(bitbuf_ & 3);
        consume(2);
      } else if (sym == 17) {
        if (bitsleft_ < 3) return false;
        rep = 3 + (int)(bitbuf_ & 7);
        consume(3);
      } else {
        if (bitsleft_ < 7) return false;
        rep = 11 + (int)(bitbuf_ & 127);
        consume(7);
      }
      if (i + rep > total) return false;
      memset(lens + i, val, (size_t)rep);
      i += rep;
    }
    if (lens[256] == 0) return false;  // no end-of-block code
    if (!build_table(lens, hlit, LBITS, dyn_lt_, LT_SIZE, describe_litlen)) return false;
    if (!build_table(lens + hlit, hdist, DBITS, dyn_dt_, DT_SIZE, describe_dist)) return false;
    build_pairs(dyn_lt_, dyn_pt_);
    return true;
  }

  bool huffman_block() {
    const uint32_t* lt = lt_;
    const uint32_t* dt = dt_;
    const uint32_t* pt = pt_;
    const uint8_t* in_next = in_next_;
    uint8_t* out_next = out_next_;
    uint64_t bitbuf = bitbuf_;
    int bitsleft = bitsleft_;
    constexpr uint32_t LMASK = (1u << LBITS) - 1, DMASK = (1u << DBITS) - 1;
    // ---- fast loop: 8-byte loads are in bounds (the BGZF footer follows the payload) and there is room for the
    // longest match plus the overshoot of the word copies
    const uint8_t* in_fast_end = in_end_;  // loads of 8 bytes at in_next <= in_end_ stay inside payload + footer
    uint8_t* const out_end = out_end_;
    bool done = false;
    while (in_next <= in_fast_end && out_end - out_next >= 258 + 16) {
      bitbuf |= load64(in_next) << bitsleft;
      in_next += 7 - ((bitsleft >> 3) & 7);
      bitsleft |= 56;
      uint32_t p = pt[bitbuf & LMASK];
      if (p & 0x30u) {
        // literal run: up to four lookups (<= LBITS bits each, >= 56 in the buffer), one or two literals per lookup;
        // two bytes are always stored, the cursor moves by the count
        store16(out_next, (uint16_t)(p >> 8));
        out_next += (p >> 4) & 3;
        bitbuf >>= (p & 15); bitsleft -= (int)(p & 15);
        p = pt[bitbuf & LMASK];
        if (p & 0x30u) {
          store16(out_next, (uint16_t)(p >> 8));
          out_next += (p >> 4) & 3;
          bitbuf >>= (p & 15); bitsleft -= (int)(p & 15);
          p = pt[bitbuf & LMASK];
          if (p & 0x30u) {
            store16(out_next, (uint16_t)(p >> 8));
            out_next += (p >> 4) & 3;
            bitbuf >>= (p & 15); bitsleft -= (int)(p & 15);
            p = pt[bitbuf & LMASK];
            if (p & 0x30u) {
              store16(out_next, (uint16_t)(p >> 8));
              out_next += (p >> 4) & 3;
              bitbuf >>= (p & 15); bitsleft -= (int)(p & 15);
            }
          }
        }
        continue;
      }
      uint32_t e = lt[bitbuf & LMASK];
      if (e_kind(e) == K_SUB) e = lt[e_value(e) + ((bitbuf >> LBITS) & ((1u << e_extra(e)) - 1))];
      bitbuf >>= e_len(e); bitsleft -= e_len(e);
      const int kind = e_kind(e);
      if (kind == K_LITERAL) {
        *out_next++ = (uint8_t)e_value(e);
        continue;
      }
      if (kind == K_EOB) {
        done = true;
        break;
      }
      if (kind != K_BASE) return false;
      const int length = e_value(e) + (int)(bitbuf & ((1u << e_extra(e)) - 1));
      bitbuf >>= e_extra(e); bitsleft -= e_extra(e);
      uint32_t d = dt[bitbuf & DMASK];
      if (e_kind(d) == K_SUB) d = dt[e_value(d) + ((bitbuf >> DBITS) & ((1u << e_extra(d)) - 1))];
      if (e_kind(d) != K_BASE) return false;
      bitbuf >>= e_len(d); bitsleft -= e_len(d);
      const int64_t dist = e_value(d) + (int64_t)(bitbuf & ((1u << e_extra(d)) - 1));
      bitbuf >>= e_extra(d); bitsleft -= e_extra(d);
      if (bitsleft < 0 || dist > out_next - out_begin_) return false;
      uint8_t* dst = out_next;
      uint8_t* const end = out_next + length;
      const uint8_t* src = out_next - dist;
      if (dist >= 8) {
        do {
          store64(dst, load64(src));
          dst += 8; src += 8;
        } while (dst < end);
      } else if (dist == 1) {
        const uint64_t v = 0x0101010101010101ULL * src[0];
        do {
          store64(dst, v);
          dst += 8;
        } while (dst < end);
      } else {
        do {
          *dst++ = *src++;
        } while (dst < end);
      }
      out_next = end;
    }
    // ---- careful loop near the ends of the buffers: byte-wise refill, every write bounds-checked
    if (bitsleft < 64) bitbuf &= (bitsleft > 0 ? (~0ULL >> (64 - bitsleft)) : 0ULL);  // keep only the counted bits
    in_next_ = in_next;
    out_next_ = out_next;
    bitbuf_ = bitbuf;
    bitsleft_ = bitsleft;
    while (!done) {
      refill_careful();
      uint32_t e = lt[bitbuf_ & LMASK];
      if (e_kind(e) == K_SUB) e = lt[e_value(e) + ((bitbuf_ >> LBITS) & ((1u << e_extra(e)) - 1))];
      const int kind = e_kind(e);
      if (kind == K_INVALID || e_len(e) > bitsleft_) return false;
      consume(e_len(e));
      if (kind == K_LITERAL) {
        if (out_next_ >= out_end_) return false;
        *out_next_++ = (uint8_t)e_value(e);
        continue;
      }
      if (kind == K_EOB) break;
      if (kind != K_BASE || e_extra(e) > bitsleft_) return false;
      const int length = e_value(e) + (int)(bitbuf_ & ((1u << e_extra(e)) - 1));
      consume(e_extra(e));
      refill_careful();
      uint32_t d = dt[bitbuf_ & DMASK];
      if (e_kind(d) == K_SUB) d = dt[e_value(d) + ((bitbuf_ >> DBITS) & ((1u << e_extra(d)) - 1))];
      if (e_kind(d) != K_BASE || e_len(d) + e_extra(d) > bitsleft_) return false;
      consume(e_len(d));
      const int64_t dist = e_value(d) + (int64_t)(bitbuf_ & ((1u << e_extra(d)) - 1));
      consume(e_extra(d));
      if (dist > out_next_ - out_begin_ || length > out_end_ - out_next_) return false;
      const uint8_t* src = out_next_ - dist;
      for (int k = 0; k < length; ++k) out_next_[k] = src[k];
      out_next_ += length;
    }
    return true;
  }

  const uint8_t* in_next_ = nullptr;
  const uint8_t* in_end_ = nullptr;
  uint8_t* out_begin_ = nullptr;
  uint8_t* out_next_ = nullptr;
  uint8_t* out_end_ = nullptr;
  uint64_t bitbuf_ = 0;
  int bitsleft_ = 0;
  const uint32_t* lt_ = nullptr;
  const uint32_t* dt_ = nullptr;
  const uint32_t* pt_ = nullptr;
  bool fixed_ready_ = false;
  uint32_t dyn_lt_[LT_SIZE], dyn_dt_[DT_SIZE], fixed_lt_[LT_SIZE], fixed_dt_[DT_SIZE];
  uint32_t dyn_pt_[1 << LBITS], fixed_pt_[1 << LBITS];
};

}  // namespace ccsm
