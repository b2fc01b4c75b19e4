// Host-side BGZF block codec on a thread team (SAM/BAM spec section 4.1).  The reference gets this from htslib
// through pysam (`pysam.AlignmentFile(..., threads=)`, reference extract_features.py:60-73,
// call_modifications.py:410-462); the call_mods pipeline needs it at GPU speed, so blocks are inflated /
// deflated in parallel with zlib.  Pure host code: no CUDA calls.
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <zlib.h>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include <atomic>
#include <thread>
#include <vector>

#include <memory>

#include "ccsm_internal.h"
#include "deflate_rle.h"
#include "inflate_fast.h"

namespace ccsm {

static std::atomic<int64_t> g_inflate_fast{0}, g_inflate_zlib{0};

#if defined(__x86_64__)
// CRC-32 (gzip polynomial, reflected) by carry-less multiplication: folds 64 bytes per iteration, then reduces
// (Gopal et al., "Fast CRC computation for generic polynomials using PCLMULQDQ").  len >= 64, multiple of 16.
__attribute__((target("pclmul,sse4.1")))
static uint32_t crc32_clmul(const uint8_t* buf, size_t len, uint32_t crc) {
  static const uint64_t __attribute__((aligned(16))) k1k2[] = {0x0154442bd4ULL, 0x01c6e41596ULL};
  static const uint64_t __attribute__((aligned(16))) k3k4[] = {0x01751997d0ULL, 0x00ccaa009eULL};
  static const uint64_t __attribute__((aligned(16))) k5k0[] = {0x0163cd6124ULL, 0x0000000000ULL};
  static const uint64_t __attribute__((aligned(16))) poly[] = {0x01db710641ULL, 0x01f7011641ULL};
  __m128i x0, x1, x2, x3, x4, x5, x6, x7, x8, y5, y6, y7, y8;
  x1 = _mm_loadu_si128((const __m128i*)(buf + 0x00));
  x2 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
  x3 = _mm_loadu_si128((const __m128i*)(buf + 0x20));
  x4 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
  x1 = _mm_xor_si128(x1, _mm_cvtsi32_si128((int)crc));
  x0 = _mm_load_si128((const __m128i*)k1k2);
  buf += 64; len -= 64;
  while (len >= 64) {
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x6 = _mm_clmulepi64_si128(x2, x0, 0x00);
    x7 = _mm_clmulepi64_si128(x3, x0, 0x00); x8 = _mm_clmulepi64_si128(x4, x0, 0x00);
    x1 = _mm_clmulepi64_si128(x1, x0, 0x11); x2 = _mm_clmulepi64_si128(x2, x0, 0x11);
    x3 = _mm_clmulepi64_si128(x3, x0, 0x11); x4 = _mm_clmulepi64_si128(x4, x0, 0x11);
    y5 = _mm_loadu_si128((const __m128i*)(buf + 0x00)); y6 = _mm_loadu_si128((const __m128i*)(buf + 0x10));
    y7 = _mm_loadu_si128((const __m128i*)(buf + 0x20)); y8 = _mm_loadu_si128((const __m128i*)(buf + 0x30));
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x5), y5); x2 = _mm_xor_si128(_mm_xor_si128(x2, x6), y6);
    x3 = _mm_xor_si128(_mm_xor_si128(x3, x7), y7); x4 = _mm_xor_si128(_mm_xor_si128(x4, x8), y8);
    buf += 64; len -= 64;
  }
  x0 = _mm_load_si128((const __m128i*)k3k4);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x3), x5);
  x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
  x1 = _mm_xor_si128(_mm_xor_si128(x1, x4), x5);
  while (len >= 16) {
    x2 = _mm_loadu_si128((const __m128i*)buf);
    x5 = _mm_clmulepi64_si128(x1, x0, 0x00); x1 = _mm_clmulepi64_si128(x1, x0, 0x11);
    x1 = _mm_xor_si128(_mm_xor_si128(x1, x2), x5);
    buf += 16; len -= 16;
  }
  x2 = _mm_clmulepi64_si128(x1, x0, 0x10);
  x3 = _mm_setr_epi32(~0, 0, ~0, 0);
  x1 = _mm_srli_si128(x1, 8);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_loadl_epi64((const __m128i*)k5k0);
  x2 = _mm_srli_si128(x1, 4);
  x1 = _mm_and_si128(x1, x3);
  x1 = _mm_clmulepi64_si128(x1, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  x0 = _mm_load_si128((const __m128i*)poly);
  x2 = _mm_and_si128(x1, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x10);
  x2 = _mm_and_si128(x2, x3);
  x2 = _mm_clmulepi64_si128(x2, x0, 0x00);
  x1 = _mm_xor_si128(x1, x2);
  return (uint32_t)_mm_extract_epi32(x1, 1);
}

#endif

// gzip CRC-32 of one BGZF block.  x86-64 with PCLMULQDQ: carry-less-multiply folding over the 16-byte-multiple body
// (5x zlib's table code), zlib for the tail; the folding path is checked against zlib once per process before use.
static uint32_t block_crc32(const uint8_t* buf, size_t len) {
#if defined(__x86_64__)
  static const bool use_clmul = [] {
    if (!__builtin_cpu_supports("pclmul") || !__builtin_cpu_supports("sse4.1")) return false;
    uint8_t probe[272];
    for (int i = 0; i < 272; ++i) probe[i] = (uint8_t)(i * 131 + 7);
    for (size_t n : {(size_t)64, (size_t)80, (size_t)256, (size_t)272})
      if (~crc32_clmul(probe, n, ~0u) != (uint32_t)crc32(crc32(0L, Z_NULL, 0), probe, (uInt)n)) return false;
    return true;
  }();
  if (use_clmul && len >= 64) {
    const size_t body = len & ~(size_t)15;
    uint32_t c = ~crc32_clmul(buf, body, ~0u);
    if (len > body) c = (uint32_t)crc32(c, buf + body, (uInt)(len - body));
    return c;
  }
#endif
  return (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf, (uInt)len);
}

struct BgzfBlock {
  int64_t src_off;   // start of the deflate payload
  int32_t clen;      // payload bytes
  int32_t isize;     // inflated bytes
  int64_t dst_off;
};

// Walks complete BGZF blocks in [src, src+n).  Returns 0, or -1 on a malformed header.
static int scan_blocks(const uint8_t* src, int64_t n, std::vector<BgzfBlock>& out, int64_t* consumed, int64_t* total) {
  int64_t p = 0, d = 0;
  while (p + 18 <= n) {
    if (src[p] != 0x1f || src[p + 1] != 0x8b || src[p + 2] != 8 || !(src[p + 3] & 4)) return -1;
    const int xlen = src[p + 10] | (src[p + 11] << 8);
    if (p + 12 + xlen > n) break;
    int bsize = -1;
    for (int i = 0; i + 4 <= xlen;) {
      const uint8_t* e = src + p + 12 + i;
      const int slen = e[2] | (e[3] << 8);
      if (i + 4 + slen > xlen) return -1;  // a subfield must lie inside the extra field
      if (e[0] == 66 && e[1] == 67 && slen == 2) bsize = (e[4] | (e[5] << 8)) + 1;
      i += 4 + slen;
    }
    if (bsize < 12 + xlen + 8) return -1;  // header + extra field + CRC32 + ISIZE at the very least
    if (p + bsize > n) break;
    BgzfBlock b;
    b.src_off = p + 12 + xlen;
    b.clen = bsize - xlen - 20;
    const uint8_t* t = src + p + bsize - 4;
    b.isize = (int32_t)(t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24));
    b.dst_off = d;
    if (b.clen < 0 || b.isize < 0 || b.isize > 65536) return -1;
    d += b.isize;
    p += bsize;
    out.push_back(b);
  }
  *consumed = p;
  *total = d;
  return 0;
}

// A team of `threads` workers (the caller is one of them); each worker pulls item indices from `next` itself, so it
// can keep per-thread state (a zlib stream) across items.
template <class W>
static void run_workers(int threads, int64_t n_items, W&& worker) {
  if (threads < 1) threads = 1;
  if ((int64_t)threads > n_items) threads = (int)(n_items > 0 ? n_items : 1);
  std::atomic<int64_t> next{0};
  std::vector<std::thread> team;
  for (int t = 1; t < threads; ++t)
    team.emplace_back([&]() {
      // codec helpers yield to the thread that feeds the GPU: with reader and writer teams both as wide as the host, the
      // kernel-launch thread otherwise waits for a core (profiles/r01_demo_pipeline_sweep_before.json: forward stage
      // 0.14 s with 5 codec threads, 0.20-0.36 s with 16).  Linux applies nice values per thread.
      // CCSM_CODEC_NICE overrides the value (0 = leave the helpers at the caller's priority).
      static const int nice_by = [] {
        const char* e = getenv("CCSM_CODEC_NICE");
        return e ? atoi(e) : 10;
      }();
      if (nice_by > 0) setpriority(PRIO_PROCESS, (id_t)syscall(SYS_gettid), nice_by);
      worker(next);
    });
  worker(next);
  for (auto& t : team) t.join();
}

template <class F>
static void run_team(int threads, int64_t n_items, F&& fn) {
  run_workers(threads, n_items, [&](std::atomic<int64_t>& next) {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_items) break;
      fn(i);
    }
  });
}

}  // namespace ccsm

using namespace ccsm;

extern "C" {

int64_t ccsm_bgzf_inflated_size(const uint8_t* src, int64_t src_bytes, int64_t* consumed) {
  if (!src || src_bytes < 0 || !consumed) {
    set_error("ccsm_bgzf_inflated_size: bad argument");
    return CCSM_EINVAL;
  }
  std::vector<BgzfBlock> blocks;
  int64_t total = 0;
  if (scan_blocks(src, src_bytes, blocks, consumed, &total) != 0) {
    set_error("ccsm_bgzf_inflated_size: not a BGZF block");
    return CCSM_EINVAL;
  }
  return total;
}

int64_t ccsm_bgzf_inflate(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_cap, int32_t threads,
                          int64_t* consumed) {
  if (!src || src_bytes < 0 || !consumed || (!dst && dst_cap > 0)) {
    set_error("ccsm_bgzf_inflate: bad argument");
    return CCSM_EINVAL;
  }
  std::vector<BgzfBlock> blocks;
  int64_t total = 0;
  if (scan_blocks(src, src_bytes, blocks, consumed, &total) != 0) {
    set_error("ccsm_bgzf_inflate: not a BGZF block");
    return CCSM_EINVAL;
  }
  if (total > dst_cap) {
    set_error("ccsm_bgzf_inflate: destination too small (%lld > %lld)", (long long)total, (long long)dst_cap);
    return CCSM_EINVAL;
  }
  std::atomic<int> bad{0};
  const int64_t nblk = (int64_t)blocks.size();
  // CCSM_INFLATE=zlib keeps every block on zlib (A/B timing, cross-checks)
  const char* env = getenv("CCSM_INFLATE");
  const bool use_fast = !(env && strcmp(env, "zlib") == 0);
  run_workers(threads, nblk, [&](std::atomic<int64_t>& next) {  // one decoder state per team thread
    std::unique_ptr<FastInflate> fast(use_fast ? new (std::nothrow) FastInflate() : nullptr);
    z_stream zs;
    bool zs_ready = false;
    int64_t n_fast = 0, n_zlib = 0;
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= nblk) break;
      const BgzfBlock& b = blocks[i];
      if (b.isize == 0) continue;
      // the CRC32 of the block guards against silent corruption, like htslib
      const uint8_t* t = src + b.src_off + b.clen;
      const uint32_t want = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
      if (fast && fast->run(src + b.src_off, b.clen, dst + b.dst_off, b.isize) &&
          block_crc32(dst + b.dst_off, (size_t)b.isize) == want) {
        ++n_fast;
        continue;
      }
      // zlib decodes whatever the table decoder rejected (and reports real corruption)
      if (!zs_ready) {
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; break; }
        zs_ready = true;
      } else if (inflateReset(&zs) != Z_OK) {
        bad = 1;
        break;
      }
      zs.next_in = const_cast<Bytef*>(src + b.src_off);
      zs.avail_in = (uInt)b.clen;
      zs.next_out = dst + b.dst_off;
      zs.avail_out = (uInt)b.isize;
      const int rc = inflate(&zs, Z_FINISH);
      if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
      if (block_crc32(dst + b.dst_off, (size_t)b.isize) != want) bad = 1;
      ++n_zlib;
    }
    if (zs_ready) inflateEnd(&zs);
    g_inflate_fast += n_fast;
    g_inflate_zlib += n_zlib;
  });
  if (bad) {
    set_error("ccsm_bgzf_inflate: corrupt BGZF block (inflate or CRC32 failed)");
    return CCSM_EINVAL;
  }
  return total;
}

void ccsm_bgzf_inflate_stats(int64_t* fast_blocks, int64_t* zlib_blocks) {
  if (fast_blocks) *fast_blocks = g_inflate_fast.load();
  if (zlib_blocks) *zlib_blocks = g_inflate_zlib.load();
}

int64_t ccsm_bgzf_deflate_bound(int64_t src_bytes) {
  const int64_t nblk = (src_bytes + 65279) / 65280;
  return nblk * (65280 + 1024) + 64;
}

int64_t ccsm_bgzf_deflate(const uint8_t* src, int64_t src_bytes, uint8_t* dst, int64_t dst_cap, int32_t level,
                          int32_t threads) {
  if (src_bytes < 0 || (!src && src_bytes > 0) || !dst || dst_cap < ccsm_bgzf_deflate_bound(src_bytes)) {
    set_error("ccsm_bgzf_deflate: bad argument (dst_cap must be >= ccsm_bgzf_deflate_bound)");
    return CCSM_EINVAL;
  }
  const int strategy = (level & CCSM_BGZF_RLE) ? Z_RLE : Z_DEFAULT_STRATEGY;
  level &= 0xff;
  if (level > 9) {
    set_error("ccsm_bgzf_deflate: level %d out of range", level);
    return CCSM_EINVAL;
  }
  const int64_t kBlock = 65280, kSlot = 65280 + 1024;
  const int64_t nblk = (src_bytes + kBlock - 1) / kBlock;
  std::vector<int32_t> sizes((size_t)nblk, 0);
  // block i is compressed into its own slot dst + i * kSlot (the bound reserves one slot per block), then the blocks
  // are compacted to the front in order; one deflate state per team thread, reset between blocks
  std::atomic<int> bad{0};
  // CCSM_BGZF_RLE runs the library's own run-length + Huffman encoder (csrc/deflate_rle.h; `level` is ignored);
  // CCSM_DEFLATE=zlib keeps zlib's Z_RLE for A/B timing and cross-checks
  const char* env = getenv("CCSM_DEFLATE");
  const bool own_rle = strategy == Z_RLE && !(env && strcmp(env, "zlib") == 0);
  run_workers(threads, nblk, [&](std::atomic<int64_t>& next) {
    std::unique_ptr<RleDeflate> rle(own_rle ? new (std::nothrow) RleDeflate() : nullptr);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (!rle && deflateInit2(&zs, level, Z_DEFLATED, -15, 9, strategy) != Z_OK) { bad = 1; return; }
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= nblk) break;
      const uint8_t* in = src + i * kBlock;
      const int64_t in_n = std::min<int64_t>(kBlock, src_bytes - i * kBlock);
      uint8_t* out = dst + i * kSlot;
      static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
      memcpy(out, hdr, 16);
      int64_t clen;
      if (rle) {
        clen = (int64_t)rle->run(in, (int)in_n, out + 18, (size_t)(kSlot - 18 - 8));
        if (clen <= 0) { bad = 1; break; }
      } else {
        if (deflateReset(&zs) != Z_OK) { bad = 1; break; }
        zs.next_in = const_cast<Bytef*>(in);
        zs.avail_in = (uInt)in_n;
        zs.next_out = out + 18;
        zs.avail_out = (uInt)(kSlot - 18 - 8);
        if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad = 1; break; }
        clen = (int64_t)zs.total_out;
      }
      const int64_t bsize = clen + 26;  // whole block; header stores bsize - 1
      if (bsize > 65536) { bad = 1; break; }
      out[16] = (uint8_t)((bsize - 1) & 0xff);
      out[17] = (uint8_t)((bsize - 1) >> 8);
      const uint32_t crc = block_crc32(in, (size_t)in_n);
      uint8_t* t = out + 18 + clen;
      for (int k = 0; k < 4; ++k) t[k] = (uint8_t)(crc >> (8 * k));
      for (int k = 0; k < 4; ++k) t[4 + k] = (uint8_t)((uint32_t)in_n >> (8 * k));
      sizes[(size_t)i] = (int32_t)bsize;
    }
    if (!rle) deflateEnd(&zs);
  });
  if (bad) {
    set_error("ccsm_bgzf_deflate: deflate failed");
    return CCSM_EINVAL;
  }
  int64_t o = 0;
  for (int64_t i = 0; i < nblk; ++i) {
    if (o != i * kSlot) memmove(dst + o, dst + i * kSlot, (size_t)sizes[(size_t)i]);
    o += sizes[(size_t)i];
  }
  return o;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// BAM record indexing and re-tagging (SAM/BAM spec 4.2): the per-read host work of the call_mods pipeline that
// remains once feature extraction runs on the device -- finding each read's sequence / kinetics arrays inside the
// inflated records (reference extract_features.py:88-126 through pysam) and writing the records back with MM/ML
// (reference call_modifications.py:230-266, _bam2modbam.py:211-226).
// ------------------------------------------------------------------------------------------------------------
namespace ccsm {

static inline int32_t rd_i32(const uint8_t* p) {
  return (int32_t)((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24));
}
static inline uint32_t rd_u16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }

// size in bytes of the aux field starting at p (p[0..1] = tag, p[2] = type), or -1 if malformed / beyond end
static int64_t aux_size(const uint8_t* p, const uint8_t* end) {
  if (p + 3 > end) return -1;
  const uint8_t ty = p[2];
  int64_t sz;
  switch (ty) {
    case 'A': case 'c': case 'C': sz = 3 + 1; break;
    case 's': case 'S': sz = 3 + 2; break;
    case 'i': case 'I': case 'f': sz = 3 + 4; break;
    case 'Z': case 'H': {
      const uint8_t* q = p + 3;
      while (q < end && *q) ++q;
      if (q >= end) return -1;
      sz = (q - p) + 1;
      break;
    }
    case 'B': {
      if (p + 8 > end) return -1;
      const uint8_t sub = p[3];
      const int64_t cnt = (uint32_t)rd_i32(p + 4);
      const int w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
      if (!w) return -1;
      sz = 8 + cnt * w;
      break;
    }
    default: return -1;
  }
  return (p + sz <= end) ? sz : -1;
}

static bool aux_int(const uint8_t* p, int32_t* v) {
  switch (p[2]) {
    case 'c': *v = (int8_t)p[3]; return true;
    case 'C': *v = p[3]; return true;
    case 's': *v = (int16_t)rd_u16(p + 3); return true;
    case 'S': *v = (int32_t)rd_u16(p + 3); return true;
    case 'i': case 'I': *v = rd_i32(p + 3); return true;
    default: return false;
  }
}

}  // namespace ccsm

extern "C" {

int ccsm_bam_index(const uint8_t* buf, int64_t n_bytes, const ccsm_bam_filter* f, ccsm_bam_rec* recs, int32_t max_recs,
                   ccsm_read* descs, int32_t* n_recs, int32_t* n_descs, int64_t* consumed) {
  if (!buf || n_bytes < 0 || !f || !recs || !descs || !n_recs || !n_descs || !consumed || max_recs < 0) {
    set_error("ccsm_bam_index: bad argument");
    return CCSM_EINVAL;
  }
  int64_t p = 0;
  int32_t nr = 0, nd = 0;
  while (p + 4 <= n_bytes && nr < max_recs) {
    const int64_t len = rd_i32(buf + p);
    if (len < 32) {
      set_error("ccsm_bam_index: malformed record (block_size %lld) at offset %lld", (long long)len, (long long)p);
      return CCSM_EINVAL;
    }
    if (p + 4 + len > n_bytes) break;  // incomplete: the caller carries the tail over
    const uint8_t* r = buf + p + 4;
    const uint8_t* end = r + len;
    ccsm_bam_rec& rec = recs[nr];
    rec.off = p;
    rec.len = (int32_t)len;
    const int l_name = r[8];
    rec.mapq = r[9];
    rec.n_cigar = (int32_t)rd_u16(r + 12);
    rec.flag = (int32_t)rd_u16(r + 14);
    rec.l_seq = rd_i32(r + 16);
    const int64_t seq_off = 32 + l_name + 4LL * rec.n_cigar;
    const int64_t aux_off = seq_off + (rec.l_seq + 1) / 2 + rec.l_seq;
    if (rec.l_seq < 0 || aux_off > len) {
      set_error("ccsm_bam_index: record at offset %lld is inconsistent", (long long)p);
      return CCSM_EINVAL;
    }
    rec.aux_off = (int32_t)aux_off;
    rec.read_idx = -1;
    // read-level filters of the reference extractor (extract_features.py:269-288)
    bool use = true;
    if (f->mode_align) {
      if (rec.flag & (0x4 | 0x100 | 0x400)) use = false;
      if (f->no_supplementary && (rec.flag & 0x800)) use = false;
      if (rec.mapq < f->mapq) use = false;
      if (use && f->identity > 0.0) {
        // compute_pct_identity (process_utils.py:174-186), applied at extract_features.py:283-286
        const uint8_t* cig = r + 32 + l_name;
        double nalign = 0, nmatch = 0;
        for (int c = 0; c < rec.n_cigar; ++c) {
          const uint32_t v = (uint32_t)rd_i32(cig + 4 * c);
          const int op = v & 15;
          if (op != 4 && op != 5 && op <= 9) nalign += v >> 4;
          if (op == 0 || op == 7) nmatch += v >> 4;
        }
        if ((nalign > 0 ? nmatch / nalign : 0.0) < f->identity) use = false;
      }
    }
    // aux scan: kinetics arrays (B:C of l_seq entries), pass counts, sn
    const uint8_t* a = r + aux_off;
    int64_t koff[4] = {-1, -1, -1, -1};
    int32_t fn = 0, rn = 0;
    bool has_fn = false, has_rn = false;
    float sn[4] = {0.f, 0.f, 0.f, 0.f};
    bool bad_kin = false;
    while (a < end) {
      const int64_t sz = aux_size(a, end);
      if (sz < 0) {
        set_error("ccsm_bam_index: bad aux field in the record at offset %lld", (long long)p);
        return CCSM_EINVAL;
      }
      const char t0 = (char)a[0], t1 = (char)a[1];
      int k = -1;
      if (t0 == 'f' && t1 == 'i') k = 0;
      else if (t0 == 'r' && t1 == 'i') k = 1;
      else if (t0 == 'f' && t1 == 'p') k = 2;
      else if (t0 == 'r' && t1 == 'p') k = 3;
      if (k >= 0) {
        if (a[2] == 'B' && a[3] == 'C') {
          if (rd_i32(a + 4) == rec.l_seq) koff[k] = (a + 8) - buf;
          else bad_kin = true;  // incomplete kinetics: the reference skips the read (extract_features.py:321-326)
        } else {
          // not CodecV1 bytes (e.g. raw-frame B:S arrays): this read cannot be called here; like a read with
          // incomplete kinetics it is passed through to the output untouched instead of failing the whole run
          bad_kin = true;
        }
      } else if (t0 == 'f' && t1 == 'n') {
        has_fn = aux_int(a, &fn);
      } else if (t0 == 'r' && t1 == 'n') {
        has_rn = aux_int(a, &rn);
      } else if (t0 == 's' && t1 == 'n' && a[2] == 'B' && a[3] == 'f') {
        const int cnt = rd_i32(a + 4);
        for (int q = 0; q < 4 && q < cnt; ++q) memcpy(&sn[q], a + 8 + 4 * q, 4);
      }
      a += sz;
    }
    if (use && !bad_kin && rec.l_seq > 0 && koff[0] >= 0 && koff[1] >= 0 && koff[2] >= 0 && koff[3] >= 0) {
      ccsm_read& d = descs[nd];
      memset(&d, 0, sizeof(d));
      d.seq_off = (r + seq_off) - buf;
      d.fi_off = koff[0]; d.ri_off = koff[1]; d.fp_off = koff[2]; d.rp_off = koff[3];
      d.len = rec.l_seq;
      if (has_fn && has_rn) { d.fn = fn; d.rn = rn; }  // both or neither (extract_features.py:113-117)
      const bool reverse = rec.flag & 0x10;
      d.flags = CCSM_READ_SEQ_4BIT | (reverse ? CCSM_READ_REVERSE : 0);
      int32_t lo = 0, hi = rec.l_seq;
      if (f->mode_align && f->skip_unmapped && rec.n_cigar > 0) {
        // aligned part of the query = sequence minus the soft clips (pysam query_alignment_start / _end)
        const uint8_t* c = r + 32 + l_name;
        int32_t qs = 0, qe = rec.l_seq;
        for (int i = 0; i < rec.n_cigar; ++i) {
          const uint32_t v = (uint32_t)rd_i32(c + 4 * i);
          if ((v & 15) == 5) continue;
          if ((v & 15) == 4) qs += (int32_t)(v >> 4);
          else break;
        }
        for (int i = rec.n_cigar - 1; i >= 0; --i) {
          const uint32_t v = (uint32_t)rd_i32(c + 4 * i);
          if ((v & 15) == 5) continue;
          if ((v & 15) == 4) qe -= (int32_t)(v >> 4);
          else break;
        }
        if (reverse) { lo = rec.l_seq - qe; hi = rec.l_seq - qs; }  // extract_features.py:296-301
        else { lo = qs; hi = qe; }
      }
      d.win_lo = lo;
      d.win_hi = hi;
      if (f->want_sn) memcpy(d.sn, sn, sizeof(sn));
      rec.read_idx = nd++;
    }
    p += 4 + len;
    ++nr;
  }
  *n_recs = nr;
  *n_descs = nd;
  *consumed = p;
  return CCSM_OK;
}

int64_t ccsm_bam_tag_records(const uint8_t* buf, const ccsm_bam_rec* recs, int32_t n_recs, int32_t keep_pulse,
                             const int64_t* site_begin, const int32_t* mm, const uint8_t* ml, uint8_t* out,
                             int64_t out_cap, int32_t* n_with_mm) {
  if (!buf || !recs || n_recs < 0 || !site_begin || !out || !n_with_mm) {
    set_error("ccsm_bam_tag_records: bad argument");
    return CCSM_EINVAL;
  }
  int64_t o = 0;
  int32_t with_mm = 0;
  for (int32_t i = 0; i < n_recs; ++i) {
    const ccsm_bam_rec& rec = recs[i];
    const uint8_t* r = buf + rec.off + 4;
    const uint8_t* end = r + rec.len;
    const int64_t ns = rec.read_idx >= 0 ? site_begin[rec.read_idx + 1] - site_begin[rec.read_idx] : 0;
    // worst case: the record itself + "MMZC+m?," + 11 chars per delta + ";\0" + "MLBC" + count + bytes
    if (o + 4 + rec.len + 32 + ns * 13 > out_cap) {
      set_error("ccsm_bam_tag_records: output buffer too small");
      return CCSM_EINVAL;
    }
    uint8_t* w0 = out + o;
    uint8_t* w = w0 + 4;
    memcpy(w, r, (size_t)rec.aux_off);
    w += rec.aux_off;
    const uint8_t* a = r + rec.aux_off;
    while (a < end) {
      const int64_t sz = aux_size(a, end);
      if (sz < 0) {
        set_error("ccsm_bam_tag_records: bad aux field in record %d", i);
        return CCSM_EINVAL;
      }
      const char t0 = (char)a[0], t1 = (char)a[1];
      const bool is_mod = t0 == 'M' && (t1 == 'M' || t1 == 'L');                       // _bam2modbam.py:215-216
      const bool is_pulse = (t0 == 'f' || t0 == 'r') && (t1 == 'i' || t1 == 'p');      // :217-218
      if (!(is_mod || (is_pulse && !keep_pulse))) {
        memcpy(w, a, (size_t)sz);
        w += sz;
      }
      a += sz;
    }
    if (ns > 0 && mm && ml) {
      const int64_t b = site_begin[rec.read_idx];
      memcpy(w, "MMZC+m?,", 8);
      w += 8;
      for (int64_t k = 0; k < ns; ++k) {
        char tmp[12];
        int len = 0;
        uint32_t v = (uint32_t)mm[b + k];
        do { tmp[len++] = (char)('0' + v % 10); v /= 10; } while (v);
        while (len) *w++ = (uint8_t)tmp[--len];
        *w++ = (k + 1 < ns) ? ',' : ';';
      }
      *w++ = 0;
      memcpy(w, "MLBC", 4);
      w += 4;
      const uint32_t cnt = (uint32_t)ns;
      for (int k = 0; k < 4; ++k) *w++ = (uint8_t)(cnt >> (8 * k));
      memcpy(w, ml + b, (size_t)ns);
      w += ns;
      ++with_mm;
    }
    const uint32_t blen = (uint32_t)(w - w0 - 4);
    for (int k = 0; k < 4; ++k) w0[k] = (uint8_t)(blen >> (8 * k));
    o += 4 + blen;
  }
  *n_with_mm = with_mm;
  return o;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// call_freqb, host half: per aligned read, the modification calls carried by its MM/ML tags projected onto the
// reference through the CIGAR (reference call_mods_freq_bam.py:118-168 _get_moddict_in_tags, :457-540 the read
// loop of _readmods_to_bed_of_one_region).  Output: one tuple per call that lands on an aligned reference base.
// ------------------------------------------------------------------------------------------------------------
extern "C" {

int64_t ccsm_bam_modcalls(const uint8_t* buf, const ccsm_bam_rec* recs, int32_t n_recs, const ccsm_modcall_opts* o,
                          int32_t* ref_id, int32_t* ref_pos, uint8_t* ml, uint8_t* hap, uint8_t* strand, int64_t cap,
                          int32_t* n_reads_used) {
  if (!buf || !recs || n_recs < 0 || !o || !n_reads_used || cap < 0 ||
      (cap > 0 && (!ref_id || !ref_pos || !ml || !hap || !strand))) {
    set_error("ccsm_bam_modcalls: bad argument");
    return CCSM_EINVAL;
  }
  int64_t n_out = 0;
  int32_t used = 0;
  std::vector<int32_t> cpos;       // positions of 'C' in the forward (original) read
  std::vector<int32_t> q_ml;       // query position (alignment orientation) -> ML value, -1 = no call
  for (int32_t i = 0; i < n_recs; ++i) {
    const ccsm_bam_rec& rec = recs[i];
    // read filters (:470-480)
    if (rec.flag & (0x4 | 0x100 | 0x400)) continue;
    if (o->no_supplementary && (rec.flag & 0x800)) continue;
    if (rec.mapq < o->mapq) continue;
    const uint8_t* r = buf + rec.off + 4;
    const uint8_t* end = r + rec.len;
    const int l_name = r[8];
    const int32_t rid = rd_i32(r + 0), pos0 = rd_i32(r + 4);
    const uint8_t* cig = r + 32 + l_name;
    const uint8_t* seq = cig + 4LL * rec.n_cigar;
    const bool reverse = rec.flag & 0x10;
    // aux scan: MM:Z, ML:B:C, hap tag, NM
    const uint8_t* a = r + rec.aux_off;
    const char* mm = nullptr;
    const uint8_t* mlp = nullptr;
    int64_t ml_n = -1;
    int32_t hp = 0, nm = 0;
    bool ok = true;
    while (a < end) {
      const int64_t sz = aux_size(a, end);
      if (sz < 0) { ok = false; break; }
      const char t0 = (char)a[0], t1 = (char)a[1];
      if (t0 == 'M' && t1 == 'M' && a[2] == 'Z') mm = (const char*)a + 3;
      else if (t0 == 'M' && t1 == 'L' && a[2] == 'B' && (a[3] == 'C' || a[3] == 'c')) { mlp = a + 8; ml_n = (uint32_t)rd_i32(a + 4); }
      else if (t0 == o->hap_tag[0] && t1 == o->hap_tag[1]) {
        int32_t v;
        if (aux_int(a, &v)) hp = v;
        else if (a[2] == 'A' && a[3] >= '0' && a[3] <= '9') hp = a[3] - '0';  // the reference does int(get_tag(..))
        else if (a[2] == 'Z') {
          const char* z = (const char*)a + 3;
          int32_t acc = 0;
          bool dig = *z != 0;
          for (; *z; ++z) {
            if (*z < '0' || *z > '9' || acc > 100000000) { dig = false; break; }
            acc = acc * 10 + (*z - '0');
          }
          if (dig) hp = acc;
        }
      }
      else if (t0 == 'N' && t1 == 'M') { int32_t v; if (aux_int(a, &v)) nm = v; }
      a += sz;
    }
    if (!ok) {
      set_error("ccsm_bam_modcalls: bad aux field in record %d", i);
      return CCSM_EINVAL;
    }
    (void)nm;
    if (o->identity > 0.0) {
      // compute_pct_identity (process_utils.py:174-186): matches (M, =) over all aligned ops but clips
      double nalign = 0, nmatch = 0;
      for (int c = 0; c < rec.n_cigar; ++c) {
        const uint32_t v = (uint32_t)rd_i32(cig + 4 * c);
        const int op = v & 15;
        if (op != 4 && op != 5 && op <= 9) nalign += v >> 4;
        if (op == 0 || op == 7) nmatch += v >> 4;
      }
      const double ident = nalign > 0 ? nmatch / nalign : 0.0;
      if (ident < o->identity) continue;
    }
    ++used;
    // The read's calls: query position (alignment orientation) -> ML byte; any defect of the tags leaves the read
    // without calls (the reference's {} results, :127-168) -- with --refsites_all it still spans reference sites.
    const int32_t L = rec.l_seq;
    q_ml.assign((size_t)L, -1);
    bool have_calls = false;
    do {
      if (!mm || !mlp) break;
      // the first MM entry for C+m (optionally followed by '?' or '.'), :131-141
      const char* x = mm;
      const char* hit = nullptr;
      while (*x) {
        const char* e = x;
        while (*e && *e != ';') ++e;
        if (e - x >= 3 && x[0] == 'C' && x[1] == '+' && x[2] == 'm') { hit = x; break; }
        x = *e ? e + 1 : e;
      }
      if (!hit) break;
      const char* q = hit + 3;
      if (*q == '?' || *q == '.') ++q;
      if (*q != ',') break;  // no positions listed
      ++q;
      // C positions of the forward read
      cpos.clear();
      for (int32_t f = 0; f < L; ++f) {
        const int32_t j = reverse ? L - 1 - f : f;
        const int nib = (j & 1) ? (seq[j >> 1] & 15) : (seq[j >> 1] >> 4);
        // forward base is C  <=>  stored base is C (forward strand) or G (reverse strand)
        if (nib == (reverse ? 4 : 2)) cpos.push_back(f);
      }
      int64_t base_count = 0, n_mod = 0;
      bool bad = false;
      while (true) {
        char* endp = nullptr;
        const long d = strtol(q, &endp, 10);
        if (endp == q) { bad = true; break; }
        base_count += d + 1;  // _get_mm_position_iters (:110-116)
        if (base_count - 1 >= (int64_t)cpos.size() || base_count < 1) { bad = true; break; }  // IndexError -> {}
        if (n_mod >= ml_n) { bad = true; break; }                                             // MM longer than ML
        const int32_t fpos = cpos[(size_t)(base_count - 1)];
        const int32_t qpos = reverse ? L - 1 - fpos : fpos;
        q_ml[(size_t)qpos] = mlp[n_mod];
        ++n_mod;
        q = endp;
        if (*q == ',') { ++q; continue; }
        break;
      }
      if (bad || n_mod != ml_n) {  // assertion len(modbases) == len(mltag) (:147)
        q_ml.assign((size_t)L, -1);
        break;
      }
      have_calls = true;
    } while (false);
    if (!have_calls && !o->refsites_all) continue;
    const int hv = (hp == 1 || hp == 2) ? hp : 0;
    // aligned pairs -- matches only (M, =, X), or with --refsites_all every pair pysam lists (soft clips and
    // insertions pair with no reference base, deletions / skips with no query base) -- then the optional clip of
    // the PAIR list (:487-491)
    const bool all_pairs = o->refsites_all != 0;
    int64_t n_pairs = 0;
    for (int c = 0; c < rec.n_cigar; ++c) {
      const uint32_t v = (uint32_t)rd_i32(cig + 4 * c);
      const int op = v & 15;
      if (op == 0 || op == 7 || op == 8 || (all_pairs && (op == 1 || op == 2 || op == 3 || op == 4))) n_pairs += v >> 4;
    }
    const int64_t p_lo = o->base_clip > 0 ? o->base_clip : 0;
    const int64_t p_hi = o->base_clip > 0 ? n_pairs - o->base_clip : n_pairs;
    const uint8_t* site_mask = nullptr;  // reference motif sites of this read's strand (--refsites_all)
    int64_t mask_len = 0;
    if (all_pairs && rid >= 0 && rid < o->n_refs) {
      site_mask = (reverse ? o->sites_rev : o->sites_fwd) + o->ref_off[rid];
      mask_len = o->ref_off[rid + 1] - o->ref_off[rid];
    }
    auto emit = [&](int32_t rpos, int mlv, int zero_call) {
      if (n_out < cap) {
        ref_id[n_out] = rid;
        ref_pos[n_out] = rpos;
        ml[n_out] = (uint8_t)mlv;
        hap[n_out] = (uint8_t)hv;
        strand[n_out] = (uint8_t)((reverse ? 1 : 0) | (zero_call ? 2 : 0));
      }
      ++n_out;
    };
    int64_t pi = 0;
    int32_t qp = 0, rp = pos0;
    for (int c = 0; c < rec.n_cigar; ++c) {
      const uint32_t v = (uint32_t)rd_i32(cig + 4 * c);
      const int op = v & 15;
      const int32_t ln = (int32_t)(v >> 4);
      if (op == 0 || op == 7 || op == 8) {
        for (int32_t k = 0; k < ln; ++k, ++pi) {
          if (pi < p_lo || pi >= p_hi) continue;
          if (qp + k < L && q_ml[(size_t)(qp + k)] >= 0) emit(rp + k, q_ml[(size_t)(qp + k)], 0);
          else if (site_mask && rp + k >= 0 && rp + k < mask_len && site_mask[rp + k]) emit(rp + k, 0, 1);  // (0.0, hap) :505-509
        }
        qp += ln;
        rp += ln;
      } else if (op == 1 || op == 4) {
        if (all_pairs) pi += ln;  // (q, None): never lands on the reference
        qp += ln;
      } else if (op == 2 || op == 3) {
        if (all_pairs)
          for (int32_t k = 0; k < ln; ++k, ++pi)
            if (pi >= p_lo && pi < p_hi && site_mask && rp + k >= 0 && rp + k < mask_len && site_mask[rp + k]) emit(rp + k, 0, 1);
        rp += ln;
      }  // H, P: neither
    }
  }
  *n_reads_used = used;
  return n_out;  // > cap: the caller retries with a larger buffer
}

}  // extern "C"

extern "C" {

// Replaces the record walk of `samtools sort` / `samtools index` (reference call_modifications.py:592-607 runs both
// through pysam): per complete alignment record of an inflated BAM stream the coordinate key
// (uint32(refID) << 32 | (pos + 1) << 1 | reverse-strand), its byte range, and the fields the BAI needs.
int64_t ccsm_bam_scan_records(const uint8_t* buf, int64_t n_bytes, int64_t max_recs, uint64_t* key, int64_t* off,
                              int32_t* len, int32_t* ref_id, int32_t* pos, int32_t* end, int32_t* flag,
                              int64_t* consumed) {
  if (!buf || n_bytes < 0 || max_recs < 0 || !key || !off || !len || !ref_id || !pos || !end || !flag || !consumed) {
    ccsm::set_error("ccsm_bam_scan_records: bad argument");
    return CCSM_EINVAL;
  }
  int64_t p = 0, n = 0;
  while (p + 4 <= n_bytes && n < max_recs) {
    const int64_t bs = rd_i32(buf + p);
    if (bs < 32) {
      ccsm::set_error("ccsm_bam_scan_records: record at byte %lld has block_size %lld", (long long)p, (long long)bs);
      return CCSM_EINVAL;
    }
    if (p + 4 + bs > n_bytes) break;
    const uint8_t* r = buf + p + 4;
    const int32_t tid = rd_i32(r), ps = rd_i32(r + 4);
    const int l_name = r[8];
    const uint32_t n_cig = rd_u16(r + 12), fl = rd_u16(r + 14);
    if (32 + (int64_t)l_name + 4 * (int64_t)n_cig > bs) {
      ccsm::set_error("ccsm_bam_scan_records: record at byte %lld overruns its block_size", (long long)p);
      return CCSM_EINVAL;
    }
    int64_t rl = 0;
    const uint8_t* cg = r + 32 + l_name;
    for (uint32_t i = 0; i < n_cig; ++i) {
      const uint32_t v = (uint32_t)rd_i32(cg + 4 * i), op = v & 15;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += v >> 4;
    }
    if ((fl & 4) || rl == 0) rl = 1;
    key[n] = ((uint64_t)(uint32_t)tid << 32) | ((uint64_t)(uint32_t)(ps + 1) << 1) | ((fl & 16) ? 1u : 0u);
    off[n] = p;
    len[n] = (int32_t)(bs + 4);
    ref_id[n] = tid;
    pos[n] = ps;
    end[n] = (int32_t)(ps + rl);
    flag[n] = (int32_t)fl;
    ++n;
    p += 4 + bs;
  }
  *consumed = p;
  return n;
}

}  // extern "C"
