// Host-side BGZF block codec on a thread team (SAM/BAM spec section 4.1).  The reference gets this from htslib
// through pysam (`pysam.AlignmentFile(..., threads=)`, reference extract_features.py:60-73,
// call_modifications.py:410-462); the call_mods pipeline needs it at GPU speed, so blocks are inflated /
// deflated in parallel with zlib.  Pure host code: no CUDA calls.
#include <string.h>
#include <zlib.h>

#include <atomic>
#include <thread>
#include <vector>

#include "ccsm_internal.h"

namespace ccsm {

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
      if (e[0] == 66 && e[1] == 67 && slen == 2) bsize = (e[4] | (e[5] << 8)) + 1;
      i += 4 + slen;
    }
    if (bsize < 0) return -1;
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

template <class F>
static void run_team(int threads, int64_t n_items, F&& fn) {
  if (threads < 1) threads = 1;
  if ((int64_t)threads > n_items) threads = (int)(n_items > 0 ? n_items : 1);
  std::atomic<int64_t> next{0};
  auto body = [&]() {
    for (;;) {
      const int64_t i = next.fetch_add(1);
      if (i >= n_items) break;
      fn(i);
    }
  };
  std::vector<std::thread> team;
  for (int t = 1; t < threads; ++t) team.emplace_back(body);
  body();
  for (auto& t : team) t.join();
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
  run_team(threads, (int64_t)blocks.size(), [&](int64_t i) {
    const BgzfBlock& b = blocks[i];
    if (b.isize == 0) return;
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) { bad = 1; return; }
    zs.next_in = const_cast<Bytef*>(src + b.src_off);
    zs.avail_in = (uInt)b.clen;
    zs.next_out = dst + b.dst_off;
    zs.avail_out = (uInt)b.isize;
    const int rc = inflate(&zs, Z_FINISH);
    if (rc != Z_STREAM_END || zs.avail_out != 0) bad = 1;
    inflateEnd(&zs);
    // the CRC32 of the block guards against silent corruption, like htslib
    const uint8_t* t = src + b.src_off + b.clen;
    const uint32_t want = t[0] | (t[1] << 8) | (t[2] << 16) | ((uint32_t)t[3] << 24);
    if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), dst + b.dst_off, (uInt)b.isize) != want) bad = 1;
  });
  if (bad) {
    set_error("ccsm_bgzf_inflate: corrupt BGZF block (inflate or CRC32 failed)");
    return CCSM_EINVAL;
  }
  return total;
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
  const int64_t kBlock = 65280, kSlot = 65280 + 1024;
  const int64_t nblk = (src_bytes + kBlock - 1) / kBlock;
  std::vector<int32_t> sizes((size_t)nblk, 0);
  // each block is compressed into its own slot at the END of dst, then compacted to the front in order
  std::vector<uint8_t> scratch((size_t)(nblk * kSlot));
  std::atomic<int> bad{0};
  run_team(threads, nblk, [&](int64_t i) {
    const uint8_t* in = src + i * kBlock;
    const int64_t in_n = std::min<int64_t>(kBlock, src_bytes - i * kBlock);
    uint8_t* out = scratch.data() + i * kSlot;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(out, hdr, 16);
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
    zs.next_in = const_cast<Bytef*>(in);
    zs.avail_in = (uInt)in_n;
    zs.next_out = out + 18;
    zs.avail_out = (uInt)(kSlot - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) bad = 1;
    const int64_t clen = (int64_t)zs.total_out;
    deflateEnd(&zs);
    const int64_t bsize = clen + 26;  // whole block; header stores bsize - 1
    if (bsize > 65536) { bad = 1; return; }
    out[16] = (uint8_t)((bsize - 1) & 0xff);
    out[17] = (uint8_t)((bsize - 1) >> 8);
    const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, (uInt)in_n);
    uint8_t* t = out + 18 + clen;
    for (int k = 0; k < 4; ++k) t[k] = (uint8_t)(crc >> (8 * k));
    for (int k = 0; k < 4; ++k) t[4 + k] = (uint8_t)((uint32_t)in_n >> (8 * k));
    sizes[(size_t)i] = (int32_t)bsize;
  });
  if (bad) {
    set_error("ccsm_bgzf_deflate: deflate failed");
    return CCSM_EINVAL;
  }
  int64_t o = 0;
  for (int64_t i = 0; i < nblk; ++i) {
    memcpy(dst + o, scratch.data() + i * kSlot, (size_t)sizes[(size_t)i]);
    o += sizes[(size_t)i];
  }
  return o;
}

}  // extern "C"
