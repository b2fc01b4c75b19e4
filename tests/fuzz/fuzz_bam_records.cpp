// ASan/UBSan fuzz of the BAM record helpers (ccsm_bam_index / ccsm_bam_tag_records / ccsm_bam_modcalls) on mutated records.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <random>
#include "ccsm.h"
int main(int argc, char** argv) {
  FILE* f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> bam(n); if (fread(bam.data(), 1, n, f) != (size_t)n) return 1; fclose(f);
  int64_t used = 0; int64_t total = ccsm_bgzf_inflated_size(bam.data(), n, &used);
  std::vector<uint8_t> raw(total);
  ccsm_bgzf_inflate(bam.data(), n, raw.data(), total, 2, &used);
  // skip the BAM header
  int32_t l_text; memcpy(&l_text, &raw[4], 4); size_t p = 8 + l_text; int32_t n_ref; memcpy(&n_ref, &raw[p], 4); p += 4;
  for (int i = 0; i < n_ref; ++i) { int32_t l; memcpy(&l, &raw[p], 4); p += 8 + l; }
  const int mode = argc > 2 ? atoi(argv[2]) : 0;
  size_t len = std::min<size_t>(raw.size() - p, (size_t)3 << 20);
  std::mt19937 rng(5);
  long calls = 0, okc = 0;
  // clean index once: record offsets to aim the mutations at structural fields
  std::vector<ccsm_bam_rec> clean(8192); std::vector<ccsm_read> cleanr(8192);
  int32_t cn = 0, cd = 0; int64_t cc = 0;
  { ccsm_bam_filter flt0 = {0, 0, 0, 0, 0}; ccsm_bam_index(&raw[p], (int64_t)len, &flt0, clean.data(), 8192, cleanr.data(), &cn, &cd, &cc); }
  printf("clean records %d\n", cn);
  for (int trial = 0; trial < 400; ++trial) {
    size_t cut = trial % 3 == 0 ? rng() % len : len;   // truncated tails too
    uint8_t* buf = (uint8_t*)malloc(cut ? cut : 1);   // exact size: ASan sees any overread
    memcpy(buf, &raw[p], cut);
    int k = trial % 7 == 0 ? 0 : 1 + rng() % 24;
    for (int j = 0; j < k && cut; ++j) {
      size_t at = rng() % cut;
      const ccsm_bam_rec& r = clean[rng() % cn];
      switch (rng() % 4) {
        case 0: at = r.off + rng() % 40; break;                       // block_size + fixed header fields
        case 1: at = r.off + 4 + r.aux_off + rng() % 64; break;         // first tags: tag / type / array count bytes
        case 2: at = r.off + 4 + r.aux_off + rng() % (r.len > r.aux_off ? r.len - r.aux_off : 1); break;  // anywhere in tags
        default: break;
      }
      if (at < cut) buf[at] = (rng() % 4 == 0) ? (uint8_t)0xff : (uint8_t)rng();
    }
    ccsm_bam_filter flt = {trial & 1, 1, (trial >> 1) & 1, 1, (trial >> 2) & 1};
    int cap = 4096;
    std::vector<ccsm_bam_rec> recs(cap); std::vector<ccsm_read> reads(cap);
    int32_t nr = 0, nd = 0; int64_t consumed = 0;
    int rc = ccsm_bam_index(buf, (int64_t)cut, &flt, recs.data(), cap, reads.data(), &nr, &nd, &consumed);
    calls++;
    if (rc == 0 && nr > 0) {
      okc++;
      // re-tag with a plausible site table
      std::vector<int64_t> sb(nd + 1, 0);
      for (int i = 0; i < nd; ++i) sb[i + 1] = sb[i] + (rng() % 4);
      std::vector<int32_t> mm(sb[nd] + 1, 1); std::vector<uint8_t> ml(sb[nd] + 1, 200);
      int64_t sum = 0; for (int i = 0; i < nr; ++i) sum += recs[i].len;
      int64_t ocap = sum + 36LL * nr + 13 * sb[nd] + 64;
      uint8_t* out = (uint8_t*)malloc(ocap);
      int32_t with_mm = 0;
      ccsm_bam_tag_records(buf, recs.data(), nr, trial & 1, sb.data(), mm.data(), ml.data(), out, ocap, &with_mm);
      free(out);
      if (mode == 1) {
        ccsm_modcall_opts o; memset(&o, 0, sizeof(o)); o.mapq = 0; memcpy(o.hap_tag, "HP", 2);
        int64_t mcap = 1 << 20;
        std::vector<int32_t> rid(mcap); std::vector<int32_t> pos(mcap); std::vector<uint8_t> mlv(mcap), hap(mcap), st(mcap);
        int32_t usedr = 0;
        ccsm_bam_modcalls(buf, recs.data(), nr, &o, rid.data(), pos.data(), mlv.data(), hap.data(), st.data(), mcap, &usedr);
      }
    }
    free(buf);
  }
  printf("calls %ld indexed-ok %ld\n", calls, okc);
}
