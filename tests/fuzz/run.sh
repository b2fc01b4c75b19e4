#!/bin/bash
# AddressSanitizer / UBSan fuzz of the host-side parsers that see untrusted bytes:
#   fuzz_inflate      csrc/inflate_fast.h on corrupted and truncated DEFLATE payloads (exact-size heap buffers)
#   roundtrip_rle     csrc/deflate_rle.h: thousands of synthetic payloads (runs of every length, skewed / geometric /
#                     Fibonacci frequencies, random bytes) and real records -> zlib inflate + the table decoder
#   fuzz_bam_records  ccsm_bam_index / ccsm_bam_tag_records / ccsm_bam_modcalls on records with mutated header and tag fields
# Usage: tests/fuzz/run.sh            (needs g++ with -fsanitize=address,undefined; about two minutes)
set -e
cd "$(dirname "$0")/../.."
OUT=${TMPDIR:-/tmp}/ccsm_fuzz
mkdir -p "$OUT"
FLAGS="-O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer"
g++ $FLAGS -I ccsmeth_b200/csrc -o "$OUT/fuzz_inflate" tests/fuzz/fuzz_inflate.cpp
g++ $FLAGS -std=c++17 -I ccsmeth_b200/csrc -o "$OUT/roundtrip_rle" tests/fuzz/roundtrip_rle.cpp -lz
g++ $FLAGS -I include -I /usr/local/cuda/include -x c++ ccsmeth_b200/csrc/hostio.cu tests/fuzz/stub.cpp \
    tests/fuzz/fuzz_bam_records.cpp -o "$OUT/fuzz_bam_records" -lz -lpthread
export ASAN_OPTIONS=detect_leaks=0
"$OUT/fuzz_inflate" tests/golden/demo/hg002.chr20_demo.hifi.bam
"$OUT/fuzz_inflate" tests/golden/freqb/synth.aligned.modbam.bam
"$OUT/roundtrip_rle" tests/golden/demo/hg002.chr20_demo.hifi.bam 3000
"$OUT/roundtrip_rle" SURVEY.md 10
"$OUT/fuzz_bam_records" tests/golden/demo/hg002.chr20_demo.hifi.bam 0
"$OUT/fuzz_bam_records" tests/golden/freqb/synth.aligned.modbam.bam 1
echo "fuzz: no sanitizer findings"
