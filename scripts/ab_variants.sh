#!/bin/bash
# A/B of GRU kernel variants in one session: parity tests first (bounded), then short benches.
# usage: scripts/ab_variants.sh <outdir> <precision> <variant> [<variant> ...]
out=$1; prec=$2; shift 2
mkdir -p $out
for v in "$@"; do
  CCSM_TC_VARIANT=$v timeout 300 python -m pytest tests/test_tc_gpu.py tests/test_parity_gpu.py -x -q > $out/pytest_$v.log 2>&1
  echo "variant $v tests: $(tail -1 $out/pytest_$v.log)"
  if grep -q passed $out/pytest_$v.log && ! grep -q failed $out/pytest_$v.log; then
    CCSM_TC_VARIANT=$v timeout 300 python bench.py --steps 3 --warmup 3 --precision $prec --no-cpu-baseline > $out/bench_${prec}_$v.json 2> $out/bench_${prec}_$v.err
    python - <<PY
import json
try:
    d = json.load(open("$out/bench_${prec}_$v.json"))
    r = d["roofline"]
    print("variant $v $prec: %.0f sites/s  frac %.3f  kernel_ms %s  clocks %s" % (d["value"], r["frac"], r["kernel_ms"], d["clocks"]))
except Exception as e:
    print("variant $v bench failed:", e)
PY
  fi
done
