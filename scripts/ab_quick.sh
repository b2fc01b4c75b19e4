#!/bin/bash
# quick A/B of GRU kernel variants without the parity tests: scripts/ab_quick.sh <precision> <variant> [<variant> ...]
prec=$1; shift
for v in "$@"; do
  CCSM_TC_VARIANT=$v timeout 300 python bench.py --precision $prec --steps 3 --warmup 2 --no-cpu-baseline --no-configs --no-throughput-mode --parity-sites 256 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$prec $v', round(d['value']), d['roofline']['kernel_ms'], d['clocks']['sm_mhz'], d['clocks']['power_w_max'], d['max_abs_dprob_vs_cpu_port'])"
done
