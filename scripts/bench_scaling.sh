#!/bin/bash
# scripts/bench_scaling.sh N [weak|strong|both]: the driver-style weak-scaling bench line and/or BASELINE config 4 (64 M sites per
# step split over the ranks) on N GPUs of one box; lines land in gpurun_out/BENCH_r02_<N>gpu[_strong].json
N=$1; what=${2:-both}
run() {  # args: output file, bench.py arguments
  out=$1; shift
  if [ "$N" = 1 ]; then timeout 1500 python bench.py "$@" > $out 2> ${out%.json}.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" > $out 2> ${out%.json}.err; fi
  python - $out <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1]); print(sys.argv[1], d['n_gpus'], d['scaling'], round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d.get('e2e',{}).get('value'))
PY
}
if [ $what != strong ]; then run gpurun_out/BENCH_r02_${N}gpu.json --steps 5 --warmup 3 --no-cpu-baseline; fi
if [ $what != weak ]; then run gpurun_out/BENCH_r02_${N}gpu_strong.json --scaling strong --total-sites 67108864 --steps 2 --warmup 3 --no-cpu-baseline --no-throughput-mode; fi
