N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/BENCH_r02_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --scaling strong --total-sites 67108864 --steps 2 --warmup 3 --no-cpu-baseline --no-throughput-mode > gpurun_out/BENCH_r02_${N}gpu_strong.json 2> gpurun_out/bench_${N}gpu_strong.err
for f in gpurun_out/BENCH_r02_${N}gpu.json gpurun_out/BENCH_r02_${N}gpu_strong.json; do python - $f <<'PY'
import json,sys
l=[x for x in open(sys.argv[1]) if x.startswith('{')]
d=json.loads(l[-1]); print(sys.argv[1], d['n_gpus'], d['scaling'], round(d['value']), d['ms_per_step'], d['clocks']['sm_mhz'], d.get('e2e',{}).get('value'))
PY
done
