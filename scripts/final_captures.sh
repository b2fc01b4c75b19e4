# final single-GPU captures of the round: ncu --set full per precision, launch list of the bench command, the bench lines
set -x
for prec in fp16c8 bf16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ -c 5 -o /tmp/full_$prec python scripts/ncu_target.py $prec 75776 1 > /dev/null 2>&1
  python scripts/summarize_ncu.py report /tmp/full_$prec.ncu-rep gpurun_out/r02_ncu_full_$prec.md
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-configs > gpurun_out/launches_r02.log 2>&1
python scripts/summarize_ncu.py launches gpurun_out/launches_r02.csv gpurun_out/r02_launches.md
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_r02_reference.json 2> gpurun_out/bench_ref.err
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/BENCH_r02.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/BENCH_r02.json
