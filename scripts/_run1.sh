CCSM_TC_VARIANT=d3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair -c 1 -o /tmp/pair python scripts/ncu_target.py bf16 75776 1 > /dev/null 2>&1
ncu -i /tmp/pair.ncu-rep --page source --csv > gpurun_out/pair_source.csv 2>/dev/null
ncu -i /tmp/pair.ncu-rep --page raw --csv > gpurun_out/pair_raw.csv 2>/dev/null
ls -la gpurun_out/pair_*.csv
scripts/ab_quick.sh bf16 ld l3 2>&1 | tail -2
