#!/bin/bash
# DRAM bytes, L2 hit rate and duration per kernel of one 75,776-site chunk: scripts/ncu_dram.sh <precision> <tag> [env assignments ...]
prec=$1; tag=$2; shift 2
out=gpurun_out/dram_${tag}.csv
env "$@" timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
  --clock-control none --csv --log-file $out python scripts/ncu_target.py $prec 75776 1 > /dev/null 2>&1
python - "$out" "$tag" <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index("Kernel Name"); mi = hdr.index("Metric Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:40]), {})[r[mi]] = float(r[vi].replace(",", ""))
for (i, k), m in d.items():
    if "tc_" not in k: continue
    print(sys.argv[2], k, "ms %.3f" % (m["gpu__time_duration.sum"] / 1e6), "rd %.2f GB" % (m["dram__bytes_read.sum"] / 1e9),
          "wr %.2f GB" % (m["dram__bytes_write.sum"] / 1e9), "l2hit %.1f" % m["lts__t_sector_hit_rate.pct"],
          "tensor %.1f" % m["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"])
PY
