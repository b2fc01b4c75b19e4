#!/usr/bin/env python
"""Turn ncu outputs into the small text summaries committed under profiles/.

  summarize_ncu.py launches <launches.csv> <out.md>     per-kernel share of device time (launch list pass)
  summarize_ncu.py report <file.ncu-rep> <out.md>       key metrics of every captured launch (--set full pass)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_active.avg",
]


def launches(path, out):
    rows = []
    with open(path) as f:
        txt = "".join(l for l in f if l.startswith('"'))
    for r in csv.DictReader(io.StringIO(txt)):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            rows.append((r["Kernel Name"], float(r["Metric Value"]), r["Grid Size"], r["Block Size"]))
    tot = sum(t for _, t, _, _ in rows)
    agg = OrderedDict()
    for k, t, g, b in rows:
        k = re.sub(r"\(.*", "", k)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    with open(out, "w") as f:
        f.write("# ncu launch list (gpu__time_duration.sum, --clock-control none): %d launches, %.3f ms total\n\n" % (len(rows), tot / 1e6))
        f.write("Times are cold-cache and serialised (profiler replay): compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k[:110], n, t / 1e6, 100 * t / tot))
    print("wrote", out)


def report(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# ncu --set full summary of %s\n\n" % path.split("/")[-1])
        for r in data:
            f.write("## %s  grid %s block %s\n\n| metric | value | unit |\n|---|---:|---|\n" %
                    (r[idx["Kernel Name"]][:100], r[idx.get("Grid Size", 0)], r[idx.get("Block Size", 0)]))
            for k in KEYS:
                if k in idx:
                    f.write("| %s | %s | %s |\n" % (k, r[idx[k]], units[idx[k]]))
            if "dram__bytes_read.sum" in idx:
                f.write("\n")
        f.write("\n")
    print("wrote", out)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
