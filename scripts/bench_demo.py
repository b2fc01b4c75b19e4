#!/usr/bin/env python
"""Config 1/3 timing: call_mods end to end (BAM in -> device feature extraction -> forward -> MM/ML -> modbam out).

  demo    the reference's demo BAM (116 reads, 12,691 CpG sites); reference h0 stream (parity configuration)
  big     the same reads replicated REP times with distinct names (REP x 12,691 sites): steady-state throughput of the
          whole pipeline, --h0 device and --h0 reference
Prints one JSON line; the reference CPU chain's time on the build container is in demo_callmods.npz."""
import json, os, struct, sys, time
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200 import call_mods as cm, _lib
from ccsmeth_b200.bamio import BamReader, BamWriter

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "demo_callmods.npz")))
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
import tempfile
out_dir = tempfile.mkdtemp(prefix="ccsm_demo_")  # nothing large goes under gpurun_out/ (it travels back)
ckpt = os.path.join(out_dir, "model_v3.ckpt")
torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ck.items()), ckpt)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 40
modes = sys.argv[3].split(",") if len(sys.argv) > 3 else ["device", "reference"]  # h0 modes of the replicated run
# optional sweep over (host threads, --device_batch) for the replicated run, e.g. "16:16,8:16,8:32"
# a third field picks --bam_compress (rle | zlib), e.g. "16:16:zlib"
sweep = [tuple(kv.split(":")) for kv in sys.argv[4].split(",")] if len(sys.argv) > 4 else []
demo = os.path.join(ROOT, "tests", "golden", "demo", "hg002.chr20_demo.hifi.bam")


def run(inp, extra, reps, out_prefix=None, threads=None):
    times = []
    for _ in range(reps):
        args = cm.build_parser().parse_args(["-i", inp, "-m", ckpt, "-o", out_prefix or os.path.join(out_dir, "demo_out"),
                                             "--precision", prec, "--threads", str(threads or os.cpu_count())] + extra)
        t0 = time.perf_counter()
        counts, path = cm.call_mods(args)
        times.append(time.perf_counter() - t0)
    return counts, times


counts, times = run(demo, [], 4)
best = min(times[1:])
res = {"workload": "demo/hg002.chr20_demo.hifi.bam call_mods end to end (BAM in -> modbam out), 1xB200",
       "precision": prec, "sites": counts["sites"], "seconds_runs": times, "seconds_best_warm": best,
       "sites_per_s": counts["sites"] / best,
       "reference_cpu_chain_seconds_build_container": float(g["ref_cpu_seconds"]),
       "reference_cpu_threads": int(g["ref_cpu_threads"]),
       "speedup_vs_reference_chain": float(g["ref_cpu_seconds"]) / best, "host_cores": os.cpu_count()}

if rep > 0:
    tmp = out_dir
    big = os.path.join(tmp, "demo_x%d.bam" % rep)
    rd = BamReader(demo)
    recs = list(rd)
    wr = BamWriter(big, rd.header_text, rd.references, threads=os.cpu_count())
    for k in range(rep):
        for r in recs:
            nm = r.raw[32:32 + r.l_read_name - 1] + b"/%d" % k + b"\x00"
            raw = bytearray(r.raw[:32]) + nm + r.raw[32 + r.l_read_name:]
            raw[8] = len(nm)
            wr.write_raw(bytes(raw))
    wr.close()
    res["big"] = {"workload": "demo reads x%d (%d reads, %.1f MB BAM)" % (rep, rep * len(recs), os.path.getsize(big) / 1e6)}
    for mode in modes:
        c, t = run(big, ["--h0", mode], 2, os.path.join(tmp, "out"))
        res["big"]["h0_" + mode] = {"sites": c["sites"], "seconds_runs": t, "sites_per_s": c["sites"] / min(t)}
    for item in sweep:
        thr, db, comp = int(item[0]), int(item[1]), (item[2] if len(item) > 2 else "rle")
        c, t = run(big, ["--h0", "device", "--device_batch", str(db), "--bam_compress", comp], 3, os.path.join(tmp, "out"),
                   threads=thr)
        res["big"]["threads%d_device_batch%d_%s" % (thr, db, comp)] = {
            "seconds_runs": t, "sites_per_s": c["sites"] / min(t), "stages": dict(cm.TIMING),
            "out_bytes": os.path.getsize(os.path.join(tmp, "out.modbam.bam"))}
print(json.dumps(res))
