#!/usr/bin/env python
"""Config 1/3 timing: call_mods end to end on the demo BAM (116 reads, 12,691 CpG sites), 1 GPU.
Prints one JSON line; the reference CPU chain's time on the build container is in demo_callmods.npz."""
import json, os, sys, time
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200 import call_mods as cm, _lib

g = dict(np.load(os.path.join(ROOT, "tests", "golden", "demo_callmods.npz")))
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
ckpt = os.path.join(ROOT, "gpurun_out", "model_v3.ckpt")
torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ck.items()), ckpt)
prec = sys.argv[1] if len(sys.argv) > 1 else "fp16x3"
demo = os.path.join(ROOT, "tests", "golden", "demo", "hg002.chr20_demo.hifi.bam")
times = []
for rep in range(4):
    args = cm.build_parser().parse_args(["-i", demo, "-m", ckpt, "-o", os.path.join(ROOT, "gpurun_out", "demo_out"),
                                         "--precision", prec])
    t0 = time.perf_counter()
    counts, path = cm.call_mods(args)
    times.append(time.perf_counter() - t0)
best = min(times[1:])
print(json.dumps({"workload": "demo/hg002.chr20_demo.hifi.bam call_mods end to end (BAM in -> modbam out), 1xB200",
                  "precision": prec, "sites": counts["sites"], "seconds_runs": times, "seconds_best_warm": best,
                  "sites_per_s": counts["sites"] / best,
                  "reference_cpu_chain_seconds_build_container": float(g["ref_cpu_seconds"]),
                  "reference_cpu_threads": int(g["ref_cpu_threads"]),
                  "speedup_vs_reference_chain": float(g["ref_cpu_seconds"]) / best}))
