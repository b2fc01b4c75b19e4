#!/usr/bin/env python
"""Real multi-process check of the read-stream sharding: runs call_mods (and call_freqb in count mode)
once with one rank and once under torchrun with N ranks (NCCL count all-reduce), and verifies that the rank shards
together equal the one-rank output.  Prints one JSON line.  Usage: check_multirank.py [N]"""
import json, os, subprocess, sys, tempfile, time
from collections import OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from ccsmeth_b200.bamio import BamReader

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
tmp = tempfile.mkdtemp(prefix="ccsm_mr_")
ck = dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))
ckpt = os.path.join(tmp, "m.ckpt")
torch.save(OrderedDict((k, torch.from_numpy(v)) for k, v in ck.items()), ckpt)
demo = os.path.join(ROOT, "tests", "golden", "demo", "hg002.chr20_demo.hifi.bam")
env = dict(os.environ, PYTHONPATH=ROOT)


def tags(path):
    return {r.query_name: (r.get_tag("MM"), r.get_tag("ML").tobytes()) if r.has_tag("MM") else None for r in BamReader(path)}


base = ["-i", demo, "-m", ckpt, "--h0", "zeros", "--holes_batch", "10", "--device_batch", "2"]
t0 = time.time()
subprocess.run([sys.executable, "-m", "ccsmeth_b200.call_mods"] + base + ["-o", os.path.join(tmp, "one")], check=True, env=env,
               stderr=subprocess.PIPE)
t1 = time.time()
p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(N), "--master-addr",
                    "127.0.0.1", "--master-port", "29533", "-m", "ccsmeth_b200.call_mods"] + base +
                   ["-o", os.path.join(tmp, "many")], env=env, stderr=subprocess.PIPE, text=True)
t2 = time.time()
assert p.returncode == 0, p.stderr[-2000:]
one = tags(os.path.join(tmp, "one.modbam.bam"))
# without --no_sort rank 0 merges the rank shards into one coordinate-sorted, indexed modbam (bamsort.py)
many = tags(os.path.join(tmp, "many.modbam.bam"))
assert os.path.exists(os.path.join(tmp, "many.modbam.bam.bai")) and not os.path.exists(os.path.join(tmp, "many.rank0.modbam.bam"))
log = [l for l in p.stderr.splitlines() if l.startswith("[call_mods]")]
res = {"ranks": N, "reads": len(one), "call_mods_shards_equal_single": many == one, "single_s": t1 - t0, "multi_s": t2 - t1,
       "rank0_log": log[-1] if log else ""}
# call_freqb, count mode, on the synthetic aligned modbam
d = os.path.join(ROOT, "tests", "golden", "freqb")
fb = ["--input_bam", os.path.join(d, "synth.aligned.modbam.bam"), "--ref", os.path.join(d, "synth.fa"), "--chunk_len", "10000"]
subprocess.run([sys.executable, "-m", "ccsmeth_b200.call_freqb"] + fb + ["-o", os.path.join(tmp, "f1")], check=True, env=env,
               stderr=subprocess.PIPE)
p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(N), "--master-addr",
                    "127.0.0.1", "--master-port", "29534", "-m", "ccsmeth_b200.call_freqb"] + fb +
                   ["-o", os.path.join(tmp, "fN")], env=env, stderr=subprocess.PIPE, text=True)
assert p.returncode == 0, p.stderr[-2000:]
a = open(os.path.join(tmp, "f1.count.all.freq.txt")).read().splitlines()
b = sum((open(os.path.join(tmp, "fN.rank%d.count.all.freq.txt" % r)).read().splitlines() for r in range(N)), [])
res["call_freqb_shards_equal_single"] = sorted(a) == sorted(b)
res["call_freqb_log"] = [l for l in p.stderr.splitlines() if l.startswith("[call_freqb]")][-1:]
print(json.dumps(res))
assert res["call_mods_shards_equal_single"] and res["call_freqb_shards_equal_single"]
