#!/usr/bin/env python
"""Per-kernel SASS evidence of libccsm.so -> profiles/sass_summary.md.

Counts, per kernel, the mnemonics that prove the Blackwell path (B200_PROFILING.md): tcgen05.mma -> UTC*MMA
(UTCHMMA = kind::f16, UTCQMMA = kind::f8f6f4), tcgen05.ld/st -> LDTM / STTM, bulk (TMA engine) copies -> UBLKCP
(1-D; UTMALDG would be tensor-map loads: none, every streamed operand is a pre-tiled image that one 1-D bulk copy lands
MMA-ready), tcgen05.commit -> UTCBAR, plus MUFU / FFMA / HFMA2 and the legacy tensor path (HMMA: none).  Also records
the sha256 of the library and of its sources so that a loaded binary can be tied to the tree.

    python scripts/sass_summary.py [--out profiles/sass_summary.md]
"""
import argparse
import hashlib
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "MUFU", "FFMA", "HFMA2", "HMMA", "LDG", "STG"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "sass_summary.md"))
    args = ap.parse_args()
    from ccsmeth_b200 import _lib
    so = _lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
    kernels = OrderedDict()
    cur = None
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = dict.fromkeys(MNEMONICS, 0)
            kernels[cur]["_instr"] = 0
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        kernels[cur]["_instr"] += 1
        for k in MNEMONICS:
            if op == k or op.startswith(k + "."):
                kernels[cur][k] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    tot = dict.fromkeys(MNEMONICS, 0)
    for c in kernels.values():
        for k in MNEMONICS:
            tot[k] += c[k]
    with open(args.out, "w") as f:
        f.write("# SASS summary of ccsmeth_b200/libccsm.so (sm_100a)\n\n")
        f.write("* library sha256 `%s`\n* sources sha256 `%s` (`ccsmeth_b200._lib.source_hash()`, printed by `__graft_entry__.build()`)\n"
                % (hashlib.sha256(open(so, "rb").read()).hexdigest(), _lib.source_hash()))
        f.write("* `cuobjdump -sass`, %d kernels; totals: %s\n\n" % (len(kernels), ", ".join("%s %d" % (k, tot[k]) for k in MNEMONICS)))
        f.write("tcgen05.mma -> `UTCHMMA` (kind::f16) / `UTCQMMA` (kind::f8f6f4, the e4m3 correction passes of fp16c8); "
                "tcgen05.ld / st -> `LDTM` / `STTM`; cp.async.bulk -> `UBLKCP` (1-D bulk copies of pre-tiled images); "
                "cp.async.bulk.tensor -> `UTMALDG` (the CTA-pair kernel's tensor-map loads, .cta_group::2); tcgen05.commit -> `UTCBAR`; mbarrier -> `SYNCS`.  `HMMA` (legacy mma.sync) must be 0.\n\n")
        f.write("| kernel | instr | " + " | ".join(MNEMONICS) + " |\n|---|---:|" + "---:|" * len(MNEMONICS) + "\n")
        for k, c in sorted(kernels.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCQMMA"]) * 100000 - kv[1]["_instr"]):
            nm = names[k].replace("ccsm::", "")
            i = nm.rfind(">(")
            nm = nm[:i + 1] if i > 0 else re.sub(r"\(.*", "", nm)
            nm = nm.replace("(int)", "").replace("(bool)", "").replace("void ", "")
            f.write("| `%s` | %d | %s |\n" % (nm[:140], c["_instr"], " | ".join(str(c[m]) for m in MNEMONICS)))
    print("wrote", args.out, "kernels", len(kernels), "totals", {k: v for k, v in tot.items() if v})


if __name__ == "__main__":
    main()
