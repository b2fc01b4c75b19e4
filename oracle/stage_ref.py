#!/usr/bin/env python
"""Stage the UNMODIFIED reference package for the CPU baseline arm -- test / measurement infrastructure only.

The reference is pure Python (no build step), so "building" it for the GPU box is copying its package directory,
byte for byte, from the read-only tree to ``oracle/_ref/ccsmeth`` (git-ignored, so no reference source enters the
history; not gpurun-ignored, so it travels to the GPU box next to libccsm.so).  ``bench.py --impl reference`` and
the ``cpu_baseline`` legs then time the reference's own ``ModelAttRNN.forward`` / ``_call_mods2s`` on the box's host
cores (``cpu_baseline.kind == "reference"``); without the staged copy they fall back to ``oracle/torch_port.py``
(``kind == "port"``).  Only tests/, bench.py's baseline legs and __graft_entry__.build() use this; nothing under
ccsmeth_b200/ imports it.

    python oracle/stage_ref.py            # no-op when /root/reference is absent
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/ccsmeth"
DST = os.path.join(HERE, "_ref", "ccsmeth")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("stage_ref: %s not present (GPU box?) -- using whatever is already staged" % SRC)
        return os.path.isdir(DST)
    n = 0
    h = hashlib.sha256()
    for root, _dirs, files in os.walk(SRC):
        rel = os.path.relpath(root, SRC)
        if "__pycache__" in rel:
            continue
        os.makedirs(os.path.join(DST, rel), exist_ok=True)
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s, d = os.path.join(root, f), os.path.join(DST, rel, f)
            data = open(s, "rb").read()
            h.update(f.encode() + data)
            if not os.path.exists(d) or open(d, "rb").read() != data:
                with open(d, "wb") as fo:
                    fo.write(data)
            n += 1
    with open(os.path.join(HERE, "_ref", "STAGED_FROM"), "w") as fo:
        fo.write("%s\nsha256(all .py) %s\nfiles %d\n" % (SRC, h.hexdigest(), n))
    if verbose:
        print("stage_ref: %d reference files -> %s (sha256 %s)" % (n, DST, h.hexdigest()[:16]))
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
