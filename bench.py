#!/usr/bin/env python
"""Throughput benchmark of the call_mods attbigru2s inference path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16] [--sites S] [--impl ours|reference]

A "step" = one pass of the hot path (ModelAttRNN forward: two-strand embedding + 3-layer BiGRU +
attention + FC/softmax) over one batch of S synthetic CpG sites per GPU (config "synthetic 1M x 21 x
feat feature tensor, attbigru2s forward, 1xB200, bf16-in/fp32-accum": S = 2^20 per GPU, weak scaling).

  value      sites/s, whole job, inputs (features + explicit h0) resident in HBM, CUDA-event timed,
             max over ranks.
  e2e        the same metric through the host-buffer C-ABI call (ccsm_forward_att2s_host via
             ModelAttRNN.forward_host): pinned host features + h0 in, probabilities out, copies inside the
             timed region.
  roofline   tensor-core bound: 244.23 MFLOP of GEMM work per site / measured bf16 peak.
  cpu_baseline  the oracle's torch-CPU port (same ATen calls as the reference forward) on the host cores.

`--impl reference` times that CPU port only (the reference's own CPU implementation of the path cannot
travel to the GPU box: it is Python importing absent I/O deps; oracle/torch_port.py issues the same ATen
calls).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SITE = 244233216.0      # SURVEY.md section 8d (MAC x 2, both strands)
FLOP_GRU_L0 = 2.0 * (709632 + 16515072)      # per site: layer-0 input + recurrent GEMMs, both strands, both dirs
FLOP_GRU_LN = 2.0 * (33030144 + 16515072)    # per site per layer >= 1
ALG_BYTES_PER_SITE = 720.0 + 12288.0  # reference 16-tensor fp32 layout + explicit fp32 h0
FEATS = ("kmer", "kpass", "ipd", "pw", "kmer2", "kpass2", "ipd2", "pw2")
METRIC = "CpG sites/sec call_mods attbigru2s seq21"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d.get("bf16_tflops_sustained", 1369.6), "bf16_tflops_burst": d.get("bf16_tflops", 1629.8),
                "hbm_gbs": d.get("hbm_gbs", 6550.7), "source": "measured"}
    return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


def load_ckpt():
    return dict(np.load(os.path.join(ROOT, "tests", "golden", "ckpt_att2s_v3.npz")))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_port_throughput(ck, batches, batch_size=512, warmup=1, threads=None):
    """The oracle's torch-CPU port on the host cores: forward over `batches` x `batch_size` sites
    (the reference's own per-call batch, call_modifications.py:668), h0 drawn per call like the reference."""
    from oracle import torch_port
    from ccsmeth_b200 import synth
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    m = torch_port.load_numpy_state(torch_port.Att2sPort(), ck)
    b = synth.make_batch(batch_size, seed=synth.SEED, with_h0=False)
    a = [b[k] for k in FEATS]
    for _ in range(warmup):
        m(*a)
    t0 = time.perf_counter()
    for _ in range(batches):
        m(*a)
    dt = time.perf_counter() - t0
    return batches * batch_size / dt, dt, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ck = load_ckpt()
    bs, per_step = 512, 4
    cores = os.cpu_count()
    cpu_port_throughput(ck, 1, bs, warmup=max(1, args.warmup) - 1 if args.warmup > 1 else 1)
    t0 = time.perf_counter()
    sites = 0
    for _ in range(args.steps):
        v, dt, _ = cpu_port_throughput(ck, per_step, bs, warmup=0)
        sites += per_step * bs
    dt = time.perf_counter() - t0
    val = sites / dt
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "sites/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "synthetic (batch,21,feat) attbigru2s forward, bounded sample: %d batches x %d sites "
                                  "per step on host cores" % (per_step, bs), "kmer_len": 21},
           "cpu_baseline": {"value": val, "unit": "sites/s", "cores": cores, "kind": "port",
                            "sample": "%d steps x %d batches x %d sites, torch %s CPU, %d threads" %
                                      (args.steps, per_step, bs, torch.__version__, cores)},
           "e2e": {"value": val, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("CCSM_BENCH_PRECISION", "bf16"))
    ap.add_argument("--sites", type=int, default=1 << 20, help="sites per GPU per step")
    ap.add_argument("--e2e-sites", type=int, default=1 << 18, help="sites per GPU per e2e step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-mode", action="store_true", help="skip the extra fp16x3 timing leg")
    ap.add_argument("--parity-sites", type=int, default=2048)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from ccsmeth_b200 import _lib, synth
    from ccsmeth_b200.models import ModelAttRNN
    from ccsmeth_b200 import parallel

    rank, world, local = parallel.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    ck = load_ckpt()
    m = ModelAttRNN(21, 3, 2, 0, 256, is_npass=True, model_type="attbigru2s", device=local)
    m.load_state_dict({k: torch.from_numpy(v) for k, v in ck.items()})
    m = m.cuda(local).eval()
    prec = args.precision
    try:
        m.set_precision(prec)
        m._ensure_handle()
    except _lib.CcsmError as e:
        if e.code != _lib.EUNSUPPORTED:
            raise
        prec = "fp32"
        m.set_precision(prec)
    S = args.sites
    # inputs resident in HBM: features (n,21) fp32 x8 + explicit h0 (6,n,256) fp32 x2, per-rank seed
    b = synth.make_batch(S, seed=synth.SEED + rank, device=dev, with_h0=False)
    g = torch.Generator(device=dev).manual_seed(synth.SEED + 1000 + rank)
    h0 = []
    for _ in range(2):
        h = torch.empty((6, S, 256), device=dev)
        for l in range(6):
            h[l].normal_(generator=g)
        h0.append(h)
    fargs = synth.to_forward_args(b)

    def step():
        return m(*fargs, h0=(h0[0], h0[1]))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    parallel.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m.profile(True)
    m.profile_read()
    l0 = _lib.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        logits, probs = step()
    ev1.record()
    torch.cuda.synchronize()
    parallel.barrier()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.kernel_launches() - l0
    ms_max = parallel.allreduce_max(ms)
    clocks = sampler.stop() if rank == 0 else None
    value = world * S * args.steps / (ms_max * 1e-3)

    # per-kernel device times (library-side CUDA events on the launching stream, recorded inside the timed region)
    prof = m.profile_read()
    m.profile(False)

    # ---- the same step in the <= 1e-4 parity mode (3-pass fp16 split), so that one line carries both numbers
    parity_leg = None
    if prec in ("bf16", "fp16") and not args.no_parity_mode:
        m.set_precision("fp16x3")
        for _ in range(2):
            step()
        torch.cuda.synchronize()
        parallel.barrier()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(max(2, args.steps // 2)):
            _, probs_x3 = step()
        pe1.record()
        torch.cuda.synchronize()
        pms = parallel.allreduce_max(pe0.elapsed_time(pe1))
        parity_leg = {"precision": "fp16x3", "value": world * S * max(2, args.steps // 2) / (pms * 1e-3), "unit": "sites/s",
                      "steps": max(2, args.steps // 2)}
        m.set_precision(prec)

    # ---- e2e: host buffers through the C-ABI host entry (ccsm_forward_att2s_host).
    # Like the reference's forward, the model draws h0 itself (models.py:77-87,125-130) -- here on the device
    # (Philox, CCSM_H0_DEVICE_RANDOM) -- so the caller hands over only the 8 feature tensors and reads back probs.
    E = min(args.e2e_sites, S)
    hb = synth.make_batch(E, seed=synth.SEED + 77 + rank, with_h0=False)
    hfeats = {k: hb[k].pin_memory() for k in FEATS}
    m.set_h0_mode("device", seed=synth.SEED + rank)
    for _ in range(max(1, args.warmup - 1)):
        m.forward_host(hfeats)
    parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        _, p_host = m.forward_host(hfeats)
    e2e_s = time.perf_counter() - t0
    e2e_s = parallel.allreduce_max(e2e_s)
    e2e_val = world * E * args.steps / e2e_s
    h2d = E * (8 * 21 * 4)
    d2h = E * 2 * 2 * 4
    # same call with an explicit host-resident h0 (parity-style use): +12,288 B/site over PCIe
    m.set_h0_mode("reference")
    E2 = min(E, 1 << 17)
    hh0 = (torch.randn(6, E2, 256).pin_memory(), torch.randn(6, E2, 256).pin_memory())
    hf2 = {k: v[:E2] for k, v in hfeats.items()}
    m.forward_host(hf2, h0=hh0)
    parallel.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.forward_host(hf2, h0=hh0)
    e2e2_s = parallel.allreduce_max(time.perf_counter() - t0)
    e2e_host_h0 = world * E2 * args.steps / e2e2_s

    # ---- end-of-run count all-reduce (the path's only collective: SURVEY.md section 8e)
    counts = parallel.allreduce_counts([S * args.steps, -(-S // 512) * args.steps, 0, 0])

    if rank != 0:
        parallel.finalize()
        return

    # ---- parity of the timed configuration on a slice, vs the CPU oracle port with the same h0
    from oracle import torch_port
    P = min(args.parity_sites, S)
    port = torch_port.load_numpy_state(torch_port.Att2sPort(), ck)
    torch.set_num_threads(os.cpu_count())
    with torch.no_grad():
        _, ref = port(*[b[k][:P].cpu() for k in FEATS], h0[0][:, :P].cpu().contiguous(), h0[1][:, :P].cpu().contiguous())
    dprob = float((probs[:P].cpu() - ref).abs().max())
    if parity_leg is not None:
        parity_leg["max_abs_dprob_vs_cpu_port"] = float((probs_x3[:P].cpu() - ref).abs().max())
        parity_leg["roofline_issued_frac"] = 3.0 * parity_leg["value"] / world * FLOP_PER_SITE / 1e12 / load_peaks()["bf16_tflops"]

    peaks = load_peaks()
    tflops = value / world * FLOP_PER_SITE / 1e12
    roof = {"bound": "tensor", "achieved": tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
            "frac": tflops / peaks["bf16_tflops"], "traffic": None, "kernel": "whole forward (all kernels)",
            "note": "per GPU vs %s sustained bf16 peak; 244.23 MFLOP/site" % peaks["source"]}
    g_ms = prof["gru_l0"][0] + prof["gru_ln"][0]
    if g_ms > 0:
        # dominant kernel = tc_gru_layer_kernel (3 launches per chunk: layer 0 with K_in=16, layers 1-2 with K_in=512)
        g_flop = prof["gru_l0"][1] * FLOP_GRU_L0 + prof["gru_ln"][1] * FLOP_GRU_LN
        g_launch = prof["gru_l0"][2] + prof["gru_ln"][2]
        ach = g_flop / (g_ms * 1e-3) / 1e12
        tot_ms = sum(v[0] for v in prof.values())
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_gru_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get(prec)
        roof = {"bound": "tensor", "kernel": "tc_gru_layer_kernel<%s>" % prec, "achieved": ach,
                "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops"],
                "traffic": traffic, "flop_per_launch": g_flop / g_launch, "ms_per_launch": g_ms / g_launch,
                "launches": g_launch, "share_of_step": g_ms / tot_ms,
                "issued_frac": (3.0 if prec.endswith("x3") else 1.0) * ach / peaks["bf16_tflops"],
                "whole_forward_tflops": tflops, "whole_forward_frac": tflops / peaks["bf16_tflops"],
                "kernel_ms": {k: round(v[0], 3) for k, v in prof.items() if v[2]},
                "note": "algorithmic GEMM FLOPs (SURVEY.md 8d) of the GRU layer launches / their CUDA-event time, "
                        "vs %s sustained bf16 cuBLAS peak; x3 modes issue 3 MMAs per algorithmic MAC (issued_frac)"
                        % peaks["source"]}
    out = {"metric": METRIC, "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp16": "f16", "fp16x3": "f16x3",
                                          "fp16c8": "f16+e4m3", "fp32": "f32"}[prec],
           "data": "synthetic",
           "config": {"workload": "synthetic %dx21xfeat per GPU, attbigru2s forward (v3 checkpoint weights), %s" % (S, prec),
                      "sites_per_gpu_per_step": S, "kmer_len": 21, "precision": prec,
                      "l2": "inputs (%.1f GB/step) larger than L2" % (S * ALG_BYTES_PER_SITE / 1e9),
                      "parallelism": "dp%d (reads sharded per rank, no data-path collective)" % world},
           "max_abs_dprob_vs_cpu_port": dprob, "parity_sites": P, "parity_mode": parity_leg,
           "e2e": {"value": e2e_val, "unit": "sites/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "sites_per_step": E, "timer": "host wall clock around ccsm_forward_att2s_host, max over ranks",
                   "h0": "drawn on device by the library (the reference's forward also draws h0 internally)",
                   "with_explicit_host_h0": {"value": e2e_host_h0, "sites_per_step": E2,
                                             "h2d_bytes_per_step": E2 * (8 * 21 * 4 + 2 * 6 * 256 * 4)}},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
           "allreduce_counts": {"sites": counts[0], "model_batches": counts[1]}}
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count()
        v, dt, th = cpu_port_throughput(ck, 64, 512, warmup=2)
        out["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": th, "kind": "port",
                               "sample": "64 batches x 512 sites (%.1f s), torch %s CPU ATen path of the reference forward"
                                         % (dt, torch.__version__)}
    print(json.dumps(out))
    parallel.finalize()


if __name__ == "__main__":
    main()
