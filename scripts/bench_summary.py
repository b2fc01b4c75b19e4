"""Prints a one-line summary of a bench.py JSON line read from stdin (tag = argv[1])."""
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print(sys.argv[1] if len(sys.argv) > 1 else "", d["config"]["precision"], "value %.0f sites/s" % d["value"],
      "gru frac %.3f" % r["frac"], r.get("kernel_ms"), "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
      "e2e %.0f" % d["e2e"]["value"], "dprob %.2e" % d["max_abs_dprob_vs_cpu_port"])
