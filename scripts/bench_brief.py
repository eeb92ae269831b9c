"""Reads one bench.py JSON line on stdin and prints the handful of numbers an A/B sweep compares."""
import json, sys
line = sys.stdin.read().strip()
try:
    d = json.loads(line)
except Exception:
    print("NOT JSON:", line[-300:]); sys.exit(0)
r = d.get("roofline", {})
t = r.get("fp64_tensor", {})
print(f"gates/s {d['value']:.0f}  ms/step {d['ms_per_step']:.1f}  sweeps {d['sweeps_per_step']} rounds {d['rounds_per_step']} passes {d.get('passes_per_step')} pairs {d.get('paired_passes_per_step')}  "
      f"ms/sweep {r.get('avg_launch_ms', 0):.2f}  hbm {r.get('frac', 0):.3f}  fp64t {t.get('frac', 0):.3f}  norm {d.get('norm')}  clocks {d.get('clocks')}")
