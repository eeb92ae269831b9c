"""Reads the [tile-prof] lines and the bench JSON line of a profiling-build run on stdin; prints one compact block."""
import json, sys
prof, line = [], None
for ln in sys.stdin:
    ln = ln.strip()
    if ln.startswith("[tile-prof]"):
        prof.append(ln)
    elif ln.startswith("{"):
        line = ln
if line:
    try:
        d = json.loads(line)
        print(f"ms/step {d['ms_per_step']:.1f}  sweeps {d['sweeps_per_step']} rounds {d['rounds_per_step']} passes {d.get('passes_per_step')} "
              f"pairs {d.get('paired_passes_per_step')}  gates/s {d['value']:.0f}")
    except Exception:
        print("NOT JSON:", line[-200:])
# the profile accumulates over all launches of the process: print the last dump only
n = 9
for ln in prof[-n:]:
    print("   ", ln)
