"""Plan statistics for a benchmark circuit (host only, uses the test emulator's plan export).
Usage: python scripts/plan_stats.py [qubits] [key=value scheduler knobs ...]"""
import ctypes as C
import sys
from collections import Counter

import numpy as np

sys.path.insert(0, ".")
from qclojure_b200 import circuits as CI
from tests.emu import emu as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kw = dict(a.split("=") for a in sys.argv[2:])
kw = {k: int(v) for k, v in kw.items()}
depth = kw.pop("depth", 20)
circ = CI.random_brickwork_circuit(n, depth)
ops = circ["operations"]
p = E.EmuPlan(n, ops, **kw)
nw = E.lib().emu_program_words(p.h, None, 0)
buf = (C.c_uint64 * nw)()
E.lib().emu_program_words(p.h, buf, nw)
w = np.frombuffer(buf, dtype=np.uint64)
pos, ns = 4, int(w[1])
kc, rounds, matbytes, interp = Counter(), [], [], 0
for s in range(ns):
    kind = int(w[pos]); pos += 2
    if kind != 0:
        continue
    st = w[pos:]
    nr = int(st[3]); total = int(st[40])
    rounds.append(nr)
    mb = 0
    for r in range(nr):
        rd = st[48 + 40 * r: 48 + 40 * (r + 1)]
        if int(rd[17]) in (1, 2):
            k = int(rd[29]); kc[k] += 1; mb += (1 << k) * 2048
        else:
            interp += 1
    matbytes.append(mb)
    pos += total
print(f"n={n} gates={len(ops)} stages={ns} tile_sweeps={len(rounds)} rounds={sum(rounds)} interp_rounds={interp}")
print("rounds/stage histogram:", sorted(Counter(rounds).items()))
print("cond-bit k histogram:", sorted(kc.items()))
print("matrix bytes/stage: max", max(matbytes), "mean", sum(matbytes) / len(matbytes), "hist(KB)", sorted(Counter(b // 1024 for b in matbytes).items()))
# local (tile-local) vs tile-id condition bits per tensor-core round: a warp's batches share one matrix variant when the
# local condition bits fit in the bits of the batch index that select the warp (2 for 4 warps per group, 3 for 8)
pos, kl_hist = 4, Counter()
for s in range(ns):
    kind = int(w[pos]); pos += 2
    if kind != 0:
        continue
    st = w[pos:]
    nr = int(st[3]); total = int(st[40]); m = int(st[1])
    for r in range(nr):
        rd = st[48 + 40 * r: 48 + 40 * (r + 1)]
        if int(rd[17]) in (1, 2):
            k = int(rd[29])
            kl = sum(1 for j in range(k) if int(rd[30 + j]) < m)
            kl_hist[(kl, k - kl)] += 1
    pos += total
print("(local, tile-id) condition bits per round:", sorted(kl_hist.items()))
