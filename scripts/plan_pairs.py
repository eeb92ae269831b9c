"""Passes / rounds of the benchmark plan with paired rounds (host only, uses the test emulator's plan export).
Usage: python scripts/plan_pairs.py [qubits] [key=value scheduler knobs ...]   (env knobs: QCB_PAIR_ROUNDS, QCB_PAIR_YIELD_PCT, QCB_PAIR_COST_Q)"""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from qclojure_b200 import circuits as CI
from tests.emu import emu as E

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
kw = dict(a.split("=") for a in sys.argv[2:])
kw = {k: int(v) for k, v in kw.items()}
world = kw.pop("world", 1)
verbose = kw.pop("verbose", 1)
circ = CI.random_brickwork_circuit(n, kw.pop("depth", 20))
t0 = time.time()
p = E.EmuPlan(n, circ["operations"], world=world, **kw)
dt = time.time() - t0
nw = E.lib().emu_program_words(p.h, None, 0)
buf = (C.c_uint64 * nw)()
E.lib().emu_program_words(p.h, buf, nw)
w = np.frombuffer(buf, dtype=np.uint64)
pos, ns = 4, int(w[1])
sweeps = passes = rounds = pairs = exch = 0
cost = 0.0
for s in range(ns):
    kind = int(w[pos]); pos += 2
    if kind == 1:
        exch += 1
    if kind != 0:
        continue
    st = w[pos:]
    nr = int(st[3]); total = int(st[40]); m = int(st[1])
    desc = []
    for r in range(nr):
        rd = st[48 + 40 * r: 48 + 40 * (r + 1)]
        k = int(rd[29]); cond = sorted(int(rd[30 + j]) for j in range(k))
        s1 = sorted(int(rd[4 + j]) for j in range(3))
        if int(rd[17]) == 3:
            s2 = [int(rd[19 + j]) for j in range(3)]
            desc.append(f"[{s1}+{s2} c{cond}]"); pairs += 1; rounds += 2
        else:
            desc.append(f"{s1}c{cond}k{int(rd[17])}"); rounds += 1
    sweeps += 1; passes += nr
    if verbose:
        print(f"stage {s}: conflicts {p.max_conflict(s)} passes {nr}: " + " ".join(desc))
    pos += total
print(f"n={n} world={world} sweeps {sweeps} passes {passes} rounds {rounds} pairs {pairs} exchanges {exch} plan {dt * 1e3:.1f} ms")
