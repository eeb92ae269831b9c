"""Summarises the [tile-trace] lines of a -DQCB_TILE_TRACE build (kernels.cu: TileTrace): per launch and per consumer warp the
cycles between the hand-over points of a pass, and how the two consumer groups and the mover interleave.
Usage: python scripts/trace_summary.py trace.log"""
import sys
from collections import defaultdict

NAMES = {0: "tile wanted", 1: "tile arrived", 9: "pass top", 2: "barrier passed", 3: "setup done", 4: "first loads issued", 5: "2 peeled calls done",
         6: "prefetch issued", 7: "loop left", 8: "results stored", 10: "tile released", 20: "mover: loads issued", 21: "mover: tile done seen",
         22: "mover: written back"}
rows = []
for ln in open(sys.argv[1]):
    if ln.startswith("[tile-trace]"):
        _, i, w, j, r, c = ln.split()
        rows.append((int(i), int(w), int(j), int(r), int(c)))
# launches: the 40-bit clock restarts / jumps between launches; split where the clock of consecutive records goes back by a lot or
# forward by more than 5e6 cycles
launches, cur, last = [], [], None
for rec in rows:
    c = rec[4]
    if last is not None and (c < last - (1 << 20) or c > last + 5_000_000):
        launches.append(cur); cur = []
    cur.append(rec); last = c
if cur:
    launches.append(cur)
print(f"{len(rows)} records, {len(launches)} launches")
seg_tot = defaultdict(list)
for li, L in enumerate(launches):
    by = defaultdict(list)          # (warp, tile, pass) -> [(id, clk)]
    for i, w, j, r, c in L:
        by[(w, j, r)].append((i, c))
    t0 = min(c for *_, c in L)
    # per consumer warp and pass: segments
    for (w, j, r), ev in sorted(by.items()):
        if w >= 8 or r == 15:
            continue
        d = dict(ev)
        def seg(a, b):
            return d[b] - d[a] if a in d and b in d else None
        for name, a, b in (("barrier wait", 9, 2), ("setup", 2, 3), ("to first loads", 3, 4), ("first block + 2 calls", 4, 5), ("prefetch", 5, 6), ("loop", 6, 7),
                           ("tail calls + stores", 7, 8), ("pass total", 9, 8)):
            v = seg(a, b)
            if v is not None:
                seg_tot[name].append(v)
    if li in (2, len(launches) // 2):
        print(f"-- launch {li}: timeline of warps 0 (group 0), 4 (group 1), 8 (mover), cycles from the launch's first record")
        for w in (0, 4, 8):
            ev = sorted((c - t0, j, r, i) for (ww, j, r), e in by.items() if ww == w for i, c in e)
            print(f"   warp {w}: " + "  ".join(f"{c}:{'t%d' % j}{'p%d' % r if r != 15 else ''}:{i}" for c, j, r, i in ev[:60]))
print("-- consumer segments over all launches (cycles): median / mean / p90")
for name in ("barrier wait", "setup", "to first loads", "first block + 2 calls", "prefetch", "loop", "tail calls + stores", "pass total"):
    v = sorted(seg_tot[name])
    if v:
        print(f"   {name:24s} n={len(v):5d}  {v[len(v) // 2]:7d} / {sum(v) / len(v):9.1f} / {v[int(0.9 * len(v))]:7d}")
# tile wait per tile
w = defaultdict(list)
for L in launches:
    d = defaultdict(dict)
    for i, ww, j, r, c in L:
        if ww < 8 and i in (0, 1, 10):
            d[(ww, j)][i] = c
    for k, e in d.items():
        if 0 in e and 1 in e:
            w["wait for tile"].append(e[1] - e[0])
        if 1 in e and 10 in e:
            w["tile residence"].append(e[10] - e[1])
for name, v in w.items():
    v = sorted(v)
    print(f"   {name:24s} n={len(v):5d}  {v[len(v) // 2]:7d} / {sum(v) / len(v):9.1f} / {v[int(0.9 * len(v))]:7d}")
