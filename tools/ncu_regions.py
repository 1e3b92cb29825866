#!/usr/bin/env python
"""Per-region view of an `ncu --page source --csv` dump: contiguous SASS ranges grouped by how many
threads execute them, with warp instructions per unit of work.

    python tools/ncu_regions.py source.csv <units (e.g. warp-paths)> [--dump lo hi]
"""
import csv
import itertools
import sys

rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2])
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
ix = {h: i for i, h in enumerate(rows[hi])}
data = rows[hi + 1:]
out = []
for n, r in enumerate(data):
    ie = int(r[ix["Instructions Executed"]] or 0)
    at = float(r[ix["Avg. Threads Executed"]] or 0)
    sm = int(r[ix["# Samples"]] or 0)
    out.append((n, r[ix["Source"]], ie / units, at, sm))
tot_s = sum(o[4] for o in out) or 1
print(f"warp instructions per unit: {sum(o[2] for o in out):.1f}")
if "--dump" in sys.argv:
    a, b = int(sys.argv[sys.argv.index("--dump") + 1]), int(sys.argv[sys.argv.index("--dump") + 2])
    for o in out[a:b]:
        print(f"{o[0]:5d} {o[2]:8.2f} {o[3]:5.1f} {o[4]:6d}  {o[1][:90]}")
    sys.exit(0)


def cls(o):
    return -1 if o[2] == 0 else int(o[2] * 4 + 0.5) * 100 + int(o[3] / 4)


for k, g in itertools.groupby(out, key=cls):
    g = list(g)
    if k < 0 and len(g) < 8:
        continue
    ins = sum(x[2] for x in g)
    print(f"rows {g[0][0]:5d}-{g[-1][0]:5d} n={len(g):4d}  exec/unit {g[0][2]:7.2f}  instr/unit {ins:8.1f}  "
          f"samples {100 * sum(x[4] for x in g) / tot_s:5.1f}%  avg threads {sum(x[3] * x[2] for x in g) / max(1e-9, ins):5.1f}")
