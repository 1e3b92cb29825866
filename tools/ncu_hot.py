#!/usr/bin/env python
"""Per-opcode and per-source-line stall sample breakdown from an .ncu-rep (needs --import-source on, -lineinfo)."""
import csv
import io
import subprocess
import sys
import collections

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[hi], rows[hi + 1:]
ix = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[ix["# Samples"]] or 0) for r in data) or 1
ops = collections.Counter()
inst = collections.Counter()
for r in data:
    src = r[ix["Source"]]
    if not src:
        continue
    t = src.split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[ix["# Samples"]] or 0)
    inst[op] += int(r[ix["Instructions Executed"]] or 0)
ti = sum(inst.values()) or 1
print(f"total samples {tot}, warp instructions {ti}")
for op, n in ops.most_common(16):
    print(f"  {op:10s} samples {100 * n / tot:6.2f}%   executed {100 * inst[op] / ti:6.2f}%")
# address ranges with the most samples (64-instruction windows)
win = collections.Counter()
for r in data:
    try:
        a = int(r[ix["Address"]], 16)
    except ValueError:
        continue
    win[a // 1024] += int(r[ix["# Samples"]] or 0)
print("hottest 1 KiB code windows (64 instructions each):")
for w, n in win.most_common(8):
    print(f"  0x{w * 1024:06x}  {100 * n / tot:6.2f}%")
