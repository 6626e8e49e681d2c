#!/usr/bin/env python
"""Top warp-stall SASS lines of one launch in an ncu report (needs --import-source on / -lineinfo).
    python profiles/ncu_hotspots.py <rep> <launch index> [top N]
"""
import csv, io, subprocess, sys

rep, idx = sys.argv[1], int(sys.argv[2])
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(txt)))
hi = [i for i, x in enumerate(r) if x and x[0] == "Address"][0]
print(r[hi - 1][:2] if hi else "")
h, rows = r[hi], r[hi + 1:]
si, src = h.index("# Samples"), h.index("Source")
ex = h.index("Instructions Executed")
tot = sum(int(x[si]) for x in rows if len(x) > si and x[si].isdigit())
print("total samples", tot, " SASS lines", len(rows))
best = sorted([(int(x[si]), i, x[src], x[ex]) for i, x in enumerate(rows) if len(x) > si and x[si].isdigit()], reverse=True)
for s, i, t, e in best[:top_n]:
    print(f"{s:6d} {100 * s / tot:5.1f}%  line{i:5d} exec={e:>8s}  {t.strip()[:100]}")
