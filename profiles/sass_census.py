#!/usr/bin/env python
"""cuobjdump -sass of the shipped library -> per-kernel counts of the Blackwell-specific instructions (profiles/r2_sass_census.md)."""
import collections
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "..", "ddpo_b200", "libddpo_b200.so")
KEYS = ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS")
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
rows, fn = collections.defaultdict(collections.Counter), None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1)
        continue
    for k in KEYS:
        if re.search(r"\b" + k + r"\b|\b" + k + r"\.", line):
            rows[fn][k] += 1
print("| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---:|" * len(KEYS))
for fn, c in sorted(rows.items()):
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip().split("(")[0]
    print(f"| `{name}` | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
