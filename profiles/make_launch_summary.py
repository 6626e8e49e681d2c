#!/usr/bin/env python
"""ncu launch list (CSV) -> per-kernel markdown table + profiles/r1_traffic.json.

Capture (one GPU, profiler brackets exactly one eager step; per-launch times are cold-cache and serialised, so compare
SHARES with bench.py's CUDA-event breakdown, not absolutes):

    ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/launches_sample.csv \
        --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        python bench.py --ncu sample --steps 1 --warmup 3 --no-cpu --phase sample
    (same with `--ncu train --phase ppo` -> launches_train.csv)

    python profiles/make_launch_summary.py sample=gpurun_out/launches_sample.csv train=gpurun_out/launches_train.csv
"""
import collections
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCALE = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
GROUP = [("igemm", "igemm"), ("wgrad2_kernel", "wgrad"), ("wgrad_kernel", "wgrad"), ("wgrad_reduce", "wgrad_reduce"),
         ("attention_fwd", "attention_fwd"), ("attention_bwd", "attention_bwd"), ("attention_delta", "attention_bwd")]


def short(name):
    n = name.split("(")[0].replace("ddpo::", "").replace("void ", "")
    for key, g in GROUP:
        if key in n:
            return g
    return n.split("<")[0]


def parse(path):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    r = csv.DictReader(lines)
    per = collections.defaultdict(dict)
    for row in r:
        key = row["ID"]
        per[key]["name"] = row["Kernel Name"]
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        per[key][row["Metric Name"]] = v * SCALE.get(row["Metric Unit"], 1.0)
    for key in sorted(per, key=lambda k: int(k)):
        d = per[key]
        rows.append((short(d["name"]), d.get("gpu__time_duration.sum", 0.0),
                     d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)))
    return rows


def main():
    out = {}
    prefix, traffic = "r2_launches_summary", "r2_launches_summary"
    if len(sys.argv) > 2 and sys.argv[1] == "--out":        # --out NAME: write NAME.md / NAME.json instead
        prefix = traffic = sys.argv[2]
        del sys.argv[1:3]
    md = [f"# {prefix}: ncu launch lists of one eager step (current build)\n",
          "Per-launch times under ncu are cold-cache and serialised: the SHARE column is what must agree with the CUDA-event",
          "breakdown `bench.py` prints (`kernels` / `ppo.kernels`).  DRAM bytes are `dram__bytes_read.sum + dram__bytes_write.sum`.\n"]
    for arg in sys.argv[1:]:
        tag, path = arg.split("=", 1)
        rows = parse(path)
        agg = collections.OrderedDict()
        for name, ms, b in rows:
            a = agg.setdefault(name, [0, 0.0, 0.0])
            a[0] += 1
            a[1] += ms
            a[2] += b
        tot = sum(a[1] for a in agg.values())
        md.append(f"\n## {tag}: {len(rows)} launches, {tot:.2f} ms summed\n")
        md.append("| kernel | launches | total ms | share | DRAM MB / launch | DRAM GB/s |")
        md.append("|---|---:|---:|---:|---:|---:|")
        out[tag] = {}
        for name, (n, ms, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            md.append(f"| {name} | {n} | {ms:.3f} | {100 * ms / tot:.1f}% | {b / n / 1e6:.1f} | {b / ms / 1e6 if ms else 0:.0f} |")
            out[tag][name] = {"launches": n, "ms": round(ms, 4), "share": round(ms / tot, 4),
                              "dram_bytes_per_launch": round(b / n), "dram_gbs": round(b / ms / 1e6, 1) if ms else None}
    with open(os.path.join(HERE, prefix + ".md"), "w") as f:
        f.write("\n".join(md) + "\n")
    with open(os.path.join(HERE, traffic + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    print("\n".join(md))


if __name__ == "__main__":
    main()
