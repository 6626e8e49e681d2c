#!/usr/bin/env python
"""Turn an `ncu --set full` report into the per-kernel table quoted in profiles/README.md.

    python profiles/summarize_ncu.py profiles/r1_igemm_full.ncu-rep [more.ncu-rep ...] > table.md

Reads the report with `ncu -i <rep> --page raw --csv` (ncu is in the image; no GPU needed) and prints, per launch:
duration, DRAM read+write bytes (the `roofline.traffic` figure of bench.py), DRAM throughput as % of the measured 6576 GB/s copy peak (MEASURED_PEAKS.json), tensor-pipe active %,
XU (MUFU) pipe %, SM throughput %, achieved occupancy, registers/thread, grid x block.
"""
import csv
import io
import subprocess
import sys

COLS = {
    "name": "Kernel Name",
    "grid": "Grid Size",
    "block": "Block Size",
    "dur": "gpu__time_duration.sum",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
    "sm": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "occ": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
}
UNIT_SCALE = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def num(x):
    try:
        return float(str(x).replace(",", ""))
    except ValueError:
        return float("nan")


def rows(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    r = list(csv.reader(io.StringIO(txt)))
    head, units, body = r[0], r[1], r[2:]
    idx = {}
    for k, v in COLS.items():  # some metrics carry a section prefix ("FBSP.TriageCompute.dram__throughput...")
        hit = [i for i, h in enumerate(head) if h == v] or [i for i, h in enumerate(head) if h.endswith("." + v)]
        if hit:
            idx[k] = hit[0]
    for b in body:
        d = {k: b[i] for k, i in idx.items()}
        u = {k: units[i] for k, i in idx.items()}
        d["dur_us"] = num(d["dur"]) * UNIT_SCALE.get(u["dur"], 1.0)
        d["bytes"] = num(d["rd"]) * UNIT_SCALE.get(u["rd"], 1.0) + num(d["wr"]) * UNIT_SCALE.get(u["wr"], 1.0)
        yield d


def main():
    for rep in sys.argv[1:]:
        print(f"\n### {rep}\n")
        print("| kernel | grid x block | regs | us | DRAM MB | GB/s | % of 6576 GB/s | tensor % | XU % | SM % | warps active % |")
        print("|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
        for d in rows(rep):
            name = d["name"].split("(")[0][:44]
            gbs = d["bytes"] / d["dur_us"] / 1e3 if d["dur_us"] > 0 else float("nan")
            f = lambda k: f"{num(d.get(k, 'nan')):.1f}"
            print(f"| {name} | {d['grid']} x {d['block']} | {d.get('regs', '')} | {d['dur_us']:.1f} | "
                  f"{d['bytes'] / 1e6:.1f} | {gbs:.0f} | {100 * gbs / 6576:.1f} | {f('tensor')} | {f('xu')} | {f('sm')} | {f('occ')} |")


if __name__ == "__main__":
    main()
