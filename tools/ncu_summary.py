#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box) into the tracked files under profiles/:
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_v7_bf16_ncu_full [key-prefix]
writes <out>.md (per-kernel table) and <out>.json, and refreshes profiles/ncu_traffic.json
(kernel name -> DRAM bytes per launch at the default bench workload; bench.py's roofline.traffic).
key-prefix "5/tf32x3/" files the entries under "<workload>/<precision>/<kernel>" (what bench.py looks up first)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read_MB"),
    ("dram__bytes_write.sum", "dram_write_MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__inst_executed_pipe_xu.sum", "xu_inst"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]


def to_float(v, unit):
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return None
    u = unit.lower()
    scale = {"gbyte": 1e3, "mbyte": 1.0, "kbyte": 1e-3, "byte": 1e-6}.get(u)
    if scale is not None:
        return x * scale
    if u in ("ms", "msecond"):
        return x * 1e3
    if u in ("ns", "nsecond"):
        return x * 1e-3
    if u in ("s", "second"):
        return x * 1e6
    return x


def main(rep, out, prefix=""):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[idx["Kernel Name"]]).replace("void ", "").strip()
        name = re.sub(r"<unnamed>::|unnamed>::|regt::|\(anonymous namespace\)::", "", name)
        d = {"kernel": name}
        for m, k in METRICS:
            if m in idx:
                d[k] = to_float(r[idx[m]], units[idx[m]])
        res.append(d)
    with open(out + ".json", "w") as f:
        json.dump(res, f, indent=1)
    cols = ["kernel"] + [k for _, k in METRICS]
    with open(out + ".md", "w") as f:
        f.write(f"ncu --set full --clock-control none, report {os.path.basename(rep)} (cold-cache, serialised replays: compare shares)\n\n")
        f.write("| " + " | ".join(cols) + " |\n|" + "---|" * len(cols) + "\n")
        for d in res:
            f.write("| " + " | ".join(("%.4g" % d[c]) if isinstance(d.get(c), float) else str(d.get(c, "")) for c in cols) + " |\n")
    traffic_path = os.path.join(os.path.dirname(os.path.abspath(out)), "ncu_traffic.json")
    traffic = {}
    if os.path.exists(traffic_path):
        traffic = json.load(open(traffic_path))
    for d in res:
        if d.get("dram_read_MB") is not None:
            key = re.sub(r"<.*", "", d["kernel"])
            traffic[prefix + key] = {"bytes_per_launch": int((d["dram_read_MB"] + d["dram_write_MB"]) * 1e6), "source": os.path.basename(out)}
    json.dump(traffic, open(traffic_path, "w"), indent=1)
    print(open(out + ".md").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")
