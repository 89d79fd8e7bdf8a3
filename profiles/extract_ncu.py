#!/usr/bin/env python
"""Per-kernel averages of an `ncu --set full` report: duration, DRAM bytes, instructions, issue-slot utilisation.

    python profiles/extract_ncu.py gpurun_out/r1q_chain.ncu-rep > profiles/r01_chain_ncu.json
"""
import collections
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "dur",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "smsp__inst_executed.sum": "warp_inst",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "lts__t_bytes.sum": "l2_bytes",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows[2:]:
        name = r[idx["Kernel Name"]].split("(")[0]
        for m, key in WANT.items():
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            agg[name][key].append(v * SCALE.get(units[idx[m]], 1.0))
    out = {}
    for name, d in agg.items():
        out[name] = {k: sum(v) / len(v) for k, v in d.items()}
        out[name]["launches_profiled"] = len(d["dur"])
    print(json.dumps({"source": path, "units": {"dur": "us", "dram_read": "bytes", "dram_write": "bytes", "l2_bytes": "bytes"}, "kernels": out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
