#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of bench.py: per kernel, the per-launch DRAM traffic and a few headline
counters.  Reads the raw-page CSV (`ncu -i X.ncu-rep --page raw --csv > X.csv`), writes profiles/traffic.json
(read by bench.py for `roofline.traffic`) and prints a markdown table.

    python tools/ncu_traffic.py gpurun_out/X_full_raw.csv [profiles/traffic.json]
"""
import csv
import json
import re
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main():
    src = sys.argv[1]
    dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/traffic.json"
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in data:
        name = re.sub(r"<.*", "", r[col["Kernel Name"]]).split("::")[-1].replace("void ", "").strip()
        rec = {}
        for k in KEEP:
            if k in col and r[col[k]] != "":
                rec[k] = float(r[col[k]].replace(",", "")) * UNIT.get(units[col[k]], 1.0)
        per.setdefault(name, []).append(rec)
    out = {}
    for name, recs in per.items():
        n = len(recs)
        avg = {k: sum(x.get(k, 0.0) for x in recs) / n for k in recs[0]}
        out[name] = {"launches_captured": n,
                     "dram_bytes_per_launch": avg.get("dram__bytes_read.sum", 0) + avg.get("dram__bytes_write.sum", 0),
                     "dram_read_bytes": avg.get("dram__bytes_read.sum", 0), "dram_write_bytes": avg.get("dram__bytes_write.sum", 0),
                     "duration_s_under_ncu": avg.get("gpu__time_duration.sum", 0),
                     "warp_instructions": avg.get("smsp__inst_executed.sum", 0),
                     "warps_active_pct": avg.get("sm__warps_active.avg.pct_of_peak_sustained_active"),
                     "issue_active_pct": avg.get("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                     "registers_per_thread": avg.get("launch__registers_per_thread"), "grid": avg.get("launch__grid_size"),
                     "smem_bank_conflicts": avg.get("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                     "source": src}
    json.dump(out, open(dst, "w"), indent=1)
    print("| kernel | launches | time (us, under ncu) | DRAM read | DRAM write | GB/s | warp instr | issue active % | warps active % |")
    print("|---|---|---|---|---|---|---|---|---|")
    for name, o in out.items():
        t = o["duration_s_under_ncu"]
        print(f"| {name} | {o['launches_captured']} | {t * 1e6:.1f} | {o['dram_read_bytes'] / 1e6:.1f} MB | "
              f"{o['dram_write_bytes'] / 1e6:.1f} MB | {o['dram_bytes_per_launch'] / t / 1e9 if t else 0:.0f} | "
              f"{o['warp_instructions'] / 1e6:.1f} M | {o['issue_active_pct']:.1f} | {o['warps_active_pct']:.1f} |")


if __name__ == "__main__":
    main()
