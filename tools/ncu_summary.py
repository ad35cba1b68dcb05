#!/usr/bin/env python
"""Condense an `ncu -i X.ncu-rep --page raw --csv` dump into a small JSON (one record per profiled launch).
Usage: python tools/ncu_summary.py raw.csv out.json [kernel-substring]"""
import csv
import json
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
]


def main():
    src, dst = sys.argv[1], sys.argv[2]
    pat = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(src)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in body:
        name = r[col["Kernel Name"]]
        if pat and pat not in name:
            continue
        rec = {"kernel": name[:90]}
        for k in KEYS:
            if k in col and r[col[k]] != "":
                rec[k] = (r[col[k]] + " " + units[col[k]]).strip()
        # warp-stall breakdown (top 6)
        stalls = []
        for h, i in col.items():
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") and r[i]:
                try:
                    stalls.append((float(r[i]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        rec["top_stalls_per_issue"] = {n: round(v, 3) for v, n in stalls[:6]}
        out.append(rec)
    json.dump(out, open(dst, "w"), indent=1)
    print(f"{len(out)} launches -> {dst}")


if __name__ == "__main__":
    main()
