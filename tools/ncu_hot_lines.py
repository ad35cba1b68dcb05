#!/usr/bin/env python
"""Per-source-line summary of `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`:
samples (all / not issued), executed instructions and the dominant stall reasons of the hottest lines.
Usage: python tools/ncu_hot_lines.py src.csv [top=40] [file-substring]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    only = sys.argv[3] if len(sys.argv) > 3 else ""
    rows = list(csv.reader(open(path)))
    cur_file, head, lines = "", None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Line No":
            head = {h: i for i, h in enumerate(r)}
            continue
        if head is None or r[0] in ("", "Function Name", "Kernel Name") or not r[0].isdigit():
            continue
        lines.append((cur_file, head, r))
    def num(r, head, k):
        try:
            return float(r[head[k]])
        except (KeyError, ValueError, IndexError):
            return 0.0
    tot_s = sum(num(r, h, "# Samples") for _, h, r in lines) or 1.0
    tot_i = sum(num(r, h, "Instructions Executed") for _, h, r in lines) or 1.0
    print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
    stall_keys = [k for k in lines[0][1] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {}
    for f, h, r in lines:
        for k in stall_keys:
            agg[k] = agg.get(k, 0.0) + num(r, h, k)
    tot_st = sum(agg.values()) or 1.0
    print("stall mix: " + ", ".join(f"{k[6:]} {100 * v / tot_st:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    sel = [x for x in lines if only in x[0]]
    sel.sort(key=lambda x: -num(x[2], x[1], "# Samples"))
    for f, h, r in sel[:top]:
        st = sorted(((num(r, h, k), k[6:]) for k in stall_keys), reverse=True)[:2]
        print(f"{f.split('/')[-1]:>24s}:{r[0]:>4s} smp {100 * num(r, h, '# Samples') / tot_s:5.1f}% inst {100 * num(r, h, 'Instructions Executed') / tot_i:5.1f}% "
              f"{st[0][1]}/{st[1][1]:<14s} {r[1].strip()[:110]}")


if __name__ == "__main__":
    main()
