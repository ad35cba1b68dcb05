#!/usr/bin/env python
"""Dynamic profile of one kernel by *phase*: joins the per-SASS-instruction counters of
`ncu -i X.ncu-rep --page source --csv --print-source sass` with the inline chains of `nvdisasm -g` on the same
object (instruction order is the key) and sums executed warp instructions / stall samples / L1 tag requests between
marker lines (comments containing '=====' or '---- ') of the kernel's source file.
Usage: python tools/ncu_phase_profile.py sass.csv obj.o <kernel-substring> <source.cu>"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main():
    sass_csv, obj, pat, srcfile = sys.argv[1], os.path.abspath(sys.argv[2]), sys.argv[3], sys.argv[4]
    base = os.path.basename(srcfile)
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    lines_of = []            # outermost source line (in srcfile) of every SASS instruction, in order
    inside, cur = False, None
    for l in dis.split("\n"):
        if l.startswith(".text."):
            inside = pat in l
            continue
        if not inside:
            continue
        if "//##" in l:
            t = [int(n) for f, n in re.findall(r'"([^"]+)", line (\d+)', l) if f.endswith(base)]
            if t:
                cur = t[-1]
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines_of.append(cur)
    rows = list(csv.reader(open(sass_csv)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    col = {h: i for i, h in enumerate(rows[hi])}
    body = [r for r in rows[hi + 1:] if len(r) > 5]
    if len(body) != len(lines_of):
        print(f"warning: {len(body)} profiled instructions vs {len(lines_of)} disassembled; object differs from the profiled build?")
    src = open(srcfile).read().split("\n")
    marks = [(i + 1, l.strip()[:72]) for i, l in enumerate(src) if "=====" in l or "// ---- " in l]
    marks = [(1, "(prologue)")] + marks + [(len(src) + 1, "end")]
    STALLS = ["stall_long_sb", "stall_barrier", "stall_wait", "stall_short_sb", "stall_branch_resolving", "stall_mio", "stall_lg",
              "stall_math", "stall_no_inst", "stall_not_selected", "stall_selected"]
    agg = collections.defaultdict(lambda: [0.0] * (4 + len(STALLS)))

    def num(r, k):
        try:
            return float(r[col[k]])
        except (KeyError, ValueError):
            return 0.0
    for r, ln in zip(body, lines_of):
        if ln is None:
            ln = 1
        ph = max(i for i, (a, _) in enumerate(marks) if a <= ln)
        a = agg[ph]
        a[0] += num(r, "Instructions Executed")
        a[1] += num(r, "# Samples")
        a[2] += num(r, "L1 Tag Requests Global")
        a[3] += num(r, "L1 Wavefronts Shared")
        for j, k in enumerate(STALLS):
            a[4 + j] += num(r, k)
    ti, ts = sum(a[0] for a in agg.values()) or 1, sum(a[1] for a in agg.values()) or 1
    print(f"{'phase':72s} {'inst%':>6s} {'smp%':>6s} {'L1 tags':>10s} {'smem wf':>10s}")
    for ph in sorted(agg):
        a = agg[ph]
        print(f"{marks[ph][1]:72s} {100 * a[0] / ti:6.1f} {100 * a[1] / ts:6.1f} {a[2] / 1e6:9.1f}M {a[3] / 1e6:9.1f}M")
    print(f"total warp instructions {ti / 1e6:.1f}M, samples {ts:.0f}")
    print("stall samples per phase (% of all samples): " + " ".join(k.replace("stall_", "") for k in STALLS))
    for ph in sorted(agg):
        a = agg[ph]
        print(f"{marks[ph][1][:40]:40s} " + " ".join(f"{100 * x / ts:6.1f}" for x in a[4:]))
    tot = [sum(agg[ph][4 + j] for ph in agg) for j in range(len(STALLS))]
    print(f"{'total':40s} " + " ".join(f"{100 * x / ts:6.1f}" for x in tot))


if __name__ == "__main__":
    main()
