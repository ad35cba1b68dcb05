#!/usr/bin/env python
"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list (one NVTX-filtered timed bench step or more)
into per-kernel totals and shares.  Usage: python tools/launch_list.py launches.csv [steps] [out.json]"""
import collections, csv, json, sys

rows = list(csv.reader(open(sys.argv[1])))
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
col = {h: i for i, h in enumerate(rows[hi])}
agg = collections.OrderedDict()
tot = 0.0
for r in rows[hi + 1:]:
    if len(r) < len(col) or r[col["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[col["Metric Value"]].replace(",", ""))
    u = r[col["Metric Unit"]]
    v = v / 1e6 if u == "ns" else v / 1e3 if u == "us" else v
    k = r[col["Kernel Name"]][:120]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
out = [{"kernel": k, "launches_per_step": n / steps, "ms_per_step": ms / steps, "share_pct": 100 * ms / tot} for k, (n, ms) in
       sorted(agg.items(), key=lambda kv: -kv[1][1])]
print(f"{tot / steps:.3f} ms per step over {sum(a[0] for a in agg.values()) / steps:.0f} launches (serialised, cold cache)")
for o in out[:30]:
    print(f"{o['ms_per_step']:8.3f} ms {o['share_pct']:5.1f}% {o['launches_per_step']:5.1f}x  {o['kernel'][:100]}")
if len(sys.argv) > 3:
    json.dump({"ms_per_step": tot / steps, "kernels": out}, open(sys.argv[3], "w"), indent=1)
