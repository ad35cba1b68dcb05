#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (code-size budget; needs -lineinfo).
Usage: python tools/sass_by_line.py obj.o <kernel-substring> [bucket=10]"""
import collections
import os
import re
import subprocess
import sys
import tempfile


def main():
    obj, pat = os.path.abspath(sys.argv[1]), sys.argv[2]
    bucket = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", obj], cwd=td, check=True, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    cnt, cur, inside = collections.Counter(), None, False
    for l in dis.split("\n"):
        if l.startswith(".text."):
            inside = pat in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l) and cur:
            cnt[cur] += 1
    print("total", sum(cnt.values()))
    b = collections.Counter()
    for (f, ln), c in cnt.items():
        b[(f, ln // bucket * bucket)] += c
    for k, c in sorted(b.items(), key=lambda kv: -kv[1])[:45]:
        print(f"{k[0]:>26s}:{k[1]:<5d} {c}")


if __name__ == "__main__":
    main()
