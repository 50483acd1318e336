#!/usr/bin/env python
"""Executed warp instructions per cell by SOURCE FUNCTION (and the hottest source lines) of one captured kernel, from the
cuda,sass source page of an `ncu --set full --import-source on` report (inlined code is attributed to the line it was
written on).  usage: ncu_lines.py <report.ncu-rep> <cells> [top-lines]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def enclosing_function(path, line, cache={}):
    local = os.path.join(ROOT, path.split("/repo/", 1)[1]) if "/repo/" in path else path
    if local not in cache:
        try:
            cache[local] = open(local, encoding="utf-8").read().split("\n")
        except OSError:
            cache[local] = []
    src = cache[local]
    for i in range(min(line - 1, len(src) - 1), -1, -1):
        m = re.match(r"^(?:template\s*<[^>]*>\s*)?(?:static\s+|inline\s+|extern\s+\"C\"\s+)*(?:__device__|__global__|__host__)[^;{]*?\b(\w+)\s*\(", src[i])
        if m and not src[i].startswith(" "):
            return m.group(1)
        m = re.match(r"^\s{4}static __device__ __forceinline__ \S+ (\w+)\(", src[i])  # members of the arithmetic-policy structs
        if m:
            return m.group(1)
    return "?"


def main():
    report, cells = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    txt = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True, check=True).stdout
    path, hdr = None, None
    by_line = collections.Counter()
    text = {}
    for r in csv.reader(txt.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1]
        elif r and r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
        elif hdr and r and r[0].isdigit() and len(r) > ie and r[ie].isdigit():
            by_line[(path, int(r[0]))] += int(r[ie])
            text[(path, int(r[0]))] = r[1].strip()
    warps = cells / 32
    by_fn = collections.Counter()
    for (p, l), n in by_line.items():
        by_fn[(os.path.basename(p), enclosing_function(p, l))] += n
    total = sum(by_line.values())
    print(f"# executed warp instructions per cell: {total / warps:.1f} (by the function the source line belongs to)")
    for (f, fn), n in by_fn.most_common(30):
        print(f"{n / warps:9.1f}  {fn}  ({f})")
    print(f"# hottest {top} source lines")
    for (p, l), n in by_line.most_common(top):
        print(f"{n / warps:9.1f}  {os.path.basename(p)}:{l}  {text[(p, l)][:110]}")


if __name__ == "__main__":
    main()
