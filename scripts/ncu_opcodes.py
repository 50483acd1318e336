#!/usr/bin/env python
"""Dynamic opcode histogram of one captured kernel: executed (thread-level, predicated-on) instructions per cell by SASS
opcode, from the source page of an `ncu --set full --import-source on` report.
usage: ncu_opcodes.py <report.ncu-rep> <cells> [--json]   (run where the report is; needs the ncu CLI)"""
import collections
import csv
import json
import subprocess
import sys


def histogram(report, cells):
    txt = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(r for r in rows if "Source" in r and "Instructions Executed" in r)
    si, wi, ti = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Predicated-On Thread Instructions Executed")
    warp, thread = collections.Counter(), collections.Counter()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= ti or not r[wi].isdigit():
            continue
        s = r[si].strip()
        if s.startswith("@"):
            s = s.split(None, 1)[1]
        op = s.split()[0]
        warp[op] += int(r[wi])
        thread[op] += int(r[ti])
    return ({k: v * 32 / cells for k, v in warp.items()}, {k: v / cells for k, v in thread.items()})


if __name__ == "__main__":
    report, cells = sys.argv[1], int(sys.argv[2])
    warp, thread = histogram(report, cells)
    fp64 = {k: v for k, v in thread.items() if k.split(".")[0] in ("DFMA", "DMUL", "DADD", "DSETP")}
    if "--json" in sys.argv:
        by = collections.Counter()
        for k, v in fp64.items():
            by[k.split(".")[0]] += v
        print(json.dumps({"fp64_instr_per_cell": round(sum(fp64.values()), 1), "by_opcode": {k: round(v, 1) for k, v in by.items()},
                          "warp_instr_per_cell_slot": round(sum(warp.values()), 1)}))
    else:
        print(f"# executed warp instructions per cell (x32 lanes / cells): {sum(warp.values()):.1f}; FP64-pipe thread instructions per cell: {sum(fp64.values()):.1f}")
        for k, v in sorted(warp.items(), key=lambda kv: -kv[1])[:40]:
            print(f"{k:28s} {v:9.1f}")
