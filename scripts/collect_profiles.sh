#!/bin/bash
# Turn what scripts/gpu_round.sh left in gpurun_out/ into the tracked evidence files of profiles/.
# usage: bash scripts/collect_profiles.sh r02
set -eu
R=${1:?round tag, e.g. r02}
share() { python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1])); print("%.3f" % d["roofline"]["kernel_share_of_step"])
PY
}
for W in pisces_c4 lobster_c3; do cp gpurun_out/launches_$W.csv profiles/${R}_launches_$W.csv; done
{ echo "# $R — ncu launch list, pisces_c4 (PISCES 1024x1024x128, full size), command:"
  echo "# ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<our kernels> -c 400 python bench.py --workload pisces_c4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
  echo "# (cold-cache, serialised: compare SHARES with bench.py's roofline.kernel_share_of_step = $(share gpurun_out/bench_pisces_c4.json), profiles/${R}_bench_pisces_c4.json)"
  python scripts/ncu_summary.py launches gpurun_out/launches_pisces_c4.csv; } > profiles/${R}_launches_pisces_c4.txt
{ echo "# $R — ncu launch list, lobster_c3 (LOBSTER + carbonates + O2, 512x512x64), same command with --workload lobster_c3"
  echo "# (compare SHARES with roofline.kernel_share_of_step = $(share gpurun_out/bench_lobster_c3.json) in profiles/${R}_bench_lobster_c3.json)"
  python scripts/ncu_summary.py launches gpurun_out/launches_lobster_c3.csv; } > profiles/${R}_launches_lobster_c3.txt
{ echo "# $R — ncu --set full --clock-control none --import-source on, one launch per kernel, pisces_c4 at full size (134 M cells), accumulate mode"
  echo "# (scripts/gpu_profile.sh; summary by scripts/ncu_summary.py full <report>)"
  for K in pisces_tendency scale_negative_calcite par_multiband; do python scripts/ncu_summary.py full gpurun_out/prof_pisces_c4_$K.ncu-rep; done; } > profiles/${R}_full_pisces_c4.txt
{ echo "# $R — ncu --set full --clock-control none --import-source on, one launch per kernel, lobster_c3 at full size (16.8 M cells), accumulate mode"
  for K in npd_tendency par_twoband scale_negative; do python scripts/ncu_summary.py full gpurun_out/prof_lobster_c3_$K.ncu-rep; done; } > profiles/${R}_full_lobster_c3.txt
for W in pisces_c4 lobster_c3 lobster_c2 npzd_c1 carbon_c5; do cp gpurun_out/bench_$W.json profiles/${R}_bench_$W.json; done
cp gpurun_out/bench_ref_pisces_c4.json profiles/${R}_bench_reference_pisces_c4.json
for W in pisces_c4 lobster_c3; do cp gpurun_out/time_kernels_$W.json profiles/${R}_time_kernels_$W.json; done
cp gpurun_out/pcie_bw.json profiles/${R}_pcie_bw.json
cp gpurun_out/stream_pattern.json profiles/${R}_stream_pattern.json
cp gpurun_out/dfma.log profiles/${R}_fp64_peak_dfma.txt
python - "$R" <<'PY'
import json, re, sys
R = sys.argv[1]
def traffic(path, kernel):
    txt = open(path).read()
    blk = txt[txt.index("===== " ) :]
    for part in txt.split("===== ")[1:]:
        if kernel in part.splitlines()[0]:
            rd = float(re.search(r"dram__bytes_read.sum\s+([\d.]+) Gbyte", part).group(1))
            wr = float(re.search(r"dram__bytes_write.sum\s+([\d.]+) Gbyte", part).group(1))
            return (rd + wr) * 1e9
    raise SystemExit(f"{kernel} not in {path}")
t = {"pisces_c4": {"kernel": "pisces_tendency_kernel", "bytes_per_launch": traffic(f"profiles/{R}_full_pisces_c4.txt", "pisces_tendency_kernel"),
                   "note": "captured at full size (134 M cells), accumulate mode", "cells": 134217728, "source": f"profiles/{R}_full_pisces_c4.txt"},
     "lobster_c3": {"kernel": "npd_tendency_kernel", "bytes_per_launch": traffic(f"profiles/{R}_full_lobster_c3.txt", "npd_tendency_kernel"),
                    "note": "full size (16.8 M cells), accumulate mode", "cells": 16777216, "source": f"profiles/{R}_full_lobster_c3.txt"}}
# the other roof: FP64-pipe thread instructions per cell of the dominant PISCES kernel (source page of the same capture) and
# the DFMA rate measured by obm_fp64_peak_dfma_per_s on the same visit
import subprocess
ops = json.loads(subprocess.run([sys.executable, "scripts/ncu_opcodes.py", "gpurun_out/prof_pisces_c4_pisces_tendency.ncu-rep",
                                 str(t["pisces_c4"]["cells"]), "--json"], capture_output=True, text=True, check=True).stdout)
t["pisces_c4"]["fp64_instr_per_cell"] = ops["fp64_instr_per_cell"]
t["pisces_c4"]["fp64_source"] = ("ncu source page of the same capture: " + " + ".join(f"{k} {v}" for k, v in ops["by_opcode"].items())
                                 + f" thread instructions per cell (scripts/ncu_opcodes.py; profiles/{R}_opcodes_pisces_c4.txt)")
rates = [float(x) for x in re.findall(r"DFMA/s\s+([\d.]+)", open(f"profiles/{R}_fp64_peak_dfma.txt").read())]
t["fp64_peak_instr_per_s"] = max(rates) if rates else 16.9e12
t["fp64_peak_source"] = f"profiles/{R}_fp64_peak_dfma.txt (obm_fp64_peak_dfma_per_s: 8 independent DFMA chains per thread, best of three)"
json.dump(t, open("profiles/traffic.json", "w"), indent=1)
PY
{ echo "# $R — executed instructions per cell by SASS opcode (scripts/ncu_opcodes.py on the full captures above; 134 M cells)"
  for K in pisces_tendency scale_negative_calcite par_multiband; do echo "===== $K"; python scripts/ncu_opcodes.py gpurun_out/prof_pisces_c4_$K.ncu-rep 134217728; done; } > profiles/${R}_opcodes_pisces_c4.txt
{ echo "# $R — executed warp instructions per cell by the source function / line they were written in (scripts/ncu_lines.py on the full captures above)"
  for K in scale_negative_calcite par_multiband pisces_tendency; do echo "===== $K"; python scripts/ncu_lines.py gpurun_out/prof_pisces_c4_$K.ncu-rep 134217728 20; done; } > profiles/${R}_functions_pisces_c4.txt
echo "profiles/${R}_* refreshed"
