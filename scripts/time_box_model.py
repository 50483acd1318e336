"""The reference's box-model benchmark (benchmark/box_model.jl:22-69: NPZD, N = 10, P = 0.1, Z = 0.01, prescribed PAR,
1000 RK3 steps of 20 minutes; published 23.5 ms without outputs, 34 ms with `SpeedyOutput` every 20 steps, for ONE box
on a CPU) run as a device-resident ensemble of n boxes: wall time of `run(...)` (tabulation of the PAR series in Python,
upload, capture and 1000 replays — or ONE launch for the whole run, `device_loop=True` — snapshots every 20 steps kept on
the device) and the device time of the 1000 steps alone.  Run on the GPU box:
    python scripts/time_box_model.py > gpurun_out/time_box_model.json"""
import json
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import oceanbiome_b200 as ob  # noqa: E402

day, minutes = 86400.0, 60.0
year = 365 * day


def PAR_func(t):
    PAR0 = 60 * (1 - math.cos((t + 15 * day) * 2 * math.pi / year)) \
        * (1 / (1 + 0.2 * math.exp(-(((t % year) - 200 * day) / (50 * day)) ** 2))) + 2
    return PAR0 * math.exp(0.2 * -10)


def build(n, sweep, fused=False):
    grid = ob.BoxModelGrid(n, device="cuda")
    PAR = ob.CenterField(grid, "PAR")
    kw = {}
    if sweep:
        rng = np.random.default_rng(0)
        kw["parameter_ensemble"] = {
            "phytoplankton_maximum_growth_rate": 0.6989 / day * rng.uniform(0.7, 1.3, n),
            "nitrate_half_saturation": 2.3868 * rng.uniform(0.7, 1.3, n),
            "light_half_saturation": 3.58 * rng.uniform(0.7, 1.3, n),
            "phytoplankton_mortality_rate": 0.0761 / day * rng.uniform(0.7, 1.3, n),
            "phytoplankton_solid_waste_fraction": 0.1327 * rng.uniform(0.7, 1.3, n)}
    bgc = ob.NPZD(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR), **kw)
    model = ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_func}, fused_step=fused)  # T stays 0, as in the benchmark
    model.set(N=10.0, P=0.1, Z=0.01)
    return model


def main():
    steps, rows = 1000, []
    build(1, False).run(20 * minutes, 20, graph=True)  # module loads, allocator
    build(1, False, True).run(20 * minutes, 20, graph=True)
    build(1, False, True).run(20 * minutes, 20, device_loop=True)
    cases = []
    for n, sweep in ((1, False), (8, True), (4096, True), (262144, True), (1048576, True)):
        cases += [(n, sweep, False, "graph"), (n, sweep, True, "graph"), (n, sweep, True, "device_loop")]
    for n, sweep, fused, mode in cases:
        model = build(n, sweep, fused)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = model.run(20 * minutes, steps, graph=mode == "graph", device_loop=mode == "device_loop", output_every=20)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        replay = model.replay_events[0].elapsed_time(model.replay_events[1]) * 1e-3  # the 1000 steps alone, on the device
        P = out["P"]
        rows.append({"boxes": n, "parameter_sweep": sweep, "fused_tendency_and_substep": fused,
                     "mode": "one launch for the whole run" if mode == "device_loop" else "CUDA graph of one step, replayed",
                     "steps": steps, "run_wall_s": round(wall, 4),
                     "device_s": round(replay, 5), "us_per_rk3_stage": round(replay / steps / 3 * 1e6, 2),
                     "box_steps_per_s": round(n * steps / wall, 1),
                     "reference_one_box_s": 0.0235, "speedup_vs_reference_sequential": round(0.0235 * n / wall, 1),
                     "speedup_device_time_vs_reference_sequential": round(0.0235 * n / replay, 1),
                     "finite": bool(torch.isfinite(P).all()), "P_end_member0": float(P[-1, 0])})
        del model, out
    print(json.dumps({"benchmark": "benchmark/box_model.jl: NPZD box, 1000 RK3 steps of 20 min, snapshots every 20 steps",
                      "reference_published": "23.5 ms no outputs / 34 ms SpeedyOutput, one box, unstated CPU",
                      "rows": rows}, indent=1))


if __name__ == "__main__":
    main()
