"""BASELINE configs[1] as a RUN: LOBSTER + carbonates + O₂, 4096 independent columns × 64 levels, sinking POM (UpwindBiased(3)),
ScaleNegativeTracers, RK3 — stepped on the device by the eager loop of `time_step` and by `run(graph=True)` (one captured
time step replayed).  Prints one JSON line: ms per time step either way.
    python scripts/time_column_ensemble.py > gpurun_out/time_column_ensemble.json"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import oceanbiome_b200 as ob  # noqa: E402
from oceanbiome_b200 import synthetic  # noqa: E402


def build(columns=4096, levels=64):
    grid = ob.RectilinearGrid(size=(columns, levels), extent=(float(columns), 200.0), topology=("Periodic", "Flat", "Bounded"), device="cuda")
    bgc = ob.LOBSTER(grid, carbonate_system=ob.CarbonateSystem(), oxygen=ob.Oxygen(), scale_negatives=True,
                     surface_photosynthetically_active_radiation=100.0)
    model = ob.BiogeochemicalModel(grid, bgc, sinking_advection="UpwindBiased3")
    for n, f in model.tracers.items():
        synthetic.fill_torch(f, n, *synthetic.lobster_range(n))
    return model


def main():
    steps, dt, rows = 300, 120.0, {}
    for mode in ("eager", "graph"):
        m = build()
        m.run(dt, 5, graph=mode == "graph")  # lazily built fields, allocator, module loads
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m.run(dt, steps, graph=mode == "graph")
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        rows[mode] = {"ms_per_step_wall": round(wall / steps * 1e3, 4), "cells": m.grid.ncells,
                      "Gcell_updates_per_s_incl_every_launch_of_a_stage": round(m.grid.ncells * 3 * steps / wall / 1e9, 3),
                      "P_mean": float(m.tracers["P"].interior.mean())}
        if mode == "graph":
            rows[mode]["ms_per_step_device"] = round(m.replay_events[0].elapsed_time(m.replay_events[1]) / (steps - 1), 4)
    print(json.dumps({"workload": "LOBSTER + carbonates + O2, 4096 columns x 64 levels, RK3 + sinking, 300 steps", **rows}))


if __name__ == "__main__":
    main()
