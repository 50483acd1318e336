"""LOBSTER in a box — the reference's examples/box.jl: one well-mixed box at a nominal depth of 10 m under a seasonal PAR
cycle, 5 years at Δt = 5 minutes, every field saved every 10 days.

The reference steps the box on the CPU; here the box (or `--boxes n` of them side by side) lives on the device and one
captured time step is replayed, so the 525 600 steps cost launch latency only.  The time series is written as `box.npz`
(the reference writes JLD2) and, if matplotlib is installed, plotted like the reference's figure.

    python examples/box.py [--years 5] [--boxes 1] [--out box.npz]
"""
import argparse
import math
import sys

import numpy as np

sys.path.insert(0, ".")
import oceanbiome_b200 as ob  # noqa: E402

minutes, day = 60.0, 86400.0
year = 365 * day
z = -10.0  # nominal depth of the box for the PAR profile


def PAR0(t):  # box.jl:24
    return 60 * (1 - math.cos((t + 15 * day) * 2 * math.pi / year)) \
        * (1 / (1 + 0.2 * math.exp(-(((t % year) - 200 * day) / (50 * day)) ** 2))) + 2


def PAR_func(t):
    return PAR0(t) * math.exp(0.2 * z)


def build(boxes=1, device="cuda"):
    grid = ob.BoxModelGrid(boxes, device=device)
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.LOBSTER(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR))
    model = ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_func}, fused_step=True)
    model.set(**{"NO₃": 10.0, "NH₄": 0.1, "P": 0.1, "Z": 0.01})
    return model


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--years", type=float, default=5.0)
    ap.add_argument("--boxes", type=int, default=1)
    ap.add_argument("--out", default="box.npz")
    args = ap.parse_args()
    model = build(args.boxes)
    dt = 5 * minutes
    every = int(round(10 * day / dt))  # TimeInterval(10days)
    steps = int(round(args.years * year / dt))
    series = model.run(dt, steps, device_loop=True, output_every=every)  # the whole run is one launch (obm_npd_box_run)
    times = (np.arange(steps // every) + 1) * every * dt
    np.savez(args.out, t=times, **{n: v.cpu().numpy() for n, v in series.items()})
    print(f"{steps} steps of {args.boxes} box(es): {args.out} holds {len(times)} outputs of {', '.join(series)}")
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
    except ImportError:
        return
    fig, axs = plt.subplots((len(series) + 1) // 2, 2, figsize=(12, 12), squeeze=False)
    for ax, (name, v) in zip(axs.ravel(), series.items()):
        ax.plot(times / year, v[:, 0].cpu().numpy(), linewidth=3)
        ax.set_xlabel("Year")
        ax.set_ylabel(name)
    fig.savefig("box.png")


if __name__ == "__main__":
    main()
