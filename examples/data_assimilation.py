"""Calibrating NPZD parameters in a box model with an ensemble Kalman inversion — the workflow of the reference's
examples/data_assimilation.jl, with the whole ensemble stepped in ONE device-resident box model.

The reference builds one `BoxModel` per parameter vector (`run_box_simulation`, :26-62) and runs the N_ensemble = 8
members of each of the 5 iterations on CPU threads (:118-130).  Here the members are the boxes of `BoxModelGrid(n)`:
`NPZD(grid, parameter_ensemble={…})` gives every box its own PhytoZoo parameters, and `run(..., device_loop=True)` integrates every
member through the whole run in ONE launch (`obm_npd_box_run`; `graph=True` would replay one captured time step for all of
them).  The Kalman update itself (EnsembleKalmanProcesses.jl's `Inversion()`, a
third-party package) is the textbook perturbed-observation update, written out in NumPy below.

    python examples/data_assimilation.py            # 2 model years per forward run, as in the reference
"""
import math
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import oceanbiome_b200 as ob  # noqa: E402

day, minutes, hours = 86400.0, 60.0, 3600.0
year = 365 * day
z = -10.0  # nominal depth of the box for the PAR profile


def PAR_func(t):  # data_assimilation.jl:22-25
    PAR0 = 60 * (1 - math.cos((t + 15 * day) * 2 * math.pi / year)) \
        * (1 / (1 + 0.2 * math.exp(-(((t % year) - 200 * day) / (50 * day)) ** 2))) + 2
    return PAR0 * math.exp(0.2 * z)


def phytozoo_parameters(u):
    """(α, μ₀, k_N, m_P), one column per member → the PhytoZoo keywords `run_box_simulation` sets (:39-43), written as
    they are there (m_P, a rate per second, is divided by `day` once more in the mortality rate)."""
    alpha, mu, kN, mP = u
    return {"phytoplankton_maximum_growth_rate": mu, "nitrate_half_saturation": kN, "light_half_saturation": mu / alpha,
            "phytoplankton_mortality_rate": 0.066 / day + mP / day,
            "phytoplankton_solid_waste_fraction": mP * day / (0.066 + mP * day)}


def run_box_simulations(u, stop_time=2 * year, dt=20 * minutes, output_interval=8 * hours, device="cuda"):
    """All members at once: P of the last model year, shape (n_times, n_members), and the output times."""
    u = np.atleast_2d(np.asarray(u, dtype=np.float64).T).T  # (4, n)
    n = u.shape[1]
    grid = ob.BoxModelGrid(n, device=device)
    PAR = ob.CenterField(grid, "PAR")
    bgc = ob.NPZD(grid, light_attenuation=ob.PrescribedPhotosyntheticallyActiveRadiation(PAR),
                  parameter_ensemble=phytozoo_parameters(u))
    model = ob.BoxModel(biogeochemistry=bgc, grid=grid, prescribed_tracers={"PAR": PAR_func}, fused_step=True)
    model.set(N=10.0, P=0.1, Z=0.01)
    every = int(round(output_interval / dt))
    steps = int(round(stop_time / dt))
    out = model.run(dt, steps, device_loop=True, output_every=every, output_names=["P"])  # every member, every step: one launch
    times = (np.arange(out["P"].shape[0]) + 1) * every * dt
    last = min(len(times), int(round(year / output_interval)) - 2)  # the reference keeps the last 1093 outputs
    return out["P"][-last:], times[-last:]


def extract_observables(P, times):
    """(peak, winter, average, peak_timing, die_off_time) per member — data_assimilation.jl:83-100; NaN for a member
    whose P is not positive throughout."""
    t = torch.as_tensor(times, device=P.device)
    ok = (P > 0).all(dim=0)
    h = P.shape[0] // 2 - 1  # second half of the year (diff(P)[546:end] of 1093 outputs, 1-based)
    growth = P[1:] - P[:-1]
    obs = torch.stack([P.max(dim=0).values, P.min(dim=0).values, P.mean(dim=0), t[P.argmax(dim=0)] / day,
                       t[h + growth[h:].argmin(dim=0)] / day])
    obs[:, ~ok] = float("nan")
    return obs.cpu().numpy()


def G(u, **kw):
    P, times = run_box_simulations(u, **kw)
    return extract_observables(P, times), P


def constrained_gaussian(rng, mean, std, n):
    """Positive prior with the given mean and standard deviation: log-normal, sampled in the unconstrained space."""
    s2 = math.log(1 + (std / mean) ** 2)
    return rng.normal(math.log(mean) - s2 / 2, math.sqrt(s2), n)


def update_ensemble(theta, g, y, Gamma, rng):
    """Ensemble Kalman inversion, perturbed observations: θ ← θ + C^{θg}(C^{gg} + Γ)⁻¹(y + η − g); failed members
    (NaN observables) are redrawn from the Gaussian fitted to the successful ones."""
    good = np.isfinite(g).all(axis=0)
    th, gg = theta[:, good], g[:, good]
    dth, dg = th - th.mean(axis=1, keepdims=True), gg - gg.mean(axis=1, keepdims=True)
    J = th.shape[1]
    Ctg, Cgg = dth @ dg.T / J, dg @ dg.T / J
    eta = rng.multivariate_normal(np.zeros(len(y)), Gamma, J).T
    new = theta.copy()
    new[:, good] = th + Ctg @ np.linalg.solve(Cgg + Gamma, y[:, None] + eta - gg)
    if (~good).any():
        mu, cov = new[:, good].mean(axis=1), np.cov(new[:, good]) + 1e-12 * np.eye(new.shape[0])
        new[:, ~good] = rng.multivariate_normal(mu, cov, int((~good).sum())).T
    return new


def main(N_ensemble=8, N_iterations=5, seed=41, **kw):
    rng = np.random.default_rng(seed)
    Gamma = np.diag([0.001, 0.0001, 0.002, 5.0, 5.0])
    truth = np.array([0.15 / day, 0.7 / day, 2.4, 0.01 / day])
    obs, _ = G(truth[:, None], **kw)
    y = obs[:, 0] + rng.multivariate_normal(np.zeros(5), Gamma)
    # priors of data_assimilation.jl:104-107, sampled in log space
    theta = np.stack([constrained_gaussian(rng, 0.1953 / day, 0.05 / day, N_ensemble),
                      constrained_gaussian(rng, 0.6989 / day, 0.1 / day, N_ensemble),
                      constrained_gaussian(rng, 2.3868, 0.5, N_ensemble),
                      constrained_gaussian(rng, 0.0101 / day, 0.01 / day, N_ensemble)])
    history = []
    for it in range(N_iterations):
        g, P = G(np.exp(theta), **kw)  # one device run for the whole ensemble
        misfit = float(np.nanmean(((g - y[:, None]) ** 2) / np.diag(Gamma)[:, None]))
        history.append(misfit)
        print(f"iteration {it + 1}: mean normalised misfit {misfit:.3f}, ensemble mean "
              f"{np.exp(theta).mean(axis=1) * [day, day, 1, day]}")
        theta = update_ensemble(theta, g, y, Gamma, rng)
    final = np.exp(theta)
    print("truth (α, μ₀ per day, k_N, m_P per day):", truth * [day, day, 1, day])
    print("final ensemble mean                    :", final.mean(axis=1) * [day, day, 1, day])
    return truth, final, history


if __name__ == "__main__":
    main()
