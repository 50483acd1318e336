"""TEST INFRASTRUCTURE — a second, independent restatement of the light path in plain Python floats, one column at a time.

Written from the reference's
  src/Light/2band.jl:1-33                         two-band PAR, serial in z
  src/Light/multi_band.jl:136-140, 149-166        numerical_mean and the per-band scan
  src/Light/compute_euphotic_depth.jl:3-29        zₑᵤ
  src/Models/AdvectedPopulations/PISCES/mean_mixed_layer_properties.jl:24-49   mixed-layer mean
without looking at oracle/src/oracle_light.c and sharing no code with it.  Columns are 1-based like the reference:
`c[k]` for k = 1 … Nz is the interior, `c[Nz + 1]` the level above the surface; zc[k] centres, zf[k] faces (Nz + 1 of them).
Only tests/ imports this file.
"""
import math


class Col:
    """1-based view of a Python list (index 0 ↔ k = lo)."""

    def __init__(self, values, lo=1):
        self.v, self.lo = list(values), lo

    def __getitem__(self, k):
        return self.v[k - self.lo]

    def __setitem__(self, k, x):
        self.v[k - self.lo] = x


def two_band(Nz, zc, zf, P, PAR0, kr, kb, xr, xb, er, eb, r, Rcp):
    """2band.jl:1-33 → PAR[1 … Nz]."""
    PAR = Col([0.0] * Nz)
    ir = (zf[Nz + 1] - zc[Nz]) * (P[Nz] * Rcp / r) ** er
    ib = (zf[Nz + 1] - zc[Nz]) * (P[Nz] * Rcp / r) ** eb
    PAR[Nz] = PAR0 * (math.exp(kr * zc[Nz] - xr * ir) + math.exp(kb * zc[Nz] - xb * ib)) / 2
    for k in range(Nz - 1, 0, -1):
        ir += (zc[k + 1] - zf[k + 1]) * (P[k + 1] * Rcp / r) ** er + (zf[k + 1] - zc[k]) * (P[k] * Rcp / r) ** er
        ib += (zc[k + 1] - zf[k + 1]) * (P[k + 1] * Rcp / r) ** eb + (zf[k + 1] - zc[k]) * (P[k] * Rcp / r) ** eb
        PAR[k] = PAR0 * (math.exp(kr * zc[k] - xr * ir) + math.exp(kb * zc[k] - xb * ib)) / 2
    return PAR


def numerical_mean(lam, C, idx1, idx2):
    """multi_band.jl:136-140 with 1-based idx1 < idx2 into the base tables."""
    lam, C = Col(lam), Col(C)
    integral = sum((C[n] + C[n - 1]) * (lam[n] - lam[n - 1]) / 2 for n in range(idx1 + 1, idx2 + 1))
    return integral / (lam[idx2] - lam[idx1])


def band_coefficients(bands, base_bands, base):
    """multi_band.jl:96-104: findlast(base_bands .<= edge) for both edges of every band."""
    out = []
    for lo, hi in bands:
        idx1 = max(n + 1 for n, b in enumerate(base_bands) if b <= lo)
        idx2 = max(n + 1 for n, b in enumerate(base_bands) if b <= hi)
        out.append(numerical_mean(base_bands, base, idx1, idx2))
    return out


def multi_band(Nz, zc, Chl, PAR0, division, kw, e, chi):
    """multi_band.jl:149-166, one band → field[1 … Nz]."""
    f = Col([0.0] * Nz)
    f[Nz] = PAR0 * division * math.exp(zc[Nz] * (kw + chi * Chl[Nz] ** e))
    for k in range(Nz - 1, 0, -1):
        dz = zc[k] - zc[k + 1]
        f[k] = f[k + 1] * math.exp(dz * (kw + chi * Chl[k] ** e))
    return f


def euphotic_depth(Nz, zc, PAR, cutoff=1 / 1000):
    """compute_euphotic_depth.jl:3-29; PAR[Nz + 1] is read (the level above the surface), zc[0] is the fallback."""
    surface = (PAR[Nz] + PAR[Nz + 1]) / 2
    zeu = -math.inf
    for k in range(Nz - 1, 0, -1):
        if PAR[k] <= surface * cutoff and math.isinf(zeu):
            zeu = zc[k] + (math.log(surface * cutoff) - math.log(PAR[k])) * (zc[k] - zc[k + 1]) \
                / (math.log(PAR[k]) - math.log(PAR[k + 1]))
    return zeu if math.isfinite(zeu) else zc[0]


def mixed_layer_mean(Nz, zf, zmxl, C):
    """mean_mixed_layer_properties.jl:24-49."""
    total, depth = 0.0, 0.0
    for k in range(Nz, 0, -1):
        zk, zk1 = zf[k], zf[k + 1]
        dzk = zk1 - zk
        dzk1 = zk1 - zmxl if zk1 > zmxl else 0.0
        dz = dzk if zk >= zmxl else dzk1
        total += C[k] * dz
        depth += dz
    return total / depth
