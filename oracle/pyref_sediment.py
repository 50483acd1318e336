"""pyref_sediment — second, independent restatement of the sediment models' equations (TEST INFRASTRUCTURE ONLY):
src/Models/Sediments/simple_multi_G.jl:165-427 and instant_remineralisation.jl:103-125 transliterated method by method
(one function per `Val`, each re-deriving Nr, Cr, k and the Soetaert (2000) meta-model fractions, like the reference),
with the reference's default parameters (simple_multi_G.jl:83-100).  No code shared with oracle_sediment.c.
The reference's own sediment tests do not run (test/test_sediments.jl:106-163 is commented out), so there is no
reference-side golden; this pins the oracle's reading of the equations on a second reading."""
import math
import sys

EPS0 = sys.float_info.min * sys.float_info.epsilon
DAY = 86400.0


def finite_or_zero(p):
    return p if math.isfinite(p) else 0.0


def log(x):  # Julia's log(0.0) = -Inf (Python raises)
    return -math.inf if x == 0 else math.log(x)


class SimpleMultiG:
    def __init__(self, depth_of_first_cell_centre, carbon=False, **kw):
        d = dict(fast_decay_rate=2 / DAY, slow_decay_rate=0.2 / DAY, fast_fraction=0.74, slow_fraction=0.26, refactory_fraction=0.1,
                 sedimentation_rate=982 * abs(depth_of_first_cell_centre) ** (-1.548), anoxia_half_saturation=1.0,
                 nitrate_oxidation_params=(-1.9785, 0.2261, -0.0615, -0.0289, -0.36109, -0.0232),
                 denitrification_params=(-3.0790, 1.7509, 0.0593, -0.1923, 0.0604, 0.0662),
                 anoxic_params=(-3.9476, 2.6269, -0.2426, -1.3349, 0.1826, -0.0143), sinking_redfield=6.56)
        d.update(kw)
        self.__dict__.update(d)
        self.carbon = carbon

    def reactivity(self, Cs, Cf):  # :370-376
        return (self.slow_decay_rate * Cs + self.fast_decay_rate * Cf) / (Cs + Cf + EPS0)

    def ammonia_oxidation_fraction(self, Nr, Cr, k, NH4, O2):  # :378-394
        A, B, C, D, E, F = self.nitrate_oxidation_params
        ln = (A + B * log(Cr * DAY) * log(O2) + C * log(Cr * DAY) ** 2 + D * log(k * DAY) * log(NH4) + E * log(Cr * DAY)
              + F * log(Cr * DAY) * log(NH4))
        try:
            p = math.exp(ln) / (Nr * DAY) * O2 / (self.anoxia_half_saturation + O2)
        except (ZeroDivisionError, OverflowError, ValueError):
            return 0.0
        return finite_or_zero(p)

    def denitrification_fraction(self, Nr, Cr, k, NO3, O2):  # :396-410
        A, B, C, D, E, F = self.denitrification_params
        ln = (A + B * log(Cr * DAY) + C * log(NO3) ** 2 + D * log(Cr * DAY) ** 2 + E * log(k * DAY) ** 2 + F * log(O2) * log(k))
        try:
            p = math.exp(ln) / (Cr * DAY) * O2 / (self.anoxia_half_saturation + O2)
        except (ZeroDivisionError, OverflowError, ValueError):
            return 0.0
        return finite_or_zero(p)

    def anoxic_remineralisation_fraction(self, Nr, Cr, k, NO3, O2):  # :412-424
        A, B, C, D, E, F = self.anoxic_params
        ln = (A + B * log(Cr * DAY) + C * log(Cr * DAY) ** 2 + D * log(k * DAY) + E * log(O2) * log(k) + F * log(NO3) ** 2)
        try:
            p = math.exp(ln) / (Cr * DAY)
        except (ZeroDivisionError, OverflowError, ValueError):
            return 0.0
        return finite_or_zero(p)

    def solid_deposition_fraction(self):  # :426 — the literal 0.223 (solid_dep_params says 0.233 and is not used)
        return 0.223 * self.sedimentation_rate ** 0.336

    def _rates(self, pools):
        ls, lf = self.slow_decay_rate, self.fast_decay_rate
        Ns, Nf = pools[0], pools[1]
        Nr = ls * Ns + lf * Nf
        if self.carbon:
            Cs, Cf = pools[3], pools[4]
            return Nr, ls * Cs + lf * Cf, self.reactivity(Cs, Cf)
        R = self.sinking_redfield
        return Nr, Nr * R, self.reactivity(Ns * R, Nf * R)

    def __call__(self, name, pools, NO3, NH4, O2, fN, fC=0.0):
        fr, fs, ff = self.refactory_fraction, self.slow_fraction, self.fast_fraction
        if name == "Ns":
            return (1 - fr) * fs * fN - self.slow_decay_rate * pools[0]
        if name == "Nf":
            return (1 - fr) * ff * fN - self.fast_decay_rate * pools[1]
        if name == "Nr":
            return fr * fN
        if name == "Cs":
            return (1 - fr) * fs * fC - self.slow_decay_rate * pools[3]
        if name == "Cf":
            return (1 - fr) * ff * fC - self.fast_decay_rate * pools[4]
        if name == "Cr":
            return fr * fC
        Nr, Cr, k = self._rates(pools)
        if name == "DIC":
            return Cr
        pn = self.ammonia_oxidation_fraction(Nr, Cr, k, NH4, O2)
        pnp = self.denitrification_fraction(Nr, Cr, k, NO3, O2)
        if name == "NH₄":
            return (1 - pn) * Nr + 0.8 * pnp * Cr
        if name == "NO₃":
            return pn * Nr - 0.8 * pnp * Cr
        if name == "O₂":
            pa = self.anoxic_remineralisation_fraction(Nr, Cr, k, NO3, O2)
            ps = self.solid_deposition_fraction()
            return -(1 - pa * ps - pnp) * O2 / (self.anoxia_half_saturation + O2) * Cr - 2 * pn * Nr
        raise KeyError(name)
