"""TEST INFRASTRUCTURE — a second, independent restatement of the sugar-kelp individual model in plain Python floats.

Written method by method from the reference's
  src/Models/Individuals/SugarKelp/SugarKelp.jl:85-139   (keyword defaults, incl. the derived ones)
  src/Models/Individuals/SugarKelp/equations.jl:1-255    (A, N, C equations and every helper)
  src/Models/Individuals/SugarKelp/coupling.jl:3-57      (the eight coupled tracer tendencies)
  src/Utils/solvers.jl:1-23                              (NewtonRaphsonSolver, as used for the light inhibition β)
without looking at oracle/src/oracle_kelp.c, and sharing no code with it: agreement of the two (tests/test_oracle_kelp.py,
tests/golden/kelp_rates.json) is what pins the C oracle — and through it csrc/kelp.cu — on absolute kelp rates, which the
reference's own tests (test/test_sugar_kelp.jl: conservation only) do not pin.  Only tests/ and scripts/make_kelp_golden.py
import this file.
"""
import math

day = 86400.0


class LinearOptimalTemperatureRange:
    """equations.jl:203-222"""

    def __init__(self, lower_optimal=10.0, upper_optimal=15.0, lower_gradient=None, upper_gradient=-0.25):
        self.lower_optimal, self.upper_optimal = lower_optimal, upper_optimal
        self.lower_gradient = 1 / (lower_optimal + 1.8) if lower_gradient is None else lower_gradient
        self.upper_gradient = upper_gradient

    def __call__(self, T):
        Tl, Tu, al, au = self.lower_optimal, self.upper_optimal, self.lower_gradient, self.upper_gradient
        return (max(0.0, al * (T - Tl) + 1) * (T < Tl) + max(0.0, au * (T - Tu) + 1) * (T > Tu) + 1.0 * (Tl <= T <= Tu))


def newton_raphson(f, df, x0, params, max_iters=1000, atol=None):
    """solvers.jl:6-23 with the kelp's `atol = eps(1e-9)` (SugarKelp.jl:139)."""
    atol = math.ulp(1e-9) if atol is None else atol
    x, n = x0, 0
    fx = f(x, params)
    while abs(fx) > atol and n < max_iters:
        fx = f(x, params)
        x -= fx / df(x, params)
        n += 1
    return x


def maximum_photosynthesis(alpha, beta):  # equations.jl:127
    return alpha / math.log(1 + alpha / beta) * (alpha / (alpha + beta)) * (beta / (alpha + beta)) ** (beta / alpha)


def beta_residual(beta, p):  # :128
    return maximum_photosynthesis(p["alpha"], beta) - p["Pm"] / p["Is"]


def d_beta_maximum_photosynthesis(beta, p):  # :129
    a = p["alpha"]
    L = math.log(a / beta + 1)
    return (a * (beta / (beta + a)) ** (beta / a) * ((L * beta ** 2 + a * L * beta) * math.log(beta / (beta + a)) + a ** 2)) \
        / (L ** 2 * beta * (beta + a) ** 2)


def day_length(phi, n):  # :244-252
    n -= 171
    M = (356.5291 + 0.98560028 * n) % 360
    C = 1.9148 * math.sin(M * math.pi / 180) + 0.02 * math.sin(2 * M * math.pi / 180) + 0.0003 * math.sin(3 * M * math.pi / 180)
    lam = (M + C + 180 + 102.9372) % 360
    delta = math.asin(math.sin(lam * math.pi / 180) * math.sin(23.44 * math.pi / 180))
    omega = (math.sin(-0.83 * math.pi / 180) * math.sin(phi * math.pi / 180) * math.sin(delta)) \
        / (math.cos(phi * math.pi / 180) * math.cos(delta))
    return omega / 180


def normed_day_length_change(phi, n):  # :242
    return (day_length(phi, n) - day_length(phi, n - 1)) / (day_length(phi, 76) - day_length(phi, 75))


class SugarKelp:
    """SugarKelp.jl:85-139: every keyword with its default; derived defaults follow the keywords they are built from."""

    def __init__(self, **kw):
        g = kw.get
        self.temperature_limit = g("temperature_limit", LinearOptimalTemperatureRange())
        self.growth_rate_adjustment = g("growth_rate_adjustment", 4.5)
        self.photosynthetic_efficiency = g("photosynthetic_efficiency", 4.15e-5 * 24 * 10 ** 6 / (24 * 60 * 60))
        self.minimum_carbon_reserve = g("minimum_carbon_reserve", 0.01)
        self.structural_carbon = g("structural_carbon", 0.2)
        self.exudation = g("exudation", 0.5)
        self.erosion_exponent = g("erosion_exponent", 0.22)
        self.base_erosion_rate = g("base_erosion_rate", 10.0 ** -6)
        self.saturation_irradiance = g("saturation_irradiance", 90 * day / (10 ** 6))
        self.structural_dry_weight_per_area = g("structural_dry_weight_per_area", 0.5)
        self.minimum_nitrogen_reserve = g("minimum_nitrogen_reserve", 0.0126)
        self.maximum_nitrogen_reserve = g("maximum_nitrogen_reserve", 0.0216)
        r = 1 - self.minimum_nitrogen_reserve / self.maximum_nitrogen_reserve
        self.growth_adjustment_2 = g("growth_adjustment_2", 0.039 / (2 * r))
        self.growth_adjustment_1 = g("growth_adjustment_1", 0.18 / (2 * r) - self.growth_adjustment_2)
        self.maximum_specific_growth_rate = g("maximum_specific_growth_rate", 0.18)
        self.structural_nitrogen = g("structural_nitrogen", 0.0146)
        self.photosynthesis_at_ref_temp_1 = g("photosynthesis_at_ref_temp_1", 1.22e-3 * 24)
        self.photosynthesis_at_ref_temp_2 = g("photosynthesis_at_ref_temp_2", 1.3e-3 * 24)
        self.photosynthesis_ref_temp_1 = g("photosynthesis_ref_temp_1", 285.0)
        self.photosynthesis_ref_temp_2 = g("photosynthesis_ref_temp_2", 288.0)
        self.photoperiod_1 = g("photoperiod_1", 0.85)
        self.photoperiod_2 = g("photoperiod_2", 0.3)
        self.respiration_at_ref_temp_1 = g("respiration_at_ref_temp_1", 2.785e-4 * 24)
        self.respiration_at_ref_temp_2 = g("respiration_at_ref_temp_2", 5.429e-4 * 24)
        self.respiration_ref_temp_1 = g("respiration_ref_temp_1", 285.0)
        self.respiration_ref_temp_2 = g("respiration_ref_temp_2", 290.0)
        self.photosynthesis_arrhenius_temp = g(
            "photosynthesis_arrhenius_temp",
            (1 / self.photosynthesis_ref_temp_1 - 1 / self.photosynthesis_ref_temp_2) ** -1
            * math.log(self.photosynthesis_at_ref_temp_2 / self.photosynthesis_at_ref_temp_1))
        self.photosynthesis_high_temp = g("photosynthesis_high_temp", 296.0)
        self.photosynthesis_high_arrhenius_temp = g("photosynthesis_high_arrhenius_temp", 1414.87)
        self.photosynthesis_low_arrhenius_temp = g("photosynthesis_low_arrhenius_temp", 4547.89)
        self.respiration_arrhenius_temp = g(
            "respiration_arrhenius_temp",
            (1 / self.respiration_ref_temp_1 - 1 / self.respiration_ref_temp_2) ** -1
            * math.log(self.respiration_at_ref_temp_2 / self.respiration_at_ref_temp_1))
        self.nitrate_half_saturation = g("nitrate_half_saturation", 4.0)
        self.ammonia_half_saturation = g("ammonia_half_saturation", 1.3)
        self.maximum_nitrate_uptake = g("maximum_nitrate_uptake", 10 / self.structural_dry_weight_per_area * 24 * 14 / (10 ** 6))
        self.maximum_ammonia_uptake = g("maximum_ammonia_uptake", 12 / self.structural_dry_weight_per_area * 24 * 14 / (10 ** 6))
        self.current_1, self.current_2, self.current_3 = g("current_1", 0.72), g("current_2", 0.28), g("current_3", 0.045)
        self.base_activity_respiration_rate = g("base_activity_respiration_rate", 1.11e-4 * 24)
        self.base_basal_respiration_rate = g("base_basal_respiration_rate", 5.57e-5 * 24)
        self.exudation_redfield_ratio = g("exudation_redfield_ratio", math.inf)
        self.adapted_latitude = g("adapted_latitude", 57.5)

    # ---- helpers, equations.jl:38-252 ------------------------------------------------------------------------------
    def current_factor(self, u, v, w):  # :173-181
        U = math.sqrt(u ** 2 + v ** 2 + w ** 2)
        return self.current_1 * (1 - math.exp(-U / self.current_3)) + self.current_2

    def potential_ammonia_uptake(self, NH4, u, v, w):  # :85-92
        return self.maximum_ammonia_uptake * self.current_factor(u, v, w) * NH4 / (self.ammonia_half_saturation + NH4)

    def area_limitation(self, A):  # :195-201
        return self.growth_adjustment_1 * math.exp(-(A / self.growth_rate_adjustment) ** 2) + self.growth_adjustment_2

    def seasonal_limitation(self, t):  # :225-239
        n = math.floor((t % (364 * day)) / day)
        lam = normed_day_length_change(self.adapted_latitude, n)
        sign = (lam > 0) - (lam < 0)
        return self.photoperiod_1 * (1 + sign * abs(lam) ** .5) + self.photoperiod_2

    def base_growth_limitation(self, t, A, N, C, T):  # :183-193
        return self.temperature_limit(T) * self.area_limitation(A) * self.seasonal_limitation(t)

    def growth(self, t, A, N, C, T, NH4, u, v, w):  # :38-60
        f = self.base_growth_limitation(t, A, N, C, T)
        j_NH4 = self.potential_ammonia_uptake(NH4, u, v, w)
        mu_NH4 = j_NH4 / self.structural_dry_weight_per_area / (N + self.structural_nitrogen)
        mu_N = 1 - self.minimum_nitrogen_reserve / N
        mu_C = 1 - self.minimum_carbon_reserve / C
        return f * min(mu_C, max(mu_N, mu_NH4))

    def nitrate_uptake(self, N, NO3, u, v, w):  # :62-73
        Nmax, Nmin = self.maximum_nitrogen_reserve, self.minimum_nitrogen_reserve
        return max(0.0, self.maximum_nitrate_uptake * self.current_factor(u, v, w) * (Nmax - N) / (Nmax - Nmin) * NO3
                   / (self.nitrate_half_saturation + NO3))

    def ammonia_uptake(self, t, A, N, C, T, NH4, u, v, w):  # :75-83
        j = self.potential_ammonia_uptake(NH4, u, v, w)
        mu = self.growth(t, A, N, C, T, NH4, u, v, w)
        return min(j, mu * self.structural_dry_weight_per_area * (N + self.structural_nitrogen))

    def solve_for_light_inhibition(self, Pm):  # :118-125
        p = {"alpha": self.photosynthetic_efficiency, "Is": self.saturation_irradiance, "Pm": Pm}
        return newton_raphson(beta_residual, d_beta_maximum_photosynthesis, 1e-9, p)

    def photosynthesis(self, T, PAR):  # :94-116 (Tₚₗ is photosynthesis_ref_temp_1, as written there)
        PAR = PAR * (day / (3.99e-10 * 545e12))
        Tk = T + 273.15
        Ta, Tal, Tah = (self.photosynthesis_arrhenius_temp, self.photosynthesis_low_arrhenius_temp,
                        self.photosynthesis_high_arrhenius_temp)
        Tp = Tpl = self.photosynthesis_ref_temp_1
        Tph = self.photosynthesis_high_temp
        alpha, Is = self.photosynthetic_efficiency, self.saturation_irradiance
        Pmax = self.photosynthesis_at_ref_temp_1 * math.exp(Ta / Tp - Ta / Tk) \
            / (1 + math.exp(Tal / Tk - Tal / Tpl) + math.exp(Tah / Tph - Tah / Tk))
        beta = self.solve_for_light_inhibition(Pmax)
        ps = alpha * Is / math.log(1 + alpha / beta)
        return ps * (1 - math.exp(-alpha * PAR / ps)) * math.exp(-beta * PAR / ps)

    def respiration(self, t, A, N, C, T, NO3, NH4, u, v, w, mu):  # :139-157
        Tk = T + 273.15
        f = math.exp(self.respiration_arrhenius_temp / self.respiration_ref_temp_1 - self.respiration_arrhenius_temp / Tk)
        Jm = self.maximum_nitrate_uptake + self.maximum_ammonia_uptake
        J = self.nitrate_uptake(N, NO3, u, v, w) + self.ammonia_uptake(t, A, N, C, T, NH4, u, v, w)
        return f * (self.base_basal_respiration_rate
                    + self.base_activity_respiration_rate * (mu / self.maximum_specific_growth_rate + J / Jm))

    def specific_carbon_exudate(self, C):  # :159-164
        return 1 - math.exp(self.exudation * (self.minimum_carbon_reserve - C))

    def nitrogen_exudate(self, C, T, PAR):  # :166-173
        return self.photosynthesis(T, PAR) * self.specific_carbon_exudate(C) * 14 / 12 / self.exudation_redfield_ratio

    def erosion(self, t, A, N, C, T):  # :175-180
        e = math.exp(self.erosion_exponent * A)
        return self.base_erosion_rate * e / (1 + self.base_erosion_rate * (e - 1))

    # ---- the callables kelp(Val(name), t, A, N, C, u, v, w, T, NO₃, NH₄, PAR) -------------------------------------------
    def __call__(self, name, t, A, N, C, u, v, w, T, NO3, NH4, PAR):
        ka, Ns, Cs = self.structural_dry_weight_per_area, self.structural_nitrogen, self.structural_carbon
        if name == "A":  # equations.jl:1-7
            return A * (self.growth(t, A, N, C, T, NH4, u, v, w) - self.erosion(t, A, N, C, T)) / day
        if name == "N":  # :9-21
            J = self.nitrate_uptake(N, NO3, u, v, w) + self.ammonia_uptake(t, A, N, C, T, NH4, u, v, w)
            e = self.nitrogen_exudate(C, T, PAR)
            mu = self.growth(t, A, N, C, T, NH4, u, v, w)
            return ((J - e) / ka - mu * (N + Ns)) / day
        if name == "C":  # :23-36
            P = self.photosynthesis(T, PAR)
            mu = self.growth(t, A, N, C, T, NH4, u, v, w)
            R = self.respiration(t, A, N, C, T, NO3, NH4, u, v, w, mu)
            e = self.specific_carbon_exudate(C)
            return ((P * (1 - e) - R) / ka - mu * (C + Cs)) / day
        # coupling.jl:3-57
        if name == "NO₃":
            return -self.nitrate_uptake(N, NO3, u, v, w) * A / (day * 14 * 0.001)
        if name == "NH₄":
            return -self.ammonia_uptake(t, A, N, C, T, NH4, u, v, w) * A / (day * 14 * 0.001)
        if name == "DIC":
            P = self.photosynthesis(T, PAR)
            mu = self.growth(t, A, N, C, T, NH4, u, v, w)
            R = self.respiration(t, A, N, C, T, NO3, NH4, u, v, w, mu)
            return -(P - R) * A / (day * 12 * 0.001)
        if name == "O₂":
            return -self("DIC", t, A, N, C, u, v, w, T, NO3, NH4, PAR)
        if name == "DOC":
            return self.specific_carbon_exudate(C) * self.photosynthesis(T, PAR) * A / (day * 12 * 0.001)
        if name == "DON":
            return self("DOC", t, A, N, C, u, v, w, T, NO3, NH4, PAR) / self.exudation_redfield_ratio
        if name == "bPON":
            return self.erosion(t, A, N, C, T) * ka * A * (N + Ns) / (day * 14 * 0.001)
        if name == "bPOC":
            return self.erosion(t, A, N, C, T) * ka * A * (C + Cs) / (day * 12 * 0.001)
        raise KeyError(name)


NAMES = ("A", "N", "C", "NO₃", "NH₄", "DIC", "O₂", "DOC", "DON", "bPOC", "bPON")
