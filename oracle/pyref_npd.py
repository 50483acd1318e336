"""pyref_npd — a SECOND, independent restatement of the Nutrients–Plankton–Detritus tendencies (TEST INFRASTRUCTURE ONLY).

The reference holds no absolute tendency value for this family (its tests check conservation and the zero state), so
a transcription error that keeps the budgets closed — a wrong half-saturation constant, a swapped preference — would
pass every pin of the C oracle.  This module transliterates the reference's Julia methods one by one, keeping their
names, call structure (every tracer method re-evaluates what it needs) and operation order, from
src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/{nutrients,plankton,detritus,carbonate_system,oxygen}.jl; it
shares no code with oracle/src/oracle_npd.c, which was written from the same source as one fused pass per tracer.
scripts/make_npd_golden.py evaluates it at seeded states and stores the values in tests/golden/npd_tendencies.json;
tests/test_oracle_npd.py requires the C oracle to reproduce them to 1e-15 relative.  Plain Python floats (IEEE double).
"""
import math
import sys

EPS0 = sys.float_info.min * sys.float_info.epsilon  # eps(0.0) = 5e-324


def jl_eps(x):
    return math.ulp(x) if math.isfinite(x) else math.nan


def jl_max(a, b):
    return math.nan if (a != a or b != b) else max(a, b)


class NPD:
    """bgc::NutrientsPlanktonDetritus: `nutrients` in {"Nutrient", "NitrateAmmonia", "NitrateAmmoniaIron"}, `detritus`
    in {None, "Detritus", "TwoParticleAndDissolved", "VariableRedfieldDetritus"}; `pl`, `de`, `ox`, `nu` are dicts of
    the reference's @kwdef fields (defaults below)."""

    def __init__(self, nutrients, detritus, pl, de, nu, ox, mortality="Quadratic", grazing="Quadratic", light="Mondo"):
        self.nutrients, self.detritus, self.pl, self.de, self.nu, self.ox = nutrients, detritus, pl, de, nu, ox
        self.mortality_formulation, self.grazing_formulation, self.light = mortality, grazing, light

    # ---- plankton.jl:83-90
    def mortality(self, X, m):
        return m * X if self.mortality_formulation == "Linear" else m * X ** 2

    def concentration_limit(self, X, k):
        return X / (X + k) if self.grazing_formulation == "Linear" else X ** 2 / (X ** 2 + k ** 2)

    # ---- detritus.jl accessors
    def small_particulate_concentration(self, f):
        d = self.detritus
        if d is None:
            return 0.0
        if d == "Detritus":
            return f["D"] * self.de["small_particle_fraction"]
        return f["sPOM"] if d == "TwoParticleAndDissolved" else f["sPON"]

    def large_particulate_concentration(self, f):
        return f["bPOM"] if self.detritus == "TwoParticleAndDissolved" else f["bPON"]

    def dissolved_organic_nitrogen(self, f):
        return f["DOM"] if self.detritus == "TwoParticleAndDissolved" else f["DON"]

    def small_particulate_carbon_concentration(self, f):
        return f["sPOM"] * self.de["redfield_ratio"] if self.detritus == "TwoParticleAndDissolved" else f["sPOC"]

    def large_particulate_carbon_concentration(self, f):
        return f["bPOM"] * self.de["redfield_ratio"] if self.detritus == "TwoParticleAndDissolved" else f["bPOC"]

    def dissolved_organic_carbon(self, f):
        return f["DOM"] * self.de["redfield_ratio"] if self.detritus == "TwoParticleAndDissolved" else f["DOC"]

    # ---- plankton.jl
    def weighted_phytoplankton_preference(self, P, sPOM):  # :384-388
        p = self.pl["preference_for_phytoplankton"]
        return p * P / (p * P + (1 - p) * sPOM + EPS0)

    def total_grazing(self, f):  # :118-135
        kG, g = self.pl["grazing_half_saturation"], self.pl["maximum_grazing_rate"]
        Z, P, sPOM = f["Z"], f["P"], self.small_particulate_concentration(f)
        p = self.weighted_phytoplankton_preference(P, sPOM)
        food = p * P + (1 - p) * sPOM
        return g * self.concentration_limit(food, kG) * Z

    def grazing(self, f, name):  # :340-382
        kG, g = self.pl["grazing_half_saturation"], self.pl["maximum_grazing_rate"]
        Z, P, sPOM = f["Z"], f["P"], self.small_particulate_concentration(f)
        p = self.weighted_phytoplankton_preference(P, sPOM)
        food = p * P + (1 - p) * sPOM
        L = self.concentration_limit(food, kG)
        if name == "P":
            return g * p * L * P / (food + jl_eps(food)) * Z
        if name in ("sPOM", "sPON", "D"):
            return g * (1 - p) * L * sPOM / (food + jl_eps(food)) * Z
        if name == "sPOC":
            return self.grazing(f, "sPOM") * self.pl["redfield_ratio"]
        raise KeyError(name)

    def nitrogen_limitation(self, f):  # :150-166
        k3, k4, psi = self.pl["nitrate_half_saturation"], self.pl["ammonia_half_saturation"], self.pl["nitrate_ammonia_inhibition"]
        NO3, NH4 = f["NO₃"], f["NH₄"]
        nl = NO3 * math.exp(-psi * NH4) / (NO3 + k3)
        al = jl_max(0, NH4 / (k4 + NH4))
        return (nl + al) / 2

    def nutrient_limitation(self, f):  # :168-189
        if self.nutrients == "NitrateAmmonia":
            return self.nitrogen_limitation(f)
        if self.nutrients == "NitrateAmmoniaIron":
            return self.nitrogen_limitation(f) * (f["Fe"] / (self.pl["iron_half_saturation"] + f["Fe"]))
        return f["N"] / (f["N"] + self.pl["nitrate_half_saturation"])

    def light_limitation(self, PAR, kPAR):  # :216-220
        return PAR / (kPAR + PAR) if self.light == "Mondo" else PAR / math.sqrt(PAR ** 2 + kPAR ** 2)

    def temperature_limitation(self, f):  # :206-214
        Q10 = self.pl["temperature_coefficient"]
        return 1.0 if Q10 is None else Q10 ** (f["T"] / 10)

    def phytoplankton_growth(self, f):  # :191-204
        Ln = self.nutrient_limitation(f)
        Ll = self.light_limitation(f["PAR"], self.pl["light_half_saturation"])
        Lt = self.temperature_limitation(f)
        return self.pl["phytoplankton_maximum_growth_rate"] * Ll * Ln * Lt * f["P"]

    def nutrient_uptake(self, f, name):  # :222-269
        mu = self.phytoplankton_growth(f)
        a, g = self.pl["ammonia_fraction_of_exudate"], self.pl["phytoplankton_exudation_fraction"]
        if name == "Fe":
            return self.pl["iron_ratio"] * mu
        if name == "N":
            return mu * (1 - a * g)
        k3, k4, psi = self.pl["nitrate_half_saturation"], self.pl["ammonia_half_saturation"], self.pl["nitrate_ammonia_inhibition"]
        NO3, NH4 = f["NO₃"], f["NH₄"]
        nl = NO3 * math.exp(-psi * NH4) / (NO3 + k3)
        al = jl_max(0, NH4 / (k4 + NH4))
        if name == "NO₃":
            return mu * nl / (nl + al + EPS0)
        waste = a * g * mu
        return mu * al / (nl + al + EPS0) - waste

    def phytoplankton_primary_production(self, f):  # :271-280
        a, g = self.pl["ammonia_fraction_of_exudate"], self.pl["phytoplankton_exudation_fraction"]
        rho, R = self.pl["carbon_calcite_ratio"], self.pl["redfield_ratio"]
        return (1 + rho * (1 - g) - a * g) * self.phytoplankton_growth(f) * R

    def plankton_inorganic_nitrogen_waste(self, f):  # :282-296
        aP, aZ = self.pl["phytoplankton_solid_waste_fraction"], self.pl["excretion_inorganic_fraction"]
        nuP = self.mortality(f["P"], self.pl["phytoplankton_mortality_rate"])
        return aZ * self.pl["zooplankton_excretion_rate"] * f["Z"] + (1 - aP) * nuP

    def plankton_inorganic_carbon_waste(self, f):
        return self.pl["redfield_ratio"] * self.plankton_inorganic_nitrogen_waste(f)

    def plankton_organic_nitrogen_waste(self, f):  # :302-315
        aZ, mu = self.pl["excretion_inorganic_fraction"], self.pl["zooplankton_excretion_rate"]
        aP, g = self.pl["ammonia_fraction_of_exudate"], self.pl["phytoplankton_exudation_fraction"]
        return (1 - aP) * g * self.phytoplankton_growth(f) + (1 - aZ) * mu * f["Z"]

    def plankton_organic_carbon_waste(self, f):
        return self.pl["redfield_ratio"] * self.plankton_organic_nitrogen_waste(f)

    def solid_waste(self, f):  # :320-335
        aP, aZ = self.pl["phytoplankton_solid_waste_fraction"], self.pl["zooplankton_assimilation_fraction"]
        G = self.total_grazing(f)
        nuP = self.mortality(f["P"], self.pl["phytoplankton_mortality_rate"])
        return (1 - aZ) * G + aP * nuP + self.pl["zooplankton_mortality_rate"] * f["Z"] ** 2

    def solid_carbon_waste(self, f):
        return self.solid_waste(f) * self.pl["redfield_ratio"]

    def calcite_production(self, f):  # :390-403
        R, rho, eta = self.pl["redfield_ratio"], self.pl["carbon_calcite_ratio"], self.pl["zooplankton_gut_calcite_dissolution"]
        G = self.grazing(f, "P")
        nu = self.mortality(f["P"], self.pl["phytoplankton_mortality_rate"])
        return (G * (1 - eta) + nu) * rho * R

    def calcite_dissolution(self, f):  # :405-431: the fixed-Redfield override for everything but VariableRedfieldDetritus
        R, rho = self.pl["redfield_ratio"], self.pl["carbon_calcite_ratio"]
        G = self.grazing(f, "P")
        if self.detritus == "VariableRedfieldDetritus":
            return G * self.pl["zooplankton_gut_calcite_dissolution"] * rho * R
        nu = self.mortality(f["P"], self.pl["phytoplankton_mortality_rate"])
        return (G + nu) * rho * R

    def calcite_uptake(self, f):  # :433-440
        return 2 * self.pl["carbon_calcite_ratio"] * self.phytoplankton_growth(f) * self.pl["redfield_ratio"]

    # ---- detritus.jl wastes
    def detritus_inorganic_nitrogen_waste(self, f):
        d = self.detritus
        if d is None:
            return self.plankton_organic_nitrogen_waste(f) + self.solid_waste(f)
        if d == "Detritus":
            return f["D"] * self.de["remineralisation_rate"]
        a, s, b, dm = (self.de[k] for k in ("remineralisation_inorganic_fraction", "small_remineralisation_rate",
                                            "large_remineralisation_rate", "dissolved_remineralisation_rate"))
        return a * (s * self.small_particulate_concentration(f) + b * self.large_particulate_concentration(f)) \
            + dm * self.dissolved_organic_nitrogen(f)

    def detritus_organic_nitrogen_waste(self, f):
        a, s, b = (self.de[k] for k in ("remineralisation_inorganic_fraction", "small_remineralisation_rate", "large_remineralisation_rate"))
        return (1 - a) * (s * self.small_particulate_concentration(f) + b * self.large_particulate_concentration(f))

    def detritus_inorganic_carbon_waste(self, f):
        d = self.detritus
        if d is None:
            return (self.plankton_organic_nitrogen_waste(f) + self.solid_waste(f)) * self.pl["redfield_ratio"]
        if d == "Detritus":
            return f["D"] * self.de["remineralisation_rate"] * self.de["redfield_ratio"]
        a, s, b, dm = (self.de[k] for k in ("remineralisation_inorganic_fraction", "small_remineralisation_rate",
                                            "large_remineralisation_rate", "dissolved_remineralisation_rate"))
        return a * (s * self.small_particulate_carbon_concentration(f) + b * self.large_particulate_carbon_concentration(f)) \
            + dm * self.dissolved_organic_carbon(f)

    def detritus_organic_carbon_waste(self, f):
        a, s, b = (self.de[k] for k in ("remineralisation_inorganic_fraction", "small_remineralisation_rate", "large_remineralisation_rate"))
        return (1 - a) * (s * self.small_particulate_carbon_concentration(f) + b * self.large_particulate_carbon_concentration(f))

    def nitrification(self, f):  # nutrients.jl:52,91
        return 0.0 if self.nutrients == "Nutrient" else self.nu["nitrification_rate"] * f["NH₄"]

    # ---- bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)
    def __call__(self, name, f):
        d, de = self.detritus, self.de
        if name == "Fe":
            return -self.nutrient_uptake(f, "Fe")
        if name == "NO₃":
            return self.nitrification(f) - self.nutrient_uptake(f, "NO₃")
        if name == "NH₄":
            return (self.plankton_inorganic_nitrogen_waste(f) + self.detritus_inorganic_nitrogen_waste(f)
                    - self.nitrification(f) - self.nutrient_uptake(f, "NH₄"))
        if name == "N":
            return (self.plankton_inorganic_nitrogen_waste(f) + self.detritus_inorganic_nitrogen_waste(f)
                    - self.nutrient_uptake(f, "N"))
        if name == "P":
            g, m = self.pl["phytoplankton_exudation_fraction"], self.pl["phytoplankton_mortality_rate"]
            return (1 - g) * self.phytoplankton_growth(f) - self.grazing(f, "P") - self.mortality(f["P"], m)
        if name == "Z":
            a, m, mu = self.pl["zooplankton_assimilation_fraction"], self.pl["zooplankton_mortality_rate"], self.pl["zooplankton_excretion_rate"]
            return a * self.total_grazing(f) - m * f["Z"] ** 2 - mu * f["Z"]
        if name == "D":
            return (self.plankton_organic_nitrogen_waste(f) + self.solid_waste(f) - self.grazing(f, "D")
                    - de["remineralisation_rate"] * f["D"])
        if name in ("DOM", "DON"):
            return (self.plankton_organic_nitrogen_waste(f) + self.detritus_organic_nitrogen_waste(f)
                    - de["dissolved_remineralisation_rate"] * self.dissolved_organic_nitrogen(f))
        if name in ("sPOM", "sPON"):
            return (de["small_solid_waste_fraction"] * self.solid_waste(f) - self.grazing(f, name)
                    - de["small_remineralisation_rate"] * self.small_particulate_concentration(f))
        if name in ("bPOM", "bPON"):
            return ((1 - de["small_solid_waste_fraction"]) * self.solid_waste(f)
                    - de["large_remineralisation_rate"] * self.large_particulate_concentration(f))
        if name == "sPOC":
            return (de["small_solid_waste_fraction"] * self.solid_carbon_waste(f) - self.grazing(f, "sPOC")
                    - de["small_remineralisation_rate"] * self.small_particulate_carbon_concentration(f))
        if name == "bPOC":
            return ((1 - de["small_solid_waste_fraction"]) * self.solid_carbon_waste(f) + self.calcite_production(f)
                    - de["large_remineralisation_rate"] * self.large_particulate_carbon_concentration(f))
        if name == "DOC":
            return (self.plankton_organic_carbon_waste(f) + self.detritus_organic_carbon_waste(f)
                    - de["dissolved_remineralisation_rate"] * self.dissolved_organic_carbon(f))
        if name == "DIC":
            return (-self.phytoplankton_primary_production(f) + self.plankton_inorganic_carbon_waste(f)
                    + self.detritus_inorganic_carbon_waste(f) + self.calcite_dissolution(f))
        if name == "Alk":
            if self.nutrients == "Nutrient":
                return self("N", f) - 2 * self.calcite_uptake(f) + 2 * self.calcite_dissolution(f)
            return (self("NH₄", f) * (1 - 1 / 16) - self("NO₃", f) * (1 + 1 / 16)
                    - 2 * self.calcite_uptake(f) + 2 * self.calcite_dissolution(f))
        if name == "O₂":
            Rp, Rn = self.ox["respiration_oxygen_nitrogen_ratio"], self.ox["nitrification_oxygen_nitrogen_ratio"]
            mu = self.phytoplankton_growth(f)
            if self.nutrients == "Nutrient":
                # oxygen.jl:33-43 specialises on `NutrientsPlanktonDetritus{<:Any, <:Nutrient, …}` — the PLANKTON slot — so it
                # is never selected; a Nutrient model runs the generic method, where bgc(Val(:NH₄)) falls back to
                # zero(grid) (no NH₄ method for Nutrient) and nitrification(::Nutrient) = 0
                return Rp * mu - (Rp - Rn) * 0.0 - Rp * self.nitrification(f)
            return Rp * mu - (Rp - Rn) * self("NH₄", f) - Rp * self.nitrification(f)
        raise KeyError(name)


DAY = 86400.0
PHYTOZOO_DEFAULTS = dict(  # plankton.jl:19-58
    nitrate_half_saturation=0.7, ammonia_half_saturation=0.001, iron_half_saturation=2e-4, nitrate_ammonia_inhibition=3.0,
    light_half_saturation=33.0, phytoplankton_maximum_growth_rate=2.42e-5, iron_ratio=4.6375e-5,
    phytoplankton_exudation_fraction=0.05, ammonia_fraction_of_exudate=0.75, temperature_coefficient=None,
    phytoplankton_mortality_rate=5.8e-7, zooplankton_mortality_rate=2.31e-6, zooplankton_excretion_rate=5.8e-7,
    phytoplankton_solid_waste_fraction=1.0, excretion_inorganic_fraction=0.5, preference_for_phytoplankton=0.5,
    maximum_grazing_rate=9.26e-6, grazing_half_saturation=1.0, zooplankton_assimilation_fraction=0.7,
    zooplankton_calcite_dissolution=0.3, redfield_ratio=6.56, carbon_calcite_ratio=0.1, zooplankton_gut_calcite_dissolution=0.3,
    phytoplankton_chlorophyll_ratio=1.31)
TWO_PARTICLE_DEFAULTS = dict(  # detritus.jl:25-37
    remineralisation_inorganic_fraction=0.0, small_remineralisation_rate=5.88e-7, large_remineralisation_rate=5.88e-7,
    dissolved_remineralisation_rate=3.86e-7, small_solid_waste_fraction=0.5, redfield_ratio=6.56)
DETRITUS_DEFAULTS = dict(remineralisation_rate=0.1213 / DAY, small_particle_fraction=0.5, redfield_ratio=6.56)  # :264-270
OXYGEN_DEFAULTS = dict(respiration_oxygen_nitrogen_ratio=10.75, nitrification_oxygen_nitrogen_ratio=2.0)
NUTRIENT_DEFAULTS = dict(nitrification_rate=5.8e-7)
NPZD_PLANKTON = dict(PHYTOZOO_DEFAULTS, **dict(  # constructors.jl:177-227 (Kuhn et al. 2015)
    nitrate_half_saturation=2.3868, phytoplankton_maximum_growth_rate=0.6989 / DAY, phytoplankton_exudation_fraction=0.0,
    temperature_coefficient=1.88, phytoplankton_mortality_rate=(0.066 + 0.0101) / DAY, preference_for_phytoplankton=1.0,
    grazing_half_saturation=0.5573, zooplankton_mortality_rate=0.3395 / DAY, zooplankton_excretion_rate=0.0102 / DAY,
    zooplankton_assimilation_fraction=0.9116, excretion_inorganic_fraction=1.0,
    phytoplankton_solid_waste_fraction=0.0101 / (0.066 + 0.0101), maximum_grazing_rate=2.1522 / DAY,
    light_half_saturation=(0.6989 / DAY) / (0.1953 / DAY)))


def lobster(detritus="TwoParticleAndDissolved", iron=False):
    return NPD("NitrateAmmoniaIron" if iron else "NitrateAmmonia", detritus, dict(PHYTOZOO_DEFAULTS), dict(TWO_PARTICLE_DEFAULTS),
               dict(NUTRIENT_DEFAULTS), dict(OXYGEN_DEFAULTS))


def npzd():
    return NPD("Nutrient", "Detritus", dict(NPZD_PLANKTON), dict(DETRITUS_DEFAULTS), dict(NUTRIENT_DEFAULTS), dict(OXYGEN_DEFAULTS),
               mortality="Linear", grazing="Quadratic", light="Analytical")
