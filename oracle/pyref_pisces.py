"""pyref_pisces — a SECOND, independent restatement of the 24 PISCES tendencies (TEST INFRASTRUCTURE ONLY).

Method-by-method Python transliteration of src/Models/AdvectedPopulations/PISCES/ of the reference (v0.17.6), keeping
its function names, its call structure (every tracer method re-evaluates what it needs) and its operation order; it
shares no code with oracle/src/oracle_pisces.c (written as one pass per tracer over pre-gathered cell values) nor with
the CUDA kernel.  The reference holds no absolute PISCES tendency; its tests check element budgets, which are blind to a
consistently mis-read constant — this module is the second reading.  scripts/make_pisces_golden.py evaluates it at
seeded states covering both sides of every branch (above / below the mixed layer, Ω ≷ 1, southern latitude, oxic /
anoxic, zero biomass) and stores the values in tests/golden/pisces_tendencies.json; tests/test_oracle_pisces.py requires
the C oracle to reproduce them.  Plain Python floats; `f` is a dict of the cell's tracers and auxiliary values.
"""
import math
import sys

EPS0 = sys.float_info.min * sys.float_info.epsilon  # eps(0.0)
DAY, HOUR = 86400.0, 3600.0


def jmax(*a):
    return math.nan if any(x != x for x in a) else max(a)


def jmin(*a):
    return math.nan if any(x != x for x in a) else min(a)


def sind(x):  # Julia's sind / cosd reduce the argument in degrees first
    r = math.fmod(x, 360.0)
    return math.sin(math.radians(r)) if abs(r) <= 45 else math.cos(math.radians(90.0 - r)) if 45 < r <= 135 else _sind_general(r)


def _sind_general(r):
    a = abs(r)
    if a <= 135:
        v = math.cos(math.radians(90.0 - a))
    elif a < 225:
        v = math.sin(math.radians(180.0 - a))
    elif a <= 315:
        v = -math.cos(math.radians(270.0 - a))
    else:
        v = math.sin(math.radians(a - 360.0))
    return v if r >= 0 else -v


def cosd(x):
    return sind(90.0 - math.fmod(x, 360.0))


def cbm_day_length(t, phi, p=0.833):  # src/Utils/Utils.jl:18-34
    J = math.floor((t % (365 * DAY)) / DAY)
    theta = 0.216310 + 2 * math.atan(0.9671396 * math.tan(0.00860 * (J - 186)))
    decl = math.degrees(math.asin(0.39795 * math.cos(theta)))
    L = jmax(-1.0, jmin(1.0, (sind(p) + sind(phi) * sind(decl)) / (cosd(phi) * cosd(decl))))
    return (24 - 24 / 180 * math.degrees(math.acos(L))) * HOUR


class MixedMondo:  # phytoplankton/mixed_mondo.jl:23-93 + growth_rate.jl:107-115 + nutrient_limitation.jl:9-17
    def __init__(self, name, **kw):
        self.name = name  # "P" or "D"
        d = dict(exudated_fraction=0.05, mortality_half_saturation=0.2, linear_mortality_rate=0.01 / DAY,
                 base_quadratic_mortality=0.01 / DAY, minimum_chlorophyll_ratio=0.0033, maximum_iron_ratio=0.06,
                 silicate_half_saturation=2.0, enhanced_silicate_half_saturation=20.9, optimal_silicate_ratio=0.159,
                 threshold_for_size_dependency=1.0, size_ratio=3.0,
                 # GrowthRespirationLimitedProduction
                 base_growth_rate=0.6 / DAY, temperature_sensitivity=1.066, initial_slope_of_PI_curve=2.0, low_light_adaptation=0.0,
                 basal_respiration_rate=0.033 / DAY, reference_growth_rate=1.0 / DAY,
                 # NitrogenIronPhosphateSilicateLimitation
                 optimal_iron_quota=0.007, minimum_silicate_half_saturation=1.0, silicate_half_saturation_parameter=16.6)
        d.update(kw)
        self.__dict__.update(d)


def default_nano():  # mixed_mondo_nano_diatoms.jl:1-14
    return MixedMondo("P", dark_tolerance=3 * DAY, minimum_ammonium_half_saturation=0.013, minimum_nitrate_half_saturation=0.13,
                      minimum_phosphate_half_saturation=0.8, silicate_limited=False, blue_light_absorption=2.1,
                      green_light_absorption=0.42, red_light_absorption=0.4, maximum_quadratic_mortality=0.0,
                      maximum_chlorophyll_ratio=0.033, half_saturation_for_iron_uptake=1.0)


def default_diatoms():  # :15-27
    return MixedMondo("D", dark_tolerance=4 * DAY, minimum_ammonium_half_saturation=0.039, minimum_nitrate_half_saturation=0.39,
                      minimum_phosphate_half_saturation=2.4, silicate_limited=True, blue_light_absorption=1.6,
                      green_light_absorption=0.69, red_light_absorption=0.7, maximum_quadratic_mortality=0.03 / DAY,
                      maximum_chlorophyll_ratio=0.05, half_saturation_for_iron_uptake=3.0)


class Zoo:  # zooplankton/food_quality_dependant.jl:9-76, defaults.jl:1-19
    def __init__(self, name, **kw):
        self.name = name  # "Z" or "M"
        d = dict(temperature_sensitivity=1.079, food_threshold_concentration=0.3, specific_food_threshold_concentration=0.001,
                 grazing_half_saturation=20.0, non_assimilated_fraction=0.3, mortality_half_saturation=0.2,
                 dissolved_excretion_fraction=0.6)
        d.update(kw)
        self.__dict__.update(d)


def default_micro():
    return Zoo("Z", maximum_grazing_rate=3 / DAY, food_preferences=(1.0, 0.5, 0.1, 0.0), quadratic_mortality=0.004 / DAY,
               linear_mortality=0.03 / DAY, minimum_growth_efficiency=0.3, maximum_flux_feeding_rate=0.0,
               undissolved_calcite_fraction=0.5, iron_ratio=0.01)


def default_meso():
    return Zoo("M", maximum_grazing_rate=0.75 / DAY, food_preferences=(0.3, 1.0, 0.3, 1.0), quadratic_mortality=0.03 / DAY,
               linear_mortality=0.005 / DAY, minimum_growth_efficiency=0.35, maximum_flux_feeding_rate=2e3 / 1e6,
               undissolved_calcite_fraction=0.75, iron_ratio=0.015)


class PISCES:
    """bgc::PISCES with the default component types (PISCES.jl:288-408)."""

    def __init__(self, latitude=45.0, **kw):
        self.nano, self.diatoms, self.micro, self.meso = default_nano(), default_diatoms(), default_micro(), default_meso()
        d = dict(
            base_rain_ratio=0.3,
            # MicroAndMeso micro_and_meso.jl:3-15
            microzooplankton_bacteria_concentration=0.7, mesozooplankton_bacteria_concentration=1.4,
            bacteria_concentration_depth_exponent=0.684, doc_half_saturation_for_bacterial_activity=417.0,
            nitrate_half_saturation_for_bacterial_activity=0.03, ammonia_half_saturation_for_bacterial_activity=0.003,
            phosphate_half_saturation_for_bacterial_activity=0.003, iron_half_saturation_for_bacterial_activity=0.01,
            # DissolvedOrganicCarbon
            dom_remineralisation_rate=0.3 / DAY, dom_reference_bacteria_concentration=1.0, dom_temperature_sensitivity=1.066,
            dom_aggregation_parameters=tuple(x * (10 ** -6 / DAY) for x in (0.37, 102, 3530, 5095, 114)),
            # TwoCompartmentCarbonIronParticles
            pom_temperature_sensitivity=1.066, pom_base_breakdown_rate=0.025 / DAY,
            pom_aggregation_parameters=tuple(x * (10 ** -6 / DAY) for x in (25.9, 4452, 3.3, 47.1)),
            minimum_iron_scavenging_rate=3e-5 / DAY, load_specific_iron_scavenging_rate=0.005 / DAY,
            bacterial_iron_uptake_efficiency=0.16, small_fraction_of_bacterially_consumed_iron=0.12 / 0.16,
            large_fraction_of_bacterially_consumed_iron=0.04 / 0.16, base_liable_silicate_fraction=0.5,
            fast_dissolution_rate_of_silicate=0.025 / DAY, slow_dissolution_rate_of_silicate=0.003 / DAY,
            base_calcite_dissolution_rate=0.197 / DAY, calcite_dissolution_exponent=1.0, maximum_iron_ratio_in_bacteria=0.06,
            iron_half_saturation_for_bacteria=0.3, maximum_bacterial_growth_rate=0.6 / DAY,
            # NitrateAmmonia
            maximum_nitrification_rate=0.05 / DAY, maximum_fixation_rate=0.013 / DAY, iron_half_saturation_for_fixation=0.1,
            phosphate_half_saturation_for_fixation=0.8, light_saturation_for_fixation=50.0,
            # SimpleIron
            excess_scavenging_enhancement=1000.0, maximum_ligand_concentration=0.6, dissolved_ligand_ratio=0.09,
            # Oxygen
            ratio_for_respiration=133 / 122, ratio_for_nitrification=32 / 122,
            # PISCES.jl:288-310
            first_anoxia_threshold=6.0, second_anoxia_threshold=1.0, nitrogen_redfield_ratio=16 / 122,
            phosphate_redfield_ratio=1 / 122, mixed_layer_shear=1.0, background_shear=0.01)
        d.update(kw)
        self.__dict__.update(d)
        self.latitude = latitude

    # ---- common.jl
    def anoxia_factor(self, O2):  # :57-62
        return jmin(1, jmax(0, 0.4 * (self.first_anoxia_threshold - O2) / (self.second_anoxia_threshold + O2)))

    # ---- phytoplankton -------------------------------------------------------------------------------------------
    @staticmethod
    def phytoplankton_concentrations(ph, f):
        return (f["P"], f["PChl"], f["PFe"]) if ph.name == "P" else (f["D"], f["DChl"], f["DFe"])

    @staticmethod
    def size_factor(ph, I):  # mixed_mondo.jl:207-215
        I1 = jmin(I, ph.threshold_for_size_dependency)
        I2 = jmax(0, I - ph.threshold_for_size_dependency)
        return (I1 + ph.size_ratio * I2) / (I1 + I2 + EPS0)

    def nutrient_limitation(self, ph, f):  # nutrient_limitation.jl:20-73
        I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
        NO3, NH4, PO4, Si, Sip = f["NO₃"], f["NH₄"], f["PO₄"], f["Si"], f["Si_clim"]
        tFe = 0 if I == 0 else IFe / (I + EPS0)
        tChl = 0 if I == 0 else IChl / (12 * I + EPS0)
        Kbar = self.size_factor(ph, I)
        Kno, Knh = ph.minimum_nitrate_half_saturation * Kbar, ph.minimum_ammonium_half_saturation * Kbar
        Kp, Ksi = ph.minimum_phosphate_half_saturation * Kbar, ph.minimum_silicate_half_saturation * Kbar
        nl = lambda N1, N2, K1, K2: (K2 * N1) / (K1 * K2 + K1 * N2 + K2 * N1 + EPS0)  # noqa: E731
        LNO3 = nl(NO3, NH4, Kno, Knh)
        LNH4 = nl(NH4, NO3, Knh, Kno)
        LN = LNO3 + LNH4
        LPO4 = PO4 / (PO4 + Kp + EPS0)
        tm = 10 ** 3 * (0.0016 / 55.85 * 12 * tChl + 1.5 * 1.21e-5 * 14 / (55.85 * 7.625) * LN + 1.15e-4 * 14 / (55.85 * 7.625) * LNO3)
        LFe = jmin(1, jmax(0, (tFe - tm) / ph.optimal_iron_quota))
        pk = ph.silicate_half_saturation_parameter
        KSi = Ksi + 7 * Sip ** 2 / (pk ** 2 + Sip ** 2)
        LSi = Si / (Si + KSi)
        LSi = LSi if ph.silicate_limited else math.inf
        return jmin(LN, LPO4, LFe, LSi), LFe, LPO4, LN, LNO3, LNH4

    @staticmethod
    def base_production_rate(ph, T):  # growth_rate.jl:158-165
        return ph.base_growth_rate * ph.temperature_sensitivity ** T

    def growth_rate(self, ph, f, L):  # growth_rate.jl:3-47 (GrowthRespirationLimitedProduction light limitation :117-124)
        I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
        PAR = ph.blue_light_absorption * f["PAR₁"] + ph.green_light_absorption * f["PAR₂"] + ph.red_light_absorption * f["PAR₃"]
        day_length = cbm_day_length(self.latitude, f["t"])  # bgc.day_length(φ, clock.time): arguments swapped in the reference
        dark_residence_time = jmax(0, f["zₑᵤ"] - f["zₘₓₗ"]) ** 2 / f["κ"]
        mui = ph.base_growth_rate * ph.temperature_sensitivity ** f["T"]
        f1 = 1.5 * day_length / (day_length + 0.5 * DAY)
        f2 = 1 - dark_residence_time / (dark_residence_time + ph.dark_tolerance)
        alpha = ph.initial_slope_of_PI_curve * (1 + ph.low_light_adaptation * math.exp(-PAR))
        theta = IChl / (12 * I + EPS0)
        fl = 1 - math.exp(-alpha * theta * PAR / (day_length * (ph.basal_respiration_rate + ph.reference_growth_rate)))
        return mui * f1 * f2 * fl * L

    def total_production(self, ph, f):  # mixed_mondo.jl:169-175
        I = self.phytoplankton_concentrations(ph, f)[0]
        L = self.nutrient_limitation(ph, f)[0]
        return self.growth_rate(ph, f, L) * I

    def production_and_energy_assimilation_absorption_ratio(self, ph, f):  # growth_rate.jl:126-156
        I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
        PAR = ph.blue_light_absorption * f["PAR₁"] + ph.green_light_absorption * f["PAR₂"] + ph.red_light_absorption * f["PAR₃"]
        day_length = cbm_day_length(f["t"], self.latitude)  # bgc.day_length(clock.time, φ): the documented order
        f1 = 1.5 * day_length / (day_length + 0.5 * DAY)
        L = self.nutrient_limitation(ph, f)[0]
        mu = self.growth_rate(ph, f, L)
        mucheck = mu / f1 * day_length
        alpha = ph.initial_slope_of_PI_curve * (1 + ph.low_light_adaptation * math.exp(-PAR))
        return mu, 12 * mucheck * I / (alpha * IChl * PAR + EPS0) * L

    def phyto_mortality(self, ph, f):  # mixed_mondo.jl:137-167
        I = self.phytoplankton_concentrations(ph, f)[0]
        L = self.nutrient_limitation(ph, f)[0]
        linear = ph.linear_mortality_rate * I / (I + ph.mortality_half_saturation) * I
        w = ph.base_quadratic_mortality + ph.maximum_quadratic_mortality * 0.25 * (1 - L ** 2) / (0.25 + L ** 2)
        shear = self.background_shear if f["z"] < f["zₘₓₗ"] else self.mixed_layer_shear
        return linear, shear * w * I ** 2

    def iron_uptake(self, ph, f):  # mixed_mondo.jl:177-205
        I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
        tFe = IFe / (I + EPS0)
        L, LFe = self.nutrient_limitation(ph, f)[:2]
        mui = self.base_production_rate(ph, f["T"])
        K = ph.half_saturation_for_iron_uptake * self.size_factor(ph, I)
        L1 = f["Fe"] / (f["Fe"] + K + EPS0)
        L2 = 4 - 4.5 * LFe / (LFe + 1)
        tm = ph.maximum_iron_ratio
        return (1 - ph.exudated_fraction) * tm * L1 * L2 * jmax(0, (1 - tFe / tm) / (1.05 - tFe / tm)) * mui * I

    def silicate_uptake(self, ph, f):  # mixed_mondo.jl:217-248
        I = self.phytoplankton_concentrations(ph, f)[0]
        Si = f["Si"]
        L, LFe, LPO4, LN = self.nutrient_limitation(ph, f)[:4]
        mu = self.growth_rate(ph, f, L)
        mui = self.base_production_rate(ph, f["T"])
        L1 = Si / (Si + ph.silicate_half_saturation + EPS0)
        K2 = ph.enhanced_silicate_half_saturation
        L2 = Si ** 3 / (Si ** 3 + K2 ** 3) if self.latitude < 0 else 0
        F1 = jmin(mu / (mui * L + EPS0), LFe, LPO4, LN)
        F2 = jmin(1, 2.2 * jmax(0, L1 - 0.5))
        t1 = ph.optimal_silicate_ratio * L1 * jmin(5.4, (4.4 * math.exp(-4.23 * F1) * F2 + 1) * (1 + 2 * L2))
        return (1 - ph.exudated_fraction) * t1 * mu * I

    def uptake(self, ph, which, f):  # mixed_mondo.jl:250-268
        if which == "Fe":
            return self.iron_uptake(ph, f)
        nlim = self.nutrient_limitation(ph, f)
        LN, LNO3, LNH4 = nlim[3], nlim[4], nlim[5]
        muI = self.total_production(ph, f)
        return muI * (LNO3 if which == "NO₃" else LNH4) / (LN + EPS0)

    # NanoAndDiatoms sums (nano_and_diatoms.jl)
    def phyto_uptake(self, which, f):
        return self.uptake(self.nano, which, f) + self.uptake(self.diatoms, which, f)

    def phyto_total_production(self, f):
        return self.total_production(self.nano, f) + self.total_production(self.diatoms, f)

    def dissolved_exudate(self, f):
        return (self.nano.exudated_fraction * self.total_production(self.nano, f)
                + self.diatoms.exudated_fraction * self.total_production(self.diatoms, f))

    # ---- zooplankton -----------------------------------------------------------------------------------------------
    def _food(self, f):  # defaults.jl:39-56 — always (P, D, POC, Z), whatever prey_names says
        food = (f["P"], f["D"], f["POC"], f["Z"])
        iron = (f["PFe"] / (f["P"] + EPS0), f["DFe"] / (f["D"] + EPS0), f["SFe"] / (f["POC"] + EPS0), self.micro.iron_ratio)
        return food, iron

    def _specific_grazing(self, zoo, f):
        N = 3 if zoo.name == "Z" else 4  # length(prey_names): (:P, :D, :POC) / (:P, :D, :Z, :POC)
        p, J = zoo.food_preferences, zoo.specific_food_threshold_concentration
        food, iron = self._food(f)
        base = zoo.maximum_grazing_rate * zoo.temperature_sensitivity ** f["T"]
        total_food = sum(food[n] * p[n] for n in range(N))
        available = sum(jmax(0.0, food[n] - J) * p[n] for n in range(N))
        limited = jmax(0, available - jmin(available / 2, zoo.food_threshold_concentration))
        return base * limited / (zoo.grazing_half_saturation + total_food), available, total_food, N, food, iron, p, J

    def zoo_grazing(self, zoo, f):  # food_quality_dependant.jl:96-141 → (g I, growth efficiency)
        tsg, available, total_food, N, food, iron, p, J = self._specific_grazing(zoo, f)
        total_iron = sum(iron[n] * p[n] for n in range(N))
        ratio = total_iron / (zoo.iron_ratio * tsg + EPS0)
        e = jmin(1, ratio) * jmin(zoo.minimum_growth_efficiency, (1 - zoo.non_assimilated_fraction) * ratio)
        return tsg * f[zoo.name], e

    def grazing_on(self, zoo, prey, f):  # :198-233
        tsg, available, total_food, N, food, iron, p, J = self._specific_grazing(zoo, f)
        pref = {"P": p[0], "D": p[1], "POC": p[2], "Z": p[3]}.get(prey, 0)
        return pref * jmax(0, f[prey] - J) * tsg / (available + EPS0) * f[zoo.name]

    def flux_rate(self, name, f):  # two_size_class.jl:91-98 (w already interpolated to the cell centre)
        return f[name] * (f["wPOC"] if name in ("POC", "SFe") else f["wGOC"])

    def zoo_flux_feeding(self, zoo, f, prey=None):  # :143-158, :235-252
        base = zoo.maximum_flux_feeding_rate * zoo.temperature_sensitivity ** f["T"]
        flux = (self.flux_rate("POC", f) + self.flux_rate("GOC", f)) if prey is None else self.flux_rate(prey, f)
        return base * flux * f[zoo.name]

    def zoo_mortality(self, zoo, f):  # :160-176
        I = f[zoo.name]
        return zoo.temperature_sensitivity ** f["T"] * I * (
            zoo.quadratic_mortality * I + zoo.linear_mortality * (I / (I + zoo.mortality_half_saturation) + 3 * self.anoxia_factor(f["O₂"])))

    def zoo_linear_mortality(self, zoo, f):  # :178-192
        I = f[zoo.name]
        return (zoo.temperature_sensitivity ** f["T"] * zoo.linear_mortality
                * (I / (I + zoo.mortality_half_saturation) + 3 * self.anoxia_factor(f["O₂"])) * I)

    def iron_grazing(self, zoo, f):  # iron_grazing.jl:2-33
        tsg, available, total_food, N, food, iron, p, J = self._specific_grazing(zoo, f)
        s = sum(jmax(0.0, food[n] - J) * p[n] * iron[n] for n in range(N)) * tsg / (available + EPS0)
        return s * f[zoo.name]

    def iron_flux_feeding(self, zoo, f):  # :35-51
        base = zoo.maximum_flux_feeding_rate * zoo.temperature_sensitivity ** f["T"]
        return base * (self.flux_rate("SFe", f) + self.flux_rate("BFe", f)) * f[zoo.name]

    def non_assimilated_waste(self, zoo, f):  # grazing_waste.jl:3-10
        return zoo.non_assimilated_fraction * (self.zoo_grazing(zoo, f)[0] + self.zoo_flux_feeding(zoo, f))

    def excretion(self, zoo, f):  # :12-19
        gI, e = self.zoo_grazing(zoo, f)
        return (1 - zoo.non_assimilated_fraction - e) * (gI + self.zoo_flux_feeding(zoo, f))

    def inorganic_excretion(self, f):
        return sum(z.dissolved_excretion_fraction * self.excretion(z, f) for z in (self.micro, self.meso))

    def organic_excretion(self, f):
        return sum((1 - z.dissolved_excretion_fraction) * self.excretion(z, f) for z in (self.micro, self.meso))

    def non_assimilated_iron_waste(self, zoo, f):  # :33-40
        return zoo.non_assimilated_fraction * (self.iron_grazing(zoo, f) + self.iron_flux_feeding(zoo, f))

    def non_assimilated_iron(self, zoo, f):  # :42-63
        gI, e = self.zoo_grazing(zoo, f)
        assimilated = zoo.iron_ratio * e * (gI + self.zoo_flux_feeding(zoo, f))
        gIFe, gfIFe = self.iron_grazing(zoo, f), self.iron_flux_feeding(zoo, f)
        return (gIFe + gfIFe) - zoo.non_assimilated_fraction * (gIFe + gfIFe) - assimilated

    def upper_trophic_waste(self, f):  # mortality_waste.jl:27-41 (meso only)
        z = self.meso
        return 1 / (1 - z.minimum_growth_efficiency) * z.quadratic_mortality * z.temperature_sensitivity ** f["T"] * f["M"] ** 2

    def upper_trophic_respiration_product(self, f):
        z = self.meso
        return (1 - z.minimum_growth_efficiency - z.non_assimilated_fraction) * self.upper_trophic_waste(f)

    def upper_trophic_excretion(self, f):
        return (1 - self.meso.dissolved_excretion_fraction) * self.upper_trophic_respiration_product(f)

    def upper_trophic_respiration(self, f):
        return self.meso.dissolved_excretion_fraction * self.upper_trophic_respiration_product(f)

    def upper_trophic_fecal_production(self, f):
        return self.meso.non_assimilated_fraction * self.upper_trophic_waste(f)

    def grazing(self, prey, f):  # MicroAndMeso micro_and_meso.jl:50-52
        return self.grazing_on(self.micro, prey, f) + self.grazing_on(self.meso, prey, f)

    def total_grazing(self, prey, f):  # micro_meso_zoo_coupling.jl:27-32
        ff = self.zoo_flux_feeding(self.micro, f, prey) + self.zoo_flux_feeding(self.meso, f, prey)
        return ff if prey in ("GOC", "BFe", "PSi", "CaCO₃") else self.grazing(prey, f) + ff

    def bacteria_concentration(self, f):  # micro_and_meso.jl:85-105
        zm = jmin(f["zₘₓₗ"], f["zₑᵤ"])
        surface = jmin(4, self.microzooplankton_bacteria_concentration * f["Z"] + self.mesozooplankton_bacteria_concentration * f["M"])
        z = f["z"]
        return (1 if z >= zm else (zm / z) ** self.bacteria_concentration_depth_exponent) * surface

    def bacteria_activity(self, f):  # :107-132
        K3, K4 = self.nitrate_half_saturation_for_bacterial_activity, self.ammonia_half_saturation_for_bacterial_activity
        NH4, NO3, PO4, Fe, DOC = f["NH₄"], f["NO₃"], f["PO₄"], f["Fe"], f["DOC"]
        LN = (K3 * NH4 + K4 * NO3) / (K3 * K4 + K3 * NH4 + K4 * NO3)
        LP = PO4 / (PO4 + self.phosphate_half_saturation_for_bacterial_activity)
        LF = Fe / (Fe + self.iron_half_saturation_for_bacterial_activity)
        return jmin(LN, LP, LF) * (DOC / (DOC + self.doc_half_saturation_for_bacterial_activity))

    # ---- dissolved organic matter ----------------------------------------------------------------------------------
    def dom_degradation(self, f):  # dissolved_organic_carbon.jl:54-70
        return (self.dom_remineralisation_rate * self.dom_temperature_sensitivity ** f["T"] * self.bacteria_activity(f)
                * self.bacteria_concentration(f) / self.dom_reference_bacteria_concentration * f["DOC"])

    def dom_aggregation(self, f):  # :72-95
        a1, a2, a3, a4, a5 = self.dom_aggregation_parameters
        DOC, POC, GOC = f["DOC"], f["POC"], f["GOC"]
        shear = self.background_shear if f["z"] < f["zₘₓₗ"] else self.mixed_layer_shear
        P1 = shear * (a1 * DOC + a2 * POC) * DOC
        P2 = shear * (a3 * GOC) * DOC
        P3 = (a4 * POC + a5 * DOC) * DOC
        return P1 + P2 + P3, P1, P2, P3

    def free_iron(self, f):  # iron/iron.jl:25-37
        ligands = jmax(0.6, 0.09 * (f["DOC"] + 40) - 3)
        K = math.exp(16.27 - 1565.7 / jmax(f["T"] + 273.15, 5))
        D = 1 + K * ligands - K * f["Fe"]
        return (-D + math.sqrt(D ** 2 + 4 * K * f["Fe"])) / (2 * K)

    def aggregation_of_colloidal_iron(self, f):  # dissolved_organic_carbon.jl:97-113
        _, P1, P2, P3 = self.dom_aggregation(f)
        colloidal = 0.5 * (f["Fe"] - self.free_iron(f))
        C1 = (P1 + P3) * colloidal / (f["DOC"] + EPS0)
        C2 = P2 * colloidal / (f["DOC"] + EPS0)
        return C1 + C2, C1, C2

    # ---- particulate organic matter ---------------------------------------------------------------------------------
    def pom_aggregation(self, f):  # two_size_class.jl:109-125
        a1, a2, a3, a4 = self.pom_aggregation_parameters
        POC, GOC = f["POC"], f["GOC"]
        shear = self.background_shear if f["z"] < f["zₘₓₗ"] else self.mixed_layer_shear
        return shear * (a1 * POC ** 2 + a2 * POC * GOC) + a3 * POC * GOC + a4 * POC ** 2

    def specific_degradation_rate(self, f):  # :127-137
        return self.pom_base_breakdown_rate * self.pom_temperature_sensitivity ** f["T"] * (1 - 0.45 * self.anoxia_factor(f["O₂"]))

    def pom_degradation(self, name, f):
        return self.specific_degradation_rate(f) * f[name]

    def iron_scavenging_rate(self, f):  # particulate_organic_matter/iron.jl:97-106
        return self.minimum_iron_scavenging_rate + self.load_specific_iron_scavenging_rate * (f["POC"] + f["GOC"] + f["CaCO₃"] + f["PSi"])

    def bacterial_iron_uptake(self, f):  # :108-124
        mu = self.maximum_bacterial_growth_rate * self.pom_temperature_sensitivity ** f["T"]
        Fe = f["Fe"]
        return (mu * self.bacteria_activity(f) * self.maximum_iron_ratio_in_bacteria * Fe / (Fe + self.iron_half_saturation_for_bacteria)
                * self.bacteria_concentration(f) * self.bacterial_iron_uptake_efficiency)

    def rain_ratio(self, f):  # nano_diatom_coupling.jl:57-124
        nl = self.nutrient_limitation(self.nano, f)
        LPO4, LN = nl[2], nl[3]
        L_CaCO3 = jmin(LN, f["Fe"] / (f["Fe"] + 0.05), LPO4)
        T, PAR = f["T"], f["PAR"]
        return (self.base_rain_ratio * L_CaCO3 * jmax(1.0, f["P"] / 2) * (jmax(0, PAR - 1) / (4 + PAR)) * (30 / (30 + PAR))
                * jmax(0, T / (T + 0.1)) * (1 + math.exp(-(T - 10) ** 2 / 25)) * jmin(1, -50 / f["zₘₓₗ"]))

    def calcite_production(self, f):  # :101-110
        R = self.rain_ratio(f)
        lin, quad = self.phyto_mortality(self.nano, f)
        loss = (self.micro.undissolved_calcite_fraction * self.grazing_on(self.micro, "P", f)
                + self.meso.undissolved_calcite_fraction * self.grazing_on(self.meso, "P", f))
        return R * (loss + (lin + quad) / 2)

    def calcite_dissolution(self, f):  # particulate_organic_matter/calcite.jl:9-19
        return self.base_calcite_dissolution_rate * jmax(0, 1 - f["Ω"]) ** self.calcite_dissolution_exponent * f["CaCO₃"]

    def particulate_silicate_dissolution(self, f):  # particulate_organic_matter/silicate.jl:9-48
        ll, lr = self.fast_dissolution_rate_of_silicate, self.slow_dissolution_rate_of_silicate
        zm = jmin(f["zₘₓₗ"], f["zₑᵤ"])
        z, T, Si = f["z"], f["T"], f["Si"]
        chi = self.base_liable_silicate_fraction * (1 if z >= zm else math.exp((ll - lr) * (zm - z) / f["wGOC"]))
        l0 = chi * ll + (1 - chi) * lr
        eq = 10 ** (6.44 - 968 / (T + 273.15))
        sat = (eq - Si) / eq
        lam = l0 * (0.225 * (1 + T / 15) * sat + 0.775 * ((1 + T / 400) ** 4 * sat) ** 9)
        return lam * f["PSi"]

    # ---- nitrogen ---------------------------------------------------------------------------------------------------
    def nitrification(self, f):  # nitrate_ammonia.jl:49-59
        return self.maximum_nitrification_rate * f["NH₄"] / (1 + f["mixed_layer_PAR"]) * (1 - self.anoxia_factor(f["O₂"]))

    def nitrogen_fixation(self, f):  # :61-88
        LN = self.nutrient_limitation(self.nano, f)[3]
        limit = 0.01 if LN >= 0.8 else 1 - LN
        mu = self.base_production_rate(self.nano, f["T"])
        Fe, PO4 = f["Fe"], f["PO₄"]
        nutrient = jmin(Fe / (Fe + self.iron_half_saturation_for_fixation), PO4 / (PO4 + self.phosphate_half_saturation_for_fixation))
        return self.maximum_fixation_rate * jmax(0, mu - 2.15) * limit * nutrient * (1 - math.exp(-f["PAR"] / self.light_saturation_for_fixation))

    # ---- bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields) ------------------------------------------------
    def __call__(self, name, f):
        nano, dia, micro, meso = self.nano, self.diatoms, self.micro, self.meso
        if name in ("P", "D"):  # mixed_mondo_nano_diatoms.jl:45-58
            ph = nano if name == "P" else dia
            lin, quad = self.phyto_mortality(ph, f)
            return (1 - ph.exudated_fraction) * self.total_production(ph, f) - (lin + quad) - self.grazing(name, f)
        if name in ("PChl", "DChl"):  # :60-76
            ph = nano if name == "PChl" else dia
            I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
            mu, rho = self.production_and_energy_assimilation_absorption_ratio(ph, f)
            growth = (1 - ph.exudated_fraction) * 12 * (ph.minimum_chlorophyll_ratio + (ph.maximum_chlorophyll_ratio - ph.minimum_chlorophyll_ratio) * rho) * mu * I
            lin, quad = self.phyto_mortality(ph, f)
            return growth - ((lin + quad) + self.grazing(ph.name, f)) * (IChl / (12 * I + EPS0)) * 12
        if name in ("PFe", "DFe"):  # :78-94
            ph = nano if name == "PFe" else dia
            I, IChl, IFe = self.phytoplankton_concentrations(ph, f)
            lin, quad = self.phyto_mortality(ph, f)
            return self.iron_uptake(ph, f) - ((lin + quad) + self.grazing(ph.name, f)) * (IFe / (I + EPS0))
        if name == "DSi":  # :96-111
            lin, quad = self.phyto_mortality(dia, f)
            return self.silicate_uptake(dia, f) - ((lin + quad) + self.grazing("D", f)) * (f["DSi"] / (f["D"] + EPS0))
        if name in ("Z", "M"):  # micro_and_meso.jl:36-48, food_quality_dependant.jl:78-86
            zoo = micro if name == "Z" else meso
            gI, e = self.zoo_grazing(zoo, f)
            net = e * (gI + self.zoo_flux_feeding(zoo, f)) - self.zoo_mortality(zoo, f)
            return net - (self.grazing_on(meso, "Z", f) if name == "Z" else 0.0)
        if name == "DOC":  # dissolved_organic_carbon.jl:39-52
            return (self.dissolved_exudate(f) + self.upper_trophic_excretion(f) + self.organic_excretion(f) + self.pom_degradation("POC", f)
                    - self.dom_degradation(f) - self.dom_aggregation(f)[0])
        if name == "POC":  # particulate_organic_matter/carbon.jl:3-26
            _, P1, _, P3 = self.dom_aggregation(f)
            plin, pquad = self.phyto_mortality(nano, f)
            R = self.rain_ratio(f)
            dlin = self.phyto_mortality(dia, f)[0]
            small_phyto = (1 - R / 2) * (plin + pquad) + dlin / 2
            return (self.non_assimilated_waste(micro, f) + small_phyto + self.zoo_mortality(micro, f) + (P1 + P3) + self.pom_degradation("GOC", f)
                    - self.total_grazing("POC", f) - self.pom_aggregation(f) - self.pom_degradation("POC", f))
        if name == "GOC":  # :28-50
            plin, pquad = self.phyto_mortality(nano, f)
            R = self.rain_ratio(f)
            dlin, dquad = self.phyto_mortality(dia, f)
            large_phyto = R / 2 * (plin + pquad) + dlin / 2 + dquad
            return (self.non_assimilated_waste(meso, f) + large_phyto + self.zoo_linear_mortality(meso, f) + self.upper_trophic_fecal_production(f)
                    + self.pom_aggregation(f) + self.dom_aggregation(f)[2]
                    - self.total_grazing("GOC", f) - self.pom_degradation("GOC", f))
        if name in ("SFe", "BFe"):  # particulate_organic_matter/iron.jl:2-89
            POC, SFe, GOC, BFe = f["POC"], f["SFe"], f["GOC"], f["BFe"]
            tS, tB = SFe / (POC + EPS0), BFe / (GOC + EPS0)
            plin, pquad = self.phyto_mortality(nano, f)
            R = self.rain_ratio(f)
            dlin, dquad = self.phyto_mortality(dia, f)
            tP, tD = f["PFe"] / (f["P"] + EPS0), f["DFe"] / (f["D"] + EPS0)
            lFe, Fep, BactFe = self.iron_scavenging_rate(f), self.free_iron(f), self.bacterial_iron_uptake(f)
            _, C1, C2 = self.aggregation_of_colloidal_iron(f)
            if name == "SFe":
                phyto = (1 - R / 2) * (plin + pquad) * tP + dlin * tD / 2
                return (self.non_assimilated_iron_waste(micro, f) + phyto + self.zoo_mortality(micro, f) * micro.iron_ratio
                        + self.pom_degradation("BFe", f) + lFe * POC * Fep + self.small_fraction_of_bacterially_consumed_iron * BactFe + C1
                        - self.total_grazing("POC", f) * tS - self.pom_aggregation(f) * tS - self.pom_degradation("SFe", f))
            phyto = R / 2 * (plin + pquad) * tP + (dlin / 2 + dquad) * tD
            return (self.non_assimilated_iron_waste(meso, f) + phyto + self.zoo_linear_mortality(meso, f) * meso.iron_ratio
                    + self.upper_trophic_fecal_production(f) * meso.iron_ratio
                    + lFe * GOC * Fep + self.large_fraction_of_bacterially_consumed_iron * BactFe + C2 + self.pom_aggregation(f) * tS
                    - self.total_grazing("GOC", f) * tB - self.pom_degradation("BFe", f))
        if name == "PSi":  # particulate_organic_matter/silicate.jl:1-7, nano_diatom_coupling.jl:86-99
            dlin, dquad = self.phyto_mortality(dia, f)
            production = (self.grazing("D", f) + dlin + dquad) * (f["DSi"] / (f["D"] + EPS0))
            return production - self.particulate_silicate_dissolution(f)
        if name == "CaCO₃":  # particulate_organic_matter/calcite.jl:1-7
            return self.calcite_production(f) - self.calcite_dissolution(f)
        if name == "NO₃":  # nitrate_ammonia.jl:22-32
            oxic = (1 - self.anoxia_factor(f["O₂"])) * self.dom_degradation(f)
            return self.nitrification(f) + self.nitrogen_redfield_ratio * (oxic - self.phyto_uptake("NO₃", f))
        if name == "NH₄":  # :34-47
            anoxic = self.anoxia_factor(f["O₂"]) * self.dom_degradation(f)
            return (self.nitrogen_fixation(f)
                    + self.nitrogen_redfield_ratio * (anoxic + self.inorganic_excretion(f) + self.upper_trophic_respiration(f) - self.phyto_uptake("NH₄", f))
                    - self.nitrification(f))
        if name == "PO₄":  # phosphate.jl:21-33
            return self.phosphate_redfield_ratio * (self.inorganic_excretion(f) + self.upper_trophic_respiration(f) + self.dom_degradation(f)
                                                    - self.phyto_total_production(f))
        if name == "Fe":  # iron/simple_iron.jl:19-62
            Fe = f["Fe"]
            lFe, Fep = self.iron_scavenging_rate(f), self.free_iron(f)
            Lt = jmax(self.maximum_ligand_concentration, self.dissolved_ligand_ratio * f["DOC"] - self.maximum_ligand_concentration)
            ligand_aggregation = self.excess_scavenging_enhancement * lFe * jmax(0, Fe - Lt) * Fep
            scavenging = lFe * (f["POC"] + f["GOC"]) * Fep
            waste = self.non_assimilated_iron(micro, f) + self.non_assimilated_iron(meso, f)
            upper = meso.iron_ratio * self.upper_trophic_respiration_product(f)
            return (self.pom_degradation("SFe", f) + waste + upper
                    - self.phyto_uptake("Fe", f) - ligand_aggregation - self.aggregation_of_colloidal_iron(f)[0] - scavenging
                    - self.bacterial_iron_uptake(f))
        if name == "Si":  # silicate.jl:20-26
            return self.particulate_silicate_dissolution(f) - self.silicate_uptake(dia, f)
        if name == "DIC":  # inorganic_carbon.jl:32-47
            return (self.inorganic_excretion(f) + self.upper_trophic_respiration(f) + self.dom_degradation(f)
                    + self.calcite_dissolution(f) - self.calcite_production(f) - self.phyto_total_production(f))
        if name == "Alk":  # :49-58
            return self("NH₄", f) - self("NO₃", f) - 2 * self("CaCO₃", f)
        if name == "O₂":  # oxygen.jl:30-51
            tr, tn = self.ratio_for_respiration, self.ratio_for_nitrification
            d = self.anoxia_factor(f["O₂"])
            remin = (tr + tn) * ((1 - d) * self.dom_degradation(f)) + tr * (d * self.dom_degradation(f))
            return (tr * self.phyto_uptake("NH₄", f) + (tr + tn) * self.phyto_uptake("NO₃", f)
                    + tn * self.nitrogen_fixation(f) / self.nitrogen_redfield_ratio
                    - remin - tr * self.inorganic_excretion(f) - tr * self.upper_trophic_respiration(f)
                    - tn * self.nitrification(f) / self.nitrogen_redfield_ratio)
        raise KeyError(name)


TRACERS = ("P", "PChl", "PFe", "D", "DChl", "DFe", "DSi", "Z", "M", "DOC", "POC", "GOC", "SFe", "BFe", "PSi", "CaCO₃",
           "NO₃", "NH₄", "PO₄", "Fe", "Si", "DIC", "Alk", "O₂")
