"""PISCES — host-side mirror of src/Models/AdvectedPopulations/PISCES/ (constructor `PISCES(; grid, …)`
PISCES.jl:288-408 with the same keyword arguments and defaults, tracer / auxiliary-field lists, state
update order of update_state.jl:1-17, conserved groups of coupling_utils.jl:11-33).  The 24 tendencies are
evaluated by ONE fused kernel (csrc/pisces_tendencies.cu) behind `obm_pisces_tendencies`.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib
from .carbon_chemistry import CarbonChemistry
from .grids import CenterField, Field, Field2D, RectilinearGrid, ZFaceField, current_stream_ptr, require_cuda
from .light import (MultiBandPhotosyntheticallyActiveRadiation, compute_euphotic_depth, compute_mixed_layer_mean,
                    default_surface_PAR)

day = 86400.0
hour = 3600.0

TRACERS = ("P", "PChl", "PFe", "D", "DChl", "DFe", "DSi", "Z", "M", "DOC", "POC", "GOC", "SFe", "BFe", "PSi", "CaCO₃",
           "NO₃", "NH₄", "PO₄", "Fe", "Si", "DIC", "Alk", "O₂", "T", "S")  # PISCES.jl:94-105


# ---- day length (src/Utils/Utils.jl:13-34) ----------------------------------------------------------
def _sind(x: float) -> float:
    """Julia's `sind`: the argument is reduced in DEGREES first (`rem(x, 360)` is exact), then folded to |angle| ≤ 45° —
    `math.sin(math.radians(x))` would carry the rounding of x·π/180 (2e-11 absolute at x = 10⁷, which is what the
    reference's swapped `day_length(φ, t)` passes as the latitude: the clock time in seconds)."""
    r = math.fmod(x, 360.0)
    a = abs(r)
    if a < 45.0:
        v = math.sin(math.radians(a))
    elif a <= 135.0:
        v = math.cos(math.radians(90.0 - a))
    elif a < 225.0:
        v = math.sin(math.radians(180.0 - a))
    elif a <= 315.0:
        v = -math.cos(math.radians(270.0 - a))
    else:
        v = math.sin(math.radians(a - 360.0))
    return math.copysign(v, r) if v >= 0 else -math.copysign(-v, r)


def _cosd(x: float) -> float:
    """Julia's `cosd` (degree-exact reduction, see `_sind`)."""
    a = abs(math.fmod(x, 360.0))
    if a <= 45.0:
        return math.cos(math.radians(a))
    if a < 135.0:
        return math.sin(math.radians(90.0 - a))
    if a <= 225.0:
        return -math.cos(math.radians(180.0 - a))
    if a < 315.0:
        return math.sin(math.radians(a - 270.0))
    return math.cos(math.radians(360.0 - a))


@dataclass
class CBMDayLength:
    day_length_coefficient: float = 0.833

    def __call__(self, t, φ):
        p = self.day_length_coefficient
        J = math.floor((t % (365 * day)) / day)
        θ = 0.216310 + 2 * math.atan(0.9671396 * math.tan(0.00860 * (J - 186)))
        # NB: Python NFKC-normalises identifiers, so the reference's ϕ (declination) and φ (latitude) would collide
        decl = math.degrees(math.asin(0.39795 * math.cos(θ)))
        L = max(-1.0, min(1.0, (_sind(p) + _sind(φ) * _sind(decl)) / (_cosd(φ) * _cosd(decl))))
        return (24 - 24 / 180 * math.degrees(math.acos(L))) * hour


@dataclass
class PrescribedLatitude:
    latitude: float = 45.0


class ModelLatitude:
    """`ModelLatitude()` — PISCES/common.jl:9-13,27-28: the latitude of the model grid's own rows, φnode(i, j, k, grid).
    Needs a grid that has one (`LatitudeLongitudeGrid`); the day lengths then differ from row to row and the tendencies
    go through `obm_pisces_tendencies_rows`."""

    def __repr__(self):
        return "ModelLatitude()"


@dataclass
class DepthDependantSinkingSpeed:
    """common.jl:39-55 — note the literal 5000 (the `maximum_depth` field is ignored by the reference)."""
    minimum_speed: float = 30 / day
    maximum_speed: float = 200 / day
    maximum_depth: float = 5000.0

    def face_field(self, grid: RectilinearGrid, mixed_layer_depth: Field, euphotic_depth: Field) -> Field:
        w = ZFaceField(grid, "wGOC")
        zc = torch.from_numpy(grid.zc).to(grid.device).reshape(-1, 1, 1)  # znode(Center) at index k, as the reference does
        zm = torch.minimum(mixed_layer_depth.interior, euphotic_depth.interior)
        vals = -self.minimum_speed + (self.maximum_speed - self.minimum_speed) * torch.clamp(zc - zm, max=0.0) / 5000
        w.face_interior[:grid.Nz] = vals  # face Nz+1: ifelse(k == grid.Nz + 1, 0, w)
        return w


# ---- component parameter sets -------------------------------------------------------------------------
@dataclass
class GrowthRespirationLimitedProduction:  # growth_rate.jl:107-115
    dark_tolerance: float = 3 * day
    base_growth_rate: float = 0.6 / day
    temperature_sensitivity: float = 1.066
    initial_slope_of_PI_curve: float = 2.0
    low_light_adaptation: float = 0.0
    basal_respiration_rate: float = 0.033 / day
    reference_growth_rate: float = 1.0 / day
    kind = _lib_kind = 1


@dataclass
class NutrientLimitedProduction:  # growth_rate.jl:69-75
    dark_tolerance: float = 3 * day
    base_growth_rate: float = 0.6 / day
    temperature_sensitivity: float = 1.066
    initial_slope_of_PI_curve: float = 2.0
    low_light_adaptation: float = 0.0
    basal_respiration_rate: float = 0.0
    reference_growth_rate: float = 0.0
    kind = _lib_kind = 0


@dataclass
class NitrogenIronPhosphateSilicateLimitation:  # nutrient_limitation.jl:10-18
    minimum_ammonium_half_saturation: float = 0.013
    minimum_nitrate_half_saturation: float = 0.13
    minimum_phosphate_half_saturation: float = 0.8
    optimal_iron_quota: float = 0.007
    silicate_limited: bool = False
    minimum_silicate_half_saturation: float = 1.0
    silicate_half_saturation_parameter: float = 16.6


@dataclass
class MixedMondo:  # mixed_mondo.jl:23-93
    growth_rate: object = field(default_factory=GrowthRespirationLimitedProduction)
    nutrient_limitation: NitrogenIronPhosphateSilicateLimitation = field(default_factory=NitrogenIronPhosphateSilicateLimitation)
    exudated_fraction: float = 0.05
    blue_light_absorption: float = 2.1
    green_light_absorption: float = 0.42
    red_light_absorption: float = 0.4
    mortality_half_saturation: float = 0.2
    linear_mortality_rate: float = 0.01 / day
    base_quadratic_mortality: float = 0.01 / day
    maximum_quadratic_mortality: float = 0.0
    minimum_chlorophyll_ratio: float = 0.0033
    maximum_chlorophyll_ratio: float = 0.033
    maximum_iron_ratio: float = 0.06
    silicate_half_saturation: float = 2.0
    enhanced_silicate_half_saturation: float = 20.9
    optimal_silicate_ratio: float = 0.159
    half_saturation_for_iron_uptake: float = 1.0
    threshold_for_size_dependency: float = 1.0
    size_ratio: float = 3.0

    def fill(self, c: _lib.obm_pisces_phyto):
        g, n = self.growth_rate, self.nutrient_limitation
        c.growth_rate_kind = g.kind
        c.silicate_limited = 1 if n.silicate_limited else 0
        for k in ("base_growth_rate", "temperature_sensitivity", "dark_tolerance", "initial_slope_of_PI_curve",
                  "low_light_adaptation", "basal_respiration_rate", "reference_growth_rate"):
            setattr(c, k, float(getattr(g, k)))
        for k in ("minimum_ammonium_half_saturation", "minimum_nitrate_half_saturation", "minimum_phosphate_half_saturation",
                  "optimal_iron_quota", "minimum_silicate_half_saturation", "silicate_half_saturation_parameter"):
            setattr(c, k, float(getattr(n, k)))
        for k in ("exudated_fraction", "blue_light_absorption", "green_light_absorption", "red_light_absorption",
                  "mortality_half_saturation", "linear_mortality_rate", "base_quadratic_mortality",
                  "maximum_quadratic_mortality", "minimum_chlorophyll_ratio", "maximum_chlorophyll_ratio",
                  "maximum_iron_ratio", "silicate_half_saturation", "enhanced_silicate_half_saturation",
                  "optimal_silicate_ratio", "half_saturation_for_iron_uptake", "threshold_for_size_dependency",
                  "size_ratio"):
            setattr(c, k, float(getattr(self, k)))


@dataclass
class NanoAndDiatoms:  # nano_and_diatoms.jl:1-5
    nano: MixedMondo
    diatoms: MixedMondo
    base_rain_ratio: float = 0.3


def MixedMondoNanoAndDiatoms() -> NanoAndDiatoms:
    """mixed_mondo_nano_diatoms.jl:1-28"""
    nano = MixedMondo(growth_rate=GrowthRespirationLimitedProduction(dark_tolerance=3 * day),
                      nutrient_limitation=NitrogenIronPhosphateSilicateLimitation(0.013, 0.13, 0.8, silicate_limited=False),
                      blue_light_absorption=2.1, green_light_absorption=0.42, red_light_absorption=0.4,
                      maximum_quadratic_mortality=0.0, maximum_chlorophyll_ratio=0.033, half_saturation_for_iron_uptake=1.0)
    diatoms = MixedMondo(growth_rate=GrowthRespirationLimitedProduction(dark_tolerance=4 * day),
                         nutrient_limitation=NitrogenIronPhosphateSilicateLimitation(0.039, 0.39, 2.4, silicate_limited=True),
                         blue_light_absorption=1.6, green_light_absorption=0.69, red_light_absorption=0.7,
                         maximum_quadratic_mortality=0.03 / day, maximum_chlorophyll_ratio=0.05,
                         half_saturation_for_iron_uptake=3.0)
    return NanoAndDiatoms(nano, diatoms)


@dataclass
class QualityDependantZooplankton:  # food_quality_dependant.jl:68-106
    maximum_grazing_rate: float
    food_preferences: dict  # keys P, D, POC, Z (NamedTuple order of defaults.jl:4,12)
    maximum_flux_feeding_rate: float
    iron_ratio: float
    minimum_growth_efficiency: float
    quadratic_mortality: float
    linear_mortality: float
    undissolved_calcite_fraction: float
    temperature_sensitivity: float = 1.079
    food_threshold_concentration: float = 0.3
    specific_food_threshold_concentration: float = 0.001
    grazing_half_saturation: float = 20.0
    non_assimilated_fraction: float = 0.3
    mortality_half_saturation: float = 0.2
    dissolved_excretion_fraction: float = 0.6

    def fill(self, c: _lib.obm_pisces_zoo):
        for k in ("temperature_sensitivity", "maximum_grazing_rate", "food_threshold_concentration",
                  "specific_food_threshold_concentration", "grazing_half_saturation", "maximum_flux_feeding_rate",
                  "iron_ratio", "minimum_growth_efficiency", "non_assimilated_fraction", "mortality_half_saturation",
                  "quadratic_mortality", "linear_mortality", "dissolved_excretion_fraction",
                  "undissolved_calcite_fraction"):
            setattr(c, k, float(getattr(self, k)))
        for n, k in enumerate(("P", "D", "POC", "Z")):
            c.food_preferences[n] = float(self.food_preferences[k])


@dataclass
class MicroAndMeso:  # micro_and_meso.jl:3-17
    micro: QualityDependantZooplankton
    meso: QualityDependantZooplankton
    microzooplankton_bacteria_concentration: float = 0.7
    mesozooplankton_bacteria_concentration: float = 1.4
    maximum_bacteria_concentration: float = 4.0
    bacteria_concentration_depth_exponent: float = 0.684
    doc_half_saturation_for_bacterial_activity: float = 417.0
    nitrate_half_saturation_for_bacterial_activity: float = 0.03
    ammonia_half_saturation_for_bacterial_activity: float = 0.003
    phosphate_half_saturation_for_bacterial_activity: float = 0.003
    iron_half_saturation_for_bacterial_activity: float = 0.01


def MicroAndMesoZooplankton() -> MicroAndMeso:
    """zooplankton/defaults.jl:2-21"""
    micro = QualityDependantZooplankton(maximum_grazing_rate=3 / day, food_preferences=dict(P=1.0, D=0.5, POC=0.1, Z=0.0),
                                        quadratic_mortality=0.004 / day, linear_mortality=0.03 / day,
                                        minimum_growth_efficiency=0.3, maximum_flux_feeding_rate=0.0,
                                        undissolved_calcite_fraction=0.5, iron_ratio=0.01)
    meso = QualityDependantZooplankton(maximum_grazing_rate=0.75 / day, food_preferences=dict(P=0.3, D=1.0, POC=0.3, Z=1.0),
                                       quadratic_mortality=0.03 / day, linear_mortality=0.005 / day,
                                       minimum_growth_efficiency=0.35, maximum_flux_feeding_rate=2e3 / 1e6,
                                       undissolved_calcite_fraction=0.75, iron_ratio=0.015)
    return MicroAndMeso(micro, meso)


@dataclass
class DissolvedOrganicCarbon:  # dissolved_organic_carbon.jl:17-35
    remineralisation_rate: float = 0.3 / day
    bacteria_concentration_depth_exponent: float = 0.684
    reference_bacteria_concentration: float = 1.0
    temperature_sensitivity: float = 1.066
    aggregation_parameters: tuple = tuple(a * (10.0 ** -6 / day) for a in (0.37, 102, 3530, 5095, 114))


@dataclass
class TwoCompartmentCarbonIronParticles:  # two_size_class.jl:40-84
    temperature_sensitivity: float = 1.066
    base_breakdown_rate: float = 0.025 / day
    aggregation_parameters: tuple = tuple(a * (10.0 ** -6 / day) for a in (25.9, 4452, 3.3, 47.1))
    minimum_iron_scavenging_rate: float = 3e-5 / day
    load_specific_iron_scavenging_rate: float = 0.005 / day
    bacterial_iron_uptake_efficiency: float = 0.16
    small_fraction_of_bacterially_consumed_iron: float = 0.12 / 0.16
    large_fraction_of_bacterially_consumed_iron: float = 0.04 / 0.16
    base_liable_silicate_fraction: float = 0.5
    fast_dissolution_rate_of_silicate: float = 0.025 / day
    slow_dissolution_rate_of_silicate: float = 0.003 / day
    base_calcite_dissolution_rate: float = 0.197 / day
    calcite_dissolution_exponent: float = 1.0
    maximum_iron_ratio_in_bacteria: float = 0.06
    iron_half_saturation_for_bacteria: float = 0.3
    maximum_bacterial_growth_rate: float = 0.6 / day


@dataclass
class NitrateAmmonia:  # nitrogen/nitrate_ammonia.jl:10-16
    maximum_nitrification_rate: float = 0.05 / day
    maximum_fixation_rate: float = 0.013 / day
    iron_half_saturation_for_fixation: float = 0.1
    phosphate_half_saturation_for_fixation: float = 0.8
    light_saturation_for_fixation: float = 50.0


@dataclass
class SimpleIron:  # iron/simple_iron.jl:9-13
    excess_scavenging_enhancement: float = 1000.0
    maximum_ligand_concentration: float = 0.6
    dissolved_ligand_ratio: float = 0.09


@dataclass
class Oxygen:  # oxygen.jl:21-24
    ratio_for_respiration: float = 133 / 122
    ratio_for_nitrification: float = 32 / 122


class PISCESModel:
    """The underlying biogeochemistry `PISCES{…}` (PISCES.jl:53-92)."""

    # the two day lengths are host-evaluated from (clock.time, latitude) at every launch: a captured CUDA graph would freeze them
    clock_dependent_parameters = True

    def __init__(self, grid, phytoplankton, zooplankton, dissolved_organic_matter, particulate_organic_matter, nitrogen,
                 iron, oxygen, first_anoxia_threshold, second_anoxia_threshold, nitrogen_redfield_ratio,
                 phosphate_redfield_ratio, mixed_layer_shear, background_shear, latitude, day_length, mixed_layer_depth,
                 euphotic_depth, silicate_climatology, mean_mixed_layer_vertical_diffusivity, mean_mixed_layer_light,
                 carbon_chemistry, calcite_saturation, sinking_velocities):
        self.grid = grid
        self.phytoplankton, self.zooplankton = phytoplankton, zooplankton
        self.dissolved_organic_matter, self.particulate_organic_matter = dissolved_organic_matter, particulate_organic_matter
        self.nitrogen, self.iron, self.oxygen = nitrogen, iron, oxygen
        self.first_anoxia_threshold, self.second_anoxia_threshold = first_anoxia_threshold, second_anoxia_threshold
        self.nitrogen_redfield_ratio, self.phosphate_redfield_ratio = nitrogen_redfield_ratio, phosphate_redfield_ratio
        self.mixed_layer_shear, self.background_shear = mixed_layer_shear, background_shear
        self.latitude, self.day_length = latitude, day_length
        self.mixed_layer_depth, self.euphotic_depth = mixed_layer_depth, euphotic_depth
        self.silicate_climatology = silicate_climatology
        self.mean_mixed_layer_vertical_diffusivity = mean_mixed_layer_vertical_diffusivity
        self.mean_mixed_layer_light = mean_mixed_layer_light
        self.carbon_chemistry, self.calcite_saturation = carbon_chemistry, calcite_saturation
        # [H⁺] of every cell kept between stages: the Ω solve of the next stage starts from it (1–2 Newton steps)
        self.warm_start_carbonate_solve, self.carbonate_state = True, None
        self.sinking_velocities = sinking_velocities

    # ---- plugin surface -------------------------------------------------------------------------------
    def required_biogeochemical_tracers(self):
        return TRACERS

    def required_biogeochemical_auxiliary_fields(self):
        return ("zₘₓₗ", "zₑᵤ", "Si′", "Ω", "κ", "mixed_layer_PAR", "wPOC", "wGOC", "PAR", "PAR₁", "PAR₂", "PAR₃")  # PISCES.jl:107-108

    def biogeochemical_auxiliary_fields(self):  # PISCES.jl:110-118
        return {"zₘₓₗ": self.mixed_layer_depth, "zₑᵤ": self.euphotic_depth, "Si′": self.silicate_climatology,
                "Ω": self.calcite_saturation, "κ": self.mean_mixed_layer_vertical_diffusivity,
                "mixed_layer_PAR": self.mean_mixed_layer_light, "wPOC": self.sinking_velocities["POC"],
                "wGOC": self.sinking_velocities["GOC"]}

    def biogeochemical_drift_velocity(self, name):  # two_size_class.jl:100-107
        if name in ("POC", "SFe"):
            return self.sinking_velocities["POC"]
        if name in ("GOC", "BFe", "PSi", "CaCO₃"):
            return self.sinking_velocities["GOC"]
        return None

    def chlorophyll(self, model):  # coupling_utils.jl:7
        return model.tracers["PChl"], model.tracers["DChl"], 1.0

    def conserved_tracers(self, ntuple=False):
        """coupling_utils.jl:11-33 — applied in this order: carbon, iron, phosphate, silicon, nitrogen."""
        carbon = ("P", "D", "Z", "M", "DOC", "POC", "GOC", "DIC", "CaCO₃")
        iron = {"tracers": ("PFe", "DFe", "Z", "M", "SFe", "BFe", "Fe"),
                "scalefactors": (1, 1, self.zooplankton.micro.iron_ratio, self.zooplankton.meso.iron_ratio, 1, 1, 1)}
        tP = self.phosphate_redfield_ratio
        phosphate = {"tracers": ("P", "D", "Z", "M", "DOC", "POC", "GOC", "PO₄"), "scalefactors": (tP,) * 7 + (1,)}
        silicon = ("DSi", "Si", "PSi")
        tN = self.nitrogen_redfield_ratio
        nitrogen = {"tracers": ("NH₄", "NO₃", "P", "D", "Z", "M", "DOC", "POC", "GOC"), "scalefactors": (1, 1) + (tN,) * 7}
        if ntuple:
            return {"carbon": carbon, "iron": iron, "phosphate": phosphate, "silicon": silicon, "nitrogen": nitrogen}
        return (carbon, iron, phosphate, silicon, nitrogen)

    # ---- C parameter block ------------------------------------------------------------------------------
    def c_params(self, time: float = 0.0) -> _lib.obm_pisces_params:
        p = _lib.obm_pisces_params()
        ph, zo = self.phytoplankton, self.zooplankton
        ph.nano.fill(p.nano)
        ph.diatoms.fill(p.diatoms)
        p.base_rain_ratio = ph.base_rain_ratio
        zo.micro.fill(p.micro)
        zo.meso.fill(p.meso)
        for k in ("microzooplankton_bacteria_concentration", "mesozooplankton_bacteria_concentration",
                  "maximum_bacteria_concentration", "bacteria_concentration_depth_exponent",
                  "doc_half_saturation_for_bacterial_activity", "nitrate_half_saturation_for_bacterial_activity",
                  "ammonia_half_saturation_for_bacterial_activity", "phosphate_half_saturation_for_bacterial_activity",
                  "iron_half_saturation_for_bacterial_activity"):
            setattr(p, k, float(getattr(zo, k)))
        dom, pom = self.dissolved_organic_matter, self.particulate_organic_matter
        p.dom_remineralisation_rate = dom.remineralisation_rate
        p.dom_reference_bacteria_concentration = dom.reference_bacteria_concentration
        p.dom_temperature_sensitivity = dom.temperature_sensitivity
        for n in range(5):
            p.dom_aggregation_parameters[n] = dom.aggregation_parameters[n]
        p.pom_temperature_sensitivity = pom.temperature_sensitivity
        p.pom_base_breakdown_rate = pom.base_breakdown_rate
        for n in range(4):
            p.pom_aggregation_parameters[n] = pom.aggregation_parameters[n]
        for k in ("minimum_iron_scavenging_rate", "load_specific_iron_scavenging_rate", "bacterial_iron_uptake_efficiency",
                  "small_fraction_of_bacterially_consumed_iron", "large_fraction_of_bacterially_consumed_iron",
                  "base_liable_silicate_fraction", "fast_dissolution_rate_of_silicate", "slow_dissolution_rate_of_silicate",
                  "base_calcite_dissolution_rate", "calcite_dissolution_exponent", "maximum_iron_ratio_in_bacteria",
                  "iron_half_saturation_for_bacteria", "maximum_bacterial_growth_rate"):
            setattr(p, k, float(getattr(pom, k)))
        for k in ("maximum_nitrification_rate", "maximum_fixation_rate", "iron_half_saturation_for_fixation",
                  "phosphate_half_saturation_for_fixation", "light_saturation_for_fixation"):
            setattr(p, k, float(getattr(self.nitrogen, k)))
        for k in ("excess_scavenging_enhancement", "maximum_ligand_concentration", "dissolved_ligand_ratio"):
            setattr(p, k, float(getattr(self.iron, k)))
        p.ratio_for_respiration = self.oxygen.ratio_for_respiration
        p.ratio_for_nitrification = self.oxygen.ratio_for_nitrification
        for k in ("first_anoxia_threshold", "second_anoxia_threshold", "nitrogen_redfield_ratio", "phosphate_redfield_ratio",
                  "mixed_layer_shear", "background_shear"):
            setattr(p, k, float(getattr(self, k)))
        # ModelLatitude: the per-row values travel in `row_table`; the scalar members are then ignored by the kernel
        φ = 0.0 if isinstance(self.latitude, ModelLatitude) else float(self.latitude.latitude)
        p.latitude = φ
        # the reference's two call orders (growth_rate.jl:30 swapped, :143 correct) — SURVEY App. A bug 1
        p.day_length_growth = float(self.day_length(φ, time))
        p.day_length_chlorophyll = float(self.day_length(time, φ))
        p.silicate_climatology = float(self.silicate_climatology)
        return p

    def row_table(self, grid, time: float) -> torch.Tensor:
        """ModelLatitude: latitude, day_length(φ, t) (the reference's swapped call, growth_rate.jl:29-30) and
        day_length(t, φ) (:141-143) of every interior row j as a device array [3][Ny] — host-evaluated per launch like
        their scalar counterparts in `c_params`."""
        key = (float(time), id(grid), str(grid.device))
        cached = getattr(self, "_row_cache", None)
        if cached is not None and cached[0] == key:  # the slabs of one host-staged stage share the clock
            return cached[1]
        φs = [float(v) for v in grid.latitude_centers]
        rows = [φs, [float(self.day_length(φ, time)) for φ in φs], [float(self.day_length(time, φ)) for φ in φs]]
        table = torch.tensor(rows, dtype=torch.float64).to(grid.device)
        self._row_cache = (key, table)
        return table

    def c_fields(self, aux: dict) -> _lib.obm_pisces_fields:
        f = _lib.obm_pisces_fields()
        f.PAR1, f.PAR2, f.PAR3, f.PAR = aux["PAR₁"].ptr, aux["PAR₂"].ptr, aux["PAR₃"].ptr, aux["PAR"].ptr
        f.Omega = aux["Ω"].ptr
        f.wPOC, f.wGOC = aux["wPOC"].ptr, aux["wGOC"].ptr
        f.mixed_layer_depth_xy = aux["zₘₓₗ"].ptr
        f.euphotic_depth_xy = aux["zₑᵤ"].ptr
        f.mean_mixed_layer_vertical_diffusivity_xy = aux["κ"].ptr
        f.mean_mixed_layer_light_xy = aux["mixed_layer_PAR"].ptr
        return f

    # ---- update_biogeochemical_state!(model, bgc::PISCES) — update_state.jl:1-17 --------------------
    def _carbonate_state(self, model):
        if self.carbonate_state is None and self.warm_start_carbonate_solve:
            self.carbonate_state = CenterField(model.grid, "[H⁺]")  # zero-filled ⇒ first call starts from pH 8
        return self.carbonate_state

    def calcite_saturation_arguments(self, model):
        """What `apply_scalers(…, calcite=…)` needs to solve Ω in the negative-scaling launch."""
        t = model.tracers
        return (self.carbon_chemistry, t["T"], t["S"], t["DIC"], t["Alk"], t["Si"], self.calcite_saturation,
                self._carbonate_state(model))

    def column_light_state(self, model):
        """What the multi-band PAR launch needs to also leave zₑᵤ and the mixed-layer mean PAR (update_state.jl:7,11);
        None when either is a prescribed `ConstantField` (then nothing may overwrite it)."""
        if self.euphotic_depth.constant or self.mean_mixed_layer_light.constant:
            return None
        return (self.mixed_layer_depth, 1 / 1000, self.euphotic_depth, self.mean_mixed_layer_light)

    def update_biogeochemical_state(self, model, stream: Optional[int] = None, calcite_saturation_done: bool = False,
                                    light_state_done: bool = False):
        PAR = model.biogeochemistry.light_attenuation.biogeochemical_auxiliary_fields()["PAR"]
        if not light_state_done and not self.euphotic_depth.constant:
            compute_euphotic_depth(self.euphotic_depth, PAR, stream=stream)
        kappa = getattr(model, "vertical_diffusivity", None)  # closure = nothing ⇒ pre-set κ̄ is kept (:64-67)
        if kappa is not None and not self.mean_mixed_layer_vertical_diffusivity.constant:
            compute_mixed_layer_mean(self.mean_mixed_layer_vertical_diffusivity, self.mixed_layer_depth, kappa, model.grid, stream)
        if not light_state_done and not self.mean_mixed_layer_light.constant:
            compute_mixed_layer_mean(self.mean_mixed_layer_light, self.mixed_layer_depth, PAR, model.grid, stream)
        if calcite_saturation_done:  # solved in the negative-scaling launch (Biogeochemistry.update_biogeochemical_state)
            return
        t = model.tracers
        self.carbon_chemistry.calcite_saturation(model.grid, t["T"], t["S"], t["DIC"], t["Alk"], t["Si"],
                                                 self.calcite_saturation, stream, state=self._carbonate_state(model))

    # ---- fused tendencies ------------------------------------------------------------------------------------
    def compute_tendencies(self, grid, tracers, auxiliary_fields, G, accumulate=True, stream=None, time=0.0):
        require_cuda(*[tracers[n] for n in TRACERS])
        p = self.c_params(time)
        f = self.c_fields(auxiliary_fields)
        cg = grid.c_grid()
        tptr = _lib.pointer_table([tracers[n].ptr for n in TRACERS])
        gptr = _lib.pointer_table([G[n].ptr if (n in G and G[n] is not None and n not in ("T", "S")) else None for n in TRACERS])
        s = stream if stream is not None else current_stream_ptr(grid.device)
        if isinstance(self.latitude, ModelLatitude):
            rows = self.row_table(grid, time)  # (kept alive by the local name until the launch is enqueued; same stream order)
            rc = _lib.load().obm_pisces_tendencies_rows(C.byref(cg), C.byref(p), C.c_void_p(rows.data_ptr()), tptr, C.byref(f),
                                                        gptr, 1 if accumulate else 0, s)
            _lib.check(rc, "obm_pisces_tendencies_rows")
            self._row_table = rows  # the launch reads it asynchronously: keep it until the next one replaces it
            return
        rc = _lib.load().obm_pisces_tendencies(C.byref(cg), C.byref(p), tptr, C.byref(f), gptr, 1 if accumulate else 0, s)
        _lib.check(rc, "obm_pisces_tendencies")

    AUXILIARY_DEFAULTS = {"PAR₁": 0.0, "PAR₂": 0.0, "PAR₃": 0.0, "PAR": None, "Ω": 1.0, "zₘₓₗ": -10.0, "zₑᵤ": -10.0, "κ": 1.0,
                          "mixed_layer_PAR": 0.0, "wPOC": 0.0, "wGOC": 0.0}

    def __call__(self, name: str, *, z: float = -5.0, time: float = 0.0, device="cuda", **state):
        """The per-tracer form of the plugin API, `bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)`
        (PISCES.jl:120-123; the continuous form `bgc(Val(name), x, y, z, t, fields...)` of
        docs/src/model_implementation.md:34-75): the tendency of ONE tracer at the given state.  `state` holds tracer
        values (absent ones are 0) and the auxiliary fields of `required_biogeochemical_auxiliary_fields` by their
        reference names — PAR₁, PAR₂, PAR₃, PAR (default: their sum), Ω, zₘₓₗ, zₑᵤ, κ, mixed_layer_PAR, wPOC, wGOC (the
        cell-centre sinking speeds) — as scalars or as arrays of states evaluated side by side.  The fused kernel runs on
        a row of boxes centred on depth `z` at model time `time`; a float is returned for scalar input, else a device
        tensor.  NPD's counterpart: `NutrientsPlanktonDetritus.__call__`."""
        from .box_model import BoxModelGrid
        if name not in TRACERS:
            raise KeyError(f"{name} is not a tracer of PISCES {TRACERS}")
        unknown = set(state) - set(TRACERS) - set(self.AUXILIARY_DEFAULTS)
        if unknown:
            raise KeyError(f"unknown fields {sorted(unknown)}; PISCES carries {TRACERS} and reads {tuple(self.AUXILIARY_DEFAULTS)}")
        flat = lambda v: torch.as_tensor(v, dtype=torch.float64).reshape(-1)  # noqa: E731
        vals = {n: flat(state.get(n, 0.0)) for n in TRACERS}
        aux = {n: flat(state.get(n, d if d is not None else 0.0)) for n, d in self.AUXILIARY_DEFAULTS.items()}
        if "PAR" not in state:
            aux["PAR"] = aux["PAR₁"] + aux["PAR₂"] + aux["PAR₃"]
        n = max(v.numel() for v in list(vals.values()) + list(aux.values()))
        grid = BoxModelGrid(n, device=device, z=z)
        row = lambda v: v.expand(n).reshape(1, 1, n)  # noqa: E731
        tr = {k: CenterField(grid, k).set(row(v)) for k, v in vals.items()}
        fields = {}
        for k, v in aux.items():
            if k in ("zₘₓₗ", "zₑᵤ", "κ", "mixed_layer_PAR"):
                fields[k] = Field2D(grid, k).set(row(v))
            elif k in ("wPOC", "wGOC"):  # ℑzᵃᵃᶜ(w) of two equal faces is the cell-centre speed
                f = ZFaceField(grid, k)
                f.face_interior[...] = row(v).to(f.data.device).expand(2, 1, n)
                fields[k] = f
            else:
                fields[k] = CenterField(grid, k).set(row(v))
        G = {k: CenterField(grid, "G" + k) for k in TRACERS}
        self.compute_tendencies(grid, tr, fields, G, accumulate=False, time=time)
        out = G[name].interior.reshape(-1)
        scalar = n == 1 and not any(torch.is_tensor(v) for v in state.values())
        return out.item() if scalar else out

    def summary(self):
        return "PISCES biogeochemical model (24 tracers)"


def PISCES(grid: RectilinearGrid, phytoplankton=None, zooplankton=None, dissolved_organic_matter=None,
           particulate_organic_matter=None, nitrogen=None, iron=None, oxygen=None, first_anoxia_threshold=6.0,
           second_anoxia_threshold=1.0, nitrogen_redfield_ratio=16 / 122, phosphate_redfield_ratio=1 / 122,
           mixed_layer_shear=1.0, background_shear=0.01, latitude=None, day_length=None, mixed_layer_depth=None,
           euphotic_depth=None, silicate_climatology=7.5, mean_mixed_layer_vertical_diffusivity=None,
           mean_mixed_layer_light=None, carbon_chemistry=None, calcite_saturation=None,
           surface_photosynthetically_active_radiation=default_surface_PAR, light_attenuation=None, sinking_speeds=None,
           open_bottom=True, scale_negatives=False, invalid_fill_value=float("nan"), sediment=None, particles=None,
           modifiers=None):
    """`PISCES(; grid, …)` — PISCES.jl:288-408."""
    from .biogeochemistry import Biogeochemistry
    from .negative_tracers import ScaleNegativeTracers

    mixed_layer_depth = mixed_layer_depth if mixed_layer_depth is not None else Field2D(grid, "zₘₓₗ")
    euphotic_depth = euphotic_depth if euphotic_depth is not None else Field2D(grid, "zₑᵤ")
    if mean_mixed_layer_vertical_diffusivity is None:
        mean_mixed_layer_vertical_diffusivity = Field2D(grid, "κ̄", fill=1.0)  # `set!(…, 1)` PISCES.jl:373-375
    mean_mixed_layer_light = mean_mixed_layer_light if mean_mixed_layer_light is not None else Field2D(grid, "PAR̄")
    calcite_saturation = calcite_saturation if calcite_saturation is not None else CenterField(grid, "Ω")
    if light_attenuation is None:
        light_attenuation = MultiBandPhotosyntheticallyActiveRadiation(
            grid=grid, surface_PAR=surface_photosynthetically_active_radiation)
    if sinking_speeds is None:
        sinking_speeds = {"POC": 2 / day, "GOC": DepthDependantSinkingSpeed()}
    velocities = {}
    for name, w in sinking_speeds.items():  # setup_velocity_fields, sinking_velocity_fields.jl:10-35
        if isinstance(w, (int, float)):
            f = ZFaceField(grid, "w" + name)
            for k in range(grid.Nz):  # faces 1…Nz; the top face (Nz + 1) stays 0
                f.face_interior[k] = -w * (1.0 if open_bottom else (1 - math.exp((1 - (k + 1)) / 2)))
            velocities[name] = f
        elif isinstance(w, DepthDependantSinkingSpeed):
            velocities[name] = w.face_field(grid, mixed_layer_depth, euphotic_depth)  # `compute!(w)` once, at setup
        else:
            velocities[name] = w
    # PISCES.jl:360-367: a grid with its own latitude overrides a prescribed one (the reference warns and then stores
    # `nothing`, which its kernels cannot call — the stated intent, the grid's latitude, is what is built here); a
    # RectilinearGrid has none to offer
    from .grids import LatitudeLongitudeGrid
    latitude = latitude or PrescribedLatitude(45.0)
    if isinstance(latitude, PrescribedLatitude) and isinstance(grid, LatitudeLongitudeGrid):
        import warnings
        φ = grid.latitude_centers
        warnings.warn(f"A latitude of {latitude} was given but the grid has its own latitude ({min(φ)}, {max(φ)}) so the "
                      "prescribed value is ignored")
        latitude = ModelLatitude()
    elif isinstance(latitude, ModelLatitude) and not isinstance(grid, LatitudeLongitudeGrid):
        raise ValueError("You must prescribe a latitude when using a `RectilinearGrid`")
    underlying = PISCESModel(
        grid, phytoplankton or MixedMondoNanoAndDiatoms(), zooplankton or MicroAndMesoZooplankton(),
        dissolved_organic_matter or DissolvedOrganicCarbon(), particulate_organic_matter or TwoCompartmentCarbonIronParticles(),
        nitrogen or NitrateAmmonia(), iron or SimpleIron(), oxygen or Oxygen(), first_anoxia_threshold,
        second_anoxia_threshold, nitrogen_redfield_ratio, phosphate_redfield_ratio, mixed_layer_shear, background_shear,
        latitude, day_length or CBMDayLength(), mixed_layer_depth, euphotic_depth,
        silicate_climatology, mean_mixed_layer_vertical_diffusivity, mean_mixed_layer_light,
        carbon_chemistry or CarbonChemistry(newton_iterations=8), calcite_saturation, velocities)
    if scale_negatives:
        scalers = ScaleNegativeTracers.from_biogeochemistry(underlying, grid, invalid_fill_value=invalid_fill_value)
        if modifiers is None:
            modifiers = scalers
        elif isinstance(modifiers, tuple):
            modifiers = (*modifiers, *scalers)
        else:
            modifiers = (modifiers, *scalers)
    return Biogeochemistry(underlying, light_attenuation=light_attenuation, sediment=sediment, particles=particles,
                           modifiers=modifiers)


# ---- synthetic state of BASELINE config C4 -----------------------------------------------
PISCES_INITIAL_VALUES = {  # test/test_PISCES.jl:8-16
    "P": 0.5, "PChl": 0.02, "PFe": 0.005, "D": 0.1, "DChl": 0.004, "DFe": 0.001, "DSi": 0.01, "Z": 0.1, "M": 0.7,
    "DOC": 2.1, "POC": 7.8, "SFe": 0.206, "GOC": 38.0, "BFe": 1.1, "PSi": 0.1, "CaCO₃": 10.0 ** -10, "NO₃": 2.3,
    "NH₄": 0.9, "PO₄": 0.6, "Fe": 0.13, "Si": 8.5, "DIC": 2205.0, "Alk": 2566.0, "O₂": 317.0, "T": 10.0, "S": 35.0}


def synthetic_range(name: str):
    """each tracer = PISCES_INITIAL_VALUES × e^{0.5(2u−1)} (log-uniform); T ∈ [2, 28], S ∈ [33, 37]; DIC/Alk ±2 %"""
    if name == "T":
        return (2.0, 28.0, False)
    if name == "S":
        return (33.0, 37.0, False)
    v = PISCES_INITIAL_VALUES[name]
    if name in ("DIC", "Alk"):
        return (0.98 * v, 1.02 * v, False)
    if name == "CaCO₃":
        return (0.01, 1.0, True)  # 1e-10 of the test state would make the calcite terms invisible
    return (v * math.exp(-0.5), v * math.exp(0.5), True)


def fill_synthetic_auxiliary(bgc, model):
    """zₘₓₗ ∈ [−150, −10], κ̄ ∈ [1e-4, 1e-2] (log), wGOC recomputed from DepthDependantSinkingSpeed with them."""
    from . import synthetic
    u = bgc.underlying_biogeochemistry
    synthetic.fill_torch(u.mixed_layer_depth, "zₘₓₗ", -150.0, -10.0)
    synthetic.fill_torch(u.mean_mixed_layer_vertical_diffusivity, "κ̄", 1e-4, 1e-2, log=True)
    u.euphotic_depth.data.fill_(-60.0)
    u.sinking_velocities["GOC"] = DepthDependantSinkingSpeed().face_field(model.grid, u.mixed_layer_depth, u.euphotic_depth)
