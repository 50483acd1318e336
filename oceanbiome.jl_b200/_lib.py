"""ctypes binding of libobm_b200.so — a line-for-line mirror of include/obm_b200.h.

The same C ABI is what the Julia glue `ccall`s; parity through this binding
therefore certifies the ABI itself.  There is NO CPU fallback: if the CUDA library is missing the
import of any compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libobm_b200.so")

c_double_p = C.POINTER(C.c_double)
c_double_pp = C.POINTER(c_double_p)

OBM_MAX_BANDS = 8
OBM_MAX_SCALE_TRACERS = 32
OBM_MAX_SCALE_GROUPS = 8
OBM_MAX_GROUP_SIZE = 16
OBM_NPD_MAX_TRACERS = 32
OBM_NPD_MAX_VARIED = 16

# enums (include/obm_b200.h)
NUT_NUTRIENT, NUT_NITRATE_AMMONIA, NUT_NITRATE_AMMONIA_IRON = 0, 1, 2
DET_NONE, DET_DETRITUS, DET_TWO_PARTICLE, DET_VARIABLE_REDFIELD = 0, 1, 2, 3
LIGHT_MONDO, LIGHT_ANALYTICAL = 0, 1
LINEAR, QUADRATIC = 0, 1
CC_FCO2, CC_PCO2, CC_PH_FREE, CC_PH_TOTAL, CC_PH_SEAWATER, CC_CO3, CC_OMEGA_CALCITE = range(7)


class obm_grid(C.Structure):
    _fields_ = [
        ("Nx", C.c_int32), ("Ny", C.c_int32), ("Nz", C.c_int32),
        ("Hx", C.c_int32), ("Hy", C.c_int32), ("Hz", C.c_int32),
        ("i0", C.c_int32), ("i1", C.c_int32), ("j0", C.c_int32), ("j1", C.c_int32),
        ("zc", C.c_void_p), ("zf", C.c_void_p),
        ("bottom_indices_xy", C.c_void_p),  # nullable: 1-based bottom-most active cell per column (immersed boundary)
    ]


_NPD_DOUBLES = [
    "nitrate_half_saturation", "ammonia_half_saturation", "iron_half_saturation",
    "nitrate_ammonia_inhibition", "light_half_saturation", "phytoplankton_maximum_growth_rate",
    "iron_ratio", "phytoplankton_exudation_fraction", "ammonia_fraction_of_exudate",
    "temperature_coefficient", "phytoplankton_mortality_rate", "zooplankton_mortality_rate",
    "zooplankton_excretion_rate", "phytoplankton_solid_waste_fraction",
    "excretion_inorganic_fraction", "preference_for_phytoplankton", "maximum_grazing_rate",
    "grazing_half_saturation", "zooplankton_assimilation_fraction",
    "zooplankton_calcite_dissolution", "redfield_ratio", "carbon_calcite_ratio",
    "zooplankton_gut_calcite_dissolution", "phytoplankton_chlorophyll_ratio",
    "nitrification_rate",
    "remineralisation_inorganic_fraction", "small_remineralisation_rate",
    "large_remineralisation_rate", "dissolved_remineralisation_rate",
    "small_solid_waste_fraction", "detritus_redfield_ratio",
    "remineralisation_rate", "small_particle_fraction",
    "respiration_oxygen_nitrogen_ratio", "nitrification_oxygen_nitrogen_ratio",
]


class obm_npd_params(C.Structure):
    _fields_ = [
        ("nutrients", C.c_int32), ("detritus", C.c_int32), ("carbonate_replicates", C.c_int32),
        ("oxygen", C.c_int32), ("light_limitation", C.c_int32),
        ("phytoplankton_mortality_formulation", C.c_int32),
        ("grazing_concentration_formulation", C.c_int32),
        ("has_temperature_coefficient", C.c_int32),
    ] + [(n, C.c_double) for n in _NPD_DOUBLES]


class obm_twoband_params(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "water_red_attenuation", "water_blue_attenuation", "chlorophyll_red_attenuation",
        "chlorophyll_blue_attenuation", "chlorophyll_red_exponent", "chlorophyll_blue_exponent",
        "pigment_ratio", "phytoplankton_chlorophyll_ratio")]


class obm_multiband_params(C.Structure):
    _fields_ = [
        ("nbands", C.c_int32), ("_pad", C.c_int32),
        ("water_attenuation_coefficient", C.c_double * OBM_MAX_BANDS),
        ("chlorophyll_exponent", C.c_double * OBM_MAX_BANDS),
        ("chlorophyll_attenuation_coefficient", C.c_double * OBM_MAX_BANDS),
        ("surface_PAR_division", C.c_double * OBM_MAX_BANDS),
    ]


class obm_carbchem_params(C.Structure):
    _fields_ = [("newton_iterations", C.c_int32), ("_pad", C.c_int32), ("initial_pH_guess", C.c_double)]


class obm_scale_group(C.Structure):
    _fields_ = [
        ("n", C.c_int32),
        ("index", C.c_int32 * OBM_MAX_GROUP_SIZE),
        ("scalefactor", C.c_double * OBM_MAX_GROUP_SIZE),
    ]


_PHYTO_DOUBLES = [
    "base_growth_rate", "temperature_sensitivity", "dark_tolerance", "initial_slope_of_PI_curve",
    "low_light_adaptation", "basal_respiration_rate", "reference_growth_rate",
    "minimum_ammonium_half_saturation", "minimum_nitrate_half_saturation", "minimum_phosphate_half_saturation",
    "optimal_iron_quota", "minimum_silicate_half_saturation", "silicate_half_saturation_parameter",
    "exudated_fraction", "blue_light_absorption", "green_light_absorption", "red_light_absorption",
    "mortality_half_saturation", "linear_mortality_rate", "base_quadratic_mortality", "maximum_quadratic_mortality",
    "minimum_chlorophyll_ratio", "maximum_chlorophyll_ratio", "maximum_iron_ratio", "silicate_half_saturation",
    "enhanced_silicate_half_saturation", "optimal_silicate_ratio", "half_saturation_for_iron_uptake",
    "threshold_for_size_dependency", "size_ratio",
]


class obm_pisces_phyto(C.Structure):
    _fields_ = [("growth_rate_kind", C.c_int32), ("silicate_limited", C.c_int32)] + [(n, C.c_double) for n in _PHYTO_DOUBLES]


class obm_pisces_zoo(C.Structure):
    _fields_ = ([("temperature_sensitivity", C.c_double), ("maximum_grazing_rate", C.c_double),
                 ("food_preferences", C.c_double * 4)]
                + [(n, C.c_double) for n in (
                    "food_threshold_concentration", "specific_food_threshold_concentration", "grazing_half_saturation",
                    "maximum_flux_feeding_rate", "iron_ratio", "minimum_growth_efficiency", "non_assimilated_fraction",
                    "mortality_half_saturation", "quadratic_mortality", "linear_mortality",
                    "dissolved_excretion_fraction", "undissolved_calcite_fraction")])


class obm_pisces_params(C.Structure):
    _fields_ = (
        [("nano", obm_pisces_phyto), ("diatoms", obm_pisces_phyto), ("base_rain_ratio", C.c_double),
         ("micro", obm_pisces_zoo), ("meso", obm_pisces_zoo)]
        + [(n, C.c_double) for n in (
            "microzooplankton_bacteria_concentration", "mesozooplankton_bacteria_concentration",
            "maximum_bacteria_concentration", "bacteria_concentration_depth_exponent",
            "doc_half_saturation_for_bacterial_activity", "nitrate_half_saturation_for_bacterial_activity",
            "ammonia_half_saturation_for_bacterial_activity", "phosphate_half_saturation_for_bacterial_activity",
            "iron_half_saturation_for_bacterial_activity",
            "dom_remineralisation_rate", "dom_reference_bacteria_concentration", "dom_temperature_sensitivity")]
        + [("dom_aggregation_parameters", C.c_double * 5),
           ("pom_temperature_sensitivity", C.c_double), ("pom_base_breakdown_rate", C.c_double),
           ("pom_aggregation_parameters", C.c_double * 4)]
        + [(n, C.c_double) for n in (
            "minimum_iron_scavenging_rate", "load_specific_iron_scavenging_rate", "bacterial_iron_uptake_efficiency",
            "small_fraction_of_bacterially_consumed_iron", "large_fraction_of_bacterially_consumed_iron",
            "base_liable_silicate_fraction", "fast_dissolution_rate_of_silicate", "slow_dissolution_rate_of_silicate",
            "base_calcite_dissolution_rate", "calcite_dissolution_exponent", "maximum_iron_ratio_in_bacteria",
            "iron_half_saturation_for_bacteria", "maximum_bacterial_growth_rate",
            "maximum_nitrification_rate", "maximum_fixation_rate", "iron_half_saturation_for_fixation",
            "phosphate_half_saturation_for_fixation", "light_saturation_for_fixation",
            "excess_scavenging_enhancement", "maximum_ligand_concentration", "dissolved_ligand_ratio",
            "ratio_for_respiration", "ratio_for_nitrification",
            "first_anoxia_threshold", "second_anoxia_threshold", "nitrogen_redfield_ratio", "phosphate_redfield_ratio",
            "mixed_layer_shear", "background_shear",
            "latitude", "day_length_growth", "day_length_chlorophyll", "silicate_climatology")])


class obm_pisces_fields(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "PAR1", "PAR2", "PAR3", "PAR", "Omega", "wPOC", "wGOC", "mixed_layer_depth_xy", "euphotic_depth_xy",
        "mean_mixed_layer_vertical_diffusivity_xy", "mean_mixed_layer_light_xy")]


OBM_PISCES_NTRACERS = 26

SED_INSTANT_REMINERALISATION, SED_SIMPLE_MULTI_G = 0, 1
ADV_UPWIND1, ADV_CENTERED2, ADV_UPWIND3, ADV_WENO5 = 0, 1, 2, 3
OBM_MAX_SINKING_TRACERS = 8
TS_AB2, TS_RK3 = 0, 1
OBM_SED_MAX_SINKING, OBM_SED_MAX_POOLS, OBM_SED_MAX_COUPLED = 4, 6, 4


class obm_sediment_params(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("model", "carbon", "nsinking_nitrogen", "nsinking_carbon", "advection", "timestepper")]
                + [(n, C.c_double) for n in (
                    "burial_efficiency_constant1", "burial_efficiency_constant2", "burial_efficiency_half_saturation",
                    "sinking_redfield", "fast_decay_rate", "slow_decay_rate", "fast_redfield", "slow_redfield",
                    "fast_fraction", "slow_fraction", "refactory_fraction", "sedimentation_rate", "anoxia_half_saturation")]
                + [("nitrate_oxidation_params", C.c_double * 6), ("denitrification_params", C.c_double * 6),
                   ("anoxic_params", C.c_double * 6), ("solid_dep_params", C.c_double * 4)])


class obm_sediment_fields(C.Structure):
    _fields_ = [("bottom_indices_xy", C.c_void_p), ("NO3", C.c_void_p), ("NH4", C.c_void_p), ("O2", C.c_void_p),
                ("sinking", C.c_void_p * (2 * OBM_SED_MAX_SINKING)), ("sinking_w", C.c_void_p * (2 * OBM_SED_MAX_SINKING)),
                ("pools", C.c_void_p * OBM_SED_MAX_POOLS), ("Gn", C.c_void_p * OBM_SED_MAX_POOLS),
                ("Gm", C.c_void_p * OBM_SED_MAX_POOLS), ("tracked_xy", C.c_void_p * (3 + 2 * OBM_SED_MAX_SINKING)),
                ("G_coupled", C.c_void_p * OBM_SED_MAX_COUPLED)]


class obm_gas_exchange_params(C.Structure):
    _fields_ = ([(n, C.c_int32) for n in ("water_kind", "air_kind", "solubility_kind", "k660_order",
                                          "use_silicate_phosphate", "_pad")]
                + [("k660", C.c_double * 4), ("schmidt", C.c_double * 5), ("w92", C.c_double * 6)]
                + [(n, C.c_double) for n in ("air_concentration", "wind_speed", "silicate", "phosphate")]
                + [("carbon_chemistry", obm_carbchem_params)])


OBM_GE_WATER_TRACER, OBM_GE_WATER_PCO2 = 0, 1
OBM_GE_AIR_PLAIN, OBM_GE_AIR_WANNINKHOF92 = 0, 1
OBM_GE_SOLUBILITY_ONE, OBM_GE_SOLUBILITY_K0_RHO = 0, 1

KELP_DOUBLES = (
    "lower_optimal", "upper_optimal", "lower_gradient", "upper_gradient",
    "growth_rate_adjustment", "photosynthetic_efficiency", "minimum_carbon_reserve", "structural_carbon",
    "exudation", "erosion_exponent", "base_erosion_rate", "saturation_irradiance",
    "structural_dry_weight_per_area", "structural_dry_to_wet_weight", "carbon_reserve_per_carbon",
    "nitrogen_reserve_per_nitrogen", "minimum_nitrogen_reserve", "maximum_nitrogen_reserve",
    "growth_adjustment_2", "growth_adjustment_1", "maximum_specific_growth_rate", "structural_nitrogen",
    "photosynthesis_at_ref_temp_1", "photosynthesis_at_ref_temp_2", "photosynthesis_ref_temp_1",
    "photosynthesis_ref_temp_2", "photoperiod_1", "photoperiod_2",
    "respiration_at_ref_temp_1", "respiration_at_ref_temp_2", "respiration_ref_temp_1", "respiration_ref_temp_2",
    "photosynthesis_arrhenius_temp", "photosynthesis_low_temp", "photosynthesis_high_temp",
    "photosynthesis_high_arrhenius_temp", "photosynthesis_low_arrhenius_temp", "respiration_arrhenius_temp",
    "current_speed_for_0p65_uptake", "nitrate_half_saturation", "ammonia_half_saturation",
    "maximum_nitrate_uptake", "maximum_ammonia_uptake", "current_1", "current_2", "current_3",
    "base_activity_respiration_rate", "base_basal_respiration_rate", "exudation_redfield_ratio", "adapted_latitude")
OBM_TOPO_PERIODIC, OBM_TOPO_BOUNDED, OBM_TOPO_FLAT = 0, 1, 2
OBM_KELP_NCOUPLED = 8


class obm_sugar_kelp_params(C.Structure):
    _fields_ = [(n, C.c_double) for n in KELP_DOUBLES] + [("newton_iterations", C.c_int32), ("_pad", C.c_int32)]


class obm_particles(C.Structure):
    _fields_ = ([("n", C.c_int64)] + [(n, C.c_void_p) for n in ("x", "y", "z", "A", "N", "C", "scalefactors")]
                + [(n, C.c_double) for n in ("x0", "dx", "y0", "dy")] + [("topology", C.c_int32 * 3), ("_pad", C.c_int32)])


class obm_kelp_tracers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "v", "w", "T", "NO3", "NH4", "PAR")]


STRUCTS = {
    "obm_sugar_kelp_params": obm_sugar_kelp_params,
    "obm_particles": obm_particles,
    "obm_kelp_tracers": obm_kelp_tracers,
    "obm_gas_exchange_params": obm_gas_exchange_params,
    "obm_sediment_params": obm_sediment_params,
    "obm_sediment_fields": obm_sediment_fields,
    "obm_pisces_phyto": obm_pisces_phyto,
    "obm_pisces_zoo": obm_pisces_zoo,
    "obm_pisces_params": obm_pisces_params,
    "obm_pisces_fields": obm_pisces_fields,
    "obm_grid": obm_grid,
    "obm_npd_params": obm_npd_params,
    "obm_twoband_params": obm_twoband_params,
    "obm_multiband_params": obm_multiband_params,
    "obm_carbchem_params": obm_carbchem_params,
    "obm_scale_group": obm_scale_group,
}

# name → (restype, argtypes); exactly the prototypes of include/obm_b200.h
PROTOTYPES = {
    "obm_npd_tracer_names": (C.c_int, [C.POINTER(obm_npd_params), C.c_void_p]),
    "obm_npd_tendencies": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_npd_params), C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_void_p]),
    "obm_npd_param_index": (C.c_int, [C.c_char_p]),
    "obm_npd_tendencies_ensemble": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_npd_params), C.c_int, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "obm_npd_box_run": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_npd_params), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]),
    "obm_pisces_tendencies": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_pisces_params), C.c_void_p,
                                        C.POINTER(obm_pisces_fields), C.c_void_p, C.c_int, C.c_void_p]),
    "obm_pisces_tendencies_rows": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_pisces_params), C.c_void_p, C.c_void_p,
                                             C.POINTER(obm_pisces_fields), C.c_void_p, C.c_int, C.c_void_p]),
    "obm_par_twoband": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_twoband_params), C.c_void_p, C.c_void_p,
                                  C.c_double, C.c_void_p, C.c_void_p]),
    "obm_par_multiband": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_multiband_params), C.c_void_p, C.c_void_p,
                                    C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "obm_par_multiband_column_state": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_multiband_params), C.c_void_p, C.c_void_p,
                                                 C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "obm_euphotic_depth": (C.c_int, [C.POINTER(obm_grid), C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "obm_mixed_layer_mean": (C.c_int, [C.POINTER(obm_grid), C.c_void_p, C.c_void_p, C.c_double, C.c_void_p,
                                       C.c_void_p]),
    "obm_carbon_chemistry": (C.c_int, [C.c_int64, C.POINTER(obm_carbchem_params)] + [C.c_void_p] * 8
                             + [C.c_int, C.c_void_p, C.c_void_p]),
    "obm_calcite_saturation": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_carbchem_params)] + [C.c_void_p] * 8),
    "obm_scale_negative_tracers": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_int,
                                             C.POINTER(obm_scale_group), C.c_double, C.c_void_p]),
    "obm_scale_negative_tracers_calcite_saturation": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_int,
                                                                C.POINTER(obm_scale_group), C.c_double,
                                                                C.POINTER(obm_carbchem_params)] + [C.c_void_p] * 8),
    "obm_zero_negative_tracers": (C.c_int, [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "obm_kelp_update_tendencies": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_sugar_kelp_params), C.POINTER(obm_particles),
                                             C.POINTER(obm_kelp_tracers), C.c_void_p, C.c_double, C.c_void_p]),
    "obm_kelp_step": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_sugar_kelp_params), C.POINTER(obm_particles),
                                C.POINTER(obm_kelp_tracers), C.c_double, C.c_double, C.c_void_p, C.c_void_p]),
    "obm_kelp_seasonal_limitation": (C.c_double, [C.POINTER(obm_sugar_kelp_params), C.c_double]),
    "obm_sinking_tendencies": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_void_p]),
    "obm_inventory_workspace_bytes": (C.c_int64, [C.c_int]),
    "obm_inventory": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_int, C.POINTER(obm_scale_group),
                                C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "obm_sediment_update_state": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_sediment_params), C.POINTER(obm_sediment_fields),
                                            C.c_double, C.c_double, C.c_void_p]),
    "obm_sediment_update_tendencies": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_sediment_params),
                                                 C.POINTER(obm_sediment_fields), C.c_void_p]),
    "obm_find_bottom_cells": (C.c_int, [C.POINTER(obm_grid), C.c_void_p, C.c_void_p, C.c_void_p]),
    "obm_gas_exchange_flux": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_gas_exchange_params)] + [C.c_void_p] * 12),
    "obm_rk3_substep": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                  C.c_double, C.c_int, C.c_int, C.c_void_p]),
    "obm_npd_tendencies_substep": (C.c_int, [C.POINTER(obm_grid), C.POINTER(obm_npd_params), C.c_int, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_double,
                                             C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "obm_copy_slab": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "obm_copy_slab_sm": (C.c_int, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "obm_fp64_peak_dfma_per_s": (C.c_double, [C.c_void_p, C.c_int, C.c_void_p]),
    "obm_fetch_ceiling_ms": (C.c_double, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "obm_stream_pattern_gbs": (C.c_double, [C.POINTER(obm_grid), C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                            C.c_int, C.c_void_p]),
    "obm_last_error": (C.c_char_p, []),
    "obm_version": (C.c_int, []),
    "obm_sizeof": (C.c_int, [C.c_char_p]),
}

_lib = None


class ObmError(RuntimeError):
    pass


def load(path: str | None = None):
    """Load libobm_b200.so (once).  Raises ImportError — loudly — when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or os.environ.get("OBM_B200_LIB", LIB_PATH)
    if not os.path.exists(path):
        raise ImportError(
            f"libobm_b200.so not found at {path}: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  oceanbiome.jl_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().obm_last_error().decode("utf-8", "replace")
        raise ObmError(f"{what} failed with code {rc}: {msg}")


def pointer_table(ptrs):
    """Host array of device pointers (`const double* const*`)."""
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(p) if p else None for p in ptrs])
    return arr
