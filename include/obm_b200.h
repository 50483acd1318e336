/*
 * obm_b200.h — C ABI of the B200-native biogeochemical hot path.
 *
 * Every entry point replaces one piece of OceanBioME.jl (v0.17.6) arithmetic; the
 * reference interface each one stands in for is cited as `path:line` relative to the
 * reference tree.  The Julia glue `ccall`s these from the same
 * hooks OceanBioME implements today (`update_biogeochemical_state!`,
 * `update_tendencies!`, …), so nothing else in a user script changes.
 *
 * Conventions
 *  - plain C linkage, POD structs, no C++ / torch types, no exceptions;
 *  - all data pointers are DEVICE pointers owned by the caller (CUDA.jl CuArray
 *    parents); the library never allocates, frees or retains them;
 *  - pointer tables (`const double* const*`) and parameter structs are HOST memory,
 *    read synchronously at call time (copied into the kernel parameter space);
 *  - every call only enqueues work on the caller's CUDA stream (`void* stream` is a
 *    `cudaStream_t`; NULL = legacy default stream) and returns immediately;
 *  - return value: 0 = OK, negative = argument error (OBM_E*), positive = cudaError_t
 *    from the launch; `obm_last_error()` gives a thread-local message;
 *  - numerical non-convergence is not an error (the reference returns its last
 *    iterate, src/Utils/solvers.jl:118-120); NaNs propagate.
 *
 * Field layout (all 3-D fields): the parent array of an Oceananigans `Field` — Julia
 * column-major with halos.  Interior cell (i,j,k), 0-based, lives at parent index
 *     (i+Hx) + (Nx+2Hx) * ((j+Hy) + (Ny+2Hy) * (k+Hz)).
 * 2-D fields (`Field{Center,Center,Nothing}`) use the same x-y plane layout with a
 * single k plane.  `Flat` dimensions have N = 1, H = 0.
 */
#ifndef OBM_B200_H
#define OBM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OBM_VERSION 100 /* 0.1.0 */

/* error codes (negative = argument errors) */
#define OBM_OK 0
#define OBM_ENULL (-1)    /* required pointer is NULL                        */
#define OBM_ESIZE (-2)    /* bad grid size / range / count                   */
#define OBM_EENUM (-3)    /* unsupported enum / model combination            */
#define OBM_ENOTIMPL (-4) /* valid request that this build does not support  */

/* ------------------------------------------------------------------------------------
 * Grid descriptor.  Mirrors what the kernels of the reference read from an Oceananigans
 * grid: sizes, halos, and z nodes (`znodes(grid, Center/Face)` — src/Light/2band.jl:15-16).
 * ------------------------------------------------------------------------------------ */
typedef struct obm_grid {
    int32_t Nx, Ny, Nz;     /* interior size                                                  */
    int32_t Hx, Hy, Hz;     /* halo widths (0 in Flat dimensions)                             */
    int32_t i0, i1, j0, j1; /* half-open 0-based interior sub-range; i1<=0 → Nx, j1<=0 → Ny   */
    const double* zc;       /* DEVICE: parent of z centres, Nz+2Hz entries, cell k at zc[k+Hz] */
    const double* zf;       /* DEVICE: parent of z faces, Nz+1+2Hz entries, face k at zf[k+Hz] */
    /* Immersed boundary (grid-fitted bottom), nullable: DEVICE x–y plane (parent layout, halos included) of the 1-based
     * index of the bottom-most ACTIVE cell of every column, as `calculate_bottom_indices` returns it
     * (src/Sediments/bottom_indices.jl:19-26; obm_find_bottom_cells).  Cells below it are `immersed_cell`s: the kernels
     * the reference guards with `!immersed_cell(i, j, k, grid)` — ScaleNegativeTracers, src/Utils/negative_tracers.jl:194,253
     * — leave them untouched.  NULL = no immersed cells. */
    const int64_t* bottom_indices_xy;
} obm_grid;

/* ------------------------------------------------------------------------------------
 * (a1, a2) Nutrients–Plankton–Detritus family: NPZD, LOBSTER and every component mix of
 * src/Models/AdvectedPopulations/NutrientsPlanktonDetritus/ (struct :27-33).
 * ------------------------------------------------------------------------------------ */
enum { OBM_NUT_NUTRIENT = 0, OBM_NUT_NITRATE_AMMONIA = 1, OBM_NUT_NITRATE_AMMONIA_IRON = 2 };
enum { OBM_DET_NONE = 0, OBM_DET_DETRITUS = 1, OBM_DET_TWO_PARTICLE = 2, OBM_DET_VARIABLE_REDFIELD = 3 };
enum { OBM_LIGHT_MONDO = 0, OBM_LIGHT_ANALYTICAL = 1 }; /* plankton.jl:216-220 */
enum { OBM_LINEAR = 0, OBM_QUADRATIC = 1 };             /* plankton.jl:83-90   */

typedef struct obm_npd_params {
    int32_t nutrients;                           /* OBM_NUT_*  nutrients.jl:16,41,81          */
    int32_t detritus;                            /* OBM_DET_*  detritus.jl:25,71,264          */
    int32_t carbonate_replicates;                /* 0 = no CarbonateSystem, N = CarbonateSystem(N) carbonate_system.jl:39-46 */
    int32_t oxygen;                              /* 0/1        oxygen.jl:14                   */
    int32_t light_limitation;                    /* OBM_LIGHT_*                               */
    int32_t phytoplankton_mortality_formulation; /* OBM_LINEAR / OBM_QUADRATIC                */
    int32_t grazing_concentration_formulation;   /* OBM_LINEAR / OBM_QUADRATIC                */
    int32_t has_temperature_coefficient;         /* Q10 present ⇒ tracer T is required (plankton.jl:80) */
    /* PhytoZoo — plankton.jl:19-58 */
    double nitrate_half_saturation;
    double ammonia_half_saturation;
    double iron_half_saturation;
    double nitrate_ammonia_inhibition;
    double light_half_saturation;
    double phytoplankton_maximum_growth_rate;
    double iron_ratio;
    double phytoplankton_exudation_fraction;
    double ammonia_fraction_of_exudate;
    double temperature_coefficient;
    double phytoplankton_mortality_rate;
    double zooplankton_mortality_rate;
    double zooplankton_excretion_rate;
    double phytoplankton_solid_waste_fraction;
    double excretion_inorganic_fraction;
    double preference_for_phytoplankton;
    double maximum_grazing_rate;
    double grazing_half_saturation;
    double zooplankton_assimilation_fraction;
    double zooplankton_calcite_dissolution;
    double redfield_ratio;
    double carbon_calcite_ratio;
    double zooplankton_gut_calcite_dissolution;
    double phytoplankton_chlorophyll_ratio;
    /* NitrateAmmonia / NitrateAmmoniaIron — nutrients.jl:16-19,41-43 */
    double nitrification_rate;
    /* TwoParticleAndDissolved / VariableRedfieldDetritus — detritus.jl:25-37,71-81 */
    double remineralisation_inorganic_fraction;
    double small_remineralisation_rate;
    double large_remineralisation_rate;
    double dissolved_remineralisation_rate;
    double small_solid_waste_fraction;
    double detritus_redfield_ratio; /* TwoParticleAndDissolved.redfield_ratio / Detritus.redfield_ratio */
    /* Detritus (Kuhn 2015) — detritus.jl:264-270 */
    double remineralisation_rate;
    double small_particle_fraction;
    /* Oxygen — oxygen.jl:14-17 */
    double respiration_oxygen_nitrogen_ratio;
    double nitrification_oxygen_nitrogen_ratio;
} obm_npd_params;

#define OBM_NPD_MAX_TRACERS 32

/* Number of tracers and their order = `required_biogeochemical_tracers(bgc)`
 * (NutrientsPlanktonDetritus.jl:69-74).  Writes up to OBM_NPD_MAX_TRACERS NUL-terminated
 * UTF-8 names (≤ 15 bytes each) into names[i][16] when names != NULL.  Returns the count,
 * or a negative error. */
int obm_npd_tracer_names(const obm_npd_params* p, char (*names)[16]);

/* Fused replacement of the per-tracer callables
 *   bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)
 * (nutrients.jl:48-90, plankton.jl:92-116, detritus.jl:85-158,282-288,
 *  carbonate_system.jl:50-83, oxygen.jl:21-31) evaluated for EVERY tracer of a cell in one
 * pass.  `tracers[n]`, `G[n]` follow obm_npd_tracer_names order; a NULL G[n] is skipped
 * (e.g. T, which has no biogeochemical tendency).  accumulate = 0: G = tendency;
 * accumulate = 1: G += tendency (the `update_tendencies!` seam, src/OceanBioME.jl:148-152,
 * same pattern as src/Sediments/tracer_coupling.jl:37). */
int obm_npd_tendencies(const obm_grid* grid, const obm_npd_params* p,
                       const double* const* tracers, const double* PAR,
                       double* const* G, int accumulate, void* stream);

/* Parameter-sweep ensembles.  The reference calibrates by building one box model per parameter vector
 * and running them one after the other on the CPU (examples/data_assimilation.jl:26-52, 118-130: NPZD with five
 * PhytoZoo parameters per ensemble member).  Here every horizontal column (i, j) of the grid is a member: the same fused
 * kernel evaluates all members in one launch, member m = i + Nx·j reading its own value of each varied parameter.
 *   which[v]  (host, nvary ≤ OBM_NPD_MAX_VARIED) — the varied parameter, as its position among the `double` members of
 *             obm_npd_params in declaration order (obm_npd_param_index("phytoplankton_maximum_growth_rate") …);
 *   values    (DEVICE, [nvary][Nx·Ny], member fastest) — the member's value of parameter which[v].
 * Everything not named in `which` — and the structure of the model, i.e. the int32 members — comes from `p`.
 * A member's result does not depend on the other members (bit for bit); against obm_npd_tendencies run with that member's
 * block it agrees to rounding — the two kernel instantiations may contract different multiply-adds. */
#define OBM_NPD_MAX_VARIED 16
int obm_npd_param_index(const char* name); /* ≥ 0, or OBM_EENUM when obm_npd_params has no such double member */
int obm_npd_tendencies_ensemble(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which,
                                const double* values, const double* const* tracers, const double* PAR,
                                double* const* G, int accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * (a4) Two-band PAR — src/Light/2band.jl:1-33 (kernel), :35-70 (struct), :106-117 (defaults)
 * ------------------------------------------------------------------------------------ */
typedef struct obm_twoband_params {
    double water_red_attenuation;
    double water_blue_attenuation;
    double chlorophyll_red_attenuation;
    double chlorophyll_blue_attenuation;
    double chlorophyll_red_exponent;
    double chlorophyll_blue_exponent;
    double pigment_ratio;
    double phytoplankton_chlorophyll_ratio;
} obm_twoband_params;

/* Replaces `update_biogeochemical_state!(model, PAR::TwoBandPhotosyntheticallyActiveRadiation)`
 * (2band.jl:148-155).  `surface_PAR_xy` is the glue-evaluated `getbc(surface_PAR, i, j, …)`
 * (2band.jl:4) as a 2-D field in parent x-y layout; if NULL, `surface_PAR_const` is used. */
int obm_par_twoband(const obm_grid* grid, const obm_twoband_params* p, const double* P,
                    const double* surface_PAR_xy, double surface_PAR_const, double* PAR,
                    void* stream);

/* ------------------------------------------------------------------------------------
 * (a5, a6) Multi-band PAR — src/Light/multi_band.jl:147-185 — all bands in one launch,
 * plus the euphotic-depth diagnostic src/Light/compute_euphotic_depth.jl:3-29.
 * ------------------------------------------------------------------------------------ */
#define OBM_MAX_BANDS 8
typedef struct obm_multiband_params {
    int32_t nbands;
    int32_t _pad;
    double water_attenuation_coefficient[OBM_MAX_BANDS];
    double chlorophyll_exponent[OBM_MAX_BANDS];
    double chlorophyll_attenuation_coefficient[OBM_MAX_BANDS];
    double surface_PAR_division[OBM_MAX_BANDS];
} obm_multiband_params;

/* Chl = chl_scale * (chl_a + chl_b) with chl_b nullable: PISCES passes PChl, DChl with
 * scale 1 (PISCES/coupling_utils.jl:7), the NPD family passes P with
 * scale = phytoplankton_chlorophyll_ratio (NutrientsPlanktonDetritus/coupling_utils.jl:54).
 * PAR_bands[n] receives band n; PAR_total (nullable) receives Σ bands (the lazy `sum(fields)`
 * multi_band.jl:120 materialised). */
int obm_par_multiband(const obm_grid* grid, const obm_multiband_params* p, const double* chl_a,
                      const double* chl_b, double chl_scale, const double* surface_PAR_xy,
                      double surface_PAR_const, double* const* PAR_bands, double* PAR_total,
                      void* stream);

/* obm_par_multiband that also leaves the two column diagnostics PISCES derives from the total PAR in
 * the same stage — `compute_euphotic_depth!` (compute_euphotic_depth.jl:3-40, below) and
 * `compute_mixed_layer_mean!` of PAR (PISCES/mean_mixed_layer_properties.jl:10-49; PISCES/update_state.jl:7,11)
 * — while each level's total is on chip: one launch instead of three, PAR is not re-read.  Same results
 * as the three calls (the mean is summed pairwise instead of top-down: ≤ 1e-15 relative apart).
 * PAR_total is required (its top halo cell is read as found, like obm_euphotic_depth). */
int obm_par_multiband_column_state(const obm_grid* grid, const obm_multiband_params* p,
                                   const double* chl_a, const double* chl_b, double chl_scale,
                                   const double* surface_PAR_xy, double surface_PAR_const,
                                   double* const* PAR_bands, double* PAR_total,
                                   const double* mixed_layer_depth_xy, double cutoff,
                                   double* zeu_xy, double* mean_mixed_layer_PAR_xy, void* stream);

/* `compute_euphotic_depth!(euphotic_depth, PAR, cutoff)` compute_euphotic_depth.jl:31-40.
 * Reads PAR[i,j,Nz+1] — a halo cell — exactly as the reference does (:6). zeu_xy is a 2-D
 * field in parent x-y layout. */
int obm_euphotic_depth(const obm_grid* grid, const double* PAR, double cutoff, double* zeu_xy,
                       void* stream);

/* `compute_mixed_layer_mean!(Cₘₓₗ, mixed_layer_depth, C, grid)`
 * PISCES/mean_mixed_layer_properties.jl:10-49 — depth-weighted mean of C above zₘₓₗ (used for
 * κ̄ and PAR̄, PISCES/update_state.jl:9,11).  C == NULL ⇒ the constant C_const (ConstantField). */
int obm_mixed_layer_mean(const obm_grid* grid, const double* mixed_layer_depth_xy, const double* C,
                         double C_const, double* mean_xy, void* stream);

/* ------------------------------------------------------------------------------------
 * (a7) Carbonate chemistry — src/Models/CarbonChemistry/carbon_chemistry.jl:111-210,
 * alkalinity_residual.jl:18-75, equilibrium_constants.jl, calcite_concentration.jl:1-80,
 * density src/Models/seawater_density.jl:33-39 (SeawaterPolynomials TEOS-10, 55 terms).
 * All default constants of `CarbonChemistry()` (carbon_chemistry.jl:66-87).
 * ------------------------------------------------------------------------------------ */
enum {
    OBM_CC_FCO2 = 0,      /* output = Val(:fCO₂)  carbon_chemistry.jl:167  */
    OBM_CC_PCO2 = 1,      /* Val(:pCO₂)  :170-193                           */
    OBM_CC_PH_FREE = 2,   /* Val(:pHᶠ)   :168                               */
    OBM_CC_PH_TOTAL = 3,  /* Val(:pHᵗ)   :195-200                           */
    OBM_CC_PH_SEAWATER = 4, /* Val(:pHˢ) :202-210                           */
    OBM_CC_CO3 = 5,       /* carbonate_concentration  calcite_concentration.jl:1-53 */
    OBM_CC_OMEGA_CALCITE = 6 /* calcite_saturation  calcite_concentration.jl:55-80  */
};

typedef struct obm_carbchem_params {
    int32_t newton_iterations; /* upper bound on the FP64 steps of the branch-free ln[H] Newton (default 12 when <= 0).  Without a
                                * stored [H⁺] (and without phosphate) the search — carbonate-alkalinity quadratic + three Newton
                                * steps — runs in FP32, and ONE FP64 step with its second-order error removed finishes it: a fixed
                                * count for every sea-water state; a state whose FP64 step is not below 1e-5 in ln H keeps its
                                * warp in the FP64 loop (warp-uniform exit, never more than this many steps). */
    int32_t _pad;
    double initial_pH_guess;   /* default 8 (carbon_chemistry.jl:121) when <= 0: fallback start and borate term of the analytic start */
} obm_carbchem_params;

/* Flat sweep over n cells.  T °C, S PSU, DIC mmol/m³, Alk meq/m³; optional (nullable)
 * P bar, silicate & phosphate mmol/m³, pH (given ⇒ skip the solve, carbon_chemistry.jl:213). */
int obm_carbon_chemistry(int64_t n, const obm_carbchem_params* p, const double* T,
                         const double* S, const double* DIC, const double* Alk,
                         const double* P_bar, const double* silicate, const double* phosphate,
                         const double* pH, int output_kind, double* out, void* stream);

/* Gridded Ω for PISCES — `compute_calcite_saturation!`
 * (PISCES/compute_calcite_saturation.jl:9-37): P = |z|·g·1026/1e5 bar, silicate = Si.
 * H_state (nullable, 3-D parent, in/out): [H⁺] (mol/kg) of each cell as left by the previous call; a
 * plausible value (pH 2 … 13) warm-starts the Newton iteration — 1–3 steps per stage — anything else
 * (e.g. a zero-filled field on the first call) falls back to the analytic starting point every solve
 * uses.  The root found is the same to ≈ 1e-14 either way. */
int obm_calcite_saturation(const obm_grid* grid, const obm_carbchem_params* p, const double* T,
                           const double* S, const double* DIC, const double* Alk,
                           const double* Si, double* Omega, double* H_state, void* stream);

/* ------------------------------------------------------------------------------------
 * (a3) PISCES — src/Models/AdvectedPopulations/PISCES/ (struct PISCES.jl:53-92), default
 * component types: MixedMondo nano + diatoms, QualityDependant micro + meso zooplankton,
 * DissolvedOrganicCarbon, TwoCompartmentCarbonIronParticles, NitrateAmmonia, SimpleIron,
 * Silicate, Oxygen, Phosphate, InorganicCarbon.
 * ------------------------------------------------------------------------------------ */
enum { OBM_GROWTH_NUTRIENT_LIMITED = 0, OBM_GROWTH_RESPIRATION_LIMITED = 1 }; /* growth_rate.jl:69,107 */

typedef struct obm_pisces_phyto { /* MixedMondo — phytoplankton/mixed_mondo.jl:23-93 */
    int32_t growth_rate_kind;     /* OBM_GROWTH_*                                            */
    int32_t silicate_limited;     /* nutrient_limitation.jl:15                               */
    /* growth rate — growth_rate.jl:69-75,107-115 */
    double base_growth_rate, temperature_sensitivity, dark_tolerance, initial_slope_of_PI_curve,
        low_light_adaptation, basal_respiration_rate, reference_growth_rate;
    /* NitrogenIronPhosphateSilicateLimitation — nutrient_limitation.jl:10-18 */
    double minimum_ammonium_half_saturation, minimum_nitrate_half_saturation,
        minimum_phosphate_half_saturation, optimal_iron_quota, minimum_silicate_half_saturation,
        silicate_half_saturation_parameter;
    /* MixedMondo */
    double exudated_fraction, blue_light_absorption, green_light_absorption, red_light_absorption,
        mortality_half_saturation, linear_mortality_rate, base_quadratic_mortality,
        maximum_quadratic_mortality, minimum_chlorophyll_ratio, maximum_chlorophyll_ratio,
        maximum_iron_ratio, silicate_half_saturation, enhanced_silicate_half_saturation,
        optimal_silicate_ratio, half_saturation_for_iron_uptake, threshold_for_size_dependency,
        size_ratio;
} obm_pisces_phyto;

typedef struct obm_pisces_zoo { /* QualityDependantZooplankton — zooplankton/food_quality_dependant.jl:11-35 */
    double temperature_sensitivity, maximum_grazing_rate;
    double food_preferences[4]; /* NamedTuple order (P, D, POC, Z) — zooplankton/defaults.jl:4,12 */
    double food_threshold_concentration, specific_food_threshold_concentration,
        grazing_half_saturation, maximum_flux_feeding_rate, iron_ratio, minimum_growth_efficiency,
        non_assimilated_fraction, mortality_half_saturation, quadratic_mortality, linear_mortality,
        dissolved_excretion_fraction, undissolved_calcite_fraction;
} obm_pisces_zoo;

typedef struct obm_pisces_params {
    obm_pisces_phyto nano, diatoms; /* mixed_mondo_nano_diatoms.jl:1-28                         */
    double base_rain_ratio;         /* nano_and_diatoms.jl:4                                    */
    obm_pisces_zoo micro, meso;     /* zooplankton/defaults.jl:2-21                             */
    /* MicroAndMeso — zooplankton/micro_and_meso.jl:3-17 */
    double microzooplankton_bacteria_concentration, mesozooplankton_bacteria_concentration,
        maximum_bacteria_concentration, bacteria_concentration_depth_exponent,
        doc_half_saturation_for_bacterial_activity, nitrate_half_saturation_for_bacterial_activity,
        ammonia_half_saturation_for_bacterial_activity,
        phosphate_half_saturation_for_bacterial_activity, iron_half_saturation_for_bacterial_activity;
    /* DissolvedOrganicCarbon — dissolved_organic_matter/dissolved_organic_carbon.jl:9-35 */
    double dom_remineralisation_rate, dom_reference_bacteria_concentration,
        dom_temperature_sensitivity, dom_aggregation_parameters[5];
    /* TwoCompartmentCarbonIronParticles — particulate_organic_matter/two_size_class.jl:17-84 */
    double pom_temperature_sensitivity, pom_base_breakdown_rate, pom_aggregation_parameters[4],
        minimum_iron_scavenging_rate, load_specific_iron_scavenging_rate,
        bacterial_iron_uptake_efficiency, small_fraction_of_bacterially_consumed_iron,
        large_fraction_of_bacterially_consumed_iron, base_liable_silicate_fraction,
        fast_dissolution_rate_of_silicate, slow_dissolution_rate_of_silicate,
        base_calcite_dissolution_rate, calcite_dissolution_exponent,
        maximum_iron_ratio_in_bacteria, iron_half_saturation_for_bacteria,
        maximum_bacterial_growth_rate;
    /* NitrateAmmonia — nitrogen/nitrate_ammonia.jl:10-16 */
    double maximum_nitrification_rate, maximum_fixation_rate, iron_half_saturation_for_fixation,
        phosphate_half_saturation_for_fixation, light_saturation_for_fixation;
    /* SimpleIron — iron/simple_iron.jl:9-13 */
    double excess_scavenging_enhancement, maximum_ligand_concentration, dissolved_ligand_ratio;
    /* Oxygen — oxygen.jl:21-24 */
    double ratio_for_respiration, ratio_for_nitrification;
    /* PISCES.jl:69-76 */
    double first_anoxia_threshold, second_anoxia_threshold, nitrogen_redfield_ratio,
        phosphate_redfield_ratio, mixed_layer_shear, background_shear;
    /* Host-evaluated per launch (they depend on (clock.time, latitude) only; user callables cannot
     * cross the ABI).  `latitude` = PrescribedLatitude (common.jl:20-25).  The reference calls
     * `bgc.day_length(φ, clock.time)` with SWAPPED arguments in every growth rate
     * (growth_rate.jl:30) and `bgc.day_length(clock.time, φ)` in chlorophyll synthesis
     * (growth_rate.jl:143); both are reproduced: */
    double latitude;
    double day_length_growth;      /* = day_length(φ, t)  — the swapped call                     */
    double day_length_chlorophyll; /* = day_length(t, φ)                                         */
    double silicate_climatology;   /* Si′ = ConstantField(7.5) (PISCES.jl:319)                   */
} obm_pisces_params;

#define OBM_PISCES_NTRACERS 26 /* 24 prognostic + T, S — order of PISCES.jl:94-105:
   P PChl PFe D DChl DFe DSi Z M DOC POC GOC SFe BFe PSi CaCO₃ NO₃ NH₄ PO₄ Fe Si DIC Alk O₂ T S */

typedef struct obm_pisces_fields { /* biogeochemical_auxiliary_fields — PISCES.jl:107-118 + light */
    const double* PAR1;            /* 3-D centre fields                                           */
    const double* PAR2;
    const double* PAR3;
    const double* PAR;             /* total (multi_band.jl:120)                                   */
    const double* Omega;           /* calcite saturation Ω                                        */
    const double* wPOC;            /* z-FACE fields (Nz+1 levels, face k at parent level k+Hz)    */
    const double* wGOC;
    const double* mixed_layer_depth_xy;  /* 2-D fields, parent x-y layout                         */
    const double* euphotic_depth_xy;
    const double* mean_mixed_layer_vertical_diffusivity_xy;
    const double* mean_mixed_layer_light_xy;
} obm_pisces_fields;

/* Fused replacement of the 24 per-tracer callables of PISCES (files listed in SURVEY §8 a3):
 * every tracer of a cell is read once, all shared sub-models (nutrient limitation, growth rates,
 * grazing, mortalities, bacteria, aggregation, iron chemistry) are evaluated once in registers
 * and all 24 tendencies are written in one pass.  tracers[n] / G[n] in OBM_PISCES_NTRACERS
 * order; G[n] == NULL is skipped (T, S have no tendency, PISCES.jl:120). */
int obm_pisces_tendencies(const obm_grid* grid, const obm_pisces_params* p,
                          const double* const* tracers, const obm_pisces_fields* aux,
                          double* const* G, int accumulate, void* stream);

/* The same launch for `latitude = ModelLatitude()` (PISCES/common.jl:9-13,27-28: φ = φnode(i, j, k, grid), what a
 * LatitudeLongitudeGrid needs; PISCES.jl:360-367 decides between the two).  The three members of obm_pisces_params
 * that depend on the latitude then differ from row to row: `row_latitude_daylengths` is a DEVICE array [3][grid->Ny] —
 * latitude (°), day_length(φ, t) (the swapped call of growth_rate.jl:29-30) and day_length(t, φ) (:141-143) of every
 * interior row j — evaluated by the host like their scalar counterparts; the members of `p` are ignored.  Everything
 * else, and the results for a table of Ny identical rows, are those of obm_pisces_tendencies bit for bit. */
int obm_pisces_tendencies_rows(const obm_grid* grid, const obm_pisces_params* p,
                               const double* row_latitude_daylengths, const double* const* tracers,
                               const obm_pisces_fields* aux, double* const* G, int accumulate,
                               void* stream);

/* ------------------------------------------------------------------------------------
 * (a9) ScaleNegativeTracers — src/Utils/negative_tracers.jl:137-276.  All groups of a model
 * in ONE launch, applied sequentially in the given order (PISCES: carbon, iron, phosphate,
 * silicon, nitrogen — PISCES/coupling_utils.jl:31).
 * ------------------------------------------------------------------------------------ */
#define OBM_MAX_SCALE_TRACERS 32
#define OBM_MAX_SCALE_GROUPS 8
#define OBM_MAX_GROUP_SIZE 16
typedef struct obm_scale_group {
    int32_t n;                                /* tracers in this group                          */
    int32_t index[OBM_MAX_GROUP_SIZE];        /* indices into the `tracers` pointer table        */
    double scalefactor[OBM_MAX_GROUP_SIZE];   /* negative_tracers.jl:40                          */
} obm_scale_group;

int obm_scale_negative_tracers(const obm_grid* grid, int ntracers, double* const* tracers,
                               int ngroups, const obm_scale_group* groups,
                               double invalid_fill_value, void* stream);

/* The PISCES stage prologue in ONE launch: the modifiers' negative scaling (above) followed by
 * `compute_calcite_saturation!` (PISCES/compute_calcite_saturation.jl:9-37, see obm_calcite_saturation)
 * on the same cell.  Replaces the pair  update_biogeochemical_state!(model, modifiers)
 * (OceanBioME.jl:169) … compute_calcite_saturation! (PISCES/update_state.jl:13): both are pointwise
 * and nothing between them writes DIC, Alk, Si, T or S, so Ω is solved from the cell's rescaled
 * values while they are on chip; the bandwidth-bound scaling pass is hidden under the FP64-bound
 * solve.  DIC / Alk / Si are matched against `tracers` by pointer (a match uses the rescaled
 * value); T and S must not be rescaled tracers (OBM_ESIZE).  Results are those of the two
 * separate calls. */
int obm_scale_negative_tracers_calcite_saturation(const obm_grid* grid, int ntracers,
                                                  double* const* tracers, int ngroups,
                                                  const obm_scale_group* groups,
                                                  double invalid_fill_value,
                                                  const obm_carbchem_params* p, const double* T,
                                                  const double* S, const double* DIC,
                                                  const double* Alk, const double* Si,
                                                  double* Omega, double* H_state, void* stream);

/* `ZeroNegativeTracers` negative_tracers.jl:22-32: parent .= max.(0, parent) over the whole
 * parent array (halos included, as the reference does). n = parent element count. */
int obm_zero_negative_tracers(int64_t n_parent, int ntracers, double* const* tracers,
                              void* stream);

/* ------------------------------------------------------------------------------------
 * (a8) Bottom sediments — src/Sediments/ (driver) + src/Models/Sediments/ (equations).
 * The reference runs ≈ 15 tiny :xy launches per stage (K7–K11 of SURVEY §2.2); here there are
 * two fused :xy launches, one per hook:
 *   obm_sediment_update_state      ↔ update_biogeochemical_state!(model, ::BiogeochemicalSediment)
 *                                    (Sediments/update_state.jl:6-16): gather bottom-cell tracers
 *                                    (tracked_fields.jl:43-51), sinking fluxes (:60-71), step the
 *                                    sediment pools with the stored tendencies (timesteppers.jl:15-73),
 *                                    cache G⁻ ← Gⁿ (:78-94), recompute Gⁿ (compute_tendencies.jl:5-50);
 *   obm_sediment_update_tendencies ↔ update_tendencies!(bgc, sediment, model)
 *                                    (Sediments/tracer_coupling.jl:3-39): G[i,j,k_bottom] += flux / Δz.
 * PARITY UNPINNED: the reference's sediment tests are commented out (test/test_sediments.jl:106-163)
 * and the stepping order / flux operator live in Oceananigans (not in the tree) — see DESIGN.md.
 * ------------------------------------------------------------------------------------ */
enum { OBM_SED_INSTANT_REMINERALISATION = 0, OBM_SED_SIMPLE_MULTI_G = 1 };
enum { OBM_ADV_UPWIND1 = 0, OBM_ADV_CENTERED2 = 1, OBM_ADV_UPWIND3 = 2, OBM_ADV_WENO5 = 3 }; /* face reconstruction of advective_tracer_flux_z (UPWIND3, WENO5: obm_sinking_tendencies; the sediment's bottom face is first-order upwind for every upwind-biased scheme) */
enum { OBM_TS_AB2 = 0, OBM_TS_RK3 = 1 };             /* Sediments.jl:48: timestepper                   */
#define OBM_SED_MAX_SINKING 4
#define OBM_SED_MAX_POOLS 6
#define OBM_SED_MAX_COUPLED 4

typedef struct obm_sediment_params {
    int32_t model;             /* OBM_SED_*                                                          */
    int32_t carbon;            /* SimpleMultiG{Nothing}: separate C pools + sinking carbon tracers   */
    int32_t nsinking_nitrogen; /* length(sinking_nitrogen) (InstantRemineralisation: sinking_tracers) */
    int32_t nsinking_carbon;
    int32_t advection;         /* OBM_ADV_*                                                          */
    int32_t timestepper;       /* OBM_TS_*                                                           */
    /* InstantRemineralisation — Models/Sediments/instant_remineralisation.jl:13-19,83-96 */
    double burial_efficiency_constant1, burial_efficiency_constant2, burial_efficiency_half_saturation;
    /* SimpleMultiG — Models/Sediments/simple_multi_G.jl:15-38,104-132 */
    double sinking_redfield, fast_decay_rate, slow_decay_rate, fast_redfield, slow_redfield,
        fast_fraction, slow_fraction, refactory_fraction, sedimentation_rate, anoxia_half_saturation;
    double nitrate_oxidation_params[6], denitrification_params[6], anoxic_params[6],
        solid_dep_params[4];
} obm_sediment_params;

typedef struct obm_sediment_fields {
    const int64_t* bottom_indices_xy; /* 2-D Int field, 1-based k of the bottom cell
                                         (bottom_indices.jl:19-26); NULL ⇒ OneField ⇒ k = 1      */
    /* water-column side (3-D parents) */
    const double* NO3;                /* tracked tracers of SimpleMultiG (simple_multi_G.jl:42)    */
    const double* NH4;
    const double* O2;
    const double* sinking[2 * OBM_SED_MAX_SINKING];   /* sinking tracers: nitrogen…, then carbon… */
    const double* sinking_w[2 * OBM_SED_MAX_SINKING]; /* their w: z-FACE fields                    */
    /* sediment side (2-D parents): pools in required_sediment_fields order
       (InstantRemineralisation: storage; SimpleMultiG: Ns Nf Nr [Cs Cf Cr]) */
    double* pools[OBM_SED_MAX_POOLS];
    double* Gn[OBM_SED_MAX_POOLS];    /* timestepper.Gⁿ / G⁻ of the pools                         */
    double* Gm[OBM_SED_MAX_POOLS];
    double* tracked_xy[3 + 2 * OBM_SED_MAX_SINKING]; /* tracked_fields (tracers…, fluxes…), written every call */
    /* tendencies of the coupled tracers (3-D): InstantRemineralisation: the receiver;
       SimpleMultiG: NO₃ NH₄ O₂ [DIC] (simple_multi_G.jl:45-46) */
    double* G_coupled[OBM_SED_MAX_COUPLED];
} obm_sediment_fields;

/* dt = model.clock.last_stage_Δt (non-finite ⇒ pools are not stepped, update_state.jl:11-13):
 * `time_step!(sediment_model, dt)` in one launch.  AB2: one ab2_step! with chi (χ = −0.5 ⇒ Euler;
 * timesteppers.jl:16-42).  RK3: the sediment's whole three-stage step (rk3_substep! :46-73,
 * cache_previous_tendencies! :84-96 and the tendency recompute per stage) with the tracked
 * tracers and fluxes held fixed; chi is ignored. */
int obm_sediment_update_state(const obm_grid* grid, const obm_sediment_params* p,
                              const obm_sediment_fields* f, double dt, double chi, void* stream);
int obm_sediment_update_tendencies(const obm_grid* grid, const obm_sediment_params* p,
                                   const obm_sediment_fields* f, void* stream);

/* K12 `find_bottom_cell!` (bottom_indices.jl:7-17) for a grid-fitted bottom: k_bottom = first
 * (1-based) k whose centre is NOT immersed, i.e. not (z_c[k] <= bottom_height[i,j]), capped at Nz.
 * Integer work: bit-exact. */
int obm_find_bottom_cells(const obm_grid* grid, const double* bottom_height_xy,
                          int64_t* bottom_indices_xy, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-1) Air–sea gas exchange — src/Models/GasExchange/: the `GasExchange` callable
 * (gas_exchange.jl:26-38) evaluated for every surface column in one x–y launch:
 *     flux[i,j] = k(u₁₀, T, S) · (water − air),
 *     k = k₆₆₀(u₁₀) / √(Sc(T)/660) · solubility(T, S)          (gas_transfer_velocity.jl:32-33)
 * with T, S (and every tracer) read at the top cell k = Nz.  Positive = out of the ocean
 * (Oceananigans top-flux convention).  The result is the plain 2-D field Oceananigans reads as
 * a flux boundary condition; if `G_top` is given the kernel also applies it to the tendency of
 * the top cell, G[i,j,Nz] -= flux / Δzᶜ(Nz), which is what Oceananigans' `apply_z_bcs!` does
 * with the boundary value.
 * ------------------------------------------------------------------------------------ */
enum {
    OBM_GE_WATER_TRACER = 0, /* OxygenConcentration / any tracer: surface value, GasExchange.jl:37 */
    OBM_GE_WATER_PCO2 = 1    /* CarbonDioxideConcentration: carbon-chemistry pCO₂ of the surface
                                cell, carbon_dioxide_concentration.jl:48-60 */
};
enum {
    OBM_GE_AIR_PLAIN = 0,        /* number / field as is, surface_values.jl:4,33-34 */
    OBM_GE_AIR_WANNINKHOF92 = 1  /* PartiallySolubleGas: air · β(T,S)/Tk, gas_solubility.jl:24-49 */
};
enum {
    OBM_GE_SOLUBILITY_ONE = 0,     /* default `(T, S) -> 1`, gas_transfer_velocity.jl:29 */
    OBM_GE_SOLUBILITY_K0_RHO = 1   /* MolPerKgPerAtmToMMolPerCubicMPerMicroAtm: K0(T+273.15,S)·ρ(T,S)/10³,
                                      gas_solubility.jl:65 (CO₂ default, GasExchange.jl:105-106) */
};

typedef struct obm_gas_exchange_params {
    int32_t water_kind;      /* OBM_GE_WATER_*      */
    int32_t air_kind;        /* OBM_GE_AIR_*        */
    int32_t solubility_kind; /* OBM_GE_SOLUBILITY_* */
    int32_t k660_order;      /* order (0…3) of the base transfer velocity polynomial in u₁₀ */
    int32_t use_silicate_phosphate; /* 0: `nothing` → (0, 0); 1: fields if given else the two values below */
    int32_t _pad;
    double k660[4];          /* k₆₆₀ coefficients c₀…c₃ (Ho06: (0, 0, 0.266/hour/100)), gas_transfer_velocity.jl:52-135 */
    double schmidt[5];       /* order-4 Schmidt-number polynomial in T °C, schmidt_number.jl:6-16 */
    double w92[6];           /* A1 A2 A3 B1 B2 B3 of Wanninkhof92Solubility, gas_solubility.jl:46-47 */
    double air_concentration; /* used when the air field is NULL (CO₂ default 413 ppmv, O₂ 9352.7 mmol/m³) */
    double wind_speed;        /* used when the wind field is NULL (default 2 m/s) */
    double silicate, phosphate; /* constant values (NamedTuple form), carbon_dioxide_concentration.jl:63 */
    obm_carbchem_params carbon_chemistry;
} obm_gas_exchange_params;

/* T, S, tracer, DIC, Alk, silicate_f, phosphate_f: 3-D parents (surface level read).
 * wind_speed_xy, air_concentration_xy: 2-D planes in parent x–y layout, or NULL → the scalars.
 * tracer is needed for OBM_GE_WATER_TRACER, DIC/Alk for OBM_GE_WATER_PCO2.
 * flux_xy (2-D plane) and/or G_top (3-D parent) receive the result; at least one is required. */
int obm_gas_exchange_flux(const obm_grid* grid, const obm_gas_exchange_params* p,
                          const double* T, const double* S, const double* tracer,
                          const double* DIC, const double* Alk,
                          const double* silicate_f, const double* phosphate_f,
                          const double* wind_speed_xy, const double* air_concentration_xy,
                          double* flux_xy, double* G_top, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-4) Biologically active particles — src/Particles/ with the sugar-kelp individual model of
 * src/Models/Individuals/SugarKelp/ (Broch & Slagstad 2012 as updated there).  The reference runs, per
 * stage, one launch per coupled tracer (8) that re-evaluates the whole kelp model including the
 * Newton solve for the light-inhibition parameter β, then 3 tendency + 3 Euler launches for the
 * particle fields; here there are two: a scatter launch (all 8 uptake / release terms of a particle
 * from ONE evaluation, atomically added into Gⁿ of the nearest cell) and a step launch (dA, dN, dC
 * and the forward-Euler update).  A user-supplied particle biogeochemistry is a Julia callable and
 * cannot cross the C ABI; SugarKelp is the one the reference ships.
 * ------------------------------------------------------------------------------------ */
enum { OBM_TOPO_PERIODIC = 0, OBM_TOPO_BOUNDED = 1, OBM_TOPO_FLAT = 2 };

typedef struct obm_sugar_kelp_params { /* SugarKelp.jl:36-150, field order kept */
    /* temperature_limit::LinearOptimalTemperatureRange — equations.jl:206-227 */
    double lower_optimal, upper_optimal, lower_gradient, upper_gradient;
    double growth_rate_adjustment, photosynthetic_efficiency, minimum_carbon_reserve, structural_carbon;
    double exudation, erosion_exponent, base_erosion_rate, saturation_irradiance;
    double structural_dry_weight_per_area, structural_dry_to_wet_weight;
    double carbon_reserve_per_carbon, nitrogen_reserve_per_nitrogen;
    double minimum_nitrogen_reserve, maximum_nitrogen_reserve;
    double growth_adjustment_2, growth_adjustment_1, maximum_specific_growth_rate, structural_nitrogen;
    double photosynthesis_at_ref_temp_1, photosynthesis_at_ref_temp_2;
    double photosynthesis_ref_temp_1, photosynthesis_ref_temp_2, photoperiod_1, photoperiod_2;
    double respiration_at_ref_temp_1, respiration_at_ref_temp_2, respiration_ref_temp_1, respiration_ref_temp_2;
    double photosynthesis_arrhenius_temp, photosynthesis_low_temp, photosynthesis_high_temp;
    double photosynthesis_high_arrhenius_temp, photosynthesis_low_arrhenius_temp, respiration_arrhenius_temp;
    double current_speed_for_0p65_uptake, nitrate_half_saturation, ammonia_half_saturation;
    double maximum_nitrate_uptake, maximum_ammonia_uptake, current_1, current_2, current_3;
    double base_activity_respiration_rate, base_basal_respiration_rate, exudation_redfield_ratio;
    double adapted_latitude;
    int32_t newton_iterations; /* cap of the β solve (reference: 1000 with an unreachable atol); <= 0 ⇒ 100 */
    int32_t _pad;
} obm_sugar_kelp_params;

typedef struct obm_particles { /* BiogeochemicalParticles — Particles.jl:25-41 (device arrays of length n) */
    int64_t n;
    const double *x, *y, *z;
    double *A, *N, *C;            /* required_particle_fields(::SugarKelp) = (:A, :N, :C)           */
    const double* scalefactors;   /* nullable ⇒ 1                                                   */
    /* where the particles live horizontally (obm_grid carries only z): first cell CENTRE and spacing */
    double x0, dx, y0, dy;
    int32_t topology[3];          /* OBM_TOPO_* of x, y, z — get_node, tracer_interpolation.jl:5-7  */
    int32_t _pad;
} obm_particles;

typedef struct obm_kelp_tracers { /* required_tracers(::SugarKelp) = (:u, :v, :w, :T, :NO₃, :NH₄, :PAR), 3-D parents */
    const double *u, *v, *w;      /* nullable ⇒ 0 (read at the same (i, j, k) as the reference does)  */
    const double *T, *NO3, *NH4, *PAR;
} obm_kelp_tracers;

#define OBM_KELP_NCOUPLED 8 /* coupled_tracers: NO₃ NH₄ DIC O₂ DOC DON bPOC bPON (SugarKelp.jl:164) */

/* `update_tendencies!(bgc, particles, model)` — update_tracer_tendencies.jl:1-48 with NearestPoint
 * (tracer_interpolation.jl:16-72): G[c][nearest cell] += scalefactor · kelp(Val(c), t, …) / volume
 * for the 8 coupled tracers (G[c] == NULL ⇒ that tracer is not coupled in this model).  Atomic. */
int obm_kelp_update_tendencies(const obm_grid* grid, const obm_sugar_kelp_params* p,
                               const obm_particles* particles, const obm_kelp_tracers* tracers,
                               double* const* G, double t, void* stream);

/* `time_step_particle_fields!(::ForwardEuler, …)` — tendencies.jl:3-35 + time_stepping.jl:18-48:
 * dA, dN, dC from the state as it is, then field += tendency · Δt.  tendencies_out (nullable): three
 * device arrays of length n that receive dA, dN, dC (the reference's timestepper.tendencies). */
int obm_kelp_step(const obm_grid* grid, const obm_sugar_kelp_params* p, const obm_particles* particles,
                  const obm_kelp_tracers* tracers, double t, double dt, double* const* tendencies_out,
                  void* stream);

/* `seasonal_limitation(kelp, t)` (equations.jl:229-255) — a function of the clock only, evaluated on
 * the host by the two calls above. */
double obm_kelp_seasonal_limitation(const obm_sugar_kelp_params* p, double t);

/* ------------------------------------------------------------------------------------
 * (f-2) Sinking: Gⁿ[c] (+)= −∂z(w c) for every tracer with a biogeochemical drift velocity, all in
 * one launch.  w_faces[t]: the z-face field `biogeochemical_drift_velocity(bgc, Val(c)).w`
 * (`setup_velocity_fields` src/Utils/sinking_velocity_fields.jl:10-35, `DepthDependantSinkingSpeed`
 * PISCES/common.jl:39-55; face k at plane k, Nz + 1 faces).  Flux form  F_k = w_k·c̃_k,
 * G_k −= (F_{k+1} − F_k)/Δz_k  with c̃ by `advection` (OBM_ADV_UPWIND1 | CENTERED2 | UPWIND3 | WENO5;
 * UPWIND3 falls back to first order, WENO5 — fifth-order WENO with the Z weights of Borges et al. 2008,
 * uniform-grid coefficients, ε = 1e-8 — to WENO3 and then to first order where the stencil would leave
 * the interior).  This is the
 * `div_Uc(…, total_velocities, c)` term of Oceananigans' tracer tendency for models without a
 * resolved vertical velocity (column / box ensembles, the reference's sediment tests); the
 * bottom-face flux is the one obm_sediment_update_state reads.  Tracer z-halos are read as found.
 * ------------------------------------------------------------------------------------ */
#define OBM_MAX_SINKING_TRACERS 8
int obm_sinking_tendencies(const obm_grid* grid, int ntracers, const double* const* tracers,
                           const double* const* w_faces, double* const* G, int advection,
                           int accumulate, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-2 / f-3) The tracer update either side of the tendency pass, every tracer in one launch:
 *     U[f] += Δt·(γ·Gⁿ[f] + ζ·G⁻[f])   (has_zeta = 0, the first stage: U[f] += Δt·γ·Gⁿ[f])
 *     G⁻[f] ← Gⁿ[f]                     (cache_previous != 0)
 * `rk3_substep!` src/BoxModel/timesteppers.jl:66-93 and `cache_previous_tendencies!` :20-28 (the
 * box model's copies of Oceananigans' per-tracer launches); γ = 1, has_zeta = 0 is forward Euler.
 * U, Gn, Gm: host tables of nfields device parents.  Operation order as written (no FMA).
 * ------------------------------------------------------------------------------------ */
int obm_rk3_substep(const obm_grid* grid, int nfields, double* const* U, const double* const* Gn,
                    double* const* Gm, double dt, double gamma, double zeta, int has_zeta,
                    int cache_previous, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-2) Tendencies AND the tracer update in ONE launch, for models in which nothing but the
 * biogeochemistry (and a forcing already sitting in Gⁿ) moves the tracers — box models, column
 * ensembles, parameter sweeps (src/BoxModel/timesteppers.jl:30-93: compute_tendencies! →
 * rk3_substep! → cache_previous_tendencies!, three passes over every field per stage):
 *     G      = bgc_n(U) (+ Gⁿ[n] as found, when accumulate != 0: the forcing)
 *     U[n]  += Δt·(γ·G + ζ·G⁻[n])        (has_zeta = 0, first stage / Euler: U[n] += Δt·γ·G)
 *     G⁻[n] ← G                           (Gⁿ[n] ← G as well when store_Gn != 0)
 * evaluated per cell from one read of the cell's tracers, so Gⁿ never travels: 312 instead of
 * ≈ 630 B per cell and stage for LOBSTER + carbonates + O₂.  `tracers` are read AND written
 * (every thread reads its whole cell before it writes any of it: pointwise, no hazard); a
 * tracer whose Gm[n] is NULL is not stepped (prescribed fields, T); G may be NULL when neither
 * accumulate nor store_Gn asks for it.  nvary > 0: per-column parameter values exactly as in
 * obm_npd_tendencies_ensemble.  Same arithmetic, same operation order as obm_npd_tendencies
 * followed by obm_rk3_substep — the two paths agree bit for bit (tests/test_gpu_box_model.py).
 * ------------------------------------------------------------------------------------ */
int obm_npd_tendencies_substep(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which,
                               const double* values, double* const* tracers, const double* PAR, double* const* G,
                               int accumulate, int store_Gn, double* const* Gm, double dt, double gamma, double zeta,
                               int has_zeta, void* stream);

/* ------------------------------------------------------------------------------------
 * (f-3) The whole RUN of a box-model ensemble in ONE launch — `run!(simulation)` over
 * src/BoxModel/boxmodel.jl:92-110 + timesteppers.jl:30-93 for n independent boxes laid along x
 * (grid->Ny = grid->Nz = 1).  Boxes do not interact, so a thread integrates its box through
 * `nsteps` time steps of `nstages` stages each (RK3: γ = 8/15, 5/12, 3/4, ζ = NaN, −17/60,
 * −5/12; forward Euler: one stage, γ = 1, ζ = NaN — NaN marks a stage without a ζ term); a
 * stage is exactly obm_npd_tendencies_substep (store_Gn = accumulate = 0) on values held on
 * chip.  Prescribed series come from host-tabulated DEVICE tables with one row per global
 * stage r = step·nstages + stage, row r holding what `update_state!` leaves AFTER that stage
 * (the stage itself sees row r − 1, the very first one the values found in the fields):
 * `PAR_table` [nsteps·nstages][PAR_per_box ? n : 1], optional `T_table` likewise for the
 * temperature tracer.  On return tracers, G⁻, the PAR field and T hold what nsteps calls of
 * time_step! would have left; `snapshots[t]` (tracer order, nullable entries or table)
 * receives tracer t of every box after every `output_every`-th step: [nsteps/output_every][n].
 * Tables, parameter values and snapshots are indexed by the box's position along x (n = grid->Nx).
 * Results are those of the per-stage launches bit for bit (tests/test_gpu_box_model.py); the
 * reference's benchmark/box_model.jl (NPZD box, 1000 RK3 steps: 23.5 ms on its CPU for ONE
 * box) takes one launch for the whole ensemble.  Every tracer the tendencies read must be
 * stepped (Gm[t] != NULL) — OBM_ENOTIMPL otherwise: use the per-stage path for prescribed
 * biogeochemical tracers and for forcings.
 * ------------------------------------------------------------------------------------ */
int obm_npd_box_run(const obm_grid* grid, const obm_npd_params* p, int nvary, const int32_t* which,
                    const double* values, double* const* tracers, double* const* Gm, double* PAR,
                    const double* PAR_table, int PAR_per_box, const double* T_table, int T_per_box,
                    int nsteps, int nstages, const double* gamma, const double* zeta, double dt,
                    int output_every, double* const* snapshots, void* stream);

/* ------------------------------------------------------------------------------------
 * (e) Tracer inventory for conservation diagnostics: out[g] = Σ_cells Σ_f sf[g][f]·c_f·V_cell
 * (the user-side sums of test/test_NutrientsPlanktonDetritus.jl:8-21 at scale).  `out` is a
 * DEVICE array of ngroups doubles, overwritten (deterministic two-level reduction; no atomics
 * on doubles).  `workspace` is a caller-provided DEVICE buffer of
 * obm_inventory_workspace_bytes(ngroups) bytes.  The multi-GPU all-reduce of `out` is done by
 * the caller with NCCL (see oceanbiome.jl_b200/distributed.py).
 * ------------------------------------------------------------------------------------ */
int64_t obm_inventory_workspace_bytes(int ngroups);
int obm_inventory(const obm_grid* grid, int ntracers, const double* const* tracers, int ngroups,
                  const obm_scale_group* groups, const double* cell_volume /*nullable: 3-D field*/,
                  double uniform_volume, double* out, void* workspace, void* stream);

/* ------------------------------------------------------------------------------------
 * Host-buffer staging (for callers whose fields live in pinned HOST memory): copies interior
 * rows j ∈ [grid.j0, grid.j1) of `nplanes` k-planes of each field, host → device (direction 0)
 * or device → host (1), as one strided 2-D copy per field on `stream`.  Because every hot
 * kernel is pointwise or column-local, a stage can be pipelined slab by slab with these copies
 * (oceanbiome.jl_b200/host_stage.py).  Oceananigans fields are device-resident, so the Julia
 * glue does not need this; it is the end-to-end path of bench.py.
 * ------------------------------------------------------------------------------------ */
int obm_copy_slab(const obm_grid* grid, int nfields, void* const* dst, const void* const* src,
                  int nplanes, int direction, void* stream);

/* The same copy driven by a small persistent kernel (one block per SM, 16-byte accesses over PCIe)
 * instead of one DMA descriptor per (field, k-plane) row: ≈ 10⁵ descriptors per direction per stage
 * at 32 slabs cost ≈ 15 % of the PCIe bandwidth.  Host pointers must be pinned memory (mapped for
 * the device by UVA).  Same arguments and result as obm_copy_slab. */
int obm_copy_slab_sm(const obm_grid* grid, int nfields, void* const* dst, const void* const* src,
                     int nplanes, int direction, void* stream);

/* ------------------------------------------------------------------------------------
 * Diagnostic (synchronises): measured FP64-pipe peak in DFMA instructions per second, the
 * denominator of the FP64 roofline.  scratch: DEVICE buffer of >= 148*8*256 doubles.
 * ------------------------------------------------------------------------------------ */
double obm_fp64_peak_dfma_per_s(double* scratch, int iters, void* stream);

/* Diagnostic (synchronises): bandwidth in GB/s of the bare access pattern of a fused tendency kernel —
 * every thread reads its cell from `nread` (<= 40) fields and read-modify-writes (mode 0) or writes
 * (mode 1) `nrmw` (<= 26) fields, same launch geometry, no arithmetic.  The ceiling the memory system
 * sets for that many concurrent streams (a two-stream copy does not show it). */
/* Diagnostic (synchronises): run time in ms of `blocks` blocks x 128 threads x 4 096 DFMA executed as one
 * straight-line instruction stream (straight != 0) or as a 64-instruction loop body — what a single pass over
 * a long instruction stream costs on this device.  scratch: DEVICE buffer of >= blocks*128 doubles. */
double obm_fetch_ceiling_ms(double* scratch, int blocks, int straight, void* stream);

double obm_stream_pattern_gbs(const obm_grid* grid, int nread, const double* const* reads, int nrmw,
                              double* const* rmw, int mode, int reps, void* stream);

/* ------------------------------------------------------------------------------------ */
const char* obm_last_error(void);
int obm_version(void);
/* compile-time facts, for the test-suite: sizeof of each param struct */
int obm_sizeof(const char* struct_name);

#ifdef __cplusplus
}
#endif
#endif /* OBM_B200_H */
