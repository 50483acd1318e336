# OceanBioMEB200.jl — the reference-side binding a maintainer would add: thin `ccall` glue from the hooks
# OceanBioME.jl already implements (src/OceanBioME.jl:53-59,148-169) to libobm_b200.so (include/obm_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: there is no Julia in the build image.  The ctypes binding
# (oceanbiome.jl_b200/_lib.py) mirrors the same C ABI line for line and is what the parity tests exercise.
module OceanBioMEB200

using CUDA, Oceananigans, OceanBioME
using Oceananigans.Fields: interior
import Oceananigans.Biogeochemistry: update_biogeochemical_state!, update_tendencies!

const libobm = get(ENV, "OBM_B200_LIB", "libobm_b200.so")

# ---- struct mirrors: generated from the ctypes definitions the test-suite checks against the compiled library ----
include("obm_structs.jl")   # ObmGrid, ObmNpdParams, ObmTwobandParams, ObmMultibandParams, ObmCarbchemParams, ObmScaleGroup,
                            # ObmPiscesPhyto/Zoo/Params/Fields, ObmSedimentParams/Fields, ObmGasExchangeParams

function ObmGrid(grid)
    Nx, Ny, Nz = size(grid)
    Hx, Hy, Hz = Oceananigans.Grids.halo_size(grid)
    Hx, Hy = (Nx == 1 ? 0 : Hx), (Ny == 1 ? 0 : Hy)                      # Flat dimensions
    zc = parent(grid.z.cᵃᵃᶜ); zf = parent(grid.z.cᵃᵃᶠ)                     # device OffsetVectors incl. halos
    return ObmGrid(Nx, Ny, Nz, Hx, Hy, Hz, 0, 0, 0, 0, pointer(zc), pointer(zf))
end

check(rc, what) = rc == 0 || error("$what failed ($rc): " * unsafe_string(ccall((:obm_last_error, libobm), Cstring, ())))

parents(fields, names) = [pointer(parent(fields[n])) for n in names]    # Vector{CuPtr{Float64}} — a HOST pointer table

# ---- the B200 biogeochemistry wrapper: same constructor surface, arithmetic in libobm ----------------------
struct B200Biogeochemistry{B, P} <: Oceananigans.Biogeochemistry.AbstractBiogeochemistry
    reference :: B     # the unmodified OceanBioME object (LOBSTER(grid; …), NPZD(grid; …), PISCES(; grid, …))
    params    :: P     # its parameters flattened into the C struct once, at construction
end

# forwarded unchanged: required_biogeochemical_tracers, required_biogeochemical_auxiliary_fields,
# biogeochemical_auxiliary_fields, biogeochemical_drift_velocity (src/OceanBioME.jl:122-129)

# per-point callable (compute_Gc! still calls it for every tracer): the fused kernel has already done the work
@inline (::B200Biogeochemistry)(i, j, k, grid, val_name, clock, fields) = zero(grid)

function update_biogeochemical_state!(bgc::B200Biogeochemistry, model)
    g, s = Ref(ObmGrid(model.grid)), CUDA.stream().handle
    ref = bgc.reference
    # 1. modifiers — all ScaleNegativeTracers groups in ONE launch (src/Utils/negative_tracers.jl:137-176)
    #    check(ccall((:obm_scale_negative_tracers, libobm), Cint, (Ref{ObmGrid}, Cint, Ptr{CuPtr{Float64}}, Cint, Ptr{Cvoid}, Cdouble, Ptr{Cvoid}), …))
    #    PISCES: the same launch also leaves Ω (step 3's obm_calcite_saturation is then skipped) —
    #    check(ccall((:obm_scale_negative_tracers_calcite_saturation, libobm), Cint,
    #                (Ref{ObmGrid}, Cint, Ptr{CuPtr{Float64}}, Cint, Ptr{Cvoid}, Cdouble, Ref{ObmCarbchemParams},
    #                 CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
    #                g, ntracers, tracer_table, ngroups, groups, NaN, carbchem, T, S, DIC, Alk, Si, Ω, H_state, s))
    # 2. light (src/Light/2band.jl:148-155): surface_PAR evaluated here into a scalar / 2-D field, then
    PAR = ref.light_attenuation
    surface = Float64(OceanBioME.Light.default_surface_PAR(model.clock.time))
    tb = Ref(twoband_params(PAR))                                          # obm_twoband_params, 8 doubles
    check(ccall((:obm_par_twoband, libobm), Cint,
                (Ref{ObmGrid}, Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
                g, tb, pointer(parent(model.tracers.P)), CU_NULL, surface, pointer(parent(PAR.field)), s),
          "obm_par_twoband")
    # 3. underlying — PISCES (PISCES/update_state.jl:1-17): zₑᵤ, mixed-layer means, Ω with the per-cell [H⁺] warm start
    if ref.underlying_biogeochemistry isa OceanBioME.Models.PISCESModel.PISCES
        u, t = ref.underlying_biogeochemistry, model.tracers
        check(ccall((:obm_euphotic_depth, libobm), Cint, (Ref{ObmGrid}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
                    g, pointer(parent(PAR.total)), 1 / 1000, pointer(parent(u.euphotic_depth)), s), "obm_euphotic_depth")
        check(ccall((:obm_mixed_layer_mean, libobm), Cint,
                    (Ref{ObmGrid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
                    g, pointer(parent(u.mixed_layer_depth)), pointer(parent(PAR.total)), 0.0,
                    pointer(parent(u.mean_mixed_layer_light)), s), "obm_mixed_layer_mean")
        check(ccall((:obm_calcite_saturation, libobm), Cint,
                    (Ref{ObmGrid}, Ref{ObmCarbchemParams}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64},
                     CuPtr{Float64}, CuPtr{Float64}, CuPtr{Float64}, Ptr{Cvoid}),
                    g, Ref(ObmCarbchemParams(8, 0, 8.0)), pointer(parent(t.T)), pointer(parent(t.S)), pointer(parent(t.DIC)),
                    pointer(parent(t.Alk)), pointer(parent(t.Si)), pointer(parent(u.calcite_saturation)),
                    pointer(parent(bgc.hydrogen_ion_state)), s), "obm_calcite_saturation")
    end
    # 4. sediment: obm_sediment_update_state(g, params, fields, model.clock.last_stage_Δt, χ, s) — the sediment's whole
    #    time_step! (AB2 step, or all three RK3 stages) in one launch (src/Sediments/update_state.jl:6-16)
    # 5. gas-exchange boundary conditions: obm_gas_exchange_flux into the flux field of each FluxBoundaryCondition
    return nothing
end

# PISCES: the 24 tendencies in one launch (replaces 26 compute_Gc! launches); `params` is an ObmPiscesParams filled once
# from the @kwdef structs, with the two clock-dependent day lengths refreshed here (growth_rate.jl:20-22,142-143)
function update_pisces_tendencies!(bgc::B200Biogeochemistry, model, params::ObmPiscesParams, fields::ObmPiscesFields)
    names   = Oceananigans.Biogeochemistry.required_biogeochemical_tracers(bgc.reference)   # PISCES.jl:94-105 order
    tracers = parents(model.tracers, names)
    G       = [n in (:T, :S) ? CU_NULL : pointer(parent(model.timestepper.Gⁿ[n])) for n in names]
    check(ccall((:obm_pisces_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmPiscesParams}, Ptr{CuPtr{Float64}}, Ref{ObmPiscesFields}, Ptr{CuPtr{Float64}}, Cint, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), Ref(params), tracers, Ref(fields), G, 1, CUDA.stream().handle),
          "obm_pisces_tendencies")
end

function update_tendencies!(bgc::B200Biogeochemistry, model)
    names   = Oceananigans.Biogeochemistry.required_biogeochemical_tracers(bgc.reference)
    tracers = parents(model.tracers, names)
    G       = [n === :T ? CU_NULL : pointer(parent(model.timestepper.Gⁿ[n])) for n in names]
    PAR     = pointer(parent(Oceananigans.Biogeochemistry.biogeochemical_auxiliary_fields(bgc.reference).PAR))
    check(ccall((:obm_npd_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Ptr{CuPtr{Float64}}, CuPtr{Float64}, Ptr{CuPtr{Float64}}, Cint, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), Ref(bgc.params), tracers, PAR, G, 1 #= accumulate: Gⁿ += … =#, CUDA.stream().handle),
          "obm_npd_tendencies")
    return nothing
end

# Parameter-sweep ensembles (examples/data_assimilation.jl builds one NPZD box model per parameter vector): one model
# whose columns are the members.  `which` = Int32[obm_npd_param_index("phytoplankton_maximum_growth_rate"), …],
# `values` = CuArray{Float64}(n_members, n_varied) (member fastest); everything else comes from bgc.params.
function update_tendencies!(bgc::B200Biogeochemistry, model, which::Vector{Int32}, values::CuMatrix{Float64})
    names   = Oceananigans.Biogeochemistry.required_biogeochemical_tracers(bgc.reference)
    tracers = parents(model.tracers, names)
    G       = [n === :T ? CU_NULL : pointer(parent(model.timestepper.Gⁿ[n])) for n in names]
    PAR     = pointer(parent(Oceananigans.Biogeochemistry.biogeochemical_auxiliary_fields(bgc.reference).PAR))
    check(ccall((:obm_npd_tendencies_ensemble, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Cint, Ptr{Int32}, CuPtr{Float64}, Ptr{CuPtr{Float64}}, CuPtr{Float64},
                 Ptr{CuPtr{Float64}}, Cint, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), Ref(bgc.params), length(which), which, pointer(values), tracers, PAR, G, 1,
                CUDA.stream().handle),
          "obm_npd_tendencies_ensemble")
    return nothing
end

# Particles (src/Particles): `update_tendencies!(bgc, particles::BiogeochemicalParticles{<:SugarKelp}, model)` →
# obm_kelp_update_tendencies (all 8 coupled tracers, one launch);  `time_step_particle_fields!(::ForwardEuler, …)` →
# obm_kelp_step.  ObmParticles carries pointer(particles.x), …, pointer(particles.fields.A), …, the first cell centre
# and spacing in x and y, and the topology codes; ObmKelpTracers the parents of u, v, w, T, NO₃, NH₄ and PAR.
#
# Column / box models without resolved flow: `obm_sinking_tendencies(g, n, tracers, w_faces, Gⁿ, scheme, 1, s)` adds
# −∂z(w c) of every sinking tracer (w = biogeochemical_drift_velocity(bgc, Val(c)).w) in one launch.

# ---- raw bindings of the remaining entry points (include/obm_b200.h), one thin method each ---------------------------
# Device arrays are passed as CuPtr{Float64} (`pointer(parent(field))`), tables of them as host Vector{CuPtr{Float64}},
# parameter blocks by Ref; `s` is `CUDA.stream().handle`.  tests/test_abi.py checks every signature's arity against the header.
const F64 = CuPtr{Float64}

par_multiband!(g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, s) =
    check(ccall((:obm_par_multiband, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmMultibandParams}, F64, F64, Cdouble, F64, Cdouble, Ptr{F64}, F64, Ptr{Cvoid}),
                g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, s), "obm_par_multiband")

# PISCES: the PAR scan that also leaves zₑᵤ and the mixed-layer mean PAR (replaces three of the reference's launches)
par_multiband_column_state!(g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, zmxl, cutoff, zeu, mean_par, s) =
    check(ccall((:obm_par_multiband_column_state, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmMultibandParams}, F64, F64, Cdouble, F64, Cdouble, Ptr{F64}, F64, F64, Cdouble, F64, F64, Ptr{Cvoid}),
                g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, zmxl, cutoff, zeu, mean_par, s),
          "obm_par_multiband_column_state")

sediment_update_state!(g, p, f, Δt, χ, s) =
    check(ccall((:obm_sediment_update_state, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSedimentParams}, Ref{ObmSedimentFields}, Cdouble, Cdouble, Ptr{Cvoid}), g, p, f, Δt, χ, s),
          "obm_sediment_update_state")

sediment_update_tendencies!(g, p, f, s) =
    check(ccall((:obm_sediment_update_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSedimentParams}, Ref{ObmSedimentFields}, Ptr{Cvoid}), g, p, f, s),
          "obm_sediment_update_tendencies")

find_bottom_cells!(g, bottom_height_xy, bottom_indices_xy::CuPtr{Int64}, s) =
    check(ccall((:obm_find_bottom_cells, libobm), Cint, (Ref{ObmGrid}, F64, CuPtr{Int64}, Ptr{Cvoid}),
                g, bottom_height_xy, bottom_indices_xy, s), "obm_find_bottom_cells")

# gas exchange: the flux field a FluxBoundaryCondition then reads; DIC / Alk only for CO₂, silicate / phosphate optional
gas_exchange_flux!(g, p, T, S, tracer, DIC, Alk, silicate, phosphate, wind_xy, air_xy, flux_xy, G_top, s) =
    check(ccall((:obm_gas_exchange_flux, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmGasExchangeParams}, F64, F64, F64, F64, F64, F64, F64, F64, F64, F64, F64, Ptr{Cvoid}),
                g, p, T, S, tracer, DIC, Alk, silicate, phosphate, wind_xy, air_xy, flux_xy, G_top, s), "obm_gas_exchange_flux")

kelp_update_tendencies!(g, p, particles, tracers, G, t, s) =
    check(ccall((:obm_kelp_update_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSugarKelpParams}, Ref{ObmParticles}, Ref{ObmKelpTracers}, Ptr{F64}, Cdouble, Ptr{Cvoid}),
                g, p, particles, tracers, G, t, s), "obm_kelp_update_tendencies")

kelp_step!(g, p, particles, tracers, t, Δt, tendencies_out, s) =
    check(ccall((:obm_kelp_step, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSugarKelpParams}, Ref{ObmParticles}, Ref{ObmKelpTracers}, Cdouble, Cdouble, Ptr{F64}, Ptr{Cvoid}),
                g, p, particles, tracers, t, Δt, tendencies_out, s), "obm_kelp_step")

sinking_tendencies!(g, tracers, w_faces, G, advection, accumulate, s) =
    check(ccall((:obm_sinking_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Ptr{F64}, Ptr{F64}, Cint, Cint, Ptr{Cvoid}),
                g, length(tracers), tracers, w_faces, G, advection, accumulate, s), "obm_sinking_tendencies")

# box / column models: U += Δt (γ Gⁿ + ζ G⁻) and G⁻ ← Gⁿ for every field in one launch (src/BoxModel/timesteppers.jl:66-93)
rk3_substep!(g, U, Gⁿ, G⁻, Δt, γ, ζ, has_ζ, cache_previous, s) =
    check(ccall((:obm_rk3_substep, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Ptr{F64}, Ptr{F64}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Cvoid}),
                g, length(U), U, Gⁿ, G⁻, Δt, γ, ζ, has_ζ, cache_previous, s), "obm_rk3_substep")

# CarbonChemistry()(; DIC, Alk, T, S, …) over flat device arrays (src/Models/CarbonChemistry/carbon_chemistry.jl:84-181)
carbon_chemistry!(out, p, T, S, DIC, Alk, P_bar, silicate, phosphate, pH, output_kind, n, s) =
    check(ccall((:obm_carbon_chemistry, libobm), Cint,
                (Int64, Ref{ObmCarbchemParams}, F64, F64, F64, F64, F64, F64, F64, F64, Cint, F64, Ptr{Cvoid}),
                n, p, T, S, DIC, Alk, P_bar, silicate, phosphate, pH, output_kind, out, s), "obm_carbon_chemistry")

zero_negative_tracers!(n_parent, tracers, s) =
    check(ccall((:obm_zero_negative_tracers, libobm), Cint, (Int64, Cint, Ptr{F64}, Ptr{Cvoid}),
                n_parent, length(tracers), tracers, s), "obm_zero_negative_tracers")

# conservation diagnostics: Σ sf·c·V per conserved group on this GPU; follow with an all-reduce over the ranks
inventory!(out, g, tracers, groups, cell_volume, uniform_volume, workspace, s) =
    check(ccall((:obm_inventory, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Cint, Ptr{ObmScaleGroup}, F64, Cdouble, F64, CuPtr{Cvoid}, Ptr{Cvoid}),
                g, length(tracers), tracers, length(groups), groups, cell_volume, uniform_volume, out, workspace, s),
          "obm_inventory")

end # module
