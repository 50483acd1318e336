# OceanBioMEB200.jl — the reference-side binding a maintainer would add: `ccall` glue from the plugin hooks OceanBioME.jl
# implements (src/OceanBioME.jl:53-59,122-169) to libobm_b200.so (include/obm_b200.h).
#
#   using OceanBioME, Oceananigans, OceanBioMEB200
#   biogeochemistry = B200(LOBSTER(; grid, carbonates = true, scale_negatives = true), grid)     # or NPZD(…), PISCES(; grid, …)
#   model = NonhydrostaticModel(; grid, biogeochemistry, …)                                      # drops in unchanged
#
# `B200(bgc, grid)` wraps the UNMODIFIED OceanBioME object (its constructors, tracer lists, auxiliary fields, drift
# velocities, boundary conditions and show methods keep working) and replaces the arithmetic of the two hooks Oceananigans
# calls every stage — `update_biogeochemical_state!` and `update_tendencies!` — by the fused kernels; the per-point callable
# that `compute_Gc!` still evaluates for every tracer returns `zero(grid)` because `update_tendencies!` has added every
# tendency to Gⁿ already.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: there is no Julia in the build image.  What is checked mechanically instead
# (tests/test_abi.py): every `ccall` names an exported symbol with the prototype's arity and scalar / pointer kinds; the
# struct mirrors (obm_structs.jl) are generated from the ctypes definitions whose sizes the compiled library confirms;
# the parameter-fill constructors (obm_fill.jl) are generated from the Python host mirror's own `c_params()` — the code
# path of every GPU parity test — and compared field by field with the reference's struct definitions.
module OceanBioMEB200

export B200, B200Biogeochemistry, tracer_inventory

using CUDA, Oceananigans, OceanBioME
using Oceananigans.Architectures: architecture
using Oceananigans.BoundaryConditions: getbc
using Oceananigans.Grids: halo_size, topology, Flat, Center
using Oceananigans.Fields: Field, CenterField, ConstantField
using OffsetArrays: OffsetArray
using Oceananigans.Utils: launch!
using KernelAbstractions: @kernel, @index

using OceanBioME: ContinuousBiogeochemistry, DiscreteBiogeochemistry, CompleteBiogeochemistry, ScaleNegativeTracers, ZeroNegativeTracers
using OceanBioME.Models.NutrientsPlanktonDetritusModels: NutrientsPlanktonDetritus, Nutrient, NitrateAmmonia, NitrateAmmoniaIron,
    Detritus, TwoParticleAndDissolved, VariableRedfieldDetritus, CarbonateSystem, Linear, Quadratic, MondoLightLimitation
using OceanBioME.Models.PISCESModel: PISCES
using OceanBioME.Models.PISCESModel.Phytoplankton: NutrientLimitedProduction, GrowthRespirationLimitedProduction
using OceanBioME.Light: TwoBandPhotosyntheticallyActiveRadiation, MultiBandPhotosyntheticallyActiveRadiation
using OceanBioME.Sediments: BiogeochemicalSediment, sinking_fluxes, coupled_tracers, required_tracers
using OceanBioME.Models.SedimentModels: SimpleMultiG, InstantRemineralisation

import Oceananigans.Biogeochemistry: required_biogeochemical_tracers, required_biogeochemical_auxiliary_fields,
    biogeochemical_auxiliary_fields, biogeochemical_drift_velocity, update_biogeochemical_state!, update_tendencies!
import OceanBioME: chlorophyll

const libobm = get(ENV, "OBM_B200_LIB", "libobm_b200.so")
const F64 = CuPtr{Float64}

# ---- struct mirrors and their fill constructors (both generated, see the header) ----------------------------------------
include("obm_structs.jl")   # ObmGrid, ObmNpdParams, ObmTwobandParams, ObmMultibandParams, ObmCarbchemParams, ObmScaleGroup,
                            # ObmPiscesPhyto / Zoo / Params / Fields, ObmSedimentParams / Fields, ObmGasExchangeParams, …

# helpers the generated constructors use
fieldor(x, f::Symbol) = (hasproperty(x, f) && !isnothing(getproperty(x, f))) ? Float64(getproperty(x, f)) : 0.0
# the scalar latitude of a PrescribedLatitude (PISCES/common.jl:20-25); 0.0 for ModelLatitude, whose per-row values travel separately
prescribed_latitude(bgc) = hasproperty(bgc.latitude, :latitude) ? Float64(bgc.latitude.latitude) : 0.0
tupleor(x, f::Symbol, n) = (hasproperty(x, f) && length(getproperty(x, f)) ≥ n) ? Float64(getproperty(x, f)[n]) : 0.0
bandor(v, n) = n ≤ length(v) ? Float64(v[n]) : 0.0
# enumerations of include/obm_b200.h, chosen by the TYPE of the reference component
nutrient_kind(::Nutrient) = Int32(0); nutrient_kind(::NitrateAmmonia) = Int32(1); nutrient_kind(::NitrateAmmoniaIron) = Int32(2)
detritus_kind(::Nothing) = Int32(0); detritus_kind(::Detritus) = Int32(1)
detritus_kind(::TwoParticleAndDissolved) = Int32(2); detritus_kind(::VariableRedfieldDetritus) = Int32(3)
carbonate_replicates(::Nothing) = Int32(0); carbonate_replicates(::CarbonateSystem{N}) where N = Int32(N)   # carbonate_system.jl:39
light_limitation_kind(::MondoLightLimitation) = Int32(0); light_limitation_kind(_) = Int32(1)              # plankton.jl:216-220
formulation_kind(::Linear) = Int32(0); formulation_kind(::Quadratic) = Int32(1)                              # plankton.jl:83-90
growth_rate_kind(::NutrientLimitedProduction) = Int32(0); growth_rate_kind(::GrowthRespirationLimitedProduction) = Int32(1)
sediment_kind(::InstantRemineralisation) = Int32(0); sediment_kind(::SimpleMultiG) = Int32(1)
has_carbon(b::SimpleMultiG) = !isnothing(b.sinking_carbon) && length(b.sinking_carbon) > 0; has_carbon(::InstantRemineralisation) = false
sinking_nitrogen(b::SimpleMultiG) = b.sinking_nitrogen; sinking_nitrogen(b::InstantRemineralisation) = b.sinking_tracers
sinking_carbon(b::SimpleMultiG) = has_carbon(b) ? b.sinking_carbon : (); sinking_carbon(::InstantRemineralisation) = ()
timestepper_kind(ts) = Int32(ts isa Oceananigans.TimeSteppers.RungeKutta3TimeStepper)                       # OBM_TS_AB2 = 0, OBM_TS_RK3 = 1
advection_kind(::Oceananigans.Advection.Centered) = Int32(1)                                                 # OBM_ADV_CENTERED2
advection_kind(_) = Int32(0)   # OBM_ADV_UPWIND1: the sediment's bottom-face flux is first order for every upwind-biased scheme (WENO included)
# the interior faces of obm_sinking_tendencies (column / box models): UpwindBiased(order = 3) → 2, WENO(order = 5) → 3
sinking_advection_kind(a::Oceananigans.Advection.UpwindBiased) = Int32(a isa Oceananigans.Advection.UpwindBiased{1} ? 0 : 2)
sinking_advection_kind(::Oceananigans.Advection.WENO) = Int32(3)
sinking_advection_kind(a) = advection_kind(a)

include("obm_fill.jl")      # ObmNpdParams(bgc), ObmPiscesParams(bgc, clock), ObmTwobandParams(par), ObmMultibandParams(par),
                            # ObmSedimentParams(sed, advection)

function ObmGrid(grid)
    Nx, Ny, Nz = size(grid)
    Hx, Hy, Hz = halo_size(grid)
    tx, ty, tz = topology(grid)
    Hx, Hy, Hz = (tx === Flat ? 0 : Hx), (ty === Flat ? 0 : Hy), (tz === Flat ? 0 : Hz)   # Flat dimensions carry no halo
    zc = parent(grid.z.cᵃᵃᶜ); zf = parent(grid.z.cᵃᵃᶠ)                                     # device vectors incl. halos
    return ObmGrid(; Nx, Ny, Nz, Hx, Hy, Hz, zc = pointer(zc), zf = pointer(zf))
end

check(rc, what) = rc == 0 || error("$what failed ($rc): " * unsafe_string(ccall((:obm_last_error, libobm), Cstring, ())))
dptr(field) = pointer(parent(field))                       # CuPtr{Float64} of a field's parent array (halos included)
dptr(::Nothing) = CU_NULL
table(fields) = F64[dptr(f) for f in fields]               # a HOST table of device pointers, read at call time
stream() = CUDA.stream().handle                            # the caller's stream: our launches order with Oceananigans' own

# ---- the wrapper -------------------------------------------------------------------------------------------------------------
struct B200Biogeochemistry{B, S} <: Oceananigans.Biogeochemistry.AbstractBiogeochemistry
    reference :: B     # the unmodified OceanBioME object: Biogeochemistry(LOBSTER / NPZD / PISCES …; light_attenuation, sediment, …)
    scratch   :: S     # device buffers the kernels need beside the model's own fields (allocated once, here)
end

"""
    B200(biogeochemistry, grid)

Wrap a complete OceanBioME biogeochemistry (what `LOBSTER(; grid, …)`, `NPZD(; grid, …)`, `PISCES(; grid, …)` or
`Biogeochemistry(underlying; light_attenuation, sediment, particles, modifiers)` return) so that its two per-stage hooks
run on libobm_b200's fused kernels.
"""
function B200(bgc::CompleteBiogeochemistry, grid)
    plane() = CUDA.zeros(Float64, size(parent(Field{Center, Center, Nothing}(grid)))...)
    scratch = (surface_PAR = plane(),                                     # getbc(surface_PAR, i, j, …) of every column (2band.jl:4)
               hydrogen_ion = bgc.underlying_biogeochemistry isa PISCES ? CUDA.zeros(Float64, size(parent(CenterField(grid)))...) : nothing,
               inventory = CUDA.zeros(Float64, 8),
               inventory_workspace = CUDA.zeros(Float64, ccall((:obm_inventory_workspace_bytes, libobm), Int64, (Cint,), 8) ÷ 8))
    return B200Biogeochemistry(bgc, scratch)
end

const B200NPD    = B200Biogeochemistry{<:CompleteBiogeochemistry{<:NutrientsPlanktonDetritus}}
const B200PISCES = B200Biogeochemistry{<:CompleteBiogeochemistry{<:PISCES}}

# forwarded unchanged (src/OceanBioME.jl:122-131): what Oceananigans asks when it builds the model
required_biogeochemical_tracers(b::B200Biogeochemistry)          = required_biogeochemical_tracers(b.reference)
required_biogeochemical_auxiliary_fields(b::B200Biogeochemistry) = required_biogeochemical_auxiliary_fields(b.reference)
biogeochemical_auxiliary_fields(b::B200Biogeochemistry)          = biogeochemical_auxiliary_fields(b.reference)
biogeochemical_drift_velocity(b::B200Biogeochemistry, val_name)  = biogeochemical_drift_velocity(b.reference, val_name)
chlorophyll(b::B200Biogeochemistry, model)                       = chlorophyll(b.reference, model)
Base.summary(b::B200Biogeochemistry) = "B200 kernels behind " * summary(b.reference)
Base.show(io::IO, b::B200Biogeochemistry) = (print(io, "libobm_b200 → "); show(io, b.reference))

# The discrete per-point form `bgc(i, j, k, grid, Val(name), clock, fields)` — what `compute_Gc!` evaluates for every tracer
# of an AbstractBiogeochemistry (Oceananigans `biogeochemical_transition`): the fused launch of `update_tendencies!` has
# already added the tendency, so nothing is left to add.
@inline (::B200Biogeochemistry)(i, j, k, grid, val_name, clock, fields) = zero(grid)

# The continuous per-tracer form `bgc(Val(:P), x, y, z, t, fields...)` (docs/src/model_implementation.md:34-75; the form
# BoxModel evaluates, src/BoxModel/timesteppers.jl:60-61): ONE tracer's tendency at ONE state.  Evaluated by the same fused
# kernel on a single cell (all tendencies are computed, the named one is returned) — a host-side convenience for box
# models and for inspecting a tendency; inside 3-D models the hooks below are the path.  `fields` are the values of
# `(required_biogeochemical_tracers(bgc)..., required_biogeochemical_auxiliary_fields(bgc)...)` in that order.
function (b::B200NPD)(::Val{name}, x, y, z, t, fields...) where name
    u = b.reference.underlying_biogeochemistry
    names = required_biogeochemical_tracers(u)
    values = CuArray{Float64}[CuArray([Float64(v)]) for v in fields[1:length(names)]]
    PAR = CuArray([Float64(fields[length(names) + 1])])                                  # the one auxiliary field, plankton.jl:81
    G = CuArray{Float64}[CUDA.zeros(Float64, 1) for _ in names]
    zn = CuArray([Float64(z) - 0.5, Float64(z) + 0.5])
    g = Ref(ObmGrid(; Nx = 1, Ny = 1, Nz = 1, zc = pointer(CuArray([Float64(z)])), zf = pointer(zn)))
    check(ccall((:obm_npd_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Ptr{F64}, F64, Ptr{F64}, Cint, Ptr{Cvoid}),
                g, Ref(ObmNpdParams(u)), pointer.(values), pointer(PAR), pointer.(G), 0, stream()), "obm_npd_tendencies")
    return Array(G[findfirst(==(name), names)])[1]
end

function (b::B200PISCES)(::Val{name}, x, y, z, t, fields...) where name
    u = b.reference.underlying_biogeochemistry
    names = required_biogeochemical_tracers(u)                                           # PISCES.jl:94-105
    aux = NamedTuple{required_biogeochemical_auxiliary_fields(u)}(fields[length(names)+1:end])   # PISCES.jl:107-108
    one(v) = CuArray([Float64(v)])
    values = CuArray{Float64}[one(v) for v in fields[1:length(names)]]
    G = CuArray{Float64}[CUDA.zeros(Float64, 1) for _ in names]
    w(v) = CuArray([Float64(v), Float64(v)])                                             # two equal faces: ℑzᵃᵃᶜ(w) = v
    hold = (PAR₁ = one(aux.PAR₁), PAR₂ = one(aux.PAR₂), PAR₃ = one(aux.PAR₃), PAR = one(aux.PAR), Ω = one(aux.Ω),
            wPOC = w(aux.wPOC), wGOC = w(aux.wGOC), zₘₓₗ = one(aux.zₘₓₗ), zₑᵤ = one(aux.zₑᵤ), κ = one(aux.κ), mlPAR = one(aux.mixed_layer_PAR))
    f = ObmPiscesFields(; PAR1 = pointer(hold.PAR₁), PAR2 = pointer(hold.PAR₂), PAR3 = pointer(hold.PAR₃), PAR = pointer(hold.PAR),
                        Omega = pointer(hold.Ω), wPOC = pointer(hold.wPOC), wGOC = pointer(hold.wGOC),
                        mixed_layer_depth_xy = pointer(hold.zₘₓₗ), euphotic_depth_xy = pointer(hold.zₑᵤ),
                        mean_mixed_layer_vertical_diffusivity_xy = pointer(hold.κ), mean_mixed_layer_light_xy = pointer(hold.mlPAR))
    zn = CuArray([Float64(z) - 0.5, Float64(z) + 0.5])
    g = Ref(ObmGrid(; Nx = 1, Ny = 1, Nz = 1, zc = pointer(CuArray([Float64(z)])), zf = pointer(zn)))
    Gp = F64[n in (:T, :S) ? CU_NULL : pointer(G[i]) for (i, n) in enumerate(names)]
    check(ccall((:obm_pisces_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmPiscesParams}, Ptr{F64}, Ref{ObmPiscesFields}, Ptr{F64}, Cint, Ptr{Cvoid}),
                g, Ref(ObmPiscesParams(u, (; time = t))), pointer.(values), Ref(f), Gp, 0, stream()), "obm_pisces_tendencies")
    return name in (:T, :S) ? 0.0 : Array(G[findfirst(==(name), names)])[1]            # zero(grid), PISCES.jl:120
end

# ---- update_biogeochemical_state!(bgc, model): modifiers → light → underlying → sediment (src/OceanBioME.jl:161-167) ---------
function update_biogeochemical_state!(b::B200Biogeochemistry, model)
    ref = b.reference
    g = Ref(ObmGrid(model.grid))
    Ω_done = update_modifiers!(b, model, g)
    light_done = update_light!(b, ref.light_attenuation, model, g)
    update_underlying!(b, ref.underlying_biogeochemistry, model, g, Ω_done, light_done)
    update_sediment!(b, ref.sediment, model, g)
    return nothing
end

# 1. modifiers (src/OceanBioME.jl:163,169).  Every ScaleNegativeTracers of the tuple — one launch per conserved group in the
#    reference (src/Utils/negative_tracers.jl:137-176) — goes into ONE launch, groups applied in tuple order; for PISCES
#    the same launch solves Ω of `compute_calcite_saturation!` from the rescaled DIC, Alk, Si (PISCES/update_state.jl:13).
#    Any other modifier keeps the reference's own method.
function scale_groups(scalers, names)
    groups = ObmScaleGroup[]
    for s in scalers
        idx = ntuple(n -> n ≤ length(s.tracers) ? Int32(findfirst(==(s.tracers[n]), names) - 1) : Int32(0), 16)
        sf  = ntuple(n -> n ≤ length(s.tracers) ? Float64(s.scalefactors[n]) : 0.0, 16)
        push!(groups, ObmScaleGroup(; n = Int32(length(s.tracers)), index = idx, scalefactor = sf))
    end
    return groups
end

function update_modifiers!(b, model, g)
    mods = b.reference.modifiers
    mods = isnothing(mods) ? () : (mods isa Tuple ? mods : (mods,))
    scalers = filter(m -> m isa ScaleNegativeTracers, mods)
    for m in mods
        m isa ZeroNegativeTracers && zero_negative_tracers!(m, model)
        m isa Union{ScaleNegativeTracers, ZeroNegativeTracers} || update_biogeochemical_state!(model, m)
    end
    isempty(scalers) && return false
    names = unique(vcat((collect(s.tracers) for s in scalers)...))       # distinct scaled tracers, first occurrence order
    tracers = table(model.tracers[n] for n in names)
    groups = scale_groups(scalers, names)
    fill_value = Float64(first(scalers).invalid_fill_value)
    u = b.reference.underlying_biogeochemistry
    if u isa PISCES
        t = model.tracers
        check(ccall((:obm_scale_negative_tracers_calcite_saturation, libobm), Cint,
                    (Ref{ObmGrid}, Cint, Ptr{F64}, Cint, Ptr{ObmScaleGroup}, Cdouble, Ref{ObmCarbchemParams},
                     F64, F64, F64, F64, F64, F64, F64, Ptr{Cvoid}),
                    g, length(names), tracers, length(groups), groups, fill_value, Ref(ObmCarbchemParams(; newton_iterations = Int32(12), initial_pH_guess = 8.0)),
                    dptr(t.T), dptr(t.S), dptr(t.DIC), dptr(t.Alk), dptr(t.Si), dptr(u.calcite_saturation),
                    dptr(b.scratch.hydrogen_ion), stream()), "obm_scale_negative_tracers_calcite_saturation")
        return true
    end
    check(ccall((:obm_scale_negative_tracers, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Cint, Ptr{ObmScaleGroup}, Cdouble, Ptr{Cvoid}),
                g, length(names), tracers, length(groups), groups, fill_value, stream()), "obm_scale_negative_tracers")
    return false
end

function zero_negative_tracers!(m::ZeroNegativeTracers, model)   # negative_tracers.jl:26-32: every tracer but the excluded ones
    fields = [f for (n, f) in pairs(model.tracers) if !(n in m.exclude)]
    check(ccall((:obm_zero_negative_tracers, libobm), Cint, (Int64, Cint, Ptr{F64}, Ptr{Cvoid}),
                length(parent(first(fields))), length(fields), table(fields), stream()), "obm_zero_negative_tracers")
end

# 2. light.  The surface PAR boundary function (a constant, a Field, a continuous f(x, y, t) or a discrete
#    f(i, j, grid, clock, fields): all four are `getbc`-able, 2band.jl:4, multi_band.jl:151) cannot cross the C ABI, so it is
#    evaluated for every column into a 2-D device array by this three-line kernel, exactly as the reference kernels do.
@kernel function _surface_values!(out, grid, clock, surface_PAR, args)
    i, j = @index(Global, NTuple)
    @inbounds out[i, j, 1] = getbc(surface_PAR, i, j, grid, clock, args)
end

function surface_PAR!(b, par, model)
    Hx, Hy, _ = halo_size(model.grid)
    out = OffsetArray(b.scratch.surface_PAR, -Hx, -Hy, 0)              # indexed like a Field{Center, Center, Nothing}
    launch!(architecture(model.grid), model.grid, :xy, _surface_values!, out, model.grid, model.clock, par.surface_PAR,
            par isa TwoBandPhotosyntheticallyActiveRadiation ? model.tracers.P : Oceananigans.fields(model))
    return pointer(b.scratch.surface_PAR)
end

update_light!(b, ::Nothing, model, g) = false
update_light!(b, par, model, g) = (update_biogeochemical_state!(model, par); false)       # e.g. PrescribedPhotosyntheticallyActiveRadiation

function update_light!(b, par::TwoBandPhotosyntheticallyActiveRadiation, model, g)        # 2band.jl:148-155
    check(ccall((:obm_par_twoband, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmTwobandParams}, F64, F64, Cdouble, F64, Ptr{Cvoid}),
                g, Ref(ObmTwobandParams(par)), dptr(model.tracers.P), surface_PAR!(b, par, model), 0.0, dptr(par.field), stream()),
          "obm_par_twoband")
    return false
end

function update_light!(b, par::MultiBandPhotosyntheticallyActiveRadiation, model, g)      # multi_band.jl:165-185, all bands in one launch
    chl = chlorophyll(b.reference, model)                                                  # a field, or a sum of fields (PISCES: PChl + DChl, coupling_utils.jl:7)
    chl_a, chl_b = chl isa Oceananigans.AbstractOperations.BinaryOperation ? (chl.a, chl.b) : (chl, nothing)
    bands, surface = table(par.fields), surface_PAR!(b, par, model)
    p = Ref(ObmMultibandParams(par))
    u = b.reference.underlying_biogeochemistry
    if u isa PISCES && !(u.euphotic_depth isa ConstantField) && !(u.mean_mixed_layer_light isa ConstantField)
        # the scan also leaves zₑᵤ (compute_euphotic_depth.jl:3-40, cutoff 1/1000) and PAR̄ₘₓₗ (mean_mixed_layer_properties.jl:10-49)
        check(ccall((:obm_par_multiband_column_state, libobm), Cint,
                    (Ref{ObmGrid}, Ref{ObmMultibandParams}, F64, F64, Cdouble, F64, Cdouble, Ptr{F64}, F64, F64, Cdouble, F64, F64, Ptr{Cvoid}),
                    g, p, dptr(chl_a), dptr(chl_b), 1.0, surface, 0.0, bands, dptr(par.total), dptr(u.mixed_layer_depth), 1 / 1000,
                    dptr(u.euphotic_depth), dptr(u.mean_mixed_layer_light), stream()), "obm_par_multiband_column_state")
        return true
    end
    check(ccall((:obm_par_multiband, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmMultibandParams}, F64, F64, Cdouble, F64, Cdouble, Ptr{F64}, F64, Ptr{Cvoid}),
                g, p, dptr(chl_a), dptr(chl_b), 1.0, surface, 0.0, bands, dptr(par.total), stream()), "obm_par_multiband")
    return false
end

# 3. underlying biogeochemistry.  NPD models have no state of their own; PISCES: PISCES/update_state.jl:1-17.
update_underlying!(b, u, model, g, Ω_done, light_done) = nothing

function update_underlying!(b, u::PISCES, model, g, Ω_done, light_done)
    PAR = biogeochemical_auxiliary_fields(b.reference.light_attenuation).PAR
    if !light_done
        u.euphotic_depth isa ConstantField ||
            check(ccall((:obm_euphotic_depth, libobm), Cint, (Ref{ObmGrid}, F64, Cdouble, F64, Ptr{Cvoid}),
                        g, dptr(PAR), 1 / 1000, dptr(u.euphotic_depth), stream()), "obm_euphotic_depth")
        u.mean_mixed_layer_light isa ConstantField ||
            check(ccall((:obm_mixed_layer_mean, libobm), Cint, (Ref{ObmGrid}, F64, F64, Cdouble, F64, Ptr{Cvoid}),
                        g, dptr(u.mixed_layer_depth), dptr(PAR), 0.0, dptr(u.mean_mixed_layer_light), stream()), "obm_mixed_layer_mean")
    end
    # κ̄ over the mixed layer needs the closure's diffusivity field: the reference's own method finds it (mean_mixed_layer_properties.jl:23-67)
    OceanBioME.Models.PISCESModel.compute_mean_mixed_layer_vertical_diffusivity!(u.mean_mixed_layer_vertical_diffusivity, u.mixed_layer_depth, model)
    if !Ω_done
        t = model.tracers
        check(ccall((:obm_calcite_saturation, libobm), Cint,
                    (Ref{ObmGrid}, Ref{ObmCarbchemParams}, F64, F64, F64, F64, F64, F64, F64, Ptr{Cvoid}),
                    g, Ref(ObmCarbchemParams(; newton_iterations = Int32(12), initial_pH_guess = 8.0)), dptr(t.T), dptr(t.S), dptr(t.DIC),
                    dptr(t.Alk), dptr(t.Si), dptr(u.calcite_saturation), dptr(b.scratch.hydrogen_ion), stream()), "obm_calcite_saturation")
    end
    return nothing
end

# 4. sediment (src/Sediments/update_state.jl:6-16: tracked fields, then the sediment's own time_step!) in one launch
function sediment_fields(b, sed::BiogeochemicalSediment, model)
    bgc, t = sed.biogeochemistry, model.tracers
    pad(v, n, z) = ntuple(i -> i ≤ length(v) ? v[i] : z, n)
    sinking = collect(sinking_fluxes(bgc))
    w(n) = biogeochemical_drift_velocity(b.reference, Val(n)).w                           # tracked_fields.jl:52-62
    has = length(required_tracers(bgc)) > 0
    Gⁿ, G⁻ = sed.timestepper.Gⁿ, sed.timestepper.G⁻
    return ObmSedimentFields(;
        bottom_indices_xy = reinterpret(F64, pointer(parent(sed.bottom_indices))),        # Int64 plane (bottom_indices.jl:19-26)
        NO3 = has ? dptr(t.NO₃) : CU_NULL, NH4 = has ? dptr(t.NH₄) : CU_NULL, O2 = has ? dptr(t.O₂) : CU_NULL,
        sinking   = pad([dptr(t[n]) for n in sinking], 8, CU_NULL),
        sinking_w = pad([dptr(w(n)) for n in sinking], 8, CU_NULL),
        pools = pad([dptr(f) for f in sed.fields], 6, CU_NULL),
        Gn = pad([dptr(f) for f in Gⁿ], 6, CU_NULL), Gm = pad([dptr(f) for f in G⁻], 6, CU_NULL),
        tracked_xy = pad([dptr(f) for f in sed.tracked_fields], 11, CU_NULL),
        G_coupled = pad([dptr(model.timestepper.Gⁿ[n]) for n in coupled_tracers(bgc)], 4, CU_NULL))
end

update_sediment!(b, ::Nothing, model, g) = nothing
function update_sediment!(b, sed::BiogeochemicalSediment, model, g)
    Δt = model.clock.last_stage_Δt
    χ = sed.timestepper isa Oceananigans.TimeSteppers.QuasiAdamsBashforth2TimeStepper ? Float64(sed.timestepper.χ) : 0.0
    check(ccall((:obm_sediment_update_state, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSedimentParams}, Ref{ObmSedimentFields}, Cdouble, Cdouble, Ptr{Cvoid}),
                g, Ref(ObmSedimentParams(sed, model.advection)), Ref(sediment_fields(b, sed, model)), Float64(Δt), χ, stream()),
          "obm_sediment_update_state")
end

# ---- update_tendencies!(bgc, model) (src/OceanBioME.jl:148-152): every tendency of every tracer added to Gⁿ in one launch,
#      then sediment ↔ tracer fluxes (Sediments/tracer_coupling.jl:3-39), particles and modifiers as in the reference ---------
function update_tendencies!(b::B200Biogeochemistry, model)
    g = Ref(ObmGrid(model.grid))
    add_tendencies!(b, b.reference.underlying_biogeochemistry, model, g)
    sed = b.reference.sediment
    isnothing(sed) ||
        check(ccall((:obm_sediment_update_tendencies, libobm), Cint,
                    (Ref{ObmGrid}, Ref{ObmSedimentParams}, Ref{ObmSedimentFields}, Ptr{Cvoid}),
                    g, Ref(ObmSedimentParams(sed, model.advection)), Ref(sediment_fields(b, sed, model)), stream()),
              "obm_sediment_update_tendencies")
    update_tendencies!(b.reference, b.reference.particles, model)        # kelp: see kelp_update_tendencies! below for the fused form
    update_tendencies!(b.reference, b.reference.modifiers, model)
    return nothing
end

function add_tendencies!(b, u::NutrientsPlanktonDetritus, model, g)   # replaces one compute_Gc! pass per tracer (10 for LOBSTER + carbonates + O₂)
    names = required_biogeochemical_tracers(u)
    G = F64[n === :T ? CU_NULL : dptr(model.timestepper.Gⁿ[n]) for n in names]
    check(ccall((:obm_npd_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Ptr{F64}, F64, Ptr{F64}, Cint, Ptr{Cvoid}),
                g, Ref(ObmNpdParams(u)), table(model.tracers[n] for n in names), dptr(biogeochemical_auxiliary_fields(b.reference).PAR),
                G, 1 #= accumulate: Gⁿ += … =#, stream()), "obm_npd_tendencies")
end

function add_tendencies!(b, u::PISCES, model, g)                       # replaces 24 compute_Gc! passes (PISCES.jl:120-123)
    names = required_biogeochemical_tracers(u)                          # PISCES.jl:94-105 order = OBM_PISCES_NTRACERS order
    aux = biogeochemical_auxiliary_fields(b.reference)
    f = ObmPiscesFields(; PAR1 = dptr(aux.PAR₁), PAR2 = dptr(aux.PAR₂), PAR3 = dptr(aux.PAR₃), PAR = dptr(aux.PAR), Omega = dptr(aux.Ω),
                        wPOC = dptr(aux.wPOC), wGOC = dptr(aux.wGOC), mixed_layer_depth_xy = dptr(aux.zₘₓₗ),
                        euphotic_depth_xy = dptr(aux.zₑᵤ), mean_mixed_layer_vertical_diffusivity_xy = dptr(aux.κ),
                        mean_mixed_layer_light_xy = dptr(aux.mixed_layer_PAR))
    G = F64[n in (:T, :S) ? CU_NULL : dptr(model.timestepper.Gⁿ[n]) for n in names]
    if u.latitude isa OceanBioME.Models.PISCESModel.ModelLatitude       # PISCES/common.jl:27-28: φ = φnode(i, j, k, grid) — a row's own day lengths
        φ = Float64.(φnodes(model.grid, Center(), Center(), Center()))  # the Ny interior rows
        t = model.clock.time
        rows = CuArray(hcat(φ, Float64[u.day_length(x, t) for x in φ],  # growth_rate.jl:29-30 (the swapped call)
                            Float64[u.day_length(t, x) for x in φ]))    # :141-143; an Ny × 3 column-major matrix = the C array [3][Ny]
        check(ccall((:obm_pisces_tendencies_rows, libobm), Cint,
                    (Ref{ObmGrid}, Ref{ObmPiscesParams}, F64, Ptr{F64}, Ref{ObmPiscesFields}, Ptr{F64}, Cint, Ptr{Cvoid}),
                    g, Ref(ObmPiscesParams(u, model.clock)), pointer(rows), table(model.tracers[n] for n in names), Ref(f), G, 1, stream()),
              "obm_pisces_tendencies_rows")
        return nothing
    end
    check(ccall((:obm_pisces_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmPiscesParams}, Ptr{F64}, Ref{ObmPiscesFields}, Ptr{F64}, Cint, Ptr{Cvoid}),
                g, Ref(ObmPiscesParams(u, model.clock)), table(model.tracers[n] for n in names), Ref(f), G, 1, stream()),
          "obm_pisces_tendencies")
end

# Parameter-sweep ensembles (examples/data_assimilation.jl builds one NPZD box model per parameter vector): one model
# whose columns are the members.  `which` = Int32[obm_npd_param_index("phytoplankton_maximum_growth_rate"), …],
# `values` = CuArray{Float64}(n_members, n_varied) (member fastest); everything else comes from the model's parameters.
function update_tendencies!(b::B200NPD, model, which::Vector{Int32}, values::CuMatrix{Float64})
    u = b.reference.underlying_biogeochemistry
    names = required_biogeochemical_tracers(u)
    G = F64[n === :T ? CU_NULL : dptr(model.timestepper.Gⁿ[n]) for n in names]
    check(ccall((:obm_npd_tendencies_ensemble, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Cint, Ptr{Int32}, F64, Ptr{F64}, F64, Ptr{F64}, Cint, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), Ref(ObmNpdParams(u)), length(which), which, pointer(values), table(model.tracers[n] for n in names),
                dptr(biogeochemical_auxiliary_fields(b.reference).PAR), G, 1, stream()), "obm_npd_tendencies_ensemble")
    return nothing
end

# ---- the whole run of a box-model ensemble in ONE launch (`run!(simulation)` over src/BoxModel/boxmodel.jl:92-110 and
#      timesteppers.jl:30-93; the reference's benchmark/box_model.jl).  The boxes lie along x (`BoxModelGrid` with Nx members);
#      `PAR_series` / `T_series` hold what `update_state!` prescribes AFTER every stage — one row per global stage, one column
#      (shared) or Nx columns — as CuMatrix{Float64}(columns, rows), i.e. row-major [rows][columns] for the C side;
#      `snapshots[n]` is a CuMatrix (Nx, nsteps ÷ output_every) per tracer, or `nothing`.  RK3 coefficients are Oceananigans'.
function run_boxes!(b::B200NPD, model, Δt, nsteps; PAR_series::CuMatrix{Float64}, T_series = nothing, output_every = 0,
                    snapshots = nothing, which = Int32[], values = nothing,
                    γ = Float64[8/15, 5/12, 3/4], ζ = Float64[NaN, -17/60, -5/12])
    u = b.reference.underlying_biogeochemistry
    names = required_biogeochemical_tracers(u)
    U  = table(model.fields[n] for n in names)
    G⁻ = F64[n === :T ? CU_NULL : dptr(model.timestepper.G⁻[n]) for n in names]
    S  = isnothing(snapshots) ? Ptr{F64}(C_NULL) : F64[haskey(snapshots, n) ? pointer(snapshots[n]) : CU_NULL for n in names]
    check(ccall((:obm_npd_box_run, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Cint, Ptr{Int32}, F64, Ptr{F64}, Ptr{F64}, F64, F64, Cint, F64, Cint, Cint, Cint,
                 Ptr{Float64}, Ptr{Float64}, Cdouble, Cint, Ptr{F64}, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), Ref(ObmNpdParams(u)), length(which), which, isnothing(values) ? CU_NULL : pointer(values),
                U, G⁻, dptr(biogeochemical_auxiliary_fields(b.reference).PAR), pointer(PAR_series), size(PAR_series, 1) != 1,
                isnothing(T_series) ? CU_NULL : pointer(T_series), !isnothing(T_series) && size(T_series, 1) != 1,
                nsteps, length(γ), γ, ζ, Δt, output_every, S, stream()), "obm_npd_box_run")
    return nothing
end

# ---- air–sea gas exchange (src/Models/GasExchange/gas_exchange.jl:26-38): the reference evaluates `g(i, j, grid, clock, fields)`
#      per surface cell inside Oceananigans' boundary-condition kernel; here the whole x–y plane is computed by one launch into
#      a flux field, and the boundary condition reads that field: `FluxBoundaryCondition(flux_field)`.  Call it from a
#      callback / after `update_biogeochemical_state!` each stage.  `p` is an ObmGasExchangeParams (k660 / Schmidt / Wanninkhof-92
#      coefficients copied from g.transfer_velocity and g.air_concentration, gas_transfer_velocity.jl, schmidt_number.jl,
#      gas_solubility.jl:31-65); DIC / Alk only for CO₂; wind and air concentration as 2-D fields or the struct's constants.
gas_exchange_flux!(g, p, T, S, tracer, DIC, Alk, silicate, phosphate, wind_xy, air_xy, flux_xy, G_top, s = stream()) =
    check(ccall((:obm_gas_exchange_flux, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmGasExchangeParams}, F64, F64, F64, F64, F64, F64, F64, F64, F64, F64, F64, Ptr{Cvoid}),
                g, p, T, S, tracer, DIC, Alk, silicate, phosphate, wind_xy, air_xy, flux_xy, G_top, s), "obm_gas_exchange_flux")

# ---- conservation diagnostics: Σ sf·c·V per conserved group on this GPU (one fused pass), then the path's only collective —
#      an all-reduce of ≤ 8 doubles over the ranks (NCCL.Allreduce! / MPI.Allreduce! on the returned device vector) ------------
function tracer_inventory(b::B200Biogeochemistry, model, groups; cell_volume = nothing, uniform_volume = 0.0)
    names = unique(vcat((collect(t) for (t, _) in groups)...))
    cgroups = scale_groups([(; tracers = t, scalefactors = sf) for (t, sf) in groups], names)
    check(ccall((:obm_inventory, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Cint, Ptr{ObmScaleGroup}, F64, Cdouble, F64, CuPtr{Cvoid}, Ptr{Cvoid}),
                Ref(ObmGrid(model.grid)), length(names), table(model.tracers[n] for n in names), length(cgroups), cgroups,
                dptr(cell_volume), Float64(uniform_volume), pointer(b.scratch.inventory),
                reinterpret(CuPtr{Cvoid}, pointer(b.scratch.inventory_workspace)), stream()), "obm_inventory")
    return view(b.scratch.inventory, 1:length(cgroups))
end

# ---- raw bindings of the remaining entry points (include/obm_b200.h), one thin method each ---------------------------
# Device arrays are passed as CuPtr{Float64} (`dptr(field)`), tables of them as host Vector{CuPtr{Float64}}, parameter blocks
# by Ref; `s` is `stream()`.
par_multiband!(g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, s) =
    check(ccall((:obm_par_multiband, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmMultibandParams}, F64, F64, Cdouble, F64, Cdouble, Ptr{F64}, F64, Ptr{Cvoid}),
                g, p, chl_a, chl_b, chl_scale, surface_xy, surface_const, bands, total, s), "obm_par_multiband")

find_bottom_cells!(g, bottom_height_xy, bottom_indices_xy::CuPtr{Int64}, s) =
    check(ccall((:obm_find_bottom_cells, libobm), Cint, (Ref{ObmGrid}, F64, CuPtr{Int64}, Ptr{Cvoid}),
                g, bottom_height_xy, bottom_indices_xy, s), "obm_find_bottom_cells")

# Particles (src/Particles): `update_tendencies!(bgc, particles::BiogeochemicalParticles{<:SugarKelp}, model)` → all 8 coupled
# tracers in one launch; `time_step_particle_fields!(::ForwardEuler, …)` → obm_kelp_step.  ObmParticles carries
# pointer(particles.x), …, pointer(particles.fields.A), …, the first cell centre and spacing in x and y and the topology
# codes; ObmKelpTracers the parents of u, v, w, T, NO₃, NH₄ and PAR.
kelp_update_tendencies!(g, p, particles, tracers, G, t, s) =
    check(ccall((:obm_kelp_update_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSugarKelpParams}, Ref{ObmParticles}, Ref{ObmKelpTracers}, Ptr{F64}, Cdouble, Ptr{Cvoid}),
                g, p, particles, tracers, G, t, s), "obm_kelp_update_tendencies")

kelp_step!(g, p, particles, tracers, t, Δt, tendencies_out, s) =
    check(ccall((:obm_kelp_step, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmSugarKelpParams}, Ref{ObmParticles}, Ref{ObmKelpTracers}, Cdouble, Cdouble, Ptr{F64}, Ptr{Cvoid}),
                g, p, particles, tracers, t, Δt, tendencies_out, s), "obm_kelp_step")

# Column / box models without resolved flow: Gⁿ += −∂z(w c) of every sinking tracer (w = biogeochemical_drift_velocity(bgc, Val(c)).w)
sinking_tendencies!(g, tracers, w_faces, G, advection, accumulate, s) =
    check(ccall((:obm_sinking_tendencies, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Ptr{F64}, Ptr{F64}, Cint, Cint, Ptr{Cvoid}),
                g, length(tracers), tracers, w_faces, G, advection, accumulate, s), "obm_sinking_tendencies")

# box / column models: U += Δt (γ Gⁿ + ζ G⁻) and G⁻ ← Gⁿ for every field in one launch (src/BoxModel/timesteppers.jl:66-93)
rk3_substep!(g, U, Gⁿ, G⁻, Δt, γ, ζ, has_ζ, cache_previous, s) =
    check(ccall((:obm_rk3_substep, libobm), Cint,
                (Ref{ObmGrid}, Cint, Ptr{F64}, Ptr{F64}, Ptr{F64}, Cdouble, Cdouble, Cdouble, Cint, Cint, Ptr{Cvoid}),
                g, length(U), U, Gⁿ, G⁻, Δt, γ, ζ, has_ζ, cache_previous, s), "obm_rk3_substep")

# the same with the tendencies evaluated in the launch itself (NPZD / LOBSTER family): compute_tendencies! + rk3_substep! +
# cache_previous_tendencies! of src/BoxModel/timesteppers.jl:30-93 in one pass; Gⁿ may be C_NULL (no forcing, nothing stored)
npd_tendencies_substep!(g, p, nvary, which, values, U, PAR, Gⁿ, accumulate, store_Gⁿ, G⁻, Δt, γ, ζ, has_ζ, s) =
    check(ccall((:obm_npd_tendencies_substep, libobm), Cint,
                (Ref{ObmGrid}, Ref{ObmNpdParams}, Cint, Ptr{Int32}, F64, Ptr{F64}, F64, Ptr{F64}, Cint, Cint, Ptr{F64}, Cdouble, Cdouble,
                 Cdouble, Cint, Ptr{Cvoid}),
                g, p, nvary, which, values, U, PAR, Gⁿ, accumulate, store_Gⁿ, G⁻, Δt, γ, ζ, has_ζ, s), "obm_npd_tendencies_substep")

# CarbonChemistry()(; DIC, Alk, T, S, …) over flat device arrays (src/Models/CarbonChemistry/carbon_chemistry.jl:84-181)
carbon_chemistry!(out, p, T, S, DIC, Alk, P_bar, silicate, phosphate, pH, output_kind, n, s) =
    check(ccall((:obm_carbon_chemistry, libobm), Cint,
                (Int64, Ref{ObmCarbchemParams}, F64, F64, F64, F64, F64, F64, F64, F64, Cint, F64, Ptr{Cvoid}),
                n, p, T, S, DIC, Alk, P_bar, silicate, phosphate, pH, output_kind, out, s), "obm_carbon_chemistry")

end # module
