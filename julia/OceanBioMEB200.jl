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

# ---- struct mirrors (isbits; field order = include/obm_b200.h) -------------------------------------------
struct ObmGrid
    Nx::Int32; Ny::Int32; Nz::Int32
    Hx::Int32; Hy::Int32; Hz::Int32
    i0::Int32; i1::Int32; j0::Int32; j1::Int32
    zc::CuPtr{Float64}; zf::CuPtr{Float64}
end

function ObmGrid(grid)
    Nx, Ny, Nz = size(grid)
    Hx, Hy, Hz = Oceananigans.Grids.halo_size(grid)
    Hx, Hy = (Nx == 1 ? 0 : Hx), (Ny == 1 ? 0 : Hy)                      # Flat dimensions
    zc = parent(grid.z.cᵃᵃᶜ); zf = parent(grid.z.cᵃᵃᶠ)                     # device OffsetVectors incl. halos
    return ObmGrid(Nx, Ny, Nz, Hx, Hy, Hz, 0, 0, 0, 0, pointer(zc), pointer(zf))
end

struct ObmNpdParams                                                       # obm_npd_params
    nutrients::Int32; detritus::Int32; carbonate_replicates::Int32; oxygen::Int32
    light_limitation::Int32; phytoplankton_mortality_formulation::Int32
    grazing_concentration_formulation::Int32; has_temperature_coefficient::Int32
    doubles::NTuple{35, Float64}                                          # plankton.jl:19-58 … oxygen.jl:14-17, header order
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
    # 2. light (src/Light/2band.jl:148-155): surface_PAR evaluated here into a scalar / 2-D field, then
    PAR = ref.light_attenuation
    surface = Float64(OceanBioME.Light.default_surface_PAR(model.clock.time))
    tb = Ref(twoband_params(PAR))                                          # obm_twoband_params, 8 doubles
    check(ccall((:obm_par_twoband, libobm), Cint,
                (Ref{ObmGrid}, Ptr{Cvoid}, CuPtr{Float64}, CuPtr{Float64}, Cdouble, CuPtr{Float64}, Ptr{Cvoid}),
                g, tb, pointer(parent(model.tracers.P)), CU_NULL, surface, pointer(parent(PAR.field)), s),
          "obm_par_twoband")
    # 3. underlying (PISCES: obm_euphotic_depth, obm_mixed_layer_mean, obm_calcite_saturation)   4. sediment
    return nothing
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

end # module
