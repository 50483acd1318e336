# dump_goldens.jl — for owners of Julia: pin this repository's oracle on the REAL OceanBioME.jl.
#
# The build image has no Julia, so the absolute NPD / PISCES tendency values of the CPU oracle (oracle/src/*.c) rest on two
# independent readings of the reference source (DESIGN.md §4), not on reference output.  This script closes that gap on any
# machine with Julia and OceanBioME v0.17.6: it evaluates the reference's own per-tracer callables
#     bgc(i, j, k, grid, Val(name), clock, fields, auxiliary_fields)
# (PISCES.jl:122-123, NutrientsPlanktonDetritus.jl:88 and the component methods) at exactly the states stored in
# tests/golden/pisces_tendencies.json and tests/golden/npd_tendencies.json and writes the SAME schema to
# tests/golden/reference_pisces_tendencies.json / reference_npd_tendencies.json.  tests/test_reference_goldens.py then diffs
# those files against the oracle (1e-13 of the tendency's own Σ|terms|) whenever they are present.
#
#   julia --project bench_ref/dump_goldens.jl            (from the repository root; needs OceanBioME, Oceananigans, JSON)
#
# NOT RUN HERE.  Kept to the reference's public constructors and the callable forms cited above.
using OceanBioME, Oceananigans, JSON
using Oceananigans.Units: day
using Oceananigans.Fields: ConstantField
using OceanBioME.Models.PISCESModel: PISCES as PISCESUnderlying

root = normpath(joinpath(@__DIR__, ".."))
cell(v) = fill(Float64(v), 1, 1, 1)                    # a field of one cell, indexed [i, j, k]
faces(v) = fill(Float64(v), 1, 1, 2)                   # a z-face field of one cell: ℑzᵃᵃᶜ of two equal faces is v

# ---- PISCES ---------------------------------------------------------------------------------------------------------------
function pisces_rows(rows)
    out = []
    for r in rows
        f = r["state"]
        z = f["z"]
        grid = RectilinearGrid(CPU(); size = (1, 1, 1), x = (0, 1), y = (0, 1), z = (z - 0.5, z + 0.5), topology = (Periodic, Periodic, Bounded))
        bgc = PISCES(; grid, latitude = PrescribedLatitude(r["latitude"]), silicate_climatology = ConstantField(f["Si_clim"]),
                     sinking_speeds = (POC = 0.0, GOC = 0.0))        # w comes from the row (auxiliary fields below)
        u = bgc.underlying_biogeochemistry
        tracers = Oceananigans.Biogeochemistry.required_biogeochemical_tracers(u)
        fields = NamedTuple{tracers}(Tuple(cell(get(f, String(n), 0.0)) for n in tracers))
        aux = (zₘₓₗ = cell(f["zₘₓₗ"]), zₑᵤ = cell(f["zₑᵤ"]), Si′ = cell(f["Si_clim"]), Ω = cell(f["Ω"]), κ = cell(f["κ"]),
               mixed_layer_PAR = cell(f["mixed_layer_PAR"]), wPOC = faces(f["wPOC"]), wGOC = faces(f["wGOC"]),
               PAR = cell(f["PAR"]), PAR₁ = cell(f["PAR₁"]), PAR₂ = cell(f["PAR₂"]), PAR₃ = cell(f["PAR₃"]))
        clock = Clock(time = f["t"])
        tend = Dict(String(n) => Float64(u(1, 1, 1, grid, Val(n), clock, fields, aux)) for n in tracers if !(n in (:T, :S)))
        push!(out, Dict("latitude" => r["latitude"], "state" => f, "tendencies" => tend,
                        "day_length_growth" => Float64(u.day_length(r["latitude"], f["t"])),          # growth_rate.jl:30 order
                        "day_length_chlorophyll" => Float64(u.day_length(f["t"], r["latitude"]))))     # growth_rate.jl:143 order
    end
    return out
end

# ---- NutrientsPlanktonDetritus family --------------------------------------------------------------------------------------
npd_grid = RectilinearGrid(CPU(); size = (1, 1, 1), extent = (1, 1, 1))
npd_models = Dict(
    "lobster" => () -> LOBSTER(; grid = npd_grid),
    "lobster_carbonate_oxygen" => () -> LOBSTER(; grid = npd_grid, carbonate_system = CarbonateSystem(), oxygen = Oxygen()),
    "lobster_iron_variable_redfield_carbonate_oxygen" =>
        () -> LOBSTER(; grid = npd_grid, nutrients = NitrateAmmoniaIron(), detritus = VariableRedfieldDetritus(grid = npd_grid),
                      carbonate_system = CarbonateSystem(), oxygen = Oxygen()),
    "npzd" => () -> NPZD(; grid = npd_grid),
    "npzd_carbonate_oxygen" => () -> NPZD(; grid = npd_grid, carbonate_system = CarbonateSystem(), oxygen = Oxygen()))

function npd_cases(cases)
    out = Dict()
    for (name, c) in cases
        u = npd_models[name]().underlying_biogeochemistry
        tracers = Tuple(Symbol.(c["tracers"]))
        @assert tracers == Oceananigans.Biogeochemistry.required_biogeochemical_tracers(u) "tracer order of $name"
        rows = []
        for r in c["rows"]
            f = r["state"]
            fields = NamedTuple{tracers}(Tuple(cell(f[String(n)]) for n in tracers))
            aux = (PAR = cell(f["PAR"]),)
            tend = Dict(String(n) => Float64(u(1, 1, 1, npd_grid, Val(n), Clock(time = 0.0), fields, aux)) for n in tracers)
            push!(rows, Dict("state" => f, "tendencies" => tend))
        end
        out[name] = Dict("tracers" => c["tracers"], "rows" => rows)
    end
    return out
end

stamp = "bench_ref/dump_goldens.jl (OceanBioME.jl $(pkgversion(OceanBioME)), Julia $(VERSION))"
pis = JSON.parsefile(joinpath(root, "tests", "golden", "pisces_tendencies.json"))
open(joinpath(root, "tests", "golden", "reference_pisces_tendencies.json"), "w") do io
    JSON.print(io, Dict("generator" => stamp, "rows" => pisces_rows(pis["rows"])), 1)
end
npd = JSON.parsefile(joinpath(root, "tests", "golden", "npd_tendencies.json"))
open(joinpath(root, "tests", "golden", "reference_npd_tendencies.json"), "w") do io
    JSON.print(io, Dict("generator" => stamp, "cases" => npd_cases(npd["cases"])), 1)
end
println("wrote tests/golden/reference_{pisces,npd}_tendencies.json — now run: python -m pytest tests/test_reference_goldens.py")
