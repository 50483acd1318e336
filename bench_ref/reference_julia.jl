# reference_julia.jl — times the UNMODIFIED OceanBioME.jl on the same workloads as bench.py, for anyone who has Julia (this
# image does not: `which julia` fails, so bench.py's `--impl reference` arm times the C restatement under oracle/ instead
# and says `"kind": "port"`).  NOT run here; kept short so a maintainer can check it by eye.
#
#   julia --project -t auto bench_ref/reference_julia.jl [lobster_c3|pisces_c4|npzd_c1] [CPU|GPU] [scale]
#
# Prints one JSON line in bench.py's format: metric "BGC tendency Gcell-updates/s", one "step" = the biogeochemical part of
# one time-stepper stage = update_biogeochemical_state! + the tracer tendency kernels (compute_tendencies! of a model with
# no advection, closure, buoyancy or forcing).  Synthetic state: the same splitmix64 stream, field ids (crc32 of the name
# + 1) and ranges as oceanbiome.jl_b200/synthetic.py / pisces.synthetic_range, so the inputs are those of bench.py.
using OceanBioME, Oceananigans, CUDA, Printf, CRC32   # CRC32.jl: `crc32` = zlib's polynomial (CRC32c is a different one)
using Oceananigans.TimeSteppers: update_state!, compute_tendencies!

workload = length(ARGS) ≥ 1 ? ARGS[1] : "lobster_c3"
arch     = length(ARGS) ≥ 2 && ARGS[2] == "GPU" ? GPU() : CPU()
scale    = length(ARGS) ≥ 3 ? parse(Float64, ARGS[3]) : 1.0

sizes   = Dict("npzd_c1" => (160, 1, 32), "lobster_c3" => (512, 512, 64), "pisces_c4" => (1024, 1024, 128))
extents = Dict("npzd_c1" => (10e3, 1.0, 500.0), "lobster_c3" => (1000.0, 1000.0, 140.0), "pisces_c4" => (1024e3, 1024e3, 400.0))
Nx, Ny, Nz = sizes[workload]
Lx, Ly, Lz = extents[workload]
Ny = max(1, round(Int, Ny * scale))

# lobster_c3: the stretched vertical grid of paper/figures/eady.jl:11-27
h(k) = (k - 1) / Nz; ζ₀(k) = 1 + (h(k) - 1) / 1.8; Σ(k) = (1 - exp(-3h(k))) / (1 - exp(-3))
z = workload == "lobster_c3" ? (k -> Lz * (ζ₀(k) * Σ(k) - 1)) : (-Lz, 0)
grid = RectilinearGrid(arch; size = (Nx, Ny, Nz), x = (0, Lx), y = (0, Ly * Ny / sizes[workload][2]), z)

biogeochemistry =
    workload == "npzd_c1"    ? NPZD(; grid, scale_negatives = true, surface_photosynthetically_active_radiation = (x, y, t) -> 100.0) :
    workload == "lobster_c3" ? LOBSTER(; grid, carbonate_system = CarbonateSystem(), oxygen = Oxygen(), scale_negatives = true,
                                       sediment = SimpleMultiGSediment(grid),                          # BASELINE configs[2]: with sediment
                                       surface_photosynthetically_active_radiation = (x, y, t) -> 100.0) :
                               PISCES(; grid, scale_negatives = true,                                   # PISCES.jl:288: keyword `grid`
                                      surface_photosynthetically_active_radiation = (x, y, t) -> 100.0)

extra = workload == "lobster_c3" ? (:T, :S) : ()
model = NonhydrostaticModel(; grid, biogeochemistry, tracers = extra, advection = nothing, closure = nothing, buoyancy = nothing)

# deterministic synthetic state — bit-identical to synthetic.fill_numpy (linear index n = i + Nx (j + Ny k), 0-based)
splitmix(x::UInt64) = (x += 0x9E3779B97F4A7C15; x = (x ⊻ (x >> 30)) * 0xBF58476D1CE4E5B9;
                       x = (x ⊻ (x >> 27)) * 0x94D049BB133111EB; x ⊻ (x >> 31))
const SEED = UInt64(20260117)
field_id(name) = UInt64(crc32(codeunits(String(name)))) + 1       # zlib.crc32 (IEEE): use the CRC32 package's `crc32`, not crc32c
u01(f, n) = Float64(splitmix(SEED ⊻ (f * 0x9E3779B97F4A7C15) ⊻ UInt64(n)) >> 11) * 2.0^-53

pisces0 = Dict(:P => 0.5, :PChl => 0.02, :PFe => 0.005, :D => 0.1, :DChl => 0.004, :DFe => 0.001, :DSi => 0.01, :Z => 0.1, :M => 0.7,
               :DOC => 2.1, :POC => 7.8, :SFe => 0.206, :GOC => 38.0, :BFe => 1.1, :PSi => 0.1, :NO₃ => 2.3, :NH₄ => 0.9,
               :PO₄ => 0.6, :Fe => 0.13, :Si => 8.5, :DIC => 2205.0, :Alk => 2566.0, :O₂ => 317.0)       # test/test_PISCES.jl:8-16
lobster = Dict(:NO₃ => (0.0, 12.0, false), :NH₄ => (1e-3, 1.0, true), :P => (0.005, 0.5, true), :Z => (0.005, 0.5, true),
               :sPOM => (0.0, 1.0, false), :bPOM => (0.0, 1.0, false), :DOM => (0.0, 1.0, false), :DIC => (2000.0, 2300.0, false),
               :Alk => (2300.0, 2500.0, false), :O₂ => (150.0, 350.0, false), :T => (2.0, 28.0, false), :S => (33.0, 37.0, false))
npzd = Dict(:N => (0.5, 4.5, false), :P => (0.01, 0.03, false), :Z => (0.01, 0.03, false), :D => (0.0, 0.1, false), :T => (8.9, 9.1, false))
function range_of(name)                                           # (lo, hi, log-uniform?) — pisces.synthetic_range, synthetic.RANGES_*
    workload == "npzd_c1" && return npzd[name]
    workload == "lobster_c3" && return lobster[name]
    name == :T && return (2.0, 28.0, false)
    name == :S && return (33.0, 37.0, false)
    name == :CaCO₃ && return (0.01, 1.0, true)
    v = pisces0[name]
    return name in (:DIC, :Alk) ? (0.98v, 1.02v, false) : (v * exp(-0.5), v * exp(0.5), true)
end
for name in keys(model.tracers)
    lo, hi, islog = range_of(name)
    f = field_id(name)
    value(u) = islog ? exp(log(lo) + (log(hi) - log(lo)) * u) : lo + (hi - lo) * u
    vals = [value(u01(f, (i - 1) + Nx * ((j - 1) + Ny * (k - 1)))) for i in 1:Nx, j in 1:Ny, k in 1:Nz]
    set!(model.tracers[name], vals)
end

stage!() = (update_state!(model); compute_tendencies!(model, []))

for _ in 1:3; stage!(); end
steps = workload == "pisces_c4" ? 3 : 10
t0 = time_ns(); for _ in 1:steps; stage!(); end; arch isa GPU && CUDA.synchronize(); t1 = time_ns()
cells = Nx * Ny * Nz
@printf("{\"impl\": \"reference\", \"metric\": \"BGC tendency Gcell-updates/s\", \"value\": %.6g, \"unit\": \"Gcell-updates/s\", \"config\": {\"workload\": \"%s\", \"cells\": %d}, \"threads\": %d, \"arch\": \"%s\", \"ms_per_step\": %.4g}\n",
        cells * steps / ((t1 - t0) * 1e-9) / 1e9, workload, cells, Threads.nthreads(), string(typeof(arch)), (t1 - t0) * 1e-6 / steps)
