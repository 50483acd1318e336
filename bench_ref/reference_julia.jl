# reference_julia.jl — times the UNMODIFIED OceanBioME.jl on the same workloads as bench.py, for anyone
# who has Julia (this image does not: `which julia` fails, so bench.py's `--impl reference` arm times
# the C restatement under oracle/ instead and says `"kind": "port"`).  NOT run here; kept short so a
# maintainer can check it by eye.
#
#   julia --project -t auto bench_ref/reference_julia.jl [lobster_c3|pisces_c4|npzd_c1] [CPU|GPU] [scale]
#
# Prints one JSON line in bench.py's format: metric "BGC tendency Gcell-updates/s", one "step" = the
# biogeochemical part of one time-stepper stage = update_biogeochemical_state! + the tracer tendency
# kernels (compute_tendencies! of a model with no advection, closure, buoyancy or forcing).
using OceanBioME, Oceananigans, CUDA, Printf
using Oceananigans.TimeSteppers: update_state!, compute_tendencies!

workload = length(ARGS) ≥ 1 ? ARGS[1] : "lobster_c3"
arch     = length(ARGS) ≥ 2 && ARGS[2] == "GPU" ? GPU() : CPU()
scale    = length(ARGS) ≥ 3 ? parse(Float64, ARGS[3]) : 1.0

sizes = Dict("npzd_c1" => (160, 1, 32), "lobster_c3" => (512, 512, 64), "pisces_c4" => (1024, 1024, 128))
Nx, Ny, Nz = sizes[workload]
Ny = max(1, round(Int, Ny * scale))
grid = RectilinearGrid(arch; size = (Nx, Ny, Nz), extent = (Nx * 10.0, Ny * 10.0, workload == "pisces_c4" ? 400.0 : 140.0))

biogeochemistry =
    workload == "npzd_c1"    ? NPZD(grid) :
    workload == "lobster_c3" ? LOBSTER(grid; carbonate_system = CarbonateSystem(), oxygen = Oxygen()) :
                               PISCES(grid)

model = NonhydrostaticModel(grid; biogeochemistry, tracers = (:T, :S), advection = nothing, closure = nothing,
                            buoyancy = nothing)

# deterministic synthetic state: the same splitmix64 stream as oceanbiome.jl_b200/synthetic.py
splitmix(x::UInt64) = (x += 0x9E3779B97F4A7C15; x = (x ⊻ (x >> 30)) * 0xBF58476D1CE4E5B9;
                       x = (x ⊻ (x >> 27)) * 0x94D049BB133111EB; x ⊻ (x >> 31))
const SEED = UInt64(20260117)
u01(f, n) = Float64(splitmix(SEED ⊻ (UInt64(f) * 0x9E3779B97F4A7C15) ⊻ UInt64(n)) >> 11) * 2.0^-53
for (f, name) in enumerate(keys(model.tracers))
    lo, hi = name == :T ? (2.0, 28.0) : name == :S ? (33.0, 37.0) : name in (:DIC, :Alk) ? (2000.0, 2400.0) : (0.01, 1.0)
    vals = [lo + (hi - lo) * u01(f, (i - 1) + Nx * ((j - 1) + Ny * (k - 1))) for i in 1:Nx, j in 1:Ny, k in 1:Nz]
    set!(model.tracers[name], vals)
end

stage!() = (update_state!(model); compute_tendencies!(model, []))

for _ in 1:3; stage!(); end
steps = workload == "pisces_c4" ? 3 : 10
t0 = time_ns(); for _ in 1:steps; stage!(); end; arch isa GPU && CUDA.synchronize(); t1 = time_ns()
cells = Nx * Ny * Nz
@printf("{\"impl\": \"reference\", \"metric\": \"BGC tendency Gcell-updates/s\", \"value\": %.6g, \"unit\": \"Gcell-updates/s\", \"config\": {\"workload\": \"%s\", \"cells\": %d}, \"threads\": %d, \"arch\": \"%s\", \"ms_per_step\": %.4g}\n",
        cells * steps / ((t1 - t0) * 1e-9) / 1e9, workload, cells, Threads.nthreads(), string(typeof(arch)), (t1 - t0) * 1e-6 / steps)
