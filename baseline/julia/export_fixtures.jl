# export_fixtures.jl — for anyone who CAN run Julia (the build environment of this repo cannot).
# Runs the reference's own CPU path (EnhancedBayesianNetworks.jl + UncertaintyQuantification.jl 0.11)
# on the vehicle-suspension limit state and writes the sample matrix, g and the fail count in the
# exchange format of SURVEY.md §8c, so that `ebn_eval_matrix(strict=1)` can be checked bit for bit
# against REAL reference output:
#
#     julia --project=/path/to/EnhancedBayesianNetworks.jl baseline/julia/export_fixtures.jl out_dir
#     python tools/check_julia_fixture.py out_dir/vehicle_fixture.bin out_dir/vehicle_fixture.json
#
using EnhancedBayesianNetworks, Random, JSON

M, m, g = 3.2633, 0.8158, 981
function composite_model(A, b₀, V, M, m, g, C, Cₖ, K)   # demo/vehicle_suspension.jl:17-23
    g1 = 1 .- (π .* m .* V .* A) ./ (b₀ .* K .* g .^ 2) .* [(Cₖ ./ (m .+ M) .- (C ./ M)) .^ 2 .+ C .^ 2 ./ (m .* M) .+ Cₖ .* K .^ 2 ./ (m .* M .^ 2)]
    g2 = 4000 .* C .* (M .* g) .^ (-1.5) .- 8.6394
    g3 = 2 .* .√(M .* g .* (K .^ 2 .* Cₖ ./ (C .* (m .+ M)) .+ C)) .- 1
    g4 = Cₖ .- [g .* (M .+ m)] .^ 0.877
    return minimum([g1[1], g2, g3, g4[1]])
end

outdir = length(ARGS) > 0 ? ARGS[1] : "."
Random.seed!(20261019)
n = 100_000
inputs = [Parameter(0.8, :A), Parameter(0.27, :b₀), RandomVariable(Uniform(7, 12), :V),
    RandomVariable(Normal(431.7221, 10), :C), RandomVariable(Normal(1475.5503, 10), :Cₖ),
    RandomVariable(Normal(55.0406, 10), :K)]
model = Model(df -> composite_model.(df.A, df.b₀, df.V, M, m, g, df.C, df.Cₖ, df.K), :y)
df = UncertaintyQuantification.sample(inputs, n)
UncertaintyQuantification.evaluate!([model], df)
count = sum(df.y .< 0)
cols = [:A, :b₀, :V, :C, :Cₖ, :K, :y]
open(joinpath(outdir, "vehicle_fixture.bin"), "w") do io   # column-major little-endian Float64
    for c in cols
        write(io, Float64.(df[!, c]))
    end
end
open(joinpath(outdir, "vehicle_fixture.json"), "w") do io
    JSON.print(io, Dict("n" => n, "columns" => String.(cols), "count" => count, "pf" => count / n,
        "consts" => [M, m, Float64(g), (M * g)^(-1.5), (g * (M + m))^0.877]))
end
println("wrote $n rows, count = $count")
