# bench_reference.jl — times the REFERENCE's own CPU path (BASELINE.md §3 row B3). Not runnable in
# the build environment (no Julia); shipped so that the cell can be filled elsewhere:
#
#     julia --project=/path/to/EnhancedBayesianNetworks.jl baseline/julia/bench_reference.jl [n]
#
# Current-API restatement of demo/vehicle_suspension.jl (the demo file itself uses a stale API,
# SURVEY.md F4): 20 scenarios, n samples each, `evaluate!(net, true, false)`.
using EnhancedBayesianNetworks

n = length(ARGS) > 0 ? parse(Int, ARGS[1]) : 2 * 10^6
M, m, g = 3.2633, 0.8158, 981
function composite_model(A, b₀, V, M, m, g, C, Cₖ, K)
    g1 = 1 .- (π .* m .* V .* A) ./ (b₀ .* K .* g .^ 2) .* [(Cₖ ./ (m .+ M) .- (C ./ M)) .^ 2 .+ C .^ 2 ./ (m .* M) .+ Cₖ .* K .^ 2 ./ (m .* M .^ 2)]
    g2 = 4000 .* C .* (M .* g) .^ (-1.5) .- 8.6394
    g3 = 2 .* .√(M .* g .* (K .^ 2 .* Cₖ ./ (C .* (m .+ M)) .+ C)) .- 1
    g4 = Cₖ .- [g .* (M .+ m)] .^ 0.877
    return minimum([g1[1], g2, g3, g4[1]])
end

function root(name, states, probs, pars)
    cpt = DiscreteConditionalProbabilityTable{PreciseDiscreteProbability}(name)
    for (s, p) in zip(states, probs)
        cpt[name=>s] = p
    end
    DiscreteNode(name, cpt, pars)
end
function croot(name, dist, disc=ExactDiscretization())
    cpt = ContinuousConditionalProbabilityTable{PreciseContinuousInput}()
    cpt[] = dist
    ContinuousNode(name, cpt, disc)
end

A = root(:A, [:road, :offroad], [0.7, 0.3], Dict(:road => [Parameter(0.15915, :A)], :offroad => [Parameter(0.8, :A)]))
b₀ = root(:b₀, [:normal_load, :over_load], [0.7, 0.3], Dict(:normal_load => [Parameter(0.27, :b₀)], :over_load => [Parameter(0.5, :b₀)]))
V = croot(:V, Uniform(7, 12), ExactDiscretization(collect(range(8.5, 11.5, 4))))
C = croot(:C, Normal(431.7221, 10))
Cₖ = croot(:Cₖ, Normal(1475.5503, 10))
K = croot(:K, Normal(55.0406, 10))
model = Model(df -> composite_model.(df.A, df.b₀, df.V, M, m, g, df.C, df.Cₖ, df.K), :y)
E = DiscreteFunctionalNode(:E, [model], df -> df.y, MonteCarlo(n))
net = EnhancedBayesianNetwork([A, b₀, V, C, Cₖ, K, E])
for p in (A, b₀, V, C, Cₖ, K)
    add_child!(net, p, E)
end
order!(net)
evaluate!(deepcopy(net), true, false)          # compile
t = @elapsed evaluate!(net, true, false)
println("reference CPU path: 20 scenarios x $n samples in $(round(t, digits=3)) s = $(round(20n / t, sigdigits=4)) samples/s on 1 thread")
