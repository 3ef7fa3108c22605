# EBNCuda.jl — `ccall` wrapper over libebncuda.so (C ABI: include/ebncuda.h).
#
# Drop-in for the Monte Carlo hot path of EnhancedBayesianNetworks.jl. A functional node whose
# `simulation` is a `CudaMonteCarlo` (a new `AbstractMonteCarlo` subtype) is evaluated by ONE call
# into the library for all its scenarios instead of one UncertaintyQuantification.jl call per
# scenario (reference src/networks/ebn/evaluation/evaluate_node.jl:79-95 and :22-33).
#
# NOT EXECUTED in the build environment of this repository (no Julia there, SURVEY.md F3): this
# file is kept a thin, obviously-correct marshalling layer; everything it calls is exercised
# through the same C ABI by the Python tests under tests/.
#
# Integration (see INTEGRATION.md): `include("cuda/EBNCuda.jl")` at the end of
# src/EnhancedBayesianNetworks.jl, after networks/ebn/ebn.jl.

module EBNCuda

using ..EnhancedBayesianNetworks
using ..EnhancedBayesianNetworks: AbstractNode, ContinuousNode, DiscreteNode, DiscreteFunctionalNode,
    ContinuousFunctionalNode, EnhancedBayesianNetwork, Evidence, parents, discrete_ancestors, states,
    _uq_inputs, DiscreteConditionalProbabilityTable, ContinuousConditionalProbabilityTable,
    PreciseDiscreteProbability, PreciseContinuousInput, ExactDiscretization
using UncertaintyQuantification
using Distributions
using DataFrames
using Libdl

export CudaMonteCarlo, CudaModel, CudaExprModel, ebn_init, ebn_shutdown

const LIB = Ref{String}(get(ENV, "EBNCUDA_LIB", "libebncuda.so"))

# ---- constants of include/ebncuda.h ------------------------------------------------------------
const EBN_LS = Dict(:vehicle_suspension => 1, :straub_frame => 2, :hydrogen_radius => 3,
    :polynomial => 4, :min_affine => 5, :straub_frame_r45 => 6, :straub_capacity => 7, :expression => 8)
# EBN_OP_* (expression program)
const OP_INPUT, OP_PARAM, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_ABS, OP_SQRT, OP_EXP, OP_LOG,
    OP_SIN, OP_COS, OP_MIN, OP_MAX, OP_POW, OP_POWI, OP_LT, OP_SELECT, OP_STORE = 0:20
const IN_PARAM, IN_NORMAL, IN_LOGNORMAL, IN_UNIFORM, IN_GUMBEL, IN_GAMMA, IN_EXP_AFFINE, IN_TABULATED = 0:7

# typedef struct ebn_input { int32 kind; int32 flags; double p[4]; double lo; double hi; }  (56 bytes)
struct EbnInput
    kind::Int32
    flags::Int32
    p::NTuple{4,Float64}
    lo::Float64
    hi::Float64
end

# typedef struct ebn_job (104 bytes); field order and widths as in the header
struct EbnJob
    limit_state::Int32
    n_consts::Int32
    consts::Ptr{Float64}
    n_scenarios::Int32
    n_inputs::Int32
    inputs::Ptr{EbnInput}
    n_per_scenario::Int64
    seed::UInt64
    strict::Int32
    flags::UInt32
    corr_chol::Ptr{Float64}
    stream_id::Ptr{UInt32}
    tables::Ptr{Float64}
    n_tables::Int64
    sample_offset::Int64
    shard_rank::Int32
    shard_world::Int32
end

@assert sizeof(EbnInput) == 56
@assert sizeof(EbnJob) == 104

last_error() = unsafe_string(ccall((:ebn_last_error, LIB[]), Cstring, ()))

function check(rc::Integer)
    if rc != 0
        msg = last_error()
        # evaluate! rewrites every exception of _evaluate_node into one of two fixed messages
        # (evaluate_net.jl:16-24): log the real one before throwing.
        @error "libebncuda: $msg (code $rc)"
        error("libebncuda: $msg (code $rc)")
    end
    return nothing
end

"ebn_init: `devices` are CUDA ordinals; more than one → single-process multi-GPU with an NCCL all-reduce."
function ebn_init(devices::Vector{Int}=Int[])
    dev = Int32.(devices)
    check(ccall((:ebn_init, LIB[]), Cint, (Ptr{Int32}, Cint), isempty(dev) ? C_NULL : dev, length(dev)))
end
ebn_shutdown() = check(ccall((:ebn_shutdown, LIB[]), Cint, ()))

# ---- the plug-in point: a new AbstractMonteCarlo subtype ---------------------------------------
"""
    CudaMonteCarlo(n; seed=0x5EED0001, strict=false, performance=nothing, fp32_sampling=false)

Simulation object routing a functional node to libebncuda. `strict=true` evaluates the limit
state with one IEEE operation per Julia operation (bit-exact `g`), the default contracts FMAs and
hoists reciprocals (|Δg| ~ 1e-13, count differences only on ties). For a chain of `CudaExprModel`s,
`performance` is the node's performance function as a quoted expression (`:(2 .- df.C)`); the
node's own `performance::Function` cannot be translated. Default: the last model's column.
`fp32_sampling=true` sets EBN_JOB_FP32_SAMPLING (untruncated normals drawn in fp32, limit state in fp64).
"""
struct CudaMonteCarlo <: UncertaintyQuantification.AbstractMonteCarlo
    n::Int
    seed::UInt64
    strict::Bool
    performance::Any
    fp32_sampling::Bool
end
CudaMonteCarlo(n::Integer; seed::Integer=0x5EED0001, strict::Bool=false, performance=nothing, fp32_sampling::Bool=false) =
    CudaMonteCarlo(n, UInt64(seed), strict, performance, fp32_sampling)
job_flags(sim::CudaMonteCarlo) = sim.fp32_sampling ? UInt32(16) : UInt32(0)   # EBN_JOB_FP32_SAMPLING

"""
    CudaModel(name, limit_state, consts, args)

A `UQModel` naming a limit state compiled into the kernels (`:vehicle_suspension`, `:straub_frame`,
`:polynomial`, ...). `args` are the DataFrame column names in the functor's argument order.
Arbitrary closures cannot run on the GPU; there is no CPU fallback.
"""
struct CudaModel <: UncertaintyQuantification.UQModel
    name::Symbol
    limit_state::Symbol
    consts::Vector{Float64}
    args::Vector{Symbol}
end

"""
    CudaExprModel(name, :(formula); constants...)

A `UQModel` whose closure body is given as a quoted Julia expression over the DataFrame columns
(written `df.x` or plain `x`) and named constants, e.g.

    CudaExprModel(:g2, :(4000 .* df.C .* Mg15 .- 8.6394); Mg15 = (M * g)^-1.5)

A chain of these (and a quoted `performance`) is compiled, in Julia's evaluation order, into the
expression program of include/ebncuda.h and then into a GPU kernel at run time (NVRTC).
`rand(Normal(mu, sigma))` / `randn()` inside a formula become hidden standard-normal inputs.
"""
struct CudaExprModel <: UncertaintyQuantification.UQModel
    name::Symbol
    formula::Any
    constants::Dict{Symbol,Float64}
end
CudaExprModel(name::Symbol, formula; constants...) =
    CudaExprModel(name, formula, Dict{Symbol,Float64}(k => Float64(v) for (k, v) in constants))

# -- Julia Expr -> expression program (post-order = Julia's evaluation order) ------------------------
mutable struct Emitter
    columns::Dict{Symbol,Int}      # 0-based column of inputs and stored models
    params::Dict{Symbol,Int}       # 0-based runtime constant
    hidden_base::Int
    n_hidden::Int
    code::Vector{Float64}
end
emit!(e::Emitter, op, arg=0.0) = push!(e.code, Float64(op), Float64(arg))
hidden!(e::Emitter) = (emit!(e, OP_INPUT, e.hidden_base + e.n_hidden); e.n_hidden += 1)

const UNARY = Dict(:sqrt => OP_SQRT, :exp => OP_EXP, :log => OP_LOG, :abs => OP_ABS, :sin => OP_SIN, :cos => OP_COS)
const BINARY = Dict(:+ => OP_ADD, :- => OP_SUB, :* => OP_MUL, :/ => OP_DIV)
undot(f::Symbol) = (s = String(f); startswith(s, ".") && length(s) > 1 ? Symbol(s[2:end]) : f)

function emit_expr!(e::Emitter, x)
    if x isa Real
        emit!(e, OP_CONST, x)
    elseif x isa Symbol
        if haskey(e.columns, x)
            emit!(e, OP_INPUT, e.columns[x])
        elseif haskey(e.params, x)
            emit!(e, OP_PARAM, e.params[x])
        elseif x === :pi || x === :π
            emit!(e, OP_CONST, Float64(pi))
        else
            error("unknown name $x: not an input column, an earlier model, or a constant")
        end
    elseif x isa QuoteNode
        emit_expr!(e, x.value)
    elseif x isa Expr && x.head === :. && length(x.args) == 2 && x.args[1] === :df      # df.x
        emit_expr!(e, x.args[2])
    elseif x isa Expr && x.head === :. && x.args[2] isa Expr && x.args[2].head === :tuple   # f.(args...)
        emit_expr!(e, Expr(:call, x.args[1], x.args[2].args...))
    elseif x isa Expr && x.head === :ref                                                 # g1[1] on a scalar
        emit_expr!(e, x.args[1])
    elseif x isa Expr && x.head === :if && length(x.args) == 3                           # c ? a : b
        emit_expr!(e, x.args[1]); emit_expr!(e, x.args[2]); emit_expr!(e, x.args[3]); emit!(e, OP_SELECT)
    elseif x isa Expr && x.head === :call
        f, args = undot(x.args[1]), x.args[2:end]
        if f === :^ && length(args) == 2
            emit_expr!(e, args[1])
            if args[2] isa Integer
                emit!(e, OP_POWI, args[2])
            else
                emit_expr!(e, args[2]); emit!(e, OP_POW)
            end
        elseif f === :- && length(args) == 1
            args[1] isa Real ? emit!(e, OP_CONST, -args[1]) : (emit_expr!(e, args[1]); emit!(e, OP_NEG))
        elseif haskey(BINARY, f) && length(args) >= 2           # a + b + c is +(a, b, c): folds left
            emit_expr!(e, args[1])
            for a in args[2:end]
                emit_expr!(e, a); emit!(e, BINARY[f])
            end
        elseif haskey(UNARY, f) && length(args) == 1
            emit_expr!(e, args[1]); emit!(e, UNARY[f])
        elseif f in (:min, :max) && length(args) >= 2
            emit_expr!(e, args[1])
            for a in args[2:end]
                emit_expr!(e, a); emit!(e, f === :min ? OP_MIN : OP_MAX)
            end
        elseif f in (:minimum, :maximum) && length(args) == 1 && args[1] isa Expr && args[1].head in (:vect, :tuple)
            emit_expr!(e, Expr(:call, f === :minimum ? :min : :max, args[1].args...))
        elseif f === :< && length(args) == 2
            emit_expr!(e, args[1]); emit_expr!(e, args[2]); emit!(e, OP_LT)
        elseif f === :> && length(args) == 2
            emit_expr!(e, args[2]); emit_expr!(e, args[1]); emit!(e, OP_LT)
        elseif f === :ifelse && length(args) == 3
            emit_expr!(e, args[1]); emit_expr!(e, args[2]); emit_expr!(e, args[3]); emit!(e, OP_SELECT)
        elseif f === :randn && isempty(args)
            hidden!(e)
        elseif f === :rand && length(args) == 1 && args[1] isa Expr && args[1].head === :call && args[1].args[1] === :Normal
            pa = args[1].args[2:end]
            if isempty(pa)
                hidden!(e)
            else                                                # mu + sigma * randn(), as Distributions.jl
                emit_expr!(e, pa[1]); emit_expr!(e, pa[2]); hidden!(e); emit!(e, OP_MUL); emit!(e, OP_ADD)
            end
        else
            error("function $f with $(length(args)) argument(s) has no device translation")
        end
    else
        error("unsupported syntax: $x")
    end
end

count_hidden(x) = x isa Expr ? ((x.head === :call && (x.args[1] === :randn || x.args[1] === :rand) ? 1 : 0) +
                                sum(count_hidden, x.args; init=0)) : 0

"Returns (consts, n_hidden, stored names): the expression program for `inputs` (column order of the job)."
function compile_program(inputs::Vector{Symbol}, models::Vector{CudaExprModel}, performance)
    params = Dict{Symbol,Float64}()
    for m in models, (k, v) in m.constants
        get!(params, k, v) == v || error("constant $k has two different values in one model chain")
    end
    names = collect(keys(params))
    nh = sum(count_hidden(m.formula) for m in models; init=0) + count_hidden(performance)
    e = Emitter(Dict(c => i - 1 for (i, c) in enumerate(inputs)), Dict(k => i - 1 for (i, k) in enumerate(names)),
        length(inputs), 0, Float64[])
    for (t, m) in enumerate(models)
        emit_expr!(e, m.formula)
        emit!(e, OP_STORE)
        e.columns[m.name] = length(inputs) + nh + t - 1
    end
    emit_expr!(e, performance)
    consts = vcat(Float64(length(names)), Float64(length(e.code) ÷ 2), [params[k] for k in names], e.code)
    return consts, nh, [m.name for m in models]
end

# ---- marshalling: UQInput -> ebn_input_t --------------------------------------------------------
abi(p::Parameter) = EbnInput(IN_PARAM, 0, (Float64(p.value), 0.0, 0.0, 0.0), -Inf, Inf)
abi(rv::RandomVariable) = abi(rv.dist, -Inf, Inf)
abi(d::Normal, lo, hi) = EbnInput(IN_NORMAL, 0, (d.μ, d.σ, 0.0, 0.0), lo, hi)
abi(d::LogNormal, lo, hi) = EbnInput(IN_LOGNORMAL, 0, (d.μ, d.σ, 0.0, 0.0), lo, hi)
abi(d::Uniform, lo, hi) = EbnInput(IN_UNIFORM, 0, (d.a, d.b, 0.0, 0.0), lo, hi)
abi(d::Gumbel, lo, hi) = EbnInput(IN_GUMBEL, 0, (d.μ, d.θ, 0.0, 0.0), lo, hi)
abi(d::Gamma, lo, hi) = (isinf(lo) && isinf(hi)) ? EbnInput(IN_GAMMA, 0, (d.α, d.θ, 0.0, 0.0), lo, hi) :
                        error("truncated Gamma must be passed as an inverse-CDF table (EBN_IN_TABULATED)")
abi(d::Truncated, lo, hi) = abi(d.untruncated, max(lo, something(d.lower, -Inf)), min(hi, something(d.upper, Inf)))
# `±Exponential(λ) + c` built by _approximate (discretize.jl:83-91) is a LocationScale
abi(d::Distributions.AffineDistribution{<:Any,<:Any,<:Exponential}, lo, hi) =
    EbnInput(IN_EXP_AFFINE, 0, (d.ρ.θ, sign(d.σ), d.μ, 0.0), lo, hi)
abi(x, args...) = error("marginal $(typeof(x)) has no device sampler")

function input_matrix(scenario_inputs::Vector{<:Vector}, columns::Vector{Symbol}, hidden::Vector{EbnInput})
    S = length(scenario_inputs)
    d = length(columns) + length(hidden)
    M = Vector{EbnInput}(undef, S * d)
    for (s, inputs) in enumerate(scenario_inputs)
        byname = Dict(i.name => i for i in inputs)
        for (j, c) in enumerate(columns)
            haskey(byname, c) || error("limit state needs input $c which the node's parents do not provide")
            M[(s-1)*d+j] = abi(byname[c])
        end
        for (j, h) in enumerate(hidden)
            M[(s-1)*d+length(columns)+j] = h
        end
    end
    return M, S, d
end

# ---- one call for all scenarios -------------------------------------------------------------------
"Returns (fail_count, pf, var) for every scenario; replaces S calls of `probability_of_failure`."
function pf_batch(model::CudaModel, scenario_inputs::Vector{<:Vector}, sim::CudaMonteCarlo;
    hidden::Vector{EbnInput}=EbnInput[])
    M, S, d = input_matrix(scenario_inputs, model.args, hidden)
    consts = model.consts
    counts = zeros(Int64, S)
    pf = zeros(Float64, S)
    var = zeros(Float64, S)
    GC.@preserve M consts begin
        job = Ref(EbnJob(EBN_LS[model.limit_state], length(consts), pointer(consts), S, d, pointer(M), sim.n,
            sim.seed, sim.strict ? 1 : 0, job_flags(sim), C_NULL, C_NULL, C_NULL, 0, 0, 0, 1))
        check(ccall((:ebn_pf_batch, LIB[]), Cint, (Ref{EbnJob}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            job, counts, pf, var))
    end
    return counts, pf, var
end

"Expression chain: all scenarios of a node in one call. `columns` = the input names in DataFrame order."
function pf_batch(models::Vector{CudaExprModel}, performance, scenario_inputs::Vector{<:Vector}, sim::CudaMonteCarlo)
    columns = Symbol[i.name for i in scenario_inputs[1]]
    consts, nh, _ = compile_program(columns, models, performance)
    return pf_batch(CudaModel(models[end].name, :expression, consts, columns), scenario_inputs, sim;
        hidden=fill(abi(Normal(), -Inf, Inf), nh))
end

# the in-model `rand(Normal(...))` draws of the Straub capacities travel as hidden N(0,1) inputs
const N_HIDDEN = Dict(:straub_frame => 5, :straub_frame_r45 => 3, :straub_capacity => 1)
hidden_inputs(m::CudaModel) = fill(abi(Normal(), -Inf, Inf), get(N_HIDDEN, m.limit_state, 0))

# UQ.jl-style single-scenario entry (demo/vehicle_suspension.jl:68)
function UncertaintyQuantification.probability_of_failure(models::Union{CudaModel,Vector{CudaModel}},
    performance::Function, inputs::Vector{<:UQInput}, sim::CudaMonteCarlo)
    model = models isa Vector ? models[end] : models
    _, pf, var = pf_batch(model, [inputs], sim; hidden=hidden_inputs(model))
    return pf[1], sqrt(var[1]), DataFrame()
end

# ---- `_evaluate_node` specialised on the simulation type ------------------------------------------
# Same scenario enumeration, marshalling and CPT write-back as evaluate_node.jl:41-122; only the
# scenario loop is replaced by one library call.
function EnhancedBayesianNetworks._evaluate_node(net::EnhancedBayesianNetwork,
    node::DiscreteFunctionalNode{<:Any}, collect_samples::Bool, sim::CudaMonteCarlo)
    ancestors = discrete_ancestors(net, node)
    combos = sort(vec(collect(Iterators.product(states.(ancestors)...))))
    owner(s) = (hits = filter(n -> s ∈ states(n), ancestors); isempty(hits) ? node.name : hits[1].name)
    evidences = isempty(combos[1]) ? [Evidence()] : map(c -> Evidence(owner(s) => s for s in c), combos)
    cpt_names = isempty(combos[1]) ? [node.name] : vcat([a.name for a in ancestors], node.name)
    cpt = DiscreteConditionalProbabilityTable{PreciseDiscreteProbability}(cpt_names)
    scen = map(ev -> mapreduce(p -> _uq_inputs(p, ev), vcat, parents(net, node)[3]; init=UQInput[]), evidences)
    _, pf, var = if all(m -> m isa CudaExprModel, node.models)
        pf_batch(Vector{CudaExprModel}(node.models), something(sim.performance, node.models[end].name), scen, sim)
    else
        model = node.models[end]::CudaModel
        pf_batch(model, scen, sim; hidden=hidden_inputs(model))
    end
    info = Dict{AbstractVector{Symbol},Dict}()
    for (s, ev) in enumerate(evidences)
        fail_ev = copy(ev); fail_ev[node.name] = Symbol(string(node.name) * "_fail")
        safe_ev = copy(ev); safe_ev[node.name] = Symbol(string(node.name) * "_safe")
        cpt[fail_ev] = pf[s]
        cpt[safe_ev] = 1 - pf[s]
        if collect_samples
            # the n-row DataFrame cannot be materialised at 1e8+ samples: empty frame, replay with
            # ebn_sample_matrix if rows are needed (INTEGRATION.md "samples")
            info[collect(values(ev))] = Dict(:cov => sqrt(var[s]), :samples => DataFrame())
        end
    end
    return DiscreteNode(node.name, cpt, node.parameters, info)
end

# ---- ContinuousFunctionalNode: histogram instead of a KDE over n outputs ---------------------------

"ebn_hist_batch for all scenarios: (counts[S, nedges+1], sum[S], sumsq[S]); first/last column = under/overflow."
function hist_batch(model::CudaModel, scenario_inputs::Vector{<:Vector}, sim::CudaMonteCarlo, edges::Vector{Float64};
    hidden::Vector{EbnInput}=EbnInput[])
    M, S, d = input_matrix(scenario_inputs, model.args, hidden)
    consts = model.consts
    nb = length(edges) + 1
    counts = zeros(Int64, nb, S)            # column-major: scenario-major rows of the C layout
    s1 = zeros(Float64, S)
    s2 = zeros(Float64, S)
    GC.@preserve M consts begin
        job = Ref(EbnJob(EBN_LS[model.limit_state], length(consts), pointer(consts), S, d, pointer(M), sim.n,
            sim.seed, sim.strict ? 1 : 0, job_flags(sim), C_NULL, C_NULL, C_NULL, 0, 0, 0, 1))
        check(ccall((:ebn_hist_batch, LIB[]), Cint,
            (Ref{EbnJob}, Ptr{Float64}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            job, edges, length(edges), counts, s1, s2))
    end
    return permutedims(counts), s1, s2
end

"Piecewise-uniform distribution of a histogram: what stands where the reference puts `EmpiricalDistribution(y)`."
function histogram_distribution(edges::Vector{Float64}, inner_counts::Vector{Int64})
    keep = findall(>(0), inner_counts)
    w = inner_counts[keep] ./ sum(inner_counts[keep])
    return MixtureModel([Uniform(edges[i], edges[i+1]) for i in keep], w)
end

# Like the Python mirror (uq.output_histograms): a pilot of <= 65536 samples per scenario gives the
# output range from the moments, then ONE call histograms every scenario on a shared grid.
function EnhancedBayesianNetworks._evaluate_node(net::EnhancedBayesianNetwork,
    node::ContinuousFunctionalNode, collect_samples::Bool, sim::CudaMonteCarlo)
    all(EnhancedBayesianNetworks.isprecise.(parents(net, node)[3])) ||
        error("node $(node.name) is a continuousfunctionalnode with at least one parent with Interval or p-boxes in its distributions. No method for extracting failure probability p-box have been implemented yet")
    ancestors = discrete_ancestors(net, node)
    combos = sort(vec(collect(Iterators.product(states.(ancestors)...))))
    owner(s) = filter(n -> s ∈ states(n), filter(x -> isa(x, DiscreteNode), net.nodes))[1].name
    evidences = map(c -> Evidence(owner(s) => s for s in c), combos)
    if isempty(combos[1])
        cpt = ContinuousConditionalProbabilityTable{PreciseContinuousInput}()
        discretization = ExactDiscretization(node.discretization.intervals)
    else
        cpt = ContinuousConditionalProbabilityTable{PreciseContinuousInput}([i.name for i in ancestors])
        discretization = node.discretization
    end
    scen = map(ev -> mapreduce(p -> _uq_inputs(p, ev), vcat, parents(net, node)[3]; init=UQInput[]), evidences)
    model = node.models[end]::CudaModel
    hid = hidden_inputs(model)
    pilot = CudaMonteCarlo(min(sim.n, 1 << 16), sim.seed, sim.strict, nothing, sim.fp32_sampling)
    _, s1, s2 = hist_batch(model, scen, pilot, [0.0]; hidden=hid)          # moments only
    μ = s1 ./ pilot.n
    σ = sqrt.(max.(s2 ./ pilot.n .- μ .^ 2, 0.0))
    lo, hi = minimum(μ .- 8 .* σ), maximum(μ .+ 8 .* σ)
    hi > lo || ((lo, hi) = (lo - 1.0, hi + 1.0))
    edges = collect(range(lo, hi; length=1025))
    counts, _, _ = hist_batch(model, scen, sim, edges; hidden=hid)         # ONE call, all scenarios
    info = Dict{Vector{Symbol},Dict}()
    for (s, ev) in enumerate(evidences)
        cpt[ev] = histogram_distribution(edges, counts[s, 2:end-1])
        collect_samples && (info[collect(values(ev))] = Dict(:samples => DataFrame()))
    end
    return ContinuousNode(node.name, cpt, discretization, info)
end

# dispatch hook: called from _evaluate_node(net, node::DiscreteFunctionalNode, collect_samples)
# when `node.simulation isa CudaMonteCarlo` (one-line patch in evaluate_node.jl, see INTEGRATION.md)

end # module
