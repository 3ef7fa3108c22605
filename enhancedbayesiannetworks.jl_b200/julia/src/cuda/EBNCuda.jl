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
    _uq_inputs, isprecise, DiscreteConditionalProbabilityTable, ContinuousConditionalProbabilityTable,
    PreciseDiscreteProbability, ImpreciseDiscreteProbability, PreciseContinuousInput, ExactDiscretization
using UncertaintyQuantification
using Distributions
using DataFrames
using LinearAlgebra

export CudaMonteCarlo, CudaSobolSampling, CudaDoubleLoop, CudaRandomSlicing, CudaModel, CudaExprModel, ebn_init, ebn_shutdown

const LIB = Ref{String}(get(ENV, "EBNCUDA_LIB", "libebncuda.so"))

# ---- constants of include/ebncuda.h ------------------------------------------------------------
const EBN_LS = Dict(:vehicle_suspension => 1, :straub_frame => 2, :hydrogen_radius => 3,
    :polynomial => 4, :min_affine => 5, :straub_frame_r45 => 6, :straub_capacity => 7, :expression => 8)
# EBN_OP_* (expression program)
const OP_INPUT, OP_PARAM, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_NEG, OP_ABS, OP_SQRT, OP_EXP, OP_LOG,
    OP_SIN, OP_COS, OP_MIN, OP_MAX, OP_POW, OP_POWI, OP_LT, OP_SELECT, OP_STORE = 0:20
const IN_PARAM, IN_NORMAL, IN_LOGNORMAL, IN_UNIFORM, IN_GUMBEL, IN_GAMMA, IN_EXP_AFFINE, IN_TABULATED = 0:7
# EBN_JOB_* flags
const JOB_COPULA_SHARED, JOB_PARTIAL, JOB_JIT_PLAN, JOB_NO_JIT_PLAN, JOB_FP32_SAMPLING, JOB_QMC_SOBOL = UInt32.((1, 2, 4, 8, 16, 32))
const TABLE_POINTS = 4097      # points of an inverse-CDF table (uniform grid in u)

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
`qmc=true` (or the constructor `CudaSobolSampling`) sets EBN_JOB_QMC_SOBOL: UQ.jl's `SobolSampling` on the
device — sample i is point i of an Owen-scrambled Sobol sequence, every marginal through its inverse CDF.
"""
struct CudaMonteCarlo <: UncertaintyQuantification.AbstractMonteCarlo
    n::Int
    seed::UInt64
    strict::Bool
    performance::Any
    fp32_sampling::Bool
    qmc::Bool
end
CudaMonteCarlo(n::Integer; seed::Integer=0x5EED0001, strict::Bool=false, performance=nothing,
    fp32_sampling::Bool=false, qmc::Bool=false) =
    CudaMonteCarlo(n, UInt64(seed), strict, performance, fp32_sampling, qmc)
"Quasi-Monte Carlo counterpart of UQ.jl's `SobolSampling(n)`; a power of two for `n` keeps the point set a net."
CudaSobolSampling(n::Integer; kwargs...) = CudaMonteCarlo(n; qmc=true, kwargs...)
job_flags(sim::CudaMonteCarlo) = (sim.fp32_sampling ? JOB_FP32_SAMPLING : UInt32(0)) | (sim.qmc ? JOB_QMC_SOBOL : UInt32(0))

"""
    CudaDoubleLoop(lb::CudaMonteCarlo, ub = lb; method = :optimise, grid = 33)

The device counterpart of UQ.jl's `DoubleLoop` (evaluate_node.jl:104-110): for every scenario the failure
probability is minimised and maximised over the box of the imprecise inputs (an `Interval` becomes a
`Parameter`, a `ProbabilityBox{T}` a `RandomVariable(T(x...))`). All outer candidates of all scenarios
are ONE `ebn_pf_batch` call; the candidates of a scenario share its Philox stream (`stream_id`), so the
inner pf is a deterministic function of the candidate. `:grid` sweeps `grid` points per coordinate;
`:optimise` is a compass search from the box midpoint (UQ.jl starts BOBYQA there; the two optimisers
are not bit-comparable, both work on the same objective).
"""
struct CudaDoubleLoop <: UncertaintyQuantification.AbstractSimulation
    lb::CudaMonteCarlo
    ub::CudaMonteCarlo
    method::Symbol
    grid::Int
end
CudaDoubleLoop(lb::CudaMonteCarlo, ub::CudaMonteCarlo=lb; method::Symbol=:optimise, grid::Integer=33) =
    CudaDoubleLoop(lb, ub, method, grid)

"""
    CudaRandomSlicing(lb::CudaMonteCarlo, ub = lb; levels = [0.0, 0.5, 1.0])

The device counterpart of UQ.jl's `RandomSlicing` (evaluate_node.jl:111-117) for chains of
`CudaExprModel`s: per sample an `Interval` is its box, a p-box is sliced by one standard variate into
[min, max] of the family's quantile over the vertices of its parameter box, the chain is evaluated at
every combination of `levels` of the per-sample intervals inside the kernel and the indicators of the
smallest / largest g are counted (pf_ub / pf_lb). Both sides are scenarios of one call on the same stream.
"""
struct CudaRandomSlicing <: UncertaintyQuantification.AbstractSimulation
    lb::CudaMonteCarlo
    ub::CudaMonteCarlo
    levels::Vector{Float64}
end
CudaRandomSlicing(lb::CudaMonteCarlo, ub::CudaMonteCarlo=lb; levels=[0.0, 0.5, 1.0]) =
    CudaRandomSlicing(lb, ub, Float64.(levels))

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
function compile_program(inputs::Vector{Symbol}, models::Vector{CudaExprModel}, performance;
    constants::Dict{Symbol,Float64}=Dict{Symbol,Float64}())
    params = copy(constants)
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
# `pool` collects the inverse-CDF tables of the job (ebn_job_t.tables / n_tables): an input without a
# closed-form sampler on the device travels as EBN_IN_TABULATED = (offset, npts) into it.
abi(p::Parameter, pool) = EbnInput(IN_PARAM, 0, (Float64(p.value), 0.0, 0.0, 0.0), -Inf, Inf)
abi(rv::RandomVariable, pool) = abi(rv.dist, -Inf, Inf, pool)
abi(i::Union{Interval,ProbabilityBox}, pool) =
    error("input $(i.name) is imprecise: the node needs CudaDoubleLoop or CudaRandomSlicing")
abi(d::Normal, lo, hi, pool) = EbnInput(IN_NORMAL, 0, (d.μ, d.σ, 0.0, 0.0), lo, hi)
abi(d::LogNormal, lo, hi, pool) = EbnInput(IN_LOGNORMAL, 0, (d.μ, d.σ, 0.0, 0.0), lo, hi)
abi(d::Uniform, lo, hi, pool) = EbnInput(IN_UNIFORM, 0, (d.a, d.b, 0.0, 0.0), lo, hi)
abi(d::Gumbel, lo, hi, pool) = EbnInput(IN_GUMBEL, 0, (d.μ, d.θ, 0.0, 0.0), lo, hi)
# untruncated Gamma: Marsaglia-Tsang on the device; truncated Gamma (a discretised root,
# discretize.jl:54) has no closed-form quantile there: inverse-CDF table
abi(d::Gamma, lo, hi, pool) = (isinf(lo) && isinf(hi)) ? EbnInput(IN_GAMMA, 0, (d.α, d.θ, 0.0, 0.0), lo, hi) :
                              tabulated(d, lo, hi, pool)
abi(d::Truncated, lo, hi, pool) =
    abi(d.untruncated, max(lo, something(d.lower, -Inf)), min(hi, something(d.upper, Inf)), pool)
# `±Exponential(λ) + c` built by _approximate (discretize.jl:83-91) is a LocationScale
abi(d::Distributions.AffineDistribution{<:Any,<:Any,<:Exponential}, lo, hi, pool) =
    EbnInput(IN_EXP_AFFINE, 0, (d.ρ.θ, sign(d.σ), d.μ, 0.0), lo, hi)
# everything else that has cdf and quantile: the piecewise-uniform MixtureModel an evaluated
# ContinuousFunctionalNode carries (histogram_distribution), UQ.jl's EmpiricalDistribution, ...
abi(d::UnivariateDistribution, lo, hi, pool) = tabulated(d, lo, hi, pool)
abi(x, args...) = error("marginal $(typeof(x)) has no device sampler")

"Quantiles of `truncated(d, lo, hi)` on a uniform grid of `npts` probabilities, appended to the job's table pool."
function tabulated(d, lo, hi, pool::Vector{Float64}; npts::Int=TABLE_POINTS)
    a, b = max(lo, minimum(d)), min(hi, maximum(d))
    Fl, Fh = cdf(d, a), cdf(d, b)
    off = length(pool)
    if !(Fh > Fl)
        # a window the (histogram) distribution gives no mass to; its bin has probability 0 in the
        # discretised parent, so any proper law on the window serves: uniform
        (isfinite(a) && isfinite(b)) || error("$d has no mass in [$a, $b]")
        append!(pool, range(a, b; length=npts))
    else
        q = Vector{Float64}(undef, npts)
        for k in 1:npts
            pk = clamp(Fl + (k - 1) / (npts - 1) * (Fh - Fl), isfinite(a) ? Fl : 1e-12, isfinite(b) ? Fh : 1 - 1e-12)
            q[k] = quantile(d, pk)
        end
        isfinite(a) && (q[1] = a)
        isfinite(b) && (q[end] = b)
        accumulate!(max, q, q)
        append!(pool, q)
    end
    return EbnInput(IN_TABULATED, 0, (Float64(off), Float64(npts), 0.0, 0.0), -Inf, Inf)
end

# UQ.jl JointDistribution(marginals, GaussianCopula(R)) (SURVEY A3): the marginals become ordinary
# columns, the correlation of their normal scores goes into ebn_job_t.corr_chol.
flatten(inputs) = mapreduce(i -> i isa JointDistribution ? collect(i.marginals) : [i], vcat, inputs; init=UQInput[])
input_names(inputs) = Symbol[i.name for i in flatten(inputs)]

"Everything an ebn_job_t points at, owned on the Julia side for the duration of the call."
struct JobData
    inputs::Vector{EbnInput}     # [S*d], row-major by scenario
    S::Int
    d::Int
    tables::Vector{Float64}      # inverse-CDF table pool (may be empty)
    chol::Vector{Float64}        # [S*d*d] row-major lower-triangular factors, empty without a copula
    stream::Vector{UInt32}       # [S] Philox stream per scenario, empty = scenario index
end

function build_job(scenario_inputs::Vector{<:Vector}, columns::Vector{Symbol}, hidden::Int;
    stream::Vector{UInt32}=UInt32[], qmc::Bool=false)
    S = length(scenario_inputs)
    d = length(columns) + hidden
    M = Vector{EbnInput}(undef, S * d)
    pool = Float64[]
    chol = Float64[]
    any_copula = any(i -> i isa JointDistribution, scenario_inputs[1])
    std_normal = EbnInput(IN_NORMAL, 0, (0.0, 1.0, 0.0, 0.0), -Inf, Inf)   # in-model randn() draws
    for (s, inputs) in enumerate(scenario_inputs)
        byname = Dict(i.name => i for i in flatten(inputs))
        for (j, c) in enumerate(columns)
            haskey(byname, c) || error("limit state needs input $c which the node's parents do not provide")
            inp = byname[c]
            # no closed-form quantile on the device: Gamma under a copula or a Sobol sequence is shipped tabulated
            M[(s-1)*d+j] = ((any_copula || qmc) && inp isa RandomVariable && inp.dist isa Gamma) ?
                           tabulated(inp.dist, -Inf, Inf, pool; npts=8193) : abi(inp, pool)
        end
        for j in 1:hidden
            M[(s-1)*d+length(columns)+j] = std_normal
        end
        if any_copula
            R = Matrix{Float64}(I, d, d)
            for joint in filter(i -> i isa JointDistribution, inputs)
                C = joint.copula.correlation
                idx = [findfirst(==(m.name), columns) for m in joint.marginals]
                for (a, ia) in enumerate(idx), (b, ib) in enumerate(idx)
                    (ia === nothing || ib === nothing) || (R[ia, ib] = C[a, b])
                end
            end
            Lc = cholesky(Symmetric(R)).L
            append!(chol, vec(permutedims(Matrix(Lc))))      # row-major
        end
    end
    return JobData(M, S, d, pool, chol, stream)
end

function with_job(f, ls::Symbol, consts::Vector{Float64}, jd::JobData, sim::CudaMonteCarlo)
    inputs, tables, chol, stream = jd.inputs, jd.tables, jd.chol, jd.stream
    GC.@preserve inputs consts tables chol stream begin
        job = Ref(EbnJob(EBN_LS[ls], length(consts), isempty(consts) ? C_NULL : pointer(consts), jd.S, jd.d,
            pointer(inputs), sim.n, sim.seed, sim.strict ? 1 : 0, job_flags(sim),
            isempty(chol) ? C_NULL : pointer(chol), isempty(stream) ? C_NULL : pointer(stream),
            isempty(tables) ? C_NULL : pointer(tables), length(tables), 0, 0, 1))
        f(job)
    end
end

"ebn_pf_batch: (fail_count, pf, var) for every scenario of the job."
function run_pf(ls::Symbol, consts::Vector{Float64}, jd::JobData, sim::CudaMonteCarlo)
    counts = zeros(Int64, jd.S)
    pf = zeros(Float64, jd.S)
    var = zeros(Float64, jd.S)
    with_job(ls, consts, jd, sim) do job
        check(ccall((:ebn_pf_batch, LIB[]), Cint, (Ref{EbnJob}, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            job, counts, pf, var))
    end
    return counts, pf, var
end

"ebn_hist_batch: (counts[S, nedges+1], sum[S], sumsq[S]); first/last column = under/overflow (NaN counts as overflow)."
function run_hist(ls::Symbol, consts::Vector{Float64}, jd::JobData, sim::CudaMonteCarlo, edges::Vector{Float64})
    nb = length(edges) + 1
    counts = zeros(Int64, nb, jd.S)          # column-major: scenario-major rows of the C layout
    s1 = zeros(Float64, jd.S)
    s2 = zeros(Float64, jd.S)
    with_job(ls, consts, jd, sim) do job
        check(ccall((:ebn_hist_batch, LIB[]), Cint,
            (Ref{EbnJob}, Ptr{Float64}, Cint, Ptr{Int64}, Ptr{Float64}, Ptr{Float64}),
            job, edges, length(edges), counts, s1, s2))
    end
    return permutedims(counts), s1, s2
end

# ---- model chain -> (limit state, consts, columns, hidden inputs) -----------------------------------
# the in-model `rand(Normal(...))` draws of the Straub capacities travel as hidden N(0,1) inputs
const N_HIDDEN = Dict(:straub_frame => 5, :straub_frame_r45 => 3, :straub_capacity => 1)

struct Lowered
    limit_state::Symbol
    consts::Vector{Float64}
    columns::Vector{Symbol}
    hidden::Int
end

"`performance` of an expression chain: the quoted formula carried by the simulation, else the last model's column."
function lower(models::Vector, performance, inputs::Vector{Symbol})
    if all(m -> m isa CudaExprModel, models)
        consts, nh, _ = compile_program(inputs, Vector{CudaExprModel}(models), something(performance, models[end].name))
        return Lowered(:expression, consts, inputs, nh)
    end
    model = models[end]
    model isa CudaModel || error("model $(model.name) has no device functor: use CudaModel (catalogue) or " *
                                 "CudaExprModel (the closure's formula); a Julia closure cannot run on the GPU " *
                                 "and there is no CPU fallback")
    return Lowered(model.limit_state, model.consts, model.args, get(N_HIDDEN, model.limit_state, 0))
end

"All scenarios of a node in ONE call; replaces S calls of `probability_of_failure` (evaluate_node.jl:83)."
function pf_batch(models::Vector, performance, scenario_inputs::Vector{<:Vector}, sim::CudaMonteCarlo;
    stream::Vector{UInt32}=UInt32[])
    low = lower(models, performance, input_names(scenario_inputs[1]))
    return run_pf(low.limit_state, low.consts, build_job(scenario_inputs, low.columns, low.hidden; stream=stream, qmc=sim.qmc), sim)
end

# UQ.jl-style single-scenario entry (demo/vehicle_suspension.jl:68)
function UncertaintyQuantification.probability_of_failure(models::Union{CudaModel,CudaExprModel,Vector{<:Union{CudaModel,CudaExprModel}}},
    performance::Function, inputs::Vector{<:UQInput}, sim::CudaMonteCarlo)
    _, pf, var = pf_batch(models isa Vector ? collect(models) : [models], sim.performance, [inputs], sim)
    return pf[1], sqrt(var[1]), DataFrame()
end

# ---- imprecise inputs: DoubleLoop and RandomSlicing (evaluate_node.jl:104-117) -------------------------
is_imprecise(i) = i isa Interval || i isa ProbabilityBox

"(input position, parameter position or 0, lb, ub) per free coordinate of the outer loop."
function imprecise_dims(inputs)
    dims = Tuple{Int,Int,Float64,Float64}[]
    for (pos, i) in enumerate(inputs)
        if i isa Interval
            push!(dims, (pos, 0, Float64(i.lb), Float64(i.ub)))
        elseif i isa ProbabilityBox
            for (k, p) in enumerate(i.parameters)
                push!(dims, (pos, k, Float64(p.lb), Float64(p.ub)))
            end
        end
    end
    return dims
end

pbox_family(::ProbabilityBox{T}) where {T} = T

"UQ.jl `map_to_precise`: an Interval becomes Parameter(x), a p-box RandomVariable(T(x...)) (truncated to its support bounds)."
function realise(inputs, dims, x)
    out = Vector{UQInput}(collect(inputs))
    vals = Dict{Int,Dict{Int,Float64}}()
    for ((pos, k, _, _), v) in zip(dims, x)
        k == 0 ? (out[pos] = Parameter(v, inputs[pos].name)) : (get!(vals, pos, Dict{Int,Float64}())[k] = v)
    end
    for (pos, pv) in vals
        pb = inputs[pos]
        dist = pbox_family(pb)([pv[k] for k in 1:length(pb.parameters)]...)
        (isfinite(pb.lb) || isfinite(pb.ub)) && (dist = truncated(dist, pb.lb, pb.ub))
        out[pos] = RandomVariable(dist, pb.name)
    end
    return out
end

"pf of every (scenario, candidate) in ONE batch; the candidates of a scenario share its Philox stream."
function eval_candidates(models, performance, scenario_inputs, dims_per_s, cands_per_s, sim::CudaMonteCarlo)
    flat = Vector{Vector{UQInput}}()
    stream = UInt32[]
    for (s, (inputs, dims, cands)) in enumerate(zip(scenario_inputs, dims_per_s, cands_per_s)), x in cands
        push!(flat, realise(inputs, dims, x))
        push!(stream, UInt32(s - 1))
    end
    _, pf, _ = pf_batch(models, performance, flat, sim; stream=stream)
    out = Vector{Vector{Float64}}()
    k = 0
    for cands in cands_per_s
        push!(out, pf[k+1:k+length(cands)])
        k += length(cands)
    end
    return out
end

"Per scenario (pf_lb, pf_ub) over the box of the imprecise inputs."
function double_loop(models, performance, scenario_inputs, sim::CudaDoubleLoop)
    dims_per_s = map(imprecise_dims, scenario_inputs)
    # UQ.jl's DoubleLoop asserts on precise-only inputs; evaluate! turns the AssertionError into its
    # "are no longer imprecise. A prices simulation technique must be selected!" message (evaluate_net.jl:17-18)
    any(isempty, dims_per_s) && throw(AssertionError("DoubleLoop needs at least one Interval / ProbabilityBox input"))
    los = [Float64[d[3] for d in dims] for dims in dims_per_s]
    his = [Float64[d[4] for d in dims] for dims in dims_per_s]
    S = length(scenario_inputs)
    if sim.method === :grid
        grid_of(lo, hi) = vec([collect(p) for p in Iterators.product([range(a, b; length=sim.grid) for (a, b) in zip(lo, hi)]...)])
        cands = map(grid_of, los, his)
        pf = eval_candidates(models, performance, scenario_inputs, dims_per_s, cands, sim.lb)
        return [(minimum(p), maximum(p)) for p in pf]
    end
    # compass search from the box midpoint, both bounds and all scenarios in lock-step; the vertices first
    best_lo = fill(Inf, S); best_hi = fill(-Inf, S)
    x_lo = [0.5 .* (a .+ b) for (a, b) in zip(los, his)]
    x_hi = deepcopy(x_lo)
    step = [0.25 .* (b .- a) for (a, b) in zip(los, his)]
    first_round = true
    for _ in 1:14
        cands = Vector{Vector{Vector{Float64}}}()
        for s in 1:S
            c = [copy(x_lo[s]), copy(x_hi[s])]
            for x0 in (x_lo[s], x_hi[s]), k in eachindex(x0), sg in (-1.0, 1.0)
                y = copy(x0)
                y[k] = clamp(y[k] + sg * step[s][k], los[s][k], his[s][k])
                push!(c, y)
            end
            first_round && append!(c, vec([collect(v) for v in Iterators.product(zip(los[s], his[s])...)]))
            push!(cands, c)
        end
        pf = eval_candidates(models, performance, scenario_inputs, dims_per_s, cands, sim.lb)
        moved = false
        for s in 1:S
            pmin, imin = findmin(pf[s])
            pmax, imax = findmax(pf[s])
            if pmin < best_lo[s]
                moved |= isfinite(best_lo[s]); best_lo[s] = pmin; x_lo[s] = copy(cands[s][imin])
            end
            if pmax > best_hi[s]
                moved |= isfinite(best_hi[s]); best_hi[s] = pmax; x_hi[s] = copy(cands[s][imax])
            end
        end
        if !moved && !first_round
            step = [0.5 .* st for st in step]
            all(all(st .<= 1e-3 .* (b .- a) .+ 1e-300) for (st, a, b) in zip(step, los, his)) && break
        end
        first_round = false
    end
    return collect(zip(best_lo, best_hi))
end

# quantile of a p-box family as an expression of its two parameters and ONE standard variate per sample
const PBOX_QUANTILE = Dict{Any,Tuple{Symbol,Function}}(
    Normal => (:normal, (p0, p1, s) -> :($p0 + $p1 * $s)),
    LogNormal => (:normal, (p0, p1, s) -> :(exp($p0 + $p1 * $s))),
    Uniform => (:uniform, (p0, p1, s) -> :($p0 + ($p1 - $p0) * $s)),
    Gumbel => (:uniform, (p0, p1, s) -> :($p0 - $p1 * log(-log($s)))))

"Replaces column names (`x` or `df.x`) by expressions."
function subst(x, table::Dict{Symbol,Any})
    x isa Symbol && return get(table, x, x)
    x isa QuoteNode && return x
    if x isa Expr
        x.head === :. && length(x.args) == 2 && x.args[1] === :df && x.args[2] isa QuoteNode &&
            return get(table, x.args[2].value, x)
        x.head === :call && return Expr(:call, x.args[1], map(a -> subst(a, table), x.args[2:end])...)
        return Expr(x.head, map(a -> subst(a, table), x.args)...)
    end
    return x
end

"""
Per scenario ((pf_lb, pf_ub), σ_lb, σ_ub): RandomSlicing for chains of CudaExprModels, no monotonicity
assumed, p-boxes with two interval parameters. The chain is rewritten into ONE formula per candidate
(every combination of `sim.levels` of the per-sample intervals) under min / max; a Parameter column
`__side` selects max g (lower bound of pf) or min g (upper bound). One `ebn_pf_batch` call.
"""
function random_slicing(models::Vector, performance, scenario_inputs, sim::CudaRandomSlicing)
    all(m -> m isa CudaExprModel, models) ||
        error("CudaRandomSlicing needs CudaExprModel chains (the per-sample extremes are found by rewriting the formulas)")
    base = scenario_inputs[1]
    any(is_imprecise, base) || throw(AssertionError("RandomSlicing needs at least one Interval / ProbabilityBox input"))
    columns = Symbol[]
    makers = Function[]                       # scenario inputs -> the UQ input of that column
    ends = Dict{Symbol,Tuple{Any,Any}}()      # imprecise input -> (per-sample lower end, upper end) as expressions
    for (pos, inp) in enumerate(base)
        if inp isa Interval
            lo, hi = Symbol(inp.name, :__lo), Symbol(inp.name, :__hi)
            push!(columns, lo, hi)
            push!(makers, ins -> Parameter(Float64(ins[pos].lb), lo), ins -> Parameter(Float64(ins[pos].ub), hi))
            ends[inp.name] = (lo, hi)
        elseif inp isa ProbabilityBox
            fam = pbox_family(inp)
            (haskey(PBOX_QUANTILE, fam) && length(inp.parameters) == 2) ||
                error("CudaRandomSlicing has no quantile formula for a $fam p-box")
            (isinf(inp.lb) && isinf(inp.ub)) || error("CudaRandomSlicing takes untruncated p-boxes")
            kind, q = PBOX_QUANTILE[fam]
            sname = Symbol(inp.name, :__s)
            push!(columns, sname)
            push!(makers, ins -> RandomVariable(kind === :normal ? Normal() : Uniform(0.0, 1.0), sname))
            pn = Vector{Tuple{Symbol,Symbol}}()
            for k in 1:2
                nlo, nhi = Symbol(inp.name, :__p, k, :lo), Symbol(inp.name, :__p, k, :hi)
                push!(columns, nlo, nhi)
                push!(makers, ins -> Parameter(Float64(ins[pos].parameters[k].lb), nlo),
                    ins -> Parameter(Float64(ins[pos].parameters[k].ub), nhi))
                push!(pn, (nlo, nhi))
            end
            verts = [q(a, b, sname) for a in pn[1] for b in pn[2]]
            ends[inp.name] = (Expr(:call, :min, verts...), Expr(:call, :max, verts...))
        else
            push!(columns, inp.name)
            push!(makers, ins -> ins[pos])
        end
    end
    push!(columns, :__side)
    exprs = CudaExprModel[m for m in models]
    any(m -> count_hidden(m.formula) > 0, exprs) && error("CudaRandomSlicing: in-model randomness cannot be sliced")
    perf = something(sim.lb.performance, exprs[end].name)
    names = collect(keys(ends))
    cands = Any[]
    for combo in Iterators.product(fill(sim.levels, length(names))...)
        table = Dict{Symbol,Any}()
        for (name, lam) in zip(names, combo)
            lo, hi = ends[name]
            table[name] = lam == 0.0 ? lo : lam == 1.0 ? hi : :($lo + $lam * ($hi - $lo))
        end
        for m in exprs                       # a later model sees the earlier ones already substituted
            table[m.name] = subst(m.formula, table)
        end
        push!(cands, subst(perf, table))
    end
    gmin = length(cands) > 1 ? Expr(:call, :min, cands...) : cands[1]
    gmax = length(cands) > 1 ? Expr(:call, :max, cands...) : cands[1]
    final = :(ifelse(__side < 0.5, $gmax, $gmin))       # side 0: lower bound of pf, side 1: upper bound
    constants = Dict{Symbol,Float64}()
    for m in exprs
        merge!(constants, m.constants)
    end
    consts, _, _ = compile_program(columns, CudaExprModel[], final; constants=constants)
    rows = Vector{Vector{UQInput}}()
    stream = UInt32[]
    for (s, ins) in enumerate(scenario_inputs), side in (0.0, 1.0)
        push!(rows, UQInput[[mk(ins) for mk in makers]; Parameter(side, :__side)])
        push!(stream, UInt32(s - 1))
    end
    _, pf, _ = run_pf(:expression, consts, build_job(rows, columns, 0; stream=stream), sim.lb)
    n = sim.lb.n
    sd(p) = sqrt(max(p - p^2, 0) / n)
    return map(s -> ((pf[2s-1], pf[2s]), sd(pf[2s-1]), sd(pf[2s])), 1:length(scenario_inputs))
end

# ---- `_evaluate_node` specialised on the simulation type ------------------------------------------
# Same scenario enumeration, marshalling and CPT write-back as evaluate_node.jl:41-122; only the
# scenario loop is replaced by one library call (a few, for the outer loop of CudaDoubleLoop).
const CudaSimulation = Union{CudaMonteCarlo,CudaDoubleLoop,CudaRandomSlicing}

function scenarios_of(net::EnhancedBayesianNetwork, node)
    ancestors = discrete_ancestors(net, node)
    combos = sort(vec(collect(Iterators.product(states.(ancestors)...))))
    owner(s) = (hits = filter(n -> s ∈ states(n), ancestors); isempty(hits) ? node.name : hits[1].name)
    evidences = isempty(combos[1]) ? [Evidence()] : map(c -> Evidence(owner(s) => s for s in c), combos)
    scen = map(ev -> mapreduce(p -> _uq_inputs(p, ev), vcat, parents(net, node)[3]; init=UQInput[]), evidences)
    return ancestors, evidences, scen
end

function EnhancedBayesianNetworks._evaluate_node(net::EnhancedBayesianNetwork,
    node::DiscreteFunctionalNode, collect_samples::Bool, sim::CudaSimulation)
    ancestors, evidences, scen = scenarios_of(net, node)
    cpt_names = isempty(ancestors) ? [node.name] : vcat([a.name for a in ancestors], node.name)
    # evaluate_node.jl:58-75: the table is imprecise iff an imprecise ContinuousNode parent is NOT discretised
    imprecise = filter(x -> isa(x, ContinuousNode) && !isprecise(x) && isempty(x.discretization.intervals),
        parents(net, node)[3])
    P = isempty(imprecise) ? PreciseDiscreteProbability : ImpreciseDiscreteProbability
    cpt = DiscreteConditionalProbabilityTable{P}(cpt_names)
    info = Dict{AbstractVector{Symbol},Dict}()
    fail_safe(ev) = begin
        f = copy(ev); f[node.name] = Symbol(string(node.name) * "_fail")
        s = copy(ev); s[node.name] = Symbol(string(node.name) * "_safe")
        f, s
    end
    if sim isa CudaMonteCarlo
        _, pf, var = pf_batch(collect(node.models), sim.performance, scen, sim)
        for (s, ev) in enumerate(evidences)
            fail_ev, safe_ev = fail_safe(ev)
            cpt[fail_ev] = pf[s]
            cpt[safe_ev] = 1 - pf[s]
            # the n-row DataFrame cannot be materialised at 1e8+ samples: empty frame, replay with
            # ebn_sample_matrix if rows are needed (INTEGRATION.md "samples")
            collect_samples && (info[collect(values(ev))] = Dict(:cov => sqrt(var[s]), :samples => DataFrame()))
        end
    elseif sim isa CudaDoubleLoop
        bounds = double_loop(collect(node.models), sim.lb.performance, scen, sim)
        for ((lb, ub), ev) in zip(bounds, evidences)
            fail_ev, safe_ev = fail_safe(ev)
            cpt[fail_ev] = (lb, ub)
            cpt[safe_ev] = (1 - ub, 1 - lb)
            collect_samples && (info[collect(values(ev))] = Dict())
        end
    else
        res = random_slicing(collect(node.models), sim.lb.performance, scen, sim)
        for (((lb, ub), sd_lb, sd_ub), ev) in zip(res, evidences)
            fail_ev, safe_ev = fail_safe(ev)
            cpt[fail_ev] = (lb, ub)
            cpt[safe_ev] = (1 - ub, 1 - lb)
            collect_samples && (info[collect(values(ev))] = Dict(:lb => (sd_lb, DataFrame()), :ub => (sd_ub, DataFrame())))
        end
    end
    return DiscreteNode(node.name, cpt, node.parameters, info)
end

# ---- ContinuousFunctionalNode: histogram instead of a KDE over n outputs ---------------------------

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
    all(isprecise.(parents(net, node)[3])) ||
        error("node $(node.name) is a continuousfunctionalnode with at least one parent with Interval or p-boxes in its distributions. No method for extracting failure probability p-box have been implemented yet")
    ancestors, evidences, scen = scenarios_of(net, node)
    if isempty(ancestors)
        cpt = ContinuousConditionalProbabilityTable{PreciseContinuousInput}()
        discretization = ExactDiscretization(node.discretization.intervals)
    else
        cpt = ContinuousConditionalProbabilityTable{PreciseContinuousInput}([i.name for i in ancestors])
        discretization = node.discretization
    end
    # the histogrammed quantity is the last model's column: catalogue functor, or an expression chain whose
    # "performance" is that column
    low = lower(collect(node.models), node.models[end] isa CudaExprModel ? node.models[end].name : nothing,
        input_names(scen[1]))
    jd = build_job(scen, low.columns, low.hidden; qmc=sim.qmc)
    pilot = CudaMonteCarlo(min(sim.n, 1 << 16), sim.seed, sim.strict, nothing, sim.fp32_sampling, sim.qmc)
    _, s1, s2 = run_hist(low.limit_state, low.consts, jd, pilot, [0.0])        # moments only
    μ = s1 ./ pilot.n
    σ = sqrt.(max.(s2 ./ pilot.n .- μ .^ 2, 0.0))
    lo, hi = minimum(μ .- 8 .* σ), maximum(μ .+ 8 .* σ)
    hi > lo || ((lo, hi) = (lo - 1.0, hi + 1.0))
    edges = collect(range(lo, hi; length=1025))
    counts, _, _ = run_hist(low.limit_state, low.consts, jd, sim, edges)       # ONE call, all scenarios
    info = Dict{Vector{Symbol},Dict}()
    for (s, ev) in enumerate(evidences)
        outside = (counts[s, 1] + counts[s, end]) / sim.n
        outside > 1e-3 && @warn "node $(node.name), scenario $(collect(values(ev))): $(round(100outside; digits=2)) % of the outputs fall outside the histogram grid (heavy tail or NaN); they are dropped from the distribution"
        cpt[ev] = histogram_distribution(edges, counts[s, 2:end-1])
        collect_samples && (info[collect(values(ev))] = Dict(:samples => DataFrame(), :mass_outside_grid => outside))
    end
    return ContinuousNode(node.name, cpt, discretization, info)
end

# dispatch hook: called from _evaluate_node(net, node, collect_samples) of the reference when
# `node.simulation isa EBNCuda.CudaSimulation` (one-line patch per method, see INTEGRATION.md)

end # module
