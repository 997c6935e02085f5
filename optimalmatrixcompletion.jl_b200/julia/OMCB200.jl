# OMCB200.jl -- Julia glue that routes OptimalMatrixCompletion.jl's hot path through libomc_b200.so.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: neither Julia nor Mosek exists in the build image (SURVEY.md G4/G5).
# The same calls are exercised from Python (optimalmatrixcompletion.jl_b200/engine.py + host.py), which
# mirrors this file function for function.  See INTEGRATION.md for how a maintainer wires it in.
#
# The reference keeps its public API and its host loop (node queue, branching, cut bookkeeping, incumbent,
# printlist / instance outputs).  Only the BODIES of the functions below change: each builds the flat ABI
# arguments, `ccall`s the engine and returns the same Dict keys the host loop reads
# (/root/reference/src/OptimalMatrixCompletion.jl, "OMC.jl" below).
module OMCB200

using LinearAlgebra
import MathOptInterface
const MOI = MathOptInterface

const LIB = get(ENV, "OMC_B200_LIB", joinpath(@__DIR__, "..", "libomc_b200.so"))

const STATUS_TO_MOI = Dict(
    0 => MOI.OPTIMAL,        # OMC_STATUS_OPTIMAL
    1 => MOI.SLOW_PROGRESS,  # OMC_STATUS_ITERATION_LIMIT (has values -> feasible = true, OMC.jl:1871-1877)
    2 => MOI.INFEASIBLE,     # OMC_STATUS_INFEASIBLE
    3 => MOI.TIME_LIMIT,     # OMC_STATUS_TIME_LIMIT
    4 => MOI.OPTIMAL,        # OMC_STATUS_CUTOFF: objective = certified bound > incumbent, pruned at OMC.jl:797
)
const CUT_TYPE = Dict("linear" => 0, "linear2" => 1, "linear3" => 2)
const LABELS = Dict(
    "linear" => ["left", "right"],                                 # OMC.jl:2482
    "linear2" => ["left", "middle", "right"],                      # OMC.jl:2486
    "linear3" => ["left", "inner_left", "inner_right", "right"],   # OMC.jl:2490
)

struct OmcError <: Exception
    code::Int32
    msg::String
end
function check(rc::Int32)
    rc == 0 && return
    throw(OmcError(rc, unsafe_string(ccall((:omc_last_error, LIB), Cstring, ()))))
end

# mirrors omc_relax_opts (include/omc_b200.h)
mutable struct RelaxOpts
    eps_abs::Float64; eps_rel::Float64
    max_iter::Int32; check_every::Int32; adapt_every::Int32; fix_linear3_right::Int32
    rho0::Float64; sigma::Float64; alpha::Float64; cutoff::Float64; time_limit_s::Float64; jacobi_tol::Float64
    reortho_every::Int32; exact_projection::Int32
end
function default_opts()
    o = RelaxOpts(0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0)
    ccall((:omc_relax_default_opts, LIB), Cvoid, (Ref{RelaxOpts},), o)
    return o
end

init(device::Integer = 0) = check(ccall((:omc_init, LIB), Int32, (Int32,), device))

"""(A, indices, γ, k) resident in HBM + cut pool.  One per `matrix_completion_branchandbound` call."""
mutable struct Problem
    handle::Ptr{Cvoid}
    n::Int; m::Int; k::Int
    cut_type::String
    cut_ids::IdDict{Any, Int32}     # (breakpoint_vec, Û) tuple identity -> pool id (children share the tuple, OMC.jl:2522)
    function Problem(k::Int, A::Matrix{Float64}, indices::BitMatrix, γ::Float64, cut_type::String; state_pool::Int = 0)
        (n, m) = size(A)
        h = Ref{Ptr{Cvoid}}(C_NULL)
        # BitMatrix.chunks is exactly the ABI's mask layout (column-major bit index, LSB first)
        check(ccall((:omc_problem_create, LIB), Int32,
                    (Int32, Int32, Int32, Ptr{Float64}, Ptr{UInt64}, Float64, Int32, Int32, Ref{Ptr{Cvoid}}),
                    n, m, k, A, indices.chunks, γ, CUT_TYPE[cut_type], state_pool, h))
        p = new(h[], n, m, k, cut_type, IdDict{Any, Int32}())
        finalizer(x -> ccall((:omc_problem_destroy, LIB), Int32, (Ptr{Cvoid},), x.handle), p)
        return p
    end
end

function cut_id!(p::Problem, breakpoint_vec::Vector{Float64}, Û::Matrix{Float64})
    key = breakpoint_vec                      # one Vector object per cut, shared by all children (OMC.jl:2522)
    haskey(p.cut_ids, key) && return p.cut_ids[key]
    v̂ = Û' * breakpoint_vec                   # the only use of Û (OMC.jl:1577, 2053)
    id = Ref{Int32}(0)
    check(ccall((:omc_cutpool_add, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ref{Int32}),
                p.handle, breakpoint_vec, v̂, id))
    p.cut_ids[key] = id[]
    return id[]
end

function flatten(p::Problem, nodes_cuts)
    B = length(nodes_cuts)
    ptr = zeros(Int32, B + 1); ids = Int32[]; dirs = UInt8[]
    for (b, cuts) in enumerate(nodes_cuts)
        ptr[b + 1] = ptr[b] + length(cuts)
        for (x, Û, directions) in cuts
            push!(ids, cut_id!(p, x, Û))
            append!(dirs, UInt8[findfirst(isequal(d), LABELS[p.cut_type]) - 1 for d in directions])
        end
    end
    isempty(ids) && (push!(ids, 0); push!(dirs, 0))
    return ptr, ids, dirs
end

"""
Batched body of `matrix_completion_SDP_relaxation` (OMC.jl:1431-1943).  `nodes::Vector{BBNode}`; returns one Dict
per node with the reference's keys ("model", "solve_time", "termination_status", "feasible", "objective",
"Y", "U", "X", "Θ" -- OMC.jl:1860-1919).  "Θ" is returned as `nothing`: on the disjunctive path the host only
reads it through the objective, which the engine already recomputed from the primal point (OMC.jl:1890-1895).
"""
function relax_batch(p::Problem, nodes; opts::RelaxOpts = default_opts())
    B = length(nodes)
    ptr, ids, dirs = flatten(p, [nd.disjunctive_cuts.cuts for nd in nodes])
    status = zeros(Int32, B); iters = zeros(Int32, B)
    objective = zeros(B); lower = zeros(B); res = zeros(2B)
    X = zeros(p.n, p.m, B); Y = zeros(p.n, p.n, B); U = zeros(p.n, p.k, B)
    ms = Ref{Float32}(0)
    t = @elapsed check(ccall((:omc_relax_batch, LIB), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Ref{RelaxOpts},
         Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ref{Float32}),
        p.handle, B, ptr, ids, dirs, C_NULL, C_NULL, opts, status, objective, lower, iters, res, X, Y, U, C_NULL, ms))
    for b in 1:B   # OMC_STATUS_NUMERICAL: same failure mode as the reference's final else branch (OMC.jl:1936-1940)
        status[b] == 5 && error("unexpected termination status: NUMERICAL_ERROR (node $b of the batch)")
    end
    return [Dict{String, Any}(
        "model" => nothing,
        "solve_time" => t / B,
        "termination_status" => STATUS_TO_MOI[Int(status[b])],
        "feasible" => status[b] != 2,
        "objective" => objective[b],
        "Y" => Y[:, :, b], "U" => U[:, :, b], "X" => X[:, :, b], "Θ" => nothing,
        "lower_bound" => lower[b], "iterations" => iters[b],
    ) for b in 1:B]
end

"""Drop-in for `matrix_completion_SDP_relaxation(node, n, k, A, indices, γ, true; disjunctive_cuts_type, ...)`."""
matrix_completion_SDP_relaxation(p::Problem, node; kwargs...) = relax_batch(p, [node]; kwargs...)[1]

"""Batched `eigs(Symmetric(U*U' - Y), nev, which=:SR)` + the test of OMC.jl:1272-1277."""
function smallest_eigvecs_batch(Ys::Vector{Matrix{Float64}}, Us::Vector{Matrix{Float64}}, nev::Int)
    B = length(Ys); (n, k) = size(Us[1])
    Y = cat(Ys...; dims = 3); U = cat(Us...; dims = 3)
    lam = zeros(nev, B); vec = zeros(n, nev, B); bp = zeros(n, B); feas = zeros(Int32, B)
    check(ccall((:omc_smallest_eigvecs_batch, LIB), Int32,
                (Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}),
                n, k, B, Y, U, nev, lam, vec, bp, feas))
    return lam, vec, bp, feas .== 1
end
matrix_completion_master_feasible(Y, U, X, Θ, use_disjunctive_cuts::Bool) =
    smallest_eigvecs_batch([Y], [U], 1)[4][1]                                  # OMC.jl:1261-1277
breakpoint_vector(Y, U, nev::Int) = smallest_eigvecs_batch([Y], [U], nev)[3][:, 1]   # OMC.jl:2466-2477

"""Drop-in body of `alternating_minimization` (OMC.jl:1979-2279); same result Dict (OMC.jl:2249-2278)."""
function alternating_minimization(p::Problem; U_initial::Matrix{Float64}, disjunctive_cuts = [], ϵ::Float64 = 1e-5,
                                  max_iters::Int = 100, time_limit::Int = 3600)
    ids = Int32[cut_id!(p, x, Û) for (x, Û, _) in disjunctive_cuts]
    dirs = UInt8[findfirst(isequal(d), LABELS[p.cut_type]) - 1 for (_, _, ds) in disjunctive_cuts for d in ds]
    isempty(ids) && (push!(ids, 0); push!(dirs, 0))
    U = zeros(p.n, p.k); V = zeros(p.k, p.m); objectives = zeros(max_iters)
    conv = Ref{Int32}(0); nit = Ref{Int32}(0); st = Ref{Float64}(0)
    check(ccall((:omc_altmin, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Ptr{Int32}, Ptr{UInt8}, Float64, Int32, Float64, Ptr{Float64}, Ptr{Float64},
                 Ref{Int32}, Ref{Int32}, Ptr{Float64}, Ref{Float64}),
                p.handle, U_initial, length(disjunctive_cuts), ids, dirs, ϵ, max_iters, Float64(time_limit), U, V, conv, nit,
                objectives, st))
    return Dict("converged" => conv[] == 1, "U" => U, "V" => V, "solve_time" => st[], "n_iters" => Int(nit[]),
                "max_iters" => max_iters, "objectives" => objectives[1:nit[]])
end

"""`alternating_minimization` for B instances of one problem in ONE launch (`omc_altmin_batch`, one CTA per instance): the
extra root restarts `U_initial + maximum(abs.(U_initial)) * randn(n, k)` (OMC.jl:529-538) or the alt-min calls of a popped
batch of nodes (OMC.jl:867-927).  Returns one reference-style Dict per instance."""
function alternating_minimization_batch(p::Problem, U_initials::Vector{Matrix{Float64}}, cut_lists::Vector;
                                        ϵ::Float64 = 1e-5, max_iters::Int = 100, time_limit::Real = 3600)
    B = length(U_initials)
    ptr = zeros(Int32, B + 1); ids = Int32[]; dirs = UInt8[]
    for (b, cuts) in enumerate(cut_lists)
        for (x, Û, directions) in cuts
            push!(ids, cut_id!(p, x, Û)); append!(dirs, UInt8[findfirst(isequal(d), LABELS[p.cut_type]) - 1 for d in directions])
        end
        ptr[b + 1] = length(ids)
    end
    isempty(ids) && (push!(ids, 0); push!(dirs, 0))
    Uin = cat(U_initials...; dims = 3)                       # n x k x B, column-major per instance
    U = zeros(p.n, p.k, B); V = zeros(p.k, p.m, B); objectives = zeros(max_iters, B)
    conv = zeros(Int32, B); nit = zeros(Int32, B); st = Ref{Float64}(0)
    check(ccall((:omc_altmin_batch, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}, Float64, Int32, Float64, Ptr{Float64},
                 Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ref{Float64}),
                p.handle, B, Uin, ptr, ids, dirs, ϵ, max_iters, Float64(time_limit), U, V, conv, nit, objectives, st))
    return [Dict("converged" => conv[b] == 1, "U" => U[:, :, b], "V" => V[:, :, b], "solve_time" => st[],
                 "n_iters" => Int(nit[b]), "max_iters" => max_iters, "objectives" => objectives[1:nit[b], b]) for b in 1:B]
end

"""`generate_rank1_matrix_completion_Shor_constraints_indexes` (OMC.jl:2545-2612) and the SOC coordinate list of OMC.jl:656-665
on the GPU; returns 1-based `Vector{NTuple{4,Int}}` and `Vector{Tuple{Int,Int}}` exactly as the reference builds them."""
function shor_constraint_indexes(p::Problem, num_entries_present_list::Vector{Int})
    pl = Int32.(num_entries_present_list)
    cnt = Ref{Int64}(0); nsoc = Ref{Int64}(0)
    check(ccall((:omc_shor_indexes, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ref{Int64}, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Int64}),
                p.handle, pl, length(pl), cnt, C_NULL, 0, C_NULL, C_NULL))
    tuples = zeros(Int32, 4, max(1, cnt[])); soc = zeros(Int32, 2, p.n * p.m)
    check(ccall((:omc_shor_indexes, LIB), Int32, (Ptr{Cvoid}, Ptr{Int32}, Int32, Ref{Int64}, Ptr{Int32}, Int64, Ptr{Int32}, Ref{Int64}),
                p.handle, pl, length(pl), cnt, tuples, cnt[], soc, nsoc))
    return [ntuple(q -> Int(tuples[q, t]) + 1, 4) for t in 1:cnt[]], [(Int(soc[1, t]) + 1, Int(soc[2, t]) + 1) for t in 1:nsoc[]]
end

"""`add_Shor_valid_inequalities = true` (non-iterative mode): attaches the root's `BBNodeShorInfo` (OMC.jl:646-669; 1-based tuples
as the reference holds them) to the problem, so that every `relax_batch` of it carries the Shor rows of OMC.jl:1755-1828.
Empty lists remove the rows.  (UNTESTED: no Julia in the build image; the Python mirror `engine.py:Problem.set_shor` is.)"""
function set_shor!(p::Problem, constraints_indexes::Vector{NTuple{4, Int}}, SOC_constraints_indexes::Vector{Tuple{Int, Int}})
    mn = Int32[t[q] - 1 for q in 1:4, t in constraints_indexes]            # 4 x N, 0-based, column = one minor
    sc = Int32[t[q] - 1 for q in 1:2, t in SOC_constraints_indexes]
    check(ccall((:omc_problem_set_shor, LIB), Int32, (Ptr{Cvoid}, Int64, Ptr{Int32}, Int64, Ptr{Int32}),
                p.handle, size(mn, 2), mn, size(sc, 2), sc))
end

"""Relaxation of a batch of nodes WITH the Shor rows: `relax_batch` plus the result keys "W" and "Xt" (OMC.jl:1902, 1913)."""
function relax_batch_shor(p::Problem, nodes; opts::RelaxOpts = default_opts())
    B = length(nodes)
    ptr, ids, dirs = flatten(p, [nd.disjunctive_cuts.cuts for nd in nodes])
    fr = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:omc_frontier_create_ex, LIB), Int32,
                (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{UInt8}, Ptr{Int32}, Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}),
                p.handle, B, ptr, ids, dirs, C_NULL, C_NULL, 0, fr))
    status = zeros(Int32, B); iters = zeros(Int32, B); objective = zeros(B); lower = zeros(B); res = zeros(2B)
    X = zeros(p.n, p.m, B); Y = zeros(p.n, p.n, B); U = zeros(p.n, p.k, B)
    W = zeros(p.n, p.m, B); Xt = zeros(p.n, p.m, p.k, B)
    ms = Ref{Float32}(0)
    try
        t = @elapsed check(ccall((:omc_frontier_relax, LIB), Int32, (Ptr{Cvoid}, Ref{RelaxOpts}, Ref{Float32}), fr[], opts, ms))
        check(ccall((:omc_frontier_fetch, LIB), Int32,
                    (Ptr{Cvoid}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                    fr[], status, objective, lower, iters, res, X, Y, U, C_NULL))
        check(ccall((:omc_frontier_fetch_shor, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), fr[], W, Xt))
        return [Dict{String, Any}(
            "model" => nothing, "solve_time" => t / B, "termination_status" => STATUS_TO_MOI[Int(status[b])],
            "feasible" => status[b] != 2, "objective" => objective[b],
            "Y" => Y[:, :, b], "U" => U[:, :, b], "X" => X[:, :, b], "Θ" => nothing, "W" => W[:, :, b],
            "Xt" => Xt[:, :, :, b], "iterations" => iters[b],
        ) for b in 1:B]
    finally
        ccall((:omc_frontier_destroy, LIB), Int32, (Ptr{Cvoid},), fr[])
    end
end

"""Drop-in for `generate_violated_Shor_minors(X, indices, pattern, Shor_constraints_indexes, n_minors)` (OMC.jl:2614-2640): same
return value, a Vector of `(score, (i1, i2, j1, j2))` (1-based) in the reference's order.  `candidates` is the result of
`shor_constraint_indexes(p, pattern)[1]`, which depends on the mask only: compute it once per run.  (UNTESTED in Julia.)"""
function generate_violated_Shor_minors(p::Problem, X::Array{Float64, 3}, candidates::Vector{NTuple{4, Int}},
                                       Shor_constraints_indexes::Vector{NTuple{4, Int}}, n_minors::Int)
    Xs = permutedims(X, (2, 3, 1))                                          # (k, n, m) -> k column-major n x m slices
    cand = Int32[t[q] - 1 for q in 1:4, t in candidates]
    excl = Int32[t[q] - 1 for q in 1:4, t in Shor_constraints_indexes]
    cap = max(1, min(n_minors, size(cand, 2)))
    tuples = zeros(Int32, 4, cap); scores = zeros(cap); cnt = Ref{Int64}(0)
    check(ccall((:omc_shor_score_minors, LIB), Int32,
                (Ptr{Cvoid}, Ptr{Float64}, Int32, Int64, Ptr{Int32}, Int64, Ptr{Int32}, Int64, Ref{Int64}, Ptr{Int32}, Ptr{Float64}),
                p.handle, Xs, size(X, 1), size(cand, 2), cand, size(excl, 2), excl, n_minors, cnt, tuples, scores))
    return [(scores[q], ntuple(e -> Int(tuples[e, q]) + 1, 4)) for q in 1:cnt[]]
end

# ---- multi-GPU exchange inside the library (INTEGRATION.md section 4): one Julia process per GPU -----------------------------
"""Rank 0 creates the 128-byte id; ship it to the other workers by any channel, then every rank calls `comm_init`."""
function comm_unique_id()
    id = zeros(UInt8, 128)
    check(ccall((:omc_comm_unique_id, LIB), Int32, (Ptr{UInt8},), id))
    return id
end
comm_init(id::Vector{UInt8}, rank::Integer, world::Integer) =
    check(ccall((:omc_comm_init, LIB), Int32, (Int32, Int32, Ptr{UInt8}), rank, world, id))
comm_destroy() = check(ccall((:omc_comm_destroy, LIB), Int32, ()))
"""In-place all-reduce-min over the ranks, e.g. of `[incumbent, min lower bound]` after a batch."""
function allreduce_min!(v::Vector{Float64})
    check(ccall((:omc_allreduce_min, LIB), Int32, (Ptr{Float64}, Int32), v, length(v)))
    return v
end
"""All-gather of `cnt` doubles per rank (per-node iteration counts for `balanced_partition`)."""
function allgather(v::Vector{Float64}, world::Integer)
    out = zeros(length(v) * world)
    check(ccall((:omc_allgather, LIB), Int32, (Ptr{Float64}, Int32, Ptr{Float64}), v, length(v), out))
    return reshape(out, length(v), world)
end

"""Fused `evaluate_objective` (OMC.jl:2330-2359) + `compute_MSE` (OMC.jl:2373-2409): (objective, in, out, all)."""
function objective_mse(p::Problem, X::Matrix{Float64})
    out = zeros(4)
    check(ccall((:omc_objective_mse, LIB), Int32, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), p.handle, X, out))
    return out
end
evaluate_objective(p::Problem, X, A, indices, U, γ) = objective_mse(p, X)[1]
compute_MSE(p::Problem, X, A, indices; kind = "out") = objective_mse(p, X)[kind == "in" ? 2 : kind == "out" ? 3 : 4]

end # module
