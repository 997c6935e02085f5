# bnb_batched.jl -- the body of the reference's main loop (OMC.jl:700-1073, /root/reference/src/OptimalMatrixCompletion.jl)
# restructured to pop `frontier_batch` nodes, relax them in ONE engine call, and consume the results in pop order.
#
# NOT EXECUTED IN THIS REPOSITORY: neither Julia nor Mosek exists in the build image (SURVEY.md G4/G5).  It is kept
# line-aligned with the Python mirror that IS executed by the tests and the benchmark
# (optimalmatrixcompletion.jl_b200/host.py:357-500; the "host.py:NNN" tags below name the mirrored statement), so a
# maintainer can diff the two.  `frontier_batch = 1` reproduces the reference's sequence of pops, rand() draws and
# incumbent updates exactly; larger batches evaluate nodes the reference might have pruned first (SURVEY.md section 7.7).
#
# How to use: inside `matrix_completion_branchandbound` (signature and outputs unchanged), keep everything up to the
# construction of `tree` (OMC.jl:140-698), create `problem = OMCB200.Problem(k, A, indices, γ, disjunctive_cuts_type)` once,
# replace the `while` loop OMC.jl:700-1073 by `bnb_batched_loop!(...)` below, and keep the epilogue OMC.jl:1075-1145.
# Only the disjunctive path is routed to the engine (`use_disjunctive_cuts = true`, every BASELINE configuration); the
# McCormick path keeps the reference's own loop.

using LinearAlgebra
using Random

"""
    bnb_batched_loop!(tree, solution, printlist, instance, counters, problem, k, A, indices, γ; kwargs...)

`counters` is the Dict of the reference's node counters (`nodes_dominated`, `nodes_relax_infeasible`, ... OMC.jl:410-420) and
time accumulators; every other argument is the reference's own object.  Engine knobs (defaults keep the reference's behaviour):
`frontier_batch` (nodes popped and relaxed per engine call), `use_cutoff` (hand the incumbent to the engine so that a node whose
CERTIFIED bound exceeds it stops early and is pruned at OMC.jl:797), `relax_opts`.
"""
function bnb_batched_loop!(tree, solution, printlist, instance, counters, problem::OMCB200.Problem,
                           k, A, indices, γ;
                           node_selection, bestfirst_depthfirst_cutoff, gap, disjunctive_cuts_type, disjunctive_cuts_breakpoints,
                           altmin_flag, max_altmin_probability, min_altmin_probability, altmin_probability_decay_rate,
                           use_max_steps, max_steps, time_limit, update_step, verbosity, start_time, root_only,
                           frontier_batch::Int = 1, use_cutoff::Bool = true, relax_opts = OMCB200.default_opts(),
                           # Shor valid inequalities (the reference's own keyword arguments, OMC.jl:152-159); `rng` as in the caller
                           add_Shor_valid_inequalities::Bool = false, add_Shor_valid_inequalities_iterative::Bool = false,
                           Shor_valid_inequalities_noisy_rank1_num_entries_present::Vector{Int} = [1, 2, 3, 4],
                           max_update_Shor_indices_probability = 1.0, min_update_Shor_indices_probability = 0.1,
                           update_Shor_indices_probability_decay_rate = 1.1, update_Shor_indices_n_minors::Int = 100)
    nev = disjunctive_cuts_breakpoints == "smallest_1_eigvec" ? 1 : 2
    # Shor rows: the engine keeps ONE row structure per problem (OMCB200.set_shor!).  Non-iterative mode: the root's BBNodeShorInfo,
    # attached once.  Iterative mode: nodes are relaxed in groups that share their Shor_info object (all children of a split do,
    # OMC.jl:2532-2539) and the structure is swapped between groups.                                                    host.py: "relax grouping"
    shor_loaded = nothing
    shor_candidates = NTuple{4, Int}[]
    if add_Shor_valid_inequalities
        if add_Shor_valid_inequalities_iterative     # the candidate list depends on the mask only: once per run (OMC.jl:2621-2624)
            shor_candidates = OMCB200.shor_constraint_indexes(problem, Shor_valid_inequalities_noisy_rank1_num_entries_present)[1]
        end
        root_info = tree.nodes[first(keys(tree.nodes))].Shor_info       # built by the caller exactly as OMC.jl:646-676
        OMCB200.set_shor!(problem, root_info.constraints_indexes, root_info.SOC_constraints_indexes)
        shor_loaded = root_info
    end
    relax_group(nds) = add_Shor_valid_inequalities ? OMCB200.relax_batch_shor(problem, nds; opts = relax_opts) :
                                                     OMCB200.relax_batch(problem, nds; opts = relax_opts)
    while (tree.now_gap > gap && !(use_max_steps && (tree.counter ≥ max_steps)) && time() - start_time ≤ time_limit)   # host.py:357
        length(tree.nodes) == 0 && break                                                                              # host.py:359
        # ---- pop up to frontier_batch nodes (OMC.jl:709-719)                                                          host.py:363-369
        batch = BBNode[]
        while length(batch) < frontier_batch && length(tree.nodes) > 0
            node_selection_here = node_selection
            if node_selection == "bestfirst_depthfirst"
                node_selection_here = length(tree.nodes) > bestfirst_depthfirst_cutoff ? "depthfirst" : "bestfirst"
            end
            push!(batch, retrieve_node_from_tree!(tree, node_selection_here))
        end
        live = [nd for nd in batch if !(nd.LB > tree.best_upper_bound)]                                                # host.py:370
        results = Dict{Int, Any}()
        if !isempty(live)                                                                                              # host.py:372-382
            relax_opts.cutoff = use_cutoff ? tree.best_upper_bound : Inf
            relax_opts.time_limit_s = max(1.0, time_limit - (time() - start_time))
            if add_Shor_valid_inequalities && add_Shor_valid_inequalities_iterative
                groups = Dict{UInt, Vector{Int}}()                         # nodes that share one Shor_info object
                for (q, nd) in enumerate(live)
                    push!(get!(groups, objectid(nd.Shor_info), Int[]), q)
                end
                for qs in values(groups)
                    info = live[qs[1]].Shor_info
                    if shor_loaded !== info
                        OMCB200.set_shor!(problem, info.constraints_indexes, info.SOC_constraints_indexes)
                        shor_loaded = info
                    end
                    for (q, r) in zip(qs, relax_group(live[qs]))
                        results[live[q].node_id] = r
                    end
                end
            else
                res = relax_group(live)                                     # ONE engine call: B x matrix_completion_SDP_relaxation
                for (nd, r) in zip(live, res)
                    results[nd.node_id] = r
                end
            end
        end
        # separation oracle for the whole batch in one launch (OMC.jl:814, 2466-2477)                                   host.py:391-397
        eig = Dict{Int, Any}()
        cand = [nd for nd in live if results[nd.node_id]["feasible"]]
        if !isempty(cand)
            (_, _, bp, feas) = OMCB200.smallest_eigvecs_batch([results[nd.node_id]["Y"] for nd in cand],
                                                              [results[nd.node_id]["U"] for nd in cand], nev)
            for (q, nd) in enumerate(cand)
                eig[nd.node_id] = (bp[:, q], feas[q])
            end
        end
        # ---- consume the results in pop order (OMC.jl:721-1073)                                                       host.py:399
        for (pos, current_node) in enumerate(batch)
            # nodes of this batch that were popped but not consumed yet left the queue early: their bounds still count
            pending = [nd.LB for nd in batch[pos+1:end] if !(nd.LB > tree.best_upper_bound)]                            # host.py:400
            pending_min = isempty(pending) ? Inf : minimum(pending)
            split_flag = true
            relax_result = nothing
            objective_relax = NaN
            if current_node.LB > tree.best_upper_bound                                                                 # OMC.jl:725-728
                split_flag = false
                counters["nodes_dominated"] += 1
            end
            if split_flag                                                                                              # OMC.jl:745-800
                relax_result = results[current_node.node_id]
                counters["solve_time_relaxation"] += relax_result["solve_time"]
                push!(counters["dict_solve_times_relaxation"], [current_node.node_id, current_node.depth, relax_result["solve_time"]])
                if current_node.node_id == 1
                    instance["run_details"]["root_node_timeout"] = (relax_result["termination_status"] == MOI.TIME_LIMIT)
                end
                if relax_result["feasible"] == false                                                                   # OMC.jl:777-779
                    counters["nodes_relax_infeasible"] += 1
                    split_flag = false
                else
                    counters["nodes_relax_feasible"] += 1
                    objective_relax = relax_result["objective"]
                    if relax_result["termination_status"] != MOI.OPTIMAL                                               # host.py:419-427
                        # SLOW_PROGRESS / TIME_LIMIT: the primal objective of an unconverged first-order iterate is no bound.
                        # The node keeps the larger of its inherited bound and the engine's CERTIFIED lower bound.
                        cert = min(objective_relax, relax_result["lower_bound"])
                        objective_relax = isfinite(current_node.LB) ? max(current_node.LB, cert) : cert
                    end
                    current_node.LB = objective_relax
                    if current_node.node_id == 1
                        tree.best_lower_bound = objective_relax
                    end
                    if objective_relax > tree.best_upper_bound                                                         # OMC.jl:797-800
                        counters["nodes_relax_feasible_pruned"] += 1
                        split_flag = false
                    end
                end
            end
            if split_flag && relax_result["termination_status"] == MOI.OPTIMAL                                         # OMC.jl:806-837
                if eig[current_node.node_id][2]                          # matrix_completion_master_feasible (OMC.jl:814)
                    current_node.master_feasible = true
                    counters["nodes_master_feasible"] += 1
                    if objective_relax < tree.best_upper_bound
                        counters["nodes_master_feasible_improvement"] += 1
                        update_solution!(solution, objective_relax, time() - start_time,
                                         relax_result["Y"], relax_result["U"], relax_result["X"])
                        tree.best_upper_bound = objective_relax
                        add_update!(printlist, instance, tree, time() - start_time; print_message = (verbosity ≥ 1))
                    end
                    split_flag = false
                end
            elseif split_flag && relax_result["termination_status"] == MOI.TIME_LIMIT                                  # OMC.jl:838-853
                add_update!(printlist, instance, tree, time() - start_time; print_message = (verbosity ≥ 1))
                verbosity ≥ 1 && add_message!(printlist, ["Time limit reached.\n"])
            end
            # alternating minimisation heuristic (OMC.jl:856-949): rand() is consumed once per processed node        host.py:439-461
            altmin_flag_now = false
            if altmin_flag
                lim = log(altmin_probability_decay_rate, max_altmin_probability / min_altmin_probability)
                altmin_probability = current_node.depth > lim ? min_altmin_probability :
                                     max_altmin_probability / (altmin_probability_decay_rate ^ current_node.depth)
                altmin_flag_now = (rand() < altmin_probability)
            end
            if split_flag && altmin_flag_now
                U_rounded = svd(relax_result["Y"]).U[:, 1:k]                                                           # OMC.jl:873
                am = OMCB200.alternating_minimization(problem; U_initial = Matrix(U_rounded),
                                                      disjunctive_cuts = current_node.disjunctive_cuts.cuts, time_limit = time_limit)
                alternating_minimization_printout(printlist, am, current_node.node_id, altmin_probability, verbosity)
                counters["nodes_relax_feasible_split_altmin"] += 1
                counters["solve_time_altmin"] += am["solve_time"]
                push!(counters["dict_solve_times_altmin"], [current_node.node_id, current_node.depth, am["solve_time"]])
                push!(counters["dict_num_iterations_altmin"], [current_node.node_id, current_node.depth, am["n_iters"]])
                if am["converged"]                                                                                     # OMC.jl:919
                    X_local = am["U"] * am["V"]
                    U_local = svd(X_local).U[:, 1:k]
                    Y_local = U_local * U_local'
                    objective_local = OMCB200.evaluate_objective(problem, X_local, A, indices, U_local, γ)
                    if objective_local < tree.best_upper_bound
                        counters["nodes_relax_feasible_split_altmin_improvement"] += 1
                        update_solution!(solution, objective_local, time() - start_time, Y_local, U_local, X_local)
                        tree.best_upper_bound = objective_local
                        add_update!(printlist, instance, tree, time() - start_time; altmin_flag = true, print_message = (verbosity ≥ 1))
                    end
                end
            end
            if split_flag                                                                                              # OMC.jl:951-989
                counters["nodes_relax_feasible_split"] += 1
                # create_matrix_cut_child_nodes (OMC.jl:2411-2543) with the breakpoint vector of the GPU separation oracle in
                # place of its `eigs` call (OMC.jl:2466-2477); child ordering and node ids are the reference's (OMC.jl:2479-2541)
                children = create_matrix_cut_child_nodes_with_breakpoint(
                    current_node, disjunctive_cuts_type, eig[current_node.node_id][1], relax_result["U"],
                    tree.counter, objective_relax)
                if add_Shor_valid_inequalities
                    info = current_node.Shor_info
                    if add_Shor_valid_inequalities_iterative                                                           # OMC.jl:956-969
                        lim = log(update_Shor_indices_probability_decay_rate, max_update_Shor_indices_probability / min_update_Shor_indices_probability)
                        update_probability = current_node.depth > lim ? min_update_Shor_indices_probability :
                                             max_update_Shor_indices_probability / (update_Shor_indices_probability_decay_rate ^ current_node.depth)
                        if rand() < update_probability                                                                 # OMC.jl:2495-2518
                            n, m = size(A)
                            minors = OMCB200.generate_violated_Shor_minors(problem, reshape(relax_result["X"], (1, n, m)), shor_candidates,
                                                                           info.constraints_indexes, update_Shor_indices_n_minors)
                            Shor_constraints_indexes = union(info.constraints_indexes, [mn[2] for mn in minors])
                            covered = unique(vcat([[(i1, j1), (i1, j2), (i2, j1), (i2, j2)] for (i1, i2, j1, j2) in Shor_constraints_indexes]...))
                            info = BBNodeShorInfo(constraints_indexes = Shor_constraints_indexes,
                                                  SOC_constraints_indexes = setdiff(info.SOC_constraints_indexes, covered))
                        end
                    end
                    for child in children
                        child.Shor_info = info
                    end
                end
                add_nodes_to_tree!(tree, children, objective_relax, current_node.node_id)
            end
            prune_dominated_nodes!(tree)                                                                               # OMC.jl:1036
            lower_bounds_updated = update_tree_lower_bounds_pending!(tree, pending_min)                                # OMC.jl:1039, host.py:211-222
            print_update_here = (lower_bounds_updated || current_node.node_id == 1
                                 || (tree.counter ÷ update_step) > (tree.last_updated_counter ÷ update_step)
                                 || tree.now_gap ≤ gap || (use_max_steps && tree.counter ≥ max_steps)
                                 || time() - start_time > time_limit) ? (verbosity ≥ 1) : (verbosity ≥ 3)
            add_update!(printlist, instance, tree, time() - start_time; print_message = print_update_here)
            root_only && break                                                                                         # OMC.jl:1070
        end
        root_only && break
    end
    return tree
end

"""`update_tree_lower_bounds!` (OMC.jl:1207-1218) that also counts the bounds of nodes popped in the current batch but not
consumed yet; `pending_min = Inf` (frontier_batch = 1) is the reference's rule."""
function update_tree_lower_bounds_pending!(tree, pending_min::Float64)
    (isempty(tree.lower_bounds) && pending_min == Inf) && return true
    minval = pending_min
    if !isempty(tree.lower_bounds)
        minval = min(minval, peek(tree.lower_bounds)[2])
    end
    if minval > tree.best_lower_bound
        tree.best_lower_bound = minval
        return true
    end
    return false
end

"""Host part of `create_matrix_cut_child_nodes` (OMC.jl:2479-2541) given the breakpoint vector: children share the parent's cut
list plus the new cut `(x, Û, directions)`; `node_id = counter + ind` with the first direction varying fastest."""
function create_matrix_cut_child_nodes_with_breakpoint(node, disjunctive_cuts_type, breakpoint_vec, U, counter, objective_relax)
    labels = OMCB200.LABELS[disjunctive_cuts_type]
    k = size(U, 2)
    children = BBNode[]
    for (ind, directions) in enumerate(Iterators.product(repeat([labels], k)...))                                      # OMC.jl:2481-2491
        push!(children, BBNode(
            node_id = counter + ind,
            parent_id = node.node_id,
            disjunctive_cuts = BBNodeDisjunctiveCuts(cuts = vcat(node.disjunctive_cuts.cuts, [(breakpoint_vec, U, collect(directions))])),
            LB = objective_relax,
            depth = node.depth + 1,
        ))
    end
    return children
end
