"""omc_b200: B200-native (sm_100a) bounding engine for OptimalMatrixCompletion.jl's branch-and-bound.

Import as ``import omc_b200`` (the repo-root shim registers this directory, whose name
``optimalmatrixcompletion.jl_b200`` is not a valid Python identifier, under that name).
"""
from . import _lib  # noqa: F401
from .engine import (Problem, Frontier, Cut, default_opts, init, build_flags, bitmatrix_chunks,  # noqa: F401
                     matrix_completion_SDP_relaxation, evaluate_objective, compute_MSE,
                     matrix_completion_master_feasible, smallest_eigvecs_batch, psd_project_batch,
                     measure_fp64_peak, alternating_minimization, alternating_minimization_batch, shor_constraint_indexes, generate_violated_Shor_minors, LABELS, MOI_STATUS)
from .host import (BBNode, BBTree, JuliaPriorityQueue, child_directions, create_matrix_cut_child_nodes,  # noqa: F401,E402
                   expand_frontier, shard_block_cyclic, matrix_completion_branchandbound)
