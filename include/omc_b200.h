/*
 * omc_b200.h -- C ABI of libomc_b200.so, the B200 (sm_100a) bounding engine for
 * sean-lo/OptimalMatrixCompletion.jl's disjunctive branch-and-bound.
 *
 * The reference has no FFI: its seam is the set of Julia functions the main loop calls
 * (OMC.jl = /root/reference/src/OptimalMatrixCompletion.jl).  Each entry point below names the
 * reference function whose BODY it replaces; the Julia glue (optimalmatrixcompletion.jl_b200/julia/)
 * keeps the signatures and result-Dict keys and ccall's these.
 *
 * Conventions:
 *   - every call returns int32: 0 = OK, < 0 = error (message via omc_last_error()); nothing throws
 *     across the ABI;
 *   - the caller owns every host buffer; the library copies in/out and keeps no host pointer after
 *     return; handles are opaque and freed by the matching destroy;
 *   - matrices are column-major Float64 exactly as Julia stores Matrix{Float64};
 *     the mask is Julia's BitMatrix.chunks (UInt64 words, column-major bit index, LSB first);
 *   - indices inside the ABI are 0-based;
 *   - calls are blocking and made from one host thread per process (the reference is single
 *     threaded); one process drives one GPU (omc_init(device)), N GPUs = N processes.
 *   - There is NO CPU fallback: every compute entry fails with OMC_ERR_CUDA when no sm_100 device
 *     is usable.
 */
#ifndef OMC_B200_H
#define OMC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OMC_OK 0
#define OMC_ERR_ARG (-1)
#define OMC_ERR_CUDA (-2)
#define OMC_ERR_STATE (-3)
#define OMC_ERR_UNSUPPORTED (-4)

/* per-node solver status; the Julia glue maps them to the MOI codes the host loop branches on
 * (OMC.jl:780-785, 809-812, 840-842, 1866-1940) */
#define OMC_STATUS_OPTIMAL 0         /* MOI.OPTIMAL                                              */
#define OMC_STATUS_ITERATION_LIMIT 1 /* MOI.SLOW_PROGRESS with values -> feasible = true          */
#define OMC_STATUS_INFEASIBLE 2      /* MOI.INFEASIBLE                    -> feasible = false      */
#define OMC_STATUS_TIME_LIMIT 3      /* MOI.TIME_LIMIT with values                                */
#define OMC_STATUS_CUTOFF 4          /* certified lower bound > opts.cutoff: objective := that bound,
                                        reported as MOI.OPTIMAL; the host prunes it at OMC.jl:797  */
#define OMC_STATUS_NUMERICAL 5       /* a non-finite iterate was met: the glue raises "unexpected termination
                                        status" like the reference's final else branch (OMC.jl:1936-1940)  */

/* disjunctive_cuts_type (OMC.jl:1581,1603,1635) */
#define OMC_CUT_LINEAR 0
#define OMC_CUT_LINEAR2 1
#define OMC_CUT_LINEAR3 2

typedef struct omc_problem omc_problem;   /* (A, indices, gamma, k) resident in HBM + cut pool + state pool */
typedef struct omc_frontier omc_frontier; /* a batch of open nodes resident in HBM                          */

typedef struct omc_relax_opts {
  double eps_abs;          /* ADMM residual tolerances (scaled program)                    */
  double eps_rel;
  int32_t max_iter;
  int32_t check_every;     /* residual / termination check period                          */
  int32_t adapt_every;     /* rho re-balancing period (multiple of check_every)            */
  int32_t fix_linear3_right; /* 0 = replicate OMC.jl:1675 (reference quirk Q1), 1 = valid secant */
  double rho0;
  double sigma;
  double alpha;            /* over-relaxation                                              */
  double cutoff;           /* incumbent upper bound; +inf disables early pruning           */
  double time_limit_s;     /* per call; <= 0 disables                                      */
  double jacobi_tol;       /* cap of the eigensolver's relative off-diagonal tolerance (it tightens with the ADMM residual) */
  int32_t reortho_every;   /* restart the eigenvector basis from I every this many iterations (0 = never) */
  int32_t exact_projection; /* 0 (default): warm-started low-rank tracking of the minority spectral side, every termination
                               decision confirmed by an exact projection; 1: full eigendecomposition at every iteration */
} omc_relax_opts;

/* ---- library / device ------------------------------------------------------------------------ */
int32_t omc_init(int32_t device);
int32_t omc_shutdown(void);
const char* omc_last_error(void);
int32_t omc_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, int64_t* free_bytes);
/* compile-time options of this build; callable before omc_init().  Bit 0: the relaxation kernel tests nodes with cuts
 * for a primal infeasibility certificate and can return OMC_STATUS_INFEASIBLE (make EXTRA=-DOMC_INFEASIBILITY_CERTIFICATE;
 * off by default: it costs ~10 % of the kernel, and inside branch-and-bound infeasible children already end through
 * `cutoff`).  The reference has no counterpart (Mosek reports MOI.INFEASIBLE by itself, OMC.jl:1921-1935). */
#define OMC_BUILD_INFEASIBILITY_CERTIFICATE 1
#define OMC_BUILD_INFEASIBLE_BY_BOUND 2       /* default build: INFEASIBLE when the certified bound exceeds 1/2 ||P(A)||^2; -DOMC_NO_INFEASIBLE_BY_BOUND removes it */
int32_t omc_build_flags(void);
/* stream the library launches on (a cudaStream_t), so a host can bracket launches with its own events */
void* omc_stream(void);

/* ---- problem: replaces the per-call (A, indices, gamma) arguments of every function below ---- */
/* mask_chunks: BitMatrix.chunks, ceil(n*m/64) words.  state_pool_capacity: number of warm-start
 * records (ADMM state of a relaxed node, reused by its children) the problem can hold.            */
int32_t omc_problem_create(int32_t n, int32_t m, int32_t k, const double* A, const uint64_t* mask_chunks,
                           double gamma, int32_t cut_type, int32_t state_pool_capacity, omc_problem** out);
int32_t omc_problem_destroy(omc_problem* p);
/* mask compaction (replaces the indices[i,j] loops OMC.jl:2200,2220,1852,2354).  Pass NULL arrays to
 * query nnz only.  rowptr[n+1]/colidx[nnz] = row-CSR, colptr[m+1]/rowidx[nnz] = column-CSC, ascending. */
int32_t omc_problem_get_csr(omc_problem* p, int32_t* rowptr, int32_t* colidx, int32_t* colptr, int32_t* rowidx,
                            int64_t* nnz);

/* ---- cut pool: one entry per (breakpoint_vec x, Uhat) created at OMC.jl:2522; only vhat = Uhat'x
 * is ever read (OMC.jl:1577, 2053).  A node is a list of (cut id, direction codes[k]).              */
int32_t omc_cutpool_add(omc_problem* p, const double* x, const double* vhat, int32_t* cut_id);
int32_t omc_cutpool_size(omc_problem* p, int32_t* size);

/* ---- relaxation: replaces the body of matrix_completion_SDP_relaxation (OMC.jl:1431-1943) and
 * compute_SDP_relaxation_objective (OMC.jl:1945-1977) for B nodes at once.                         */
void omc_relax_default_opts(omc_relax_opts* o);
/* node b owns cut entries node_cut_ptr[b] .. node_cut_ptr[b+1]-1; entry e is pool cut node_cut_ids[e]
 * with direction codes node_cut_dirs[e*k + j] (index into ["left","right"] / ["left","middle","right"] /
 * ["left","inner_left","inner_right","right"], OMC.jl:2481-2491).
 * warm_ids[b]  : state-pool record to start from (-1 = cold start).  May be NULL.
 * save_ids[b]  : state-pool record that receives node b's final ADMM state (-1 = none).  May be NULL. */
int32_t omc_frontier_create(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                            const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                            omc_frontier** out);
/* Same with an explicit engine.  OMC_ENGINE_PERSISTENT: one CTA per node, the node's whole ADMM inside one SM (n + m <= 104
 * in shared memory, <= 208 through an L2 buffer; at most 64 / 32 cuts per node).  OMC_ENGINE_BATCHED: the frontier advances in
 * lockstep, every ADMM iteration a short sequence of kernels over (tile, node) grids with the node state streaming from
 * HBM; PSD blocks of any size (OMC.jl:1554-1556 at config 4 / 5 sizes), no cut cap other than shared memory (~100 cuts),
 * cold starts only.  OMC_ENGINE_AUTO = batched when n + m > 104. */
#define OMC_ENGINE_AUTO 0
#define OMC_ENGINE_PERSISTENT 1
#define OMC_ENGINE_BATCHED 2
int32_t omc_frontier_create_ex(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                               const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                               int32_t engine, omc_frontier** out);
/* one pass of the hot path over the resident batch; kernel_ms (may be NULL) = CUDA-event time of the
 * fused ADMM kernel on the library stream */
int32_t omc_frontier_relax(omc_frontier* f, const omc_relax_opts* opts, float* kernel_ms);
/* results: any pointer may be NULL.  X[B*n*m], Y[B*n*n], U[B*n*k], Theta[B*m*m] column-major per node.
 * res[2*b] / res[2*b+1] = primal / dual residual.                                                       */
int32_t omc_frontier_fetch(omc_frontier* f, int32_t* status, double* objective, double* lower_bound,
                           int32_t* iters, double* res, double* X, double* Y, double* U, double* Theta);
/* per-node kernel profile: prof[32*b + q], q = 0..5 SM cycles spent in (w-update + dense rows, assembling V,
 * DMMA pre-rotation or low-rank tracking step, Jacobi sweeps of the full solver, reconstruction + dual update,
 * residual checks), 6 = Jacobi sweeps, 7 = iterations, 8..13 = full-solver rotation statistics,
 * 14 / 15 = projections done by the low-rank tracker / by the full solver,
 * 16..23 = cycles in the sub-phases of the tracking step (V Z, residual, CholQR2, V R~, Gram, Jacobi, select + combine) */
int32_t omc_frontier_fetch_profile(omc_frontier* f, double* prof);
/* generate_violated_Shor_minors (OMC.jl:2614-2640; call site 2495-2502): scores every candidate minor of `cand` (n_cand x 4, 0-based,
 * e.g. the list omc_shor_indexes returns -- it depends on the mask only, so the glue computes it once per run) that is not in
 * `excl` (the node's Shor_info.constraints_indexes) by sum_t |Xt[i1,j1] Xt[i2,j2] - Xt[i1,j2] Xt[i2,j1]| on the GPU and returns
 * the n_minors largest, ordered by (score, tuple) descending like the reference's sort(...; rev = true).  Xt: n_slices
 * column-major n x m slices (omc_frontier_fetch_shor's layout; the reference's call site passes X itself as ONE slice).  *count <= n_minors entries are written to tuples[4*q..] / scores[q].   */
int32_t omc_shor_score_minors(omc_problem* p, const double* Xt, int32_t n_slices, int64_t n_cand, const int32_t* cand, int64_t n_excl,
                              const int32_t* excl, int64_t n_minors, int64_t* count, int32_t* tuples, double* scores);
/* Shor valid inequalities (OMC.jl:1503-1552 variables, 1755-1828 rows, 1838-1846 objective; chosen by
 * generate_rank1_basis_pursuit_Shor_constraints_indexes, OMC.jl:568-668): attaches the rows to every node relaxation of the
 * problem.  minors[4*q..] = (i1, i2, j1, j2), 0-based, i1 < i2, j1 < j2 (the reference's 4-tuples minus one);
 * soc[2*q..] = (i, j) coordinates on a rotated second-order cone row, none of them covered by a minor (OMC.jl:656-665).
 * Frontiers of such a problem run on the batched engine; their status is OPTIMAL or ITERATION_LIMIT (no certified bound is
 * derived with these rows: lower_bound = -1e300, no cut-off).  k <= 4.  n_minors = n_soc = 0 removes the rows.            */
int32_t omc_problem_set_shor(omc_problem* p, int64_t n_minors, const int32_t* minors, int64_t n_soc, const int32_t* soc);
/* Shor results of a relaxed frontier: W[B*n*m] column-major per node (result key "W", OMC.jl:1902), Xt[B*k*n*m] (k slices
 * per node, result key "Xt", OMC.jl:1913; consumed by generate_violated_Shor_minors, OMC.jl:1099).  Either may be NULL.  */
int32_t omc_frontier_fetch_shor(omc_frontier* f, double* W, double* Xt);
/* out8: [0] engine, [1] kernel launches of the last relax, [2] lockstep iterations, [3] node-iterations, [4] residual checks,
 * [5] rho changes, [6] bytes of device state per node (batched engine) */
int32_t omc_frontier_stats(omc_frontier* f, int64_t* out8);
/* tracker knobs of the batched engine (values <= 0 keep the default): block-LOBPCG steps per projection (max / at the first
 * iteration and when confirming a termination decision), Ritz-residual tolerances while iterating / when confirming */
int32_t omc_frontier_set_tuning(omc_frontier* f, int32_t steps_max, int32_t steps_start, double track_tol, double confirm_tol);
/* diagnostics (tests): one array of a node's record of the batched engine after a relax (scaled variables, row-major).
 * which: 0..2 V_b, 3..5 Z_b (N_b x 16), 6..8 theta_b (16), 9..11 R_b, 12..14 W_b, 15 X, 16 Y, 17 Theta, 18 U.
 * Returns the number of doubles written, < 0 on error. */
int64_t omc_frontier_debug_fetch(omc_frontier* f, int32_t node, int32_t which, double* out, int64_t cap);
int32_t omc_frontier_destroy(omc_frontier* f);
/* convenience = create + relax + fetch + destroy (host buffers in, host buffers out) */
int32_t omc_relax_batch(omc_problem* p, int32_t B, const int32_t* node_cut_ptr, const int32_t* node_cut_ids,
                        const uint8_t* node_cut_dirs, const int32_t* warm_ids, const int32_t* save_ids,
                        const omc_relax_opts* opts, int32_t* status, double* objective, double* lower_bound,
                        int32_t* iters, double* res, double* X, double* Y, double* U, double* Theta,
                        float* kernel_ms);

/* ---- separation oracle: replaces eigs(Symmetric(U*U' - Y), nev, :SR) at OMC.jl:2466-2477 and the
 * test at OMC.jl:1272-1277.  Y[B*n*n], U[B*n*k] column-major.  Outputs: lam[B*nev] ascending,
 * vec[B*n*nev] (unit, largest-|.| component positive), breakpoint[B*n] (the mixed vector of
 * OMC.jl:2471-2476 when nev == 2), master_feasible[B] (lam_min >= -1e-6).                           */
int32_t omc_smallest_eigvecs_batch(int32_t n, int32_t k, int32_t B, const double* Y, const double* U, int32_t nev,
                                   double* lam, double* vec, double* breakpoint, int32_t* master_feasible);

/* ---- alternating minimisation: replaces the body of alternating_minimization (OMC.jl:1979-2279).
 * cut entries as for a node (ncuts may be 0).  objectives[max_iters].                               */
int32_t omc_altmin(omc_problem* p, const double* U_initial, int32_t ncuts, const int32_t* cut_ids,
                   const uint8_t* cut_dirs, double eps, int32_t max_iters, double time_limit_s, double* U,
                   double* V, int32_t* converged, int32_t* n_iters, double* objectives, double* solve_time);
/* The same heuristic for B instances of one problem in ONE launch (one CTA per instance): the reference's extra root
 * restarts U_initial + max|U_initial| randn (OMC.jl:529-538) and the alt-min calls of a popped batch of nodes
 * (OMC.jl:867-927).  U_initial / U: [B][n*k] column-major; V: [B][k*m]; objectives: [B][max_iters]; converged, n_iters: [B];
 * instance b applies the cuts cut_ptr[b] .. cut_ptr[b+1] of (cut_ids, cut_dirs).  solve_time is the whole batch. */
int32_t omc_altmin_batch(omc_problem* p, int32_t B, const double* U_initial, const int32_t* cut_ptr, const int32_t* cut_ids,
                         const uint8_t* cut_dirs, double eps, int32_t max_iters, double time_limit_s, double* U, double* V,
                         int32_t* converged, int32_t* n_iters, double* objectives, double* solve_time);

/* Shor 2x2-minor index enumeration, replaces generate_rank1_matrix_completion_Shor_constraints_indexes (OMC.jl:2545-2612,
 * called at OMC.jl:648-651) and the SOC coordinate list derived from it (OMC.jl:656-665, for fraction = 1.0).  present_list:
 * the reference's Shor_valid_inequalities_noisy_rank1_num_entries_present (values in 0..4, processed in the order given).
 * *count receives the number of minors; with tuples == NULL (and soc == nsoc == NULL) the call only counts.  tuples:
 * [count][4] = (i1, i2, j1, j2), 0-based, in exactly the order the reference pushes them; soc: [*nsoc][2] = (i, j), 0-based,
 * the coordinates covered by no minor in column-major order (i fastest).  Bit-exact integer work. */
int32_t omc_shor_indexes(omc_problem* p, const int32_t* present_list, int32_t nlist, int64_t* count, int32_t* tuples,
                         int64_t cap, int32_t* soc, int64_t* nsoc);
/* ---- fused objective + MSE: replaces evaluate_objective (OMC.jl:2330-2359) and compute_MSE
 * (OMC.jl:2373-2409).  X column-major n*m.  out[0] objective, out[1] MSE in, out[2] MSE out, out[3] MSE all */
int32_t omc_objective_mse(omc_problem* p, const double* X, double* out4);
/* measurement only (bench.py secondary roofline entries): mean device time in ms over `reps` back-to-back repetitions of
 * out_ms[0] the fused objective + MSE reduction on the last staged X, out_ms[1] the mask compaction sequence (expansion,
 * counts, scans, CSR / CSC fill) into scratch buffers.  CUDA events on the library stream.                                */
int32_t omc_profile_kernels(omc_problem* p, int32_t reps, float* out_ms);

/* ---- multi-GPU exchange: one process per GPU, the frontier sharded over the processes.  No
 * reference counterpart (OMC.jl:700-1073 is a single sequential loop).  NCCL is bound at run time (dlopen libnccl.so.2).
 * Bootstrap: rank 0 calls omc_comm_unique_id, ships the 128 bytes to the other ranks by any host channel, every rank calls
 * omc_comm_init after omc_init(device).  world = 1 makes every call a no-op, so single-GPU hosts need no NCCL.           */
int32_t omc_comm_unique_id(uint8_t* id128);
int32_t omc_comm_init(int32_t rank, int32_t world, const uint8_t* id128);
int32_t omc_comm_info(int32_t* rank, int32_t* world);
/* values[i] <- min over ranks (n host doubles): [incumbent upper bound, smallest open lower bound, ...] after a batch */
int32_t omc_allreduce_min(double* values, int32_t n);
/* recv[r * n + i] <- send[i] of rank r: per-rank load figures for the deterministic re-balancing of the open list */
int32_t omc_allgather(const double* send, int32_t n, double* recv);
int32_t omc_comm_destroy(void);
const char* omc_comm_last_error(void);

/* ---- diagnostics used by tests / bench ------------------------------------------------------- */
/* batched symmetric eigendecomposition + PSD projection of B dense N x N matrices (row-major, symmetric):
 * P[B*N*N] = projection onto the PSD cone, lam[B*N] = eigenvalues (unordered), sweeps[B] */
int32_t omc_debug_psd_project_batch(int32_t N, int32_t B, const double* Vin, double* P, double* lam,
                                    int32_t* sweeps, float* kernel_ms);
/* measured FP64 peaks of this GPU: out[0] = DFMA TFLOP/s, out[1] = DMMA (mma.m8n8k4.f64) TFLOP/s */
int32_t omc_measure_fp64_peak(double* out2);

#ifdef __cplusplus
}
#endif
#endif /* OMC_B200_H */
