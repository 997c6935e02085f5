"""Host-side mirror of the reference's operator interface over the C ABI.

Function names, argument meaning and result-dict keys follow
/root/reference/src/OptimalMatrixCompletion.jl (OMC.jl) so the parity tests read like
tests of the reference; every numerical body is a call into libomc_b200.so (CUDA,
sm_100a).  NumPy is used for buffers only.  Nothing here imports ``oracle``.
"""
import ctypes as C
import time
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from . import _lib
from ._lib import RelaxOpts, check, CUT_TYPES

LABELS = {
    "linear": ["left", "right"],                                  # OMC.jl:2482
    "linear2": ["left", "middle", "right"],                       # OMC.jl:2486
    "linear3": ["left", "inner_left", "inner_right", "right"],    # OMC.jl:2490
}
# status -> the MOI termination status name the Julia glue reports (OMC.jl:1866-1940)
MOI_STATUS = {0: "OPTIMAL", 1: "SLOW_PROGRESS", 2: "INFEASIBLE", 3: "TIME_LIMIT", 4: "OPTIMAL", 5: "NUMERICAL_ERROR"}

_initialised = None


def init(device: Optional[int] = None):
    """omc_init: one process drives one GPU.  device=None keeps the GPU already selected (0 on first use)."""
    global _initialised
    lib = _lib.load()
    if device is None:
        device = _initialised if _initialised is not None else 0
    if _initialised != device:
        check(lib.omc_init(device))
        _initialised = device
    return lib


def build_flags() -> int:
    """omc_build_flags: bit 0 = the kernel was built with the primal infeasibility certificate."""
    return int(_lib.load().omc_build_flags())


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a, ty):
    return a.ctypes.data_as(C.POINTER(ty)) if a is not None else None


def bitmatrix_chunks(indices: np.ndarray) -> np.ndarray:
    """Julia ``BitMatrix.chunks`` of a bool (n, m) array: column-major bit index, LSB first."""
    n, m = indices.shape
    flat = np.asarray(indices, dtype=bool).T.reshape(-1)
    nchunks = (n * m + 63) // 64
    padded = np.zeros(nchunks * 64, dtype=np.uint8)
    padded[: n * m] = flat
    return np.packbits(padded.reshape(nchunks, 64), axis=1, bitorder="little").view(np.uint64).reshape(-1).copy()


def default_opts(**kw) -> RelaxOpts:
    o = RelaxOpts()
    _lib.load().omc_relax_default_opts(C.byref(o))
    for key, v in kw.items():
        if not hasattr(o, key):
            raise TypeError(f"unknown relaxation option {key}")
        setattr(o, key, v)
    return o


@dataclass
class Cut:
    """One entry of BBNodeDisjunctiveCuts.cuts (OMC.jl:33-35): (breakpoint_vec, Uhat, directions)."""
    cut_id: int
    x: np.ndarray
    Uhat: np.ndarray
    directions: List[str]


class Problem:
    """(k, A, indices, gamma) resident in HBM, its cut pool and warm-start state pool."""

    def __init__(self, k: int, A: np.ndarray, indices: np.ndarray, gamma: float, disjunctive_cuts_type: str = "linear",
                 state_pool_capacity: int = 0, device: Optional[int] = None):
        if disjunctive_cuts_type not in CUT_TYPES:
            raise ValueError('Disjunctive cuts type must be either "linear" or "linear2" or "linear3"; '
                             f"{disjunctive_cuts_type} supplied instead.")         # OMC.jl:1456-1462
        if A.shape != indices.shape:
            raise ValueError("Dimension mismatch. Input matrix A must have size (n, m); "
                             "Input matrix indices must have size (n, m).")        # OMC.jl:240-246
        self.lib = init(device)
        self.k, self.gamma, self.cut_type = int(k), float(gamma), disjunctive_cuts_type
        self.n, self.m = A.shape
        self.A = np.asfortranarray(A, dtype=np.float64)
        self.indices = np.asarray(indices, dtype=bool)
        self._chunks = bitmatrix_chunks(self.indices)
        self.handle = C.c_void_p()
        check(self.lib.omc_problem_create(self.n, self.m, self.k, _ptr(self.A, C.c_double),
                                          _ptr(self._chunks, C.c_uint64), self.gamma, CUT_TYPES[disjunctive_cuts_type],
                                          int(state_pool_capacity), C.byref(self.handle)))
        self.state_pool_capacity = int(state_pool_capacity)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.omc_problem_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- mask compaction ------------------------------------------------------------------
    def csr(self):
        nnz = C.c_int64()
        check(self.lib.omc_problem_get_csr(self.handle, None, None, None, None, C.byref(nnz)))
        rowptr = np.zeros(self.n + 1, np.int32); colptr = np.zeros(self.m + 1, np.int32)
        colidx = np.zeros(max(nnz.value, 1), np.int32); rowidx = np.zeros(max(nnz.value, 1), np.int32)
        check(self.lib.omc_problem_get_csr(self.handle, _ptr(rowptr, C.c_int32), _ptr(colidx, C.c_int32),
                                           _ptr(colptr, C.c_int32), _ptr(rowidx, C.c_int32), C.byref(nnz)))
        return rowptr, colidx[: nnz.value], colptr, rowidx[: nnz.value]

    def profile_kernels(self, reps: int = 20):
        """omc_profile_kernels: mean device ms of (objective + MSE reduction, mask compaction) -- measurement only."""
        out = np.zeros(2, np.float32)
        check(self.lib.omc_profile_kernels(self.handle, int(reps), _ptr(out, C.c_float)))
        return float(out[0]), float(out[1])

    # ---- Shor valid inequalities --------------------------------------------------------------
    def set_shor(self, minors, soc_coords):
        """add_Shor_valid_inequalities = true (OMC.jl:1503-1552, 1755-1846): minors (N, 4) = (i1, i2, j1, j2) and SOC coordinates
        (M, 2), both 0-based as shor_constraint_indexes returns them.  Empty lists remove the rows."""
        mn = np.ascontiguousarray(np.asarray(minors, np.int32).reshape(-1, 4))
        sc = np.ascontiguousarray(np.asarray(soc_coords, np.int32).reshape(-1, 2))
        check(self.lib.omc_problem_set_shor(self.handle, len(mn), _ptr(mn, C.c_int32) if len(mn) else None,
                                            len(sc), _ptr(sc, C.c_int32) if len(sc) else None))
        self.has_shor = len(mn) + len(sc) > 0

    # ---- cut pool -------------------------------------------------------------------------
    def add_cut(self, x: np.ndarray, Uhat: np.ndarray) -> int:
        """Registers (breakpoint_vec, Uhat) created at OMC.jl:2522; only vhat = Uhat'x is uploaded (OMC.jl:1577)."""
        x = _f64(x)
        vhat = _f64(np.asarray(Uhat).T @ x) if np.asarray(Uhat).ndim == 2 else _f64(Uhat)
        cid = C.c_int32()
        check(self.lib.omc_cutpool_add(self.handle, _ptr(x, C.c_double), _ptr(vhat, C.c_double), C.byref(cid)))
        return cid.value

    def _flatten(self, node_cuts: List[List[Cut]]):
        B = len(node_cuts)
        ptr = np.zeros(B + 1, np.int32)
        ids, dirs = [], []
        lab = LABELS[self.cut_type]
        for b, cuts in enumerate(node_cuts):
            ptr[b + 1] = ptr[b] + len(cuts)
            for c in cuts:
                ids.append(c.cut_id)
                dirs.extend(lab.index(d) for d in c.directions)
        ids = np.asarray(ids if ids else [0], np.int32)
        dirs = np.asarray(dirs if dirs else [0], np.uint8)
        return ptr, ids, dirs

    # ---- relaxation -----------------------------------------------------------------------
    def frontier(self, node_cuts: List[List[Cut]], warm_ids=None, save_ids=None, engine: str = "auto") -> "Frontier":
        return Frontier(self, node_cuts, warm_ids, save_ids, engine)

    def relax_batch(self, node_cuts: List[List[Cut]], opts: Optional[RelaxOpts] = None, warm_ids=None, save_ids=None,
                    engine: str = "auto"):
        """Bodies of B calls of matrix_completion_SDP_relaxation (OMC.jl:1431-1943) in one launch (engine "persistent")
        or one lockstep run (engine "batched"); "auto" = batched when n + m > 104."""
        f = Frontier(self, node_cuts, warm_ids, save_ids, engine)
        try:
            t0 = time.perf_counter()
            f.relax(opts)
            out = f.fetch()
            dt = time.perf_counter() - t0
        finally:
            f.close()
        for r in out:
            r["solve_time"] = dt / len(out)
        return out

    # ---- objective / MSE ------------------------------------------------------------------
    def objective_mse(self, X: np.ndarray) -> Tuple[float, float, float, float]:
        if X.shape != (self.n, self.m):
            raise ValueError("Dimension mismatch. Input matrix X must have size (n, m).")
        Xf = np.asfortranarray(X, dtype=np.float64)
        out = np.zeros(4)
        check(self.lib.omc_objective_mse(self.handle, _ptr(Xf, C.c_double), _ptr(out, C.c_double)))
        return tuple(float(v) for v in out)


class Frontier:
    """A batch of open nodes resident in HBM (omc_frontier)."""

    def __init__(self, problem: Problem, node_cuts, warm_ids=None, save_ids=None, engine: str = "auto"):
        self.p = problem
        self.B = len(node_cuts)
        ptr, ids, dirs = problem._flatten(node_cuts)
        self._keep = (ptr, ids, dirs)
        w = np.asarray(warm_ids, np.int32) if warm_ids is not None else None
        s = np.asarray(save_ids, np.int32) if save_ids is not None else None
        self.handle = C.c_void_p()
        check(problem.lib.omc_frontier_create_ex(problem.handle, self.B, _ptr(ptr, C.c_int32), _ptr(ids, C.c_int32),
                                                 _ptr(dirs, C.c_uint8), _ptr(w, C.c_int32), _ptr(s, C.c_int32),
                                                 _lib.ENGINES[engine], C.byref(self.handle)))
        self.kernel_ms = None

    def fetch_shor(self):
        """Shor results per node (OMC.jl:1902, 1913): W (B, n, m) and Xt (B, k, n, m)."""
        p, B = self.p, self.B
        W = np.zeros((B, p.m, p.n)); Xt = np.zeros((B, p.k, p.m, p.n))
        check(p.lib.omc_frontier_fetch_shor(self.handle, _ptr(W, C.c_double), _ptr(Xt, C.c_double)))
        return W.transpose(0, 2, 1).copy(), Xt.transpose(0, 1, 3, 2).copy()

    def stats(self) -> dict:
        """omc_frontier_stats: engine, kernel launches, lockstep iterations, node-iterations, checks, rho changes, bytes per node."""
        out = np.zeros(8, np.int64)
        check(self.p.lib.omc_frontier_stats(self.handle, _ptr(out, C.c_int64)))
        return dict(engine={1: "persistent", 2: "batched"}[int(out[0])], launches=int(out[1]), iterations=int(out[2]),
                    node_iterations=int(out[3]), checks=int(out[4]), rho_changes=int(out[5]), node_bytes=int(out[6]))

    def debug_fetch(self, node: int, which: int, cap: int) -> np.ndarray:
        out = np.zeros(cap)
        cnt = self.p.lib.omc_frontier_debug_fetch(self.handle, int(node), int(which), _ptr(out, C.c_double), int(cap))
        if cnt < 0:
            raise RuntimeError("omc_frontier_debug_fetch failed")
        return out[:cnt]

    def set_tuning(self, steps_max=0, steps_start=0, track_tol=0.0, confirm_tol=0.0):
        check(self.p.lib.omc_frontier_set_tuning(self.handle, int(steps_max), int(steps_start), float(track_tol), float(confirm_tol)))

    def relax(self, opts: Optional[RelaxOpts] = None) -> float:
        ms = C.c_float()
        check(self.p.lib.omc_frontier_relax(self.handle, C.byref(opts) if opts is not None else None, C.byref(ms)))
        self.kernel_ms = ms.value
        return ms.value

    def fetch(self, matrices: bool = True):
        p, B = self.p, self.B
        status = np.zeros(B, np.int32); iters = np.zeros(B, np.int32)
        obj = np.zeros(B); lb = np.zeros(B); res = np.zeros(2 * B)
        X = np.zeros((B, p.m, p.n)) if matrices else None   # column-major (n, m) per node
        Y = np.zeros((B, p.n, p.n)) if matrices else None
        U = np.zeros((B, p.k, p.n)) if matrices else None
        check(p.lib.omc_frontier_fetch(self.handle, _ptr(status, C.c_int32), _ptr(obj, C.c_double),
                                       _ptr(lb, C.c_double), _ptr(iters, C.c_int32), _ptr(res, C.c_double),
                                       _ptr(X, C.c_double), _ptr(Y, C.c_double), _ptr(U, C.c_double), None))
        out = []
        for b in range(B):
            if int(status[b]) == _lib.STATUS_NUMERICAL:               # OMC.jl:1936-1940: the reference errors here too
                raise RuntimeError(f"unexpected termination status: NUMERICAL_ERROR (node {b} of the batch)")
            r = {
                "model": None,                                        # OMC.jl:1861 (never read by the host loop)
                "termination_status": MOI_STATUS[int(status[b])],     # OMC.jl:1863
                "status_code": int(status[b]),
                "feasible": int(status[b]) != _lib.STATUS_INFEASIBLE,  # OMC.jl:1879, 1935
                "objective": float(obj[b]),                           # OMC.jl:1882-1895
                "lower_bound": float(lb[b]) if lb[b] > -1e299 else float("-inf"),   # no certified bound yet
                "iters": int(iters[b]),
                "res_p": float(res[2 * b]), "res_d": float(res[2 * b + 1]),
            }
            if matrices:
                r["X"] = X[b].T.copy(); r["Y"] = Y[b].T.copy(); r["U"] = U[b].T.copy()
            out.append(r)
        return out

    def profile(self) -> np.ndarray:
        """(B, 16): SM cycles per phase (w-update, build V, DMMA pre-rotation, Jacobi, reconstruction, residuals), sweeps, iterations."""
        prof = np.zeros((self.B, 32))
        check(self.p.lib.omc_frontier_fetch_profile(self.handle, _ptr(prof, C.c_double)))
        return prof

    def close(self):
        if self.handle:
            self.p.lib.omc_frontier_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- reference-named operators ------------------------------------------------------------------

def matrix_completion_SDP_relaxation(problem: Problem, cuts: List[Cut], opts: Optional[RelaxOpts] = None, **kw):
    """OMC.jl:1431-1943 for one node (disjunctive path)."""
    return problem.relax_batch([cuts], opts, **kw)[0]


def evaluate_objective(problem: Problem, X: np.ndarray) -> float:
    """OMC.jl:2330-2359."""
    return problem.objective_mse(X)[0]


def compute_MSE(problem: Problem, X: np.ndarray, kind: str = "out") -> float:
    """OMC.jl:2373-2409."""
    o = problem.objective_mse(X)
    if kind == "in":
        return o[1]
    if kind == "out":
        return o[2]
    if kind == "all":
        return o[3]
    raise ValueError('Input argument `kind` not recognized! Must be one of "out", "in", or "all".')


def smallest_eigvecs_batch(Y: np.ndarray, U: np.ndarray, nev: int = 1):
    """Batched OMC.jl:2466-2477 + 1272-1277.  Y (B,n,n), U (B,n,k) -> lam (B,nev), vec (B,n,nev), breakpoint (B,n), feasible (B,)"""
    lib = init()
    Y = np.asarray(Y, dtype=np.float64); U = np.asarray(U, dtype=np.float64)
    if Y.ndim == 2:
        Y = Y[None]; U = U[None]
    B, n, _ = Y.shape
    k = U.shape[2]
    Yc = np.ascontiguousarray(np.transpose(Y, (0, 2, 1)))   # column-major per node
    Uc = np.ascontiguousarray(np.transpose(U, (0, 2, 1)))
    lam = np.zeros((B, nev)); vec = np.zeros((B, nev, n)); bp = np.zeros((B, n)); feas = np.zeros(B, np.int32)
    check(lib.omc_smallest_eigvecs_batch(n, k, B, _ptr(Yc, C.c_double), _ptr(Uc, C.c_double), nev,
                                         _ptr(lam, C.c_double), _ptr(vec, C.c_double), _ptr(bp, C.c_double),
                                         _ptr(feas, C.c_int32)))
    return lam, np.transpose(vec, (0, 2, 1)).copy(), bp, feas.astype(bool)


def matrix_completion_master_feasible(Y, U, X=None, Theta=None, use_disjunctive_cuts=True) -> bool:
    """OMC.jl:1261-1292 (disjunctive path only)."""
    if not use_disjunctive_cuts:
        raise NotImplementedError("McCormick path is out of scope (SURVEY.md section 2)")
    return bool(smallest_eigvecs_batch(Y, U, 1)[3][0])


def alternating_minimization(problem: Problem, U_initial: np.ndarray, disjunctive_cuts: Optional[List[Cut]] = None,
                             eps: float = 1e-5, max_iters: int = 100, time_limit: float = 3600.0):
    """OMC.jl:1979-2279 (disjunctive path); returns the reference's result dict (OMC.jl:2249-2278)."""
    cuts = disjunctive_cuts or []
    if U_initial.shape != (problem.n, problem.k):
        raise ValueError("U_initial must have size (n, k)")
    Ui = np.asfortranarray(U_initial, dtype=np.float64)
    lab = LABELS[problem.cut_type]
    ids = np.asarray([c.cut_id for c in cuts] or [0], np.int32)
    dirs = np.asarray([lab.index(d) for c in cuts for d in c.directions] or [0], np.uint8)
    U = np.zeros((problem.k, problem.n)); V = np.zeros((problem.m, problem.k)); obj = np.zeros(max_iters)
    conv = C.c_int32(); nit = C.c_int32(); st = C.c_double()
    check(problem.lib.omc_altmin(problem.handle, _ptr(Ui, C.c_double), len(cuts), _ptr(ids, C.c_int32), _ptr(dirs, C.c_uint8),
                                 float(eps), int(max_iters), float(time_limit), _ptr(U, C.c_double), _ptr(V, C.c_double),
                                 C.byref(conv), C.byref(nit), _ptr(obj, C.c_double), C.byref(st)))
    return {"converged": bool(conv.value), "U": U.T.copy(), "V": V.T.copy(), "solve_time": st.value,
            "n_iters": nit.value, "max_iters": max_iters, "objectives": [float(v) for v in obj[: nit.value]]}


def alternating_minimization_batch(problem: Problem, U_initials, disjunctive_cuts_list=None, eps: float = 1e-5,
                                   max_iters: int = 100, time_limit: float = 3600.0):
    """B alt-min runs of one problem in one launch (one CTA per instance): the reference's extra root restarts
    (OMC.jl:529-538) or the alt-min calls of a popped batch of nodes.  Returns a list of the reference's result dicts."""
    B = len(U_initials)
    cl = disjunctive_cuts_list if disjunctive_cuts_list is not None else [[] for _ in range(B)]
    if len(cl) != B:
        raise ValueError("one cut list per instance")
    n, m, k = problem.n, problem.m, problem.k
    Ui = np.zeros((B, k, n))
    for b, U0 in enumerate(U_initials):
        if U0.shape != (n, k):
            raise ValueError("U_initial must have size (n, k)")
        Ui[b] = np.asarray(U0, dtype=np.float64).T
    lab = LABELS[problem.cut_type]
    ptr = np.zeros(B + 1, np.int32)
    ids, dirs = [], []
    for b, cuts in enumerate(cl):
        ids += [c.cut_id for c in cuts]
        dirs += [lab.index(d) for c in cuts for d in c.directions]
        ptr[b + 1] = len(ids)
    ids = np.asarray(ids or [0], np.int32); dirs = np.asarray(dirs or [0], np.uint8)
    U = np.zeros((B, k, n)); V = np.zeros((B, m, k)); obj = np.zeros((B, max_iters))
    conv = np.zeros(B, np.int32); nit = np.zeros(B, np.int32); st = C.c_double()
    check(problem.lib.omc_altmin_batch(problem.handle, B, _ptr(Ui, C.c_double), _ptr(ptr, C.c_int32), _ptr(ids, C.c_int32),
                                       _ptr(dirs, C.c_uint8), float(eps), int(max_iters), float(time_limit),
                                       _ptr(U, C.c_double), _ptr(V, C.c_double), _ptr(conv, C.c_int32), _ptr(nit, C.c_int32),
                                       _ptr(obj, C.c_double), C.byref(st)))
    return [{"converged": bool(conv[b]), "U": U[b].T.copy(), "V": V[b].T.copy(), "solve_time": st.value, "n_iters": int(nit[b]),
             "max_iters": max_iters, "objectives": [float(v) for v in obj[b, : nit[b]]]} for b in range(B)]


def shor_constraint_indexes(problem: Problem, num_entries_present_list, with_soc: bool = True):
    """generate_rank1_matrix_completion_Shor_constraints_indexes (OMC.jl:2545-2612) and the SOC coordinate list (OMC.jl:656-665)
    on the GPU.  Returns (tuples (N, 4) int32, soc (M, 2) int32), both 0-BASED (add 1 for the reference's 1-based tuples)."""
    pl = np.asarray(list(num_entries_present_list), np.int32)
    cnt = C.c_int64(0); nsoc = C.c_int64(0)
    check(problem.lib.omc_shor_indexes(problem.handle, _ptr(pl, C.c_int32), len(pl), C.byref(cnt), None, 0, None, None))
    tuples = np.zeros((max(1, cnt.value), 4), np.int32)
    soc = np.zeros((problem.n * problem.m, 2), np.int32) if with_soc else None
    check(problem.lib.omc_shor_indexes(problem.handle, _ptr(pl, C.c_int32), len(pl), C.byref(cnt), _ptr(tuples, C.c_int32), cnt.value,
                                       _ptr(soc, C.c_int32) if with_soc else None, C.byref(nsoc) if with_soc else None))
    return tuples[: cnt.value], (soc[: nsoc.value] if with_soc else None)


def generate_violated_Shor_minors(problem: Problem, Xt: np.ndarray, candidates: np.ndarray, existing, n_minors: int):
    """generate_violated_Shor_minors (OMC.jl:2614-2640) on the GPU.  Xt (slices, n, m) -- the reference's call site (OMC.jl:2496-2502)
    passes reshape(X, (1, n, m)); candidates (N, 4) and existing (M, 4) 0-based
    minors (candidates = shor_constraint_indexes(problem, pattern)[0], computed once per run).  Returns (scores, tuples) of the
    n_minors most violated candidates that are not in `existing`, in the reference's order."""
    Xt = np.asarray(Xt, np.float64)
    assert Xt.ndim == 3 and Xt.shape[1:] == (problem.n, problem.m), Xt.shape
    Xs = np.ascontiguousarray(Xt.transpose(0, 2, 1))              # k column-major n x m slices
    cand = np.ascontiguousarray(np.asarray(candidates, np.int32).reshape(-1, 4))
    ex = np.ascontiguousarray(np.asarray(existing, np.int32).reshape(-1, 4))
    cnt = C.c_int64(0)
    cap = max(1, min(int(n_minors), len(cand)))
    tuples = np.zeros((cap, 4), np.int32); scores = np.zeros(cap)
    check(problem.lib.omc_shor_score_minors(problem.handle, _ptr(Xs, C.c_double), int(Xt.shape[0]), len(cand), _ptr(cand, C.c_int32) if len(cand) else None,
                                            len(ex), _ptr(ex, C.c_int32) if len(ex) else None, int(n_minors), C.byref(cnt),
                                            _ptr(tuples, C.c_int32), _ptr(scores, C.c_double)))
    return scores[: cnt.value], tuples[: cnt.value]


def psd_project_batch(V: np.ndarray):
    """Eigensolver self-test entry: V (B,N,N) symmetric -> (P, lam, sweeps, kernel_ms)."""
    lib = init()
    V = np.ascontiguousarray(V, dtype=np.float64)
    B, N, _ = V.shape
    P = np.zeros_like(V); lam = np.zeros((B, N)); sw = np.zeros(B, np.int32); ms = C.c_float()
    check(lib.omc_debug_psd_project_batch(N, B, _ptr(V, C.c_double), _ptr(P, C.c_double), _ptr(lam, C.c_double),
                                          _ptr(sw, C.c_int32), C.byref(ms)))
    return P, lam, sw, ms.value


def measure_fp64_peak():
    lib = init()
    out = np.zeros(2)
    check(lib.omc_measure_fp64_peak(_ptr(out, C.c_double)))
    return {"dfma_tflops": float(out[0]), "dmma_tflops": float(out[1])}
