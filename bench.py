#!/usr/bin/env python
"""bench.py -- nodes relaxed / second of the B200 bounding engine.

Workload ("config.workload"): BASELINE.json config 5, the configuration the metric is quoted on -- k = 5, 1000 x 1000 noisy
Gaussian low-rank data, 200 000 observed entries (20 %), gamma = 80, linear cuts, smallest_1_eigvec breakpoints; PSD blocks
2000 / 1005 / 1000, 85.5 MB of ADMM state per node.  The 4096-node frontier batch of the north star is 350 GB of state, so
one GPU relaxes a 128-node shard of it per step (8 GPUs: 1024 nodes); OMC_BENCH_CFG=C4 / C2 run config 4 (k = 3, 100 x 100,
linear3, smallest_2_eigvec; 1184 nodes per GPU) and config 2 the same way.  A *step* is one pass of the hot path -- the
relaxation of every node of one frontier batch of open branch-and-bound nodes (replaces matrix_completion_SDP_relaxation,
OMC.jl:1431-1943), each from a cold start to eps = 1e-8 with the root heuristic's incumbent as cut-off -- through the batched
large-block engine (csrc/omc_big.cuh): the frontier advances in lockstep, one short kernel sequence per ADMM iteration.  The
frontier is a committed fixture (tests/golden/<cfg>_frontier_pool.json, built once by best-first disjunctive expansion on
the GPU, scripts/dump_frontier_pool.py); at N GPUs it is sharded, not rebuilt.

  value     nodes/s with the batch descriptors resident in HBM (CUDA events on the library's launch stream around the
            whole lockstep run); only nodes that END the pass with a terminal status count (OPTIMAL / INFEASIBLE /
            CUTOFF); nodes stopped by max_iter are reported in config.status_counts and do not count
  e2e       same through omc_relax_batch with pinned HOST buffers (descriptor H2D + result D2H inside the timing), every step
  roofline  HBM: algorithmic bytes of one lockstep iteration (every array of a node record that the iteration must read or
            write once) x node-iterations / run time, against MEASURED_PEAKS.json's hbm_gbs
  cpu_baseline / --impl reference: the CPU oracle (NumPy restatement of the same program, exact eigh projections; Mosek and
            Julia are not installable here) on a bounded sample of the SAME frontier nodes.

N > 1 (torchrun): the frontier of N*B nodes is sharded block-cyclically, one process per GPU, no data-path collective; the
tiny [incumbent, min lower bound] all-reduce-min after each step goes over NCCL.
"""
import argparse
import os as _os, sys as _sys
if "reference" in _sys.argv:   # the oracle runs one node per process: keep BLAS single-threaded inside each
    _os.environ.setdefault("OMP_NUM_THREADS", "1"); _os.environ.setdefault("OPENBLAS_NUM_THREADS", "1"); _os.environ.setdefault("MKL_NUM_THREADS", "1")
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG = os.environ.get("OMC_BENCH_CFG", "C5")
CONFIGS = {   # (k, n, m, n_indices, cut type, nev, default nodes per GPU, max_iter)
    "C2": dict(k=1, n=50, m=50, nidx=1250, ct="linear", nev=1, nodes=1184, max_iter=5000,
               name="C2: k=1, 50x50 noisy, 1250 observed, gamma=80, linear cuts, smallest_1_eigvec"),
    "C4": dict(k=3, n=100, m=100, nidx=3000, ct="linear3", nev=2, nodes=1184, max_iter=4000,
               name="C4: k=3, 100x100 noisy, 3000 observed, gamma=80, linear3 cuts, smallest_2_eigvec"),
    "C5": dict(k=5, n=1000, m=1000, nidx=200000, ct="linear", nev=1, nodes=128, max_iter=3000,
               name="C5: k=5, 1000x1000, 20% observed, gamma=80, linear cuts, smallest_1_eigvec"),
}
W = CONFIGS[CFG]
WORKLOAD = W["name"] + "; frontier batch of open B&B nodes"
EPS = 1e-8
MAX_ITER = int(os.environ.get("OMC_BENCH_MAX_ITER", str(W["max_iter"])))
GAMMA = 80.0
FIXTURE = os.path.join(ROOT, "tests", "golden", f"{CFG.lower()}_frontier_pool.json")


def c2_instance(seed=0):
    from omc_b200.synthetic import generate_matrix_completion_data
    return generate_matrix_completion_data(1, 50, 50, 1250, seed)


def instance(seed=0):
    from omc_b200.synthetic import generate_matrix_completion_data  # host-side data generator of the package (not the oracle)
    return generate_matrix_completion_data(W["k"], W["n"], W["m"], W["nidx"], seed)


def node_iteration_bytes(n, m, k, L):
    """Algorithmic HBM bytes of ONE lockstep iteration of ONE node: every dense array of the node
    record that the iteration has to read or write, counted once per read and once per write, FP64; the symmetric arrays
    in the full storage the engine keeps.  X/Theta pass: X, Theta, V1[X], V1[X'], V1[Theta] read+write, mask*A read.
    Y/U passes: Y, V1[Y], V2, V3 read+write, Y~ write+read.  Tracker (one step): V1, V2, V3 read twice (two panel products)."""
    N1, N2 = n + m, n + k
    xt = 2 * (n * m) + 2 * (m * m) + 3 * (n * m) + 2 * (m * m) + n * m          # X rw, T rw, V1 X-part r + 2w, V1 T-part rw, AM r
    yu = 2 * (n * n) + 2 * (n * n) + 2 * (N2 * N2) + 2 * (n * n) + 2 * (n * n)  # Y rw, V1 Y-part rw, V2 rw, V3 rw, Y~ w+r
    tr = 2 * (N1 * N1 + N2 * N2 + n * n)                                       # two panel products per tracker step
    return 8.0 * (xt + yu + tr) + 8.0 * L * n * 4


# -------------------------------------------------------------------------------------------------
# clocks sampler
# -------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons}


# -------------------------------------------------------------------------------------------------
# frontier construction (untimed setup) -- built ONCE for the whole job (rank 0 builds, every rank takes its shard)
# -------------------------------------------------------------------------------------------------
def build_frontier_gpu(problem, target, omc, cutoff=float("inf"), cfg=None):
    """Best-first expansion: pop the open nodes with the smallest bound, relax them in one batch (with the incumbent as
    cut-off, like the branch-and-bound loop), branch on the separation oracle's breakpoint vector (master-feasible nodes are
    leaves, pruned nodes are dropped).  Returns the open BBNodes, best bound first."""
    from omc_b200.host import BBNode, JuliaPriorityQueue, create_matrix_cut_child_nodes
    from omc_b200.engine import LABELS
    cfg = cfg or W
    opts = omc.default_opts(eps_abs=EPS, eps_rel=EPS, max_iter=MAX_ITER, cutoff=cutoff)
    nodes = {1: BBNode(node_id=1, parent_id=0, LB=-np.inf, depth=0)}
    pq = JuliaPriorityQueue([(1, -np.inf)])
    counter = 1
    per_split = len(LABELS[cfg["ct"]]) ** cfg["k"]
    while len(nodes) < target and len(pq):
        room = max(1, (target - len(nodes) + per_split - 2) // (per_split - 1))     # every split adds (children - 1) open nodes
        batch = [nodes.pop(pq.dequeue_pair()[0]) for _ in range(min(148, len(pq), room))]
        res = problem.relax_batch([nd.disjunctive_cuts for nd in batch], opts)
        ok = [i for i, r in enumerate(res) if r["status_code"] == 0 and r["objective"] <= cutoff]
        if not ok:
            continue
        lam, vec, bp, feas = omc.smallest_eigvecs_batch(np.stack([res[i]["Y"] for i in ok]), np.stack([res[i]["U"] for i in ok]), cfg["nev"])
        for q, i in enumerate(ok):
            if feas[q] or batch[i].depth >= 60:
                continue
            kids = create_matrix_cut_child_nodes(problem, batch[i], bp[q], res[i]["U"], counter, res[i]["objective"])
            counter += len(kids)
            for kd in kids:
                nodes[kd.node_id] = kd
                pq.enqueue(kd.node_id, kd.LB)
    return sorted(nodes.values(), key=lambda nd: (nd.LB, nd.node_id))[:target]


def load_frontier_fixture(limit=None):
    """Cut descriptors of the first config-2 frontier nodes (tests/golden/c2_frontier.json, round-1 format): a list of
    oracle-style cut lists [(x, vhat, dirs), ...] per node."""
    with open(os.path.join(ROOT, "tests", "golden", "c2_frontier.json")) as f:
        fx = json.load(f)
    nodes = fx["nodes"][:limit] if limit else fx["nodes"]
    return [[(np.array(c["x"]), np.array(c["vhat"]), list(c["dirs"])) for c in nd["cuts"]] for nd in nodes]


def load_frontier_pool(path=None, limit=None):
    """The committed frontier of a configuration in pool form (scripts/dump_frontier_pool.py): the cuts of a tree are shared
    (a split adds ONE cut to all its children), so the fixture stores every cut once and each node as (pool id, direction
    codes) pairs.  Returns (list of oracle-style cut lists per node, incumbent objective)."""
    from omc_b200.engine import LABELS
    with open(path or FIXTURE) as f:
        fx = json.load(f)
    lab = LABELS[fx["cut_type"]]
    pool = [(np.array(c[0]), np.array(c[1])) for c in fx["pool"]]
    nodes = fx["nodes"][:limit] if limit else fx["nodes"]
    return [[(pool[e[0]][0], pool[e[0]][1], [lab[d] for d in e[1:]]) for e in nd] for nd in nodes], float(fx["incumbent"])


def _oracle_warm(_):
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import relaxation as R
    rng = np.random.default_rng(0)
    Aw = rng.standard_normal((6, 6))
    R.solve_relaxation(Aw, np.ones((6, 6), bool), 80.0, 1, opts=R.Options(max_iter=50))
    return 0


def _oracle_worker(args):
    os.environ["OMP_NUM_THREADS"] = "1"
    A, mask, gamma, k, ct, cuts, max_iter, cutoff = args
    from oracle import bigblock as Bg
    # exact-projection oracle has no cut-off rule; the restatement of the batched engine has the same rules as the GPU arm
    r = Bg.solve_relaxation_big(A, mask, gamma, k, ct, cuts, opts=Bg.BigOptions(eps_abs=EPS, eps_rel=EPS, max_iter=max_iter,
                                                                                infeasible_by_bound=True, cutoff=cutoff))
    return r["iters"], r["status"]


# -------------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the CPU restatement of the path (oracle port, exact eigh projections) on all host cores, on a bounded
    sample of the committed frontier fixture (the first nodes of the frontier the GPU arm relaxes)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    A, mask = instance(0)
    sample = max(2, min(cores, 64))
    if os.path.exists(FIXTURE):
        cuts, incumbent = load_frontier_pool()
        want = args.nodes or W["nodes"]
        ps = max(1, len(cuts) // want) if os.environ.get("OMC_BENCH_SAMPLE") == "strided" else 1
        cuts = cuts[::ps][:want]                                   # the GPU arm's single-GPU shard ...
    else:                                                          # config 2: the committed first nodes of the round-1 frontier
        cuts, incumbent = load_frontier_fixture(), float("inf")
    stride = max(1, len(cuts) // sample)
    cuts = cuts[::stride][:sample]                                 # ... every stride-th node of it: like for like
    jobs = [(A, mask, GAMMA, W["k"], W["ct"], c, MAX_ITER, incumbent) for c in cuts]
    ctx = mp.get_context("fork")
    with ctx.Pool(processes=min(cores, len(jobs))) as pool:
        pool.map(_oracle_warm, range(min(cores, len(jobs))))       # warm-up: import + LAPACK init in every worker
        t0 = time.perf_counter()
        iters = 0; done = 0; steps_done = 0
        for _ in range(args.steps):
            res = pool.map(_oracle_worker, jobs)
            iters += sum(r[0] for r in res); done += sum(1 for r in res if r[1] != 1); steps_done += 1
            if time.perf_counter() - t0 > 240.0:                   # bounded: the whole run ends within a few minutes
                break
        dt = time.perf_counter() - t0
    value = done / dt
    line = {"impl": "reference", "metric": "nodes relaxed/sec", "value": value, "unit": "nodes/s", "n_gpus": args.gpus,
            "steps": steps_done, "warmup": args.warmup, "ms_per_step": dt / steps_done * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "nodes_per_step": len(jobs), "eps": EPS, "max_iter": MAX_ITER,
                       "nodes_terminal_per_step": done / steps_done,
                       "cutoff": incumbent if np.isfinite(incumbent) else None,
                       "note": "CPU restatement of the relaxation (NumPy/LAPACK, oracle/bigblock.py: same ADMM, same cut-off and infeasibility rules as the "
                               "GPU arm), not Mosek: Julia and Mosek are absent; only nodes ending with a terminal status count, as in the GPU arm"},
            "cpu_baseline": {"value": value, "unit": "nodes/s", "cores": min(cores, len(jobs)), "kind": "port",
                             "sample": f"every {stride}th node of the GPU arm's {CFG} shard ({len(jobs)} nodes) x {steps_done} steps, {iters} ADMM iterations, one node per core"},
            "e2e": {"value": value, "unit": "nodes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_b200(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout for the one JSON line (NCCL prints its version banner)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import omc_b200
    omc = omc_b200
    omc.init(local)
    lib = omc._lib.load()
    A, mask = instance(0)
    n, m, k = W["n"], W["m"], W["k"]
    problem = omc.Problem(k, A, mask, GAMMA, W["ct"])
    B = args.nodes or W["nodes"]
    # ---- the frontier: taken from the committed fixture when it holds enough nodes, else built ONCE on rank 0 and broadcast
    t0 = time.time()
    if os.path.exists(FIXTURE):
        all_cuts, incumbent = load_frontier_pool()
    else:
        # config 2 has no committed pool: with the incumbent as cut-off its tree closes after 3 nodes, so (as in round 1) a wide
        # frontier is expanded on the GPU with the incumbent WITHHELD -- deterministically, by every rank for itself
        need = B if args.scaling == "strong" else B * world
        built = build_frontier_gpu(problem, need, omc, cutoff=float("inf"))
        all_cuts = [[(np.asarray(c.x), np.asarray(c.Uhat).T @ np.asarray(c.x), list(c.directions)) for c in nd.disjunctive_cuts] for nd in built]
        incumbent = float("inf")
    if len(all_cuts) < B * world and args.scaling != "strong":
        raise SystemExit(f"the committed frontier holds {len(all_cuts)} nodes, {B * world} asked (scripts/dump_frontier_pool.py builds a larger one)")
    src = os.path.relpath(FIXTURE, ROOT) if os.path.exists(FIXTURE) else "expanded on the GPU at start-up (incumbent withheld)"
    # The pool is ordered best bound first -- the order in which a best-first branch-and-bound pops its batches -- so the frontier
    # of a run is the PREFIX of B x world nodes (weak) or B nodes (strong).  Nodes get harder down the list (more cuts, more
    # ADMM iterations: mean 254 over the first 128 C5 nodes, 328 over the first 512), so nodes/s of a larger job is not N x the
    # single-GPU figure even when the ranks are balanced; config.node_iterations_per_s is the mix-independent rate.
    # OMC_BENCH_SAMPLE=strided takes an evenly strided sample of the pool instead (same statistical mix per GPU at every N).
    want = B if args.scaling == "strong" else B * world
    pool_stride = 1
    if os.environ.get("OMC_BENCH_SAMPLE") == "strided" and os.path.exists(FIXTURE):
        pool_stride = max(1, len(all_cuts) // max(want, 1))
    all_cuts = all_cuts[::pool_stride][:want]
    my_ids = list(range(len(all_cuts)))[rank::world]               # block-cyclic shard of the one frontier
    comm = None
    if world > 1:                                                  # the engine's own exchange (libomc_b200.so: NCCL via dlopen)
        from omc_b200.parallel import LibraryComm

        def bootstrap(raw):
            box = [raw]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        comm = LibraryComm(rank, world, bootstrap)
    pool_ids = {}

    def make_frontier(ids):
        cuts = []
        for i in ids:
            node = []
            for x, vh, d in all_cuts[i]:
                key = (x.tobytes(), vh.tobytes())
                if key not in pool_ids:
                    pool_ids[key] = problem.add_cut(x, vh)
                node.append(omc.Cut(pool_ids[key], x, vh, d))
            cuts.append(node)
        return cuts, omc.Frontier(problem, cuts)

    opts = omc.default_opts(eps_abs=EPS, eps_rel=EPS, max_iter=MAX_ITER, cutoff=incumbent)   # the incumbent prunes, as in the B&B loop
    node_cuts, fr = make_frontier(my_ids)
    if os.environ.get("OMC_STEPS_MAX"):
        fr.set_tuning(steps_max=int(os.environ["OMC_STEPS_MAX"]))
    rebalanced = False
    if world > 1 and not args.no_rebalance:
        # frontier re-balancing: one untimed pass gives every node's ADMM iteration count; the counts are all-gathered and
        # every rank takes its part of the same longest-first greedy partition (node descriptors are replicated: nothing moves)
        from omc_b200.parallel import balanced_partition
        fr.relax(opts)
        its = np.array([o["iters"] for o in fr.fetch(matrices=False)], dtype=np.float64)
        per = (len(all_cuts) + world - 1) // world
        send = np.full(per, -1.0); send[: len(its)] = its
        got = comm.allgather(send)
        cost = np.zeros(len(all_cuts))
        for r in range(world):
            ids_r = list(range(len(all_cuts)))[r::world]
            cost[ids_r] = got[r, : len(ids_r)]
        # lockstep cost model: below ~floor nodes an iteration costs its launch latencies (measured: a lone C5 node iterates at
        # ~0.8 ms, a node inside a batch at ~0.165 ms), so the shard that holds the longest node gets fewer node-iterations
        floor_nodes = int(os.environ.get("OMC_BALANCE_FLOOR", {"C5": 5}.get(CFG, 32)))
        my_ids = balanced_partition(cost, world, floor_nodes)[rank]
        fr.close()
        node_cuts, fr = make_frontier(my_ids)
        rebalanced = True
    setup_s = time.time() - t0
    Bl = len(node_cuts)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
    red = torch.zeros(2, dtype=torch.float64, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        flush.fill_(1.0)                       # L2 flush between steps (the frontier state, 1.1 MB per node, also exceeds L2)
        torch.cuda.synchronize()
        ms = fr.relax(opts)                    # CUDA events on the library stream bracket the whole lockstep run
        if world > 1:                          # the path's only exchange: all-reduce-min of [incumbent, min LB] (omc_allreduce_min)
            comm.allreduce_min(np.array([incumbent, 1e300]))
        return ms

    for _ in range(max(args.warmup, 0)):
        step()
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter()
    kernel_ms = [step() for _ in range(args.steps)]
    barrier()
    wall = time.perf_counter() - t0
    sampler.stop_flag = True
    out = fr.fetch(matrices=False)
    stats = fr.stats()
    iters = np.array([o["iters"] for o in out]); codes = np.array([o["status_code"] for o in out])
    status = np.bincount(codes, minlength=6)
    terminal = int(np.sum((codes == 0) | (codes == 2) | (codes == 4)))
    Ls = np.array([len(c) for c in node_cuts])
    bytes_step = float(sum(it * node_iteration_bytes(n, m, k, L) for it, L in zip(iters, Ls)))
    dev_ms = float(np.sum(kernel_ms)) / args.steps

    # ---- e2e: host buffers through omc_relax_batch (pinned), H2D of descriptors + D2H of results inside the timing, every step
    ptr, ids, dirs = problem._flatten(node_cuts)
    pin = lambda t: t.pin_memory()
    h_ptr, h_ids, h_dirs = pin(torch.from_numpy(ptr.copy())), pin(torch.from_numpy(ids.copy())), pin(torch.from_numpy(dirs.copy()))
    h_status = pin(torch.zeros(Bl, dtype=torch.int32)); h_iters = pin(torch.zeros(Bl, dtype=torch.int32))
    h_obj = pin(torch.zeros(Bl, dtype=torch.float64)); h_lb = pin(torch.zeros(Bl, dtype=torch.float64)); h_res = pin(torch.zeros(2 * Bl, dtype=torch.float64))
    h_X = pin(torch.zeros(Bl * n * m, dtype=torch.float64)); h_Y = pin(torch.zeros(Bl * n * n, dtype=torch.float64)); h_U = pin(torch.zeros(Bl * n * k, dtype=torch.float64))
    P = lambda t, ty: C.cast(t.data_ptr(), C.POINTER(ty))
    h2d = ptr.nbytes + ids.nbytes + dirs.nbytes
    d2h = sum(t.numel() * t.element_size() for t in (h_status, h_iters, h_obj, h_lb, h_res, h_X, h_Y, h_U))

    def e2e_step():
        flush.fill_(1.0)
        torch.cuda.synchronize()
        rc = lib.omc_relax_batch(problem.handle, Bl, P(h_ptr, C.c_int32), P(h_ids, C.c_int32), P(h_dirs, C.c_uint8), None, None,
                                 C.byref(opts), P(h_status, C.c_int32), P(h_obj, C.c_double), P(h_lb, C.c_double),
                                 P(h_iters, C.c_int32), P(h_res, C.c_double), P(h_X, C.c_double), P(h_Y, C.c_double),
                                 P(h_U, C.c_double), None, None)
        omc._lib.check(rc)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_wall = (time.perf_counter() - t0) / args.steps
    print(f"[bench rank {rank}] device ms per step {[round(v) for v in kernel_ms]}, e2e wall ms per step {e2e_wall * 1e3:.0f}, "
          f"terminal {terminal}/{Bl}, mean cuts/node {Ls.mean():.2f}, node-iterations {int(iters.sum())}, launches/step {stats['launches']}", file=sys.stderr, flush=True)

    # ---- max over ranks
    tmax = torch.tensor([dev_ms, wall / args.steps * 1e3, e2e_wall * 1e3], dtype=torch.float64, device="cuda")
    tsum = torch.tensor([float(Bl), bytes_step, float(iters.sum()), float(terminal), float(Ls.sum())] + [float(v) for v in status[:6]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_ms_max = [float(v) for v in tmax.tolist()]
    ts = [float(v) for v in tsum.tolist()]
    total_nodes, total_bytes, total_iters, total_terminal, total_cuts = ts[:5]
    status_all = [int(v) for v in ts[5:11]]

    if rank == 0:
        # cpu_baseline: the oracle on one core, bounded sample of the same frontier (N = 1 only)
        cpu = None
        if world == 1 and not args.no_cpu:
            from oracle import bigblock as Bg
            t0 = time.perf_counter(); done = 0; cit = 0; term = 0
            mine = [all_cuts[i] for i in my_ids]
            stride = max(1, len(mine) // 6)
            for cs in mine[::stride]:
                r = Bg.solve_relaxation_big(A, mask, GAMMA, k, W["ct"], cs, opts=Bg.BigOptions(eps_abs=EPS, eps_rel=EPS, max_iter=MAX_ITER,
                                                                                                infeasible_by_bound=True, cutoff=incumbent))
                done += 1; cit += r["iters"]; term += int(r["status"] != 1)
                if time.perf_counter() - t0 > 25.0:
                    break
            dt = time.perf_counter() - t0
            cpu = {"value": term / dt, "unit": "nodes/s", "cores": 1, "kind": "port",
                   "sample": f"{done} nodes of the same frontier (every {stride}th), {term} reached a terminal status, {cit} ADMM iterations, {dt:.1f} s; "
                             "NumPy/LAPACK restatement of the same ADMM (oracle/bigblock.py), not Mosek"}
        secondary = None
        if world == 1 and not args.no_secondary:
            secondary = secondary_metrics(omc)
        peak = 6541.8
        peak_src = "fallback of B200_PROFILING.md"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
        traffic = None
        try:   # DRAM bytes per node-iteration from the committed ncu capture of the same workload (profiles/; config 5 only)
            if CFG != "C5" or stats["engine"] != "batched":
                raise KeyError("no capture for this workload")
            with open(os.path.join(ROOT, "profiles", "r02_ncu_big_summary.json")) as f:
                traffic = float(json.load(f)["dram_bytes_per_node_iteration"]) * total_iters / max(1, world)
        except Exception:
            traffic = None
        achieved = bytes_step / (dev_ms * 1e-3) * 1e-9            # this rank
        line = {
            "metric": "nodes relaxed/sec", "value": total_terminal / (dev_ms_max * 1e-3), "unit": "nodes/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "engine": stats["engine"], "nodes_per_gpu": Bl, "nodes_total": int(total_nodes),
                       "nodes_terminal": int(total_terminal), "eps": EPS, "max_iter": MAX_ITER, "cutoff": incumbent if np.isfinite(incumbent) else None,
                       "start": "cold", "l2": "flushed between steps (256 MiB fill); frontier state exceeds L2", "frontier": src + (f", every {pool_stride}th node" if pool_stride > 1 else ", prefix in best-first order"),
                       "node_iterations_per_s": total_iters / (dev_ms_max * 1e-3),
                       "parallelism": f"one frontier of {int(total_nodes)} nodes sharded over {world} GPU(s), "
                                      + ("re-balanced by measured iterations (longest-first greedy, lockstep cost model), " if rebalanced else "block-cyclic, ")
                                      + "no data-path collective; all-reduce-min of 2 doubles per step inside libomc_b200.so",
                       "iters_per_node_mean": total_iters / total_nodes, "cuts_per_node_mean": total_cuts / total_nodes,
                       "status_counts[opt,iterlim,infeas,time,cutoff,numerical]": status_all,
                       "frontier_setup_s": setup_s, "wall_ms_per_step": wall_ms_max, "lockstep_iterations": stats["iterations"],
                       "node_bytes": stats["node_bytes"]},
            "e2e": {"value": total_terminal / (e2e_ms_max * 1e-3), "unit": "nodes/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(stats["launches"]) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "note": ("persistent engine (node state lives in shared memory / L2 for the whole ADMM): the HBM byte model of the batched engine "
                                  "does not describe it -- see DESIGN.md section 5, round-1 history, for its FP64 model; " if stats["engine"] != "batched" else "")
                                 + "batched engine, whole lockstep run (its kernels stream every node record once per pass): algorithmic bytes = node-iterations x "
                                 f"{node_iteration_bytes(n, m, k, 0):.0f} B (+ cut vectors), DESIGN.md section 4b; peak = {peak_src}; traffic = DRAM bytes measured by ncu "
                                 "per node-iteration (profiles/r02_ncu_big_summary.json) x node-iterations of this run"},
            "clocks": sampler.summary(),
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if secondary is not None:
            line["secondary"] = secondary
        print(json.dumps(line))
    fr.close()
    if world > 1:
        dist.destroy_process_group()


def secondary_metrics(omc):
    """The other two parts of BASELINE.json's metric, measured outside the timed region: wall time of the whole branch-and-bound
    to gap <= 1e-4 on config 2, and alt-min sweeps/s on config 5's shape."""
    from omc_b200.synthetic import generate_matrix_completion_data
    out = {}
    try:
        A2, m2 = generate_matrix_completion_data(1, 50, 50, 1250, 0)
        t0 = time.perf_counter()
        sol, _, inst = omc.matrix_completion_branchandbound(1, A2, m2, 80.0, node_selection="bestfirst", disjunctive_cuts_type="linear",
                                                             disjunctive_cuts_breakpoints="smallest_1_eigvec", time_limit=120, verbosity=0)
        out.update({"time_to_1e-4_gap_s_c2": time.perf_counter() - t0, "bnb_gap_c2": inst["tree"].now_gap,
                    "bnb_nodes_explored_c2": inst["run_details"]["nodes_explored"], "bnb_objective_c2": sol["objective"]})
        A5, m5 = generate_matrix_completion_data(5, 1000, 1000, 200000, 0)
        p5 = omc.Problem(5, A5, m5, 80.0, "linear")
        U5 = np.linalg.svd(np.where(m5, A5, 0.0))[0][:, :5]
        am = omc.alternating_minimization(p5, U5)
        rng5 = np.random.default_rng(0)
        starts5 = [U5] + [U5 + np.abs(U5).max() * rng5.standard_normal(U5.shape) for _ in range(147)]
        amb = omc.alternating_minimization_batch(p5, starts5, max_iters=20)
        out.update({"altmin_sweeps_per_s_c5": am["n_iters"] / max(am["solve_time"], 1e-9),
                    "altmin_sweeps_per_s_c5_batch148": sum(r["n_iters"] for r in amb) / max(amb[0]["solve_time"], 1e-9),
                    "altmin_c5": {"n_iters": am["n_iters"], "converged": am["converged"], "objective": am["objectives"][-1], "solve_time_s": am["solve_time"]}})
        # roofline entries of the smaller kernels on the path (HBM-bound; algorithmic bytes stated per entry)
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak = float(json.load(f)["hbm_gbs"])
        except Exception:
            peak = 6541.8
        n5, m5_, k5, nnz5 = 1000, 1000, 5, int(m5.sum())
        p5.objective_mse(am["U"] @ am["V"])                       # stages X on the device
        ms_obj, ms_mask = p5.profile_kernels(50)
        sweep_bytes = 2.0 * nnz5 * (8 + 4) + 2.0 * (n5 + m5_) * k5 * 8
        obj_bytes = 2.0 * 8 * n5 * m5_ + n5 * m5_ / 8
        mask_bytes = 8.0 * n5 * m5_ + 5.0 * n5 * m5_ / 8 + 8.0 * nnz5 + 4.0 * (n5 + m5_)

        def entry(nbytes, ms, note):
            g = nbytes / (ms * 1e-3) * 1e-9
            return {"bound": "hbm", "achieved": g, "peak": peak, "unit": "GB/s", "frac": g / peak, "bytes": nbytes, "ms": ms, "note": note}
        ms_sweep = am["solve_time"] * 1e3 / max(am["n_iters"], 1)
        ms_sweep_b = amb[0]["solve_time"] * 1e3 / max(sum(r["n_iters"] for r in amb), 1)
        out["rooflines"] = {
            "altmin_sweep_c5": entry(sweep_bytes, ms_sweep, "one instance: observed values + indices in CSR and CSC order, U and V read + written; "
                                     "1000 rows / columns of k = 5 normal equations -- latency-bound, one instance cannot fill the GPU"),
            "altmin_sweep_c5_batch148": entry(sweep_bytes, ms_sweep_b, "148 instances in one launch, time per instance-sweep"),
            "objective_mse_c5": entry(obj_bytes, ms_obj, "X and A (FP64) + mask bits, one pass, two-stage deterministic reduction; 16 MB is L2-resident "
                                      "across the 50 timed repetitions, so this is an L2-assisted figure"),
            "mask_compaction_c5": entry(mask_bytes, ms_mask, "BitMatrix chunks -> dense mask, row / column counts, scans, CSR + CSC fill (7 launches)"),
        }
        p5.close()
    except Exception as e:     # secondary numbers never take the headline down
        out["error"] = repr(e)[:200]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nodes", type=int, default=0, help="frontier nodes per GPU (default: 1184 = 8 x 148 SMs at config 4)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: one fixed frontier of --nodes nodes shared by all GPUs")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: keep the block-cyclic shards")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
