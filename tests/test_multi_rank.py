"""CPU test of the N > 1 host path: world_size = 2 over gloo (the GPU box uses NCCL with the same calls)."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import omc_b200  # noqa: F401
    from omc_b200.parallel import shard_block_cyclic, allreduce_bounds
    nodes = list(range(101, 101 + 37))
    mine = shard_block_cyclic(nodes, rank, world)
    # every rank "relaxes" its shard: pretend bound = node id / 10, incumbent found by the rank owning node 120
    lbs = [nid / 10.0 for nid in mine]
    ub = 11.5 if 120 in mine else 99.0
    gub, glb = allreduce_bounds(ub, min(lbs))
    q.put((rank, mine, gub, glb))
    dist.barrier()
    dist.destroy_process_group()


def test_frontier_sharding_and_bound_allreduce_world2():
    world, port = 2, 29500 + (os.getpid() % 2000)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    shards = [o[1] for o in out]
    assert sorted(shards[0] + shards[1]) == list(range(101, 138)) and not set(shards[0]) & set(shards[1])
    assert abs(len(shards[0]) - len(shards[1])) <= 1
    for _, _, gub, glb in out:                      # both ranks agree on the global incumbent and lower bound
        assert gub == 11.5 and glb == 10.1
    sys.path.insert(0, ROOT)
    import omc_b200  # noqa: F401
    from omc_b200.parallel import unshard_block_cyclic, rebalance_counts
    assert unshard_block_cyclic(shards) == list(range(101, 138))
    c = rebalance_counts([1000, 3000], [10, 10])
    assert c.sum() == 20 and c[0] > c[1]


def test_balanced_partition_is_deterministic_and_even():
    sys.path.insert(0, ROOT)
    import omc_b200  # noqa: F401
    from omc_b200.parallel import balanced_partition
    rng = np.random.default_rng(0)
    cost = rng.integers(200, 4000, size=1024).astype(float)
    parts = balanced_partition(cost, 8)
    assert sorted(i for p in parts for i in p) == list(range(1024))
    loads = np.array([cost[p].sum() for p in parts])
    assert loads.max() / loads.min() < 1.02                       # per-rank iteration totals within 2 %
    assert all(cost[p[0]] >= cost[p[-1]] for p in parts)          # each rank's queue is ordered longest first
    assert parts == balanced_partition(cost, 8)


def test_balanced_partition_lockstep_cost_model_offloads_the_straggler_rank():
    """One straggler (2150 iterations among ~260-iteration nodes, the C5 frontier's profile): with floor_nodes = 5 the rank that
    holds it gets fewer node-iterations, and the predicted lockstep times of the ranks even out."""
    sys.path.insert(0, ROOT)
    import omc_b200  # noqa: F401
    from omc_b200.parallel import balanced_partition, lockstep_cost
    rng = np.random.default_rng(1)
    cost = rng.integers(200, 330, size=256).astype(float)
    cost[17] = 2150.0
    plain = balanced_partition(cost, 2)
    model = balanced_partition(cost, 2, floor_nodes=5)
    assert sorted(i for p in model for i in p) == list(range(256))
    t_plain = [lockstep_cost(sorted(cost[p], reverse=True), 5) for p in plain]
    t_model = [lockstep_cost(sorted(cost[p], reverse=True), 5) for p in model]
    assert max(t_model) < max(t_plain) and max(t_model) / min(t_model) < 1.02
    r = [q for q in range(2) if 17 in model[q]][0]
    assert cost[model[r]].sum() < cost[model[1 - r]].sum()
    assert lockstep_cost([10.0, 4.0, 2.0], 0) == 16.0 and lockstep_cost([10.0, 4.0, 2.0], 2) == 22.0


def test_library_comm_single_rank_is_a_noop():
    """world = 1: the in-library exchange (omc_comm_* / omc_allreduce_min / omc_allgather) needs neither NCCL nor a GPU."""
    sys.path.insert(0, ROOT)
    import omc_b200  # noqa: F401
    from omc_b200.parallel import LibraryComm
    c = LibraryComm(0, 1)
    assert np.array_equal(c.allreduce_min([3.0, -1.0]), [3.0, -1.0])
    assert np.array_equal(c.allgather([1.0, 2.0, 3.0]), [[1.0, 2.0, 3.0]])
    c.close()
