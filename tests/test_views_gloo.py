"""CPU, world_size 2 over gloo: the host-side multi-rank logic (view ownership, whole-job throughput
aggregation, gradient-bucket all-reduce).  The per-view compute here is the oracle, standing in for the
GPU ops which need a device."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from simplegaussiansplat_tk71_b200 import views, workloads as wl

    mine = views.views_for_rank(5, rank, world)
    elems = 0
    n_gauss = 64
    bucket = torch.zeros(n_gauss * views.PARAM_FLOATS_PER_GAUSSIAN)
    for v in mine:
        e = wl.c3("cpu", view=v, scale=0.002)
        y = orc.cumprod_fwd(e.x.numpy(), e.key.numpy(), np.float32)
        g = orc.cumprod_bwd_exact(e.x.numpy(), e.grad_out.numpy(), e.inv.numpy())
        elems += e.n
        bucket[v] += float(y.sum() + g.sum())  # stand-in for a per-Gaussian gradient contribution
    total, tmax = views.aggregate_throughput(elems, 10.0 + rank)
    views.allreduce_param_grads(bucket)
    q.put((rank, mine, elems, total, tmax, bucket[:5].tolist()))
    dist.destroy_process_group()


def test_two_rank_view_sharding_and_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, v0, e0, t0, m0, b0), (r1, v1, e1, t1, m1, b1) = res
    assert v0 == [0, 2, 4] and v1 == [1, 3]
    assert t0 == t1 == e0 + e1          # whole-job element count
    assert m0 == m1 == 11.0             # max over ranks
    assert b0 == b1 and all(abs(v) > 0 for v in b0)   # every rank holds the summed bucket


def test_views_for_rank_covers_all_views_once():
    sys.path.insert(0, ROOT)
    from simplegaussiansplat_tk71_b200 import views

    for world in (1, 2, 4, 8):
        seen = sorted(v for r in range(world) for v in views.views_for_rank(64, r, world))
        assert seen == list(range(64))


def test_balanced_view_ownership_is_a_partition_and_evens_out_the_load():
    import random

    from simplegaussiansplat_tk71_b200 import views

    rnd = random.Random(5)
    costs = [rnd.lognormvariate(0.0, 0.6) for _ in range(64)]
    for world in (1, 2, 3, 8):
        parts = [views.views_for_rank_balanced(costs, r, world) for r in range(world)]
        assert sorted(v for p in parts for v in p) == list(range(64))
        loads = [sum(costs[v] for v in p) for p in parts]
        rr = [sum(costs[v] for v in views.views_for_rank(64, r, world)) for r in range(world)]
        assert max(loads) <= max(rr) + 1e-12          # never worse than round-robin on the slowest rank
        assert max(loads) - min(loads) <= max(costs) + 1e-12
