"""World-size-2 CPU test (gloo) of the host-side sharding logic: contiguous site blocks keyed by GLOBAL site index,
one all-reduce (sum) of the statistics rows, replicated rate updates.  The device all-reduce itself (NCCL on the
chain's stream) is exercised by bench.py --gpus N; here the same partition + reduction is checked against the
single-process result with the oracle standing in for the device sweep."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import cases
from phylomap_b200 import dist as pdist


def test_shard_partition_covers_all_sites():
    for n, w in [(10, 3), (125000, 8), (7, 7), (1000001, 4)]:
        blocks = [pdist.shard(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
        for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
            assert s0 + c0 == s1
        assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import bridge
    z = cases.tree2(T=16, S=9, seed=17)
    local, start = pdist.shard_tree(z, rank, world)
    run = bridge.OracleRun(bridge.PLAIN, [local.oracle_dict()], cases.Q2, cases.PID2, 0.2, 6, rng_mode=bridge.KEYED,
                           seed=3, site_offset=start)
    rows = pdist.combine_rows_host(run.run())
    # rate-updating sampler: per-iteration reduction is emulated by reducing the per-shard rows of a fixed-Q run; the
    # replicated host update needs identical rows on every rank, which is what this asserts
    gathered = [None] * world
    dist.all_gather_object(gathered, rows.tolist())
    if rank == 0:
        q.put((rows, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_sum_to_the_single_process_run(oracle):
    world = 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    rows, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    z = cases.tree2(T=16, S=9, seed=17)
    full = oracle.OracleRun(oracle.PLAIN, [z.oracle_dict()], cases.Q2, cases.PID2, 0.2, 6, rng_mode=oracle.KEYED, seed=3).run()
    assert np.array_equal(rows[:, 2:], full[:, 2:])             # integer counts: exact
    np.testing.assert_allclose(rows[:, :2], full[:, :2], rtol=1e-13)
    assert all(np.array_equal(np.array(g), rows) for g in gathered)  # every rank sees the same reduced rows
