"""N>1 host logic on CPU: two processes over gloo plan the hash exchange of a partitioned join the way bench.py /
tgpu_exchange_partitioned do, with the oracle's partition function standing in for the device kernel.
Checks: shards cover the table, both join sides co-locate by key, and the count matrix is consistent."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    import oracle_lib as o
    from trino_b200.page import Block, Page
    from trino_b200.sharding import exchange_plan, shard_range

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_orders = 7001
    total_rows = o.synth_lineitem_rows(n_orders)
    o_first, o_count = shard_range(n_orders, world, rank)
    l_first, l_count = shard_range(total_rows, world, rank)
    okeys = o.synth_orders_keys(n_orders, o_first, o_count, 0x7C02, True)
    lkeys = o.synth_lineitem_keys(n_orders, l_first, l_count, 0x7C01, False)
    # HashBucketFunction over the join key, bucket == rank
    o_part = o.partition_ids(Page(Block.bigint(okeys)), [0], world)
    l_part = o.partition_ids(Page(Block.bigint(lkeys)), [0], world)
    send = torch.tensor([int((l_part == r).sum()) for r in range(world)], dtype=torch.int64)
    gathered = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, send)
    matrix = [g.tolist() for g in gathered]
    recv_counts, recv_offsets, total = exchange_plan(matrix, rank)
    # "exchange": every rank publishes its rows per destination (all_gather_object is the CPU stand-in for NCCL send/recv)
    payload = [lkeys[l_part == r] for r in range(world)]
    everything = [None] * world
    dist.all_gather_object(everything, payload)
    mine = np.concatenate([everything[src][rank] for src in range(world)])
    assert len(mine) == total
    for src in range(world):
        seg = mine[recv_offsets[src]:recv_offsets[src] + recv_counts[src]]
        assert (seg == everything[src][rank]).all()
    build_payload = [okeys[o_part == r] for r in range(world)]
    all_build = [None] * world
    dist.all_gather_object(all_build, build_payload)
    my_build = np.concatenate([all_build[src][rank] for src in range(world)])
    # co-location: every probe key that landed here finds its (unique) build key here
    assert np.isin(mine, my_build).all()
    stats = torch.tensor([len(mine), len(my_build), int(mine.sum() % (1 << 40))], dtype=torch.int64)
    allstats = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allstats, stats)
    # the checksum plumbing of bench.py --workload … (bench_workloads.py): 64-bit wrap-around sums and max-over-ranks timing
    import bench_workloads as bw
    big = [(1 << 63) + 12345 + rank, (1 << 64) - 1, rank]
    tot = bw._allreduce_u64(dist, 0, big)
    assert tot == [(sum((1 << 63) + 12345 + r for r in range(world))) & bw.M64, (world * ((1 << 64) - 1)) & bw.M64, sum(range(world))]
    assert bw._max_ms(dist, 0, 1.5 + rank) == 1.5 + world - 1
    if rank == 0:
        assert sum(int(s[0]) for s in allstats) == total_rows
        assert sum(int(s[1]) for s in allstats) == n_orders
        out.put("ok")
    dist.destroy_process_group()


def test_shard_ranges_cover_the_table():
    sys.path.insert(0, ROOT)
    from trino_b200.sharding import exchange_plan, shard_range
    for total in (0, 1, 7, 1000, 600_000_003):
        for world in (1, 2, 3, 8):
            ranges = [shard_range(total, world, r) for r in range(world)]
            assert ranges[0][0] == 0
            assert sum(c for _, c in ranges) == total
            for (f0, c0), (f1, _) in zip(ranges, ranges[1:]):
                assert f0 + c0 == f1
    counts, offs, total = exchange_plan([[1, 2], [3, 4]], 1)
    assert (counts, offs, total) == ([2, 4], [0, 2], 6)


def test_two_rank_exchange_plan_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(timeout=5) == "ok"
