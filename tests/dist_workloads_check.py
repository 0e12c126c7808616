"""Multi-GPU parity of the bench_workloads.py pipelines (launched with torchrun, one rank per GPU; tests/test_gpu_dist.py runs it under
pytest -m gpu): the star join with replicated (tgpu_exchange_broadcast) and partitioned builds against the oracle chain as a multiset,
the broadcast exchange itself, and the q3way / q1 runners at toy scale (their closed-form and oracle spot checks).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_workloads_check.py
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench_workloads as bw  # noqa: E402
from helpers import oracle_star_rows  # noqa: E402
from trino_b200 import abi  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402
from trino_b200.exchange import Exchange  # noqa: E402
from trino_b200.page import AbiPage, Block, Page  # noqa: E402
from trino_b200.sharding import shard_range  # noqa: E402


class NoClocks:
    def __init__(self, index):
        pass

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    # ---- broadcast exchange: every rank ends with every rank's rows, in rank order; NULLs on one rank only
    xc = Exchange(ctx, dist, rank, world, local)
    mine = Page(Block.bigint(np.arange(rank * 100, rank * 100 + 10 + rank)), Block.double(np.arange(10 + rank) * 0.5, (np.arange(10 + rank) % 3 == 0) if rank == world - 1 else None))
    got = xc.broadcast(AbiPage(mine))
    everyone = [None] * world
    dist.all_gather_object(everyone, mine.rows())
    assert got.to_host().rows() == [r for rows in everyone for r in rows]
    got.release()
    empty = xc.broadcast(AbiPage(Page(Block.bigint([]), Block.double([]))) if rank == 0 else AbiPage(Page(Block.bigint([7]), Block.double([1.0]))))
    assert empty.to_host().rows() == [(7, 1.0)] * (world - 1)
    empty.release()
    # ---- star join over `world` shards == the oracle chain over the whole fact table (as a multiset: the exchange regroups rows)
    n = 96 * 1024
    total = n * world
    first, _ = shard_range(total, world, rank)
    ptr, both, dims, keep = bw.star_tables(ctx, ops, abi, world, rank, n, first)
    probes, part, closers = bw.star_pipeline(ctx, ops, abi, xc, dims, world)
    xc.create_arenas(int(n * 1.5) * 41 + (8 << 20))
    I64, F64 = abi.INT64, abi.FLOAT64
    rows = []
    half = n // 2
    for lo, m in ((0, half), (half, n - half)):
        page = ops.DevicePage([ops.DeviceColumn(I64, ptr["date_sk"] + lo * 8, m), ops.DeviceColumn(I64, ptr["item_sk"] + lo * 8, m),
                               ops.DeviceColumn(I64, ptr["customer_sk"] + lo * 8, m, validity=ptr["customer_valid"] + lo // 8),
                               ops.DeviceColumn(I64, ptr["store_sk"] + lo * 8, m, validity=ptr["store_valid"] + lo // 8),
                               ops.DeviceColumn(F64, ptr["net_paid"] + lo * 8, m)], m)
        out, held = bw.star_chunk(ctx, ops, xc, probes, part, page)
        if out is not None:
            rows += out.to_host().rows()
            out.release()
        for p in reversed(held):
            if p:
                p.release()
    gathered = [None] * world
    dist.all_gather_object(gathered, rows)
    if rank == 0:
        want, both_o = oracle_star_rows(total, 0)
        allrows = sorted(r for part_rows in gathered for r in part_rows)
        assert len(allrows) == both_o == len(want)
        assert allrows == sorted(want)
    for p in probes.values():
        p.close()
    builders, bridges, kept = closers
    for b in builders:
        b.close()
    for br in bridges.values():
        br.lookup_source.close()
    dist.barrier()
    xc.close()
    # ---- the bench entry points at toy scale (each builds its own communicator and verifies itself)
    args = types.SimpleNamespace(sf=0.05, ds_sf=0.1, star_chunks=3, q1_sf=0.02, steps=1, warmup=1)
    for w in ("q3way", "star", "q1"):
        line = bw.RUNNERS[w](args, ctx, rank, world, local, dist, NoClocks)
        assert line["n_gpus"] == world and line["value"] > 0, line
    dist.barrier()
    if rank == 0:
        print(f"dist_workloads_check ok (broadcast, star, q3way, q1): world={world}")
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
