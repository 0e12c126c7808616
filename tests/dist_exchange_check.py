"""Multi-GPU parity check (launched with torchrun, one rank per GPU; tests/test_gpu_dist.py runs it under pytest -m gpu):
hash-partition + NCCL all-to-all of a device page through tgpu_exchange_partitioned, checked against the oracle's
partition function, then the partitioned join against a single-process oracle join.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_exchange_check.py
"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import oracle_lib as o  # noqa: E402
from trino_b200 import abi  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402
from trino_b200.page import Block, Page  # noqa: E402
from trino_b200.sharding import shard_range  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    lib = ctx.lib
    idb = (C.c_uint8 * abi.COMM_ID_BYTES)()
    if rank == 0:
        ctx.check(lib.tgpu_comm_get_unique_id(C.cast(idb, C.c_void_p)))
    t = torch.tensor(list(idb), dtype=torch.uint8, device=f"cuda:{local}")
    dist.broadcast(t, 0)
    idb = (C.c_uint8 * abi.COMM_ID_BYTES)(*t.cpu().tolist())
    ctx.check(lib.tgpu_comm_init(ctx.h, C.cast(idb, C.c_void_p), rank, world))

    n_orders = 200_003
    total_rows = o.synth_lineitem_rows(n_orders)
    o_first, o_count = shard_range(n_orders, world, rank)
    l_first, l_count = shard_range(total_rows, world, rank)
    okeys = o.synth_orders_keys(n_orders, o_first, o_count, 0x7C02, True)
    lkeys = o.synth_lineitem_keys(n_orders, l_first, l_count, 0x7C01, False)
    rng = np.random.default_rng(rank)
    lnull = rng.random(l_count) < 0.01
    lpage = Page(Block.bigint(lkeys), Block.double(lkeys * 0.25, lnull), Block.integer((lkeys % 1000).astype(np.int32)))
    opage = Page(Block.bigint(okeys), Block.bigint(okeys % 2557))

    part = ops.PartitionedOutputOperatorFactory(ctx, [0], world).create_operator()

    def exchange(page):
        from trino_b200.page import AbiPage
        ap = AbiPage(page)
        pp = abi.PP()
        ctx.check(lib.tgpu_exchange_partitioned(ctx.h, part.h, ap.ref(), C.byref(pp)))
        return ctx.page_to_host(pp)

    got_l = exchange(lpage)
    got_o = exchange(opage)
    # a column that carries NULLs on the last rank only: every other rank's page has no validity buffer at all, but must still
    # ship a NULL-byte lane because some rank's page does (the advisor's round-1 finding on the warp-granular scatter)
    onesided = Page(Block.bigint(lkeys), Block.double(lkeys * 0.25, lnull if rank == world - 1 else None))
    got_1 = exchange(onesided)
    nulls_1 = torch.tensor([sum(1 for v in got_1.get_block(1).to_pylist() if v is None), int(lnull.sum()) if rank == world - 1 else 0],
                           dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(nulls_1)
    assert nulls_1[0] == nulls_1[1]
    assert all(v is None or v == k * 0.25 for k, v in zip(got_1.get_block(0).to_pylist(), got_1.get_block(1).to_pylist()))
    # every received row belongs here (oracle partition function), and globally nothing is lost or duplicated
    assert (o.partition_ids(got_l, [0], world) == rank).all()
    assert (o.partition_ids(got_o, [0], world) == rank).all()
    stats = torch.tensor([got_l.position_count, got_o.position_count, int(np.asarray(got_l.get_block(0).values).sum() % (1 << 50)),
                          int(sum(1 for v in got_l.get_block(1).to_pylist() if v is None))], dtype=torch.int64, device=f"cuda:{local}")
    mine = torch.tensor([l_count, o_count, int(lkeys.sum() % (1 << 50)), int(lnull.sum())], dtype=torch.int64, device=f"cuda:{local}")
    dist.all_reduce(stats)
    dist.all_reduce(mine)
    assert stats[0] == mine[0] == total_rows and stats[1] == mine[1] == n_orders
    assert stats[2] % (1 << 50) == mine[2] % (1 << 50)
    assert stats[3] == mine[3]
    # rows arrive grouped by source rank, each group in the sender's row order: payload follows its key
    vals = got_l.get_block(1).to_pylist()
    keys = got_l.get_block(0).to_pylist()
    assert all(v is None or v == k * 0.25 for k, v in zip(keys, vals))
    assert (np.asarray(got_l.get_block(2).values) == (np.asarray(keys) % 1000)).all()
    # partitioned join == the oracle join restricted to this rank's keys
    from helpers import gpu_join_rows, oracle_join_rows
    rows = gpu_join_rows(ctx, [got_o], [got_l], 0, 0, [0, 2], [1], abi.JOIN_INNER, False)
    want = oracle_join_rows(got_o, got_l, 0, 0, [0, 2], [1], abi.JOIN_INNER, False)
    assert rows == want and len(rows) == got_l.position_count
    # ---- peer-memory path: same exchange through P2P stores into the destination's arena, twice (arenas alternate)
    hb = (C.c_uint8 * (abi.NUM_ARENAS * abi.IPC_HANDLE_BYTES))()
    ctx.check(lib.tgpu_comm_arena_create(ctx.h, 64 << 20, C.cast(hb, C.c_void_p)))
    mine_h = torch.tensor(list(hb), dtype=torch.uint8, device=f"cuda:{local}")
    gathered = [torch.zeros_like(mine_h) for _ in range(world)]
    dist.all_gather(gathered, mine_h)
    allh = (C.c_uint8 * (world * abi.NUM_ARENAS * abi.IPC_HANDLE_BYTES))(*torch.cat(gathered).cpu().tolist())
    ctx.check(lib.tgpu_comm_arena_open(ctx.h, C.cast(allh, C.c_void_p)))
    for _ in range(3):
        p2p_l = exchange(lpage)
        assert p2p_l.rows() == got_l.rows()          # identical rows in identical order to the NCCL path
        p2p_o = exchange(opage)
        assert p2p_o.rows() == got_o.rows()
    assert exchange(onesided).rows() == got_1.rows()
    # ---- pipelined form: the exchange runs on `ctx`, build + probe on a second context of the same GPU; the probe of page k is
    # only enqueued, its output is taken after exchange k+1 was issued (tgpu_exchange_partitioned_fenced guards the arena reuse)
    from trino_b200.page import AbiPage
    pctx = ops.Context(local)
    bridge = ops.JoinBridge()
    b = ops.HashBuilderOperatorFactory(pctx, bridge, [0], [1]).create_operator()
    b.add_input(got_o)
    b.finish()
    probe = ops.LookupJoinOperatorFactory(pctx, bridge, abi.JOIN_INNER, False, [0], [0, 2]).create_operator()
    half = l_count // 2
    pieces = [Page(Block.bigint(lkeys[sl]), Block.double(lkeys[sl] * 0.25, lnull[sl]), Block.integer((lkeys[sl] % 1000).astype(np.int32)))
              for sl in (slice(0, half), slice(half, l_count))]
    want_rows = []
    for piece in pieces:
        ap = AbiPage(piece)
        pp = abi.PP()
        ctx.check(lib.tgpu_exchange_partitioned(ctx.h, part.h, ap.ref(), C.byref(pp)))
        recv = ctx.page_to_host(pp)
        want_rows.append(oracle_join_rows(got_o, recv, 0, 0, [0, 2], [1], abi.JOIN_INNER, False))
    got_rows = [[], []]
    inflight = None
    for it in range(6):
        piece = pieces[it % 2]
        ap = AbiPage(piece)
        pp = abi.PP()
        ctx.check(lib.tgpu_exchange_partitioned_fenced(ctx.h, part.h, ap.ref(), pctx.h, C.byref(pp)))
        inp = ops.DeviceOutputPage(ctx, pp)
        if inflight is not None:
            out = probe.get_output()
            got_rows[(it - 1) % 2] = out.rows() if out is not None else []
            assert got_rows[(it - 1) % 2] == want_rows[(it - 1) % 2], f"pipelined join differs at iteration {it - 1}"
            inflight.release()
        probe.add_input(inp.as_device_page())
        inflight = inp
    out = probe.get_output()
    assert (out.rows() if out is not None else []) == want_rows[1]
    inflight.release()
    probe.close()
    b.close()
    # ---- split-phase form: two exchanges in flight, transfers on the copy engines; rows and order identical to the blocking call
    handles = []
    seq = [lpage, opage, onesided, lpage, lpage, onesided, opage]
    want_seq = [got_l.rows(), got_o.rows(), got_1.rows(), got_l.rows(), got_l.rows(), got_1.rows(), got_o.rows()]
    aps = []
    done = 0
    for page in seq:
        ap = AbiPage(page)
        aps.append(ap)
        h = C.c_void_p()
        ctx.check(lib.tgpu_exchange_begin(ctx.h, part.h, ap.ref(), C.byref(h)))
        handles.append(h)
        if len(handles) == 2:
            pp = abi.PP()
            ctx.check(lib.tgpu_exchange_end(ctx.h, handles.pop(0), C.byref(pp)))
            assert ctx.page_to_host(pp).rows() == want_seq[done], f"split-phase exchange {done} differs"
            done += 1
    while handles:
        pp = abi.PP()
        ctx.check(lib.tgpu_exchange_end(ctx.h, handles.pop(0), C.byref(pp)))
        assert ctx.page_to_host(pp).rows() == want_seq[done], f"split-phase exchange {done} differs"
        done += 1
    assert done == len(seq)
    # ---- general exchange: variable-width columns and replicated rows (nullChannel rows go to EVERY rank, the first row of the first
    # page too: PagePartitioner.java:229-241,401-416).  Every rank regenerates every sender's page and partitions it with the oracle:
    # the received page must be the senders' parts in rank order, each in the oracle's row order.
    def general_page(r, salt):
        g = np.random.default_rng(1000 * salt + r)
        m = 4000 + 137 * r
        keys = g.integers(0, 10**6, m)
        return Page(Block.bigint(keys, g.random(m) < 0.02),
                    Block.varchar([None if i % 11 == 0 else "s%d-%s" % (k, "x" * int(k % 7)) for i, k in enumerate(keys)]),
                    Block.double(keys * 0.5, g.random(m) < 0.05),
                    Block.int128([None if i % 13 == 0 else int(k) * 10**22 - 5 for i, k in enumerate(keys)]))      # a long DECIMAL column

    for null_channel, any_row in ((-1, False), (0, False), (0, True)):
        gpart = ops.PartitionedOutputOperatorFactory(ctx, [0], world, None, null_channel, any_row).create_operator()
        states = [False] * world
        for salt, split_phase in ((1, False), (2, True), (3, False)):
            want = []
            for r in range(world):
                pg = general_page(r, salt)
                lists, states[r] = o.partition_positions(pg, [0], world, None, world, null_channel, any_row, states[r])
                rows_r = pg.rows()
                want += [rows_r[i] for i in lists[rank]]
            ap = AbiPage(general_page(rank, salt))
            pp = abi.PP()
            if split_phase:
                h = C.c_void_p()
                ctx.check(lib.tgpu_exchange_begin(ctx.h, gpart.h, ap.ref(), C.byref(h)))
                ctx.check(lib.tgpu_exchange_end(ctx.h, h, C.byref(pp)))
            else:
                ctx.check(lib.tgpu_exchange_partitioned(ctx.h, gpart.h, ap.ref(), C.byref(pp)))
            got = ctx.page_to_host(pp).rows()
            assert got == want, f"general exchange differs (null_channel={null_channel}, any_row={any_row}, page {salt}): {len(got)} vs {len(want)} rows"
        gpart.close()
    # broadcast of a page with a variable-width column: every rank ends up with every rank's rows, in rank order
    pp = abi.PP()
    bap = AbiPage(general_page(rank, 9))          # (kept alive across the call: the descriptor points into its arrays)
    ctx.check(lib.tgpu_exchange_broadcast(ctx.h, bap.ref(), C.byref(pp)))
    assert ctx.page_to_host(pp).rows() == [row for r in range(world) for row in general_page(r, 9).rows()]
    part.close()
    pctx.close()
    ctx.check(lib.tgpu_comm_destroy(ctx.h))
    dist.barrier()
    if rank == 0:
        print(f"dist_exchange_check ok (NCCL, P2P, fenced, split-phase and general paths): world={world} rows={total_rows}")
    dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
