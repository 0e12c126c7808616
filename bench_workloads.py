"""Secondary bench workloads (python bench.py --workload q3way|star|q1), each runnable at N = 1 and, under torchrun, at N = 2/4/8:

  q3way  BASELINE.json configs[3]: lineitem JOIN orders JOIN customer, both sides of each join co-located by a hash exchange on the
         join key (orders + lineitem on orderkey, then the joined rows re-partitioned on custkey), SURVEY.md §8(d)/(e)
  star   BASELINE.json configs[4]: TPC-DS store_sales x date_dim, item, store (REPLICATED builds: tgpu_exchange_broadcast) and
         customer (partitioned build, fact rows exchanged on ss_customer_sk); NULL fact keys never match
         (M/operator/join/unspilled/JoinProbe.java:154-171)
  q1     TPC-H Q1 across GPUs: PARTIAL aggregation per rank -> hash exchange of the intermediate rows on the group key -> FINAL
         (M/operator/HashAggregationOperator.java:366-393)

`value` = input rows/s over all ranks, device-resident inputs, CUDA events on the library's stream, max over ranks.  Every workload ends
with an untimed verification pass (closed-form checksums of the synthetic data; the parity tests compare the same pipelines row by
row with the oracle at reduced scale: tests/test_gpu_workloads.py, tests/dist_workloads_check.py).
"""
import ctypes as C
import json
import os

import numpy as np

M64 = (1 << 64) - 1
SEED_LINEITEM, SEED_ORDERS, SEED_CUSTKEY, SEED_STORE_SALES = 0x7C01, 0x7C02, 0x7C03, 0xD501
DATE_SK0 = 2415022           # first date_dim surrogate key (TPC-DS), 73 049 days


def _project(ctx, ops, abi, col, n, expr):
    """one computed column over a device column through the FilterAndProject operator"""
    page = ops.DevicePage([col], n)
    op = ops.FilterAndProjectOperatorFactory(ctx, ops.PageProcessorProgram(None, [expr])).create_operator()
    op.add_input(page)
    out = op.get_output_device()
    op.close()
    return out          # keep the DeviceOutputPage alive as long as its column is used


def _column_sum(ctx, abi, col, mod=0):
    c = abi.Column()
    c.type, c.flags, c.length, c.data, c.offsets, c.validity = col.type, 0, col.length, col.ptr, col.offsets, col.validity
    v = C.c_int64()
    ctx.check(ctx.lib.tgpu_column_sum(ctx.h, C.byref(c), mod, C.byref(v)))
    return v.value & M64


def _allreduce_u64(dist, local, values):
    """sum of python ints modulo 2^64 over the ranks (hi / lo halves travel as int64)"""
    if dist is None:
        return [v & M64 for v in values]
    import torch
    device = f"cuda:{local}" if dist.get_backend() == "nccl" else "cpu"       # (gloo: the CPU test of this helper)
    t = torch.tensor([v >> 32 for v in values] + [v & 0xFFFFFFFF for v in values], dtype=torch.int64, device=device)
    dist.all_reduce(t)
    t = [int(x) for x in t.tolist()]
    k = len(values)
    return [((t[i] << 32) + t[k + i]) & M64 for i in range(k)]


def _max_ms(dist, local, ms):
    if dist is None:
        return ms
    import torch
    t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local}" if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _timed(ctx, dist, local, step, steps, warmup, clock_sampler):
    for _ in range(warmup):
        step(None)
    ctx.synchronize()
    if dist is not None:
        dist.barrier()
    sampler = clock_sampler(local)
    sampler.start()
    launches0 = ctx.kernel_launches + 0
    ctx.timer_start()
    for _ in range(steps):
        step(None)
    ms = ctx.timer_stop_ms()
    launches = ctx.kernel_launches + 0 - launches0
    clocks = sampler.stop()
    if dist is not None:
        dist.barrier()
    return _max_ms(dist, local, ms) / steps, launches, clocks


# ---------------------------------------------------------------------------------------------------------------------------------
def run_q3way(args, ctx, rank, world, local, dist, clocks):
    from trino_b200 import abi
    from trino_b200 import operators as ops
    from trino_b200.exchange import Exchange
    from trino_b200.sharding import shard_range
    lib = ctx.lib
    sf = args.sf
    n_orders = int(1_500_000 * sf)                       # per GPU
    n_cust = int(150_000 * sf) // 3 * 3                  # per GPU (a multiple of 3: custkey % 3 != 0 have orders)
    total_orders, total_cust = n_orders * world, n_cust * world
    total_rows = lib.tgpu_synth_lineitem_rows(total_orders)
    o_first, _ = shard_range(total_orders, world, rank)
    c_first, _ = shard_range(total_cust, world, rank)
    l_first, l_count = shard_range(total_rows, world, rank)
    I64 = abi.INT64
    d_okeys, d_ocust, d_ckeys = ctx.malloc(n_orders * 8), ctx.malloc(n_orders * 8), ctx.malloc(n_cust * 8)
    d_lkeys = ctx.malloc(l_count * 8)
    ctx.check(lib.tgpu_synth_orders_keys(ctx.h, total_orders, o_first, n_orders, SEED_ORDERS, 1, C.c_void_p(d_okeys)))
    ctx.check(lib.tgpu_synth_orders_custkeys(ctx.h, total_orders, o_first, n_orders, SEED_ORDERS, 1, total_cust, SEED_CUSTKEY, C.c_void_p(d_ocust)))
    ctx.check(lib.tgpu_synth_sequence(ctx.h, c_first + 1, n_cust, C.c_void_p(d_ckeys)))
    ctx.check(lib.tgpu_synth_lineitem_keys(ctx.h, total_orders, l_first, l_count, SEED_LINEITEM, 0, C.c_void_p(d_lkeys)))
    B, D = abi.V_BIGINT, abi.V_DOUBLE
    price = _project(ctx, ops, abi, ops.DeviceColumn(I64, d_lkeys, l_count), l_count,
                     ops.Call(abi.EX_MUL, ops.Call(abi.EX_CAST_BIGINT_TO_DOUBLE, ops.Col(0, B)), ops.Const(0.5, D)))
    nation = _project(ctx, ops, abi, ops.DeviceColumn(I64, d_ckeys, n_cust), n_cust, ops.Call(abi.EX_MOD, ops.Col(0, B), ops.Const(25, B)))
    lineitem = ops.DevicePage([ops.DeviceColumn(I64, d_lkeys, l_count), price.column(0)], l_count)
    orders = ops.DevicePage([ops.DeviceColumn(I64, d_okeys, n_orders), ops.DeviceColumn(I64, d_ocust, n_orders)], n_orders)
    customer = ops.DevicePage([ops.DeviceColumn(I64, d_ckeys, n_cust), nation.column(0)], n_cust)

    xc = Exchange(ctx, dist, rank, world, local)
    part0 = xc.partitioner([0]) if world > 1 else None           # orders / lineitem / customer on their key (channel 0)
    part2 = xc.partitioner([2]) if world > 1 else None           # (l_orderkey, l_extendedprice, o_custkey) on o_custkey
    keep = []
    o_in = xc.partitioned(part0, orders)
    c_in = xc.partitioned(part0, customer)
    keep += [o_in, c_in]
    bridge_o, bridge_c = ops.JoinBridge(), ops.JoinBridge()
    build_o = ops.HashBuilderOperatorFactory(ctx, bridge_o, [0], [1], n_orders).create_operator()
    build_o.add_input(o_in.as_device_page() if o_in else orders)
    build_o.finish()
    build_c = ops.HashBuilderOperatorFactory(ctx, bridge_c, [0], [1], n_cust).create_operator()
    build_c.add_input(c_in.as_device_page() if c_in else customer)
    build_c.finish()
    ctx.synchronize()
    probe_o = ops.LookupJoinOperatorFactory(ctx, bridge_o, abi.JOIN_INNER, False, [0], [0, 1]).create_operator()
    probe_c = ops.LookupJoinOperatorFactory(ctx, bridge_c, abi.JOIN_INNER, False, [2], [0, 1, 2]).create_operator()
    xc.create_arenas(int(l_count * 1.3) * 24 + (8 << 20))
    chk = {"rows": 0, "lkey": 0, "nation": 0, "cust_mod": 0, "spot": 0}

    def step(check):
        a = xc.partitioned(part0, lineitem)
        probe_o.add_input(a.as_device_page() if a else lineitem)
        b = probe_o.get_output_device()                          # l_orderkey, l_extendedprice, o_custkey
        c = xc.partitioned(part2, b.as_device_page())
        probe_c.add_input(c.as_device_page() if c else b.as_device_page())
        d = probe_c.get_output_device()                          # l_orderkey, l_extendedprice, o_custkey, c_nationkey
        if check is not None:
            check["rows"] += d.rows
            check["lkey"] = (check["lkey"] + _column_sum(ctx, abi, d.column(0))) & M64
            check["nation"] = (check["nation"] + _column_sum(ctx, abi, d.column(3))) & M64
            check["cust_mod"] = (check["cust_mod"] + _column_sum(ctx, abi, d.column(2), 25)) & M64
            if not check["spot"] and d.rows:
                import oracle_lib as o
                from trino_b200.page import Block, Page
                m = min(d.rows, 1 << 18)
                host = [np.empty(m, np.int64) for _ in range(4)]
                for arr, ci in zip(host, range(4)):
                    ctx.check(lib.tgpu_memcpy_d2h(ctx.h, C.c_void_p(arr.ctypes.data), C.c_void_p(d.column(ci).ptr), m * 8))
                assert (host[3] == host[2] % 25).all() and (host[1].view(np.float64) == host[0] * 0.5).all()
                if world > 1:
                    assert (o.partition_ids(Page(Block.bigint(host[2])), [0], world) == rank).all(), "row on the wrong rank after the custkey exchange"
                check["spot"] = m
        for p in (d, c, b, a):
            if p:
                p.release()

    ms, launches, sample = _timed(ctx, dist, local, step, args.steps, args.warmup, clocks)
    step(chk)
    in_key = _column_sum(ctx, abi, ops.DeviceColumn(I64, d_lkeys, l_count))
    tot = _allreduce_u64(dist, local, [chk["rows"], chk["lkey"], chk["nation"], chk["cust_mod"], in_key])
    verify = {"rows_out": tot[0], "rows_in": total_rows, "lineitem_key_sum_conserved": tot[1] == tot[4], "nation_sum_matches_custkeys": tot[2] == tot[3],
              "oracle_spot_check_rows_per_rank": chk["spot"]}
    assert tot[0] == total_rows and verify["lineitem_key_sum_conserved"] and verify["nation_sum_matches_custkeys"], verify
    line = {"workload": "q3way", "metric": "partitioned_3way_join_probe_rows_per_sec", "value": total_rows / (ms * 1e-3), "unit": "rows/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "int64", "data": "synthetic",
            "gpu_launches": int(launches), "clocks": sample, "verify": verify,
            "config": {"workload": f"lineitem JOIN orders JOIN customer, synthetic TPC-H SF{sf:g} per GPU (BASELINE.json configs[3] is SF300 over 8 GPUs = SF37.5 per GPU), "
                                   "hash exchange on orderkey, join, re-partition on custkey, join",
                       "lineitem_rows_per_gpu": l_count, "orders_rows_per_gpu": n_orders, "customer_rows_per_gpu": n_cust,
                       "exchange_bytes_per_row": "16 (orderkey exchange) + 24 (custkey exchange)" if world > 1 else "none (single GPU)"}}
    for op in (probe_o, probe_c, build_o, build_c):
        op.close()
    bridge_o.lookup_source.close()
    bridge_c.lookup_source.close()
    for p in keep:
        if p:
            p.release()
    if dist is not None:
        dist.barrier()
    xc.close()
    return line


# ---------------------------------------------------------------------------------------------------------------------------------
def star_tables(ctx, ops, abi, world, rank, n_fact, first):
    """device-resident fact shard and this rank's range shards of the four dimensions (key, payload)"""
    lib = ctx.lib
    I64, F64 = abi.INT64, abi.FLOAT64
    nb = (n_fact + 7) // 8
    ptr = {k: ctx.malloc(n_fact * 8) for k in ("date_sk", "item_sk", "customer_sk", "store_sk", "net_paid")}
    ptr["customer_valid"], ptr["store_valid"] = ctx.malloc(nb), ctx.malloc(nb)
    both = C.c_int64()
    ctx.check(lib.tgpu_synth_store_sales(ctx.h, n_fact, first, SEED_STORE_SALES, *[C.c_void_p(ptr[k]) for k in
              ("date_sk", "item_sk", "customer_sk", "customer_valid", "store_sk", "store_valid", "net_paid")], C.byref(both)))
    from trino_b200.sharding import shard_range
    B = abi.V_BIGINT
    dims = {}
    specs = {"date": (DATE_SK0, 73049, ops.Call(abi.EX_ADD, ops.Const(1900, B), ops.Call(abi.EX_DIV, ops.Call(abi.EX_SUB, ops.Col(0, B), ops.Const(DATE_SK0, B)), ops.Const(365, B)))),
             "item": (1, 300000, ops.Call(abi.EX_ADD, ops.Call(abi.EX_MOD, ops.Col(0, B), ops.Const(1000, B)), ops.Const(1, B))),
             "store": (1, 1002, ops.Call(abi.EX_MOD, ops.Call(abi.EX_MUL, ops.Col(0, B), ops.Const(7, B)), ops.Const(100, B))),
             "customer": (1, 12_000_000, ops.Call(abi.EX_ADD, ops.Const(1920, B), ops.Call(abi.EX_MOD, ops.Col(0, B), ops.Const(70, B))))}
    keep = []
    for name, (k0, total, expr) in specs.items():
        f, cnt = shard_range(total, world, rank)
        d = ctx.malloc(max(cnt, 1) * 8)
        ctx.check(lib.tgpu_synth_sequence(ctx.h, k0 + f, cnt, C.c_void_p(d)))
        keycol = ops.DeviceColumn(I64, d, cnt)
        if cnt:
            pay = _project(ctx, ops, abi, keycol, cnt, expr)
            keep.append(pay)
            dims[name] = ops.DevicePage([keycol, pay.column(0)], cnt)
        else:
            dims[name] = ops.DevicePage([keycol, ops.DeviceColumn(I64, d, 0)], 0)
    return ptr, both.value, dims, keep


def star_pipeline(ctx, ops, abi, xc, dims, world):
    """builds (three replicated, one partitioned) and the four probe operators; returns (probes, closers)"""
    bridges = {k: ops.JoinBridge() for k in dims}
    builders, keep = [], []
    for name in ("date", "item", "store"):
        page = xc.broadcast(dims[name]) if world > 1 else None           # REPLICATED: every rank builds the whole dimension
        keep.append(page)
        b = ops.HashBuilderOperatorFactory(ctx, bridges[name], [0], [1], 1024).create_operator()
        b.add_input(page.as_device_page() if page else dims[name])
        b.finish()
        builders.append(b)
    part = xc.partitioner([0]) if world > 1 else None
    cpage = xc.partitioned(part, dims["customer"])
    keep.append(cpage)
    b = ops.HashBuilderOperatorFactory(ctx, bridges["customer"], [0], [1], 1024).create_operator()
    b.add_input(cpage.as_device_page() if cpage else dims["customer"])
    b.finish()
    builders.append(b)
    ctx.synchronize()
    J = abi.JOIN_INNER
    probes = {
        # fact: date_sk, item_sk, customer_sk, store_sk, net_paid
        "date": ops.LookupJoinOperatorFactory(ctx, bridges["date"], J, False, [0], [1, 2, 3, 4]).create_operator(),      # -> item, cust, store, paid, d_year
        "item": ops.LookupJoinOperatorFactory(ctx, bridges["item"], J, False, [0], [1, 2, 3, 4]).create_operator(),      # -> cust, store, paid, d_year, i_brand
        "store": ops.LookupJoinOperatorFactory(ctx, bridges["store"], J, False, [1], [0, 2, 3, 4]).create_operator(),    # -> cust, paid, d_year, i_brand, s_val
        "customer": ops.LookupJoinOperatorFactory(ctx, bridges["customer"], J, False, [0], [0, 1, 2, 3, 4]).create_operator(),   # ... + c_birth_year
    }
    return probes, part, (builders, bridges, keep)


def star_chunk(ctx, ops, xc, probes, part, fact_page):
    """one fact page through the four joins; returns (output page, pages to release afterwards in order)"""
    held = []
    cur = fact_page
    for name in ("date", "item", "store"):
        probes[name].add_input(cur)
        out = probes[name].get_output_device()
        held.append(out)
        if out is None:
            return None, held
        cur = out.as_device_page()
    x = xc.partitioned(part, cur) if part is not None else None
    if x is not None:
        held.append(x)
        cur = x.as_device_page()
    probes["customer"].add_input(cur)
    out = probes["customer"].get_output_device()
    return out, held


def run_star(args, ctx, rank, world, local, dist, clocks):
    from trino_b200 import abi
    from trino_b200 import operators as ops
    from trino_b200.exchange import Exchange
    from trino_b200.sharding import shard_range
    n_fact = int(2_880_000 * args.ds_sf) // 1024 * 1024                 # per GPU
    total_fact = n_fact * world
    first, _ = shard_range(total_fact, world, rank)
    ptr, both, dims, keep_dims = star_tables(ctx, ops, abi, world, rank, n_fact, first)
    xc = Exchange(ctx, dist, rank, world, local)
    probes, part, closers = star_pipeline(ctx, ops, abi, xc, dims, world)
    chunks = max(1, args.star_chunks)
    chunk_rows = (n_fact // chunks + 1023) // 1024 * 1024
    xc.create_arenas(int(chunk_rows * 1.3) * (5 * 8 + 1) + (8 << 20))
    I64, F64 = abi.INT64, abi.FLOAT64
    pages = []
    for lo in range(0, n_fact, chunk_rows):
        m = min(chunk_rows, n_fact - lo)
        cols = [ops.DeviceColumn(I64, ptr["date_sk"] + lo * 8, m), ops.DeviceColumn(I64, ptr["item_sk"] + lo * 8, m),
                ops.DeviceColumn(I64, ptr["customer_sk"] + lo * 8, m, validity=ptr["customer_valid"] + lo // 8),
                ops.DeviceColumn(I64, ptr["store_sk"] + lo * 8, m, validity=ptr["store_valid"] + lo // 8),
                ops.DeviceColumn(F64, ptr["net_paid"] + lo * 8, m)]
        pages.append(ops.DevicePage(cols, m))
    chk = {"rows": 0, "birth": 0, "cust_mod": 0}

    def step(check):
        for page in pages:
            out, held = star_chunk(ctx, ops, xc, probes, part, page)
            if out is not None:
                if check is not None:
                    check["rows"] += out.rows
                    check["birth"] = (check["birth"] + _column_sum(ctx, abi, out.column(5))) & M64
                    check["cust_mod"] = (check["cust_mod"] + _column_sum(ctx, abi, out.column(0), 70)) & M64
                out.release()
            for p in reversed(held):
                if p:
                    p.release()

    ms, launches, sample = _timed(ctx, dist, local, step, args.steps, args.warmup, clocks)
    step(chk)
    tot = _allreduce_u64(dist, local, [chk["rows"], chk["birth"], chk["cust_mod"], both])
    verify = {"rows_out": tot[0], "fact_rows_with_both_nullable_keys": tot[3], "fact_rows": total_fact,
              "birth_year_sum_matches_customer_keys": tot[1] == (tot[2] + 1920 * tot[0]) & M64}
    assert tot[0] == tot[3] and verify["birth_year_sum_matches_customer_keys"], verify
    line = {"workload": "star", "metric": "star_join_fact_rows_per_sec", "value": total_fact / (ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "int64", "data": "synthetic",
            "gpu_launches": int(launches), "clocks": sample, "verify": verify,
            "config": {"workload": f"store_sales x (date_dim, item, store: replicated builds; customer: partitioned build), synthetic TPC-DS-shaped, "
                                   f"{n_fact} fact rows per GPU in {len(pages)} pages (BASELINE.json configs[4] is SF1000 = 2.88 G rows over 8 GPUs = 360 M per GPU), "
                                   "4.5 % NULL ss_store_sk / ss_customer_sk never match",
                       "dimension_rows": {"date_dim": 73049, "item": 300000, "store": 1002, "customer": 12_000_000},
                       "exchange_bytes_per_row": "41 (5 columns + NULL byte of the key) after the three local joins" if world > 1 else "none (single GPU)"}}
    for p in probes.values():
        p.close()
    builders, bridges, keep = closers
    for b in builders:
        b.close()
    for br in bridges.values():
        br.lookup_source.close()
    for p in keep:
        if p:
            p.release()
    if dist is not None:
        dist.barrier()
    xc.close()
    return line


# ---------------------------------------------------------------------------------------------------------------------------------
def run_q1(args, ctx, rank, world, local, dist, clocks):
    import sys
    from trino_b200 import abi
    from trino_b200 import operators as ops
    from trino_b200.exchange import Exchange
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests"))
    from q1 import q1_aggregators, q1_factory
    lib = ctx.lib
    n = int(6_000_000 * args.q1_sf)                       # per GPU
    spec = [(abi.INT32, 4), (abi.INT8, 1), (abi.INT8, 1), (abi.FLOAT64, 8), (abi.FLOAT64, 8), (abi.FLOAT64, 8), (abi.FLOAT64, 8)]
    ptrs = [ctx.malloc(n * sz) for _, sz in spec]
    ctx.check(lib.tgpu_synth_lineitem_q1(ctx.h, n, rank * n, SEED_LINEITEM, *[C.c_void_p(p) for p in ptrs]))
    page = ops.DevicePage([ops.DeviceColumn(t, p, n) for (t, _), p in zip(spec, ptrs)], n)
    xc = Exchange(ctx, dist, rank, world, local)
    part = xc.partitioner([0, 1]) if world > 1 else None
    partial_f = q1_factory(ctx, fused=True, step=abi.STEP_PARTIAL)
    # intermediate layout (include/trino_gpu.h): keys 0,1; sum -> 1 column; avg -> (count, sum); count(*) -> 1 column
    A = ops.Aggregator
    final_aggs = [A(abi.AGG_SUM, 2), A(abi.AGG_SUM, 3), A(abi.AGG_SUM, 4), A(abi.AGG_SUM, 5), A(abi.AGG_AVG, 6), A(abi.AGG_AVG, 8), A(abi.AGG_AVG, 10), A(abi.AGG_COUNT_STAR, 12)]
    final_f = ops.HashAggregationOperatorFactory(ctx, [0, 1], abi.STEP_FINAL, final_aggs, expected_groups=16)
    result = {"rows": None}

    def step(check):
        p = partial_f.create_operator()
        p.add_input(page)
        p.finish()
        inter = p.get_output_device()
        recv = xc.partitioned(part, inter.as_device_page()) if world > 1 else None
        f = final_f.create_operator()
        src = recv if recv else inter
        if src.rows:
            f.add_input(src.as_device_page())
        f.finish()
        out = f.get_output()
        if check is not None:
            check["rows"] = out.rows() if out is not None else []
        f.close()
        if recv:
            recv.release()
        inter.release()
        p.close()

    ms, launches, sample = _timed(ctx, dist, local, step, args.steps, args.warmup, clocks)
    step(result)
    # verification: the distributed PARTIAL -> exchange -> FINAL result equals the sum of every rank's own SINGLE-step result
    single = q1_factory(ctx, fused=True).create_operator()
    single.add_input(page)
    single.finish()
    own = single.get_output().rows()
    single.close()
    local_cnt = sum(int(r[-1]) for r in own)
    local_qty = sum(float(r[2]) for r in own)
    dist_cnt = sum(int(r[-1]) for r in result["rows"])
    dist_qty = sum(float(r[2]) for r in result["rows"])
    groups = len(result["rows"])
    if dist is not None:
        import torch
        t = torch.tensor([local_cnt, dist_cnt, groups], dtype=torch.int64, device=f"cuda:{local}")
        q = torch.tensor([local_qty, dist_qty], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t)
        dist.all_reduce(q)
        local_cnt, dist_cnt, groups = [int(x) for x in t.tolist()]
        local_qty, dist_qty = [float(x) for x in q.tolist()]
    verify = {"groups": groups, "count_order_sum": dist_cnt, "count_matches_single_step": dist_cnt == local_cnt,
              "sum_qty_relative_error": abs(dist_qty - local_qty) / max(abs(local_qty), 1e-300), "tolerance": 1e-6}
    assert verify["count_matches_single_step"] and verify["sum_qty_relative_error"] <= 1e-6 and groups == 4, verify
    total = n * world
    line = {"workload": "q1", "metric": "groupby_input_rows_per_sec", "value": total / (ms * 1e-3), "unit": "rows/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "gpu_launches": int(launches), "clocks": sample, "verify": verify,
            "config": {"workload": f"TPC-H Q1, synthetic SF{args.q1_sf:g} lineitem per GPU: fused filter + project + PARTIAL aggregation per rank, hash exchange of the "
                                   "intermediate rows on (returnflag, linestatus), FINAL aggregation (HashAggregationOperator.java:366-393)", "rows_per_gpu": n}}
    for p in ptrs:
        ctx.free(p)
    if dist is not None:
        dist.barrier()
    xc.close()
    return line


RUNNERS = {"q3way": run_q3way, "star": run_star, "q1": run_q1}


def main(args, ClockSampler, dist_env):
    rank, world, local = dist_env()
    from trino_b200 import operators as ops
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = ops.Context(local)
    line = RUNNERS[args.workload](args, ctx, rank, world, local, dist, ClockSampler)
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    ctx.close()
    return 0
