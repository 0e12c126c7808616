"""GPU parity of the bench_workloads.py pipelines at reduced scale, single GPU: generators vs the oracle's, the star join (replicated +
partitioned builds, NULL fact keys never match: M/operator/join/unspilled/JoinProbe.java:154-171) and the 3-way join against a chain
of oracle joins, the PARTIAL -> FINAL Q1 against the oracle's Q1, and the broadcast exchange's single-GPU form."""
import ctypes as C
import types

import numpy as np
import pytest

import bench_workloads as bw
import oracle_lib as o
from helpers import oracle_star_rows
from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.exchange import Exchange
from trino_b200.page import Block, Page

pytestmark = pytest.mark.gpu


def test_new_generators_identical(ctx):
    lib = ctx.lib
    n_orders, n_cust = 50_003, 15_000
    d = ctx.malloc(n_orders * 8)
    for shuffle in (0, 1):
        ctx.check(lib.tgpu_synth_orders_custkeys(ctx.h, n_orders, 100, 4000, 0x7C02, shuffle, n_cust, 0x7C03, C.c_void_p(d)))
        got = ctx.to_host(d, np.int64, 4000)
        assert (got == o.synth_orders_custkeys(n_orders, 100, 4000, 0x7C02, shuffle, n_cust, 0x7C03)).all()
        assert (got % 3 != 0).all() and got.min() >= 1 and got.max() <= n_cust        # only customers with custkey % 3 != 0 have orders
    ctx.check(lib.tgpu_synth_sequence(ctx.h, 42, 1000, C.c_void_p(d)))
    assert (ctx.to_host(d, np.int64, 1000) == np.arange(42, 1042)).all()
    ctx.free(d)
    n = 100_003
    want, both = o.synth_store_sales(n, 777, 0xD501)
    names = ("date_sk", "item_sk", "customer_sk", "customer_valid", "store_sk", "store_valid", "net_paid")
    ptrs = {k: ctx.malloc(max(want[k].nbytes, 16)) for k in names}
    got_both = C.c_int64()
    ctx.check(lib.tgpu_synth_store_sales(ctx.h, n, 777, 0xD501, *[C.c_void_p(ptrs[k]) for k in names], C.byref(got_both)))
    assert got_both.value == both
    for k in names:
        got = ctx.to_host(ptrs[k], want[k].dtype, len(want[k]))
        assert (got.view(np.uint8) == want[k].view(np.uint8)).all(), k
        ctx.free(ptrs[k])
    nulls = 1.0 - np.unpackbits(want["customer_valid"], bitorder="little")[:n].mean()
    assert 0.035 < nulls < 0.055                                                       # 4.5 % NULL customer keys


def test_star_join_matches_oracle_chain(ctx):
    n = 200 * 1024
    ptr, both, dims, keep = bw.star_tables(ctx, ops, abi, 1, 0, n, 5000)
    xc = Exchange(ctx, None, 0, 1, 0)
    probes, part, closers = bw.star_pipeline(ctx, ops, abi, xc, dims, 1)
    I64, F64 = abi.INT64, abi.FLOAT64
    got_rows = []
    half = n // 2
    for lo, m in ((0, half), (half, n - half)):                       # two pages: the operators are re-used across pages
        page = ops.DevicePage([ops.DeviceColumn(I64, ptr["date_sk"] + lo * 8, m), ops.DeviceColumn(I64, ptr["item_sk"] + lo * 8, m),
                               ops.DeviceColumn(I64, ptr["customer_sk"] + lo * 8, m, validity=ptr["customer_valid"] + lo // 8),
                               ops.DeviceColumn(I64, ptr["store_sk"] + lo * 8, m, validity=ptr["store_valid"] + lo // 8),
                               ops.DeviceColumn(F64, ptr["net_paid"] + lo * 8, m)], m)
        out, held = bw.star_chunk(ctx, ops, xc, probes, part, page)
        got_rows += out.to_host().rows()
        out.release()
        for p in reversed(held):
            p.release()
    want, both_o = oracle_star_rows(n, 5000)
    assert both == both_o
    assert len(want) == both and len(got_rows) == both
    assert got_rows == want                                            # same rows in probe order
    for p in probes.values():
        p.close()
    builders, bridges, kept = closers
    for b in builders:
        b.close()
    for br in bridges.values():
        br.lookup_source.close()
    xc.close()


def _args(**kw):
    base = dict(sf=0.05, ds_sf=0.1, star_chunks=3, q1_sf=0.02, steps=1, warmup=1)
    base.update(kw)
    return types.SimpleNamespace(**base)


class _NoClocks:
    def __init__(self, index):
        pass

    def start(self):
        pass

    def stop(self):
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}


@pytest.mark.parametrize("workload", ["q3way", "star", "q1"])
def test_workload_runners_verify_themselves_single_gpu(ctx, workload):
    # the bench entry points at toy scale: their closed-form verification passes (row counts, checksums, oracle spot checks) must hold
    line = bw.RUNNERS[workload](_args(), ctx, 0, 1, 0, None, _NoClocks)
    assert line["n_gpus"] == 1 and line["value"] > 0
    v = line["verify"]
    if workload == "q3way":
        assert v["rows_out"] == v["rows_in"] and v["nation_sum_matches_custkeys"] and v["oracle_spot_check_rows_per_rank"] > 0
    elif workload == "star":
        assert v["rows_out"] == v["fact_rows_with_both_nullable_keys"] < v["fact_rows"]
    else:
        assert v["groups"] == 4 and v["count_matches_single_step"]


def test_q3way_matches_oracle_chain(ctx):
    n_orders, n_cust = 30_000, 3_000
    okeys = o.synth_orders_keys(n_orders, 0, n_orders, 0x7C02, True)
    ocust = o.synth_orders_custkeys(n_orders, 0, n_orders, 0x7C02, True, n_cust, 0x7C03)
    rows = o.synth_lineitem_rows(n_orders)
    lkeys = o.synth_lineitem_keys(n_orders, 0, rows, 0x7C01, False)
    ckeys = np.arange(1, n_cust + 1)
    orders, customer = Page(Block.bigint(okeys), Block.bigint(ocust)), Page(Block.bigint(ckeys), Block.bigint(ckeys % 25))
    lineitem = Page(Block.bigint(lkeys), Block.double(lkeys * 0.5))
    from helpers import gpu_join_rows, oracle_join_rows
    mid = gpu_join_rows(ctx, [orders], [lineitem], 0, 0, [0, 1], [1], abi.JOIN_INNER, False)
    assert mid == oracle_join_rows(orders, lineitem, 0, 0, [0, 1], [1], abi.JOIN_INNER, False)
    mid_page = Page(Block.bigint([r[0] for r in mid]), Block.double([r[1] for r in mid]), Block.bigint([r[2] for r in mid]))
    out = gpu_join_rows(ctx, [customer], [mid_page], 0, 2, [0, 1, 2], [1], abi.JOIN_INNER, False)
    assert out == oracle_join_rows(customer, mid_page, 0, 2, [0, 1, 2], [1], abi.JOIN_INNER, False)
    assert len(out) == rows and all(r[3] == r[2] % 25 for r in out)


def test_broadcast_exchange_single_gpu_is_a_copy(ctx):
    xc = Exchange(ctx, None, 0, 1, 0)
    from trino_b200.page import AbiPage
    page = Page(Block.bigint([5, None, 7]), Block.double([1.5, 2.5, None]), Block.integer([1, 2, 3]))
    got = xc.broadcast(AbiPage(page))
    assert got.to_host().rows() == page.rows()
    got.release()
    text = Page(Block.varchar(["a", None, "ccc"]), Block.bigint([1, 2, 3]))
    got = xc.broadcast(AbiPage(text))
    assert got.to_host().rows() == text.rows()
    got.release()
