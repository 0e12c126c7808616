"""The device generators and the oracle generators produce identical columns."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as o

pytestmark = pytest.mark.gpu


def test_keys_identical(ctx):
    n_orders = 100_003
    for shuffle in (0, 1):
        d = ctx.malloc(n_orders * 8)
        ctx.check(ctx.lib.tgpu_synth_orders_keys(ctx.h, n_orders, 0, n_orders, 0x7C02, shuffle, C.c_void_p(d)))
        assert (ctx.to_host(d, np.int64, n_orders) == o.synth_orders_keys(n_orders, 0, n_orders, 0x7C02, shuffle)).all()
        ctx.free(d)
        rows = ctx.lib.tgpu_synth_lineitem_rows(n_orders)
        assert rows == o.synth_lineitem_rows(n_orders)
        d = ctx.malloc(5000 * 8)
        ctx.check(ctx.lib.tgpu_synth_lineitem_keys(ctx.h, n_orders, 1234, 5000, 0x7C01, shuffle, C.c_void_p(d)))
        assert (ctx.to_host(d, np.int64, 5000) == o.synth_lineitem_keys(n_orders, 1234, 5000, 0x7C01, shuffle)).all()
        ctx.free(d)


def test_q1_columns_identical(ctx):
    n = 200_000
    want = o.synth_lineitem_q1(n, 77, 0x7C01)
    spec = [("shipdate", np.int32), ("returnflag", np.int8), ("linestatus", np.int8), ("quantity", np.float64),
            ("extendedprice", np.float64), ("discount", np.float64), ("tax", np.float64)]
    ptrs = [ctx.malloc(n * np.dtype(t).itemsize) for _, t in spec]
    ctx.check(ctx.lib.tgpu_synth_lineitem_q1(ctx.h, n, 77, 0x7C01, *[C.c_void_p(p) for p in ptrs]))
    for (name, t), p in zip(spec, ptrs):
        got = ctx.to_host(p, t, n)
        assert (got.view(np.uint8) == want[name].view(np.uint8)).all(), name   # bit-exact incl. the double columns
        ctx.free(p)
