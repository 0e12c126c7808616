"""TPC-H Q1 through the GPU operators (fused ScanFilterAndProject -> HashAggregation), shared by tests, smoke and bench."""
import numpy as np

from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, Page

CUTOFF = 10471   # DATE '1998-12-01' - INTERVAL '90' DAY = 1998-09-02 (R/testing/trino-benchmark-queries/.../tpch/q01.sql)


def q1_program():
    """filter l_shipdate <= cutoff; projections: returnflag, linestatus, quantity, extendedprice,
    extendedprice*(1-discount), extendedprice*(1-discount)*(1+tax), discount.
    Input channels: 0 shipdate, 1 returnflag, 2 linestatus, 3 quantity, 4 extendedprice, 5 discount, 6 tax."""
    D = abi.V_DOUBLE
    ep, disc, tax = ops.Col(4, D), ops.Col(5, D), ops.Col(6, D)
    one = ops.Const(1.0, D)
    disc_price = ops.Call(abi.EX_MUL, ep, ops.Call(abi.EX_SUB, one, disc))
    charge = ops.Call(abi.EX_MUL, ops.Call(abi.EX_MUL, ep, ops.Call(abi.EX_SUB, one, disc)), ops.Call(abi.EX_ADD, one, tax))
    flt = ops.Call(abi.EX_LE, ops.Col(0, abi.V_BIGINT), ops.Const(CUTOFF, abi.V_BIGINT))
    return ops.PageProcessorProgram(flt, [1, 2, 3, 4, disc_price, charge, 5])


def q1_aggregators():
    # sum(qty), sum(ep), sum(disc_price), sum(charge), avg(qty), avg(ep), avg(disc), count(*)
    A = ops.Aggregator
    return [A(abi.AGG_SUM, 2), A(abi.AGG_SUM, 3), A(abi.AGG_SUM, 4), A(abi.AGG_SUM, 5),
            A(abi.AGG_AVG, 2), A(abi.AGG_AVG, 3), A(abi.AGG_AVG, 6), A(abi.AGG_COUNT_STAR)]


def q1_factory(ctx, fused=True, step=abi.STEP_SINGLE):
    if fused:
        return ops.HashAggregationOperatorFactory(ctx, [0, 1], step, q1_aggregators(), expected_groups=16, pre=q1_program())
    return ops.HashAggregationOperatorFactory(ctx, [0, 1], step, q1_aggregators(), expected_groups=16)


def q1_host_page(cols, lo=0, hi=None):
    s = slice(lo, hi)
    return Page(Block.integer(cols["shipdate"][s]), Block.tinyint(cols["returnflag"][s]), Block.tinyint(cols["linestatus"][s]),
                Block.double(cols["quantity"][s]), Block.double(cols["extendedprice"][s]), Block.double(cols["discount"][s]),
                Block.double(cols["tax"][s]))


def q1_gpu_rows(ctx, cols, cutoff=CUTOFF, page_rows=None, fused=True):
    """rows (returnflag, linestatus, 8 aggregates) in group-id order"""
    assert cutoff == CUTOFF
    n = len(cols["shipdate"])
    page_rows = page_rows or n
    pages = [q1_host_page(cols, lo, min(n, lo + page_rows)) for lo in range(0, n, page_rows)]
    if fused:
        op = q1_factory(ctx, True).create_operator()
        out = ops.drive(op, pages)
        op.close()
    else:
        fp = ops.FilterAndProjectOperatorFactory(ctx, q1_program()).create_operator()
        agg = q1_factory(ctx, False).create_operator()
        for p in pages:
            fp.add_input(p)
            while True:
                o = fp.get_output_device()
                if o is None:
                    break
                agg.add_input(o)
                o.release()
        agg.finish()
        out = []
        while not agg.is_finished():
            o = agg.get_output()
            if o is not None:
                out.append(o)
        fp.close()
        agg.close()
    rows = []
    for page in out:
        for r in page.rows():
            rows.append((chr(r[0]), chr(r[1])) + tuple(r[2:]))
    return rows
