"""Long DECIMAL (TGPU_INT128 / Int128ArrayBlock) through the GPU operators: the reference's DecimalSumAggregation state tests, sums against
exact integer arithmetic, 128-bit group-by / join / partition keys against the oracle, and pass-through of 128-bit channels."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import oracle_lib as o  # noqa: E402
from helpers import gpu_join_rows, oracle_join_rows  # noqa: E402
from trino_b200 import abi  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402
from trino_b200.page import Block, Page  # noqa: E402

pytestmark = pytest.mark.gpu
A = ops.Aggregator
TWO = 2


def _rows(ctx, factory, pages):
    op = factory.create_operator()
    out = ops.drive(op, pages)
    op.close()
    return [r for p in out for r in p.rows()]


def _partial_state(ctx, batches, short=False):
    """(sum, overflow) of one group after the PARTIAL step saw `batches` (one page each)"""
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(abi.AGG_SUM_DECIMAL, 1)], 16)
    mk = Block.bigint if short else Block.int128
    rows = _rows(ctx, f, [Page(Block.bigint([7] * len(b)), mk(b)) for b in batches])
    assert len(rows) == 1 and rows[0][0] == 7
    return rows[0][1], rows[0][2]


MASK128 = (1 << 128) - 1


def _wrap128(v):
    """exact integer -> the signed value its low 128 bits read as (Block.get of an INT128 block returns signed values)"""
    return ((v + (1 << 127)) & MASK128) - (1 << 127)


def test_reference_state_cases(ctx):
    # T/operator/aggregation/TestDecimalSumAggregation.java:36-122 through the PARTIAL step (state = INT128 sum, BIGINT overflow)
    assert _partial_state(ctx, [[TWO**126]]) == (TWO**126, 0)
    s, ov = _partial_state(ctx, [[TWO**126, TWO**126]])
    assert ov == 1 and s == -(1 << 127)                                                # Int128.valueOf(1L << 63, 0)
    s, ov = _partial_state(ctx, [[-(TWO**126), -(TWO**126)]])
    assert ov == 0 and s == -(1 << 127)
    s, ov = _partial_state(ctx, [[TWO**126, TWO**126, TWO**125], [-(TWO**126)] * 3])   # testUnderflowAfterOverflow, across two pages
    assert ov == 0 and s == -(TWO**125)
    s, ov = _partial_state(ctx, [[TWO**125, TWO**126], [TWO**125, TWO**126]])          # testCombineOverflow
    assert ov == 1 and s & MASK128 == 0xC000000000000000 << 64
    s, ov = _partial_state(ctx, [[-(TWO**125), -(TWO**126)], [-(TWO**125), -(TWO**126)]])
    assert ov == -1 and s & MASK128 == 0x4000000000000000 << 64


def test_decimal_overflow_on_output(ctx):
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, [A(abi.AGG_SUM_DECIMAL, 1)], 16)
    with pytest.raises(abi.TrinoGpuError) as e:
        _rows(ctx, f, [Page(Block.bigint([1, 1]), Block.int128([TWO**126, TWO**126]))])
    assert e.value.code == abi.ERR_NUMERIC_VALUE_OUT_OF_RANGE
    with pytest.raises(abi.TrinoGpuError):
        _rows(ctx, f, [Page(Block.bigint([1, 1]), Block.int128([10**38 - 1, 1]))])
    assert _rows(ctx, f, [Page(Block.bigint([1, 1, 2]), Block.int128([10**38 - 1, None, None]))]) == [(1, 10**38 - 1), (2, None)]


def _decimal_pages(rng, card, sizes):
    pages = []
    for n in sizes:
        keys = rng.integers(0, card, n)
        longs = [None if rng.random() < 0.05 else int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**48)) for _ in range(n)]      # (group sums stay inside +-10^38)
        shorts = rng.integers(-10**17, 10**17, n)
        pages.append(Page(Block.bigint(keys), Block.int128(longs), Block.bigint(shorts, rng.random(n) < 0.05), Block.boolean(rng.random(n) < 0.5)))
    return pages


def _want_sums(pages, key_of=lambda r: r[0]):
    order, sums = [], {}
    for p in pages:
        for r in p.rows():
            k = key_of(r)
            if k not in sums:
                sums[k] = [None, None, None, 0]
                order.append(k)
            st = sums[k]
            if r[1] is not None:
                st[0] = (st[0] or 0) + r[1]
            if r[2] is not None:
                st[1] = (st[1] or 0) + r[2]
            if r[1] is not None and r[3]:
                st[2] = (st[2] or 0) + r[1]
            st[3] += 1
    return [(k, sums[k][0], sums[k][1], sums[k][2], sums[k][3]) for k in order]        # (no sum here leaves +-10^38)


AGGS = [A(abi.AGG_SUM_DECIMAL, 1), A(abi.AGG_SUM_DECIMAL, 2), A(abi.AGG_SUM_DECIMAL, 1, 3), A(abi.AGG_COUNT_STAR)]


@pytest.mark.parametrize("card", [6, 5000])
def test_decimal_sums_match_exact_arithmetic(ctx, card):
    # few groups: the shared-memory path; thousands: the global-table path.  Long and short decimal inputs, NULLs, a mask channel
    rng = np.random.default_rng(31)
    pages = _decimal_pages(rng, card, (3000, 1, 4000))
    single = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, AGGS, 16), pages)
    assert single == _want_sums(pages)
    # PARTIAL per page -> FINAL over (INT128 sum, BIGINT overflow) state pairs
    pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, AGGS, 16)
    partial = []
    for p in pages:
        op = pf.create_operator()
        partial += ops.drive(op, [p])
        op.close()
    assert partial[0].channel_count == 1 + 2 + 2 + 2 + 1
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, [A(abi.AGG_SUM_DECIMAL, 1), A(abi.AGG_SUM_DECIMAL, 3), A(abi.AGG_SUM_DECIMAL, 5), A(abi.AGG_COUNT_STAR, 7)], 16)
    assert _rows(ctx, ff, partial) == single


def test_int128_group_by_keys(ctx):
    rng = np.random.default_rng(32)
    domain = [int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**62)) for _ in range(300)] + [0, -1, 2**127 - 1, -(2**127)]
    pages = []
    for n in (2000, 5, 3000):
        ks = [None if rng.random() < 0.02 else domain[int(rng.integers(0, len(domain)))] for _ in range(n)]
        small = [None if rng.random() < 0.1 else int(rng.integers(-2**62, 2**62)) * 2**40 for _ in range(n)]
        pages.append(Page(Block.int128(ks), Block.int128(small), Block.bigint(rng.integers(-5, 5, n)), Block.boolean(np.ones(n, bool))))
    got = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, [A(abi.AGG_COUNT_STAR), A(abi.AGG_SUM, 2), A(abi.AGG_SUM_DECIMAL, 1)], 16), pages)
    order, agg = [], {}
    for p in pages:
        for r in p.rows():
            if r[0] not in agg:
                agg[r[0]] = [0, 0, None]
                order.append(r[0])
            a = agg[r[0]]
            a[0] += 1
            a[1] += r[2]
            if r[1] is not None:
                a[2] = (a[2] or 0) + r[1]
    want = [(k, agg[k][0], agg[k][1], agg[k][2]) for k in order]
    assert got == want                                  # first-seen order, the NULL key is a group of its own
    # a composite key of a 128-bit and a 64-bit channel
    got = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [2, 0], abi.STEP_SINGLE, [A(abi.AGG_COUNT_STAR)], 16), pages)
    order, cnt = [], {}
    for p in pages:
        for r in p.rows():
            k = (r[2], r[0])
            if k not in cnt:
                cnt[k] = 0
                order.append(k)
            cnt[k] += 1
    assert got == [(k[0], k[1], cnt[k]) for k in order]


def test_int128_partition_join_and_pass_through(ctx):
    rng = np.random.default_rng(33)
    n = 5000
    keys = [int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**40)) for _ in range(n)]
    page = Page(Block.int128([None if z else k for k, z in zip(keys, rng.random(n) < 0.03)]), Block.bigint(np.arange(n)), Block.varchar(["r%d" % i for i in range(n)]))
    # PagePartitioner: LongDecimalType.hash through InterpretedHashGenerator
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], 8).create_operator()
    assert (op.get_partitions(page) == o.partition_ids(page, [0], 8)).all()
    op.add_input(page)
    got = {}
    while True:
        r = op.get_output_with_partition()
        if r is None:
            break
        got[r[0]] = r[1].rows()
    op.close()
    lists, _ = o.partition_positions(page, [0], 8, None, 8, -1, False, False)
    rows = page.rows()
    assert got == {p: [rows[i] for i in l] for p, l in enumerate(lists) if len(l)}
    # join on the 128-bit key (generic path: row hash + full-key verification), 128-bit payload on both sides
    build = Page(Block.int128(keys[:2000] + [keys[5], None]), Block.int128(keys[:2000] + [1, 2]))
    probe = Page(Block.int128(keys[1000:3000] + [None]), Block.bigint(np.arange(2001)))
    rows = gpu_join_rows(ctx, [build], [probe], 0, 0, [0, 1], [1], abi.JOIN_INNER, False)
    assert rows == oracle_join_rows(build, probe, 0, 0, [0, 1], [1], abi.JOIN_INNER, False) and len(rows) >= 1000
    # a BIGINT-key join whose build side carries a 128-bit payload: the fused probe + gather fast path moves at most 8-byte payloads, so this
    # shape must take the general path (count / scan / gather) and still give the oracle's rows
    bkeys = np.arange(3000, dtype=np.int64) * 7
    build = Page(Block.bigint(bkeys), Block.int128([int(k) * 10**20 for k in bkeys]))
    probe = Page(Block.bigint(rng.integers(0, 21000, 4096)), Block.bigint(np.arange(4096)))
    rows = gpu_join_rows(ctx, [build], [probe], 0, 0, [0, 1], [1], abi.JOIN_INNER, False)
    assert rows == oracle_join_rows(build, probe, 0, 0, [0, 1], [1], abi.JOIN_INNER, False) and len(rows) > 100
    # GroupByHash.getGroupIds over a 128-bit key
    gb = ops.GroupByHash(ctx, [0])
    og = o.GroupByHash(0, 16)
    kp = Page(Block.int128([None if i % 50 == 0 else keys[i % 300] for i in range(4000)]))
    assert (gb.get_group_ids(kp) == og.get_group_ids(kp, [0])).all() and gb.get_group_count() == og.group_count()
    gb.close(); og.close()
    # FilterAndProject: a filter on a BIGINT channel, the 128-bit channel passes through
    prog = ops.PageProcessorProgram(ops.Call(abi.EX_LT, ops.Col(1, abi.V_BIGINT), ops.Const(1234, abi.V_BIGINT)), [0, 1])
    fp = ops.FilterAndProjectOperatorFactory(ctx, prog).create_operator()
    out = [r for p in ops.drive(fp, [page]) for r in p.rows()]
    fp.close()
    assert out == [(r[0], r[1]) for r in page.rows() if r[1] < 1234]


def test_decimal_sum_through_a_skipped_partial_builder(ctx):
    rng = np.random.default_rng(34)
    pages = _decimal_pages(rng, 50, (500, 700))
    controller = ops.PartialAggregationController(ctx.lib, 1 << 40, 0.0)
    controller.on_flush(1 << 41, 10, 10)
    pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, AGGS, 16, partial_aggregation_controller=controller)
    op = pf.create_operator()
    partial = ops.drive(op, pages)
    assert op.rows_with_partial_aggregation_disabled() == 1200
    op.close()
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, [A(abi.AGG_SUM_DECIMAL, 1), A(abi.AGG_SUM_DECIMAL, 3), A(abi.AGG_SUM_DECIMAL, 5), A(abi.AGG_COUNT_STAR, 7)], 16)
    assert _rows(ctx, ff, partial) == _want_sums(pages)
    controller.close()


def _avg(ctx, longs=None, shorts=None):
    """avg(decimal) of one group through the SINGLE step"""
    if longs is not None:
        page = Page(Block.bigint([3] * len(longs)), Block.int128(longs))
    else:
        page = Page(Block.bigint([3] * len(shorts)), Block.bigint(shorts))
    rows = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, [A(abi.AGG_AVG_DECIMAL, 1)], 16), [page])
    assert len(rows) == 1
    return rows[0][1]


def test_reference_decimal_average_cases(ctx):
    # T/operator/aggregation/TestDecimalAverageAggregation.java:46-216 through the operator
    MIN = -(10**38 - 1)
    assert _avg(ctx, [TWO**126, TWO**126]) == TWO**126                                  # testOverflow (overflow != 0 branch)
    assert _avg(ctx, [MIN, MIN]) == MIN                                                 # testUnderflow
    assert _avg(ctx, [TWO**126, TWO**126, TWO**125] + [-(TWO**126)] * 3) == -((TWO**125) // 6)
    for numbers, want in (([10**37, 0], 5 * 10**36), ([2, 1], 2), ([0, 1], 1), ([-2, -1], -2), ([-1, 0], -1), ([-1, 0, 0], 0), ([-2, 0, 0], -1),
                          ([-2, 0], -1), ([200, 100], 150), ([0, 100], 50), ([-200, -100], -150), ([-100, 0], -50)):
        assert _avg(ctx, numbers) == want, numbers
        if all(abs(x) < 2**63 for x in numbers):
            assert _avg(ctx, None, numbers) == want, numbers                           # the same values as a short decimal: a BIGINT result
    assert _avg(ctx, [None, None]) is None


def test_decimal_average_partial_final_and_skipped_builders(ctx):
    rng = np.random.default_rng(35)
    pages = _decimal_pages(rng, 40, (2500, 1, 3000))
    aggs = [A(abi.AGG_AVG_DECIMAL, 1), A(abi.AGG_AVG_DECIMAL, 2), A(abi.AGG_AVG_DECIMAL, 1, 3)]

    def half_up(total, n):
        q, r = divmod(abs(total), n)
        return (-1 if total < 0 else 1) * (q + (1 if 2 * r >= n else 0))

    order, acc = [], {}
    for p in pages:
        for r in p.rows():
            if r[0] not in acc:
                acc[r[0]] = [[0, 0], [0, 0], [0, 0]]
                order.append(r[0])
            a = acc[r[0]]
            if r[1] is not None:
                a[0][0] += r[1]; a[0][1] += 1
            if r[2] is not None:
                a[1][0] += r[2]; a[1][1] += 1
            if r[1] is not None and r[3]:
                a[2][0] += r[1]; a[2][1] += 1
    want = [(k,) + tuple(None if n == 0 else half_up(t, n) for t, n in acc[k]) for k in order]
    single = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, aggs, 16), pages)
    assert single == want
    finals = [A(abi.AGG_AVG_DECIMAL, 1, result_type=abi.INT128), A(abi.AGG_AVG_DECIMAL, 4, result_type=abi.INT64), A(abi.AGG_AVG_DECIMAL, 7, result_type=abi.INT128)]
    for controller_off in (False, True):
        controller = ops.PartialAggregationController(ctx.lib, 1 << 40, 0.0)
        if controller_off:
            controller.on_flush(1 << 41, 10, 10)
        pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, aggs, 16, partial_aggregation_controller=controller)
        partial = []
        for p in pages:
            op = pf.create_operator()
            partial += ops.drive(op, [p])
            op.close()
        assert partial[0].channel_count == 1 + 3 * 3
        ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, finals, 16)
        assert _rows(ctx, ff, partial) == want, controller_off
        controller.close()


def test_reference_typed_states_between_partial_and_final(ctx):
    # a GPU PARTIAL feeding a FINAL through the reference's own state types: avg(double) as ROW(BIGINT, DOUBLE), the decimal states as
    # VARBINARY (LongDecimalWithOverflow[AndLong]StateSerializer); channel numbers on the FINAL factory are the Java plan's
    from trino_b200.page import RowBlock
    rng = np.random.default_rng(36)
    pages = _decimal_pages(rng, 30, (2000, 2500))
    aggs = [A(abi.AGG_SUM_DECIMAL, 1), A(abi.AGG_AVG, 2), A(abi.AGG_AVG_DECIMAL, 1), A(abi.AGG_COUNT_STAR)]
    single = _rows(ctx, ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, aggs, 16), pages)
    pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, aggs, 16, row_typed_states=True)
    partial = []
    for p in pages:
        op = pf.create_operator()
        partial += ops.drive(op, [p])
        op.close()
    first = partial[0]
    assert first.channel_count == 5 and first.get_block(1).type == abi.UTF8 and isinstance(first.get_block(2), RowBlock) and first.get_block(3).type == abi.UTF8
    finals = [A(abi.AGG_SUM_DECIMAL, 1), A(abi.AGG_AVG, 2), A(abi.AGG_AVG_DECIMAL, 3, result_type=abi.INT128), A(abi.AGG_COUNT_STAR, 4)]
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, finals, 16, row_typed_states=True)
    from helpers import rows_equal
    assert rows_equal(_rows(ctx, ff, partial), single, rel=1e-9)
