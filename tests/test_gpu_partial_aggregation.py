"""Adaptive partial aggregation on the GPU: the reference's own operator tests (T/operator/TestHashAggregationOperator.java:784-913)
through the GPU HashAggregationOperator, skipped-builder pages against the oracle restatement of SkipAggregationBuilder, and the
end-to-end property that PARTIAL (aggregated and skipped pages mixed) -> FINAL equals a SINGLE aggregation."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import oracle_lib as o  # noqa: E402
import partial_aggregation as pa  # noqa: E402
from helpers import oracle_agg_rows, rows_equal  # noqa: E402
from q1 import CUTOFF, q1_aggregators, q1_host_page, q1_program  # noqa: E402
from trino_b200 import abi  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402
from trino_b200.page import Block, Page, RunLengthEncodedBlock  # noqa: E402

pytestmark = pytest.mark.gpu
A = ops.Aggregator


def _longs(*v):
    return Page(Block.bigint(list(v)))


def _repeated(value, n):
    return Page(RunLengthEncodedBlock(Block.bigint([value]), n))


def _run(factory, pages):
    op = factory.create_operator()
    out = ops.drive(op, pages)
    skipped = op.rows_with_partial_aggregation_disabled()
    op.close()
    return [p.rows() for p in out], skipped


def test_reference_adaptive_partial_aggregation(ctx):
    controller = ops.PartialAggregationController(ctx.lib, 1, 0.8)
    # maxPartialMemory of one byte: the operator flushes after every page and the controller reacts to every flush
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(abi.AGG_MIN, 0)], 100, max_partial_memory=1,
                                           partial_aggregation_controller=controller)
    assert not controller.is_partial_aggregation_disabled()
    got, _ = _run(f, [_longs(0, 1, 2, 3, 4, 5, 6, 7, 8, 8), _repeated(1, 10)])
    assert got == [[(i, i) for i in range(9)], [(1, 1)] * 10]          # the last position was aggregated; the second page passes raw
    assert controller.is_partial_aggregation_disabled()
    got, _ = _run(f, [_repeated(1, 10), _repeated(2, 10)])
    assert got == [[(1, 1)] * 10, [(2, 2)] * 10]
    for i in range(1, 5):
        got, _ = _run(f, [_longs(*range(9))])
        assert got == [[(k, k) for k in range(9)]]
        assert controller.is_partial_aggregation_disabled() == (i <= 3)
    controller.on_flush(1_000_000, 1_000_000, None)                     # a late flush from a disabled builder
    got, _ = _run(f, [_repeated(1, 100), _repeated(2, 100)])
    assert got == [[(1, 1)], [(2, 2)]]
    assert not controller.is_partial_aggregation_disabled()
    controller.close()


def test_reference_adaptive_partial_aggregation_triggered_only_on_flush(ctx):
    controller = ops.PartialAggregationController(ctx.lib, 1, 0.8)
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(abi.AGG_MIN, 0)], 10, max_partial_memory=16 << 20,
                                           partial_aggregation_controller=controller)
    got, skipped = _run(f, [_longs(*range(10)), _repeated(1, 2)])
    assert got == [[(k, k) for k in range(10)]]                         # the second page is squashed into the first
    assert controller.is_partial_aggregation_disabled() and skipped == 0
    got, skipped = _run(f, [_repeated(1, 10), _repeated(2, 10)])
    assert got == [[(1, 1)] * 10, [(2, 2)] * 10] and skipped == 20
    controller.close()


def test_controller_only_for_partial_steps(ctx):
    controller = ops.PartialAggregationController(ctx.lib, 1, 0.8)
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, [A(abi.AGG_MIN, 0)], 10, partial_aggregation_controller=controller)
    with pytest.raises(abi.TrinoGpuError):
        f.create_operator()
    controller.close()


AGGS = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_COUNT, 1, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_MAX, 1, -1),
        (abi.AGG_SUM, 2, -1), (abi.AGG_AVG, 2, -1), (abi.AGG_MIN, 2, -1), (abi.AGG_MAX, 2, -1), (abi.AGG_SUM, 1, 3), (abi.AGG_COUNT_STAR, -1, 3)]


def _pages(rng, card, sizes, varchar_key=False):
    pages = []
    for n in sizes:
        keys = rng.integers(0, card, n)
        knull = rng.random(n) < 0.01
        kb = Block.varchar([None if z else "key-%d" % k for k, z in zip(keys, knull)]) if varchar_key else Block.bigint(keys, knull)
        pages.append(Page(kb,
                          Block.double(rng.normal(size=n) * 100, rng.random(n) < 0.1),
                          Block.bigint(rng.integers(-1000, 1000, n), rng.random(n) < 0.1),
                          Block.boolean(rng.random(n) < 0.5, rng.random(n) < 0.05)))
    return pages


@pytest.mark.parametrize("varchar_key", [False, True])
def test_skipped_pages_match_oracle(ctx, varchar_key):
    rng = np.random.default_rng(21)
    pages = _pages(rng, 1000, (1, 777, 20_000), varchar_key)
    controller = ops.PartialAggregationController(ctx.lib, 1 << 40, 0.0)
    controller.on_flush(1 << 41, 10, 10)          # force it off (it stays off until 300 x 2^40 bytes went by)
    assert controller.is_partial_aggregation_disabled()
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(fn, ch, m) for fn, ch, m in AGGS], 100,
                                           partial_aggregation_controller=controller)
    got, skipped = _run(f, pages)
    assert skipped == sum(p.position_count for p in pages) and len(got) == len(pages)
    for page, rows in zip(pages, got):
        want = pa.skip_aggregation_rows(page.rows(), [0], AGGS, double_channels=(1,))
        assert rows_equal(rows, want, rel=0.0)
    controller.close()


# (FINAL combines counts and BIGINT sums as 128-bit pairs: a plan holds at most MAX_ACCS accumulator words, hence the shorter list)
MIXED = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_SUM, 2, -1), (abi.AGG_MAX, 2, -1),
         (abi.AGG_SUM, 1, 3), (abi.AGG_COUNT, 2, 3)]


def _final_aggs(aggs=None):
    # state channels behind the key: avg takes two
    out, ch = [], 1
    for fn, _, _ in (aggs or AGGS):
        out.append(A(fn, ch))
        ch += 2 if fn == abi.AGG_AVG else 1
    return out


def test_mixed_partial_pages_then_final_equals_single(ctx):
    rng = np.random.default_rng(22)
    # high-cardinality pages switch partial aggregation off, ~90 KB of skipped pages switch it back on: both kinds of page reach FINAL
    pages = _pages(rng, 5000, [600] * 40) + _pages(rng, 5, [600] * 20)
    controller = ops.PartialAggregationController(ctx.lib, 300, 0.5)
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(fn, ch, m) for fn, ch, m in MIXED], 100, max_partial_memory=300,
                                           partial_aggregation_controller=controller)
    op = f.create_operator()
    partial = ops.drive(op, pages)
    skipped = op.rows_with_partial_aggregation_disabled()
    op.close()
    assert 0 < skipped < sum(p.position_count for p in pages)
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, _final_aggs(MIXED), 100)
    fop = ff.create_operator()
    final = [r for p in ops.drive(fop, partial) for r in p.rows()]
    fop.close()
    want = oracle_agg_rows(pages, [0], MIXED)
    key = lambda r: (r[0] is None, r[0])
    assert rows_equal(sorted(final, key=key), sorted(want, key=key), rel=1e-9)
    controller.close()


def test_intermediate_step_passes_states_through_when_disabled(ctx):
    rng = np.random.default_rng(23)
    pages = _pages(rng, 50, (3000, 3000))
    pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(fn, ch, m) for fn, ch, m in AGGS], 100)
    partial = []
    for p in pages:
        op = pf.create_operator()
        partial += ops.drive(op, [p])
        op.close()
    controller = ops.PartialAggregationController(ctx.lib, 1 << 40, 0.0)
    controller.on_flush(1 << 41, 10, 10)
    inter = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_INTERMEDIATE, _final_aggs(), 100, partial_aggregation_controller=controller)
    got, skipped = _run(inter, partial)
    assert skipped == sum(p.position_count for p in partial)
    assert all(rows_equal(g, p.rows(), rel=0.0) for g, p in zip(got, partial))
    controller.close()


def test_fused_q1_partial_with_controller_off_then_final(ctx):
    # the fused ScanFilterAndProject -> HashAggregation operator: a skipped builder runs the pre-stage as its own FilterAndProject
    cols = o.synth_lineitem_q1(200_000, 0, 0x7C01)
    _, want = o.q1_run(cols, CUTOFF, 1)
    pages = [q1_host_page(cols, lo, lo + 50_000) for lo in range(0, 200_000, 50_000)]
    controller = ops.PartialAggregationController(ctx.lib, 1 << 40, 0.0)
    f = ops.HashAggregationOperatorFactory(ctx, [0, 1], abi.STEP_PARTIAL, q1_aggregators(), 16, pre=q1_program(),
                                           max_partial_memory=1, partial_aggregation_controller=controller)
    op = f.create_operator()
    out = []
    for i, p in enumerate(pages):
        if i == 2:
            controller.on_flush(1 << 41, 10, 10)      # another driver's flush turns partial aggregation off half way
        op.add_input(p)
        while not op.needs_input():
            out.append(op.get_output())
    op.finish()
    while not op.is_finished():
        pg = op.get_output()
        if pg is not None:
            out.append(pg)
    skipped = op.rows_with_partial_aggregation_disabled()
    op.close()
    assert skipped == 100_000
    assert out[0].position_count <= 6 and out[2].position_count > 40_000          # aggregated groups, then filtered raw rows
    finals = [A(abi.AGG_SUM, 2), A(abi.AGG_SUM, 3), A(abi.AGG_SUM, 4), A(abi.AGG_SUM, 5), A(abi.AGG_AVG, 6), A(abi.AGG_AVG, 8), A(abi.AGG_AVG, 10), A(abi.AGG_COUNT_STAR, 12)]
    ff = ops.HashAggregationOperatorFactory(ctx, [0, 1], abi.STEP_FINAL, finals, 16)
    fop = ff.create_operator()
    got = [(chr(r[0]), chr(r[1])) + tuple(r[2:]) for p in ops.drive(fop, out) for r in p.rows()]
    fop.close()
    assert [(g[0], g[1], g[9]) for g in got] == [(w[0], w[1], w[9]) for w in want]
    for g, w in zip(got, want):
        for a, b in zip(g[2:9], w[2:9]):
            assert abs(a - b) <= 1e-6 * abs(b)
    controller.close()


def test_row_typed_intermediate_states_round_trip(ctx):
    # AccumulatorCompiler.java:687-760: a two-field state travels as ONE ROW channel between the reference's PARTIAL and FINAL steps;
    # channel numbers on both factories are the Java plan's (key, count, sum, avg ROW, min)
    from trino_b200.page import RowBlock
    rng = np.random.default_rng(24)
    pages = _pages(rng, 40, (5000, 5000))
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_MIN, 2, -1)]
    pf = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(fn, ch, m) for fn, ch, m in aggs], 100, row_typed_states=True)
    partial = []
    for p in pages:
        op = pf.create_operator()
        partial += ops.drive(op, [p])
        op.close()
    assert all(p.channel_count == 5 and isinstance(p.get_block(3), RowBlock) and len(p.get_block(3).fields) == 2 for p in partial)
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, [A(abi.AGG_COUNT_STAR, 1), A(abi.AGG_SUM, 2), A(abi.AGG_AVG, 3), A(abi.AGG_MIN, 4)], 100,
                                            row_typed_states=True)
    fop = ff.create_operator()
    final = [r for p in ops.drive(fop, partial) for r in p.rows()]
    fop.close()
    want = oracle_agg_rows(pages, [0], aggs)
    assert rows_equal(final, want, rel=1e-9)
