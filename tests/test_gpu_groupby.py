"""GPU parity: GroupByHash ids (first-seen order) and HashAggregationOperator results vs the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as o
from helpers import aggregation_known_answer_cases, hash_aggregation_operator_case, rows_equal
from q1 import CUTOFF, q1_gpu_rows
from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, DictionaryBlock, Page, RunLengthEncodedBlock

pytestmark = pytest.mark.gpu
A = ops.Aggregator


# ---------------------------------------------------------------- GroupByHash
def test_group_ids_reference_cases(ctx):
    # TestGroupByHash.testGetGroupIds :187-204 (MAX_GROUP_ID shortened to 60 single-row pages x 2 tries)
    g = ops.GroupByHash(ctx, [0], 100)
    for tries in range(2):
        for value in range(60):
            ids = g.get_group_ids(Page(Block.bigint([value])))
            assert list(ids) == [value]
            assert g.get_group_count() == (value + 1 if tries == 0 else 60)
    g.close()
    # testNullGroup :163-184 incl. the forced rehash
    g = ops.GroupByHash(ctx, [0], 100)
    assert list(g.get_group_ids(Page(Block.bigint([0, None])))) == [0, 1]
    ids = g.get_group_ids(Page(Block.bigint(np.arange(1, 132749))))
    assert (ids == np.arange(2, 132750)).all()
    assert list(g.get_group_ids(Page(Block.bigint([None])))) == [1]
    g.close()
    # testDictionaryInputPage :133-160 / testRunLengthEncodedInputPage :111-131
    g = ops.GroupByHash(ctx, [0], 100)
    assert list(g.get_group_ids(Page(DictionaryBlock(Block.bigint([0, 1]), [0, 0, 1, 1])))) == [0, 0, 1, 1]
    assert g.get_group_count() == 2
    g.close()
    g = ops.GroupByHash(ctx, [0], 100)
    assert list(g.get_group_ids(Page(RunLengthEncodedBlock(Block.bigint([0]), 2)))) == [0, 0]
    assert g.get_group_count() == 1
    g.close()


@pytest.mark.parametrize("card,expected", [(50, 16), (20000, 100), (300000, 10)])
def test_group_ids_random_pages_match_oracle(ctx, card, expected):
    rng = np.random.default_rng(card)
    g = ops.GroupByHash(ctx, [0], expected)
    og = o.GroupByHash(1, expected)
    for n in (1000, 1, 65536, 300000):
        page = Page(Block.bigint(rng.integers(-card, card, n), rng.random(n) < 0.01))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()


def test_group_ids_packed_multi_column_and_other_types(ctx):
    rng = np.random.default_rng(17)
    n = 50000
    page = Page(Block.integer(rng.integers(0, 300, n), rng.random(n) < 0.02), Block.tinyint(rng.integers(65, 70, n)), Block.smallint(rng.integers(-3, 3, n), rng.random(n) < 0.02))
    g = ops.GroupByHash(ctx, [0, 1, 2], 1000)
    og = o.GroupByHash(0, 1000)
    assert (g.get_group_ids(page) == og.get_group_ids(page, [0, 1, 2])).all()
    g.close(); og.close()
    # DOUBLE key: IDENTICAL semantics, INT64_MIN-valued key, NULL
    d = Block.double([0.0, -0.0, float("nan"), 1.0, float("nan"), None, 1.0, None])
    g = ops.GroupByHash(ctx, [0], 10)
    og = o.GroupByHash(0, 10)
    assert list(g.get_group_ids(Page(d))) == list(og.get_group_ids(Page(d), [0])) == [0, 0, 1, 2, 1, 3, 2, 3]
    g.close(); og.close()
    k = Block.bigint([-2**63, 5, -2**63, None, 5])
    g = ops.GroupByHash(ctx, [0], 10)
    assert list(g.get_group_ids(Page(k))) == [0, 1, 0, 2, 1]
    g.close()


def test_group_ids_wide_composite_keys_use_fingerprints(ctx):
    # keys that do not pack into 63 bits (FlatHash territory): 64-bit fingerprint table + verification against stored keys
    rng = np.random.default_rng(23)
    g = ops.GroupByHash(ctx, [0, 1, 2], 1000)
    og = o.GroupByHash(0, 1000)
    for n in (30000, 1, 70000):
        d = rng.integers(0, 5, n).astype(np.float64)
        d[rng.random(n) < 0.05] = np.nan
        d[rng.random(n) < 0.05] = -0.0
        page = Page(Block.bigint(rng.integers(-2**62, 2**62, 40)[rng.integers(0, 40, n)], rng.random(n) < 0.02), Block.bigint(rng.integers(0, 30, n)),
                    Block.double(d, rng.random(n) < 0.02))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0, 1, 2])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()


def test_aggregation_with_wide_composite_keys(ctx):
    rng = np.random.default_rng(31)
    n = 20000
    pages = [Page(Block.bigint(rng.integers(0, 40, n)), Block.bigint(rng.integers(10**12, 10**12 + 25, n), rng.random(n) < 0.05),
                  Block.double(rng.normal(size=n)), Block.bigint(rng.integers(-9, 9, n))) for _ in range(2)]
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 2, -1), (abi.AGG_SUM, 3, -1), (abi.AGG_MAX, 2, -1)]
    got = _gpu_agg(ctx, pages, [0, 1], aggs)
    want = _oracle_agg(pages, [0, 1], aggs)
    assert rows_equal(got, want, rel=1e-6)
    assert [r[:3] for r in got] == [r[:3] for r in want]


# ---------------------------------------------------------------- aggregation
from helpers import oracle_agg_rows as _oracle_agg  # noqa: E402


def _gpu_agg(ctx, pages, key_channels, aggs, step=abi.STEP_SINGLE, expected=100, max_partial=0):
    f = ops.HashAggregationOperatorFactory(ctx, key_channels, step, [A(fn, ch, m) for fn, ch, m in aggs], expected, max_partial)
    op = f.create_operator()
    out = ops.drive(op, pages)
    op.close()
    rows = []
    for p in out:
        rows.extend(p.rows())
    return rows


AGGS = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_COUNT, 1, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_MAX, 1, -1),
        (abi.AGG_SUM, 2, -1), (abi.AGG_AVG, 2, -1), (abi.AGG_MIN, 2, -1), (abi.AGG_MAX, 2, -1), (abi.AGG_SUM, 1, 3), (abi.AGG_COUNT_STAR, -1, 3)]


def _agg_pages(rng, card, sizes):
    pages = []
    for n in sizes:
        pages.append(Page(Block.bigint(rng.integers(0, card, n), rng.random(n) < 0.01),
                          Block.double(rng.normal(size=n) * 100, rng.random(n) < 0.1),
                          Block.bigint(rng.integers(-1000, 1000, n), rng.random(n) < 0.1),
                          Block.boolean(rng.random(n) < 0.5, rng.random(n) < 0.05)))
    return pages


@pytest.mark.parametrize("card", [5, 40, 3000])
def test_aggregation_matches_oracle(ctx, card):
    # small cardinalities run the fused shared-memory path, large ones the global-table path; both must give the
    # reference's rows in first-seen group order (DOUBLE aggregates within 1e-6 relative, everything else exact)
    rng = np.random.default_rng(card)
    pages = _agg_pages(rng, card, (5000, 1, 40000, 333))
    got = _gpu_agg(ctx, pages, [0], AGGS)
    want = _oracle_agg(pages, [0], AGGS)
    assert rows_equal(got, want, rel=1e-6)
    assert [r[0] for r in got] == [r[0] for r in want]
    assert [(r[1], r[4], r[7], r[12]) for r in got] == [(r[1], r[4], r[7], r[12]) for r in want]     # counts and BIGINT sum exact


def test_general_path_slice_by_slice_matches_oracle(ctx, monkeypatch):
    # tables larger than the L2 are visited slice by slice (rows regrouped by the slot index's top bits); force that path on a
    # small table: ids stay first-seen ordered across pages, growth in the middle of a page replays deferred rows
    monkeypatch.setenv("TGPU_AGG_ROWLIST_SLICES", "1")          # the row-list form (one launch per slice), kept as the fallback for shapes the copy cannot take
    monkeypatch.setenv("TGPU_AGG_SLICE_MIN_BYTES", "0")
    monkeypatch.setenv("TGPU_AGG_SLICE_BYTES", str(256 << 10))
    rng = np.random.default_rng(77)
    pages = _agg_pages(rng, 20000, (60000, 7, 90000)) + _agg_pages(rng, 400000, (150000,))
    got = _gpu_agg(ctx, pages, [0], AGGS, expected=30000)
    want = _oracle_agg(pages, [0], AGGS)
    assert rows_equal(got, want, rel=1e-6)
    assert [r[0] for r in got] == [r[0] for r in want]
    assert [(r[1], r[4], r[7], r[12]) for r in got] == [(r[1], r[4], r[7], r[12]) for r in want]


def test_accumulator_known_answers(ctx):
    # the reference's AbstractTestAggregationFunction sequences (see helpers.aggregation_known_answer_cases): exact here, the sums are
    # of small integers
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_COUNT, 1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_MAX, 1, -1),
            (abi.AGG_SUM, 2, -1), (abi.AGG_COUNT, 2, -1)]
    for case in aggregation_known_answer_cases():
        n = len(case["values"])
        page = Page(Block.bigint(np.zeros(n, dtype=np.int64)), Block.double(case["values"].astype(np.float64), case["nulls"]), Block.bigint(case["values"], case["nulls"]))
        got = _gpu_agg(ctx, [page], [0], aggs)
        assert got == [(0, case["count_star"], case["count"], case["sum_double"], case["avg_double"], case["min"], case["max"], case["sum_bigint"], case["count"])], case["name"]


def test_hash_aggregation_operator_reference_case(ctx):
    # TestHashAggregationOperator.testHashAggregation :138-188 (restated over BIGINT channels), full size: 40 000 groups in 3 pages
    pages, keys, aggs, expected = hash_aggregation_operator_case()
    assert _gpu_agg(ctx, pages, keys, aggs, expected=100_000) == expected


def test_general_path_physical_slices_match_oracle(ctx, monkeypatch):
    # the default for tables beyond the L2: one pass over a slice-ordered COPY of the page (multi-split scatter of the channels the plan
    # reads + page row numbers for the stamps); also with the interpreted kernel (hosts without NVRTC)
    monkeypatch.setenv("TGPU_AGG_SLICE_MIN_BYTES", "0")
    monkeypatch.setenv("TGPU_AGG_SLICE_BYTES", str(256 << 10))
    rng = np.random.default_rng(78)
    pages = _agg_pages(rng, 20000, (60000, 7, 90000)) + _agg_pages(rng, 400000, (150000,))
    want = _oracle_agg(pages, [0], AGGS)
    for interpreted, stable in ((False, False), (True, False), (False, True)):
        if interpreted:
            monkeypatch.setenv("TGPU_AGG_GENERAL_INTERPRETED", "1")
        else:
            monkeypatch.delenv("TGPU_AGG_GENERAL_INTERPRETED", raising=False)
        if stable:
            monkeypatch.setenv("TGPU_AGG_STABLE_SCATTER", "1")       # the order-preserving multi-split instead of the any-order one
        got = _gpu_agg(ctx, pages, [0], AGGS, expected=30000)
        assert rows_equal(got, want, rel=1e-6)
        assert [r[0] for r in got] == [r[0] for r in want]
        assert [(r[1], r[4], r[7], r[12]) for r in got] == [(r[1], r[4], r[7], r[12]) for r in want]


def test_small_path_spills_into_general_path(ctx):
    # first page has few groups (path S), the next one thousands: state must migrate without losing ids or sums
    rng = np.random.default_rng(99)
    pages = _agg_pages(rng, 6, (2000,)) + _agg_pages(rng, 5000, (30000,)) + _agg_pages(rng, 6, (100,))
    got = _gpu_agg(ctx, pages, [0], AGGS, expected=16)
    want = _oracle_agg(pages, [0], AGGS)
    assert rows_equal(got, want, rel=1e-6)


def test_partial_then_final_equals_single(ctx):
    rng = np.random.default_rng(5)
    pages = _agg_pages(rng, 30, (4000, 4000))
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_SUM, 2, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_MAX, 2, -1), (abi.AGG_COUNT, 2, -1)]
    single = _gpu_agg(ctx, pages, [0], aggs)
    # two PARTIAL operators (one page each) -> FINAL over the intermediate state columns
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(fn, ch, m) for fn, ch, m in aggs], 100)
    partial_pages = []
    for p in pages:
        op = f.create_operator()
        partial_pages += ops.drive(op, [p])
        op.close()
    assert partial_pages[0].channel_count == 1 + 8      # avg carries (count, sum)
    final_aggs = [A(abi.AGG_COUNT_STAR, 1), A(abi.AGG_SUM, 2), A(abi.AGG_AVG, 3), A(abi.AGG_SUM, 5), A(abi.AGG_MIN, 6), A(abi.AGG_MAX, 7), A(abi.AGG_COUNT, 8)]
    ff = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, final_aggs, 100)
    op = ff.create_operator()
    out = ops.drive(op, partial_pages)
    op.close()
    final = [r for p in out for r in p.rows()]
    assert rows_equal(final, single, rel=1e-9)


def test_partial_flush_when_memory_exceeded(ctx):
    # HashAggregationOperator.needsInput :346-355 / getOutput :478-483: a full partial builder flushes and restarts
    rng = np.random.default_rng(6)
    pages = _agg_pages(rng, 2000, (10000, 10000))
    f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, [A(abi.AGG_COUNT_STAR)], 100, max_partial_memory=1024)
    op = f.create_operator()
    op.add_input(pages[0])
    assert not op.needs_input()
    first = op.get_output()
    assert first is not None and op.needs_input()
    op.add_input(pages[1])
    second = op.get_output()
    op.finish()
    assert op.get_output() is None and op.is_finished()
    total = sum(r[1] for r in first.rows()) + sum(r[1] for r in second.rows())
    assert total == 20000
    op.close()


def test_partial_flush_with_a_fused_pre_stage_and_many_groups(ctx):
    # The advisor's round-1 finding: a PARTIAL step with a fused filter + projection that overflows the shared-memory path un-fuses the
    # pre-stage (the plan then reads projection OUTPUT channels); after a maxPartialMemory flush the operator must keep routing pages
    # through that FilterAndProject instead of running the raw page through the re-pointed plan.  PARTIAL pages -> FINAL == SINGLE.
    rng = np.random.default_rng(17)
    n = 60_000
    pages = [Page(Block.bigint(rng.integers(0, 5000, n)), Block.bigint(rng.integers(-50, 50, n)), Block.double(rng.normal(size=n))) for _ in range(3)]
    keep = ops.Call(abi.EX_GE, ops.Col(1, abi.V_BIGINT), ops.Const(-10, abi.V_BIGINT))
    doubled = ops.Call(abi.EX_MUL, ops.Col(2, abi.V_DOUBLE), ops.Const(2.0, abi.V_DOUBLE))

    def program():
        return ops.PageProcessorProgram(keep, [0, 1, doubled])          # channels of the aggregation input: key, value, 2 * x

    aggs = [A(abi.AGG_COUNT_STAR), A(abi.AGG_SUM, 1), A(abi.AGG_SUM, 2)]
    single = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, aggs, 16, pre=program())
    op = single.create_operator()
    want = [r for p in ops.drive(op, pages) for r in p.rows()]
    op.close()
    partial = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_PARTIAL, aggs, 16, max_partial_memory=64 << 10, pre=program())
    op = partial.create_operator()
    flushed = ops.drive(op, pages)
    op.close()
    assert len(flushed) >= 3                                              # every page overflowed the partial memory limit
    final = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_FINAL, [A(abi.AGG_COUNT_STAR, 1), A(abi.AGG_SUM, 2), A(abi.AGG_SUM, 3)], 16)
    op = final.create_operator()
    got = [r for p in ops.drive(op, flushed) for r in p.rows()]
    op.close()
    assert [r[:3] for r in got] == [r[:3] for r in want] and rows_equal(got, want, rel=1e-9)
    expected_rows = sum(int((np.asarray(p.get_block(1).values) >= -10).sum()) for p in pages)
    assert sum(r[1] for r in got) == expected_rows


def test_bigint_sum_overflow_raises(ctx):
    page = Page(Block.bigint([1, 1]), Block.bigint([2**62, 2**62]))
    with pytest.raises(abi.TrinoGpuError) as e:
        _gpu_agg(ctx, [page], [0], [(abi.AGG_SUM, 1, -1)])
    assert e.value.code == abi.ERR_NUMERIC_VALUE_OUT_OF_RANGE


def test_empty_input_and_all_null_inputs(ctx):
    assert _gpu_agg(ctx, [], [0], [(abi.AGG_COUNT_STAR, -1, -1)]) == []
    page = Page(Block.bigint([7, 7, 8]), Block.double([None, None, 1.0]))
    rows = _gpu_agg(ctx, [page], [0], [(abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_COUNT, 1, -1), (abi.AGG_MIN, 1, -1)])
    assert rows == [(7, None, None, 0, None), (8, 1.0, 1.0, 1, 1.0)]


# ---------------------------------------------------------------- Q1
@pytest.mark.parametrize("fused", [True, False])
def test_q1_matches_oracle(ctx, fused):
    cols = o.synth_lineitem_q1(1_000_000, 0, 0x7C01)
    _, want = o.q1_run(cols, CUTOFF, 1)
    for page_rows in (None, 250_000):
        got = q1_gpu_rows(ctx, cols, CUTOFF, page_rows, fused)
        assert [(g[0], g[1], g[9]) for g in got] == [(w[0], w[1], w[9]) for w in want]      # groups, first-seen order, counts: exact
        for g, w in zip(got, want):
            for a, b in zip(g[2:9], w[2:9]):
                assert abs(a - b) <= 1e-6 * abs(b)                                           # north_star: 1e-6 relative for DOUBLE


def test_q1_fused_is_run_to_run_deterministic(ctx):
    cols = o.synth_lineitem_q1(500_000, 0, 0x7C01)
    a = q1_gpu_rows(ctx, cols, CUTOFF)
    b = q1_gpu_rows(ctx, cols, CUTOFF)
    assert a == b


# ---------------------------------------------------------------- variable-width keys (FlatHash territory: M/operator/FlatHash.java:309-348)
def _words(rng, n, vocab):
    return [vocab[i] for i in rng.integers(0, len(vocab), n)]


def test_group_ids_varchar_keys_match_oracle(ctx):
    rng = np.random.default_rng(41)
    short = ["", "A", "N", "R", "ab", "abc", "1234567"]                        # <= 7 bytes: keyed by their bytes
    long_ = ["12345678", "a much longer string than seven bytes", "x" * 40, "x" * 41, "naïve café", "日本語のキー"]   # hashed + compared
    vocab = short + long_ + [f"key-{i:06d}" for i in range(3000)]
    g = ops.GroupByHash(ctx, [0], 100)
    og = o.GroupByHash(0, 100)
    for n in (1, 5000, 70000, 3):
        vals = _words(rng, n, vocab)
        nulls = rng.random(n) < 0.03
        page = Page(Block.varchar([None if z else v for v, z in zip(vals, nulls)]))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()
    # TestGroupByHash shapes over VARCHAR: dictionary and run-length encoded inputs give the ids of the flat block (:111-160)
    g = ops.GroupByHash(ctx, [0], 100)
    assert list(g.get_group_ids(Page(DictionaryBlock(Block.varchar(["x", "yy"]), [0, 0, 1, 1, 0])))) == [0, 0, 1, 1, 0]
    assert list(g.get_group_ids(Page(RunLengthEncodedBlock(Block.varchar(["yy"]), 3)))) == [1, 1, 1]
    assert list(g.get_group_ids(Page(Block.varchar(["zzz", "x", None, "yy"])))) == [2, 0, 3, 1]
    g.close()


def test_group_ids_mixed_varchar_and_fixed_keys(ctx):
    rng = np.random.default_rng(43)
    g = ops.GroupByHash(ctx, [0, 1, 2], 1000)
    og = o.GroupByHash(0, 1000)
    for n in (20000, 1, 60000):
        page = Page(Block.varchar(_words(rng, n, ["A", "N", "R"])), Block.varchar(_words(rng, n, ["F", "O", "a-long-line-status"])),
                    Block.integer(rng.integers(0, 40, n), rng.random(n) < 0.02))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0, 1, 2])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()


@pytest.mark.parametrize("step_pages", [1, 3])
def test_aggregation_with_varchar_keys_matches_oracle(ctx, step_pages):
    rng = np.random.default_rng(47)
    vocab = ["", "A", "N", "R", "a much longer string than seven bytes", "12345678"] + [f"k{i}" for i in range(500)]
    pages = []
    for _ in range(step_pages):
        n = 30000
        keys = _words(rng, n, vocab)
        nulls = rng.random(n) < 0.02
        pages.append(Page(Block.varchar([None if z else v for v, z in zip(keys, nulls)]), Block.double(rng.normal(size=n), rng.random(n) < 0.05),
                          Block.bigint(rng.integers(-50, 50, n))))
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_SUM, 2, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_COUNT, 0, -1)]
    got = _gpu_agg(ctx, pages, [0], aggs)
    want = _oracle_agg(pages, [0], aggs)
    assert rows_equal(got, want, rel=1e-6)
    assert [r[0] for r in got] == [r[0] for r in want]                          # key strings come back, in first-seen order
    # PARTIAL -> FINAL over the string keys of the intermediate pages
    partial = _gpu_agg(ctx, pages, [0], aggs[:3], step=abi.STEP_PARTIAL)
    ppage = Page(Block.varchar([r[0] for r in partial]), Block.bigint([r[1] for r in partial]), Block.double([r[2] for r in partial]), Block.bigint([r[3] for r in partial]))
    final = _gpu_agg(ctx, [ppage], [0], [(abi.AGG_COUNT_STAR, 1, -1), (abi.AGG_SUM, 2, -1), (abi.AGG_SUM, 3, -1)], step=abi.STEP_FINAL)
    assert rows_equal(final, [r[:4] for r in want], rel=1e-6)


def test_q1_with_varchar_keys_matches_oracle(ctx):
    # the reference's Q1 groups by VARCHAR(1) l_returnflag / l_linestatus: same rows as with INT8 codes
    cols = o.synth_lineitem_q1(300_000, 0, 0x7C01)
    _, want = o.q1_run(cols, CUTOFF, 1)
    flags = [chr(c) for c in cols["returnflag"]]
    status = [chr(c) for c in cols["linestatus"]]
    page = Page(Block.integer(cols["shipdate"]), Block.varchar(flags), Block.varchar(status), Block.double(cols["quantity"]), Block.double(cols["extendedprice"]),
                Block.double(cols["discount"]), Block.double(cols["tax"]))
    from q1 import q1_factory
    for fused in (True, False):
        if fused:
            op = q1_factory(ctx, True).create_operator()
            out = ops.drive(op, [page])
            op.close()
        else:
            from q1 import q1_program
            fp = ops.FilterAndProjectOperatorFactory(ctx, q1_program()).create_operator()
            agg = q1_factory(ctx, False).create_operator()
            fp.add_input(page)
            o1 = fp.get_output_device()
            agg.add_input(o1)
            o1.release()
            agg.finish()
            out = [agg.get_output()]
            fp.close(); agg.close()
        rows = [r for p in out for r in p.rows()]
        assert len(rows) == len(want)
        for g_, w in zip(rows, want):
            assert (g_[0].decode(), g_[1].decode()) == (w[0], w[1]) and g_[9] == w[9], (g_, w)
            for a, b in zip(g_[2:9], w[2:9]):
                assert abs(a - b) <= 1e-6 * abs(b)


def test_global_aggregation_default_rows(ctx):
    # HashAggregationOperator.getGlobalAggregationOutput :537-567 (TestHashAggregationOperator.testHashAggregationWithGlobals :191-230):
    # no input, global grouping sets {42, 49}: one row per set, $group_id = id, other keys NULL, count 0, everything else NULL
    A_ = ops.Aggregator
    types = [abi.UTF8, abi.INT64, abi.FLOAT64, abi.INT64]
    f = ops.HashAggregationOperatorFactory(ctx, [0, 1], abi.STEP_SINGLE, [A_(abi.AGG_COUNT_STAR), A_(abi.AGG_SUM, 2), A_(abi.AGG_AVG, 3), A_(abi.AGG_MAX, 3), A_(abi.AGG_COUNT, 2)],
                                           expected_groups=100, global_aggregation_group_ids=[42, 49], group_id_channel=1, input_types=types)
    op = f.create_operator()
    out = ops.drive(op, [])
    op.close()
    assert [r for p in out for r in p.rows()] == [(None, 42, 0, None, None, None, 0), (None, 49, 0, None, None, None, 0)]
    # with input the default rows do not appear; a PARTIAL step never makes them
    op = f.create_operator()
    out = ops.drive(op, [Page(Block.varchar(["x"]), Block.bigint([7]), Block.double([1.5]), Block.bigint([3]))])
    op.close()
    assert [r for p in out for r in p.rows()] == [(b"x", 7, 1, 1.5, 3.0, 3, 1)]
    pf = ops.HashAggregationOperatorFactory(ctx, [0, 1], abi.STEP_PARTIAL, [A_(abi.AGG_COUNT_STAR)], global_aggregation_group_ids=[42], group_id_channel=1, input_types=types)
    op = pf.create_operator()
    assert ops.drive(op, []) == []
    op.close()


def test_group_keys_sharing_a_64_bit_fingerprint_are_kept_apart(ctx):
    # wide composite keys are addressed by a 64-bit fingerprint; two tuples that share it used to fail the query, now the later one
    # rehashes (full-key comparison on every hit, like FlatHash.valueIdentical, M/operator/FlatHash.java:445-469)
    from helpers import colliding_groupby_pairs
    t1 = (2**40 + 1, 2**41 + 5)
    t2 = (2**42 + 9, colliding_groupby_pairs(t1[0], t1[1], 2**42 + 9))
    t3 = (2**43 + 3, colliding_groupby_pairs(t1[0], t1[1], 2**43 + 3))
    rng = np.random.default_rng(9)
    g = ops.GroupByHash(ctx, [0, 1], 100)
    og = o.GroupByHash(0, 100)
    for tuples in ([t2, t1, t2, t1, t3], [t3, t3, t1], [t1]):
        a = np.concatenate([[t[0] for t in tuples], rng.integers(2**50, 2**50 + 300, 4000)]).astype(np.int64)
        b = np.concatenate([[t[1] for t in tuples], rng.integers(2**51, 2**51 + 300, 4000)]).astype(np.int64)
        page = Page(Block.bigint(a), Block.bigint(b))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0, 1])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()
    # and through the operator (sums per colliding tuple stay separate)
    a = np.array([t1[0], t2[0], t1[0], t3[0], t2[0]], dtype=np.int64)
    b = np.array([t1[1], t2[1], t1[1], t3[1], t2[1]], dtype=np.int64)
    pages = [Page(Block.bigint(a), Block.bigint(b), Block.bigint([1, 10, 100, 1000, 10000]))]
    aggs = [(abi.AGG_SUM, 2, -1), (abi.AGG_COUNT_STAR, -1, -1)]
    got = _gpu_agg(ctx, pages, [0, 1], aggs)
    assert got == _oracle_agg(pages, [0, 1], aggs) == [(t1[0], t1[1], 101, 2), (t2[0], t2[1], 10010, 2), (t3[0], t3[1], 1000, 1)]
