"""Pins the oracle's GroupByHash / join / partitioner restatements on the behaviour the reference's own unit
tests assert (cases restated as data in tests/golden/reference_cases.json, file:line cited there)."""
import numpy as np
import pytest

import oracle_lib as o
from helpers import aggregation_known_answer_cases, hash_aggregation_operator_case, join_type_of, oracle_agg_rows, oracle_outer_rows, oracle_join_rows, reference_cases, rows_equal
from trino_b200 import abi
from trino_b200.page import Block, DictionaryBlock, Page, RunLengthEncodedBlock

KINDS = [1, 2]   # BigintGroupByHash, FlatGroupByHash (T/operator/TestGroupByHash.java:60-80 runs both)


@pytest.mark.parametrize("kind", KINDS)
def test_groupby_add_page_and_get_group_ids(kind):
    # TestGroupByHash.testAddPage :84-109 / testGetGroupIds :187-204 (MAX_GROUP_ID = 500)
    g = o.GroupByHash(kind, 100)
    for tries in range(2):
        for value in range(500):
            ids = g.get_group_ids(Page(Block.bigint([value])), [0])
            assert g.group_count() == (value + 1 if tries == 0 else 500)
            assert list(ids) == [value]
    g.close()


@pytest.mark.parametrize("kind", KINDS)
def test_groupby_null_group_survives_rehash(kind):
    # TestGroupByHash.testNullGroup :163-184
    g = o.GroupByHash(kind, 100)
    assert list(g.get_group_ids(Page(Block.bigint([0, None])), [0])) == [0, 1]
    g.get_group_ids(Page(Block.bigint(np.arange(1, 132749))), [0])
    assert list(g.get_group_ids(Page(Block.bigint([None])), [0])) == [1]
    g.close()


@pytest.mark.parametrize("kind", KINDS)
def test_groupby_dictionary_and_rle_inputs(kind):
    # TestGroupByHash.testDictionaryInputPage :133-160, testRunLengthEncodedInputPage :111-131
    g = o.GroupByHash(kind, 100)
    page = Page(DictionaryBlock(Block.bigint([0, 1]), [0, 0, 1, 1]))
    assert list(g.get_group_ids(page, [0])) == [0, 0, 1, 1]
    assert g.group_count() == 2
    g.close()
    g = o.GroupByHash(kind, 100)
    page = Page(RunLengthEncodedBlock(Block.bigint([0]), 2))
    assert list(g.get_group_ids(page, [0])) == [0, 0]
    assert g.group_count() == 1
    g.close()


def test_groupby_force_rehash_preserves_ids_and_capacity():
    # TestGroupByHash.testForceRehash :260-277: expectedSize 100 -> values 0..(100*2-1); ids stay first-seen
    for kind in KINDS:
        g = o.GroupByHash(kind, 100)
        cap0 = None
        ids = g.get_group_ids(Page(Block.bigint([7])), [0])
        cap0 = g.capacity()
        ids = g.get_group_ids(Page(Block.bigint(np.arange(0, 1000))), [0])
        assert g.capacity() > cap0
        expected = np.arange(0, 1000)
        expected = np.where(expected == 7, 0, np.where(expected < 7, expected + 1, expected))
        assert (ids == expected).all()
        g.close()


def test_bigint_and_flat_agree_on_random_input():
    rng = np.random.default_rng(5)
    vals = rng.integers(-50, 50, size=5000)
    nulls = rng.random(5000) < 0.05
    page = Page(Block.bigint(vals, nulls))
    a, b = o.GroupByHash(1, 16), o.GroupByHash(2, 16)
    ia, ib = a.get_group_ids(page, [0]), b.get_group_ids(page, [0])
    assert (ia == ib).all()
    # dense, first-seen order
    seen = {}
    for i, (v, isn) in enumerate(zip(vals, nulls)):
        key = None if isn else int(v)
        if key not in seen:
            seen[key] = len(seen)
        assert ia[i] == seen[key]
    a.close(); b.close()


def test_flat_multi_column_keys_with_varchar_and_double():
    rf = Block.varchar(["A", "N", "N", "R", "A", None, None])
    ls = Block.varchar(["F", "O", "F", "F", "F", "F", "F"])
    d = Block.double([0.0, -0.0, 1.0, float("nan"), 0.0, float("nan"), float("nan")])
    g = o.GroupByHash(0, 4)
    ids = g.get_group_ids(Page(rf, ls, d), [0, 1, 2])
    assert list(ids) == [0, 1, 2, 3, 0, 4, 4]   # -0.0 IDENTICAL +0.0 only within the same other keys; NaN IDENTICAL NaN
    g.close()


def _case_pages(case):
    build = Page(Block.bigint(case["build"])) if case["build"] else Page(Block.bigint([]), position_count=0)
    probe = Page(Block.bigint(case["probe"]))
    jt = join_type_of(case)
    return build, probe, jt


@pytest.mark.parametrize("force_default", [False, True])
def test_join_reference_cases(force_default):
    # BigintPagesHash and DefaultPagesHash must both reproduce the reference's expected rows
    for case in reference_cases()["join"]:
        build, probe, jt = _case_pages(case)
        rows = oracle_join_rows(build, probe, 0, 0, [0], [0], jt, case["single_match"], force_default)
        expected = [tuple(r) for r in case["expected"]]
        assert rows == expected, case["source"]


def test_join_probe_outer_sequence_case_with_varchar_key():
    # TestHashJoinOperator.testProbeOuterJoin :481-530 exactly (VARCHAR key -> DefaultPagesHash)
    c = reference_cases()["probe_outer_sequence"]
    b0, b1, b2 = c["build_initial"]
    p0, p1, p2 = c["probe_initial"]
    nb, np_ = c["build_rows"], c["probe_rows"]
    build = Page(Block.varchar([str(b0 + i) for i in range(nb)]), Block.bigint([b1 + i for i in range(nb)]), Block.bigint([b2 + i for i in range(nb)]))
    probe = Page(Block.varchar([str(p0 + i) for i in range(np_)]), Block.bigint([p1 + i for i in range(np_)]), Block.bigint([p2 + i for i in range(np_)]))
    rows = oracle_join_rows(build, probe, 0, 0, [0, 1, 2], [0, 1, 2], abi.JOIN_PROBE_OUTER, False)
    expected = []
    for i in range(np_):
        k = p0 + i
        if b0 <= k < b0 + nb:
            expected.append((str(k).encode(), p1 + i, p2 + i, str(k).encode(), b1 + (k - b0), b2 + (k - b0)))
        else:
            expected.append((str(k).encode(), p1 + i, p2 + i, None, None, None))
    assert rows == expected


def test_join_duplicate_chain_is_descending_row_order():
    # BigintPagesHash.insertValue :122-141 + ArrayPositionLinks.link :45-50: head = last inserted, then earlier rows
    build = Page(Block.bigint([5, 7, 5, 5, 7]))
    j = o.Join(build, [0])
    pos = j.positions(Page(Block.bigint([5, 7, 6])), [0])
    assert list(pos) == [3, 4, -1]
    assert list(j.links()) == [-1, -1, 0, 2, 1]
    pi, bi = j.expand(pos)
    assert list(zip(pi, bi)) == [(0, 3), (0, 2), (0, 0), (1, 4), (1, 1)]
    j.close()


def test_partitioner_every_row_lands_in_exactly_one_partition():
    # T/operator/output/TestPagePartitioner.java:212-260 (testOutputForSimplePage...): partition = f(row) for every row
    rng = np.random.default_rng(11)
    page = Page(Block.bigint(rng.integers(0, 1000, 4096)), Block.double(rng.normal(size=4096)))
    lists, _ = o.partition_positions(page, [0], 16, None, 16, -1, False, False)
    ids = o.partition_ids(page, [0], 16)
    allpos = np.sort(np.concatenate(lists))
    assert (allpos == np.arange(4096)).all()
    for p, l in enumerate(lists):
        assert (ids[l] == p).all()
        assert (np.diff(l) > 0).all()


def test_partitioner_null_channel_replicates_and_any_row_once():
    # TestPagePartitioner.java:330-480: NULL-channel rows go to all partitions; replicate-any-row sends row 0 everywhere once
    keys = Block.bigint([3, None, 5, 6, None, 8, 9, 10, 11, 12], None)
    page = Page(keys)
    lists, flag = o.partition_positions(page, [0], 4, None, 4, 0, True, False)
    assert flag
    ids = o.partition_ids(page, [0], 4)
    for p, l in enumerate(lists):
        l = list(l)
        assert l[0] == 0                      # replicated row first
        assert l[1:3] == [1, 4]               # then NULL rows
        own = [i for i in range(1, 10) if i not in (1, 4) and ids[i] == p]
        assert l[3:] == own
    # second page: nothing replicated any more
    lists2, flag2 = o.partition_positions(page, [0], 4, None, 4, 0, True, True)
    assert flag2
    assert all(list(l)[:2] == [1, 4] for l in lists2)


def test_partitioner_row_wise_strategy_keeps_row_order():
    # positions < 2 x partitions -> partitionPageByRow :229-271: replicated rows interleave in row order
    keys = Block.bigint([None, 1, None])
    lists, _ = o.partition_positions(Page(keys), [0], 4, None, 4, 0, False, False)
    ids = o.partition_ids(Page(keys), [0], 4)
    for p, l in enumerate(lists):
        assert list(l) == [i for i in range(3) if keys.is_null(i) or ids[i] == p]


def test_q1_pipeline_matches_numpy():
    cols = o.synth_lineitem_q1(200_000, 0, 0x7C01)
    secs, rows = o.q1_run(cols, 10471, 4)
    sel = cols["shipdate"] <= 10471
    assert len(rows) == 4
    for rf, ls, sq, sp, sdp, sc, aq, ap, ad, cnt in rows:
        m = sel & (cols["returnflag"] == ord(rf)) & (cols["linestatus"] == ord(ls))
        assert cnt == int(m.sum())
        q, e, d, t = cols["quantity"][m], cols["extendedprice"][m], cols["discount"][m], cols["tax"][m]
        assert abs(sq - q.sum()) <= 1e-9 * abs(sq)
        assert abs(sp - e.sum()) <= 1e-9 * abs(sp)
        assert abs(sdp - (e * (1 - d)).sum()) <= 1e-9 * abs(sdp)
        assert abs(sc - (e * (1 - d) * (1 + t)).sum()) <= 1e-9 * abs(sc)
        assert abs(aq - q.mean()) <= 1e-9 * abs(aq)
        assert abs(ad - d.mean()) <= 1e-9 * abs(ad)


def test_synth_shapes():
    keys = o.synth_orders_keys(1000, 0, 1000, 0x7C02, True)
    assert len(set(keys.tolist())) == 1000
    plain = o.synth_orders_keys(1000, 0, 1000, 0x7C02, False)
    assert sorted(keys.tolist()) == plain.tolist()
    assert plain[:9].tolist() == [1, 2, 3, 4, 5, 6, 7, 8, 33]       # TPC-H sparse order keys
    rows = o.synth_lineitem_rows(1000)
    li = o.synth_lineitem_keys(1000, 0, rows, 0x7C01, False)
    assert set(li.tolist()) == set(plain.tolist())                   # 100 % match rate
    assert (np.diff(li) >= 0).all()                                  # order-key clustered
    counts = np.unique(li, return_counts=True)[1]
    assert counts.min() >= 1 and counts.max() <= 7 and abs(counts.mean() - 4) < 0.05
    # slices agree with the whole
    assert (o.synth_lineitem_keys(1000, 100, 50, 0x7C01, True) == o.synth_lineitem_keys(1000, 0, rows, 0x7C01, True)[100:150]).all()


def test_semi_join_reference_cases():
    # TestHashSemiJoinOperator: testSemiJoin, testBuildSideNulls, testProbeSideNulls, testProbeAndBuildNulls
    for case in reference_cases()["semi_join"]:
        got = o.semi_join_bigint(Block.bigint(case["set"]), Block.bigint(case["probe"]))
        assert got == case["expected"], case["source"]
    # an empty set answers false even for a NULL probe (HashSemiJoinOperator.java:184-187)
    assert o.semi_join_bigint(Block.bigint([]), Block.bigint([1, None])) == [False, False]


def test_semi_join_over_floats_is_identical_membership():
    """The ChannelSet is a FlatSet compared with IDENTICAL (M/operator/FlatSet.java:54,374; DoubleType.java:218-229, RealType.java:172-185): the
    BIGINT reference cases (TestHashSemiJoinOperator) hold verbatim over DOUBLE and REAL values, a NaN of any encoding is one member, -0.0 is +0.0."""
    for case in reference_cases()["semi_join"]:
        for make in (Block.double, lambda v: Block.real([None if x is None else np.float32(x) for x in v])):
            assert o.semi_join_float(make(case["set"]), make(case["probe"])) == case["expected"], case["source"]
    nan_a = np.array([0x7FF8000000000000], dtype=np.uint64).view(np.float64)[0]
    nan_b = np.array([0xFFF0000000000001], dtype=np.uint64).view(np.float64)[0]
    assert o.semi_join_float(Block.double([1.0, nan_a, -0.0]), Block.double([nan_b, 0.0, 2.0, None])) == [True, True, False, None]
    assert o.semi_join_float(Block.double([1.0, None]), Block.double([nan_b, 1.0])) == [None, True]          # not contained + the set holds a NULL -> NULL
    assert o.semi_join_float(Block.double([]), Block.double([nan_b, None])) == [False, False]                # empty set
    rn = np.array([0x7FC00000, 0xFF800001], dtype=np.uint32).view(np.float32)
    assert o.semi_join_float(Block.real(np.array([rn[0], 4.0], dtype=np.float32)), Block.real(np.array([rn[1], -4.0, 4.0], dtype=np.float32))) == [True, False, True]


def test_local_partition_function_and_constants_follow_the_row_hash():
    """LocalPartitionGenerator (M/operator/exchange/LocalPartitionGenerator.java:45-77) over the row hash; a partition constant is a
    RunLengthEncodedBlock of the value in the function page (PagePartitioner.java:436-451): one partition for a lone constant."""
    rng = np.random.default_rng(3)
    n = 5000
    page = Page(Block.bigint(rng.integers(-10**9, 10**9, n)), Block.varchar(["k%d" % (i % 50) for i in range(n)]))
    for P in (1, 2, 16):
        ids = o.local_partition_ids(page, [0, 1], P)
        assert ids.min() >= 0 and ids.max() < P
        if P == 16:
            assert len(set(ids.tolist())) == 16
    assert (o.local_partition_ids(page, [0], 8) != o.partition_ids(page, [0], 8)).any()      # not the inter-stage function
    const = Page(RunLengthEncodedBlock(Block.bigint([1]), n), page.get_block(0))
    assert len(set(o.partition_ids(const, [0], 37).tolist())) == 1
    mixed = o.partition_ids(const, [0, 1], 37)
    assert len(set(mixed.tolist())) > 30 and (mixed != o.partition_ids(page, [0], 37)).any()


def test_accumulator_known_answers():
    # AbstractTestAggregationFunction sequences with the expected values of TestDoubleSum/TestDoubleAverage/TestCount/TestLongSum
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_COUNT, 1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_AVG, 1, -1), (abi.AGG_MIN, 1, -1), (abi.AGG_MAX, 1, -1),
            (abi.AGG_SUM, 2, -1), (abi.AGG_COUNT, 2, -1)]
    for case in aggregation_known_answer_cases():
        n = len(case["values"])
        page = Page(Block.bigint(np.zeros(n, dtype=np.int64)), Block.double(case["values"].astype(np.float64), case["nulls"]), Block.bigint(case["values"], case["nulls"]))
        rows = oracle_agg_rows([page], [0], aggs)
        assert rows == [(0, case["count_star"], case["count"], case["sum_double"], case["avg_double"], case["min"], case["max"], case["sum_bigint"], case["count"])], case["name"]


def test_hash_aggregation_operator_reference_case():
    pages, keys, aggs, expected = hash_aggregation_operator_case(4000)
    assert oracle_agg_rows(pages, keys, aggs) == expected


def test_min_max_double_nan_ordering():
    # max uses COMPARISON_UNORDERED_FIRST (NaN lowest), min COMPARISON_UNORDERED_LAST (NaN highest): the orderings pinned by
    # TestMinMaxByAggregation.testMaxDoubleVarchar :270-291 (max_by over (NaN,1,2) / (1,NaN,2) / (1,2,NaN) picks the 2.0 row) and
    # testMinRealVarchar :316-337 (min_by picks the 1.0 row); MaxAggregationFunction.java:49 / MinAggregationFunction.java:49
    nan = float("nan")
    aggs = [(abi.AGG_MAX, 1, -1), (abi.AGG_MIN, 1, -1)]
    for values in ([nan, 1.0, 2.0], [1.0, nan, 2.0], [1.0, 2.0, nan]):
        page = Page(Block.bigint(np.zeros(3, dtype=np.int64)), Block.double(values))
        assert oracle_agg_rows([page], [0], aggs) == [(0, 2.0, 1.0)], values
    # only NaN rows: the first value sets the state, nothing replaces it (state.isNull() branch)
    rows = oracle_agg_rows([Page(Block.bigint([0, 0]), Block.double([nan, nan]))], [0], aggs)
    assert rows[0][1] != rows[0][1] and rows[0][2] != rows[0][2]


def test_partitioned_lookup_source_matches_single_table():
    # PartitionedLookupSource over P partitions == one table: decode(partition, position) lands on the same build row
    # (M/operator/join/unspilled/PartitionedLookupSource.java:149-186,259-275)
    n_orders = 30_011
    okeys = o.synth_orders_keys(n_orders, 0, n_orders, 0x7C02, True)
    payload = (okeys % 2557).astype(np.int32)
    rows = o.synth_lineitem_rows(n_orders)
    lkeys = o.synth_lineitem_keys(n_orders, 0, rows, 0x7C01, True)
    lkeys[::97] = -5                                             # misses
    single = o.Join(Page(Block.bigint(okeys)), [0])
    want = single.positions(Page(Block.bigint(lkeys)), [0])
    for partitions, threads in ((1, 1), (4, 3), (16, 8)):
        pj = o.PartitionedJoin(okeys, payload, partitions, threads)
        assert pj.partitions == partitions and pj.threads == threads
        pos = pj.alloc(rows, np.int64)
        pay = pj.alloc(rows, np.int32)
        for _ in range(2):                                       # the pool is reused across passes
            assert pj.probe(lkeys, pos, pay) > 0
        got = pj.decode(pos)
        assert (got == want).all()
        assert (pay == np.where(want >= 0, payload[np.maximum(want, 0)], 0)).all()
        # the partition field of the encoded position is LocalPartitionGenerator.getPartition(rawHash) (:139,159)
        lib = o.load()
        for k, e in list(zip(lkeys[want >= 0], pos[want >= 0]))[:64]:
            raw = lib.orc_hash_long(int(k))
            raw = raw - (1 << 64) if raw >= (1 << 63) else raw
            assert (int(e) & (pj.partitions - 1)) == lib.orc_local_partition(raw, pj.partitions)
        pj.close()
    single.close()
