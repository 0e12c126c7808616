"""GPU parity for REAL columns (TGPU_FLOAT32: the float's raw bits in an IntArrayBlock, S/type/RealType.java:104-121) through the C ABI:
partition key (hash :151-159), join key (EQUAL :145-149 - NaN matches nothing, -0.0 matches +0.0), group-by key (IDENTICAL :172-185 -
NaN is one group, -0.0 and +0.0 are one group whose output value is the first one seen), pass-through payload everywhere."""
import numpy as np
import pytest

import oracle_lib as o
from helpers import gpu_join_rows, oracle_agg_rows, oracle_join_rows, rows_equal
from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, Page

pytestmark = pytest.mark.gpu

NAN_BITS = [0x7FC00000, 0xFFC00000, 0x7FC00001, 0x7F800001, 0xFFFFFFFF]


def _reals(rng, n, card=50):
    """float32 values with few distinct magnitudes, both zeros, several NaN encodings, infinities"""
    v = (rng.integers(-card, card, n) * 0.25).astype(np.float32)
    v[rng.random(n) < 0.05] = np.float32(-0.0)
    special = rng.random(n) < 0.05
    bits = v.view(np.uint32).copy()
    bits[special] = rng.choice(np.array(NAN_BITS + [0x7F800000, 0xFF800000], dtype=np.uint32), int(special.sum()))
    return bits.view(np.float32)


def _bits(block):
    return block.values.view(np.int32).tolist()


def test_partition_ids_and_pages_with_real_columns(ctx):
    rng = np.random.default_rng(21)
    n = 40000
    page = Page(Block.real(_reals(rng, n), rng.random(n) < 0.05), Block.bigint(rng.integers(0, 1000, n)), Block.real(rng.normal(size=n).astype(np.float32)))
    for keys in ([0], [0, 1], [1, 2]):
        for buckets in (2, 8, 37):
            op = ops.PartitionedOutputOperatorFactory(ctx, keys, buckets).create_operator()
            want = o.partition_ids(page, keys, buckets)
            assert (op.get_partitions(page) == want).all(), (keys, buckets)
            # the partitioned pages: REAL payload bits unchanged, rows in page order
            op.add_input(page)
            while True:
                r = op.get_output_with_partition()
                if r is None:
                    break
                p, out = r
                idx = np.flatnonzero(want == p)
                for c in (0, 2):
                    got_block, src = out.get_block(c).flatten(), page.get_block(c)
                    keep = np.ones(len(idx), bool) if src.nulls is None else ~src.nulls[idx]
                    assert (got_block.values.view(np.int32)[keep] == src.values.view(np.int32)[idx][keep]).all()
                    assert (got_block.nulls is None and keep.all()) or (got_block.nulls == ~keep).all()
                assert out.get_block(1).flatten().values.tolist() == page.get_block(1).values[idx].tolist()
            op.close()
    gen = ops.LocalPartitionGenerator(ctx, [0], 16)
    assert (gen.get_partitions(page) == o.local_partition_ids(page, [0], 16)).all()
    gen.close()


@pytest.mark.parametrize("join_type", [abi.JOIN_INNER, abi.JOIN_PROBE_OUTER])
@pytest.mark.parametrize("channels", [1, 2])
def test_join_on_real_keys(ctx, join_type, channels):
    rng = np.random.default_rng(22)
    nb, npr = 3000, 20000
    build = Page(Block.real(_reals(rng, nb, 400), rng.random(nb) < 0.03), Block.bigint(rng.integers(0, 3, nb)), Block.real(rng.normal(size=nb).astype(np.float32)),
                 Block.bigint(np.arange(nb)))
    probe = Page(Block.real(_reals(rng, npr, 500), rng.random(npr) < 0.03), Block.bigint(rng.integers(0, 3, npr)), Block.real(rng.normal(size=npr).astype(np.float32)))
    keys = [0] if channels == 1 else [0, 1]
    got = gpu_join_rows(ctx, [build], [probe], keys, keys, [0, 2], [0, 2, 3], join_type, False)
    want = oracle_join_rows(build, probe, keys, keys, [0, 2], [0, 2, 3], join_type, False)
    assert rows_equal(got, want)
    assert len(got) > npr // 4
    # NaN keys never match (S/type/RealType.java:145-149): under INNER no output row carries a NaN key
    if join_type == abi.JOIN_INNER:
        assert all(r[0] == r[0] for r in got)


@pytest.mark.parametrize("card", [6, 3000])
def test_group_by_real_keys(ctx, card):
    rng = np.random.default_rng(card)
    pages = []
    for n in (7000, 1, 30000):
        pages.append(Page(Block.real(_reals(rng, n, card), rng.random(n) < 0.02), Block.bigint(rng.integers(-1000, 1000, n), rng.random(n) < 0.1),
                          Block.integer(rng.integers(0, 3, n))))
    aggs = [(abi.AGG_COUNT_STAR, -1, -1), (abi.AGG_SUM, 1, -1), (abi.AGG_MIN, 1, -1)]
    for keys in ([0], [2, 0]):
        f = ops.HashAggregationOperatorFactory(ctx, keys, abi.STEP_SINGLE, [ops.Aggregator(fn, ch, m) for fn, ch, m in aggs], 100, 0)
        op = f.create_operator()
        out = ops.drive(op, pages)
        op.close()
        got_rows, got_bits = [], []
        kc = keys.index(0)
        for p in out:
            got_rows.extend(p.rows())
            blk = p.get_block(kc).flatten()
            assert blk.type == abi.FLOAT32
            got_bits.extend(None if blk.is_null(i) else b for i, b in enumerate(_bits(blk)))
        want = oracle_agg_rows(pages, keys, aggs)
        assert rows_equal(got_rows, want), keys
        # the key VALUE of a group is the first one seen, bit for bit: -0.0 vs +0.0, and the NaN encoding (FlatHash keeps the first row's bytes)
        first_bits = {}
        og = o.GroupByHash(0, 16)
        for page in pages:
            ids = og.get_group_ids(page, keys)
            blk = page.get_block(0)
            raw = _bits(blk)
            for i, g in enumerate(ids):
                first_bits.setdefault(int(g), None if blk.is_null(i) else raw[i])
        og.close()
        assert got_bits == [first_bits[g] for g in range(len(got_bits))], keys


def test_group_ids_over_real_keys(ctx):
    rng = np.random.default_rng(23)
    g = ops.GroupByHash(ctx, [0, 1], 100)
    og = o.GroupByHash(0, 100)
    for n in (1000, 1, 50000):
        page = Page(Block.real(_reals(rng, n, 300), rng.random(n) < 0.02), Block.integer(rng.integers(0, 4, n)))
        assert (g.get_group_ids(page) == og.get_group_ids(page, [0, 1])).all()
        assert g.get_group_count() == og.group_count()
    g.close(); og.close()


def test_real_is_refused_where_it_is_not_built(ctx):
    page = Page(Block.real(np.array([1.0, 2.0], dtype=np.float32)), Block.bigint([1, 2]))
    f = ops.HashAggregationOperatorFactory(ctx, [1], abi.STEP_SINGLE, [ops.Aggregator(abi.AGG_SUM, 0, -1)], 100, 0)
    op = f.create_operator()
    with pytest.raises(Exception) as e:
        op.add_input(page)
        op.finish()
        op.get_output()
    assert "REAL" in str(e.value)
    op.close()
