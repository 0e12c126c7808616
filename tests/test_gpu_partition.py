"""GPU parity: PagePartitioner (partition ids, per-partition rows and their order) vs the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as o
from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, DictionaryBlock, Page, RunLengthEncodedBlock

pytestmark = pytest.mark.gpu


def gpu_partition(ctx, op, page):
    op.add_input(page)
    got = {}
    while True:
        r = op.get_output_with_partition()
        if r is None:
            break
        got[r[0]] = r[1].rows()
    return got


def oracle_partition(page, keys, buckets, b2p, pcount, null_channel, any_row, state):
    lists, state = o.partition_positions(page, keys, buckets, b2p, pcount, null_channel, any_row, state)
    rows = page.rows()
    return {p: [rows[i] for i in l] for p, l in enumerate(lists) if len(l)}, state


def test_partition_ids_all_key_types(ctx):
    rng = np.random.default_rng(4)
    n = 20000
    page = Page(Block.bigint(rng.integers(-10**12, 10**12, n), rng.random(n) < 0.05), Block.double(rng.normal(size=n)), Block.integer(rng.integers(-99, 99, n)),
                Block.varchar(["k%d" % (i % 97) if i % 13 else None for i in range(n)]), Block.tinyint(rng.integers(-3, 3, n)))
    for keys in ([0], [1], [2, 4], [3], [0, 1, 2, 3, 4]):
        for buckets in (1, 2, 8, 37, 256):
            op = ops.PartitionedOutputOperatorFactory(ctx, keys, buckets).create_operator()
            assert (op.get_partitions(page) == o.partition_ids(page, keys, buckets)).all(), (keys, buckets)
            op.close()
    b2p = [3, 1, 0, 2, 1, 0, 3, 2]
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], 8, b2p).create_operator()
    assert (op.get_partitions(page) == o.partition_ids(page, [0], 8, b2p)).all()
    op.close()


@pytest.mark.parametrize("null_channel,any_row", [(-1, False), (0, False), (-1, True), (0, True)])
def test_partitioned_rows_and_order(ctx, null_channel, any_row):
    rng = np.random.default_rng(8)
    P = 8
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], P, None, null_channel, any_row).create_operator()
    state = False
    for n in (5000, 3, 1, 15, 16, 4096):        # 3, 1, 15 rows < 2 x partitions take the row-wise strategy
        page = Page(Block.bigint(rng.integers(0, 10**6, n), rng.random(n) < 0.2), Block.double(rng.normal(size=n)),
                    Block.varchar(["v%d" % i if i % 5 else None for i in range(n)]))
        want, state = oracle_partition(page, [0], P, None, P, null_channel, any_row, state)
        assert gpu_partition(ctx, op, page) == want, n
    op.close()


def test_single_partition_and_dictionary_keys(ctx):
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], 1).create_operator()
    page = Page(Block.bigint([3, 1, 2]), Block.varchar(["a", "b", None]))
    assert gpu_partition(ctx, op, page) == {0: page.rows()}
    op.close()
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], 4).create_operator()
    d = Page(DictionaryBlock(Block.bigint([100, 200, 300]), np.arange(3000) % 3))
    want, _ = oracle_partition(d, [0], 4, None, 4, -1, False, False)
    assert gpu_partition(ctx, op, d) == want
    op.close()


def test_partition_constants(ctx):
    """A negative partition channel takes its value from partitionConstants: the reference hashes a RunLengthEncodedBlock of the constant
    in that position of the function page (M/operator/output/PagePartitioner.java:436-451; TestPagePartitioner.java:309-379 runs the shape
    with a test function) - so does the oracle here; the library takes the constant's type hash once."""
    rng = np.random.default_rng(12)
    n = 9000
    page = Page(Block.bigint(rng.integers(-10**9, 10**9, n), rng.random(n) < 0.1), Block.double(rng.normal(size=n)))
    constants = {"bigint": Block.bigint([1]), "null": Block.bigint([7], [True]), "double": Block.double([-0.0]), "varchar": Block.varchar(["x-ray"]),
                 "integer": Block.integer(np.array([-5], dtype=np.int32)), "decimal": Block.int128([-(10**30)])}
    for name, const in constants.items():
        for channels, consts in (([-1], [const]), ([0, -1], [None, const]), ([-1, 0, 1], [const, None, None])):
            function_page = Page(*[RunLengthEncodedBlock(consts[i], n) if ch < 0 else page.get_block(ch) for i, ch in enumerate(channels)])
            for buckets in (2, 8, 37):
                want = o.partition_ids(function_page, list(range(len(channels))), buckets)
                op = ops.PartitionedOutputOperatorFactory(ctx, channels, buckets, partition_constants=consts).create_operator()
                assert (op.get_partitions(page) == want).all(), (name, channels, buckets)
                if len(channels) == 1:
                    assert len(set(want.tolist())) == 1          # a lone constant sends every row to one partition (:315-330)
                # the partitioned pages follow the same ids (rows and their order per partition)
                got = gpu_partition(ctx, op, page)
                rows = page.rows()
                assert got == {p: [rows[i] for i in np.flatnonzero(want == p)] for p in sorted(set(want.tolist()))}, (name, channels, buckets)
                op.close()
    with pytest.raises(Exception):      # a constant channel without constants (PagePartitioner.java:111 checkArgument)
        ops.PartitionedOutputOperatorFactory(ctx, [-1], 4).create_operator()


@pytest.mark.parametrize("P", [1, 2, 8, 64])
def test_local_partition_generator(ctx, P):
    """LocalPartitionGenerator (M/operator/exchange/LocalPartitionGenerator.java:45-77): (int) XxHash64.hash(Long.reverse(rawHash)) & (P - 1)
    over the row hash of the hash channels - ids, and the pages a PartitioningExchanger-style split makes of them."""
    rng = np.random.default_rng(13)
    n = 30000
    page = Page(Block.bigint(rng.integers(-10**12, 10**12, n), rng.random(n) < 0.05), Block.double(rng.normal(size=n)), Block.integer(rng.integers(-99, 99, n)),
                Block.varchar(["k%d" % (i % 97) if i % 13 else None for i in range(n)]))
    for channels in ([0], [1], [3], [0, 2, 3]):
        gen = ops.LocalPartitionGenerator(ctx, channels, P)
        want = o.local_partition_ids(page, channels, P)
        assert (gen.get_partitions(page) == want).all(), channels
        gen.close()
    op = ops.PartitionedOutputOperatorFactory(ctx, [0], P, partition_function=abi.PARTITION_LOCAL).create_operator()
    fixed = Page(page.get_block(0), page.get_block(1))          # fixed-width: the multi-split path recomputes the id from the key
    want = o.local_partition_ids(fixed, [0], P)
    rows = fixed.rows()
    assert gpu_partition(ctx, op, fixed) == {p: [rows[i] for i in np.flatnonzero(want == p)] for p in sorted(set(want.tolist()))}
    op.close()
    with pytest.raises(ValueError):
        ops.LocalPartitionGenerator(ctx, [0], 6)


@pytest.mark.parametrize("P", [2, 8, 37, 64, 100])
def test_fixed_width_pages_take_the_multisplit_path(ctx, P):
    """Fixed-width pages without replicated rows are split by one stable histogram/scatter pass (<= 64 partitions) or by the
    sort path (100): identical per-partition rows and order either way, NULLs of value columns preserved."""
    rng = np.random.default_rng(P)
    n = 300_000
    page = Page(Block.bigint(rng.integers(0, 10**7, n), rng.random(n) < 0.1), Block.double(rng.normal(size=n), rng.random(n) < 0.3),
                Block.integer(rng.integers(-5, 5, n)), Block.tinyint(rng.integers(0, 3, n), rng.random(n) < 0.5), Block.smallint(rng.integers(0, 999, n)))
    b2p = [int(x) for x in rng.integers(0, P, 4 * P)]
    for keys, buckets, mapping in (([0], P, None), ([0, 2], 4 * P, b2p)):
        op = ops.PartitionedOutputOperatorFactory(ctx, keys, buckets, mapping, -1, False).create_operator()
        want, _ = oracle_partition(page, keys, buckets, mapping, P, -1, False, False)
        assert gpu_partition(ctx, op, page) == want
        op.close()
