"""GPU parity: PagePartitioner (partition ids, per-partition rows and their order) vs the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as o
from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, DictionaryBlock, Page

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
