"""Host-side ROW blocks (S/block/RowBlock.java): the flatten / compose pair that carries multi-field aggregation states across the C ABI."""
import numpy as np

from trino_b200.page import Block, Page, RowBlock, compose_row_blocks, flatten_row_blocks


def test_flatten_then_compose_is_identity_on_values():
    n = 6
    row = RowBlock([Block.bigint(np.arange(n)), Block.double(np.arange(n) * 0.5, np.array([0, 1, 0, 0, 0, 0], bool))])
    page = Page(Block.bigint(np.arange(n) + 10), row, Block.varchar(["a", None, "c", "d", "e", "f"]))
    flat, first = flatten_row_blocks(page)
    assert first == [0, 1, 3] and flat.channel_count == 4
    back = compose_row_blocks(flat, [1, 2, 1])
    assert back.rows() == page.rows()
    assert back.rows()[1] == (11, (1, None), None)


def test_null_rows_read_as_null_fields():
    row = RowBlock([Block.bigint([1, 2, 3]), Block.double([1.0, 2.0, 3.0])], row_nulls=[False, True, False])
    assert row.to_pylist() == [(1, 1.0), None, (3, 3.0)]
    flat, _ = flatten_row_blocks(Page(row))
    assert flat.rows() == [(1, 1.0), (None, None), (3, 3.0)]
