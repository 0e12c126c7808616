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


def test_decimal_states_take_the_reference_wire_form():
    # M/operator/aggregation/state/LongDecimalWithOverflowStateSerializer.java:36-96 and …AndLongStateSerializer.java:36-113
    from trino_b200.page import (decode_decimal_avg_states, decode_decimal_sum_states, encode_decimal_avg_states, encode_decimal_sum_states)
    sums = Block.int128([5, -3, 2**100, None, -(2**70), 0])
    over = Block.bigint([0, 0, 0, 0, 2, 0])
    enc = encode_decimal_sum_states(sums, over)
    # low only; low + high; low + high; NULL; low + high + overflow; low only (zero)
    assert [None if v is None else len(v) for v in enc.to_pylist()] == [8, 16, 16, None, 24, 8]
    assert enc.get(0) == (5).to_bytes(8, "little") and enc.get(1) == (2**64 - 3).to_bytes(8, "little") + (2**64 - 1).to_bytes(8, "little")
    back = decode_decimal_sum_states(enc)
    assert back[0].to_pylist() == sums.to_pylist() and back[1].to_pylist() == [0, 0, 0, 0, 2, 0]
    counts = Block.bigint([1, 2, 1, 0, 7, 1])
    enc = encode_decimal_avg_states(sums, over, counts)
    # count == 1 and no overflow: the decimal alone; else count and overflow follow the decimal's words; count == 0: NULL
    assert [None if v is None else len(v) for v in enc.to_pylist()] == [8, 32, 16, None, 32, 8]
    assert enc.get(1)[16:24] == (2).to_bytes(8, "little") and enc.get(1)[24:32] == bytes(8)
    s, o, c = decode_decimal_avg_states(enc)
    assert s.to_pylist() == sums.to_pylist() and o.to_pylist() == [0, 0, 0, 0, 2, 0] and c.to_pylist() == [1, 2, 1, 0, 7, 1]
    # a 24-byte average state is (low, count, overflow) with a zero high word
    s, o, c = decode_decimal_avg_states(Block.varchar([(9).to_bytes(8, "little") + (4).to_bytes(8, "little") + bytes(8)]))
    assert (s.to_pylist(), o.to_pylist(), c.to_pylist()) == ([9], [0], [4])
