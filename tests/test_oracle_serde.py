"""The page wire format restatement (oracle/serde.py) against the reference's golden sizes (TestPagesSerde) and its own round trips.
SURVEY.md §8(f) rank 1; the device kernels for it are next round's work, the oracle is pinned now."""
import os
import struct
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import serde  # noqa: E402

from trino_b200 import abi  # noqa: E402
from trino_b200.page import Block, DictionaryBlock, Page  # noqa: E402


def test_bigint_serialized_size_goldens():
    # TestPagesSerde.testBigintSerializedSize :183-204
    assert len(serde.serialize_page(Page(Block.bigint([]), position_count=0))) == 35
    assert len(serde.serialize_page(Page(Block.bigint([123])))) == 35 + 8
    assert len(serde.serialize_page(Page(Block.bigint([123, 456])))) == 35 + 8 + 8


def test_varchar_serialized_size_goldens():
    # TestPagesSerde.testVarcharSerializedSize :207-228
    assert len(serde.serialize_page(Page(Block.varchar([]), position_count=0))) == 43
    assert len(serde.serialize_page(Page(Block.varchar(["alice"])))) == 43 + 4 + 5
    assert len(serde.serialize_page(Page(Block.varchar(["alice", "bob"])))) == 43 + 4 + 5 + 4 + 3


def test_layout_field_by_field():
    data = serde.serialize_page(Page(Block.bigint([7, None, -1, None, 5, 6, 8, 9, None])))
    n, unc, comp = struct.unpack_from("<iii", data, 0)
    assert (n, unc, comp) == (9, len(data) - 12, len(data) - 12)                     # PagesSerdeUtil.java:44-48
    assert struct.unpack_from("<i", data, 12) == (1,)                                # channel count
    assert struct.unpack_from("<i", data, 16) == (10,) and data[20:30] == b"LONG_ARRAY"
    assert struct.unpack_from("<i", data, 30) == (9,) and data[34] == 1
    assert data[35:37] == bytes([0b01010000, 0b10000000])                            # EncoderUtil.java:46-67: MSB first
    assert struct.unpack_from("<i", data, 37) == (6,)                                # non-null count, LongArrayBlockEncoding.java:131
    assert struct.unpack_from("<6q", data, 41) == (7, -1, 5, 6, 8, 9)
    assert len(data) == 41 + 48


def test_round_trip_all_encodings():
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 8, 9, 1000):
        nulls = lambda p: (rng.random(n) < p) if n else None   # noqa: E731
        words = ["w%d" % i * (i % 4) for i in range(n)]
        page = Page(Block.bigint(rng.integers(-2**62, 2**62, n), nulls(0.3)), Block.double(rng.normal(size=n), nulls(0.0)), Block.integer(rng.integers(-9, 9, n), nulls(0.5)),
                    Block.smallint(rng.integers(-300, 300, n), nulls(1.0)), Block.tinyint(rng.integers(-5, 5, n)), Block.varchar([None if i % 3 == 0 else w for i, w in enumerate(words)]),
                    DictionaryBlock(Block.bigint([10, 20, 30]), rng.integers(0, 3, n)), position_count=n)
        data = serde.serialize_page(page)
        count, cols = serde.deserialize_columns(data)
        assert count == n and [c[0] for c in cols] == ["LONG_ARRAY", "LONG_ARRAY", "INT_ARRAY", "SHORT_ARRAY", "BYTE_ARRAY", "VARIABLE_WIDTH", "LONG_ARRAY"]
        flat = [page.get_block(c).flatten() for c in range(7)]
        for (name, values, got_nulls), blk in zip(cols, flat):
            want_nulls = blk.nulls if blk.nulls is not None and blk.nulls.any() else None
            assert (got_nulls is None) == (want_nulls is None)
            if want_nulls is not None:
                assert (got_nulls == want_nulls).all()
            if name == "VARIABLE_WIDTH":
                offsets, payload = values
                got = [None if (got_nulls is not None and got_nulls[i]) else payload[offsets[i]:offsets[i + 1]] for i in range(n)]
                assert got == blk.to_pylist()
            else:
                keep = np.ones(n, dtype=bool) if want_nulls is None else ~want_nulls
                assert (values.view(blk.values.dtype)[keep] == blk.values[keep]).all()
        # serialising what was read gives the same bytes (idempotence)
        again = []
        for name, values, got_nulls in cols:
            again.append(("var", values[0], values[1], got_nulls) if name == "VARIABLE_WIDTH" else ("fixed", values, got_nulls))
        assert serde.serialize_columns(count, again) == data


def test_int128_and_real_blocks():
    """long DECIMAL travels as INT128_ARRAY (S/block/Int128ArrayBlockEncoding.java:52-84: positionCount | nulls | [nonNullCount] | two
    longs per non-NULL position, the high word first), REAL as INT_ARRAY of the raw float bits."""
    import struct
    page = Page(Block.int128([1, None, (1 << 64) + 5]), Block.real(np.array([1.5, -0.0, 2.0], dtype=np.float32)))
    data = serde.serialize_page(page)
    body = data[serde.HEADER_SIZE:]
    name = b"INT128_ARRAY"
    want = struct.pack("<i", 2) + struct.pack("<i", len(name)) + name + struct.pack("<i", 3) + b"\x01" + bytes([0b01000000]) + struct.pack("<i", 2) \
        + struct.pack("<qqqq", 0, 1, 1, 5)
    assert body[:len(want)] == want
    name2 = b"INT_ARRAY"
    want2 = struct.pack("<i", len(name2)) + name2 + struct.pack("<i", 3) + b"\x00" + np.array([1.5, -0.0, 2.0], dtype=np.float32).tobytes()
    assert body[len(want):] == want2
    count, cols = serde.deserialize_columns(data)
    assert count == 3 and [c[0] for c in cols] == ["INT128_ARRAY", "INT_ARRAY"]
    assert cols[0][1].tolist() == [[0, 1], [0, 0], [1, 5]] and cols[0][2].tolist() == [False, True, False]
    assert serde.serialize_columns(count, [("fixed", cols[0][1], cols[0][2]), ("fixed", cols[1][1], cols[1][2])]) == data
    # without NULLs: no count, 16 bytes per position
    plain = serde.serialize_page(Page(Block.int128([-1, 1 << 100])))
    assert len(plain) == serde.HEADER_SIZE + 4 + 4 + len(name) + 4 + 1 + 32
    assert plain[-32:] == struct.pack("<qqqq", -1, -1, 1 << 36, 0)
