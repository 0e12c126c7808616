"""The page wire format through the C ABI against oracle/serde.py
(byte-exact for pages whose NULL-free columns carry no validity bitmap) and as a round trip."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import serde  # noqa: E402

from trino_b200 import abi  # noqa: E402
from trino_b200.page import AbiPage, Block, DictionaryBlock, Page  # noqa: E402

pytestmark = pytest.mark.gpu


def gpu_serialize(ctx, page):
    ap = AbiPage(page)
    cap = 64 + sum(16 + 9 * page.position_count for _ in range(page.channel_count)) + 64 * page.position_count
    buf = np.zeros(cap, dtype=np.uint8)
    n = C.c_int64()
    ctx.check(ctx.lib.tgpu_page_serialize(ctx.h, ap.ref(), buf.ctypes.data_as(C.c_void_p), cap, C.byref(n)))
    return bytes(buf[:n.value])


def gpu_deserialize(ctx, data, types):
    arr = np.frombuffer(data, dtype=np.uint8).copy()
    t = (C.c_int32 * len(types))(*types)
    pp = abi.PP()
    ctx.check(ctx.lib.tgpu_page_deserialize(ctx.h, arr.ctypes.data_as(C.c_void_p), len(arr), t, len(types), C.byref(pp)))
    return ctx.page_to_host(pp)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 64, 1000, 100_003])
def test_wire_bytes_match_oracle_and_round_trip(ctx, n):
    rng = np.random.default_rng(n)
    nulls = lambda p: (rng.random(n) < p) if n else None   # noqa: E731
    words = ["w%d" % i * (i % 4) for i in range(n)]
    page = Page(Block.bigint(rng.integers(-2**62, 2**62, n), nulls(0.3)), Block.double(rng.normal(size=n)), Block.integer(rng.integers(-9, 9, n), nulls(0.5)),
                Block.smallint(rng.integers(-300, 300, n), nulls(1.0)), Block.tinyint(rng.integers(-5, 5, n)), Block.varchar([None if i % 3 == 0 else w for i, w in enumerate(words)]),
                DictionaryBlock(Block.bigint([10, 20, 30]), rng.integers(0, 3, n)), position_count=n)
    want = serde.serialize_page(page)
    got = gpu_serialize(ctx, page)
    assert got == want
    types = [abi.INT64, abi.FLOAT64, abi.INT32, abi.INT16, abi.INT8, abi.UTF8, abi.INT64]
    back = gpu_deserialize(ctx, want, types)
    assert back.rows() == page.rows()


def test_golden_sizes(ctx):
    # TestPagesSerde.testBigintSerializedSize / testVarcharSerializedSize
    assert len(gpu_serialize(ctx, Page(Block.bigint([123, 456])))) == 35 + 16
    assert len(gpu_serialize(ctx, Page(Block.varchar(["alice", "bob"])))) == 43 + 9 + 7


@pytest.mark.parametrize("n", [0, 1, 9, 5000])
def test_int128_and_real_columns(ctx, n):
    """INT128_ARRAY (long DECIMAL, S/block/Int128ArrayBlockEncoding.java:52-84) and REAL as INT_ARRAY: byte-exact against the oracle, and back"""
    rng = np.random.default_rng(n + 77)
    wide = [int(x) * (1 << 70) + int(y) for x, y in zip(rng.integers(-2**50, 2**50, n), rng.integers(0, 2**62, n))]
    page = Page(Block.int128(wide, (rng.random(n) < 0.4) if n else None), Block.int128(wide[::-1]), Block.real(rng.normal(size=n).astype(np.float32), (rng.random(n) < 0.2) if n else None),
                position_count=n)
    want = serde.serialize_page(page)
    assert gpu_serialize(ctx, page) == want
    back = gpu_deserialize(ctx, want, [abi.INT128, abi.INT128, abi.FLOAT32])
    assert back.rows() == page.rows()
