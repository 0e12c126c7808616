"""Pins the oracle's hash arithmetic: reference golden vectors, published XXH64 vectors, and an independent
big-integer Python restatement of the formulas in SURVEY.md Appendix A."""
import numpy as np

import oracle_lib as o
from helpers import reference_cases
from trino_b200.page import Block, DictionaryBlock, Page, RunLengthEncodedBlock

M = (1 << 64) - 1
P1, P2 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F


def rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M


def py_hash_long(v):
    return (rotl(((v & M) * P2) & M, 31) * P1) & M


def py_murmur3(x):
    x &= M
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & M
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & M
    x ^= x >> 33
    return x


def test_xxh64_golden_vectors():
    for case in reference_cases()["xxhash64"]:
        got = o.xxh64(case["input_utf8"].encode(), case.get("seed", 0))
        assert got == int(case["expected_hex"], 16), case["source"]


def test_xxh64_long_is_xxh64_of_le_bytes():
    lib = o.load()
    for v in [0, 1, -1, 42, 2**62, -2**63]:
        assert lib.orc_xxh64_long(v) == o.xxh64(int(v & M).to_bytes(8, "little"))


def test_xxh64_all_length_classes():
    # exercises the >=32-byte stripe loop, the 8/4/1-byte tails, against a pure-Python restatement
    P3, P4, P5 = 0x165667B19E3779F9, 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5

    def py_xxh64(b, seed=0):
        n = len(b)
        i = 0
        if n >= 32:
            v = [(seed + P1 + P2) & M, (seed + P2) & M, seed, (seed - P1) & M]
            while i <= n - 32:
                for k in range(4):
                    lane = int.from_bytes(b[i + 8 * k:i + 8 * k + 8], "little")
                    v[k] = (rotl((v[k] + lane * P2) & M, 31) * P1) & M
                i += 32
            h = (rotl(v[0], 1) + rotl(v[1], 7) + rotl(v[2], 12) + rotl(v[3], 18)) & M
            for k in range(4):
                h ^= (rotl((v[k] * P2) & M, 31) * P1) & M
                h = (h * P1 + P4) & M
        else:
            h = (seed + P5) & M
        h = (h + n) & M
        while i + 8 <= n:
            lane = int.from_bytes(b[i:i + 8], "little")
            h ^= (rotl((lane * P2) & M, 31) * P1) & M
            h = (rotl(h, 27) * P1 + P4) & M
            i += 8
        if i + 4 <= n:
            h ^= (int.from_bytes(b[i:i + 4], "little") * P1) & M
            h = (rotl(h, 23) * P2 + P3) & M
            i += 4
        while i < n:
            h ^= (b[i] * P5) & M
            h = (rotl(h, 11) * P1) & M
            i += 1
        h ^= h >> 33
        h = (h * P2) & M
        h ^= h >> 29
        h = (h * P3) & M
        h ^= h >> 32
        return h

    for case in reference_cases()["xxhash64"]:
        assert py_xxh64(case["input_utf8"].encode(), case.get("seed", 0)) == int(case["expected_hex"], 16)
    rng = np.random.default_rng(7)
    for n in list(range(0, 70)) + [127, 128, 129, 1000]:
        b = bytes(rng.integers(0, 256, size=n, dtype=np.uint8))
        assert o.xxh64(b) == py_xxh64(b), n


def test_hash_long_and_murmur3_formulas():
    lib = o.load()
    rng = np.random.default_rng(1)
    vals = [0, 1, -1, 2**63 - 1, -2**63] + [int(x) for x in rng.integers(-2**63, 2**63 - 1, size=200, dtype=np.int64)]
    for v in vals:
        assert lib.orc_hash_long(v) == py_hash_long(v)
        assert lib.orc_murmur3(v & M) == py_murmur3(v)
    # PagesHash.getHashPosition known points (mix(0) == 0 by construction)
    assert lib.orc_murmur3(0) == 0


def test_hash_double_canonicalises_zero_and_nan():
    lib = o.load()
    assert lib.orc_hash_double(0.0) == lib.orc_hash_double(-0.0) == py_hash_long(0)
    nan1 = np.frombuffer(np.uint64(0x7FF8000000000001).tobytes(), dtype=np.float64)[0]
    assert lib.orc_hash_double(float("nan")) == lib.orc_hash_double(float(nan1)) == py_hash_long(0x7FF8000000000000)
    bits = np.frombuffer(np.float64(1.5).tobytes(), dtype=np.int64)[0]
    assert lib.orc_hash_double(1.5) == py_hash_long(int(bits))


def test_hash_real_follows_realtype():
    """S/type/RealType.java:151-159: AbstractLongType.hash(floatToIntBits(v)), -0.0 collapsed to +0.0; floatToIntBits makes every NaN
    0x7fc00000 and the int widens to long WITH its sign.  The reference holds no numeric golden for it (parity of the value unpinned,
    as for hash_long); its own test pins the property that all NaN encodings hash alike (M/test/.../type/TestRealType.java:63-79),
    with the very bit patterns used here."""
    lib = o.load()
    from trino_b200.page import Block, Page
    assert lib.orc_hash_real(0.0) == lib.orc_hash_real(-0.0) == py_hash_long(0)
    nan_bits = [0x7FC00000, 0xFFC00000, 0x7FC00001, 0x7F800001, (-0x400000) & 0xFFFFFFFF]        # -0x400000, 0x7fc00000: TestRealType.java:70-71
    nans = np.array(nan_bits, dtype=np.uint32).view(np.float32)
    page = Page(Block.real(nans))
    hashes = o.row_hashes(page, [0])
    assert len(set(hashes.tolist())) == 1 and (int(hashes[0]) & M) == py_hash_long(0x7FC00000)
    for v in (1.5, -1.5, 3.4028235e38, 1e-45, -2.0):
        bits = int(np.float32(v).view(np.int32))                  # negative floats: the int is negative and sign-extends
        assert lib.orc_hash_real(v) == py_hash_long(bits), v
    # a REAL column hashes through the page path like the scalar does, NULL -> 0
    vals = np.array([1.5, -0.0, 0.0, -7.25], dtype=np.float32)
    hashes = o.row_hashes(Page(Block.real(vals, [False, False, False, True])), [0])
    assert [int(h) & M for h in hashes] == [lib.orc_hash_real(1.5), py_hash_long(0), py_hash_long(0), 0]


def test_array_size_and_load_factors():
    lib = o.load()
    # HashCommon.arraySize(n, f) = max(2, nextPowerOfTwo(ceil(n / f)))
    assert lib.orc_array_size(1, 0.75) == 2
    assert lib.orc_array_size(3, 0.75) == 4
    assert lib.orc_array_size(100, 0.75) == 256
    assert lib.orc_array_size(96, 0.75) == 128
    assert lib.orc_array_size(2**30, 0.75) == -1
    # IncrementalLoadFactorHashArraySizeSupplier thresholds (T/operator/TestIncrementalLoadFactorHashArraySizeSupplier.java)
    assert lib.orc_join_hash_array_size(65536) == lib.orc_array_size(65536, 0.25) == 262144
    assert lib.orc_join_hash_array_size(65537) == lib.orc_array_size(65537, 0.5)
    assert lib.orc_join_hash_array_size(1048576) == lib.orc_array_size(1048576, 0.5)
    assert lib.orc_join_hash_array_size(1048577) == lib.orc_array_size(1048577, 0.75)


def test_process_raw_hash_range_and_formula():
    lib = o.load()
    rng = np.random.default_rng(2)
    for h in rng.integers(-2**63, 2**63 - 1, size=500, dtype=np.int64):
        h = int(h)
        for count in (1, 2, 7, 8, 256, 1000):
            x = ((h & M) ^ ((h & M) >> 32)) & 0xFFFFFFFF
            assert lib.orc_process_raw_hash(h, count) == (x * count) >> 32
            p = lib.orc_local_partition(h, 8)
            assert 0 <= p < 8


def test_row_hash_equals_manual_fold_incl_rle_and_dictionary():
    # T/operator/TestInterpretedHashGenerator.java:60-89: batched == single-position == manual fold for flat, RLE, dictionary
    lib = o.load()
    n = 64
    rng = np.random.default_rng(3)
    a = Block.bigint(rng.integers(-100, 100, n), rng.random(n) < 0.2)
    b = Block.double(rng.normal(size=n), rng.random(n) < 0.2)
    c = Block.varchar([None if i % 7 == 0 else "v%d" % (i % 5) for i in range(n)])
    d = Block.integer(rng.integers(-5, 5, n))
    page = Page(a, b, c, d)
    hashes = o.row_hashes(page, [0, 1, 2, 3])
    for i in range(n):
        h = 0
        for blk, kind in ((a, "long"), (b, "double"), (c, "varchar"), (d, "long")):
            v = blk.get(i)
            if v is None:
                th = 0
            elif kind == "long":
                th = py_hash_long(v)
            elif kind == "double":
                th = lib.orc_hash_double(v)
            else:
                th = o.xxh64(v)
            h = (31 * h + th) & M
        assert int(hashes[i]) & M == h
    dict_page = Page(DictionaryBlock(Block.bigint([5, 6, 7]), [0, 2, 1, 1, 0]), RunLengthEncodedBlock(Block.bigint([9]), 5))
    flat_page = Page(Block.bigint([5, 7, 6, 6, 5]), Block.bigint([9] * 5))
    assert (o.row_hashes(dict_page, [0, 1]) == o.row_hashes(flat_page, [0, 1])).all()
