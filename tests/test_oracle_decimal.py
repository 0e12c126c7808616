"""Long DECIMAL on the CPU side: the oracle's DecimalSumAggregation restatement replays the reference's own state tests
(T/operator/aggregation/TestDecimalSumAggregation.java:36-137), agrees with exact integer arithmetic, and LongDecimalType's hash is the
XOR of two XXH64 values (S/type/LongDecimalType.java:203-229)."""
import struct

import numpy as np
import pytest

import oracle_lib as o
from trino_b200.page import Block, Page

TWO = 2


def words(state):
    return int(state.decimal[0]), int(state.decimal[1])


def test_reference_overflow_and_underflow_cases():
    s = o.DecimalSumState().add([TWO**126])                       # testOverflow :36-50
    assert s.overflow[0] == 0 and s.value == TWO**126
    s.add([TWO**126])
    assert s.overflow[0] == 1 and words(s) == (-(1 << 63), 0)     # Int128.valueOf(1L << 63, 0)
    s = o.DecimalSumState().add([-(TWO**126)])                    # testUnderflow :52-66
    assert s.overflow[0] == 0 and s.value == -(TWO**126)
    s.add([-(TWO**126)])
    assert s.overflow[0] == 0 and words(s) == (-(1 << 63), 0)     # Int128.valueOf(0x8000000000000000L, 0)
    s = o.DecimalSumState().add([TWO**126, TWO**126, TWO**125])   # testUnderflowAfterOverflow :68-86
    assert s.overflow[0] == 1 and words(s) == (((1 << 63) | (1 << 61)) - (1 << 64), 0)
    s.add([-(TWO**126)] * 3)
    assert s.overflow[0] == 0 and s.value == -(TWO**125)


def test_reference_combine_cases():
    a = o.DecimalSumState().add([TWO**125, TWO**126])             # testCombineOverflow :88-104
    b = o.DecimalSumState().add([TWO**125, TWO**126])
    a.combine(b)
    assert a.overflow[0] == 1 and words(a) == (0xC000000000000000 - (1 << 64), 0)
    a = o.DecimalSumState().add([-(TWO**125), -(TWO**126)])       # testCombineUnderflow :106-122
    b = o.DecimalSumState().add([-(TWO**125), -(TWO**126)])
    a.combine(b)
    assert a.overflow[0] == -1 and words(a) == (0x4000000000000000, 0)


def test_reference_overflow_on_output():
    s = o.DecimalSumState().add([TWO**126, TWO**126])             # testOverflowOnOutput :124-137
    with pytest.raises(OverflowError):
        s.output()
    assert o.DecimalSumState().output() is None
    assert o.DecimalSumState().add([10**38 - 1]).output() == 10**38 - 1
    with pytest.raises(OverflowError):
        o.DecimalSumState().add([10**38 - 1, 1]).output()         # Decimals.overflows: outside +-(10^38 - 1)
    with pytest.raises(OverflowError):
        o.DecimalSumState().add([-(10**38 - 1), -1]).output()


def test_state_is_the_exact_total_split_at_128_bits():
    # total = signed128(decimal) + overflow * 2^128 after every step, for long and short inputs
    rng = np.random.default_rng(3)
    for _ in range(200):
        vals = [int(rng.integers(-2**62, 2**62)) * int(rng.integers(1, 2**62)) * int(rng.integers(1, 8)) for _ in range(int(rng.integers(1, 40)))]
        vals = [max(-(2**127), min(2**127 - 1, v)) for v in vals]
        s = o.DecimalSumState().add(vals)
        total = sum(vals)
        signed = s.value
        assert -(2**127) <= signed < 2**127 and signed + int(s.overflow[0]) * 2**128 == total
        shorts = [int(x) for x in rng.integers(-2**63, 2**63 - 1, 50)]
        t = o.DecimalSumState().add(shorts, short=True)
        assert t.value == sum(shorts) and t.overflow[0] == 0


def test_long_decimal_hash_is_the_xor_of_two_xxh64():
    for v in (0, 1, -1, 10**30, -(10**37) - 12345, 2**127 - 1, -(2**127)):
        high, low = o.int128_words(v)
        want = o.xxh64(struct.pack("<q", high)) ^ o.xxh64(struct.pack("<q", low))
        page = Page(Block.int128([v]))
        assert int(o.row_hashes(page, [0])[0]) & o.M64 == want
        assert o.load().orc_xxh64_long(high) == o.xxh64(struct.pack("<q", high))


def test_reference_decimal_average_cases():
    # T/operator/aggregation/TestDecimalAverageAggregation.java:46-216
    MIN = -(10**38 - 1)
    assert o.DecimalSumState().add([TWO**126, TWO**126]).average(2) == TWO**126                       # testOverflow
    s = o.DecimalSumState().add([MIN, MIN])                                                           # testUnderflow
    assert s.overflow[0] == -1 and words(s) == (0x698966AF4AF2770B, 0xECEBBB8000000002 - (1 << 64)) and s.average(2) == MIN
    s = o.DecimalSumState().add([TWO**126, TWO**126, TWO**125] + [-(TWO**126)] * 3)                  # testUnderflowAfterOverflow
    assert s.average(6) == -((TWO**125) // 6)                                                         # BigInteger.divide truncates; the fraction is 1/3
    a = o.DecimalSumState().add([TWO**125, TWO**126]).combine(o.DecimalSumState().add([TWO**125, TWO**126]))
    assert a.average(4) == (2 * TWO**126 + 2 * TWO**125) // 4                                         # testCombineOverflow
    a = o.DecimalSumState().add([-(TWO**125), -(TWO**126)]).combine(o.DecimalSumState().add([-(TWO**125), -(TWO**126)]))
    assert a.overflow[0] == -1 and words(a) == (1 << 62, 0) and a.average(4) == -((2 * TWO**126 + 2 * TWO**125) // 4)
    # testNoOverflow: HALF_UP of the exact quotient
    for numbers, want in (([10**37, 0], 5 * 10**36), ([2, 1], 2), ([0, 1], 1), ([-2, -1], -2), ([-1, 0], -1), ([-1, 0, 0], 0), ([-2, 0, 0], -1),
                          ([-2, 0], -1), ([200, 100], 150), ([0, 100], 50), ([-200, -100], -150), ([-100, 0], -50)):
        assert o.DecimalSumState().add(numbers).average(len(numbers)) == want, numbers
    assert o.DecimalSumState().average(0) is None
