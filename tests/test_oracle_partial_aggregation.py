"""Adaptive partial aggregation, CPU side: the oracle controller replays the trajectory of the reference's own tests
(T/operator/TestHashAggregationOperator.java:784-913) and the C-ABI controller (no GPU needed) agrees with it call by call."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "oracle"))
import partial_aggregation as pa  # noqa: E402
from trino_b200 import abi  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402

LONG_PAGE_10 = pa.page_size_in_bytes([("INT64", 10)])
LONG_PAGE_9 = pa.page_size_in_bytes([("INT64", 9)])


def reference_trajectory():
    """(bytes, rows, unique or None, disabled afterwards) - the onFlush calls testAdaptivePartialAggregation makes, page by page:
    maxPartialMemory is 1 byte, so every page flushes; LongArrayBlock and its run-length wrapper both cost 9 bytes a position"""
    t = []
    # operator 1: ten rows with nine distinct keys -> 0.9 > 0.8 and 90 >= 1.5 bytes: disabled; the second page goes through a skipped builder
    t.append((LONG_PAGE_10, 10, 9, True))
    t.append((LONG_PAGE_10, 10, None, True))
    # operator 2: two skipped pages; the second brings the total to 360 >= 300 bytes: re-enabled, counters reset
    t.append((LONG_PAGE_10, 10, None, True))
    t.append((LONG_PAGE_10, 10, None, False))
    # loop i = 1..4 (:832-846): nine unique rows disable it again, three more skipped pages reach 324 >= 300
    t.append((LONG_PAGE_9, 9, 9, True))
    t.append((LONG_PAGE_9, 9, None, True))
    t.append((LONG_PAGE_9, 9, None, True))
    t.append((LONG_PAGE_9, 9, None, False))
    # :849 a late flush from a disabled builder is ignored while enabled
    t.append((1_000_000, 1_000_000, None, False))
    # operator 3: 100 rows -> 1 unique row, twice: stays enabled
    t.append((pa.page_size_in_bytes([("INT64", 100)]), 100, 1, False))
    t.append((pa.page_size_in_bytes([("INT64", 100)]), 100, 1, False))
    return t


def test_page_size_matches_the_reference_constants():
    assert LONG_PAGE_10 == 90 and LONG_PAGE_9 == 81            # (Long.BYTES + Byte.BYTES) * positionCount


def test_oracle_controller_replays_the_reference_test():
    c = pa.PartialAggregationController(1, 0.8)
    assert not c.is_partial_aggregation_disabled()
    for b, r, u, want in reference_trajectory():
        c.on_flush(b, r, u)
        assert c.is_partial_aggregation_disabled() == want


def test_only_flush_triggers_the_switch():
    # testAdaptivePartialAggregationTriggeredOnlyOnFlush :865-913: 12 rows -> 10 unique in ONE flush (10/12 > 0.8)
    c = pa.PartialAggregationController(1, 0.8)
    c.on_flush(pa.page_size_in_bytes([("INT64", 10)]) + pa.page_size_in_bytes([("INT64", 2)]), 12, 10)
    assert c.is_partial_aggregation_disabled()


def _lib():
    return abi.load_library()


def test_c_abi_controller_matches_oracle_call_by_call():
    lib = _lib()
    rng = np.random.default_rng(11)
    for trial in range(50):
        max_mem = int(rng.integers(1, 5000))
        thr = float(rng.random())
        oc = pa.PartialAggregationController(max_mem, thr)
        gc = ops.PartialAggregationController(lib, max_mem, thr)
        for _ in range(200):
            rows = int(rng.integers(1, 10_000))
            b = rows * int(rng.integers(2, 40))
            unique = None if (oc.is_partial_aggregation_disabled() or rng.random() < 0.1) else int(rng.integers(0, rows + 1))
            oc.on_flush(b, rows, unique)
            gc.on_flush(b, rows, unique)
            assert gc.is_partial_aggregation_disabled() == oc.is_partial_aggregation_disabled()
        gc.close()
    gc = ops.PartialAggregationController(lib, 1, 0.8)
    for b, r, u, want in reference_trajectory():
        gc.on_flush(b, r, u)
        assert gc.is_partial_aggregation_disabled() == want
    gc.close()


def test_skip_rows_restatement_known_answers():
    rows = [(1, 5, None, True), (1, None, 2.5, None), (None, 7, 1.0, False)]
    aggs = [(pa.COUNT_STAR, -1, -1), (pa.COUNT, 1, -1), (pa.SUM, 1, -1), (pa.AVG, 2, -1), (pa.MIN, 2, -1), (pa.SUM, 1, 3), (pa.COUNT_STAR, -1, 3)]
    got = pa.skip_aggregation_rows(rows, [0], aggs, double_channels=(2,))
    assert got == [(1, 1, 1, 5, 0, 0.0, None, 5, 1),
                   (1, 1, 0, None, 1, 2.5, 2.5, None, 0),
                   (None, 1, 1, 7, 1, 1.0, 1.0, None, 0)]
