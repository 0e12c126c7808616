"""CPU restatement of the reference's adaptive partial aggregation — TEST INFRASTRUCTURE ONLY (like everything under oracle/).

PartialAggregationController follows M/operator/aggregation/partial/PartialAggregationController.java:35-103 statement by statement;
skip_aggregation_rows follows SkipAggregationBuilder.buildOutputPage (M/operator/aggregation/partial/SkipAggregationBuilder.java:103-131):
every position is its own group, the output row is the key values followed by each aggregate's intermediate state over that one row,
in the flat state layout include/trino_gpu.h documents (count: BIGINT; sum/min/max: the value or NULL; avg: BIGINT count, DOUBLE sum).
page_size_in_bytes is Page.getSizeInBytes() for the block kinds the tests use (S/block/LongArrayBlock.java:35,75-78,
S/block/RunLengthEncodedBlock.java:103-106).  Pinned on T/operator/TestHashAggregationOperator.java:784-913
(tests/test_oracle_partial_aggregation.py replays the controller trajectory of those two tests).
"""

DISABLE_FACTOR = 1.5
ENABLE_FACTOR = DISABLE_FACTOR * 200

COUNT_STAR, COUNT, SUM, AVG, MIN, MAX = 0, 1, 2, 3, 4, 5      # tgpu_agg_function


class PartialAggregationController:
    def __init__(self, max_partial_memory, unique_rows_ratio_threshold):
        self.max_partial_memory = max_partial_memory
        self.threshold = unique_rows_ratio_threshold
        self.disabled = False
        self.total_bytes = self.total_rows = self.total_unique = 0

    def is_partial_aggregation_disabled(self):
        return self.disabled

    def on_flush(self, bytes_processed, rows_processed, unique_rows_produced=None):
        if not self.disabled and unique_rows_produced is None:
            return                                   # :69-72 when PA is re-enabled, ignore stats from disabled flushes
        self.total_bytes += bytes_processed
        self.total_rows += rows_processed
        if unique_rows_produced is not None:
            self.total_unique += unique_rows_produced
        if not self.disabled and self._should_disable():
            self.disabled = True
        if self.disabled and self.total_bytes >= self.max_partial_memory * ENABLE_FACTOR:
            self.total_bytes = self.total_rows = self.total_unique = 0
            self.disabled = False

    def _should_disable(self):
        return (self.total_bytes >= self.max_partial_memory * DISABLE_FACTOR
                and (self.total_unique / self.total_rows) > self.threshold)


_BYTES_PER_POSITION = {"INT64": 9, "FLOAT64": 9, "INT32": 5, "INT16": 3, "INT8": 2}


def page_size_in_bytes(columns):
    """columns: [(type name, positions)] - a run-length block counts its value's size once per position"""
    return sum(_BYTES_PER_POSITION[t] * n for t, n in columns)


def skip_aggregation_rows(rows, key_channels, aggs, double_channels=()):
    """rows: tuples with None for NULL; aggs: (function, input channel or -1, mask channel or -1); -> one output tuple per row"""
    out = []
    for r in rows:
        o = [r[c] for c in key_channels]
        for fn, ch, mask in aggs:
            on = True
            if mask >= 0:
                on = r[mask] is not None and bool(r[mask])
            v = r[ch] if ch >= 0 else None
            if fn != COUNT_STAR and v is None:
                on = False
            if fn in (COUNT_STAR, COUNT):
                o.append(1 if on else 0)
            elif fn == AVG:
                o.append(1 if on else 0)
                o.append(float(v) if on else 0.0)
            else:
                o.append((float(v) if ch in double_channels else v) if on else None)
        out.append(tuple(o))
    return out
