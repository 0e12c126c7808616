"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from trino_b200 import abi
from trino_b200.page import AbiPage

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "liboracle.so")


class Q1Result(C.Structure):
    _fields_ = [("num_groups", C.c_int32), ("returnflag", C.c_int8 * 16), ("linestatus", C.c_int8 * 16),
                ("sum_qty", C.c_double * 16), ("sum_base_price", C.c_double * 16), ("sum_disc_price", C.c_double * 16),
                ("sum_charge", C.c_double * 16), ("avg_qty", C.c_double * 16), ("avg_price", C.c_double * 16),
                ("avg_disc", C.c_double * 16), ("count_order", C.c_int64 * 16)]


_lib = None
VP = C.c_void_p
PP = C.POINTER(abi.Page)
I32P = C.POINTER(C.c_int32)


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"])
    lib = C.CDLL(LIB)
    sig = {
        "orc_hash_long": (C.c_uint64, [C.c_int64]),
        "orc_hash_double": (C.c_uint64, [C.c_double]),
        "orc_hash_real": (C.c_uint64, [C.c_float]),
        "orc_xxh64": (C.c_uint64, [VP, C.c_int64, C.c_uint64]),
        "orc_xxh64_long": (C.c_uint64, [C.c_int64]),
        "orc_murmur3": (C.c_uint64, [C.c_uint64]),
        "orc_combine_hash": (C.c_uint64, [C.c_uint64, C.c_uint64]),
        "orc_array_size": (C.c_int32, [C.c_int64, C.c_double]),
        "orc_join_hash_array_size": (C.c_int32, [C.c_int64]),
        "orc_process_raw_hash": (C.c_int32, [C.c_int64, C.c_int32]),
        "orc_local_partition": (C.c_int32, [C.c_int64, C.c_int32]),
        "orc_row_hashes": (None, [PP, VP, C.c_int32, VP]),
        "orc_groupby_create": (VP, [C.c_int32, C.c_int32]),
        "orc_groupby_destroy": (None, [VP]),
        "orc_groupby_get_group_ids": (C.c_int32, [VP, PP, VP, C.c_int32, VP]),
        "orc_groupby_group_count": (C.c_int32, [VP]),
        "orc_groupby_capacity": (C.c_int32, [VP]),
        "orc_agg_sum_double": (None, [VP, C.c_int64, VP, VP, VP, VP, VP]),
        "orc_agg_avg_double": (None, [VP, C.c_int64, VP, VP, VP, VP, VP]),
        "orc_agg_count": (None, [VP, C.c_int64, VP, VP, VP]),
        "orc_agg_sum_bigint": (C.c_int32, [VP, C.c_int64, VP, VP, VP, VP, VP]),
        "orc_agg_sum_decimal": (None, [VP, C.c_int64, VP, C.c_int32, VP, VP, VP, VP, VP]),
        "orc_agg_sum_decimal_combine": (None, [VP, VP, VP, VP, C.c_int64]),
        "orc_decimal_sum_overflows": (C.c_int32, [C.c_int64, C.c_int64, C.c_int64]),
        "orc_agg_minmax_double": (None, [VP, C.c_int64, VP, VP, C.c_int32, VP, VP]),
        "orc_agg_minmax_bigint": (None, [VP, C.c_int64, VP, VP, C.c_int32, VP, VP]),
        "orc_join_build": (VP, [PP, VP, C.c_int32, C.c_int32]),
        "orc_join_destroy": (None, [VP]),
        "orc_join_hash_size": (C.c_int32, [VP]),
        "orc_join_has_links": (C.c_int32, [VP]),
        "orc_join_copy_links": (None, [VP, VP]),
        "orc_join_positions": (None, [VP, PP, VP, VP]),
        "orc_join_expand": (C.c_int64, [VP, VP, C.c_int64, C.c_int32, C.c_int32, VP, VP, C.c_int64]),
        "orc_semi_join_bigint": (None, [VP, VP, C.c_int64, VP, VP, C.c_int64, VP, VP]),
        "orc_semi_join_float": (None, [C.c_int32, VP, VP, C.c_int64, VP, VP, C.c_int64, VP, VP]),
        "orc_join_probe_timed": (C.c_double, [VP, VP, C.c_int64, C.c_int32, VP, VP, VP]),
        "orc_partition_ids": (None, [PP, VP, C.c_int32, C.c_int32, VP, VP]),
        "orc_partition_positions": (None, [PP, VP, C.c_int32, C.c_int32, VP, C.c_int32, C.c_int32, C.c_int32, VP, VP, VP]),
        "orc_q1_run": (C.c_double, [C.c_int64, VP, VP, VP, VP, VP, VP, VP, C.c_int32, C.c_int32, C.POINTER(Q1Result)]),
        "orc_splitmix64": (C.c_uint64, [C.c_uint64]),
        "orc_synth_orders_keys": (None, [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int32, VP]),
        "orc_synth_lineitem_rows": (C.c_int64, [C.c_int64]),
        "orc_synth_lineitem_keys": (None, [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int32, VP]),
        "orc_synth_lineitem_q1": (None, [C.c_int64, C.c_int64, C.c_uint64, VP, VP, VP, VP, VP, VP, VP]),
        "orc_synth_orders_custkeys": (None, [C.c_int64, C.c_int64, C.c_int64, C.c_uint64, C.c_int32, C.c_int64, C.c_uint64, VP]),
        "orc_synth_store_sales": (C.c_int64, [C.c_int64, C.c_int64, C.c_uint64, VP, VP, VP, VP, VP, VP, VP]),
        "orc_hardware_threads": (C.c_int32, []),
        "orc_pjoin_build": (VP, [VP, VP, C.c_int64, C.c_int32, C.c_int32, VP]),
        "orc_pjoin_destroy": (None, [VP]),
        "orc_pjoin_partitions": (C.c_int32, [VP]),
        "orc_pjoin_threads": (C.c_int32, [VP]),
        "orc_pjoin_touch": (None, [VP, VP, C.c_int64, C.c_int64]),
        "orc_pjoin_probe": (C.c_double, [VP, VP, C.c_int64, VP, VP]),
        "orc_pjoin_decode": (None, [VP, VP, C.c_int64, VP]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _p(arr):
    return C.c_void_p(arr.ctypes.data) if arr is not None else None


def _chan(channels):
    return np.ascontiguousarray(channels, dtype=np.int32)


def xxh64(data: bytes, seed=0):
    buf = np.frombuffer(data, dtype=np.uint8) if data else np.zeros(1, dtype=np.uint8)
    return load().orc_xxh64(_p(buf), len(data), seed)


def row_hashes(page, channels):
    ap = AbiPage(page, nulls_as_bytemap=True)
    ch = _chan(channels)
    out = np.empty(page.position_count, dtype=np.int64)
    load().orc_row_hashes(ap.ref(), _p(ch), len(ch), _p(out))
    return out


class GroupByHash:
    def __init__(self, kind=0, expected=100):
        self.lib = load()
        self.h = self.lib.orc_groupby_create(kind, expected)

    def get_group_ids(self, page, channels):
        ap = AbiPage(page)
        ch = _chan(channels)
        out = np.empty(page.position_count, dtype=np.int32)
        rc = self.lib.orc_groupby_get_group_ids(self.h, ap.ref(), _p(ch), len(ch), _p(out))
        assert rc == 0, rc
        return out

    def group_count(self):
        return self.lib.orc_groupby_group_count(self.h)

    def capacity(self):
        return self.lib.orc_groupby_capacity(self.h)

    def close(self):
        if self.h:
            self.lib.orc_groupby_destroy(self.h)
            self.h = None


class Join:
    def __init__(self, build_page, key_channels, force_default=False):
        self.lib = load()
        self.ap = AbiPage(build_page)      # the oracle keeps views into these buffers
        self.build_page = build_page
        self.key_channels = _chan(key_channels)
        self.h = self.lib.orc_join_build(self.ap.ref(), _p(self.key_channels), len(self.key_channels), int(force_default))  # 0 auto, 1 DefaultPagesHash, 2 BigintPagesHash

    def positions(self, probe_page, key_channels):
        ap = AbiPage(probe_page)
        ch = _chan(key_channels)
        out = np.empty(probe_page.position_count, dtype=np.int32)
        self.lib.orc_join_positions(self.h, ap.ref(), _p(ch), _p(out))
        return out

    def links(self):
        out = np.empty(self.build_page.position_count, dtype=np.int32)
        self.lib.orc_join_copy_links(self.h, _p(out))
        return out

    def has_links(self):
        return bool(self.lib.orc_join_has_links(self.h))

    def expand(self, positions, join_type=0, single_match=False):
        n = len(positions)
        positions = np.ascontiguousarray(positions, dtype=np.int32)
        cnt = self.lib.orc_join_expand(self.h, _p(positions), n, join_type, int(single_match), None, None, 0)
        op = np.empty(max(cnt, 1), dtype=np.int32)
        ob = np.empty(max(cnt, 1), dtype=np.int32)
        self.lib.orc_join_expand(self.h, _p(positions), n, join_type, int(single_match), _p(op), _p(ob), cnt)
        return op[:cnt], ob[:cnt]

    def probe_timed(self, keys, threads, build_payload=None):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        out = np.empty(len(keys), dtype=np.int32)
        out_payload = np.empty(len(keys), dtype=np.int32) if build_payload is not None else None
        secs = self.lib.orc_join_probe_timed(self.h, _p(keys), len(keys), threads, _p(out), _p(build_payload), _p(out_payload))
        return secs, out, out_payload

    def close(self):
        if self.h:
            self.lib.orc_join_destroy(self.h)
            self.h = None


class PartitionedJoin:
    """The stable CPU timing arm (oracle.h: orc_pjoin_*): PartitionedLookupSource of P lookup sources built in parallel, probed by a
    persistent pool of drivers; every buffer of the timed region is allocated and first-touched before it."""

    def __init__(self, build_keys, build_payload, partitions, threads):
        self.lib = load()
        self.keys = np.ascontiguousarray(build_keys, dtype=np.int64)
        self.payload = None if build_payload is None else np.ascontiguousarray(build_payload, dtype=np.int32)
        secs = C.c_double()
        self.h = self.lib.orc_pjoin_build(_p(self.keys), _p(self.payload), len(self.keys), partitions, threads, C.cast(C.byref(secs), VP))
        self.build_seconds = secs.value
        self.partitions = self.lib.orc_pjoin_partitions(self.h)
        self.threads = self.lib.orc_pjoin_threads(self.h)

    def alloc(self, n, dtype):
        """probe-side array of n rows, first-touched by the drivers that will use its pages"""
        a = np.empty(n, dtype=dtype)
        self.lib.orc_pjoin_touch(self.h, _p(a), a.itemsize, n)
        return a

    def probe(self, keys, out_positions, out_payload=None):
        return self.lib.orc_pjoin_probe(self.h, _p(keys), len(keys), _p(out_positions), _p(out_payload))

    def decode(self, positions):
        out = np.empty(len(positions), dtype=np.int32)
        self.lib.orc_pjoin_decode(self.h, _p(positions), len(positions), _p(out))
        return out

    def close(self):
        if self.h:
            self.lib.orc_pjoin_destroy(self.h)
            self.h = None


def partition_ids(page, key_channels, bucket_count, bucket_to_partition=None):
    ap = AbiPage(page)
    ch = _chan(key_channels)
    b2p = None if bucket_to_partition is None else _chan(bucket_to_partition)
    out = np.empty(page.position_count, dtype=np.int32)
    load().orc_partition_ids(ap.ref(), _p(ch), len(ch), bucket_count, _p(b2p), _p(out))
    return out


def local_partition_ids(page, hash_channels, partition_count):
    """LocalPartitionGenerator.getPartitions (M/operator/exchange/LocalPartitionGenerator.java:53-66)"""
    lib = load()
    return np.array([lib.orc_local_partition(int(h), partition_count) for h in row_hashes(page, hash_channels)], dtype=np.int32)


def partition_positions(page, key_channels, bucket_count, bucket_to_partition, partition_count, null_channel, replicates_any_row, any_row_replicated):
    """returns (list of position arrays per partition, new any_row_replicated)"""
    ap = AbiPage(page)
    ch = _chan(key_channels)
    b2p = None if bucket_to_partition is None else _chan(bucket_to_partition)
    n = page.position_count
    offsets = np.zeros(partition_count + 1, dtype=np.int64)
    positions = np.zeros((n + 1) * (partition_count + 1), dtype=np.int32)
    flag = C.c_int32(int(any_row_replicated))
    load().orc_partition_positions(ap.ref(), _p(ch), len(ch), bucket_count, _p(b2p), partition_count, null_channel, int(replicates_any_row),
                                   C.cast(C.byref(flag), C.c_void_p), _p(offsets), _p(positions))
    return [positions[offsets[p]:offsets[p + 1]].copy() for p in range(partition_count)], bool(flag.value)


def synth_orders_keys(n_total, first, count, seed, shuffle):
    out = np.empty(count, dtype=np.int64)
    load().orc_synth_orders_keys(n_total, first, count, seed, int(shuffle), _p(out))
    return out


def synth_lineitem_rows(n_orders):
    return load().orc_synth_lineitem_rows(n_orders)


def synth_lineitem_keys(n_orders, first, count, seed, shuffle):
    out = np.empty(count, dtype=np.int64)
    load().orc_synth_lineitem_keys(n_orders, first, count, seed, int(shuffle), _p(out))
    return out


def synth_lineitem_q1(n, first, seed):
    cols = dict(shipdate=np.empty(n, np.int32), returnflag=np.empty(n, np.int8), linestatus=np.empty(n, np.int8),
                quantity=np.empty(n, np.float64), extendedprice=np.empty(n, np.float64), discount=np.empty(n, np.float64),
                tax=np.empty(n, np.float64))
    load().orc_synth_lineitem_q1(n, first, seed, *[_p(cols[k]) for k in
                                                   ("shipdate", "returnflag", "linestatus", "quantity", "extendedprice", "discount", "tax")])
    return cols


def synth_orders_custkeys(n_total, first, count, seed, shuffle, n_customers, cust_seed):
    out = np.empty(count, dtype=np.int64)
    load().orc_synth_orders_custkeys(n_total, first, count, seed, int(shuffle), n_customers, cust_seed, _p(out))
    return out


def synth_store_sales(n, first, seed):
    """dict of columns (+ Arrow validity bitmaps of the two nullable keys) and the count of rows with both keys present"""
    cols = dict(date_sk=np.empty(n, np.int64), item_sk=np.empty(n, np.int64), customer_sk=np.empty(n, np.int64),
                customer_valid=np.empty((n + 7) // 8, np.uint8), store_sk=np.empty(n, np.int64), store_valid=np.empty((n + 7) // 8, np.uint8),
                net_paid=np.empty(n, np.float64))
    both = load().orc_synth_store_sales(n, first, seed, *[_p(cols[k]) for k in
                                                          ("date_sk", "item_sk", "customer_sk", "customer_valid", "store_sk", "store_valid", "net_paid")])
    return cols, both


def q1_run(cols, cutoff, threads):
    res = Q1Result()
    n = len(cols["shipdate"])
    secs = load().orc_q1_run(n, *[_p(cols[k]) for k in ("shipdate", "returnflag", "linestatus", "quantity", "extendedprice", "discount", "tax")],
                             cutoff, threads, C.byref(res))
    rows = []
    for g in range(res.num_groups):
        rows.append((chr(res.returnflag[g]), chr(res.linestatus[g]), res.sum_qty[g], res.sum_base_price[g], res.sum_disc_price[g],
                     res.sum_charge[g], res.avg_qty[g], res.avg_price[g], res.avg_disc[g], res.count_order[g]))
    return secs, rows


def hardware_threads():
    return load().orc_hardware_threads()


def semi_join_bigint(set_block, probe_block):
    """rows of the BOOLEAN column HashSemiJoinOperator appends: True / False / None"""
    lib = load()
    sv = np.ascontiguousarray(set_block.values, dtype=np.int64)
    pv = np.ascontiguousarray(probe_block.values, dtype=np.int64)
    svalid = None if set_block.nulls is None else np.packbits(~set_block.nulls, bitorder="little")
    pvalid = None if probe_block.nulls is None else np.packbits(~probe_block.nulls, bitorder="little")
    val = np.zeros(max(len(pv), 1), dtype=np.int8)
    isnull = np.zeros(max(len(pv), 1), dtype=np.uint8)
    lib.orc_semi_join_bigint(_p(sv), _p(svalid), len(sv), _p(pv), _p(pvalid), len(pv), _p(val), _p(isnull))
    return [None if isnull[i] else bool(val[i]) for i in range(len(pv))]


def semi_join_float(set_block, probe_block):
    """HashSemiJoinOperator's BOOLEAN column over a DOUBLE or REAL channel (IDENTICAL membership: a NaN in the set answers NaN probes)"""
    from trino_b200 import abi
    lib = load()
    kind = 1 if set_block.type == abi.FLOAT64 else 2
    assert set_block.type == probe_block.type and set_block.type in (abi.FLOAT64, abi.FLOAT32)
    sv = np.ascontiguousarray(set_block.values)
    pv = np.ascontiguousarray(probe_block.values)
    svalid = None if set_block.nulls is None else np.packbits(~set_block.nulls, bitorder="little")
    pvalid = None if probe_block.nulls is None else np.packbits(~probe_block.nulls, bitorder="little")
    val = np.zeros(max(len(pv), 1), dtype=np.int8)
    isnull = np.zeros(max(len(pv), 1), dtype=np.uint8)
    lib.orc_semi_join_float(kind, _p(sv), _p(svalid), len(sv), _p(pv), _p(pvalid), len(pv), _p(val), _p(isnull))
    return [None if isnull[i] else bool(val[i]) for i in range(len(pv))]


# ---- long DECIMAL sums (DecimalSumAggregation over LongDecimalWithOverflowState) --------------------------------------------------
M64 = (1 << 64) - 1


def int128_words(x):
    """python int -> (high, low) as signed 64-bit words of its 128-bit two's complement"""
    u = x & ((1 << 128) - 1)
    hi, lo = u >> 64, u & M64
    return (hi - (1 << 64) if hi >> 63 else hi), (lo - (1 << 64) if lo >> 63 else lo)


def int128_value(high, low):
    return (int(high) << 64) | (int(low) & M64)


class DecimalSumState:
    """one LongDecimalWithOverflowState driven through the oracle's restatement of inputLongDecimal / inputShortDecimal / combine"""

    def __init__(self):
        self.decimal = np.zeros(2, dtype=np.int64)
        self.overflow = np.zeros(1, dtype=np.int64)
        self.nonnull = np.zeros(1, dtype=np.uint8)

    def add(self, values, short=False):
        n = len(values)
        gids = np.zeros(n, dtype=np.int32)
        if short:
            v = np.ascontiguousarray(values, dtype=np.int64)
        else:
            v = np.ascontiguousarray([w for x in values for w in int128_words(x)], dtype=np.int64)
        load().orc_agg_sum_decimal(_p(gids), n, _p(v), int(short), None, None, _p(self.decimal), _p(self.overflow), _p(self.nonnull))
        return self

    def combine(self, other):
        load().orc_agg_sum_decimal_combine(_p(self.decimal), _p(self.overflow), _p(self.nonnull), _p(other.decimal), int(other.overflow[0]))
        return self

    @property
    def value(self):
        return int128_value(self.decimal[0], self.decimal[1])

    def average(self, rows):
        """DecimalAverageAggregation.average :152-175 over this state and its row counter: (decimal + overflow * 2^128) / rows, HALF_UP
        (Int128Math.divideRoundUp when overflow == 0 - then the result must stay inside +-(10^38 - 1) - else BigDecimal.divide(…, HALF_UP),
        whose result must fit 128 bits)"""
        if rows == 0:
            return None
        total = self.value + int(self.overflow[0]) * (1 << 128)
        q, r = divmod(abs(total), rows)
        if 2 * r >= rows:
            q += 1
        result = -q if total < 0 else q
        if int(self.overflow[0]) == 0:
            if abs(result) >= 10**38:
                raise OverflowError("Decimal overflow")
        elif not -(1 << 127) <= result < (1 << 127):
            raise OverflowError("Decimal overflow")
        return result

    def output(self):
        """outputDecimal: the value, None for an empty state, or raises OverflowError("Decimal overflow")"""
        if not self.nonnull[0]:
            return None
        if load().orc_decimal_sum_overflows(int(self.decimal[0]), int(self.decimal[1]), int(self.overflow[0])) != 0:
            raise OverflowError("Decimal overflow")
        return self.value
