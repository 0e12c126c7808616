"""Host-side mirror of Trino's Page / Block data model (S/Page.java:31, S/block/*), numpy backed.

Blocks keep the reference's shapes — value blocks (LongArrayBlock, IntArrayBlock, ShortArrayBlock,
ByteArrayBlock, VariableWidthBlock), DictionaryBlock and RunLengthEncodedBlock — and convert to the
Arrow-layout `tgpu_column` structs of the C ABI without copying the value buffers.
Nulls are held the way Java holds them (boolean valueIsNull[], True = NULL) and handed to the library
either as that byte map (TGPU_COL_NULLS_BYTEMAP) or as an Arrow validity bitmap.
"""
import ctypes as C

import numpy as np

from . import abi

_NP_OF_TYPE = {abi.INT64: np.int64, abi.INT32: np.int32, abi.INT16: np.int16, abi.INT8: np.int8, abi.FLOAT64: np.float64, abi.FLOAT32: np.float32}


class Block:
    """A value block.  `type` is a tgpu_type; `values` a numpy array (UTF8: uint8 bytes + int32 offsets)."""

    def __init__(self, type_, values, nulls=None, offsets=None):
        self.type = type_
        self.values = values
        self.offsets = offsets
        self.nulls = None if nulls is None else np.ascontiguousarray(nulls, dtype=np.bool_)
        self.position_count = (len(offsets) - 1) if type_ == abi.UTF8 else len(values)
        if self.nulls is not None and not self.nulls.any():
            self.nulls = None

    # --- constructors named after the SQL types of the reference
    @staticmethod
    def _fixed(type_, values, nulls):
        v = list(values) if not isinstance(values, np.ndarray) else values
        if not isinstance(v, np.ndarray):
            if nulls is None and any(x is None for x in v):
                nulls = [x is None for x in v]
            v = [0 if x is None else x for x in v]
        arr = np.ascontiguousarray(np.asarray(v, dtype=_NP_OF_TYPE[type_]))
        return Block(type_, arr, nulls)

    @staticmethod
    def bigint(values, nulls=None):
        return Block._fixed(abi.INT64, values, nulls)

    @staticmethod
    def int128(values, nulls=None):
        """Int128ArrayBlock (long DECIMAL): python ints (None = NULL) -> int64[n][2] = (high, low), the layout of the block's long[]
        (S/block/Int128ArrayBlock.java:123-133)"""
        v = list(values)
        if nulls is None and any(x is None for x in v):
            nulls = [x is None for x in v]
        arr = np.zeros((len(v), 2), dtype=np.int64)
        for i, x in enumerate(v):
            x = 0 if x is None else int(x)
            assert -(1 << 127) <= x < (1 << 127), "value does not fit 128 bits"
            u = x & ((1 << 128) - 1)
            hi, lo = u >> 64, u & ((1 << 64) - 1)
            arr[i, 0] = hi - (1 << 64) if hi >= (1 << 63) else hi
            arr[i, 1] = lo - (1 << 64) if lo >= (1 << 63) else lo
        return Block(abi.INT128, arr, nulls)

    @staticmethod
    def integer(values, nulls=None):
        return Block._fixed(abi.INT32, values, nulls)

    date = integer

    @staticmethod
    def smallint(values, nulls=None):
        return Block._fixed(abi.INT16, values, nulls)

    @staticmethod
    def tinyint(values, nulls=None):
        return Block._fixed(abi.INT8, values, nulls)

    @staticmethod
    def boolean(values, nulls=None):
        v = [None if x is None else int(bool(x)) for x in values] if not isinstance(values, np.ndarray) else values.astype(np.int8)
        return Block._fixed(abi.INT8, v, nulls)

    @staticmethod
    def double(values, nulls=None):
        return Block._fixed(abi.FLOAT64, values, nulls)

    @staticmethod
    def real(values, nulls=None):
        """REAL: float32 values; the buffer is the IntArrayBlock of raw float bits the reference keeps (S/type/RealType.java:104-121)"""
        return Block._fixed(abi.FLOAT32, values, nulls)

    @staticmethod
    def varchar(values):
        nulls = [v is None for v in values]
        enc = [b"" if v is None else (v.encode() if isinstance(v, str) else bytes(v)) for v in values]
        offsets = np.zeros(len(enc) + 1, dtype=np.int32)
        np.cumsum([len(e) for e in enc], out=offsets[1:])
        data = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, dtype=np.uint8)
        if len(data) == 0:
            data = np.zeros(1, dtype=np.uint8)
        return Block(abi.UTF8, data, nulls if any(nulls) else None, offsets)

    # --- accessors
    def is_null(self, i):
        return self.nulls is not None and bool(self.nulls[i])

    def get(self, i):
        if self.is_null(i):
            return None
        if self.type == abi.UTF8:
            return bytes(self.values[self.offsets[i]:self.offsets[i + 1]])
        v = self.values[i]
        if self.type == abi.INT128:
            return (int(v[0]) << 64) | (int(v[1]) & ((1 << 64) - 1))
        return float(v) if self.type in (abi.FLOAT64, abi.FLOAT32) else int(v)

    def to_pylist(self):
        return [self.get(i) for i in range(self.position_count)]

    def flatten(self):
        return self

    def get_positions(self, idx):
        idx = np.asarray(idx, dtype=np.int64)
        nulls = None if self.nulls is None else self.nulls[idx]
        if self.type == abi.UTF8:
            return Block.varchar([self.get(int(i)) for i in idx])
        return Block(self.type, np.ascontiguousarray(self.values[idx]), nulls)


class DictionaryBlock:
    """S/block/DictionaryBlock.java:37-40"""

    def __init__(self, dictionary, ids):
        self.dictionary = dictionary
        self.ids = np.ascontiguousarray(ids, dtype=np.int32)
        self.type = abi.DICT32
        self.position_count = len(self.ids)

    def get(self, i):
        return self.dictionary.get(int(self.ids[i]))

    def is_null(self, i):
        return self.dictionary.is_null(int(self.ids[i]))

    def to_pylist(self):
        return [self.get(i) for i in range(self.position_count)]

    def flatten(self):
        return self.dictionary.get_positions(self.ids)


class RunLengthEncodedBlock:
    """S/block/RunLengthEncodedBlock.java:71-72"""

    def __init__(self, value, position_count):
        assert value.position_count == 1
        self.value = value
        self.type = abi.RLE
        self.position_count = position_count

    def get(self, i):
        return self.value.get(0)

    def is_null(self, i):
        return self.value.is_null(0)

    def to_pylist(self):
        return [self.value.get(0)] * self.position_count

    def flatten(self):
        return self.value.get_positions(np.zeros(self.position_count, dtype=np.int64))


class RowBlock:
    """S/block/RowBlock.java:37-45,91-112: equally long field blocks plus optional row-level NULL flags.  This is the shape of a
    multi-field aggregation state on the reference's wire (AccumulatorCompiler.java:687-760 writes one ROW entry per group through
    RowBlockBuilder.buildEntry); the C ABI carries the fields as consecutive flat columns, so ROW blocks exist only on the host side
    of the boundary: flatten_row_blocks() before add_input, compose_row_blocks() after get_output."""
    type = -1      # no tgpu_type: never crosses the C ABI

    def __init__(self, fields, row_nulls=None):
        assert fields, "a row block needs at least one field"
        self.fields = list(fields)
        self.position_count = self.fields[0].position_count
        for f in self.fields:
            assert f.position_count == self.position_count, "field position counts differ"
        self.row_nulls = None if row_nulls is None else np.asarray(row_nulls, dtype=bool)

    def is_null(self, i):
        return self.row_nulls is not None and bool(self.row_nulls[i])

    def get(self, i):
        return None if self.is_null(i) else tuple(f.get(i) for f in self.fields)

    def to_pylist(self):
        return [self.get(i) for i in range(self.position_count)]

    def flatten(self):
        return self

    def null_suppressed_fields(self):
        """RowBlock.getFieldBlocks of a block with NULL rows: a NULL row reads as NULL in every field"""
        if self.row_nulls is None or not self.row_nulls.any():
            return [f.flatten() for f in self.fields]
        out = []
        for f in self.fields:
            f = f.flatten()
            nulls = self.row_nulls if f.nulls is None else (f.nulls | self.row_nulls)
            out.append(Block(f.type, f.values, nulls, f.offsets) if f.type == abi.UTF8 else Block(f.type, f.values, nulls))
        return out


def flatten_row_blocks(page):
    """-> (page without ROW blocks, first flat channel of every original channel)"""
    blocks, first = [], []
    for b in page.blocks:
        first.append(len(blocks))
        if isinstance(b, RowBlock):
            blocks.extend(b.null_suppressed_fields())
        else:
            blocks.append(b)
    return Page(*blocks, position_count=page.position_count), first


def compose_row_blocks(page, widths):
    """widths[c] = number of consecutive flat channels that make up output channel c (1 = a plain block)"""
    assert sum(widths) == page.channel_count, "row layout does not cover the page"
    blocks, at = [], 0
    for w in widths:
        blocks.append(page.blocks[at] if w == 1 else RowBlock(page.blocks[at:at + w]))
        at += w
    return Page(*blocks, position_count=page.position_count)


# ---- decimal aggregation states as the reference serialises them (VARBINARY) ----------------------------------------------------------
# The C ABI carries a decimal sum state as two flat columns (INT128 sum, BIGINT overflow) and a decimal average state as three (+ BIGINT
# row count); between a Java PARTIAL and a Java FINAL step they travel as ONE VARBINARY channel of 8 to 32 bytes per group:
#   sum: low, [high, [overflow]]          - LongDecimalWithOverflowStateSerializer.serialize :36-60 / deserialize :62-96
#   avg: low, [high,] [count, overflow]   - LongDecimalWithOverflowAndLongStateSerializer.serialize :36-70 / deserialize :72-113
# (M/operator/aggregation/state/).  These four functions are the marshaller's side of that (PageMarshaller.java has the same pair).
def _words_of(block, i):
    v = block.values[i]
    return int(v[0]), int(v[1])          # high, low


def _le64(*words):
    return b"".join(int(w & ((1 << 64) - 1)).to_bytes(8, "little") for w in words)


def _i64(b):
    return int.from_bytes(b, "little", signed=True)


def encode_decimal_sum_states(sum_block, overflow_block):
    out = []
    for i in range(sum_block.position_count):
        if sum_block.is_null(i):
            out.append(None)                                              # !state.isNotNull() -> appendNull
            continue
        high, low = _words_of(sum_block, i)
        overflow = int(overflow_block.values[i])
        words = [low] + ([high] if high != 0 else [])
        if overflow != 0:
            words = [low, high, overflow]
        out.append(_le64(*words))
    return Block.varchar(out)


def decode_decimal_sum_states(block):
    sums, overflows = [], []
    for i in range(block.position_count):
        b = block.get(i)
        if b is None:
            sums.append(None)
            overflows.append(0)
            continue
        low = _i64(b[0:8])
        high = _i64(b[8:16]) if len(b) >= 16 else 0
        overflow = _i64(b[16:24]) if len(b) == 24 else 0
        sums.append((high << 64) | (low & ((1 << 64) - 1)))
        overflows.append(overflow)
    return Block.int128(sums), Block.bigint(overflows)


def encode_decimal_avg_states(sum_block, overflow_block, count_block):
    out = []
    for i in range(sum_block.position_count):
        count = int(count_block.values[i])
        if count == 0:
            out.append(None)
            continue
        high, low = (0, 0) if sum_block.is_null(i) else _words_of(sum_block, i)
        overflow = int(overflow_block.values[i])
        words = [low] + ([high] if high != 0 else [])
        if not (overflow == 0 and count == 1):
            words += [count, overflow]
        out.append(_le64(*words))
    return Block.varchar(out)


def decode_decimal_avg_states(block):
    sums, overflows, counts = [], [], []
    for i in range(block.position_count):
        b = block.get(i)
        if b is None:
            sums.append(None)
            overflows.append(0)
            counts.append(0)
            continue
        low, high, overflow, count = _i64(b[0:8]), 0, 0, 1
        if len(b) == 32:
            high, count, overflow = _i64(b[8:16]), _i64(b[16:24]), _i64(b[24:32])
        elif len(b) == 16:
            high = _i64(b[8:16])
        elif len(b) == 24:
            count, overflow = _i64(b[8:16]), _i64(b[16:24])
        sums.append((high << 64) | (low & ((1 << 64) - 1)))
        overflows.append(overflow)
        counts.append(count)
    return Block.int128(sums), Block.bigint(overflows), Block.bigint(counts)


def compose_state_blocks(page, layout):
    """layout[c] of output channel c: 1 (a plain block), an int k > 1 (ROW of k flat channels), "decimal_sum" (2 flat channels ->
    VARBINARY) or "decimal_avg" (3 flat channels -> VARBINARY)"""
    blocks, at = [], 0
    for item in layout:
        if item == "decimal_sum":
            blocks.append(encode_decimal_sum_states(page.blocks[at], page.blocks[at + 1]))
            at += 2
        elif item == "decimal_avg":
            blocks.append(encode_decimal_avg_states(page.blocks[at], page.blocks[at + 1], page.blocks[at + 2]))
            at += 3
        elif item == 1:
            blocks.append(page.blocks[at])
            at += 1
        else:
            blocks.append(RowBlock(page.blocks[at:at + item]))
            at += item
    assert at == page.channel_count, "state layout does not cover the page"
    return Page(*blocks, position_count=page.position_count)


def flatten_state_blocks(page, decimal_channels):
    """decimal_channels: {input channel: "decimal_sum" | "decimal_avg"}; ROW blocks flatten by themselves.  -> flat page"""
    blocks = []
    for c, b in enumerate(page.blocks):
        kind = decimal_channels.get(c)
        if kind == "decimal_sum":
            blocks.extend(decode_decimal_sum_states(b.flatten()))
        elif kind == "decimal_avg":
            blocks.extend(decode_decimal_avg_states(b.flatten()))
        elif isinstance(b, RowBlock):
            blocks.extend(b.null_suppressed_fields())
        else:
            blocks.append(b)
    return Page(*blocks, position_count=page.position_count)


class Page:
    """S/Page.java:31"""

    def __init__(self, *blocks, position_count=None):
        self.blocks = list(blocks)
        if position_count is None:
            position_count = self.blocks[0].position_count if self.blocks else 0
        self.position_count = position_count
        for b in self.blocks:
            assert b.position_count == position_count, "block position counts differ"

    @property
    def channel_count(self):
        return len(self.blocks)

    def get_block(self, channel):
        return self.blocks[channel]

    def get_columns(self, channels):
        return Page(*[self.blocks[c] for c in channels], position_count=self.position_count)

    def rows(self):
        cols = [b.to_pylist() for b in self.blocks]
        return [tuple(c[i] for c in cols) for i in range(self.position_count)]


def _ptr(arr):
    return arr.ctypes.data_as(C.c_void_p).value if arr is not None and arr.size else (arr.ctypes.data if arr is not None else None)


def _pack_validity(nulls):
    return np.packbits(~nulls, bitorder="little")


class AbiPage:
    """A tgpu_page view over a host Page.  Holds every buffer alive for the duration of the call."""

    def __init__(self, page, nulls_as_bytemap=True):
        self.keep = []
        self.ncols = len(page.blocks)
        self.columns = (abi.Column * max(1, self.ncols))()
        for c, b in enumerate(page.blocks):
            self._fill(self.columns[c], b, nulls_as_bytemap)
        self.page = abi.Page(self.ncols, 0, page.position_count, C.cast(self.columns, C.POINTER(abi.Column)))

    def _fill(self, col, b, bytemap):
        col.length = b.position_count
        col.flags = 0
        if isinstance(b, DictionaryBlock):
            col.type = abi.DICT32
            col.data = b.ids.ctypes.data
            sub = abi.Column()
            self._fill(sub, b.dictionary, bytemap)
            self.keep.append(sub)
            col.dictionary = C.pointer(sub)
            return
        if isinstance(b, RunLengthEncodedBlock):
            col.type = abi.RLE
            sub = abi.Column()
            self._fill(sub, b.value, bytemap)
            self.keep.append(sub)
            col.dictionary = C.pointer(sub)
            return
        col.type = b.type
        self.keep.append(b.values)
        col.data = b.values.ctypes.data
        if b.type == abi.UTF8:
            self.keep.append(b.offsets)
            col.offsets = b.offsets.ctypes.data
        if b.nulls is not None:
            if bytemap:
                self.keep.append(b.nulls)
                col.validity = b.nulls.ctypes.data
                col.flags = abi.COL_NULLS_BYTEMAP
            else:
                bits = _pack_validity(b.nulls)
                self.keep.append(bits)
                col.validity = bits.ctypes.data

    def ref(self):
        return C.byref(self.page)
