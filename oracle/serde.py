"""CPU restatement of the reference's page wire format (test infrastructure only; never imported by trino_b200/).

SURVEY.md §8(f) rank 1: the format GPU stages must speak to exchange pages with Java tasks.  Uncompressed, unencrypted form
(`CompressionCodec.NONE`, no cipher):

  serialized page  = int32 positionCount | int32 uncompressedSize | int32 compressedSize | raw page
                     (M/execution/buffer/PagesSerdeUtil.java:44-48 header offsets, CompressingEncryptingPageSerializer.java:173-181,351-361:
                      both sizes are the byte length of the raw page when nothing is compressed)
  raw page         = int32 channelCount | block*                      (PagesSerdeUtil.writeRawPage :58-64)
  block            = int32 nameLength | name bytes | body              (M/metadata/InternalBlockEncodingSerde.java:73-94,143-148)
  LONG_ARRAY body  = int32 positionCount | nulls | values              (S/block/LongArrayBlockEncoding.java:61-90)
        nulls      = byte hasNulls [| ceil(n/8) bytes, position i -> bit (7 - i % 8) of byte i / 8]   (S/block/EncoderUtil.java:35-70)
        values     = n x int64 LE when hasNulls == 0, else int32 nonNullCount | nonNullCount x int64   (:120-133)
  INT_ARRAY / SHORT_ARRAY / BYTE_ARRAY: the same with 4 / 2 / 1-byte values (IntArrayBlockEncoding.java, ShortArray..., ByteArray...)
  VARIABLE_WIDTH   = int32 positionCount | nulls | int32 nonNullCount | nonNullCount x int32 ending offsets (from 0) | bytes
                                                                       (S/block/VariableWidthBlockEncoding.java:57-79,112-146)

All integers little-endian (Slice).  BIGINT, DOUBLE and the other 8-byte types travel as LONG_ARRAY; INTEGER/DATE/REAL as INT_ARRAY; long DECIMAL as INT128_ARRAY;
SMALLINT as SHORT_ARRAY; TINYINT/BOOLEAN as BYTE_ARRAY; VARCHAR/VARBINARY as VARIABLE_WIDTH.  The wire carries no SQL type: the
reader is told the channel types (the planner knows them).

Pinned on the reference's golden sizes: TestPagesSerde.testBigintSerializedSize (:183-204: 35 bytes empty, +8 per value) and
testVarcharSerializedSize (:207-228: 43 bytes empty, +9 for "alice", +7 for "bob").  Byte-level content beyond those sizes has no
golden vector in the reference's tests (they are round trips): the null-bit order is pinned on the formula of EncoderUtil.java:46-67.
"""
import struct

import numpy as np

HEADER_SIZE = 12

# 16: Int128ArrayBlockEncoding (S/block/Int128ArrayBlockEncoding.java:52-84): the same body as LONG_ARRAY with two longs per position,
# the high word first - values are int64[n][2] here
_NAMES = {16: b"INT128_ARRAY", 8: b"LONG_ARRAY", 4: b"INT_ARRAY", 2: b"SHORT_ARRAY", 1: b"BYTE_ARRAY", 0: b"VARIABLE_WIDTH"}


def pack_null_bits(nulls):
    """EncoderUtil.encodeNullsAsBitsScalar :46-67: position i sets bit (7 - i % 8) of byte i / 8"""
    return np.packbits(np.asarray(nulls, dtype=np.uint8), bitorder="big").tobytes()


def unpack_null_bits(data, n):
    return np.unpackbits(np.frombuffer(data, dtype=np.uint8), bitorder="big")[:n].astype(bool)


def _block_body_fixed(values, nulls, width):
    n = len(values)
    out = [struct.pack("<i", n)]
    if nulls is None or not np.any(nulls):
        # a block without a valueIsNull array writes hasNulls = 0; a block WITH an all-false array would write 1 and the bits:
        # the restatement (and the GPU, whose columns carry no validity when nothing is NULL) always emits the first form
        out.append(b"\x00")
        out.append(np.ascontiguousarray(values).tobytes())
    else:
        nulls = np.asarray(nulls, dtype=bool)
        out.append(b"\x01")
        out.append(pack_null_bits(nulls))
        kept = np.ascontiguousarray(np.asarray(values)[~nulls])
        out.append(struct.pack("<i", len(kept)))
        out.append(kept.tobytes())
    return b"".join(out)


def _block_body_varwidth(offsets, data, nulls):
    n = len(offsets) - 1
    out = [struct.pack("<i", n)]
    has_nulls = nulls is not None and np.any(nulls)
    out.append(b"\x01" + pack_null_bits(nulls) if has_nulls else b"\x00")
    start = int(offsets[0]) if n >= 0 else 0
    ends = np.asarray(offsets[1:], dtype=np.int64) - start
    if has_nulls:
        ends = ends[~np.asarray(nulls, dtype=bool)]
    out.append(struct.pack("<i", len(ends)))
    out.append(ends.astype("<i4").tobytes())
    out.append(bytes(data[start:int(offsets[n])]) if n > 0 else b"")
    return b"".join(out)


def serialize_columns(position_count, columns):
    """columns: list of ("fixed", numpy values, nulls or None) / ("var", int32 offsets[n+1], bytes-like, nulls or None)"""
    raw = [struct.pack("<i", len(columns))]
    for col in columns:
        if col[0] == "fixed":
            _, values, nulls = col
            width = np.asarray(values).dtype.itemsize * (2 if np.asarray(values).ndim == 2 else 1)
            name = _NAMES[width]
            body = _block_body_fixed(np.asarray(values), nulls, width)
        else:
            _, offsets, data, nulls = col
            name = _NAMES[0]
            body = _block_body_varwidth(np.asarray(offsets), data, nulls)
        raw.append(struct.pack("<i", len(name)) + name + body)
    raw = b"".join(raw)
    return struct.pack("<iii", position_count, len(raw), len(raw)) + raw


def serialize_page(page):
    """trino_b200.page.Page (flat, dictionary or RLE blocks) -> wire bytes"""
    cols = []
    for c in range(page.channel_count):
        b = page.get_block(c).flatten()
        if b.offsets is not None:
            cols.append(("var", b.offsets, b.values, b.nulls))
        else:
            cols.append(("fixed", b.values, b.nulls))
    return serialize_columns(page.position_count, cols)


def deserialize_columns(data):
    """wire bytes -> (position_count, [(encoding name, values / (offsets, bytes), nulls or None)])"""
    position_count, uncompressed, compressed = struct.unpack_from("<iii", data, 0)
    if compressed != len(data) - HEADER_SIZE or uncompressed != compressed:
        raise ValueError("not an uncompressed single-block page")
    pos = HEADER_SIZE
    (channels,) = struct.unpack_from("<i", data, pos)
    pos += 4
    out = []
    widths = {v: k for k, v in _NAMES.items()}
    for _ in range(channels):
        (ln,) = struct.unpack_from("<i", data, pos)
        pos += 4
        name = bytes(data[pos:pos + ln])
        pos += ln
        (n,) = struct.unpack_from("<i", data, pos)
        pos += 4
        has_nulls = data[pos] != 0
        pos += 1
        nulls = None
        if has_nulls:
            nb = (n + 7) // 8
            nulls = unpack_null_bits(data[pos:pos + nb], n)
            pos += nb
        width = widths[name]
        if width:
            per = 2 if width == 16 else 1
            dt = np.dtype("<i%d" % (width // per))
            shape = (lambda rows: (rows, 2)) if per == 2 else (lambda rows: (rows,))
            if nulls is None:
                values = np.frombuffer(data, dtype=dt, count=n * per, offset=pos).reshape(shape(n)).copy()
                pos += n * width
            else:
                (k,) = struct.unpack_from("<i", data, pos)
                pos += 4
                kept = np.frombuffer(data, dtype=dt, count=k * per, offset=pos).reshape(shape(k))
                pos += k * width
                values = np.zeros(shape(n), dtype=dt)        # LongArrayBlockEncoding.expandLongsWithNulls: NULL positions read as 0
                values[~nulls] = kept
            out.append((name.decode(), values, nulls))
        else:
            (k,) = struct.unpack_from("<i", data, pos)
            pos += 4
            ends = np.frombuffer(data, dtype="<i4", count=k, offset=pos).astype(np.int64)
            pos += 4 * k
            offsets = np.zeros(n + 1, dtype=np.int32)
            if nulls is None:
                offsets[1:] = ends
            else:
                # VariableWidthBlockEncoding.readOffsetsWithNullsCompacted: a NULL position repeats the previous ending offset
                lens = np.zeros(n, dtype=np.int64)
                lens[~nulls] = np.diff(np.concatenate([[0], ends]))
                offsets[1:] = np.cumsum(lens)
            total = int(offsets[n]) if n else 0
            payload = bytes(data[pos:pos + total])
            pos += total
            out.append((name.decode(), (offsets, payload), nulls))
    if pos != len(data):
        raise ValueError("trailing bytes")
    return position_count, out
