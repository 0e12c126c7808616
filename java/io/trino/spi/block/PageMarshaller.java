/*
 * Page <-> tgpu_page marshalling for the GPU operators.  Lives in io.trino.spi.block because the raw array getters of
 * LongArrayBlock / ShortArrayBlock / VariableWidthBlock are package-private (S/block/LongArrayBlock.java:235-248,
 * S/block/ShortArrayBlock.java:234-244, S/block/VariableWidthBlock.java:288-302); IntArrayBlock and ByteArrayBlock expose theirs publicly
 * (S/block/IntArrayBlock.java:240-245, S/block/ByteArrayBlock.java:79-87).
 *
 * NOT compiled in this repository (no JDK in the build image).  tests/harness/driver_loop.cpp drives libtrino_gpu.so exactly the way
 * this class does (8192-row pages with boolean[] null maps, copied into one pinned staging region per batch, one tgpu_op_add_input per
 * batch) and is what the end-to-end numbers of bench.py / DESIGN.md stand on.
 *
 * Memory layout written here == include/trino_gpu.h:
 *   tgpu_column { int32 type; int32 flags; int64 length; void* data; int32* offsets; uint8* validity; tgpu_column* dictionary }  (48 bytes)
 *   tgpu_page   { int32 num_columns; int32 flags; int64 num_rows; tgpu_column* columns }                                         (24 bytes)
 * Java null maps are boolean[] (one byte per position, 1 = NULL): they are handed over as they are with TGPU_COL_NULLS_BYTEMAP and
 * packed into Arrow bitmaps on the device.
 */
package io.trino.spi.block;

import io.airlift.slice.Slice;
import io.airlift.slice.Slices;
import io.trino.spi.Page;

import java.lang.foreign.Arena;
import java.lang.foreign.MemoryLayout;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.StructLayout;
import java.lang.foreign.ValueLayout;
import java.util.ArrayList;
import java.util.List;
import java.util.Optional;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;
import static java.lang.foreign.ValueLayout.JAVA_SHORT;

public final class PageMarshaller
{
    // tgpu_type (include/trino_gpu.h)
    public static final int INT64 = 1;
    public static final int INT32 = 2;
    public static final int INT16 = 3;
    public static final int INT8 = 4;
    public static final int FLOAT64 = 5;
    public static final int UTF8 = 7;
    public static final int DICT32 = 8;
    public static final int RLE = 9;
    public static final int INT128 = 10;      // long DECIMAL: Int128ArrayBlock, the high word first (Int128ArrayBlock.java:123-133)
    public static final int FLOAT32 = 11;     // REAL: IntArrayBlock of raw float bits (RealType.java:104-121)
    public static final int COL_NULLS_BYTEMAP = 1;
    public static final int PAGE_DEVICE = 1;

    public static final StructLayout COLUMN = MemoryLayout.structLayout(
            JAVA_INT.withName("type"), JAVA_INT.withName("flags"), JAVA_LONG.withName("length"),
            ADDRESS.withName("data"), ADDRESS.withName("offsets"), ADDRESS.withName("validity"), ADDRESS.withName("dictionary"));
    public static final StructLayout PAGE = MemoryLayout.structLayout(
            JAVA_INT.withName("num_columns"), JAVA_INT.withName("flags"), JAVA_LONG.withName("num_rows"), ADDRESS.withName("columns"));

    /** rows one native call should carry at least: a Java page is <= 8192 rows (PageProcessor.java:58), a GPU launch wants ~1M */
    public static final int BATCH_ROWS = 1 << 20;

    private final MemorySegment staging;      // pinned host memory from tgpu_host_alloc_pinned, owned by GpuContexts
    private long stagingUsed;
    private final List<Page> batch = new ArrayList<>();
    private int batchRows;
    /** tgpu_type per channel; DOUBLE travels as FLOAT64 although both are LongArrayBlocks (S/type/DoubleType.java:205) */
    private final int[] channelTypes;

    public PageMarshaller(MemorySegment pinnedStaging, int[] channelTypes)
    {
        this.staging = pinnedStaging;
        this.channelTypes = channelTypes.clone();
    }

    // ------------------------------------------------------------------------------------------------ ROW-typed aggregation states
    /**
     * Multi-field aggregation states travel between the reference's PARTIAL and FINAL steps as ONE RowBlock channel
     * (AccumulatorCompiler.java:687-760: RowBlockBuilder.buildEntry over the state serializers); the C ABI carries the fields as
     * consecutive flat columns (include/trino_gpu.h, "Intermediate state layout").  flattenRows() runs on every input page of a
     * FINAL / INTERMEDIATE GPU step, composeRows() on every output page of a PARTIAL / INTERMEDIATE one; `channelTypes` of this
     * marshaller and the channel numbers handed to NativeSpecs are those of the FLATTENED page (firstFlatChannel maps the plan's).
     * trino_b200/page.py (RowBlock, flatten_row_blocks, compose_row_blocks) is the tested mirror of the pair.
     */
    public static Page flattenRows(Page page)
    {
        boolean any = false;
        for (int channel = 0; channel < page.getChannelCount(); channel++) {
            any |= flat(page.getBlock(channel)) instanceof RowBlock;
        }
        if (!any) {
            return page;
        }
        List<Block> blocks = new ArrayList<>();
        for (int channel = 0; channel < page.getChannelCount(); channel++) {
            Block block = flat(page.getBlock(channel));
            if (block instanceof RowBlock row) {
                // a NULL row reads as NULL in every field (RowBlock.getNullSuppressedRowFieldsFromBlock, S/block/RowBlock.java:413-440)
                blocks.addAll(RowBlock.getNullSuppressedRowFieldsFromBlock(row));
            }
            else {
                blocks.add(block);
            }
        }
        return new Page(page.getPositionCount(), blocks.toArray(Block[]::new));
    }

    /** first flattened channel of every plan channel; width[c] = fields of a ROW-typed channel, 1 otherwise */
    public static int[] firstFlatChannel(int[] width)
    {
        int[] first = new int[width.length];
        for (int channel = 0, at = 0; channel < width.length; channel++) {
            first[channel] = at;
            at += width[channel];
        }
        return first;
    }

    public static Page composeRows(Page flat, int[] width)
    {
        Block[] blocks = new Block[width.length];
        for (int channel = 0, at = 0; channel < width.length; at += width[channel], channel++) {
            if (width[channel] == 1) {
                blocks[channel] = flat.getBlock(at);
                continue;
            }
            Block[] fields = new Block[width[channel]];
            for (int field = 0; field < fields.length; field++) {
                fields[field] = flat.getBlock(at + field);
            }
            blocks[channel] = RowBlock.fromFieldBlocks(flat.getPositionCount(), fields);      // states are never NULL rows
        }
        return new Page(flat.getPositionCount(), blocks);
    }

    // ------------------------------------------------------------------------------------------------ decimal aggregation states
    /**
     * LongDecimalWithOverflowState travels as VARBINARY (LongDecimalWithOverflowStateSerializer.java:36-96): low, [high, [overflow]] as
     * little-endian longs, NULL for an empty state.  The C ABI wants an Int128ArrayBlock (sum) and a LongArrayBlock (overflow).
     * trino_b200/page.py (decode_decimal_sum_states / encode_decimal_sum_states, …_avg_…) is the tested mirror of these.
     */
    public static Block[] decodeDecimalSumStates(Block states)
    {
        int positions = states.getPositionCount();
        long[] words = new long[positions * 2];
        long[] overflow = new long[positions];
        boolean[] isNull = new boolean[positions];
        VariableWidthBlock flat = (VariableWidthBlock) flat(states);
        for (int position = 0; position < positions; position++) {
            if (flat.isNull(position)) {
                isNull[position] = true;
                continue;
            }
            Slice slice = flat.getRawSlice();
            int at = flat.getRawSliceOffset(position);
            int length = flat.getSliceLength(position);
            words[2 * position + 1] = slice.getLong(at);                                   // low
            words[2 * position] = length >= 16 ? slice.getLong(at + 8) : 0;                // high (Int128ArrayBlock: high word first)
            overflow[position] = length == 24 ? slice.getLong(at + 16) : 0;
        }
        return new Block[] {new Int128ArrayBlock(positions, Optional.of(isNull), words), new LongArrayBlock(positions, Optional.empty(), overflow)};
    }

    public static Block encodeDecimalSumStates(Int128ArrayBlock sums, LongArrayBlock overflows)
    {
        VariableWidthBlockBuilder out = new VariableWidthBlockBuilder(null, sums.getPositionCount(), sums.getPositionCount() * 24);
        for (int position = 0; position < sums.getPositionCount(); position++) {
            if (sums.isNull(position)) {
                out.appendNull();
                continue;
            }
            long high = sums.getInt128High(position);
            long low = sums.getInt128Low(position);
            long overflow = overflows.getLong(position);
            Slice buffer = Slices.allocate(24);
            buffer.setLong(0, low);
            buffer.setLong(8, high);
            buffer.setLong(16, overflow);
            out.writeEntry(buffer, 0, overflow != 0 ? 24 : (high != 0 ? 16 : 8));
        }
        return out.build();
    }

    /** LongDecimalWithOverflowAndLongStateSerializer.java:36-113: low, [high,] [count, overflow]; NULL when count == 0 */
    public static Block[] decodeDecimalAverageStates(Block states)
    {
        int positions = states.getPositionCount();
        long[] words = new long[positions * 2];
        long[] overflow = new long[positions];
        long[] count = new long[positions];
        boolean[] isNull = new boolean[positions];
        VariableWidthBlock flat = (VariableWidthBlock) flat(states);
        for (int position = 0; position < positions; position++) {
            if (flat.isNull(position)) {
                isNull[position] = true;
                continue;
            }
            Slice slice = flat.getRawSlice();
            int at = flat.getRawSliceOffset(position);
            int length = flat.getSliceLength(position);
            words[2 * position + 1] = slice.getLong(at);
            count[position] = 1;
            switch (length) {
                case 32 -> {
                    words[2 * position] = slice.getLong(at + 8);
                    count[position] = slice.getLong(at + 16);
                    overflow[position] = slice.getLong(at + 24);
                }
                case 24 -> {
                    count[position] = slice.getLong(at + 8);
                    overflow[position] = slice.getLong(at + 16);
                }
                case 16 -> words[2 * position] = slice.getLong(at + 8);
                default -> {}
            }
        }
        return new Block[] {new Int128ArrayBlock(positions, Optional.of(isNull), words), new LongArrayBlock(positions, Optional.empty(), overflow),
                new LongArrayBlock(positions, Optional.empty(), count)};
    }

    public static Block encodeDecimalAverageStates(Int128ArrayBlock sums, LongArrayBlock overflows, LongArrayBlock counts)
    {
        VariableWidthBlockBuilder out = new VariableWidthBlockBuilder(null, sums.getPositionCount(), sums.getPositionCount() * 32);
        for (int position = 0; position < sums.getPositionCount(); position++) {
            long count = counts.getLong(position);
            if (count == 0) {
                out.appendNull();
                continue;
            }
            long high = sums.isNull(position) ? 0 : sums.getInt128High(position);
            long low = sums.isNull(position) ? 0 : sums.getInt128Low(position);
            long overflow = overflows.getLong(position);
            Slice buffer = Slices.allocate(32);
            buffer.setLong(0, low);
            buffer.setLong(8, high);
            int countOffset = high == 0 ? 1 : 2;
            buffer.setLong(8 * countOffset, count);
            buffer.setLong(8 * (countOffset + 1), overflow);
            out.writeEntry(buffer, 0, 8 * (countOffset + ((overflow == 0 && count == 1) ? 0 : 2)));
        }
        return out.build();
    }

    // ------------------------------------------------------------------------------------------------ batching (Operator.addInput side)
    /** returns true when the batch should be flushed into the native operator now */
    public boolean append(Page page)
    {
        page = flattenRows(page);
        batch.add(page);
        batchRows += page.getPositionCount();
        return batchRows >= BATCH_ROWS;
    }

    public boolean hasBatch()
    {
        return batchRows > 0;
    }

    /**
     * Concatenates the batched pages column by column into the pinned staging region and describes them as ONE tgpu_page.
     * The native add_input copies host pages to the device before it returns, so the staging region is free again afterwards.
     */
    public MemorySegment flush(Arena arena)
    {
        stagingUsed = 0;
        int channels = channelTypes.length;
        MemorySegment columns = arena.allocate(COLUMN, channels);
        for (int channel = 0; channel < channels; channel++) {
            writeColumn(columns.asSlice(channel * COLUMN.byteSize(), COLUMN.byteSize()), channel, arena);
        }
        MemorySegment page = arena.allocate(PAGE);
        page.set(JAVA_INT, 0, channels);
        page.set(JAVA_INT, 4, 0);
        page.set(JAVA_LONG, 8, batchRows);
        page.set(ADDRESS, 16, columns);
        batch.clear();
        batchRows = 0;
        return page;
    }

    private MemorySegment reserve(long bytes)
    {
        long aligned = (stagingUsed + 63) & ~63L;
        if (aligned + bytes > staging.byteSize()) {
            throw new IllegalStateException("pinned staging region too small for one batch: " + (aligned + bytes) + " > " + staging.byteSize());
        }
        stagingUsed = aligned + bytes;
        return staging.asSlice(aligned, bytes);
    }

    private void writeColumn(MemorySegment column, int channel, Arena arena)
    {
        int type = channelTypes[channel];
        boolean anyNulls = false;
        for (Page page : batch) {
            anyNulls |= page.getBlock(channel).mayHaveNull();
        }
        MemorySegment nulls = anyNulls ? reserve(batchRows) : MemorySegment.NULL;
        MemorySegment data;
        MemorySegment offsets = MemorySegment.NULL;
        if (type == UTF8) {
            long bytes = 0;
            for (Page page : batch) {
                VariableWidthBlock block = (VariableWidthBlock) flat(page.getBlock(channel));
                bytes += block.getRawSliceOffset(block.getPositionCount()) - block.getRawSliceOffset(0);
            }
            data = reserve(Math.max(bytes, 1));
            offsets = reserve(4L * (batchRows + 1));
            long row = 0;
            long at = 0;
            for (Page page : batch) {
                VariableWidthBlock block = (VariableWidthBlock) flat(page.getBlock(channel));
                int count = block.getPositionCount();
                int first = block.getRawSliceOffset(0);
                int[] rawOffsets = block.getRawOffsets();
                int base = block.getRawArrayBase();
                for (int i = 0; i <= count; i++) {
                    offsets.setAtIndex(JAVA_INT, row + i, (int) (at + rawOffsets[base + i] - first));
                }
                int length = block.getRawSliceOffset(count) - first;
                Slice slice = block.getRawSlice();
                MemorySegment.copy(slice.byteArray(), slice.byteArrayOffset() + first, data, JAVA_BYTE, at, length);
                copyNulls(block.getRawValueIsNull(), base, count, nulls, row);
                row += count;
                at += length;
            }
        }
        else {
            int width = type == INT128 ? 16 : type == INT64 || type == FLOAT64 ? 8 : type == INT32 || type == FLOAT32 ? 4 : type == INT16 ? 2 : 1;
            data = reserve((long) width * batchRows);
            long row = 0;
            for (Page page : batch) {
                Block block = flat(page.getBlock(channel));
                int count = block.getPositionCount();
                switch (block) {
                    case LongArrayBlock longs -> {
                        MemorySegment.copy(longs.getRawValues(), longs.getRawValuesOffset(), data, JAVA_LONG, row * 8, count);
                        copyNulls(longs.getRawValueIsNull(), longs.getRawValuesOffset(), count, nulls, row);
                    }
                    case IntArrayBlock ints -> {
                        MemorySegment.copy(ints.getRawValues(), ints.getRawValuesOffset(), data, JAVA_INT, row * 4, count);
                        copyNulls(ints.getRawValueIsNull(), ints.getRawValuesOffset(), count, nulls, row);
                    }
                    case ShortArrayBlock shorts -> {
                        MemorySegment.copy(shorts.getRawValues(), shorts.getRawValuesOffset(), data, JAVA_SHORT, row * 2, count);
                        copyNulls(shorts.getRawValueIsNull(), shorts.getRawValuesOffset(), count, nulls, row);
                    }
                    case ByteArrayBlock bytes -> {
                        MemorySegment.copy(bytes.getRawValues(), bytes.getRawValuesOffset(), data, JAVA_BYTE, row, count);
                        copyNulls(bytes.getRawValueIsNull(), bytes.getRawValuesOffset(), count, nulls, row);
                    }
                    case Int128ArrayBlock wide -> {
                        MemorySegment.copy(wide.getRawValues(), 2 * wide.getRawOffset(), data, JAVA_LONG, row * 16, 2 * count);
                        copyNulls(wide.getRawValueIsNull(), wide.getRawOffset(), count, nulls, row);
                    }
                    default -> throw new IllegalArgumentException("block type without a GPU mapping: " + block.getClass().getSimpleName());
                }
                row += count;
            }
        }
        column.set(JAVA_INT, 0, type);
        column.set(JAVA_INT, 4, anyNulls ? COL_NULLS_BYTEMAP : 0);
        column.set(JAVA_LONG, 8, batchRows);
        column.set(ADDRESS, 16, data);
        column.set(ADDRESS, 24, offsets);
        column.set(ADDRESS, 32, nulls);
        column.set(ADDRESS, 40, MemorySegment.NULL);
    }

    /**
     * Dictionary and run-length encoded blocks are flattened here when several pages are batched (each page has its own dictionary);
     * a single-page call could pass TGPU_DICT32 / TGPU_RLE through instead - the library decodes both on ingest (Appendix B.5 of the
     * survey: values, not encodings, are the operator contract).
     */
    private static Block flat(Block block)
    {
        return switch (block) {
            case DictionaryBlock dictionary -> dictionary.getDictionary().copyPositions(dictionary.getRawIds(), dictionary.getRawIdsOffset(), dictionary.getPositionCount());
            case RunLengthEncodedBlock rle -> rle.getValue().copyPositions(new int[rle.getPositionCount()], 0, rle.getPositionCount());
            default -> block;
        };
    }

    private static void copyNulls(boolean[] valueIsNull, int offset, int count, MemorySegment nulls, long row)
    {
        if (nulls.equals(MemorySegment.NULL)) {
            return;
        }
        if (valueIsNull == null) {
            nulls.asSlice(row, count).fill((byte) 0);
            return;
        }
        for (int i = 0; i < count; i++) {
            nulls.set(JAVA_BYTE, row + i, (byte) (valueIsNull[offset + i] ? 1 : 0));
        }
    }

    // ------------------------------------------------------------------------------------------------ Operator.getOutput side
    /** what the caller must know to size host buffers for tgpu_page_copy_to_host */
    public record OutputShape(int[] types, long rows, long[] utf8Bytes) {}

    /**
     * Describes host landing buffers (inside the pinned staging region) for a device page of the given shape; the caller then invokes
     * tgpu_page_copy_to_host(ctx, devicePage, hostPage) and {@link #toPages} turns the landed bytes into <= 8192-row Java pages.
     * Columns for which tgpu_page_passthrough_channel names an input channel get data == NULL: they are not copied back
     * (LookupJoinPageBuilder.build :144-150 returns the probe blocks themselves).
     */
    public MemorySegment describeLanding(OutputShape shape, int[] passthroughChannel, Arena arena)
    {
        stagingUsed = 0;
        int channels = shape.types().length;
        MemorySegment columns = arena.allocate(COLUMN, channels);
        for (int channel = 0; channel < channels; channel++) {
            MemorySegment column = columns.asSlice(channel * COLUMN.byteSize(), COLUMN.byteSize());
            int type = shape.types()[channel];
            boolean skip = passthroughChannel[channel] >= 0;
            int width = type == INT128 ? 16 : type == INT64 || type == FLOAT64 ? 8 : type == INT32 || type == FLOAT32 ? 4 : type == INT16 ? 2 : 1;
            column.set(JAVA_INT, 0, type);
            column.set(JAVA_INT, 4, 0);
            column.set(JAVA_LONG, 8, shape.rows());
            column.set(ADDRESS, 16, skip ? MemorySegment.NULL : reserve(type == UTF8 ? Math.max(shape.utf8Bytes()[channel], 1) : width * shape.rows()));
            column.set(ADDRESS, 24, type == UTF8 && !skip ? reserve(4 * (shape.rows() + 1)) : MemorySegment.NULL);
            column.set(ADDRESS, 32, reserve((shape.rows() + 7) / 8 + 8));       // Arrow validity bitmap written by the copy
            column.set(ADDRESS, 40, MemorySegment.NULL);
        }
        MemorySegment page = arena.allocate(PAGE);
        page.set(JAVA_INT, 0, channels);
        page.set(JAVA_INT, 4, 0);
        page.set(JAVA_LONG, 8, shape.rows());
        page.set(ADDRESS, 16, columns);
        return page;
    }

    /** landed host page -> Java pages of at most `maxRows` positions; `hasNulls[c]` from the device page's validity pointers */
    public static List<Page> toPages(MemorySegment hostPage, OutputShape shape, boolean[] hasNulls, Block[][] passthrough, int maxRows)
    {
        MemorySegment columns = hostPage.get(ADDRESS, 16).reinterpret(COLUMN.byteSize() * shape.types().length);
        List<Page> pages = new ArrayList<>();
        for (long first = 0; first < shape.rows(); first += maxRows) {
            int count = (int) Math.min(maxRows, shape.rows() - first);
            Block[] blocks = new Block[shape.types().length];
            for (int channel = 0; channel < blocks.length; channel++) {
                if (passthrough[channel] != null) {
                    blocks[channel] = passthrough[channel][(int) (first / maxRows)];     // the caller's own input block, unchanged
                    continue;
                }
                MemorySegment column = columns.asSlice(channel * COLUMN.byteSize(), COLUMN.byteSize());
                blocks[channel] = readBlock(column, shape.types()[channel], first, count, hasNulls[channel]);
            }
            pages.add(new Page(count, blocks));
        }
        return pages;
    }

    private static Block readBlock(MemorySegment column, int type, long first, int count, boolean hasNulls)
    {
        Optional<boolean[]> nulls = Optional.empty();
        if (hasNulls) {
            MemorySegment bitmap = column.get(ADDRESS, 32).reinterpret((first + count + 7) / 8 + 8);
            boolean[] isNull = new boolean[count];
            for (int i = 0; i < count; i++) {
                long bit = first + i;
                isNull[i] = ((bitmap.get(JAVA_BYTE, bit >> 3) >> (bit & 7)) & 1) == 0;
            }
            nulls = Optional.of(isNull);
        }
        MemorySegment data = column.get(ADDRESS, 16);
        switch (type) {
            case INT64, FLOAT64 -> {
                long[] values = new long[count];
                MemorySegment.copy(data.reinterpret((first + count) * 8), JAVA_LONG, first * 8, values, 0, count);
                return new LongArrayBlock(count, nulls, values);
            }
            case INT32, FLOAT32 -> {
                int[] values = new int[count];
                MemorySegment.copy(data.reinterpret((first + count) * 4), JAVA_INT, first * 4, values, 0, count);
                return new IntArrayBlock(count, nulls, values);
            }
            case INT16 -> {
                short[] values = new short[count];
                MemorySegment.copy(data.reinterpret((first + count) * 2), JAVA_SHORT, first * 2, values, 0, count);
                return new ShortArrayBlock(count, nulls, values);
            }
            case INT8 -> {
                byte[] values = new byte[count];
                MemorySegment.copy(data.reinterpret(first + count), JAVA_BYTE, first, values, 0, count);
                return new ByteArrayBlock(count, nulls, values);
            }
            case INT128 -> {
                long[] values = new long[2 * count];
                MemorySegment.copy(data.reinterpret((first + count) * 16), JAVA_LONG, first * 16, values, 0, 2 * count);
                return new Int128ArrayBlock(count, nulls, values);
            }
            case UTF8 -> {
                MemorySegment offsets = column.get(ADDRESS, 24).reinterpret((first + count + 1) * 4);
                int[] newOffsets = new int[count + 1];
                int start = offsets.getAtIndex(ValueLayout.JAVA_INT, first);
                for (int i = 0; i <= count; i++) {
                    newOffsets[i] = offsets.getAtIndex(ValueLayout.JAVA_INT, first + i) - start;
                }
                byte[] bytes = new byte[newOffsets[count]];
                MemorySegment.copy(data.reinterpret((long) start + bytes.length), JAVA_BYTE, start, bytes, 0, bytes.length);
                return new VariableWidthBlock(count, Slices.wrappedBuffer(bytes), newOffsets, nulls);
            }
            default -> throw new IllegalArgumentException("tgpu_type " + type);
        }
    }
}
