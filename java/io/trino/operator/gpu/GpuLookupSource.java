/*
 * The device-resident join table as it travels through the reference's join bridge.  A GPU HashBuilderOperator lends it through
 * PartitionedLookupSourceFactory.lendPartitionLookupSource(partition, supplier) (M/operator/join/unspilled/PartitionedLookupSourceFactory.java:100)
 * and a GPU probe unwraps it from the LookupSource the bridge's createLookupSource() future delivers (:126); on one GPU there is one
 * partition (P = 1: no LocalPartitionGenerator is needed in front of the table, SURVEY.md §8 a14).
 *
 * Only the bookkeeping methods of LookupSource are meaningful on the host - positions are resolved by the native probe operator
 * (tgpu_join_probe_create binds the same tgpu_lookup*), so the row-at-a-time methods refuse to run.  NOT compiled here (no JDK).
 */
package io.trino.operator.gpu;

import io.trino.operator.join.LookupSource;
import io.trino.spi.Page;
import io.trino.spi.PageBuilder;

import java.lang.foreign.MemorySegment;
import java.util.concurrent.atomic.AtomicBoolean;

public final class GpuLookupSource
        implements LookupSource
{
    private final MemorySegment lookup;      // tgpu_lookup*, reference counted by the library (tgpu_lookup_release)
    private final AtomicBoolean closed = new AtomicBoolean();

    public GpuLookupSource(MemorySegment lookup)
    {
        this.lookup = lookup;
    }

    public MemorySegment handle()
    {
        return lookup;
    }

    @Override
    public long getInMemorySizeInBytes()
    {
        try {
            return (long) TrinoGpuLibrary.LOOKUP_MEMORY_BYTES.invokeExact(lookup);
        }
        catch (Throwable e) {
            throw new RuntimeException(e);
        }
    }

    @Override
    public long getJoinPositionCount()
    {
        try {
            return (long) TrinoGpuLibrary.LOOKUP_POSITION_COUNT.invokeExact(lookup);
        }
        catch (Throwable e) {
            throw new RuntimeException(e);
        }
    }

    @Override
    public boolean isEmpty()
    {
        return getJoinPositionCount() == 0;
    }

    @Override
    public long joinPositionWithinPartition(long joinPosition)
    {
        return joinPosition;    // one partition
    }

    @Override
    public long getJoinPosition(int position, Page hashChannelsPage, Page allChannelsPage, long rawHash)
    {
        throw new UnsupportedOperationException("a GPU lookup source is probed by GpuLookupJoinOperatorFactory's operator (whole pages), not row by row");
    }

    @Override
    public long getJoinPosition(int position, Page hashChannelsPage, Page allChannelsPage)
    {
        throw new UnsupportedOperationException("a GPU lookup source is probed by GpuLookupJoinOperatorFactory's operator (whole pages), not row by row");
    }

    @Override
    public long getNextJoinPosition(long currentJoinPosition, int probePosition, Page allProbeChannelsPage)
    {
        throw new UnsupportedOperationException();
    }

    @Override
    public void appendTo(long position, PageBuilder pageBuilder, int outputChannelOffset)
    {
        throw new UnsupportedOperationException();
    }

    @Override
    public boolean isJoinPositionEligible(long currentJoinPosition, int probePosition, Page allProbeChannelsPage)
    {
        throw new UnsupportedOperationException();
    }

    @Override
    public void close()
    {
        if (closed.compareAndSet(false, true)) {
            try {
                TrinoGpuLibrary.LOOKUP_RELEASE.invokeExact(lookup);
            }
            catch (Throwable e) {
                throw new RuntimeException(e);
            }
        }
    }
}
