/*
 * One tgpu_ctx (CUDA stream, stream-ordered memory pool, error slot) per driver thread, plus the pinned staging region its
 * PageMarshallers use.  A tgpu handle is used by one thread at a time (include/trino_gpu.h), exactly like an Operator instance
 * (M/operator/Driver.java:298); drivers of one task run concurrently on different threads (TaskManagerConfig.java:65), each gets its own
 * context, all share the device and - through GpuLookupSource - the join tables.  NOT compiled here (no JDK).
 */
package io.trino.operator.gpu;

import io.trino.operator.DriverContext;
import io.trino.spi.TrinoException;
import io.trino.spi.block.PageMarshaller;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.concurrent.ConcurrentHashMap;

import static io.trino.spi.StandardErrorCode.GENERIC_INTERNAL_ERROR;
import static java.lang.foreign.ValueLayout.ADDRESS;

public final class GpuContexts
{
    /** pinned staging per driver: two batches of BATCH_ROWS rows of 64 bytes (input side) and the landing zone of one output page */
    private static final long STAGING_BYTES = 256L << 20;
    private static final ConcurrentHashMap<Thread, Handle> HANDLES = new ConcurrentHashMap<>();
    private static final int DEVICE = Integer.getInteger("trino.gpu.device", 0);

    private GpuContexts() {}

    public record Handle(MemorySegment context, MemorySegment staging)
    {
        public PageMarshaller marshaller(int[] channelTypes)
        {
            return new PageMarshaller(staging, channelTypes);
        }
    }

    /** the context of the thread that runs this driver; created on first use, destroyed by {@link #release} when the task ends */
    public static Handle forCurrentDriver(DriverContext driverContext)
    {
        return HANDLES.computeIfAbsent(Thread.currentThread(), thread -> create());
    }

    private static Handle create()
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment out = arena.allocate(ADDRESS);
            int status = (int) TrinoGpuLibrary.CTX_CREATE.invokeExact(DEVICE, out);
            if (status != 0) {
                // no CPU fallback inside the library: without a device the planner must not have chosen the GPU factories
                throw new TrinoException(GENERIC_INTERNAL_ERROR, "tgpu_ctx_create failed: " + status);
            }
            MemorySegment context = out.get(ADDRESS, 0);
            status = (int) TrinoGpuLibrary.HOST_ALLOC_PINNED.invokeExact(STAGING_BYTES, out);
            if (status != 0) {
                throw new TrinoException(GENERIC_INTERNAL_ERROR, "tgpu_host_alloc_pinned failed: " + status);
            }
            return new Handle(context, out.get(ADDRESS, 0).reinterpret(STAGING_BYTES));
        }
        catch (RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new RuntimeException(e);
        }
    }

    public static void release(Thread thread)
    {
        Handle handle = HANDLES.remove(thread);
        if (handle == null) {
            return;
        }
        try {
            TrinoGpuLibrary.HOST_FREE_PINNED.invokeExact(handle.staging());
            TrinoGpuLibrary.CTX_DESTROY.invokeExact(handle.context());
        }
        catch (Throwable e) {
            throw new RuntimeException(e);
        }
    }
}
