/*
 * Base of every GPU-backed operator: forwards the Operator protocol to a tgpu_op handle.
 * Replaces nothing in the reference by itself; the concrete factories below are what LocalExecutionPlanner instantiates.
 * NOT compiled here (no JDK).  Class lives in io.trino.operator so it can use package-private operator APIs.
 */
package io.trino.operator.gpu;

import io.trino.memory.context.LocalMemoryContext;
import io.trino.operator.Operator;
import io.trino.operator.OperatorContext;
import io.trino.spi.Page;
import io.trino.spi.TrinoException;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import static io.trino.spi.StandardErrorCode.DIVISION_BY_ZERO;
import static io.trino.spi.StandardErrorCode.GENERIC_INSUFFICIENT_RESOURCES;
import static io.trino.spi.StandardErrorCode.GENERIC_INTERNAL_ERROR;
import static io.trino.spi.StandardErrorCode.NUMERIC_VALUE_OUT_OF_RANGE;
import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_INT;

public class GpuOperator
        implements Operator
{
    protected final OperatorContext operatorContext;
    protected final LocalMemoryContext memoryContext;
    protected final MemorySegment ctx;       // tgpu_ctx*, one per driver thread (GpuContexts.forCurrentDriver())
    protected final MemorySegment op;        // tgpu_op*
    private final PageMarshaller marshaller; // Page <-> tgpu_page (pinned staging, boolean[] nulls passed as byte maps)

    protected GpuOperator(OperatorContext operatorContext, MemorySegment ctx, MemorySegment op, PageMarshaller marshaller)
    {
        this.operatorContext = operatorContext;
        this.memoryContext = operatorContext.localUserMemoryContext();
        this.ctx = ctx;
        this.op = op;
        this.marshaller = marshaller;
    }

    @Override
    public OperatorContext getOperatorContext()
    {
        return operatorContext;
    }

    @Override
    public boolean needsInput()
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment out = arena.allocate(JAVA_INT);
            check((int) TrinoGpuLibrary.OP_NEEDS_INPUT.invokeExact(op, out));
            return out.get(JAVA_INT, 0) != 0;
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    @Override
    public void addInput(Page page)
    {
        try (Arena arena = Arena.ofConfined()) {
            // batches several 8192-row Java pages into one >= 1M-row tgpu_page before crossing (see INTEGRATION.md §3)
            MemorySegment nativePage = marshaller.toNative(page, arena);
            check((int) TrinoGpuLibrary.OP_ADD_INPUT.invokeExact(op, nativePage));
            memoryContext.setBytes((long) TrinoGpuLibrary.OP_MEMORY_BYTES.invokeExact(op));
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    @Override
    public Page getOutput()
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment out = arena.allocate(ADDRESS);
            check((int) TrinoGpuLibrary.OP_GET_OUTPUT.invokeExact(op, out));
            MemorySegment devicePage = out.get(ADDRESS, 0);
            if (devicePage.equals(MemorySegment.NULL)) {
                return null;
            }
            try {
                return marshaller.toJava(ctx, devicePage);   // tgpu_page_copy_to_host into Block arrays
            }
            finally {
                TrinoGpuLibrary.PAGE_RELEASE.invokeExact(ctx, devicePage);
            }
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    @Override
    public void finish()
    {
        try {
            check((int) TrinoGpuLibrary.OP_FINISH.invokeExact(op));
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    @Override
    public boolean isFinished()
    {
        try (Arena arena = Arena.ofConfined()) {
            MemorySegment out = arena.allocate(JAVA_INT);
            check((int) TrinoGpuLibrary.OP_IS_FINISHED.invokeExact(op, out));
            return out.get(JAVA_INT, 0) != 0;
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    @Override
    public void close()
    {
        try {
            TrinoGpuLibrary.OP_CLOSE.invokeExact(op);
            memoryContext.setBytes(0);
        }
        catch (Throwable e) {
            throw propagate(e);
        }
    }

    protected void check(int status)
    {
        if (status == 0) {
            return;
        }
        String message = TrinoGpuLibrary.lastError(ctx);
        throw switch (status) {
            case -3 -> new TrinoException(GENERIC_INSUFFICIENT_RESOURCES, message);
            case -4 -> new TrinoException(NUMERIC_VALUE_OUT_OF_RANGE, message);
            case -5 -> new TrinoException(DIVISION_BY_ZERO, message);
            case -1 -> new IllegalArgumentException(message);
            case -7 -> new IllegalStateException(message);
            default -> new TrinoException(GENERIC_INTERNAL_ERROR, message);   // -2 device failure, -6 handled at plan time
        };
    }

    private static RuntimeException propagate(Throwable e)
    {
        if (e instanceof RuntimeException runtimeException) {
            return runtimeException;
        }
        return new RuntimeException(e);
    }
}
