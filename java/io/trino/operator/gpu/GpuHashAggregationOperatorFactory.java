/*
 * Drop-in for HashAggregationOperator.HashAggregationOperatorFactory (M/operator/HashAggregationOperator.java:63-200),
 * instantiated inside LocalExecutionPlanner.createHashAggregationOperatorFactory (:4040-4118) where the
 * Map<Symbol, Aggregation> still names the resolved functions (see INTEGRATION.md §2).  NOT compiled here.
 */
package io.trino.operator.gpu;

import io.trino.operator.DriverContext;
import io.trino.operator.Operator;
import io.trino.operator.OperatorContext;
import io.trino.operator.OperatorFactory;
import io.trino.sql.planner.plan.AggregationNode.Step;
import io.trino.sql.planner.plan.PlanNodeId;

import java.lang.foreign.MemorySegment;
import java.util.List;
import java.util.OptionalInt;

import static com.google.common.base.Preconditions.checkState;

public class GpuHashAggregationOperatorFactory
        implements OperatorFactory
{
    /** function id (tgpu_agg_function) + input/mask channels of one aggregate (+ the tgpu_type of an avg(decimal) result): the serialisable part of an AggregatorFactory */
    public record GpuAggregate(int function, int inputChannel, int maskChannel, int resultType)
    {
        public GpuAggregate(int function, int inputChannel, int maskChannel)
        {
            this(function, inputChannel, maskChannel, 0);
        }
    }

    private final int operatorId;
    private final PlanNodeId planNodeId;
    private final int[] inputTypes;              // tgpu_type per source channel (source.getTypes() at :4086-4089)
    private final int[] outputTypes;             // group-by types, then one (SINGLE/FINAL) or the state columns (PARTIAL) per aggregate
    private final List<Integer> groupByChannels;
    private final List<Integer> globalAggregationGroupIds;
    private final Step step;
    private final List<GpuAggregate> aggregates;
    private final OptionalInt groupIdChannel;    // index among the group-by keys, as in the reference (:552)
    private final int expectedGroups;
    private final long maxPartialMemory;
    // the native PartialAggregationController shared by this plan node's drivers (MemorySegment.NULL = none).  Created by the planner hook
    // from task.max-partial-aggregation-memory and adaptive-partial-aggregation.unique-rows-ratio-threshold where the reference creates
    // its own (LocalExecutionPlanner.java:4101-4110), destroyed with the task.
    private final MemorySegment partialAggregationController;
    private final double uniqueRowsRatioThreshold;
    private boolean closed;

    public GpuHashAggregationOperatorFactory(int operatorId, PlanNodeId planNodeId, int[] inputTypes, int[] outputTypes, List<Integer> groupByChannels,
            List<Integer> globalAggregationGroupIds, Step step, List<GpuAggregate> aggregates, OptionalInt groupIdChannel, int expectedGroups, long maxPartialMemory,
            boolean adaptivePartialAggregation, double uniqueRowsRatioThreshold)
    {
        this.operatorId = operatorId;
        this.planNodeId = planNodeId;
        this.inputTypes = inputTypes.clone();
        this.outputTypes = outputTypes.clone();
        this.groupByChannels = List.copyOf(groupByChannels);
        this.globalAggregationGroupIds = List.copyOf(globalAggregationGroupIds);
        this.step = step;
        this.aggregates = List.copyOf(aggregates);
        this.groupIdChannel = groupIdChannel;
        this.expectedGroups = expectedGroups;
        this.maxPartialMemory = maxPartialMemory;
        this.uniqueRowsRatioThreshold = uniqueRowsRatioThreshold;
        this.partialAggregationController = adaptivePartialAggregation && step.isOutputPartial()
                ? createController(maxPartialMemory, uniqueRowsRatioThreshold)
                : MemorySegment.NULL;
    }

    private static MemorySegment createController(long maxPartialMemory, double uniqueRowsRatioThreshold)
    {
        try (java.lang.foreign.Arena arena = java.lang.foreign.Arena.ofConfined()) {
            MemorySegment out = arena.allocate(java.lang.foreign.ValueLayout.ADDRESS);
            int status = (int) TrinoGpuLibrary.PA_CONTROLLER_CREATE.invokeExact(maxPartialMemory, uniqueRowsRatioThreshold, out);
            checkState(status == 0, "tgpu_partial_agg_controller_create failed: %s", status);
            return out.get(java.lang.foreign.ValueLayout.ADDRESS, 0);
        }
        catch (RuntimeException e) {
            throw e;
        }
        catch (Throwable e) {
            throw new RuntimeException(e);
        }
    }

    @Override
    public Operator createOperator(DriverContext driverContext)
    {
        checkState(!closed, "Factory is already closed");
        OperatorContext operatorContext = driverContext.addOperatorContext(operatorId, planNodeId, "GpuHashAggregationOperator");
        GpuContexts.Handle gpu = GpuContexts.forCurrentDriver(driverContext);
        MemorySegment op = NativeSpecs.createAggregation(gpu, groupByChannels, step, aggregates, expectedGroups, maxPartialMemory, globalAggregationGroupIds,
                groupIdChannel.orElse(-1), inputTypes, MemorySegment.NULL, partialAggregationController);
        return new GpuOperator(operatorContext, gpu.context(), op, gpu.marshaller(inputTypes), outputTypes);
    }

    @Override
    public void noMoreOperators()
    {
        closed = true;
    }

    @Override
    public OperatorFactory duplicate()
    {
        return new GpuHashAggregationOperatorFactory(operatorId, planNodeId, inputTypes, outputTypes, groupByChannels, globalAggregationGroupIds, step, aggregates,
                groupIdChannel, expectedGroups, maxPartialMemory,
                // a duplicated factory gets its own controller (HashAggregationOperatorFactory.duplicate :238)
                !partialAggregationController.equals(MemorySegment.NULL), uniqueRowsRatioThreshold);
    }
}
