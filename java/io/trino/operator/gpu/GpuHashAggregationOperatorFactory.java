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

import java.util.List;

import static com.google.common.base.Preconditions.checkState;

public class GpuHashAggregationOperatorFactory
        implements OperatorFactory
{
    /** function id + input/mask channels of one aggregate: the serialisable part of an AggregatorFactory */
    public record GpuAggregate(int function, int inputChannel, int maskChannel) {}

    private final int operatorId;
    private final PlanNodeId planNodeId;
    private final List<Integer> groupByChannels;
    private final Step step;
    private final List<GpuAggregate> aggregates;
    private final int expectedGroups;
    private final long maxPartialMemory;
    private boolean closed;

    public GpuHashAggregationOperatorFactory(int operatorId, PlanNodeId planNodeId, List<Integer> groupByChannels, Step step,
            List<GpuAggregate> aggregates, int expectedGroups, long maxPartialMemory)
    {
        this.operatorId = operatorId;
        this.planNodeId = planNodeId;
        this.groupByChannels = List.copyOf(groupByChannels);
        this.step = step;
        this.aggregates = List.copyOf(aggregates);
        this.expectedGroups = expectedGroups;
        this.maxPartialMemory = maxPartialMemory;
    }

    @Override
    public Operator createOperator(DriverContext driverContext)
    {
        checkState(!closed, "Factory is already closed");
        OperatorContext operatorContext = driverContext.addOperatorContext(operatorId, planNodeId, "GpuHashAggregationOperator");
        GpuContexts.Handle gpu = GpuContexts.forCurrentDriver(driverContext);
        // fills a tgpu_agg_spec {num_keys, key_channels, step, num_aggs, aggs, expected_groups, max_partial_bytes, pre} and calls tgpu_agg_create
        return new GpuOperator(operatorContext, gpu.context(), NativeSpecs.createAggregation(gpu, groupByChannels, step, aggregates, expectedGroups, maxPartialMemory), gpu.marshaller());
    }

    @Override
    public void noMoreOperators()
    {
        closed = true;
    }

    @Override
    public OperatorFactory duplicate()
    {
        return new GpuHashAggregationOperatorFactory(operatorId, planNodeId, groupByChannels, step, aggregates, expectedGroups, maxPartialMemory);
    }
}
