"""Host-side planning for the multi-GPU (one process per GPU) path: range shards of the synthetic tables and the
receive layout of the all-to-all exchange.  Mirrors what tgpu_exchange_partitioned does inside the library
(csrc/partition.cu) so the N>1 logic can be tested on CPU with the gloo backend."""


def shard_range(total_rows, world, rank):
    """contiguous range shard [first, first+count) of a table, last rank takes the remainder"""
    per = total_rows // world
    first = per * rank
    count = per if rank < world - 1 else total_rows - first
    return first, count


def exchange_plan(count_matrix, rank):
    """count_matrix[src][dst] = rows rank `src` sends to rank `dst` (the all-gathered send counts).
    Returns (recv_counts, recv_offsets, total): rows arrive grouped by source rank, in rank order."""
    world = len(count_matrix)
    recv_counts = [count_matrix[src][rank] for src in range(world)]
    offsets = [0]
    for c in recv_counts:
        offsets.append(offsets[-1] + c)
    return recv_counts, offsets[:-1], offsets[-1]
