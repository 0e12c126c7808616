"""Exchange between GPU stages of one box: the host-side plumbing (rank discovery, communicator id and IPC handle
distribution) around tgpu_comm_* / tgpu_exchange_*.

Stands in for the pair PartitionedOutputOperator -> OutputBuffer ... DirectExchangeClient -> ExchangeOperator
(M/operator/output/PartitionedOutputOperator.java:335-357, M/operator/ExchangeOperator.java) when both ends are GPU stages:
  partitioned(...)  FIXED_HASH_DISTRIBUTION (M/sql/planner/SystemPartitioningHandle.java:49): partition p goes to rank p
  broadcast(...)    FIXED_BROADCAST_DISTRIBUTION (:51): every rank receives every page (the build side of a REPLICATED join)
`dist` is torch.distributed (or any module with broadcast / all_gather over device tensors): it only carries the 128-byte
communicator id and the arena handles, never data.
"""
import ctypes as C

from . import abi
from . import operators as ops


class Exchange:
    def __init__(self, ctx, dist, rank, world, local):
        self.ctx, self.dist, self.rank, self.world, self.local = ctx, dist, rank, world, local
        self.arenas = False
        if world == 1:
            return
        import torch
        lib = ctx.lib
        idb = (C.c_uint8 * abi.COMM_ID_BYTES)()
        if rank == 0:
            ctx.check(lib.tgpu_comm_get_unique_id(C.cast(idb, C.c_void_p)))
        t = torch.tensor(list(idb), dtype=torch.uint8, device=f"cuda:{local}")
        dist.broadcast(t, 0)
        idb = (C.c_uint8 * abi.COMM_ID_BYTES)(*t.cpu().tolist())
        ctx.check(lib.tgpu_comm_init(ctx.h, C.cast(idb, C.c_void_p), rank, world))

    def create_arenas(self, arena_bytes):
        """receive arenas for the peer-memory / split-phase transports (create them AFTER the build sides were exchanged: a page received
        through an arena aliases it, and a hash build keeps its input)"""
        if self.world == 1 or self.arenas:
            return
        import torch
        lib, ctx = self.ctx.lib, self.ctx
        hb = (C.c_uint8 * (abi.NUM_ARENAS * abi.IPC_HANDLE_BYTES))()
        ctx.check(lib.tgpu_comm_arena_create(ctx.h, int(arena_bytes), C.cast(hb, C.c_void_p)))
        mine = torch.tensor(list(hb), dtype=torch.uint8, device=f"cuda:{self.local}")
        gathered = [torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(gathered, mine)
        allh = (C.c_uint8 * (self.world * abi.NUM_ARENAS * abi.IPC_HANDLE_BYTES))(*torch.cat(gathered).cpu().tolist())
        ctx.check(lib.tgpu_comm_arena_open(ctx.h, C.cast(allh, C.c_void_p)))
        self.arenas = True

    def partitioner(self, channels):
        return ops.PartitionedOutputOperatorFactory(self.ctx, list(channels), self.world).create_operator()

    def partitioned(self, partitioner, page):
        """hash exchange of a device page; returns a DeviceOutputPage (world == 1: None, the caller keeps its page)"""
        if self.world == 1:
            return None
        ctx, lib = self.ctx, self.ctx.lib
        pp = abi.PP()
        if self.arenas:
            h = C.c_void_p()
            ctx.check(lib.tgpu_exchange_begin(ctx.h, partitioner.h, page.ref(), C.byref(h)))
            ctx.check(lib.tgpu_exchange_end(ctx.h, h, C.byref(pp)))
        else:
            ctx.check(lib.tgpu_exchange_partitioned(ctx.h, partitioner.h, page.ref(), C.byref(pp)))
        return ops.DeviceOutputPage(ctx, pp)

    def broadcast(self, page):
        ctx = self.ctx
        pp = abi.PP()
        ctx.check(ctx.lib.tgpu_exchange_broadcast(ctx.h, page.ref(), C.byref(pp)))
        return ops.DeviceOutputPage(ctx, pp)

    def close(self):
        if self.world > 1:
            self.ctx.check(self.ctx.lib.tgpu_comm_destroy(self.ctx.h))
