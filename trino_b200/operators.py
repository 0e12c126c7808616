"""Host-side mirror of the reference's Operator / OperatorFactory interface over the C ABI.

The reference's host language is Java and there is no JDK in this environment, so the drop-in boundary is
the C ABI (include/trino_gpu.h); these classes are the Python equivalent of the thin Java operators in
java/  and keep the reference's names, argument meaning and error behaviour:

  Operator            M/operator/Operator.java:21-102  (needsInput/addInput/getOutput/finish/isFinished/close)
  OperatorFactory     M/operator/OperatorFactory.java:16-30 (createOperator/noMoreOperators/duplicate)
  HashAggregationOperatorFactory   M/operator/HashAggregationOperator.java:63-200
  HashBuilderOperatorFactory       M/operator/join/unspilled/HashBuilderOperator.java:55-140
  LookupJoinOperatorFactory        M/operator/join/unspilled/LookupJoinOperatorFactory.java
  FilterAndProjectOperatorFactory  M/operator/FilterAndProjectOperator.java:97-150
  PartitionedOutputOperatorFactory M/operator/output/PartitionedOutputOperator.java:52-140

Every method ends in a libtrino_gpu.so call; nothing here computes on the CPU.
"""
import ctypes as C

import numpy as np

from . import abi
from .page import AbiPage, Block, Page, RowBlock, compose_row_blocks, compose_state_blocks, flatten_row_blocks, flatten_state_blocks

_NP_OF_TYPE = {abi.INT64: np.int64, abi.INT32: np.int32, abi.INT16: np.int16, abi.INT8: np.int8, abi.FLOAT64: np.float64, abi.FLOAT32: np.float32}
_ELEM = {abi.INT64: 8, abi.INT32: 4, abi.INT16: 2, abi.INT8: 1, abi.FLOAT64: 8, abi.FLOAT32: 4}


def _i32(values):
    arr = (C.c_int32 * max(1, len(values)))(*values)
    return arr


class Context:
    """tgpu_ctx: one per (process, device, driver thread)."""

    def __init__(self, device=0):
        self.lib = abi.load_library()
        h = C.c_void_p()
        st = self.lib.tgpu_ctx_create(device, C.byref(h))
        if st != 0:
            raise abi.TrinoGpuError(st, self.lib.tgpu_status_name(st).decode(), self.lib.tgpu_last_error(None).decode())
        self.h = h
        self.device = device

    def check(self, st):
        if st != 0:
            raise abi.TrinoGpuError(st, self.lib.tgpu_status_name(st).decode(), self.lib.tgpu_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.lib.tgpu_ctx_destroy(self.h)
            self.h = None

    def synchronize(self):
        self.check(self.lib.tgpu_ctx_synchronize(self.h))

    @property
    def kernel_launches(self):
        return self.lib.tgpu_ctx_kernel_launches(self.h)

    # ---- device memory
    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.tgpu_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        self.check(self.lib.tgpu_free(self.h, C.c_void_p(ptr)))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        p = self.malloc(max(arr.nbytes, 16))
        if arr.nbytes:
            self.check(self.lib.tgpu_memcpy_h2d(self.h, C.c_void_p(p), C.c_void_p(arr.ctypes.data), arr.nbytes))
        return p

    def to_host(self, ptr, dtype, count):
        out = np.empty(count, dtype=dtype)
        if out.nbytes:
            self.check(self.lib.tgpu_memcpy_d2h(self.h, C.c_void_p(out.ctypes.data), C.c_void_p(ptr), out.nbytes))
        return out

    def pinned_empty(self, count, dtype):
        p = C.c_void_p()
        nbytes = int(count) * np.dtype(dtype).itemsize
        st = self.lib.tgpu_host_alloc_pinned(max(nbytes, 16), C.byref(p))
        self.check(st)
        buf = (C.c_char * max(nbytes, 16)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=count)
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(p)   # freed with the process; pinned buffers live as long as the bench
        return arr

    def flush_l2(self):
        self.check(self.lib.tgpu_flush_l2(self.h))

    def timer_start(self):
        self.check(self.lib.tgpu_timer_start(self.h))

    def timer_stop_ms(self):
        ms = C.c_float()
        self.check(self.lib.tgpu_timer_stop_ms(self.h, C.byref(ms)))
        return ms.value

    def last_kernel_ms(self):
        ms = C.c_float()
        self.check(self.lib.tgpu_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return ms.value

    # ---- output pages
    def page_to_host(self, pp, release=True, views_of=None):
        """device tgpu_page* -> host Page (numpy).  `views_of`: the host input Page of a by-reference probe; output columns
        without device data are views of its blocks (tgpu_page_passthrough_channel)."""
        dp = pp.contents
        n = dp.num_rows
        host_cols = (abi.Column * max(1, dp.num_columns))()
        keep = []
        for c in range(dp.num_columns):
            d = dp.columns[c]
            h = host_cols[c]
            h.type = d.type
            h.length = n
            if views_of is not None and n > 0 and not d.data:
                src = C.c_int32(-1)
                self.check(self.lib.tgpu_page_passthrough_channel(pp, c, C.byref(src)))
                if src.value < 0:
                    raise RuntimeError(f"output column {c} has no device data and is not a pass-through view")
                h.data = None
                keep.append(("view", views_of.get_block(src.value), None, None))
                continue
            valid = np.empty((n + 7) // 8 + 1, dtype=np.uint8)
            h.validity = valid.ctypes.data
            if d.type == abi.UTF8:
                nbytes = max(1, self.lib.tgpu_page_utf8_bytes(self.h, pp, c))
                # offsets may be absolute into a larger buffer: size the host buffer by the last offset
                offs = np.empty(n + 1, dtype=np.int32)
                if n:
                    self.check(self.lib.tgpu_memcpy_d2h(self.h, C.c_void_p(offs.ctypes.data), C.c_void_p(d.offsets), (n + 1) * 4))
                    nbytes = max(nbytes, int(offs[n]))
                data = np.zeros(nbytes, dtype=np.uint8)
                h.offsets = offs.ctypes.data
                h.data = data.ctypes.data
                keep.append((d.type, data, valid, offs))
            elif d.type == abi.INT128:
                data = np.empty((n, 2), dtype=np.int64)
                h.data = data.ctypes.data
                keep.append((d.type, data, valid, None))
            else:
                data = np.empty(n, dtype=_NP_OF_TYPE[d.type])
                h.data = data.ctypes.data
                keep.append((d.type, data, valid, None))
        hp = abi.Page(dp.num_columns, 0, n, C.cast(host_cols, C.POINTER(abi.Column)))
        self.check(self.lib.tgpu_page_copy_to_host(self.h, pp, C.byref(hp)))
        blocks = []
        for type_, data, valid, offs in keep:
            if type_ == "view":
                blocks.append(data)
                continue
            bits = np.unpackbits(valid, bitorder="little")[:n].astype(np.bool_)
            nulls = ~bits
            blocks.append(Block(type_, data, nulls if nulls.any() else None, offs))
        if release:
            self.lib.tgpu_page_release(self.h, pp)
        return Page(*blocks, position_count=n)


class DeviceColumn:
    """A column that already lives in HBM (bench / GPU->GPU chaining)."""

    def __init__(self, type_, ptr, length, validity=None, offsets=None):
        self.type, self.ptr, self.length, self.validity, self.offsets = type_, ptr, length, validity, offsets


class DevicePage:
    def __init__(self, columns, rows):
        self.columns = columns
        self.rows = rows
        self.abi_cols = (abi.Column * max(1, len(columns)))()
        for i, c in enumerate(columns):
            a = self.abi_cols[i]
            a.type, a.flags, a.length = c.type, 0, c.length
            a.data, a.offsets, a.validity = c.ptr, c.offsets, c.validity
        self.page = abi.Page(len(columns), abi.PAGE_DEVICE, rows, C.cast(self.abi_cols, C.POINTER(abi.Column)))

    def ref(self):
        return C.byref(self.page)


class DeviceOutputPage:
    """A library-owned output page left on the device."""

    def __init__(self, ctx, pp):
        self.ctx, self.pp = ctx, pp
        self.rows = pp.contents.num_rows
        self.num_columns = pp.contents.num_columns

    def column(self, c):
        d = self.pp.contents.columns[c]
        return DeviceColumn(d.type, d.data, d.length, d.validity, d.offsets)

    def as_device_page(self):
        return DevicePage([self.column(c) for c in range(self.num_columns)], self.rows)

    def to_host(self):
        return self.ctx.page_to_host(self.pp, release=False)

    def release(self):
        if self.pp:
            self.ctx.lib.tgpu_page_release(self.ctx.h, self.pp)
            self.pp = None


def _as_abi_page(page):
    if isinstance(page, (DevicePage, AbiPage)):
        return page
    if isinstance(page, DeviceOutputPage):
        return page.as_device_page()
    return AbiPage(page)


# =====================================================================================================
# Operator / OperatorFactory
# =====================================================================================================
class Operator:
    """M/operator/Operator.java:21-102"""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle

    def needs_input(self):
        v = C.c_int()
        self.ctx.check(self.ctx.lib.tgpu_op_needs_input(self.h, C.byref(v)))
        return bool(v.value)

    def add_input(self, page):
        ap = _as_abi_page(page)
        self._last_input = page
        self.ctx.check(self.ctx.lib.tgpu_op_add_input(self.h, ap.ref()))

    def set_passthrough_by_reference(self, enable=True):
        """LookupJoinOperator only: host probe pages upload their join key alone; 1:1 outputs return the input blocks as views"""
        self.ctx.check(self.ctx.lib.tgpu_join_probe_set_passthrough_by_reference(self.h, int(enable)))
        self._by_reference = bool(enable)

    def get_output_device(self):
        pp = abi.PP()
        self.ctx.check(self.ctx.lib.tgpu_op_get_output(self.h, C.byref(pp)))
        return DeviceOutputPage(self.ctx, pp) if pp else None

    def get_output(self):
        pp = abi.PP()
        self.ctx.check(self.ctx.lib.tgpu_op_get_output(self.h, C.byref(pp)))
        views = getattr(self, "_last_input", None) if getattr(self, "_by_reference", False) else None
        if views is not None and not hasattr(views, "get_block"):
            views = None      # device-resident input pages are never by-reference
        return self.ctx.page_to_host(pp, views_of=views) if pp else None

    def finish(self):
        self.ctx.check(self.ctx.lib.tgpu_op_finish(self.h))

    def is_finished(self):
        v = C.c_int()
        self.ctx.check(self.ctx.lib.tgpu_op_is_finished(self.h, C.byref(v)))
        return bool(v.value)

    def memory_bytes(self):
        return self.ctx.lib.tgpu_op_memory_bytes(self.h)

    def close(self):
        if self.h:
            self.ctx.lib.tgpu_op_close(self.h)
            self.h = None


class OperatorFactory:
    """M/operator/OperatorFactory.java:16-30"""

    def __init__(self):
        self.closed = False

    def create_operator(self):
        if self.closed:
            raise RuntimeError("Factory is already closed")   # checkState(!closed) in every reference factory
        return self._create()

    def no_more_operators(self):
        self.closed = True

    def duplicate(self):
        raise NotImplementedError


# ---- expressions ----------------------------------------------------------------------------------
class Col:
    def __init__(self, channel, vtype):
        self.channel, self.vtype = channel, vtype


class Const:
    def __init__(self, value, vtype):
        self.value, self.vtype = value, vtype


class Null:
    def __init__(self, vtype):
        self.vtype = vtype


class Call:
    """op: one of abi.EX_*; args: expressions.  vtype is the OPERAND type (result of comparisons is BOOLEAN)."""

    def __init__(self, op, *args, in_list=None):
        self.op, self.args, self.in_list = op, list(args), in_list
        a0 = args[0]
        self.operand_vtype = a0.vtype if not isinstance(a0, Call) else a0.result_vtype
        boolean_result = op in (abi.EX_EQ, abi.EX_NE, abi.EX_LT, abi.EX_LE, abi.EX_GT, abi.EX_GE, abi.EX_AND, abi.EX_OR, abi.EX_NOT,
                                abi.EX_IS_NULL, abi.EX_IS_NOT_NULL, abi.EX_BETWEEN, abi.EX_IN)
        if boolean_result:
            self.result_vtype = abi.V_BOOLEAN
        elif op == abi.EX_CAST_BIGINT_TO_DOUBLE:
            self.result_vtype = abi.V_DOUBLE
        elif op == abi.EX_CAST_DOUBLE_TO_BIGINT:
            self.result_vtype = abi.V_BIGINT
        else:
            self.result_vtype = self.operand_vtype

    @property
    def vtype(self):
        return self.result_vtype


class PageProcessorProgram:
    """Compiles expression trees (the RowExpressions of LocalExecutionPlanner.java:2111-2114) to the three-address
    tgpu_expr_program.  `projections`: ints pass a channel through, expressions are computed."""

    def __init__(self, filter_expr, projections):
        self.insns = []
        self.in_lists = []
        self.live = set()
        self.filter_temp = -1
        self.num_filter_insns = 0
        if filter_expr is not None:
            opnd = self._emit(filter_expr)
            if opnd[0] != abi.OPND_TEMP:
                t = self._alloc()
                self._push(abi.EX_MOV, abi.V_BOOLEAN, t, opnd)
                opnd = (abi.OPND_TEMP, t, 0)
            self.filter_temp = opnd[1]
            self.num_filter_insns = len(self.insns)
        self.projections = []
        for p in projections:
            if isinstance(p, int):
                self.projections.append((0, p, 0))
                continue
            vt = p.vtype
            opnd = self._emit(p)
            if opnd[0] != abi.OPND_TEMP or opnd[1] == self.filter_temp:
                t = self._alloc()
                self._push(abi.EX_MOV, vt, t, opnd)
                opnd = (abi.OPND_TEMP, t, 0)
            self.projections.append((1, opnd[1], vt))
        self._build()

    def _alloc(self):
        for t in range(8):
            if t not in self.live:
                self.live.add(t)
                return t
        raise ValueError("expression needs more than 8 temporaries")

    def _push(self, op, vtype, dst, a, b=None, c=None):
        self.insns.append((op, vtype, dst, a, b or (abi.OPND_NONE, 0, 0), c or (abi.OPND_NONE, 0, 0)))

    def _emit(self, e):
        if isinstance(e, Col):
            return (abi.OPND_COLUMN, e.channel, 0)
        if isinstance(e, Const):
            imm = abi.Imm()
            if e.vtype == abi.V_DOUBLE:
                imm.f64 = float(e.value)
            else:
                imm.i64 = int(e.value)
            return (abi.OPND_CONST, 0, imm.i64)
        if isinstance(e, Null):
            return (abi.OPND_NULL, 0, 0)
        ops = [self._emit(a) for a in e.args]
        b = None
        if e.op == abi.EX_IN:
            vals = []
            for v in e.in_list:
                imm = abi.Imm()
                if e.operand_vtype == abi.V_DOUBLE:
                    imm.f64 = float(v)
                else:
                    imm.i64 = int(v)
                vals.append(imm.i64)
            self.in_lists.append(vals)
            b = (abi.OPND_CONST, 0, len(self.in_lists) - 1)
        for o in ops:   # operand temps die here (projection temps are never passed as operands twice)
            if o[0] == abi.OPND_TEMP and o[1] != self.filter_temp:
                self.live.discard(o[1])
        dst = self._alloc()
        self._push(e.op, e.operand_vtype, dst, ops[0], b if b else (ops[1] if len(ops) > 1 else None), ops[2] if len(ops) > 2 else None)
        return (abi.OPND_TEMP, dst, 0)

    def _build(self):
        n = len(self.insns)
        self._insns = (abi.ExprInsn * max(1, n))()
        for i, (op, vt, dst, a, b, c) in enumerate(self.insns):
            ins = self._insns[i]
            ins.op, ins.vtype, ins.dst = op, vt, dst
            for fld, o in (("a", a), ("b", b), ("c", c)):
                f = getattr(ins, fld)
                f.kind, f.index = o[0], o[1]
                f.imm.i64 = o[2]
        self._projs = (abi.Projection * max(1, len(self.projections)))()
        for i, (k, idx, vt) in enumerate(self.projections):
            self._projs[i].kind, self._projs[i].index, self._projs[i].vtype = k, idx, vt
        self._lists = (abi.InList * max(1, len(self.in_lists)))()
        self._list_bufs = []
        for i, vals in enumerate(self.in_lists):
            buf = (C.c_int64 * max(1, len(vals)))(*vals)
            self._list_bufs.append(buf)
            self._lists[i].count = len(vals)
            self._lists[i].values = C.cast(buf, C.POINTER(C.c_int64))
        self.struct = abi.ExprProgram(n, C.cast(self._insns, C.POINTER(abi.ExprInsn)), self.filter_temp, self.num_filter_insns,
                                      len(self.projections), C.cast(self._projs, C.POINTER(abi.Projection)),
                                      len(self.in_lists), C.cast(self._lists, C.POINTER(abi.InList)))


class FilterAndProjectOperatorFactory(OperatorFactory):
    def __init__(self, ctx, program):
        super().__init__()
        self.ctx, self.program = ctx, program

    def _create(self):
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_filter_project_create(self.ctx.h, C.byref(self.program.struct), C.byref(h)))
        return Operator(self.ctx, h)

    def duplicate(self):
        return FilterAndProjectOperatorFactory(self.ctx, self.program)


# ---- aggregation ------------------------------------------------------------------------------------
class Aggregator:
    """One AggregatorFactory (M/operator/aggregation/AggregatorFactory.java:40-58): function + input/mask channels."""

    def __init__(self, function, input_channel=-1, mask_channel=-1, result_type=0):
        """result_type: tgpu_type of an avg(decimal) result (INT64 short / INT128 long decimal) where the input does not tell (FINAL step)"""
        self.function, self.input_channel, self.mask_channel, self.result_type = function, input_channel, mask_channel, result_type


class HashAggregationOperator(Operator):
    # ROW-typed intermediate states (AccumulatorCompiler.java:687-760): set by the factory when the neighbouring stage is a Java
    # operator that speaks the reference's state types.  in_first: flat channel of each original input channel; out_widths: see
    # page.compose_row_blocks
    _out_widths = None
    _in_decimal = None            # {input channel (the plan's numbering): "decimal_sum" | "decimal_avg"}: VARBINARY decimal states to unpack

    def add_input(self, page):
        if hasattr(page, "blocks") and (self._in_decimal or any(isinstance(b, RowBlock) for b in page.blocks)):
            page = flatten_state_blocks(page, self._in_decimal or {})
        super().add_input(page)

    def get_output(self):
        out = super().get_output()
        if out is not None and self._out_widths is not None:
            out = compose_state_blocks(out, self._out_widths)
        return out

    def group_count(self):
        v = C.c_int64()
        self.ctx.check(self.ctx.lib.tgpu_agg_group_count(self.h, C.byref(v)))
        return v.value

    def rows_with_partial_aggregation_disabled(self):
        """AggregationMetrics.INPUT_ROWS_WITH_PARTIAL_AGGREGATION_DISABLED_METRIC_NAME"""
        v = C.c_int64()
        self.ctx.check(self.ctx.lib.tgpu_agg_rows_with_partial_aggregation_disabled(self.h, C.byref(v)))
        return v.value


class PartialAggregationController:
    """M/operator/aggregation/partial/PartialAggregationController.java:35-103 - the object lives in the native library (the operators
    report their flushes to it themselves); needs no GPU."""

    def __init__(self, lib, max_partial_memory, unique_rows_ratio_threshold):
        self.lib, self.max_partial_memory, self.threshold = lib, int(max_partial_memory), float(unique_rows_ratio_threshold)
        h = C.c_void_p()
        rc = lib.tgpu_partial_agg_controller_create(self.max_partial_memory, self.threshold, C.byref(h))
        if rc != 0:
            raise abi.TrinoGpuError(rc, "INVALID_ARGUMENT", "tgpu_partial_agg_controller_create failed")
        self.h = h

    def is_partial_aggregation_disabled(self):
        return bool(self.lib.tgpu_partial_agg_controller_is_disabled(self.h))

    def on_flush(self, bytes_processed, rows_processed, unique_rows_produced=None):
        """unique_rows_produced None = OptionalLong.empty()"""
        self.lib.tgpu_partial_agg_controller_on_flush(self.h, int(bytes_processed), int(rows_processed),
                                                      -1 if unique_rows_produced is None else int(unique_rows_produced))

    def duplicate(self):
        return PartialAggregationController(self.lib, self.max_partial_memory, self.threshold)

    def close(self):
        if self.h:
            self.lib.tgpu_partial_agg_controller_destroy(self.h)
            self.h = None


class HashAggregationOperatorFactory(OperatorFactory):
    def __init__(self, ctx, group_by_channels, step, aggregators, expected_groups=10_000, max_partial_memory=0, pre=None,
                 global_aggregation_group_ids=(), group_id_channel=None, input_types=None, partial_aggregation_controller=None,
                 row_typed_states=False):
        """global_aggregation_group_ids / group_id_channel (a group-by CHANNEL, like the reference's groupIdChannel) / input_types (tgpu_type
        per input channel): the default rows of global grouping sets over empty input (HashAggregationOperator.java:537-567)"""
        super().__init__()
        self.ctx, self.group_by_channels, self.step, self.aggregators = ctx, list(group_by_channels), step, list(aggregators)
        self.expected_groups, self.max_partial_memory, self.pre = expected_groups, max_partial_memory, pre
        self.global_ids, self.group_id_channel, self.input_types = list(global_aggregation_group_ids), group_id_channel, input_types
        self.controller = partial_aggregation_controller
        self.row_typed_states = row_typed_states

    _FLAT = {abi.AGG_AVG: 2, abi.AGG_SUM_DECIMAL: 2, abi.AGG_AVG_DECIMAL: 3}           # flat state columns per function (default 1)
    _DECIMAL = {abi.AGG_SUM_DECIMAL: "decimal_sum", abi.AGG_AVG_DECIMAL: "decimal_avg"}

    def _state_widths(self):
        """the reference's state type per aggregate: ROW(BIGINT, DOUBLE) for avg (2 flat columns), VARBINARY for the decimal states"""
        return [self._DECIMAL.get(a.function, self._FLAT.get(a.function, 1)) for a in self.aggregators]

    def _create(self):
        from_state = self.step in (abi.STEP_FINAL, abi.STEP_INTERMEDIATE)
        to_state = self.step in (abi.STEP_PARTIAL, abi.STEP_INTERMEDIATE)
        keys_list, agg_inputs = list(self.group_by_channels), [a.input_channel for a in self.aggregators]
        if self.row_typed_states and from_state:
            # channels are numbered as the Java plan numbers them (one channel per ROW state); the library sees the flattened page
            wide = {a.input_channel: self._FLAT[a.function] for a in self.aggregators if a.function in self._FLAT}
            top = max(keys_list + agg_inputs + [0])
            first, at = [], 0
            for c in range(top + 1):
                first.append(at)
                at += wide.get(c, 1)
            keys_list = [first[c] for c in keys_list]
            agg_inputs = [first[c] if c >= 0 else c for c in agg_inputs]
        keys = _i32(keys_list)
        fns = (abi.AggFn * max(1, len(self.aggregators)))()
        for i, a in enumerate(self.aggregators):
            fns[i].function, fns[i].input_channel, fns[i].mask_channel = a.function, agg_inputs[i], a.mask_channel
            fns[i].reserved = getattr(a, "result_type", 0)
        gids = _i32(self.global_ids)
        types = _i32(list(self.input_types or []))
        spec = abi.AggSpec(len(self.group_by_channels), C.cast(keys, C.POINTER(C.c_int32)), self.step, len(self.aggregators),
                           C.cast(fns, C.POINTER(abi.AggFn)), self.expected_groups, self.max_partial_memory,
                           C.pointer(self.pre.struct) if self.pre is not None else None,
                           len(self.global_ids), C.cast(gids, C.POINTER(C.c_int32)),
                           self.group_by_channels.index(self.group_id_channel) if self.group_id_channel is not None else -1,
                           len(self.input_types or []), C.cast(types, C.POINTER(C.c_int32)),
                           self.controller.h if self.controller is not None else None)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_agg_create(self.ctx.h, C.byref(spec), C.byref(h)))
        op = HashAggregationOperator(self.ctx, h)
        if self.row_typed_states and to_state:
            op._out_widths = [1] * len(self.group_by_channels) + self._state_widths()
        if self.row_typed_states and from_state:
            op._in_decimal = {a.input_channel: self._DECIMAL[a.function] for a in self.aggregators if a.function in self._DECIMAL}
        return op

    def duplicate(self):
        return HashAggregationOperatorFactory(self.ctx, self.group_by_channels, self.step, self.aggregators, self.expected_groups,
                                              self.max_partial_memory, self.pre, self.global_ids, self.group_id_channel, self.input_types,
                                              # HashAggregationOperatorFactory.duplicate :238: a fresh controller for the duplicated plan node
                                              self.controller.duplicate() if self.controller is not None else None, self.row_typed_states)


class GroupByHash:
    """M/operator/GroupByHash.java: getGroupIds / getGroupCount over a persistent device table."""

    def __init__(self, ctx, key_channels, expected_size=10_000):
        self.ctx = ctx
        keys = _i32(list(key_channels))
        h = C.c_void_p()
        ctx.check(ctx.lib.tgpu_groupby_hash_create(ctx.h, len(key_channels), C.cast(keys, C.POINTER(C.c_int32)), expected_size, C.byref(h)))
        self.h = h

    def get_group_ids(self, page):
        ap = _as_abi_page(page)
        out = np.empty(page.position_count, dtype=np.int32)
        self.ctx.check(self.ctx.lib.tgpu_groupby_hash_get_group_ids(self.h, ap.ref(), C.c_void_p(out.ctypes.data)))
        return out

    def get_group_count(self):
        v = C.c_int64()
        self.ctx.check(self.ctx.lib.tgpu_agg_group_count(self.h, C.byref(v)))
        return v.value

    def close(self):
        if self.h:
            self.ctx.lib.tgpu_op_close(self.h)
            self.h = None


# ---- join -------------------------------------------------------------------------------------------
class LookupSource:
    """M/operator/join/LookupSource.java:24-68 (device table handle)."""

    def __init__(self, ctx, handle):
        self.ctx, self.h = ctx, handle

    def get_join_position_count(self):
        return self.ctx.lib.tgpu_lookup_position_count(self.h)

    def get_in_memory_size_in_bytes(self):
        return self.ctx.lib.tgpu_lookup_memory_bytes(self.h)

    def has_position_links(self):
        return bool(self.ctx.lib.tgpu_lookup_has_duplicates(self.h))

    def get_join_positions(self, keys_page):
        ap = _as_abi_page(keys_page)
        n = keys_page.position_count
        out = np.empty(n, dtype=np.int32)
        self.ctx.check(self.ctx.lib.tgpu_lookup_get_join_positions(self.ctx.h, self.h, ap.ref(), C.c_void_p(out.ctypes.data)))
        return out

    def get_join_positions_device(self, device_page, out_ptr):
        self.ctx.check(self.ctx.lib.tgpu_lookup_get_join_positions(self.ctx.h, self.h, device_page.ref(), C.c_void_p(out_ptr)))

    def key_domain(self, max_values):
        """DynamicFilterSourceOperator / JoinDomainBuilder: (min, max, distinct count, sorted values or None, has_null) of the
        build-side join key; values are None when there are more than max_values distinct keys (range fallback)."""
        lo, hi, cnt, has_null = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int32()
        vals = np.empty(max(max_values, 1), dtype=np.int64)
        self.ctx.check(self.ctx.lib.tgpu_lookup_key_domain(self.ctx.h, self.h, max_values, C.byref(lo), C.byref(hi), C.byref(cnt),
                                                            vals.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(has_null)))
        values = vals[:cnt.value].copy() if cnt.value <= max_values else None
        return lo.value, hi.value, cnt.value, values, bool(has_null.value)

    def position_links(self):
        out = np.empty(self.get_join_position_count(), dtype=np.int32)
        self.ctx.check(self.ctx.lib.tgpu_lookup_copy_position_links(self.ctx.h, self.h, C.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if self.h:
            self.ctx.lib.tgpu_lookup_release(self.h)
            self.h = None


class JoinBridge:
    """The JoinBridgeManager / PartitionedLookupSourceFactory hand-off (PartitionedLookupSourceFactory.java:100,126):
    the build operator lends its LookupSource, probe operators take it once it is there."""

    def __init__(self):
        self.lookup_source = None

    def lend(self, lookup_source):
        self.lookup_source = lookup_source

    def is_built(self):
        return self.lookup_source is not None


class HashBuilderOperator(Operator):
    def __init__(self, ctx, handle, bridge):
        super().__init__(ctx, handle)
        self.bridge = bridge

    def finish(self):
        super().finish()
        if not self.bridge.is_built():
            lk = C.c_void_p()
            self.ctx.check(self.ctx.lib.tgpu_join_build_get_lookup(self.h, C.byref(lk)))
            self.bridge.lend(LookupSource(self.ctx, lk))


class HashBuilderOperatorFactory(OperatorFactory):
    def __init__(self, ctx, bridge, hash_channels, output_channels, expected_positions=10_000):
        super().__init__()
        self.ctx, self.bridge = ctx, bridge
        self.hash_channels, self.output_channels, self.expected_positions = list(hash_channels), list(output_channels), expected_positions

    def _create(self):
        kc, oc = _i32(self.hash_channels), _i32(self.output_channels)
        spec = abi.JoinBuildSpec(len(self.hash_channels), C.cast(kc, C.POINTER(C.c_int32)), len(self.output_channels),
                                 C.cast(oc, C.POINTER(C.c_int32)), self.expected_positions)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_join_build_create(self.ctx.h, C.byref(spec), C.byref(h)))
        return HashBuilderOperator(self.ctx, h, self.bridge)

    def duplicate(self):
        raise RuntimeError("Parallel hash build can not be duplicated")   # HashBuilderOperator.java:131-134


class LookupJoinOperatorFactory(OperatorFactory):
    def __init__(self, ctx, bridge, join_type, output_single_match, probe_join_channels, probe_output_channels):
        super().__init__()
        self.ctx, self.bridge, self.join_type, self.output_single_match = ctx, bridge, join_type, output_single_match
        self.probe_join_channels, self.probe_output_channels = list(probe_join_channels), list(probe_output_channels)

    def _create(self):
        if not self.bridge.is_built():
            # PageJoiner.process blocks on lookupSourceFuture (PageJoiner.java:102-105); here the driver must build first
            raise RuntimeError("lookup source is not built yet")
        kc, oc = _i32(self.probe_join_channels), _i32(self.probe_output_channels)
        spec = abi.JoinProbeSpec(self.join_type, int(self.output_single_match), len(self.probe_join_channels), C.cast(kc, C.POINTER(C.c_int32)),
                                 len(self.probe_output_channels), C.cast(oc, C.POINTER(C.c_int32)))
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_join_probe_create(self.ctx.h, C.byref(spec), self.bridge.lookup_source.h, C.byref(h)))
        return Operator(self.ctx, h)

    def duplicate(self):
        return LookupJoinOperatorFactory(self.ctx, self.bridge, self.join_type, self.output_single_match, self.probe_join_channels,
                                         self.probe_output_channels)


class LookupOuterOperatorFactory(OperatorFactory):
    """M/operator/join/LookupOuterOperator.java:38-95.  `probe_output_types`: tgpu types of the probe output channels."""

    def __init__(self, ctx, bridge, probe_output_types):
        super().__init__()
        self.ctx, self.bridge, self.probe_output_types = ctx, bridge, list(probe_output_types)

    def _create(self):
        if not self.bridge.is_built():
            raise RuntimeError("lookup source is not built yet")
        t = _i32(self.probe_output_types)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_join_outer_create(self.ctx.h, self.bridge.lookup_source.h, C.cast(t, C.POINTER(C.c_int32)), len(self.probe_output_types),
                                                            C.byref(h)))
        return Operator(self.ctx, h)


class SetBuilderOperatorFactory(OperatorFactory):
    """M/operator/SetBuilderOperator.java: the ChannelSet of a semi-join is a lookup source over one channel without outputs."""

    def __init__(self, ctx, bridge, set_channel, expected_positions=0):
        super().__init__()
        self.inner = HashBuilderOperatorFactory(ctx, bridge, [set_channel], [], expected_positions)

    def _create(self):
        return self.inner.create_operator()


class HashSemiJoinOperatorFactory(OperatorFactory):
    """M/operator/HashSemiJoinOperator.java:43-100"""

    def __init__(self, ctx, bridge, probe_join_channel):
        super().__init__()
        self.ctx, self.bridge, self.probe_join_channel = ctx, bridge, probe_join_channel

    def _create(self):
        if not self.bridge.is_built():
            raise RuntimeError("channel set is not built yet")
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_semi_join_create(self.ctx.h, self.bridge.lookup_source.h, self.probe_join_channel, C.byref(h)))
        return Operator(self.ctx, h)


# ---- partitioned output -----------------------------------------------------------------------------
class ColumnDomain:
    """One column's Domain of a dynamic filter's TupleDomain (S/predicate/Domain.java): nullAllowed + a value set.
    Constructors mirror the reference's factories: all / none / only_null / single_value / multiple_values / range (end exclusive,
    like Range.range(type, lo, true, hi, false) in the reference's tests)."""

    def __init__(self, channel, kind, null_allowed=False, lo=0, hi=0, values=None):
        self.channel, self.kind, self.null_allowed, self.lo, self.hi, self.values = channel, kind, null_allowed, lo, hi, values

    @staticmethod
    def all(channel):
        return ColumnDomain(channel, abi.DOMAIN_ALL, True)

    @staticmethod
    def none(channel):
        return ColumnDomain(channel, abi.DOMAIN_NONE, False)

    @staticmethod
    def only_null(channel):
        return ColumnDomain(channel, abi.DOMAIN_NONE, True)

    @staticmethod
    def single_value(channel, value, null_allowed=False):
        return ColumnDomain(channel, abi.DOMAIN_DISCRETE, null_allowed, values=[value])

    @staticmethod
    def multiple_values(channel, values, null_allowed=False):
        return ColumnDomain(channel, abi.DOMAIN_DISCRETE, null_allowed, values=list(values))

    @staticmethod
    def range(channel, lo, hi_exclusive, null_allowed=False):
        return ColumnDomain(channel, abi.DOMAIN_RANGE, null_allowed, lo=lo, hi=hi_exclusive - 1)


class DynamicFilterOperator(Operator):
    def update(self, domains):
        arr, keep = _domains(domains)
        self.ctx.check(self.ctx.lib.tgpu_dynamic_filter_update(self.h, C.cast(arr, C.c_void_p), len(domains)))

    def is_effective(self, index):
        v = C.c_int32()
        self.ctx.check(self.ctx.lib.tgpu_dynamic_filter_is_effective(self.h, index, C.byref(v)))
        return bool(v.value)


def _domains(domains):
    arr = (abi.Domain * max(1, len(domains)))()
    keep = []
    for i, d in enumerate(domains):
        arr[i].channel, arr[i].null_allowed, arr[i].kind = d.channel, int(d.null_allowed), d.kind
        arr[i].min, arr[i].max = d.lo, d.hi
        if d.values is not None:
            vals = (C.c_int64 * len(d.values))(*d.values)
            keep.append(vals)
            arr[i].num_values = len(d.values)
            arr[i].values = C.cast(vals, C.POINTER(C.c_int64))
    return arr, keep


class DynamicFilterOperatorFactory(OperatorFactory):
    """DynamicPageFilter (M/sql/gen/columnar/DynamicPageFilter.java:47-211) as an operator in front of the probe: drops the rows the
    build side's key domain rules out.  `domains`: ColumnDomains in evaluation order; none = TupleDomain.all()."""

    def __init__(self, ctx, domains, selectivity_threshold=1.0):
        super().__init__()
        self.ctx, self.domains, self.threshold = ctx, list(domains), selectivity_threshold

    def _create(self):
        arr, keep = _domains(self.domains)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_dynamic_filter_create(self.ctx.h, C.cast(arr, C.c_void_p), len(self.domains), self.threshold, C.byref(h)))
        return DynamicFilterOperator(self.ctx, h)

    def duplicate(self):
        return DynamicFilterOperatorFactory(self.ctx, self.domains, self.threshold)


class PartitionedOutputOperator(Operator):
    def get_output_with_partition(self):
        """(partition, Page) or None — the OutputBuffer.enqueue(partition, pages) call of PagePartitioner.java:484-487"""
        page = self.get_output()
        if page is None:
            return None
        p = C.c_int32()
        self.ctx.check(self.ctx.lib.tgpu_partition_last_output_partition(self.h, C.byref(p)))
        return p.value, page

    def get_partitions(self, page):
        ap = _as_abi_page(page)
        out = np.empty(page.position_count, dtype=np.int32)
        self.ctx.check(self.ctx.lib.tgpu_partition_get_partitions(self.h, ap.ref(), C.c_void_p(out.ctypes.data)))
        return out


class PartitionedOutputOperatorFactory(OperatorFactory):
    """PartitionedOutputFactory.createOutputOperator (M/operator/output/PartitionedOutputOperator.java:69-118).
    `partition_constants`: one entry per partition channel, None or a one-position Block - read where the channel is negative
    (PagePartitioner.java:78-101).  `partition_function`: abi.PARTITION_HASH_BUCKET, or abi.PARTITION_LOCAL for the
    LocalPartitionGenerator of the local exchange (M/operator/exchange/LocalPartitionGenerator.java:45-77)."""

    def __init__(self, ctx, partition_channels, bucket_count, bucket_to_partition=None, null_channel=-1, replicates_any_row=False,
                 partition_constants=None, partition_function=abi.PARTITION_HASH_BUCKET):
        super().__init__()
        self.ctx, self.partition_channels, self.bucket_count = ctx, list(partition_channels), bucket_count
        self.bucket_to_partition, self.null_channel, self.replicates_any_row = bucket_to_partition, null_channel, replicates_any_row
        self.partition_constants, self.partition_function = partition_constants, partition_function

    def _create(self):
        kc = _i32(self.partition_channels)
        b2p = _i32(list(self.bucket_to_partition)) if self.bucket_to_partition is not None else None
        constants = None
        if self.partition_constants is not None:
            blocks = [b if b is not None else Block.bigint(np.zeros(1, dtype=np.int64)) for b in self.partition_constants]
            constants = AbiPage(Page(*blocks))      # alive until the create call returns: the library keeps only the hashes
        spec = abi.PartitionSpec(len(self.partition_channels), C.cast(kc, C.POINTER(C.c_int32)), self.bucket_count,
                                 C.cast(b2p, C.POINTER(C.c_int32)) if b2p is not None else None, self.null_channel, int(self.replicates_any_row),
                                 self.partition_function, C.cast(constants.columns, C.c_void_p) if constants is not None else None)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.tgpu_partition_create(self.ctx.h, C.byref(spec), C.byref(h)))
        return PartitionedOutputOperator(self.ctx, h)

    def duplicate(self):
        return PartitionedOutputOperatorFactory(self.ctx, self.partition_channels, self.bucket_count, self.bucket_to_partition, self.null_channel,
                                                self.replicates_any_row, self.partition_constants, self.partition_function)


class LocalPartitionGenerator:
    """M/operator/exchange/LocalPartitionGenerator.java:23-77 over the library's partitioner: getPartitions of a page's hash channels."""

    def __init__(self, ctx, hash_channels, partition_count):
        if partition_count < 1 or partition_count & (partition_count - 1):
            raise ValueError("partitionCount must be a power of 2")                 # LocalPartitionGenerator.java:33
        self.partition_count = partition_count
        self._op = PartitionedOutputOperatorFactory(ctx, hash_channels, partition_count, partition_function=abi.PARTITION_LOCAL).create_operator()

    def get_partitions(self, page):
        return self._op.get_partitions(page)

    def close(self):
        self._op.close()


def drive(operator, pages):
    """The relevant slice of Driver.processInternal (M/operator/Driver.java:391-424) for one operator:
    feed pages while needsInput, drain getOutput, then finish and drain.  Returns the output pages."""
    out = []
    for p in pages:
        while not operator.needs_input():
            o = operator.get_output()
            if o is not None:
                out.append(o)
        operator.add_input(p)
        while True:
            o = operator.get_output()
            if o is None:
                break
            out.append(o)
    operator.finish()
    while not operator.is_finished():
        o = operator.get_output()
        if o is not None:
            out.append(o)
        elif operator.is_finished():
            break
    return out
