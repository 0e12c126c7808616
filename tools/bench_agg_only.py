import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from trino_b200 import abi, operators as ops
ctx = ops.Context(0); lib = ctx.lib
m = int(float(sys.argv[1])) if len(sys.argv) > 1 else 150_000_000
groups = int(float(sys.argv[2])) if len(sys.argv) > 2 else 10_000_000
d_keys = ctx.malloc(m * 8)
for lo in range(0, m, groups):
    cnt = min(groups, m - lo)
    ctx.check(lib.tgpu_synth_orders_keys(ctx.h, groups, 0, cnt, 0x77 + lo, 1, C.c_void_p(d_keys + lo * 8)))
d_val = ctx.malloc(m * 8)
ctx.check(lib.tgpu_synth_lineitem_keys(ctx.h, m, 0, m, 1, 0, C.c_void_p(d_val)))
gpage = ops.DevicePage([ops.DeviceColumn(abi.INT64, d_keys, m), ops.DeviceColumn(abi.INT64, d_val, m)], m)
f = ops.HashAggregationOperatorFactory(ctx, [0], abi.STEP_SINGLE, [ops.Aggregator(abi.AGG_SUM, 1), ops.Aggregator(abi.AGG_COUNT_STAR)], expected_groups=groups)
for i in range(3):
    op = f.create_operator()
    ctx.timer_start()
    op.add_input(gpage)
    ms = ctx.timer_stop_ms()
    print("run", i, "ms", ms, "groups", op.group_count())
    op.close()
