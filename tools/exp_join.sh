#!/bin/bash
# experiment matrix for the join probe layout (run under gpurun); results in gpurun_out/exp_join_*.{log,csv}
set -u
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-rows 1000 --steps 3 --warmup 2 --q1-sf 0 > gpurun_out/exp_join_$name.log 2>&1
  env "$@" timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:join_probe -c 6 --csv --log-file gpurun_out/exp_join_$name.csv python bench.py --no-cpu-baseline --e2e-rows 1000 --steps 1 --warmup 1 --q1-sf 0 > /dev/null 2>&1
  python - "$name" <<'PY'
import json, sys, csv, collections
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/exp_join_{name}.log").read().strip().splitlines()[-1])
    print(name, "op Grows/s", round(d["value"] / 1e9, 1), "step ms", round(d["ms_per_step"], 2), "index kernel ms", round(d["roofline"]["kernel_ms"], 3), "build s", round(d["build_seconds"], 3))
except Exception as e:
    print(name, "bench failed", e)
try:
    rows = list(csv.reader(open(f"gpurun_out/exp_join_{name}.csv")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        key = (r[h.index("Kernel Name")][:48], r[h.index("Metric Name")])
        agg.setdefault(key, []).append(float(r[h.index("Metric Value")].replace(",", "")))
    for (k, m), v in agg.items():
        print("   ", k, m, round(sum(v) / len(v) / (1e9 if "bytes" in m or "inst" in m else 1e6), 3), "G" if ("bytes" in m or "inst" in m) else "ms")
except Exception as e:
    print(name, "ncu parse failed", e)
PY
}
run murmur TGPU_JOIN_HASH=0
run line TGPU_JOIN_HASH=1
run line_cap2 TGPU_JOIN_HASH=1 TGPU_JOIN_CAP_SHIFT=1
run line_bucket TGPU_JOIN_HASH=1 TGPU_JOIN_BUCKET_PROBE=1
run line_bucket_cap2 TGPU_JOIN_HASH=1 TGPU_JOIN_BUCKET_PROBE=1 TGPU_JOIN_CAP_SHIFT=1
run line_slotpayload TGPU_JOIN_HASH=1 TGPU_JOIN_CAP_SHIFT=1 TGPU_JOIN_PAYLOAD_BY_SLOT=1
