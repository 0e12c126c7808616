#!/bin/bash
set -u
run() {
  name=$1; sf=$2; shift; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-rows 1000 --steps 3 --warmup 2 --q1-sf 0 --sf $sf > gpurun_out/exp4_$name.log 2>&1
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/exp4_{name}.log").read().strip().splitlines()[-1])
    ip = d["roofline_index_probe"]
    print(name, "op Grows/s", round(d["value"] / 1e9, 1), "fused ms", round(d["roofline"]["kernel_ms"], 3), "index ms", round(ip["kernel_ms"], 3), "index Grows/s", round(ip["kernel_rows_per_sec"] / 1e9, 1))
except Exception as e:
    print(name, "bench failed", e, open(f"gpurun_out/exp4_{name}.log").read()[-300:])
PY
}

run sf10 10 X=1
run sf100 100 X=1
run sf100_generic 100 TGPU_JOIN_GENERIC_KERNELS=1

run sf100_murmur 100 TGPU_JOIN_HASH=0

