#!/bin/bash
set -u
mkdir -p gpurun_out
run() {
  name=$1; shift
  extra=""
  if [[ "$name" == *shuf* ]]; then extra="--shuffle-probe"; fi
  env "$@" timeout 300 python bench.py --no-cpu-baseline --e2e-rows 1000 --steps 3 --warmup 2 --q1-sf 0 $extra > gpurun_out/exp2_$name.log 2>&1
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/exp2_{name}.log").read().strip().splitlines()[-1])
    print(name, "op Grows/s", round(d["value"] / 1e9, 1), "step ms", round(d["ms_per_step"], 2), "index kernel ms", round(d["roofline"]["kernel_ms"], 3), "frac", round(d["roofline"]["frac"], 3), "build s", round(d["build_seconds"], 3))
except Exception as e:
    print(name, "bench failed", e, open(f"gpurun_out/exp2_{name}.log").read()[-500:])
PY
}
run default X=1
run rows8 TGPU_JOIN_ROWS=8

run murmur TGPU_JOIN_HASH=0

run default_shuf X=1
run murmur_shuf TGPU_JOIN_HASH=0
