set -x
O=gpurun_out/r3o
mkdir -p $O
B="python bench.py --shuffle-probe --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-groupby-bigint --q1-sf 0 --no-secondary"
run() {   # name, env..., -- extra bench args
  name=$1; shift
  env "$@" timeout 300 $B $EXTRA > $O/$name.json 2> $O/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/$name.json").read().strip().splitlines()[-1])
    print("$name: step ms %.3f  G rows/s %.2f  l2g %s  index-probe ms %.3f" % (d["ms_per_step"], d["value"] / 1e9, d.get("l2_fetch_granularity"), d["roofline_index_probe"]["kernel_ms"]))
except Exception as e:
    print("$name failed", e)
PY
}
EXTRA="" run ld0 TGPU_JOIN_WIDE_LOAD=0
EXTRA="" run ld3 TGPU_JOIN_WIDE_LOAD=3
EXTRA="" run ld5 TGPU_JOIN_WIDE_LOAD=5
EXTRA="--l2-fetch 32" run ld0_l2f32 TGPU_JOIN_WIDE_LOAD=0
EXTRA="--l2-fetch 32" run ld3_l2f32 TGPU_JOIN_WIDE_LOAD=3
EXTRA="--l2-fetch 128" run ld0_l2f128 TGPU_JOIN_WIDE_LOAD=0
EXTRA="--l2-fetch 32" run narrow_l2f32 TGPU_JOIN_WIDE=never
EXTRA="" run narrow TGPU_JOIN_WIDE=never
