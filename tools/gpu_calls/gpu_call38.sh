set -x
O=gpurun_out/r3m
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_join.py tests/test_gpu_workloads.py -m gpu -q -x --timeout 500 > $O/pytest_join.log 2>&1; tail -4 $O/pytest_join.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-groupby-bigint --q1-sf 0 --no-secondary > $O/bench_auto.json 2> $O/bench_auto.err
python - <<PY
import json
d = json.loads(open("$O/bench_auto.json").read().strip().splitlines()[-1])
sh = d.get("roofline_shuffled", {})
print("auto: step ms", d["ms_per_step"], "value", d["value"], "launches", d.get("gpu_launches"), "e2e", d.get("e2e", {}).get("value"), "shuffled", sh.get("ms_per_step"), sh.get("rows_per_sec"))
PY
