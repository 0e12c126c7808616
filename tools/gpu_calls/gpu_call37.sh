set -x
O=gpurun_out/r3l
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_join.py -m gpu -q -x --timeout 500 > $O/pytest_join.log 2>&1; tail -4 $O/pytest_join.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-groupby-bigint --q1-sf 0 --no-secondary"
for shape in 28 26 45 44 18; do
  TGPU_JOIN_WIDE_SHAPE=$shape timeout 300 $B > $O/bench_wide_$shape.json 2> $O/bench_wide_$shape.err
  python - <<PY
import json
try:
    d = json.loads(open("$O/bench_wide_$shape.json").read().strip().splitlines()[-1])
    print("wide $shape: step ms", d["ms_per_step"], "value", d["value"], "kernel frac", d["roofline"]["frac"], "shuffled", d.get("roofline_shuffled", {}).get("ms_per_step"), d.get("roofline_shuffled", {}).get("rows_per_sec"))
except Exception as e:
    print("wide $shape failed", e)
PY
done
TGPU_JOIN_NO_WIDE=1 timeout 300 $B > $O/bench_narrow.json 2> $O/bench_narrow.err
python - <<PY
import json
d = json.loads(open("$O/bench_narrow.json").read().strip().splitlines()[-1])
print("narrow: step ms", d["ms_per_step"], "value", d["value"], "shuffled", d.get("roofline_shuffled", {}).get("ms_per_step"), d.get("roofline_shuffled", {}).get("rows_per_sec"))
PY
