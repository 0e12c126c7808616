set -x
O=gpurun_out/r2x
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_workloads.py tests/test_gpu_groupby.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python tools/bench_ops.py 3e8 > $O/ops_lean3.log 2>&1; tail -1 $O/ops_lean3.log
TGPU_XCHG_LEAN_MINB4=1 python tools/bench_ops.py 3e8 > $O/ops_lean4.log 2>&1; tail -1 $O/ops_lean4.log
TGPU_XCHG_NO_LEAN=1 python tools/bench_ops.py 3e8 > $O/ops_nolean.log 2>&1; tail -1 $O/ops_nolean.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ops.csv python tools/bench_ops.py 3e8 > $O/launches_ops.log 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; python -c "import json;d=json.load(open('$O/bench_ref.json'));print(d['value'], d['cpu_baseline']['spread'])"
