set -x
O=gpurun_out/r3h
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_filter_project.py tests/test_gpu_dynamic_filter.py tests/test_gpu_decimal.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -5 $O/pytest.log
python tools/bench_ops.py 3e8 > $O/ops.log 2>&1; head -1 $O/ops.log | cut -c1-260
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ops.csv python tools/bench_ops.py 3e8 > $O/launches_ops.log 2>&1
