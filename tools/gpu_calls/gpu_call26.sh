set -x
O=gpurun_out/r2z
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_decimal.py tests/test_gpu_groupby.py tests/test_gpu_partial_aggregation.py tests/test_gpu_join.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -30 $O/pytest.log
python tools/bench_agg_only.py 150000000 10000000 > $O/agg.log 2>&1; tail -1 $O/agg.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_agg.csv python tools/bench_agg_only.py 150000000 10000000 > $O/launches_agg.log 2>&1
