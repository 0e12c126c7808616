set -x
mkdir -p gpurun_out/r2h
python -m pytest tests/test_gpu_groupby.py tests/test_gpu_harness.py -m gpu -q --timeout 900 > gpurun_out/r2h/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h/pytest.log
grep -v "^$" gpurun_out/r2h/pytest.log | tail -30 | cut -c1-250
python tools/bench_agg_only.py > gpurun_out/r2h/agg_only.log 2>&1; cat gpurun_out/r2h/agg_only.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h/agg_launches.csv python tools/bench_agg_only.py 150000000 10000000 > gpurun_out/r2h/agg_ncu.log 2>&1
python bench.py --no-cpu-baseline --no-e2e --no-shuffled --no-groupby-bigint > gpurun_out/r2h/bench_q1.json 2> gpurun_out/r2h/bench_q1.err; tail -2 gpurun_out/r2h/bench_q1.err
tests/harness/driver_loop q1 600000000 4 > gpurun_out/r2h/harness_q1.json 2>&1; cat gpurun_out/r2h/harness_q1.json
tests/harness/driver_loop join 60000000 4 > gpurun_out/r2h/harness_join.json 2>&1; cat gpurun_out/r2h/harness_join.json
