set -x
O=gpurun_out/r3n
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_partition.py -m gpu -q -x --timeout 500 > $O/pytest_partition.log 2>&1; tail -15 $O/pytest_partition.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:join_probe_wide -s 1 -c 1 -o $O/prof_wide python bench.py --shuffle-probe --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-groupby-bigint --q1-sf 0 --no-secondary > $O/ncu_wide.log 2>&1; tail -3 $O/ncu_wide.log
python tools/bench_ops.py > $O/bench_ops.log 2>&1; tail -12 $O/bench_ops.log
