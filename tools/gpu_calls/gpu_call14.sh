set -x
O=gpurun_out/r2n
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_partial_aggregation.py -m gpu -q --timeout 200 2>&1 | tail -3
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-shuffled --no-groupby-bigint --q1-sf 0"
# launch list of the default bench command (the share of the step each kernel takes)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
# full captures of the dominant kernels
ncu --set full --clock-control none --import-source on -k regex:join_probe_lean -s 3 -c 1 -o $O/prof_probe $B > $O/ncu_probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_agg_small_jit -s 1 -c 1 -o $O/prof_q1 python tools/bench_q1_only.py 300 > $O/ncu_q1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xchg_ -s 6 -c 3 -o $O/prof_xchg python tools/bench_ops.py 3e8 > $O/ncu_xchg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_fp_ -s 4 -c 2 -o $O/prof_fp python tools/bench_ops.py 3e8 > $O/ncu_fp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ops.csv python tools/bench_ops.py 3e8 > $O/launches_ops.log 2>&1
ls -la $O
