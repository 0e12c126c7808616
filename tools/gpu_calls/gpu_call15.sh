set -x
O=gpurun_out/r2o
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_partial_aggregation.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python tools/bench_agg_only.py 150000000 10000000 > $O/agg.log 2>&1; tail -2 $O/agg.log
TGPU_AGG_G_SIZE_PCT=75 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_pct75.log 2>&1; tail -1 $O/agg_pct75.log
TGPU_AGG_SLICE_BYTES=33554432 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_slice32m.log 2>&1; tail -1 $O/agg_slice32m.log
TGPU_AGG_G_SIZE_PCT=75 TGPU_AGG_SLICE_BYTES=33554432 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_pct75_slice32m.log 2>&1; tail -1 $O/agg_pct75_slice32m.log
TGPU_AGG_STABLE_SCATTER=1 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_stable.log 2>&1; tail -1 $O/agg_stable.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_agg.csv python tools/bench_agg_only.py 150000000 10000000 > $O/launches_agg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_agg_general_jit -s 1 -c 1 -o $O/prof_gjit python tools/bench_agg_only.py 150000000 10000000 > $O/ncu_gjit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xchg_scatter_unordered -s 1 -c 1 -o $O/prof_unordered python tools/bench_agg_only.py 150000000 10000000 > $O/ncu_unordered.log 2>&1
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-shuffled --no-groupby-bigint --q1-sf 0"
ncu --set full --clock-control none --import-source on -k regex:join_probe_lean -s 1 -c 1 -o $O/prof_probe $B > $O/ncu_probe.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xchg_scatter_warp -s 1 -c 1 -o $O/prof_scatter_warp python tools/bench_ops.py 3e8 > $O/ncu_scatter_warp.log 2>&1
ls -la $O
