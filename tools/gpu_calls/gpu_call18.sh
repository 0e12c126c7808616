set -x
O=gpurun_out/r2r
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_partial_aggregation.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
for mb in 2 3 4; do TGPU_AGG_G_MINB=$mb python tools/bench_agg_only.py 150000000 10000000 > $O/agg_minb$mb.log 2>&1; echo minb $mb; tail -1 $O/agg_minb$mb.log; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_agg.csv python tools/bench_agg_only.py 150000000 10000000 > $O/launches_agg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_agg_general_jit -s 1 -c 1 -o $O/prof_gjit python tools/bench_agg_only.py 150000000 10000000 > $O/ncu_gjit.log 2>&1
