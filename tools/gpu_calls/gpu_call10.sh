set -x
mkdir -p gpurun_out/r2j
python -m pytest tests/test_gpu_groupby.py -m gpu -q --timeout 900 > gpurun_out/r2j/pytest.log 2>&1; tail -3 gpurun_out/r2j/pytest.log
python tools/bench_agg_only.py > gpurun_out/r2j/agg_default.log 2>&1; cat gpurun_out/r2j/agg_default.log
TGPU_AGG_NO_SLICES=1 python tools/bench_agg_only.py > gpurun_out/r2j/agg_noslices.log 2>&1; cat gpurun_out/r2j/agg_noslices.log
TGPU_AGG_NO_SLICES=1 python tools/bench_agg_only.py 150000000 100000 > gpurun_out/r2j/agg_100k.log 2>&1; cat gpurun_out/r2j/agg_100k.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2j/agg_launches.csv python tools/bench_agg_only.py 150000000 10000000 > gpurun_out/r2j/agg_ncu.log 2>&1
TGPU_AGG_NO_SLICES=1 ncu --set full --clock-control none --import-source on -k regex:tg_agg_general_jit -s 1 -c 1 -o gpurun_out/r2j/prof_gjit python tools/bench_agg_only.py 150000000 10000000 > gpurun_out/r2j/ncu_gjit.log 2>&1
