set -x
mkdir -p gpurun_out/r2i
ncu --set full --clock-control none --import-source on -k regex:tg_agg_general_jit -s 1 -c 1 -o gpurun_out/r2i/prof_gjit python tools/bench_agg_only.py 150000000 10000000 > gpurun_out/r2i/ncu_gjit.log 2>&1
TGPU_AGG_GENERAL_INTERPRETED=1 python tools/bench_agg_only.py > gpurun_out/r2i/agg_interp.log 2>&1; cat gpurun_out/r2i/agg_interp.log
TGPU_AGG_NO_SLICES=1 python tools/bench_agg_only.py > gpurun_out/r2i/agg_noslices.log 2>&1; cat gpurun_out/r2i/agg_noslices.log
python tools/bench_agg_only.py 150000000 1000000 > gpurun_out/r2i/agg_1m.log 2>&1; cat gpurun_out/r2i/agg_1m.log
python tools/bench_agg_only.py 150000000 100000 > gpurun_out/r2i/agg_100k.log 2>&1; cat gpurun_out/r2i/agg_100k.log
python bench.py --no-cpu-baseline --no-e2e --no-shuffled --no-groupby-bigint > gpurun_out/r2i/bench_q1.json 2> gpurun_out/r2i/bench_q1.err; tail -2 gpurun_out/r2i/bench_q1.err
python -m pytest tests/test_gpu_groupby.py -m gpu -q --timeout 900 -k "varchar" > gpurun_out/r2i/pytest.log 2>&1; tail -3 gpurun_out/r2i/pytest.log
python -m pytest tests/test_gpu_dynamic_filter.py -m gpu -q --timeout 900 > gpurun_out/r2i/pytest_df.log 2>&1; tail -15 gpurun_out/r2i/pytest_df.log | cut -c1-250
