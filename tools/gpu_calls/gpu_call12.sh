set -x
mkdir -p gpurun_out/r2l
python -m pytest tests/test_gpu_join.py tests/test_gpu_groupby.py tests/test_gpu_workloads.py -m gpu -q --timeout 900 > gpurun_out/r2l/pytest.log 2>&1; tail -5 gpurun_out/r2l/pytest.log
python bench.py --no-cpu-baseline --no-e2e --no-groupby-bigint > gpurun_out/r2l/bench.json 2> gpurun_out/r2l/bench.err; tail -3 gpurun_out/r2l/bench.err
TGPU_JOIN_NO_SPAN=1 python bench.py --no-cpu-baseline --no-e2e --no-groupby-bigint --no-shuffled --q1-sf 0 > gpurun_out/r2l/bench_nospan.json 2> gpurun_out/r2l/bench_nospan.err
TGPU_JOIN_NO_DENSE=1 python bench.py --no-cpu-baseline --no-e2e --no-groupby-bigint --no-shuffled --q1-sf 0 > gpurun_out/r2l/bench_nodense.json 2> gpurun_out/r2l/bench_nodense.err
python tools/bench_agg_only.py > gpurun_out/r2l/agg.log 2>&1; tail -1 gpurun_out/r2l/agg.log
