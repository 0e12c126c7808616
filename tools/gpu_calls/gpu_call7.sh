set -x
mkdir -p gpurun_out/r2g
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g/pytest.log
grep -v "^$" gpurun_out/r2g/pytest.log | tail -40 | cut -c1-250
python tools/bench_ops.py > gpurun_out/r2g/ops.log 2>&1; tail -3 gpurun_out/r2g/ops.log
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err; tail -3 gpurun_out/r2g/bench.err
