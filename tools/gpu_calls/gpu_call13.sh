set -x
mkdir -p gpurun_out/r2m
timeout 600 python -m pytest tests/test_gpu_partial_aggregation.py -m gpu -q --timeout 300 > gpurun_out/r2m/pa.log 2>&1; tail -15 gpurun_out/r2m/pa.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2m/pytest.log 2>&1; tail -8 gpurun_out/r2m/pytest.log
