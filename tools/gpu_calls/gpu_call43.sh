set -x
O=gpurun_out/r3r
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_serde.py tests/test_gpu_join.py tests/test_gpu_real.py -m gpu -q --timeout 500 > $O/pytest.log 2>&1; tail -15 $O/pytest.log
