set -x
O=gpurun_out/r2t
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_decimal.py -m gpu -q --timeout 600 > $O/decimal.log 2>&1; tail -30 $O/decimal.log
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_partial_aggregation.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
