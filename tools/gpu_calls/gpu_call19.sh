set -x
O=gpurun_out/r2s
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_decimal.py -m gpu -q --timeout 600 -x > $O/decimal.log 2>&1; tail -30 $O/decimal.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_decimal.py > $O/pytest.log 2>&1; tail -8 $O/pytest.log
python tools/bench_agg_only.py 150000000 10000000 > $O/agg.log 2>&1; tail -1 $O/agg.log
