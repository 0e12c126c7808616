set -x
mkdir -p gpurun_out/r2f
python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log
tail -40 gpurun_out/r2f/pytest.log
