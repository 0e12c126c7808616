set -x
mkdir -p gpurun_out/r2d
python -m pytest tests/test_gpu_workloads.py tests/test_gpu_synth.py -m gpu -q -x --timeout 900 > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
tail -15 gpurun_out/r2d/pytest.log
python bench.py --workload q3way > gpurun_out/r2d/q3way_n1.json 2> gpurun_out/r2d/q3way_n1.err; tail -3 gpurun_out/r2d/q3way_n1.err
python bench.py --workload star > gpurun_out/r2d/star_n1.json 2> gpurun_out/r2d/star_n1.err; tail -3 gpurun_out/r2d/star_n1.err
python bench.py --workload q1 > gpurun_out/r2d/q1_n1.json 2> gpurun_out/r2d/q1_n1.err; tail -3 gpurun_out/r2d/q1_n1.err
cat gpurun_out/r2d/*_n1.json | cut -c1-400
