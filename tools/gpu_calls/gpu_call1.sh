set -x
mkdir -p gpurun_out/r2a
nvidia-smi -L > gpurun_out/r2a/gpus.txt
python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
tail -5 gpurun_out/r2a/pytest.log
python bench.py --no-cpu-baseline --q1-sf 0 --no-e2e > gpurun_out/r2a/bench_mode2.json 2> gpurun_out/r2a/bench_mode2.err
TGPU_JOIN_HASH=1 python bench.py --no-cpu-baseline --q1-sf 0 --no-e2e > gpurun_out/r2a/bench_mode1.json 2> gpurun_out/r2a/bench_mode1.err
python bench.py --no-cpu-baseline --q1-sf 0 --no-e2e --shuffle-probe > gpurun_out/r2a/bench_mode2_shuf.json 2> gpurun_out/r2a/bench_mode2_shuf.err
python tools/bench_ops.py > gpurun_out/r2a/ops.log 2>&1
TGPU_AGG_PHYSICAL_SLICES=1 python tools/bench_ops.py > gpurun_out/r2a/ops_physical.log 2>&1
python bench.py > gpurun_out/r2a/bench_full.json 2> gpurun_out/r2a/bench_full.err
python bench.py --impl reference > gpurun_out/r2a/bench_ref.json 2> gpurun_out/r2a/bench_ref.err
tail -c 600 gpurun_out/r2a/bench_mode2.json; tail -c 300 gpurun_out/r2a/bench_mode1.json
