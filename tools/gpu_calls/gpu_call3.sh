set -x
mkdir -p gpurun_out/r2c
python -m pytest tests/test_gpu_dist.py tests/test_gpu_partition.py -m gpu -q -x --timeout 900 > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
tail -5 gpurun_out/r2c/pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2c/bench_n2.json 2> gpurun_out/r2c/bench_n2.err
TGPU_XCHG_PID_ARRAY=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-e2e > gpurun_out/r2c/bench_n2_pidarray.json 2> gpurun_out/r2c/bench_n2_pidarray.err
python tools/bench_ops.py 300000000 > gpurun_out/r2c/ops.log 2>&1
tail -c 900 gpurun_out/r2c/bench_n2.json; tail -5 gpurun_out/r2c/bench_n2.err; cat gpurun_out/r2c/ops.log | tail -1
