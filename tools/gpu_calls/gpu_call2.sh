set -x
mkdir -p gpurun_out/r2b
nvidia-smi -L > gpurun_out/r2b/gpus.txt
python -m pytest tests/test_gpu_dist.py -m gpu -q -x --timeout 900 > gpurun_out/r2b/pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b/pytest_dist.log
tail -5 gpurun_out/r2b/pytest_dist.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2b/bench_n2.json 2> gpurun_out/r2b/bench_n2.err
TGPU_JOIN_HASH=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --no-e2e > gpurun_out/r2b/bench_n2_mode1.json 2> gpurun_out/r2b/bench_n2_mode1.err
TGPU_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --no-e2e --steps 2 --warmup 1 > gpurun_out/r2b/bench_n2_trace.json 2> gpurun_out/r2b/bench_n2_trace.err
tail -c 1500 gpurun_out/r2b/bench_n2.json; tail -5 gpurun_out/r2b/bench_n2.err
