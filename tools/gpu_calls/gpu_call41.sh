set -x
O=gpurun_out/r3p
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x --timeout 500 > $O/pytest_dist.log 2>&1; tail -5 $O/pytest_dist.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29588"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.json; tail -3 $O/bench_n2.err
