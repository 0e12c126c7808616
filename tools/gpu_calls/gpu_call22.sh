set -x
O=gpurun_out/r2v
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555"
timeout 900 $T bench.py --gpus 4 --steps 5 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; tail -c 1200 $O/bench_n4.json; tail -3 $O/bench_n4.err
timeout 600 $T bench.py --gpus 4 --workload q1 > $O/q1_n4.json 2> $O/q1_n4.err; tail -c 500 $O/q1_n4.json
T2="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556"
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 500 > $O/pytest_dist.log 2>&1; tail -5 $O/pytest_dist.log
timeout 600 $T2 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 600 $O/bench_n2.json
