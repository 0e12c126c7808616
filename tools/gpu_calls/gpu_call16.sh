set -x
O=gpurun_out/r2p
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 800 -x > $O/pytest_dist.log 2>&1; tail -25 $O/pytest_dist.log
timeout 600 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_workloads.py tests/test_gpu_partition.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python tools/bench_agg_only.py 150000000 10000000 > $O/agg.log 2>&1; tail -2 $O/agg.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
$T bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 1500 $O/bench_n2.json
$T bench.py --gpus 2 --workload q1 > $O/q1_n2.json 2> $O/q1_n2.err; tail -c 800 $O/q1_n2.json; tail -3 $O/q1_n2.err
