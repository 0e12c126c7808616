set -x
O=gpurun_out/r3d
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 800 > $O/pytest_dist.log 2>&1; tail -4 $O/pytest_dist.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577"
timeout 600 $T bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -c 400 $O/bench_n2.json; tail -2 $O/bench_n2.err
timeout 600 $T bench.py --gpus 2 --workload q3way > $O/q3way_n2.json 2> $O/q3way_n2.err; tail -c 300 $O/q3way_n2.json
timeout 600 $T bench.py --gpus 2 --workload star > $O/star_n2.json 2> $O/star_n2.err; tail -c 300 $O/star_n2.json
