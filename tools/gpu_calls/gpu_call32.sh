set -x
O=gpurun_out/r3g
mkdir -p $O
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29599"
timeout 900 $T bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; tail -c 600 $O/bench_n8.json; tail -3 $O/bench_n8.err
