set -x
O=gpurun_out/r2u
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $T bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; tail -c 2500 $O/bench_n8.json; tail -3 $O/bench_n8.err
timeout 600 $T bench.py --gpus 8 --workload q1 > $O/q1_n8.json 2> $O/q1_n8.err; tail -c 700 $O/q1_n8.json; tail -2 $O/q1_n8.err
timeout 900 $T bench.py --gpus 8 --workload q3way > $O/q3way_n8.json 2> $O/q3way_n8.err; tail -c 900 $O/q3way_n8.json; tail -2 $O/q3way_n8.err
timeout 900 $T bench.py --gpus 8 --workload star > $O/star_n8.json 2> $O/star_n8.err; tail -c 900 $O/star_n8.json; tail -2 $O/star_n8.err
