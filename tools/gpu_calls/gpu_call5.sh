set -x
mkdir -p gpurun_out/r2e
python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 900 > gpurun_out/r2e/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e/pytest.log
tail -30 gpurun_out/r2e/pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29533 bench.py --gpus 2 > gpurun_out/r2e/bench_n2.json 2> gpurun_out/r2e/bench_n2.err; tail -3 gpurun_out/r2e/bench_n2.err
$TR --master-port 29534 bench.py --gpus 2 --workload q3way > gpurun_out/r2e/q3way_n2.json 2> gpurun_out/r2e/q3way_n2.err; tail -3 gpurun_out/r2e/q3way_n2.err
$TR --master-port 29535 bench.py --gpus 2 --workload star > gpurun_out/r2e/star_n2.json 2> gpurun_out/r2e/star_n2.err; tail -3 gpurun_out/r2e/star_n2.err
$TR --master-port 29536 bench.py --gpus 2 --workload q1 > gpurun_out/r2e/q1_n2.json 2> gpurun_out/r2e/q1_n2.err; tail -3 gpurun_out/r2e/q1_n2.err
cat gpurun_out/r2e/*_n2.json | cut -c1-300
