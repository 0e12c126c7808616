set -x
O=gpurun_out/r2y
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_partition.py tests/test_gpu_workloads.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
for mb in 2 3 4; do TGPU_XCHG_LEAN_MINB=$mb python tools/bench_ops.py 3e8 > $O/ops_minb$mb.log 2>&1; tail -1 $O/ops_minb$mb.log | cut -c1-160; done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_ops.csv python tools/bench_ops.py 3e8 > $O/launches_ops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xchg_scatter_lean8 -s 1 -c 1 -o $O/prof_lean8 python tools/bench_ops.py 3e8 > $O/ncu_lean8.log 2>&1
