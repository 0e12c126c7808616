set -x
O=gpurun_out/r3s
mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1; tail -c 300 $O/launches_bench.log; wc -l $O/launches_bench.csv
