O=gpurun_out/r3u
mkdir -p $O
timeout 70 python bench.py --workload q3way --steps 5 --warmup 3 > $O/q3way.json 2> $O/q3way.err; tail -c 700 $O/q3way.json; tail -2 $O/q3way.err
