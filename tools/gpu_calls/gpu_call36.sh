set -x
O=gpurun_out/r3k
mkdir -p $O
ncu --set full --clock-control none --import-source on -k regex:tg_agg_small_jit -s 1 -c 1 -o $O/prof_q1 python tools/bench_q1_only.py 300 > $O/ncu_q1.log 2>&1
ls -la $O
