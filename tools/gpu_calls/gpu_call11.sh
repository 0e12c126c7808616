set -x
mkdir -p gpurun_out/r2k
for minb in 2 3 4; do for pct in 75 37; do
  echo "minb=$minb pct=$pct"; TGPU_AGG_G_MINB=$minb TGPU_AGG_G_SIZE_PCT=$pct python tools/bench_agg_only.py 2>&1 | tail -1
  echo "noslices minb=$minb pct=$pct"; TGPU_AGG_NO_SLICES=1 TGPU_AGG_G_MINB=$minb TGPU_AGG_G_SIZE_PCT=$pct python tools/bench_agg_only.py 2>&1 | tail -1
done; done > gpurun_out/r2k/agg_sweep.log 2>&1
cat gpurun_out/r2k/agg_sweep.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k/q1_launches.csv python tools/bench_q1_only.py 300 > gpurun_out/r2k/q1_ncu.log 2>&1
tail -c 600 gpurun_out/r2k/q1_ncu.log
