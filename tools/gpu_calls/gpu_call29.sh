O=gpurun_out/r3c
mkdir -p $O
for pct in 25 35 50; do TGPU_AGG_G_SIZE_PCT=$pct python tools/bench_agg_only.py 150000000 10000000 > $O/agg_p$pct.log 2>&1; echo "pct $pct: $(tail -1 $O/agg_p$pct.log)"; done
for pct in 25 50; do TGPU_AGG_G_SIZE_PCT=$pct python tools/bench_agg_only.py 150000000 1000000 > $O/agg1m_p$pct.log 2>&1; echo "1M groups pct $pct: $(tail -1 $O/agg1m_p$pct.log)"; done
