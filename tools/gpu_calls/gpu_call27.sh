set -x
O=gpurun_out/r3a
mkdir -p $O
for cfg in "4 3" "2 4" "2 5" "2 6" "1 8" "8 2" "4 4"; do set -- $cfg; TGPU_AGG_G_ROWS=$1 TGPU_AGG_G_MINB=$2 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_r$1_m$2.log 2>&1; echo "rows $1 minb $2: $(tail -1 $O/agg_r$1_m$2.log)"; done
