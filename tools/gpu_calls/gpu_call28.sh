set -x
O=gpurun_out/r3b
mkdir -p $O
for cfg in "50 16777216" "75 16777216" "100 16777216" "50 33554432" "75 33554432" "50 8388608"; do set -- $cfg; TGPU_AGG_G_SIZE_PCT=$1 TGPU_AGG_SLICE_BYTES=$2 python tools/bench_agg_only.py 150000000 10000000 > $O/agg_p$1_s$2.log 2>&1; echo "pct $1 slice $2: $(tail -1 $O/agg_p$1_s$2.log)"; done
timeout 600 python -m pytest tests/test_gpu_groupby.py -m gpu -q --timeout 600 2>&1 | tail -2
