set -x
O=gpurun_out/r2q
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_join.py tests/test_gpu_groupby.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -4 $O/pytest.log
python tools/bench_q1_only.py 300 > $O/q1_vec.json 2> $O/q1_vec.err; python -c "import json;d=json.load(open('$O/q1_vec.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['utf8_keys']['ms_per_step'])"
TGPU_AGG_S_NO_VEC=1 python tools/bench_q1_only.py 300 > $O/q1_novec.json 2> $O/q1_novec.err; python -c "import json;d=json.load(open('$O/q1_novec.json'));print(d['ms_per_step'], d['roofline']['kernel_ms'], d['utf8_keys']['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_agg.csv python tools/bench_agg_only.py 150000000 10000000 > $O/launches_agg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_agg_general_jit -s 1 -c 1 -o $O/prof_gjit python tools/bench_agg_only.py 150000000 10000000 > $O/ncu_gjit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tg_agg_small_jit -s 1 -c 1 -o $O/prof_q1 python tools/bench_q1_only.py 300 > $O/ncu_q1.log 2>&1
ls -la $O
