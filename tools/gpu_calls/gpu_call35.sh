set -x
O=gpurun_out/r3j
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_groupby.py tests/test_gpu_partial_aggregation.py tests/test_gpu_decimal.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; tail -3 $O/pytest.log
python tools/bench_q1_only.py 300 > $O/q1_minb4.json 2> $O/q1_minb4.err; python -c "import json;d=json.load(open('$O/q1_minb4.json'));print('minb4', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['utf8_keys']['ms_per_step'])"
TGPU_AGG_S_MINB=3 python tools/bench_q1_only.py 300 > $O/q1_minb3.json 2> $O/q1_minb3.err; python -c "import json;d=json.load(open('$O/q1_minb3.json'));print('minb3', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['utf8_keys']['ms_per_step'])"
