set -x
O=gpurun_out/r3f
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; tail -6 $O/pytest.log
python bench.py --impl reference > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 900 $O/bench_ref.json
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json; tail -3 $O/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/launches_bench.log 2>&1
