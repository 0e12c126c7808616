set -x
O=gpurun_out/r3i
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest.log 2>&1; tail -6 $O/pytest.log
python bench.py > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -2 $O/bench.err
