"""The C++ Driver-loop harness (tests/harness/driver_loop.cpp) under pytest: 8192-row "Java pages" with boolean[] null maps, marshalled batch by
batch through pinned staging into the C ABI exactly as java/io/trino/spi/block/PageMarshaller.java does, driven with the Operator protocol of
M/operator/Driver.java:391-424 by several driver threads, results checked against the oracle inside the harness (mismatches == 0)."""
import json
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    import __graft_entry__ as g
    exe = g.build_harness()
    r = subprocess.run([exe, *[str(a) for a in args]], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_q1_driver_loop_matches_oracle():
    line = _run("q1", 3_000_000, 3)
    assert line["mismatches"] == 0 and line["groups"] == 4 and line["rows_per_s"] > 0


def test_join_driver_loop_matches_oracle():
    line = _run("join", 400_000, 3)
    assert line["mismatches"] == 0 and line["output_rows"] == line["probe_rows"]
