#!/usr/bin/env python
"""Q1 GROUP-BY block of bench.py alone (INT8 codes and VARCHAR(1) keys) - for ncu launch lists of the aggregation path."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from trino_b200 import operators as ops  # noqa: E402

ctx = ops.Context(0)
args = types.SimpleNamespace(q1_sf=float(sys.argv[1]) if len(sys.argv) > 1 else 300.0)
print(json.dumps(bench.bench_q1(ctx, args)))
ctx.close()
