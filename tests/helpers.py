"""Shared helpers: run a case through the CPU oracle and through the GPU operators, as materialised rows."""
import json
import os

import numpy as np

from trino_b200 import abi
from trino_b200.page import Block, Page

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def reference_cases():
    with open(os.path.join(GOLDEN, "reference_cases.json")) as f:
        return json.load(f)


def oracle_join_rows(build_page, probe_page, build_key, probe_key, probe_out, build_out, join_type, single_match, force_default=False):
    """Rows the reference's LookupJoinOperator emits (probe output channels then build output channels)."""
    import oracle_lib as o
    bk = list(build_key) if isinstance(build_key, (list, tuple)) else [build_key]
    pk = list(probe_key) if isinstance(probe_key, (list, tuple)) else [probe_key]
    j = o.Join(build_page, bk, force_default=force_default)
    pos = j.positions(probe_page, pk)
    pi, bi = j.expand(pos, join_type, single_match)
    pcols = [probe_page.get_block(c).flatten().to_pylist() for c in probe_out]
    bcols = [build_page.get_block(c).flatten().to_pylist() for c in build_out]
    rows = []
    for p, b in zip(pi, bi):
        rows.append(tuple(c[p] for c in pcols) + tuple((c[b] if b >= 0 else None) for c in bcols))
    j.close()
    return rows


def join_type_of(case):
    from trino_b200 import abi
    return {"inner": abi.JOIN_INNER, "lookup_outer": abi.JOIN_LOOKUP_OUTER, "full_outer": abi.JOIN_FULL_OUTER}.get(case["join_type"], abi.JOIN_PROBE_OUTER)


def oracle_outer_rows(build_page, probe_pages, build_key, probe_key, num_probe_out, build_out, join_type, single_match):
    """Rows of the LookupOuterOperator after all probe pages: unvisited build positions in order (OuterLookupSource.java:109-139)"""
    import oracle_lib as o
    bk = list(build_key) if isinstance(build_key, (list, tuple)) else [build_key]
    pk = list(probe_key) if isinstance(probe_key, (list, tuple)) else [probe_key]
    j = o.Join(build_page, bk)
    visited = set()
    for p in probe_pages:
        pos = j.positions(p, pk)
        _, bi = j.expand(pos, join_type, single_match)
        visited.update(int(b) for b in bi if b >= 0)
    j.close()
    bcols = [build_page.get_block(c).flatten().to_pylist() for c in build_out]
    return [tuple([None] * num_probe_out) + tuple(c[b] for c in bcols) for b in range(build_page.position_count) if b not in visited]


def gpu_join_rows(ctx, build_pages, probe_pages, build_key, probe_key, probe_out, build_out, join_type, single_match, by_reference=False):
    from trino_b200 import operators as ops
    bridge = ops.JoinBridge()
    bk = list(build_key) if isinstance(build_key, (list, tuple)) else [build_key]
    pk = list(probe_key) if isinstance(probe_key, (list, tuple)) else [probe_key]
    bf = ops.HashBuilderOperatorFactory(ctx, bridge, bk, build_out)
    b = bf.create_operator()
    for p in build_pages:
        assert b.needs_input()
        b.add_input(p)
    b.finish()
    assert b.is_finished()
    pf = ops.LookupJoinOperatorFactory(ctx, bridge, join_type, single_match, pk, probe_out)
    j = pf.create_operator()
    if by_reference:
        j.set_passthrough_by_reference(True)
    out = ops.drive(j, probe_pages)
    rows = []
    for page in out:
        rows.extend(page.rows())
    j.close()
    b.close()
    bridge.lookup_source.close()
    return rows


def rows_equal(a, b, rel=0.0):
    if len(a) != len(b):
        return False
    for ra, rb in zip(a, b):
        if len(ra) != len(rb):
            return False
        for x, y in zip(ra, rb):
            if x is None or y is None:
                if x is not y:
                    return False
            elif isinstance(x, float) or isinstance(y, float):
                if x != x and y != y:
                    continue
                if rel == 0.0:
                    if x != y:
                        return False
                elif abs(x - y) > rel * max(abs(x), abs(y), 1e-300):
                    return False
            elif x != y:
                return False
    return True


def random_bigint_block(rng, n, lo, hi, null_frac=0.0):
    v = rng.integers(lo, hi, size=n, dtype=np.int64)
    nulls = rng.random(n) < null_frac if null_frac > 0 else None
    return Block.bigint(v, nulls)
