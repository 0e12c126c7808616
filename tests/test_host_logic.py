"""Host-side logic that needs no GPU: Page/Block model, ABI page views, expression compilation, factories."""
import ctypes as C

import numpy as np
import pytest

from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import AbiPage, Block, DictionaryBlock, Page, RunLengthEncodedBlock


def test_block_roundtrip_and_nulls():
    b = Block.bigint([1, None, 3])
    assert b.to_pylist() == [1, None, 3]
    assert b.nulls.tolist() == [False, True, False]
    v = Block.varchar(["ab", None, ""])
    assert v.to_pylist() == [b"ab", None, b""]
    assert v.offsets.tolist() == [0, 2, 2, 2]
    d = DictionaryBlock(Block.double([1.5, 2.5]), [1, 1, 0])
    assert d.to_pylist() == [2.5, 2.5, 1.5]
    assert d.flatten().to_pylist() == [2.5, 2.5, 1.5]
    r = RunLengthEncodedBlock(Block.integer([7]), 4)
    assert r.to_pylist() == [7, 7, 7, 7]


def test_abi_page_view_points_at_numpy_buffers():
    a = Block.bigint([1, 2, None])
    page = Page(a, DictionaryBlock(Block.integer([5, 6]), [0, 1, 1]))
    ap = AbiPage(page, nulls_as_bytemap=True)
    assert ap.page.num_rows == 3 and ap.page.num_columns == 2
    c0 = ap.columns[0]
    assert c0.type == abi.INT64 and c0.data == a.values.ctypes.data and c0.flags == abi.COL_NULLS_BYTEMAP
    c1 = ap.columns[1]
    assert c1.type == abi.DICT32 and c1.dictionary.contents.type == abi.INT32 and c1.dictionary.contents.length == 2
    ap2 = AbiPage(page, nulls_as_bytemap=False)
    bits = C.cast(ap2.columns[0].validity, C.POINTER(C.c_uint8))[0]
    assert bits == 0b011


def test_page_position_count_mismatch_is_rejected():
    with pytest.raises(AssertionError):
        Page(Block.bigint([1, 2]), Block.bigint([1]))


def test_program_compiler_q1_shape():
    # q01.sql: shipdate <= DATE '1998-09-02'; projections rf, ls, qty, ep, ep*(1-d), ep*(1-d)*(1+t), disc
    D = abi.V_DOUBLE
    ep, disc, tax = ops.Col(4, D), ops.Col(5, D), ops.Col(6, D)
    disc_price = ops.Call(abi.EX_MUL, ep, ops.Call(abi.EX_SUB, ops.Const(1.0, D), disc))
    charge = ops.Call(abi.EX_MUL, ops.Call(abi.EX_MUL, ep, ops.Call(abi.EX_SUB, ops.Const(1.0, D), disc)), ops.Call(abi.EX_ADD, ops.Const(1.0, D), tax))
    prog = ops.PageProcessorProgram(ops.Call(abi.EX_LE, ops.Col(0, abi.V_BIGINT), ops.Const(10471, abi.V_BIGINT)), [1, 2, 3, 4, disc_price, charge, 5])
    s = prog.struct
    assert s.filter_temp >= 0 and s.num_filter_insns == 1
    assert s.num_projections == 7
    kinds = [prog.projections[i][0] for i in range(7)]
    assert kinds == [0, 0, 0, 0, 1, 1, 0]
    temps = {prog.projections[4][1], prog.projections[5][1], s.filter_temp}
    assert len(temps) == 3                      # live results never share a temporary
    assert s.num_insns == 1 + 2 + 4
    assert all(0 <= prog.insns[i][2] < 8 for i in range(s.num_insns))


def test_program_compiler_rejects_too_many_temps():
    e = ops.Col(0, abi.V_BIGINT)
    projections = [ops.Call(abi.EX_ADD, e, ops.Const(i, abi.V_BIGINT)) for i in range(9)]
    with pytest.raises(ValueError):
        ops.PageProcessorProgram(None, projections)


def test_factory_protocol():
    class F(ops.OperatorFactory):
        def _create(self):
            return "op"
    f = F()
    assert f.create_operator() == "op"
    f.no_more_operators()
    with pytest.raises(RuntimeError):
        f.create_operator()
    with pytest.raises(RuntimeError):
        ops.HashBuilderOperatorFactory(None, ops.JoinBridge(), [0], []).duplicate()
    with pytest.raises(RuntimeError):
        ops.LookupJoinOperatorFactory(None, ops.JoinBridge(), abi.JOIN_INNER, False, [0], [0]).create_operator()


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) on a tiny scale factor: one JSON line with the
    contract's keys; under torchrun only rank 0 prints."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--sf", "0.02", "--steps", "1", "--warmup", "1", "--cpu-sample-rows", "50000"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env={**os.environ, "RANK": "0", "WORLD_SIZE": "1"})
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "hash_join_probe_rows_per_sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
    quiet = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env={**os.environ, "RANK": "1", "WORLD_SIZE": "2"})
    assert quiet.returncode == 0 and quiet.stdout.strip() == ""
