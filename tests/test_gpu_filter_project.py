"""GPU parity: FilterAndProjectOperator / PageProcessor vs a numpy restatement of the reference semantics
(NULL-rejecting filters M/sql/gen/columnar/ColumnarFilter.java:27-30, Kleene AND/OR, checked BIGINT arithmetic
M/type/BigintOperators.java:52-110, unfused IEEE DOUBLE arithmetic M/type/DoubleOperators.java:66-86)."""
import numpy as np
import pytest

from trino_b200 import abi
from trino_b200 import operators as ops
from trino_b200.page import Block, DictionaryBlock, Page

pytestmark = pytest.mark.gpu
B, D, BOOL = abi.V_BIGINT, abi.V_DOUBLE, abi.V_BOOLEAN


def run(ctx, program, pages):
    op = ops.FilterAndProjectOperatorFactory(ctx, program).create_operator()
    out = ops.drive(op, pages)
    op.close()
    return [r for p in out for r in p.rows()]


def test_q1_projection_is_bit_exact_and_filter_rejects_null(ctx):
    rng = np.random.default_rng(1)
    n = 100_000
    ship = Block.integer(rng.integers(8036, 10592, n), rng.random(n) < 0.01)
    ep = Block.double(rng.uniform(900, 105000, n).round(2))
    disc = Block.double(rng.integers(0, 11, n) / 100.0)
    tax = Block.double(rng.integers(0, 9, n) / 100.0, rng.random(n) < 0.01)
    one = ops.Const(1.0, D)
    dp = ops.Call(abi.EX_MUL, ops.Col(1, D), ops.Call(abi.EX_SUB, one, ops.Col(2, D)))
    ch = ops.Call(abi.EX_MUL, ops.Call(abi.EX_MUL, ops.Col(1, D), ops.Call(abi.EX_SUB, one, ops.Col(2, D))), ops.Call(abi.EX_ADD, one, ops.Col(3, D)))
    prog = ops.PageProcessorProgram(ops.Call(abi.EX_LE, ops.Col(0, B), ops.Const(10471, B)), [0, dp, ch])
    rows = run(ctx, prog, [Page(ship, ep, disc, tax)])
    sel = (ship.values <= 10471) & ~ship.nulls
    e, d, t = ep.values[sel], disc.values[sel], tax.values[sel]
    want_dp = e * (1.0 - d)                      # numpy never fuses: same IEEE operations in the same order
    want_ch = e * (1.0 - d) * (1.0 + t)
    tn = tax.nulls[sel]
    assert len(rows) == int(sel.sum())
    assert [r[0] for r in rows] == ship.values[sel].tolist()
    assert np.array([r[1] for r in rows]).tobytes() == want_dp.tobytes()
    got_ch = [r[2] for r in rows]
    assert all((g is None) == bool(isn) for g, isn in zip(got_ch, tn))
    assert np.array([g for g in got_ch if g is not None]).tobytes() == want_ch[~tn].tobytes()


def test_three_valued_logic_between_in_and_passthrough(ctx):
    a = Block.bigint([1, 5, None, 7, 10, None, 3, 8])
    b = Block.boolean([True, None, False, None, True, None, False, True])
    s = Block.varchar(["x", None, "zz", "", "q", "w", "e", "r"])
    A_, B_ = ops.Col(0, B), ops.Col(1, BOOL)
    between = ops.Call(abi.EX_BETWEEN, A_, ops.Const(3, B), ops.Const(8, B))
    isin = ops.Call(abi.EX_IN, A_, in_list=[1, 8, 10])
    flt = ops.Call(abi.EX_OR, ops.Call(abi.EX_AND, between, B_), isin)
    prog = ops.PageProcessorProgram(flt, [0, 2, ops.Call(abi.EX_AND, between, B_), ops.Call(abi.EX_OR, between, B_), ops.Call(abi.EX_NOT, B_),
                                          ops.Call(abi.EX_IS_NULL, A_)])
    rows = run(ctx, prog, [Page(a, b, s)])

    def k_and(x, y):
        if x is False or y is False:
            return False
        if x is None or y is None:
            return None
        return True

    def k_or(x, y):
        if x is True or y is True:
            return True
        if x is None or y is None:
            return None
        return False

    want = []
    for av, bv, sv in zip(a.to_pylist(), [None if x is None else bool(x) for x in b.to_pylist()], s.to_pylist()):
        btw = None if av is None else (3 <= av <= 8)
        inn = None if av is None else av in (1, 8, 10)
        f = k_or(k_and(btw, bv), inn)
        if f is True:
            conv = lambda v: None if v is None else int(v)
            want.append((av, sv, conv(k_and(btw, bv)), conv(k_or(btw, bv)), conv(None if bv is None else not bv), int(av is None)))
    assert rows == want


def test_bigint_arithmetic_checked_and_division(ctx):
    a = Block.bigint([7, -7, 2**62, 5, None])
    b = Block.bigint([2, 2, 2, -1, 3])
    A_, B_ = ops.Col(0, B), ops.Col(1, B)
    prog = ops.PageProcessorProgram(None, [ops.Call(abi.EX_DIV, A_, B_), ops.Call(abi.EX_MOD, A_, B_), ops.Call(abi.EX_SUB, A_, B_), ops.Call(abi.EX_NEG, A_)])
    rows = run(ctx, prog, [Page(a, b)])
    assert rows == [(3, 1, 5, -7), (-3, -1, -9, 7), (2**61, 0, 2**62 - 2, -2**62), (-5, 0, 6, -5), (None, None, None, None)]   # Java truncating / and %
    with pytest.raises(abi.TrinoGpuError) as e:
        run(ctx, ops.PageProcessorProgram(None, [ops.Call(abi.EX_ADD, A_, A_)]), [Page(a, b)])
    assert e.value.code == abi.ERR_NUMERIC_VALUE_OUT_OF_RANGE
    with pytest.raises(abi.TrinoGpuError) as e:
        run(ctx, ops.PageProcessorProgram(None, [ops.Call(abi.EX_MUL, A_, ops.Const(4, B))]), [Page(a, b)])
    assert e.value.code == abi.ERR_NUMERIC_VALUE_OUT_OF_RANGE
    with pytest.raises(abi.TrinoGpuError) as e:
        run(ctx, ops.PageProcessorProgram(None, [ops.Call(abi.EX_DIV, A_, ops.Const(0, B))]), [Page(a, b)])
    assert e.value.code == abi.ERR_DIVISION_BY_ZERO
    # errors in projections are raised only for rows the filter selected (PageProcessor filters first)
    flt = ops.Call(abi.EX_LT, A_, ops.Const(100, B))
    rows = run(ctx, ops.PageProcessorProgram(flt, [ops.Call(abi.EX_ADD, A_, A_)]), [Page(a, b)])
    assert rows == [(14,), (-14,), (10,)]


def test_double_comparisons_casts_and_dictionary_input(ctx):
    x = Block.double([1.5, float("nan"), -0.0, 2.5, None])
    X = ops.Col(0, D)
    prog = ops.PageProcessorProgram(None, [ops.Call(abi.EX_EQ, X, X), ops.Call(abi.EX_NE, X, X), ops.Call(abi.EX_LT, X, ops.Const(2.0, D)),
                                           ops.Call(abi.EX_CAST_DOUBLE_TO_BIGINT, ops.Call(abi.EX_ADD, X, ops.Const(0.0, D))),
                                           ops.Call(abi.EX_DIV, X, ops.Const(0.0, D))])
    page = Page(Block.double([1.5, 2.5, -2.5, 2.4999, None]))
    rows = run(ctx, prog, [page])
    assert [r[3] for r in rows] == [2, 3, -3, 2, None]                       # HALF_UP rounding like DoubleMath.roundToLong
    rows = run(ctx, ops.PageProcessorProgram(None, [ops.Call(abi.EX_EQ, X, X), ops.Call(abi.EX_NE, X, X), ops.Call(abi.EX_DIV, X, ops.Const(0.0, D))]), [Page(x)])
    assert rows[0][:2] == (1, 0) and rows[1][:2] == (0, 1) and rows[4] == (None, None, None)
    assert rows[0][2] == float("inf") and rows[1][2] != rows[1][2] and rows[2][2] != rows[2][2]
    # dictionary block in, values out
    d = DictionaryBlock(Block.bigint([10, 20, None]), [2, 1, 0, 1])
    rows = run(ctx, ops.PageProcessorProgram(ops.Call(abi.EX_GE, ops.Col(0, B), ops.Const(15, B)), [0, ops.Call(abi.EX_CAST_BIGINT_TO_DOUBLE, ops.Col(0, B))]), [Page(d)])
    assert rows == [(20, 20.0), (20, 20.0)]


def test_empty_and_all_filtered_pages(ctx):
    prog = ops.PageProcessorProgram(ops.Call(abi.EX_GT, ops.Col(0, B), ops.Const(100, B)), [0])
    assert run(ctx, prog, [Page(Block.bigint([1, 2, 3])), Page(Block.bigint([]), position_count=0)]) == []
    assert run(ctx, prog, [Page(Block.bigint([1, 200, 3]))]) == [(200,)]


@pytest.mark.parametrize("selectivity", [0.001, 0.3, 0.97])
def test_chunked_two_pass_form_matches_selection_vector_form(ctx, selectivity, monkeypatch):
    """Fixed-width pass-through channels + a filter run without a selection vector (flags + chunk ranks); the rows, their order
    and NULLs must be those of the select/gather form and of numpy, across several chunks and pages."""
    rng = np.random.default_rng(int(selectivity * 1000))
    pages = []
    for n in (700_000, 1, 1023, 300_001):
        pages.append(Page(Block.double(rng.random(n), rng.random(n) < 0.02), Block.bigint(rng.integers(-10**9, 10**9, n), rng.random(n) < 0.1),
                          Block.integer(rng.integers(0, 100, n)), Block.tinyint(rng.integers(-5, 5, n), rng.random(n) < 0.3),
                          Block.smallint(rng.integers(0, 30000, n))))
    flt = ops.Call(abi.EX_LT, ops.Col(0, D), ops.Const(selectivity, D))
    prog = ops.PageProcessorProgram(flt, [1, ops.Call(abi.EX_ADD, ops.Col(1, B), ops.Col(2, B)), 3, 4, ops.Call(abi.EX_MUL, ops.Col(0, D), ops.Const(2.0, D)), 2])
    got = run(ctx, prog, pages)
    want = []
    def nulls_of(blk):
        return blk.nulls if blk.nulls is not None else np.zeros(len(blk.values), dtype=bool)

    for p in pages:
        x, k, i, t, s = (p.get_block(c) for c in range(5))
        xn, kn_, tn = nulls_of(x), nulls_of(k), nulls_of(t)
        sel = (x.values < selectivity) & ~xn
        for r in np.nonzero(sel)[0]:
            kn = bool(kn_[r])
            want.append((None if kn else int(k.values[r]), None if kn else int(k.values[r]) + int(i.values[r]), None if tn[r] else int(t.values[r]),
                         int(s.values[r]), float(x.values[r]) * 2.0, int(i.values[r])))
    assert got == want
    monkeypatch.setenv("TGPU_FP_SELECTION_VECTOR", "1")
    assert run(ctx, prog, pages) == want


def test_page_processor_reference_cases(ctx):
    """TestPageProcessor.java :126-175 restated with expression filters: testPartialFilter (positionsRange(25, 50) of a 0..99 sequence
    -> rows 25..74 through InputPageProjection(0)), testSelectAllFilter (the page itself), testSelectNoneFilter (no output page)."""
    seq = Page(Block.bigint(np.arange(100)))
    c = ops.Col(0, B)
    partial = ops.Call(abi.EX_AND, ops.Call(abi.EX_GE, c, ops.Const(25, B)), ops.Call(abi.EX_LT, c, ops.Const(75, B)))
    assert run(ctx, ops.PageProcessorProgram(partial, [0]), [seq]) == [(i,) for i in range(25, 75)]
    assert run(ctx, ops.PageProcessorProgram(ops.Call(abi.EX_GE, c, ops.Const(0, B)), [0]), [seq]) == [(i,) for i in range(100)]
    op = ops.FilterAndProjectOperatorFactory(ctx, ops.PageProcessorProgram(ops.Call(abi.EX_LT, c, ops.Const(0, B)), [0])).create_operator()
    op.add_input(seq)
    assert op.get_output() is None
    op.close()
