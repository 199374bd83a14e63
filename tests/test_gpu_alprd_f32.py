"""ALP-RD float containers (float64 / float32) and float32 value columns through the C ABI, against the oracle's
restatement of internal/encode/float_alprd.go + alp/rd.go and of the reducers."""
import numpy as np
import pytest

import knoxdb_b200 as kb
import oracle as ko

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(31337)
KT = {ko.F64: kb.FLOAT64, ko.F32: kb.FLOAT32}
NPF = {ko.F64: np.float64, ko.F32: np.float32}
UINT = {ko.F64: np.uint64, ko.F32: np.uint32}


@pytest.fixture(scope="module")
def ctx():
    c = kb.Context(0)
    yield c
    c.close()


def _real_floats(t, n):
    v = (RNG.normal(0, 1, n) * RNG.choice([1.0, 1e-6, 1e9], n)).astype(NPF[t])
    if n > 16:
        v[1], v[2], v[3], v[4], v[5] = np.nan, np.inf, -np.inf, 0.0, -0.0
        v[7:12] = v[6]          # duplicates: Equal has something to find
    return v


@pytest.mark.parametrize("t", [ko.F64, ko.F32])
@pytest.mark.parametrize("n", [1, 31, 1025, 70_001])
def test_alprd_blocks_decode_match_gather_and_reduce(ctx, t, n):
    v = _real_floats(t, n)
    shifts = [-1] + ([48, 52, 60] if t == ko.F64 else [16, 20, 28])
    for si, shift in enumerate(shifts):
        blob = ko.store("alprd", t, v, shift=shift)
        assert blob[0] == 14
        oc = ko.Container(t, blob)
        # decode (AppendTo(dst, nil)) is bit-exact, NaN payloads and signed zeros included
        got = ctx.container_decode(KT[t], blob, n)
        assert (got.view(UINT[t]) == v.view(UINT[t])).all()
        ops = [(kb.EQ, v[6 % n], 0), (kb.NE, v[6 % n], 0), (kb.LT, v[n // 2], 0), (kb.LE, v[n // 2], 0), (kb.GT, v[n // 3], 0), (kb.GE, v[n // 3], 0),
               (kb.RANGE, min(v[0], v[n // 2]), max(v[0], v[n // 2])), (kb.LT, np.nan, 0), (kb.NE, np.nan, 0), (kb.GE, -np.inf, 0), (kb.RANGE, -1.0, 1.0)]
        for mode, a, b in ops:
            want = oc.match(mode, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
            bits, cnt = ctx.container_match(KT[t], blob, mode, a, b, nrows=n)
            assert (bits == want).all() and cnt == int(np.unpackbits(want).sum()), (t, n, shift, mode, a, b)
        # resident block: filter on it, aggregate over it, gather from it
        pack = 40_000 + 10 * si + (0 if t == ko.F64 else 5)
        assert ctx.block_put(pack, 1, 1, KT[t], blob) == n
        key = RNG.integers(0, 100, n).astype(np.uint64)
        ctx.block_put(pack, 1, 2, kb.UINT64, ko.store("best", ko.U64, key))
        fin = np.where(np.isfinite(v), v, NPF[t](0)).astype(NPF[t])          # NaN / Inf stay out of the aggregate parity (reducer order dependence)
        ctx.block_put(pack, 1, 3, KT[t], ko.store("alprd", t, fin, shift=shift))
        prog = kb.Program(ctx, [kb.Leaf(1, KT[t], kb.GT, -0.5), kb.Leaf(2, kb.UINT64, kb.LT, 60)])
        r = ctx.scan(prog, [(pack, 1)], nrows=[n], want_bitsets=True, aggs=[(3, KT[t])])
        want = ko.tree_eval([0, 1, 0xFE], [oc.match(ko.GT, ko.scalar_u64(t, -0.5), 0), ko.Container(ko.U64, ko.store("best", ko.U64, key)).match(ko.LT, 60, 0)], n)
        assert (r["bitsets"][0] == want).all()
        sel = np.flatnonzero(np.unpackbits(want, bitorder="little")[:n])
        g = r["aggs"][0]
        assert g.count == sel.size
        if sel.size:
            x = fin[sel].astype(np.float64)
            got_sum = float(np.uint64(g.sum_bits).view(np.float64)) if t == ko.F64 else float(np.uint32(g.sum_bits).view(np.float32))
            exact = float(np.sum(x.astype(np.longdouble)))
            assert abs(got_sum - exact) <= (1e-12 if t == ko.F64 else 1e-6) * max(abs(exact), float(np.abs(x).max()))
            mn = np.uint64(g.min_bits).view(np.float64) if t == ko.F64 else np.uint32(g.min_bits).view(np.float32)
            mx = np.uint64(g.max_bits).view(np.float64) if t == ko.F64 else np.uint32(g.max_bits).view(np.float32)
            assert mn == fin[sel].min() and mx == fin[sel].max()
        s = ctx.scan_select(prog, [(pack, 1)])
        gathered = ctx.gather([(pack, 1)], 1, KT[t], s["sel"], s["sel_off"])
        assert (gathered.view(UINT[t]) == v[sel].view(UINT[t])).all()
        prog.close()
        for f in (1, 2, 3):
            ctx.block_drop(pack, 1, f)


def test_float32_raw_columns_reduce_and_prune(ctx):
    """float32 aggregates are accumulated in float64 and rounded once: min / max equal the reference's, the sum is within
    float32 rounding of the exact sum (the reference's running float32 sum is further away); float32 zone maps prune
    like MatchRangeVectors on float32 (internal/operator/filter/match_num.go:357-588)"""
    sizes = [50_000, 8_192, 33]
    vals = [(RNG.integers(-10**6, 10**6, n) / 128.0).astype(np.float32) for n in sizes]
    for p, v in enumerate(vals):
        ctx.block_put(41_000 + p, 1, 1, kb.FLOAT32, ko.store("raw", ko.F32, v))
    refs = [(41_000 + p, 1) for p in range(len(sizes))]
    prog = kb.Program(ctx, [kb.Leaf(1, kb.FLOAT32, kb.GE, -1000.0)])
    r = ctx.scan(prog, refs, nrows=sizes, aggs=[(1, kb.FLOAT32)])
    g = r["aggs"][0]
    sel = [v[v >= -1000.0] for v in vals]
    allv = np.concatenate(sel)
    assert g.count == allv.size == int(r["counts"].sum())
    got = float(np.uint32(g.sum_bits).view(np.float32))
    exact = float(allv.astype(np.float64).sum())          # multiples of 1/128 below 2^24: exact in float64
    assert got == float(np.float32(exact))                 # correctly rounded exact sum
    st = None
    for v, m in zip(vals, sel):
        st = ko.reduce(ko.F32, v, np.packbits((v >= -1000.0).astype(np.uint8), bitorder="little"), st)
    ref = float(np.uint32(st.sum_bits).view(np.float32))
    assert abs(got - ref) <= 1e-3 * abs(ref)               # the reference's sequential float32 sum drifts; ours does not
    assert np.uint32(g.min_bits).view(np.float32) == allv.min() and np.uint32(g.max_bits).view(np.float32) == allv.max()
    assert (g.min_bits, g.max_bits) == (st.min_bits, st.max_bits)
    prog.close()
    # zone maps
    mins = np.array([[float(v.min())] for v in vals], dtype=np.float32).view(np.uint32).astype(np.uint64)
    maxs = np.array([[float(v.max())] for v in vals], dtype=np.float32).view(np.uint32).astype(np.uint64)
    for mode, a, b in ((kb.GT, float(vals[1].max()), 0), (kb.LE, float(vals[2].min()), 0), (kb.RANGE, 1e7, 2e7), (kb.EQ, float(vals[0][5]), 0)):
        prog = kb.Program(ctx, [kb.Leaf(1, kb.FLOAT32, mode, a, b)])
        bits, n = ctx.prune(prog, mins, maxs)
        want = [bool(ko.lib().ko_match_range(ko.F64, mode, ko.scalar_u64(ko.F64, a), ko.scalar_u64(ko.F64, b), ko.scalar_u64(ko.F64, float(v.min())),
                                             ko.scalar_u64(ko.F64, float(v.max())))) for v in vals]
        got_b = np.unpackbits(bits, bitorder="little")[:len(sizes)].astype(bool).tolist()
        truth = [{kb.GT: v.max() > np.float32(a), kb.LE: v.min() <= np.float32(a), kb.RANGE: v.min() <= np.float32(b) and v.max() >= np.float32(a),
                  kb.EQ: v.min() <= np.float32(a) <= v.max()}[mode] for v in vals]
        assert got_b == [bool(x) for x in truth] and n == sum(got_b), (mode, a, b, got_b, truth, want)
        prog.close()
    for p in range(len(sizes)):
        ctx.block_drop(41_000 + p, 1, 1)
