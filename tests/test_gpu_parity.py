"""GPU tier (-m gpu): the CUDA path behind the C ABI against the CPU oracle, bit for bit.

Every call below goes through libknoxgpu.so's exported C symbols (ctypes), i.e. the same
entry points the cgo adapters bind.  Integer / byte / index results must be bit-exact;
float64 sums must agree with the oracle's sequential sum within 1e-12 relative.
"""
import json
import os

import numpy as np
import pytest

import kxtest as kt
import oracle as ko

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")
RNG = np.random.default_rng(1234)


@pytest.fixture(scope="module")
def ctx():
    import knoxdb_b200 as kb
    c = kb.Context(0)     # raises KnoxError(KX_ENODEV) without a CUDA device: no CPU fallback
    yield c
    c.close()


def typed_src(tname, patterns):
    t = ko.TYPE_BY_NAME[tname]
    if t == ko.F64:
        return np.array(patterns, dtype=np.uint64).view(np.float64)
    if t == ko.F32:
        return np.array(patterns, dtype=np.uint32).view(np.float32)
    return np.array(patterns, dtype=np.int64 if t <= ko.I8 else np.uint64).astype(ko.NP[t])


def from_pattern(t, p):
    if t == ko.F64:
        return np.uint64(p).view(np.float64)
    if t == ko.F32:
        return np.uint32(p).view(np.float32)
    return int(np.uint64(p & (2**64 - 1)).view(np.int64)) if t <= ko.I8 else int(p)


def test_cmp_golden_vectors(ctx):
    """the reference's own known-answer vectors (internal/cmp/tests/<type>.go) through kx_cmp"""
    cmpv = json.load(open(os.path.join(G, "cmp_vectors.json")))["types"]
    ncases = 0
    for tname, ops in cmpv.items():
        t = ko.TYPE_BY_NAME[tname]
        for opname, cases in ops.items():
            for c in cases:
                src = typed_src(tname, c["src"])
                bits, cnt = ctx.cmp(t, ko.OP_BY_NAME[opname], src, from_pattern(t, c["a"]), from_pattern(t, c["b"]))
                assert bits.tobytes().hex() == c["bits"], (tname, opname, c["name"])
                assert cnt == c["count"], (tname, opname, c["name"])
                ncases += 1
    assert ncases == 1476


def test_bitset_golden_vectors(ctx):
    bs = json.load(open(os.path.join(G, "bitset_vectors.json")))
    for c in bs["pop"]:
        buf = np.frombuffer(bytes.fromhex(c["source"]), dtype=np.uint8)
        assert ctx.bitset_popcount(buf, c["size"]) == c["count"], c["name"]
    for c in bs["index"]:
        buf = np.frombuffer(bytes.fromhex(c["buf"]), dtype=np.uint8)
        if c["size"] == 0:
            continue
        assert ctx.bitset_indexes(buf, c["size"]).tolist() == c["idx"], c["name"]


@pytest.mark.parametrize("t", [ko.I64, ko.U64, ko.I32, ko.U32, ko.I16, ko.U16, ko.I8, ko.U8, ko.F64, ko.F32])
def test_cmp_random_vs_oracle(ctx, t):
    for n in (1, 31, 32, 33, 1000, 8191, 8192, 8193, 70001):
        if t in (ko.F64, ko.F32):
            vals = (RNG.integers(0, 2**40, n) / 100.0).astype(ko.NP[t])
            vals[::53] = np.nan
            picks = [(vals[n // 2], vals[n // 2] * 2), (np.nan, 1.0), (0.0, np.inf)]
        else:
            vals = kt.typed_rand(RNG, t, n)
            info = np.iinfo(ko.NP[t])
            picks = [(int(vals[n // 2]), min(info.max, int(vals[n // 2]) + 1000)), (info.min, info.max), (info.max, info.min), (0, 0)]
        for a, b in picks:
            for op in kt.OPS:
                want, wc = ko.cmp(t, op, vals, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                got, gc = ctx.cmp(t, op, vals, a, b)
                assert (got == want).all() and gc == wc, (t, n, op, a, b)


@pytest.mark.parametrize("w", list(range(0, 65)))
def test_bitpack_cmp_all_widths(ctx, w):
    """bitpack.Equal…Between on packed words for every width (reference: bitpack/tests/tests.go:144-298)"""
    L = ko.lib()
    for n in (1, 63, 64, 65, 1025, 8192 + 77, 3 * 8192):
        vals = kt.rnd_bits(RNG, n, w)
        packed = np.zeros(L.ko_bitpack_size(w, n) // 8 + 1, dtype=np.uint64)
        L.ko_bitpack_encode(ko._p(packed), ko._p(vals), n, w, 0)
        top = (1 << w) - 1 if w < 64 else 2**64 - 1
        for a in (int(vals[0]), int(vals[n // 2]), 0, top, min(top + 1, 2**64 - 1)):
            for op in kt.OPS:
                b = min(a + (1 << max(w - 2, 0)), 2**64 - 1)
                want = np.zeros(ko.nbytes(n) + 8, dtype=np.uint8)
                L.ko_bitpack_cmp(op, ko._p(packed), w, a, b, n, ko._p(want))
                got, cnt = ctx.bitpack_cmp(op, packed[:-1] if packed.size > 1 else packed, w, a, b, n)
                assert (got == want[:ko.nbytes(n)]).all(), (w, n, op, a, b)
                assert cnt == int(np.unpackbits(want).sum())
        # decode
        base = 12345
        out = ctx.bitpack_decode(ko.U64, packed, w, base, n)
        assert (out == vals + np.uint64(base)).all(), (w, n)


@pytest.mark.parametrize("t", kt.INT_TYPES)
def test_container_match_and_decode(ctx, t):
    """types.NumberMatcher[T] on every container scheme (EnsureBits: encode/tests/tests.go:140-205)"""
    for n in (1, 67, 1025, 9000):
        for name, vals in kt.shapes(RNG, t, n).items():
            for kind in kt.container_kinds(t, vals):
                blob = ko.store(kind, t, vals)
                oc = ko.Container(t, blob)
                assert (ctx.container_decode(t, blob, n) == vals).all(), (name, kind)
                for a in kt.operands(t, vals)[:5]:
                    b = min(np.iinfo(ko.NP[t]).max, a + 5)
                    for op in kt.OPS:
                        want = oc.match(op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                        got, cnt = ctx.container_match(t, blob, op, a, b, nrows=n)
                        assert (got == want).all(), (ko.NP[t].__name__, n, name, kind, op, a, b)
                        assert cnt == int(np.unpackbits(got).sum())
                setv = np.unique(np.concatenate([vals[: min(3, n)], kt.typed_rand(RNG, t, 3)]))
                su = ko.as_u64(t, setv)
                for neg, op in ((False, ko.IN), (True, ko.NI)):
                    got, _ = ctx.container_match(t, blob, op, values=su, nrows=n)
                    assert (got == oc.match_set(su, negate=neg)).all(), (name, kind, "set", neg)


@pytest.mark.parametrize("t", [ko.U64, ko.I64, ko.I32, ko.U16])
def test_in_sets_of_every_size_on_packed_and_raw_blocks(ctx, t):
    """MatchInSet / MatchNotInSet on bit-packed and raw blocks (int_bitpack.go:249-291, int_raw.go:339-380) through the
    prefilter + hash-table path: set sizes that keep the exact table in shared memory and sizes that leave it in
    global memory, members and non-members, tiles with several passes per warp."""
    n = 150_001
    info = np.iinfo(ko.NP[t])
    span = min(int(info.max) - 1000, 1 << 38)
    lo = max(int(info.min), -span // 2) if info.min < 0 else 0
    vals = RNG.integers(lo, lo + span, n, dtype=np.int64).astype(ko.NP[t])
    for kind in ("bitpack", "raw"):
        blob = ko.store(kind, t, vals)
        oc = ko.Container(t, blob)
        for nset in (1, 5, 64, 700, 5000, 40000):
            members = RNG.choice(vals, min(nset, n) // 2 + 1)
            others = kt.typed_rand(RNG, t, nset // 2 + 1)
            su = ko.as_u64(t, np.unique(np.concatenate([members, others])))
            for neg, op in ((False, ko.IN), (True, ko.NI)):
                got, cnt = ctx.container_match(t, blob, op, values=su, nrows=n)
                want = oc.match_set(su, negate=neg)
                assert (got == want).all(), (kind, nset, neg)
                assert cnt == int(np.unpackbits(want).sum())


def test_run_end_blocks_of_every_run_length(ctx):
    """RunEndContainer.Match* + applyMatch (int_runend.go:224-318) for runs shorter than a bitset word, runs that span
    many words, and both mixed in one block (the run-fill pre-pass sets every matching run's row range: atomics at the
    edge words, plain stores in between); row counts around word boundaries."""
    rng = np.random.default_rng(21)
    for n, lens in ((100_003, (1, 4)), (100_003, (1, 70)), (250_000, (60, 3000)), (70_001, (1, 1)), (4096, (5000, 5001)), (33, (1, 3)), (1, (1, 1))):
        chunks, total = [], 0
        while total < n:
            ln = int(rng.integers(lens[0], lens[1] + 1))
            chunks.append(np.full(ln, int(rng.integers(-50, 50)), dtype=np.int64))
            total += ln
        vals = np.concatenate(chunks)[:n]
        if n < 2:
            continue
        blob = ko.store("runend", ko.I64, vals)
        oc = ko.Container(ko.I64, blob)
        for op, a, b in ((ko.EQ, 7, 0), (ko.NE, 7, 0), (ko.LT, 0, 0), (ko.GE, -10, 0), (ko.RG, -5, 5), (ko.GT, 1000, 0), (ko.LE, 1000, 0)):
            want = oc.match(op, ko.scalar_u64(ko.I64, a), ko.scalar_u64(ko.I64, b))
            got, cnt = ctx.container_match(ko.I64, blob, op, a, b, nrows=n)
            assert (got == want).all(), (n, lens, op)
            assert cnt == int(np.unpackbits(want).sum())
        su = ko.as_u64(ko.I64, np.array([-3, 7, 11, 49], dtype=np.int64))
        got, _ = ctx.container_match(ko.I64, blob, ko.IN, values=su, nrows=n)
        assert (got == oc.match_set(su)).all(), (n, lens, "in")


def test_run_end_blocks_with_affine_run_values_follow_the_delta_matchers(ctx):
    """A run-end block whose Values child is a DeltaContainer (block heights with several rows each): the reference
    matches the child with DeltaContainer's closed-form index arithmetic (int_runend.go:224-283 → int_delta.go:149-449),
    including MatchBetween's rounding quirk for ranges that start below For between grid points.  The product does the
    same arithmetic over the RUNS (LM_RUNRANGE → run-fill pre-pass); bit for bit against the oracle, through the narrow
    drop-in and through a two-leaf scan with aggregates."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(33)
    for t, base, delta, nruns in ((ko.I64, 100, 10, 23), (ko.U64, 1000, 3, 4000), (ko.I32, -50, 4, 700), (ko.I64, -10**12, 977, 9000)):
        runs = (base + delta * np.arange(nruns)).astype(ko.NP[t])
        vals = np.repeat(runs, rng.integers(1, 70, nruns))
        n = len(vals)
        blob = ko.store("runend", t, vals)
        oc = ko.Container(t, blob)
        assert oc.ctype == ko.TRUNEND and oc.value_delta_sequences()
        lo_all, hi_all = int(runs.min()), int(runs.max())
        quirk_hits = 0
        for a in np.unique(np.r_[lo_all - 2 * delta - 1, lo_all - 1, lo_all, rng.integers(lo_all - 3 * delta, hi_all + 3 * delta, 12)]):
            a = int(a)
            if t == ko.U64 and a < 0:
                continue
            for b in (a, a + 5, a + 37 * delta + 1, hi_all + 50):
                for op in kt.OPS:
                    want = oc.match(op, ko.scalar_u64(t, a), ko.scalar_u64(t, b))
                    got, cnt = ctx.container_match(t, blob, op, a, b, nrows=n)
                    assert (got == want).all(), (ko.NP[t].__name__, delta, op, a, b)
                    assert cnt == int(np.unpackbits(want).sum())
                    if op == ko.RG and not (want == kt.pack_bits(kt.OPS[op](vals, ko.NP[t](a), ko.NP[t](b)))).all():
                        quirk_hits += 1
        assert quirk_hits > 0
    # the same block as one leaf of a scan: height BETWEEN (quirk domain) AND amount < 0 → count + sum(amount)
    t, base, delta, nruns = ko.I64, 100, 10, 5000
    vals = np.repeat((base + delta * np.arange(nruns)).astype(np.int64), rng.integers(1, 40, nruns))
    n = len(vals)
    amount = rng.integers(-10**6, 10**6, n).astype(np.int64)
    hb = ko.store("runend", ko.I64, vals)
    ctx.block_put(950, 1, 1, kb.INT64, hb)
    ctx.block_put(950, 1, 2, kb.INT64, ko.store("best", ko.I64, amount))
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, 95, 30_004), kb.Leaf(2, kb.INT64, kb.LT, 0)])
    res = ctx.scan(prog, [(950, 1)], nrows=[n], want_bitsets=True, aggs=[(2, kb.INT64)])
    want = ko.tree_eval([0, 1, 0xFE], [ko.Container(ko.I64, hb).match(ko.RG, 95, 30_004), kt.pack_bits(amount < 0)], n)
    assert (res["bitsets"][0] == want).all()
    st = ko.reduce(ko.I64, amount, want, None)
    g = res["aggs"][0]
    assert (g.count, g.sum_bits, g.min_bits, g.max_bits) == (st.count, st.sum_bits, st.min_bits, st.max_bits) and st.count > 100
    assert not kt.unpack_bits(want, n)[0] and vals[0] == 100   # the quirk: row 0 (height 100 >= 95) is dropped, as in the reference
    prog.close()
    ctx.block_drop(950, 1, 1); ctx.block_drop(950, 1, 2)


def test_bitset_ops(ctx):
    import knoxdb_b200 as kb
    L = ko.lib()
    import ctypes as C
    for size in (1, 7, 8, 9, 63, 64, 65, 1000, 4099, 300007):
        l = ko.nbytes(size)
        a = RNG.integers(0, 256, l, dtype=np.uint8)
        b = RNG.integers(0, 256, l, dtype=np.uint8)
        for op, name, flag in ((0, "ko_bitset_and_flag", True), (1, "ko_bitset_andnot", False), (2, "ko_bitset_or_flag", True), (3, "ko_bitset_xor", False)):
            d = a.copy()
            any_, all_ = C.c_int(), C.c_int()
            if flag:
                getattr(L, name)(ko._p(d), ko._p(b), size, C.byref(any_), C.byref(all_))
            else:
                getattr(L, name)(ko._p(d), ko._p(b), size)
            got, gany, gall = ctx.bitset_op(op, a, b, size)
            assert (got == d).all(), (name, size)
            if flag:
                assert gany == bool(any_.value) and gall == bool(all_.value), (name, size)
        d = a.copy(); L.ko_bitset_neg(ko._p(d), size)
        assert (ctx.bitset_neg(a, size) == d).all()
        assert ctx.bitset_popcount(a, size) == L.ko_bitset_popcount(ko._p(a), size)
        idx = np.zeros(size + 8, dtype=np.uint32)
        k = L.ko_bitset_indexes(ko._p(a), size, ko._p(idx))
        assert (ctx.bitset_indexes(a, size) == idx[:k]).all()


def make_table(rng, npacks, nrows_list):
    """3-column packs of BASELINE config 3: ts (sorted int64), acct (uint64 dups), amount (int64 + float64)"""
    packs = []
    t0 = 1_700_000_000
    accts = kt.rnd_bits(rng, 500, 40)
    for p in range(npacks):
        n = nrows_list[p]
        ts = (t0 + np.cumsum(rng.integers(0, 3, n))).astype(np.int64)
        t0 = int(ts[-1]) + 1 if n else t0
        acct = rng.choice(accts, n).astype(np.uint64)
        amount = rng.integers(-10**9, 10**9, n).astype(np.int64)
        famount = (rng.integers(0, 2**50, n) / 100.0).astype(np.float64)
        packs.append({"ts": ts, "acct": acct, "amount": amount, "famount": famount})
    return packs, accts


F_TS, F_ACCT, F_AMT, F_FAMT = 1, 2, 3, 4


def put_table(ctx, packs, acct_kind="dict"):
    blobs = []
    for p, cols in enumerate(packs):
        enc = {F_TS: ko.store("best", ko.I64, cols["ts"]), F_ACCT: ko.store(acct_kind, ko.U64, cols["acct"]),
               F_AMT: ko.store("best", ko.I64, cols["amount"]), F_FAMT: ko.store("raw", ko.F64, cols["famount"])}
        for f, t in ((F_TS, ko.I64), (F_ACCT, ko.U64), (F_AMT, ko.I64), (F_FAMT, ko.F64)):
            assert ctx.block_put(p, 1, f, t, enc[f]) == cols["ts"].size
        blobs.append(enc)
    return blobs


def oracle_query(packs, blobs, t_lo, t_hi, set_u64, postfix):
    counts, bitsets = [], []
    agg_i, agg_f = ko.Agg(), ko.Agg()
    for cols, enc in zip(packs, blobs):
        n = cols["ts"].size
        l0 = ko.Container(ko.I64, enc[F_TS]).match(ko.RG, ko.scalar_u64(ko.I64, t_lo), ko.scalar_u64(ko.I64, t_hi))
        l1 = ko.Container(ko.U64, enc[F_ACCT]).match_set(set_u64)
        bits = ko.tree_eval(postfix, [l0, l1], n)
        bitsets.append(bits)
        counts.append(int(np.unpackbits(bits).sum()))
        agg_i = ko.reduce(ko.I64, cols["amount"], bits, agg_i)
        agg_f = ko.reduce(ko.F64, cols["famount"], bits, agg_f)
    return counts, bitsets, agg_i, agg_f


def _random_tree(rng, nleaves):
    """random binary AND/OR tree over leaves 0..nleaves-1 in postfix form (leaf order shuffled: left-deep, right-deep
    and bushy shapes all occur, i.e. every stack depth from 2 to nleaves)"""
    import knoxdb_b200 as kb
    items = [[int(i)] for i in rng.permutation(nleaves)]
    while len(items) > 1:
        i = int(rng.integers(0, len(items) - 1))
        if rng.random() < 0.35:   # right-deep step: keeps operands on the stack
            i = len(items) - 2
        a, b = items[i], items[i + 1]
        items[i:i + 2] = [a + b + [kb.OP_AND if rng.random() < 0.5 else kb.OP_OR]]
    return items[0]


def test_predicate_trees_of_every_shape_over_every_container(ctx):
    """filter.Match on arbitrary AND/OR trees (match_core.go:14-215): eight leaves over eight different containers
    (bit-packed, dictionary, run-end, affine, constant, raw, raw float, ALP float), random tree shapes with stack
    depths up to 8, ragged packs (1 row … several tiles), with and without aggregates."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(77)
    nrows = [1, 31, 32, 33, 4097, 70_003, 8192, 140_000]
    cols, blobs = [], []
    T = [(1, ko.I64, kb.INT64), (2, ko.U64, kb.UINT64), (3, ko.I32, kb.INT32), (4, ko.U16, kb.UINT16), (5, ko.I8, kb.INT8),
         (6, ko.U64, kb.UINT64), (7, ko.F64, kb.FLOAT64), (8, ko.F64, kb.FLOAT64)]
    accts = kt.rnd_bits(rng, 300, 40)
    for p, n in enumerate(nrows):
        c = {1: (1_700_000_000 + np.cumsum(rng.integers(0, 3, n))).astype(np.int64),
             2: rng.choice(accts, n).astype(np.uint64),
             3: np.repeat(rng.integers(-1000, 1000, n // 7 + 1), 7)[:n].astype(np.int32),
             4: (100 + 3 * np.arange(n)).astype(np.uint16) if 100 + 3 * n < 65535 else (np.arange(n) % 60000).astype(np.uint16),
             5: np.full(n, -5, dtype=np.int8),
             6: kt.rnd_bits(rng, n, 50),
             7: (rng.integers(0, 2**40, n) / 100.0).astype(np.float64),
             8: np.round(rng.uniform(0, 500, n), 2)}
        e = {1: ko.store("best", ko.I64, c[1]), 2: ko.store("dict" if n >= 2 else "raw", ko.U64, c[2]),
             3: ko.store("runend" if n >= 2 else "raw", ko.I32, c[3]),
             4: ko.store("delta", ko.U16, base=100, delta=3, n=n) if 100 + 3 * n < 65535 else ko.store("bitpack", ko.U16, c[4]),
             5: ko.store("const", ko.I8, val=-5, n=n), 6: ko.store("raw", ko.U64, c[6]), 7: ko.store("raw", ko.F64, c[7]),
             8: ko.store("alp", ko.F64, c[8])}
        for f, kot, kbt in T:
            assert ctx.block_put(p, 1, f, kbt, e[f]) == n
        cols.append(c); blobs.append(e)
    big = cols[5]
    setv = rng.choice(accts, 40, replace=False)
    leaf_defs = [
        (kb.Leaf(1, kb.INT64, kb.RANGE, int(big[1][1000]), int(big[1][40000])), lambda oc: oc.match(ko.RG, ko.scalar_u64(ko.I64, int(big[1][1000])), ko.scalar_u64(ko.I64, int(big[1][40000])))),
        (kb.Leaf(2, kb.UINT64, kb.IN, values=setv), lambda oc: oc.match_set(setv)),
        (kb.Leaf(3, kb.INT32, kb.GT, 100), lambda oc: oc.match(ko.GT, ko.scalar_u64(ko.I32, 100), 0)),
        (kb.Leaf(4, kb.UINT16, kb.LE, 30000), lambda oc: oc.match(ko.LE, ko.scalar_u64(ko.U16, 30000), 0)),
        (kb.Leaf(5, kb.INT8, kb.NE, 7), lambda oc: oc.match(ko.NE, ko.scalar_u64(ko.I8, 7), 0)),
        (kb.Leaf(6, kb.UINT64, kb.LT, 1 << 49), lambda oc: oc.match(ko.LT, ko.scalar_u64(ko.U64, 1 << 49), 0)),
        (kb.Leaf(7, kb.FLOAT64, kb.GE, 2**39 / 100.0), lambda oc: oc.match(ko.GE, ko.scalar_u64(ko.F64, 2**39 / 100.0), 0)),
        (kb.Leaf(8, kb.FLOAT64, kb.LT, 250.0), lambda oc: oc.match(ko.LT, ko.scalar_u64(ko.F64, 250.0), 0)),
    ]
    kot_of = {f: kot for f, kot, _ in T}
    leaf_bits = [[fn(ko.Container(kot_of[lf.field], blobs[p][lf.field])) for lf, fn in leaf_defs] for p in range(len(nrows))]
    refs = [(p, 1) for p in range(len(nrows))]
    depths = set()
    for trial in range(24):
        nl = int(rng.integers(2, 9))
        pick = [int(i) for i in rng.permutation(8)[:nl]]
        pf = _random_tree(rng, nl)
        d = sp = 0
        for op in pf:
            sp += 1 if op < 0x80 else -1
            d = max(d, sp)
        depths.add(d)
        prog = kb.Program(ctx, [leaf_defs[i][0] for i in pick], pf)
        aggs = [(6, kb.UINT64), (7, kb.FLOAT64)] if trial % 2 else []
        res = ctx.scan(prog, refs, nrows=nrows, want_bitsets=True, aggs=aggs)
        st_u = st_f = None
        for p, n in enumerate(nrows):
            want = ko.tree_eval(pf, [leaf_bits[p][i] for i in pick], n)
            assert (res["bitsets"][p] == want).all(), (trial, pf, p)
            assert int(res["counts"][p]) == int(np.unpackbits(want).sum())
            st_u = ko.reduce(ko.U64, cols[p][6], want, st_u)
            st_f = ko.reduce(ko.F64, cols[p][7], want, st_f)
        if aggs:
            gu, gf = res["aggs"]
            assert (gu.count, gu.sum_bits, gu.min_bits, gu.max_bits) == (st_u.count, st_u.sum_bits, st_u.min_bits, st_u.max_bits), (trial, pf)
            if st_f.valid:
                want_sum = float(np.uint64(st_f.sum_bits).view(np.float64))
                assert abs(gf.value("sum", kb.FLOAT64) - want_sum) <= 1e-12 * abs(want_sum)
                assert (gf.min_bits, gf.max_bits) == (st_f.min_bits, st_f.max_bits)
        prog.close()
    assert max(depths) >= 5 and min(depths) == 2, depths
    for p in range(len(nrows)):
        for f, _, _ in T:
            ctx.block_drop(p, 1, f)


@pytest.fixture
def agg_stage(request, monkeypatch):
    """KX_AGG_STAGE: how the fused reduce reads value columns (on demand / staged through the ring / by selectivity)"""
    mode = getattr(request, "param", "auto")
    if mode == "auto":
        monkeypatch.delenv("KX_AGG_STAGE", raising=False)
    else:
        monkeypatch.setenv("KX_AGG_STAGE", mode)
    return mode


@pytest.mark.parametrize("agg_stage", ["auto", "always", "never"], indirect=True)
@pytest.mark.parametrize("acct_kind", ["dict", "bitpack", "raw"])
def test_multi_predicate_scan_with_aggregates(ctx, acct_kind, agg_stage):
    """BASELINE config 3: ts BETWEEN AND acct IN {…} → count/sum/min/max over int64 and float64"""
    import knoxdb_b200 as kb
    nrows = [4096, 70000, 8192, 1, 33333, 16384]
    packs, accts = make_table(RNG, len(nrows), nrows)
    blobs = put_table(ctx, packs, acct_kind)
    all_ts = np.concatenate([c["ts"] for c in packs])
    for sel in (0.001, 0.1, 0.9):
        t_lo = int(all_ts[int(all_ts.size * 0.05)])
        t_hi = int(all_ts[min(all_ts.size - 1, int(all_ts.size * (0.05 + sel)))])
        setv = RNG.choice(accts, 64, replace=False)
        for postfix in ([0, 1, kb.OP_AND], [0, 1, kb.OP_OR]):
            prog = kb.Program(ctx, [kb.Leaf(F_TS, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(F_ACCT, kb.UINT64, kb.IN, values=setv)], postfix)
            res = ctx.scan(prog, [(p, 1) for p in range(len(nrows))], nrows=nrows, want_bitsets=True,
                           aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64)])
            counts, bitsets, agg_i, agg_f = oracle_query(packs, blobs, t_lo, t_hi, setv, postfix)
            assert res["counts"].tolist() == counts
            for got, want in zip(res["bitsets"], bitsets):
                assert (got == want).all()
            gi, gf = res["aggs"]
            assert gi.count == agg_i.count and gf.count == agg_f.count
            assert bool(gi.valid) == bool(agg_i.valid)
            if agg_i.valid:
                assert gi.sum_bits == agg_i.sum_bits and gi.min_bits == agg_i.min_bits and gi.max_bits == agg_i.max_bits
                want_sum = np.uint64(agg_f.sum_bits).view(np.float64)
                got_sum = np.uint64(gf.sum_bits).view(np.float64)
                assert abs(got_sum - want_sum) <= 1e-12 * abs(want_sum)          # north-star tolerance
                assert gf.min_bits == agg_f.min_bits and gf.max_bits == agg_f.max_bits
            prog.close()
    # the same query over HOST blocks (cold device cache path)
    prog = kb.Program(ctx, [kb.Leaf(F_TS, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(F_ACCT, kb.UINT64, kb.IN, values=setv)])
    fields = [(F_TS, kb.INT64), (F_ACCT, kb.UINT64), (F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64)]
    hb = [[np.frombuffer(enc[f], np.uint8) for f, _ in fields] for enc in blobs]
    res_h = ctx.scan_host(prog, fields, hb, nrows=nrows, want_bitsets=True, aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64)])
    res_d = ctx.scan(prog, [(p, 1) for p in range(len(nrows))], nrows=nrows, want_bitsets=True, aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64)])
    assert res_h["counts"].tolist() == res_d["counts"].tolist()
    for x, y in zip(res_h["bitsets"], res_d["bitsets"]):
        assert (x == y).all()
    for x, y in zip(res_h["aggs"], res_d["aggs"]):
        assert (x.count, x.sum_bits, x.min_bits, x.max_bits) == (y.count, y.sum_bits, y.min_bits, y.max_bits)
    for p in range(len(nrows)):
        for f in (F_TS, F_ACCT, F_AMT, F_FAMT):
            ctx.block_drop(p, 1, f)
    assert ctx.store_stats()["blocks"] == 0


def test_staged_and_on_demand_reduce_agree_bit_for_bit(ctx, monkeypatch):
    """The producer decides per tile, from the selectivity it has seen so far, whether a value column is staged
    through the ring or read on demand — a timing-dependent choice.  Both paths must therefore give the SAME bits,
    float sums included (same lane/row assignment and order of additions), on a scan long enough (many tiles per
    CTA) for the choice to flip inside one launch."""
    import knoxdb_b200 as kb
    n, npacks = 300_000, 96
    base, accts = make_table(RNG, 2, [n, n])
    base[1]["ts"] = base[0]["ts"]   # same time range in both blocks: the selectivity is uniform over the scan
    encs = []
    for cols in base:
        encs.append({F_TS: ko.store("best", ko.I64, cols["ts"]), F_AMT: ko.store("best", ko.I64, cols["amount"]),
                     F_FAMT: ko.store("raw", ko.F64, cols["famount"]), F_ACCT: ko.store("dict", ko.U64, cols["acct"])})
    for p in range(npacks):
        for f, t in ((F_TS, ko.I64), (F_AMT, ko.I64), (F_FAMT, ko.F64), (F_ACCT, ko.U64)):
            ctx.block_put(p, 1, f, t, encs[p % 2][f])
    ts = base[0]["ts"]
    for sel in (0.02, 0.5, 0.95):
        t_lo, t_hi = int(ts[int(n * 0.01)]), int(ts[int(n * (0.01 + sel))])
        prog = kb.Program(ctx, [kb.Leaf(F_TS, kb.INT64, kb.RANGE, t_lo, t_hi)])
        got = {}
        for mode in ("never", "always", "3", "auto"):
            if mode == "auto":
                monkeypatch.delenv("KX_AGG_STAGE", raising=False)
            else:
                monkeypatch.setenv("KX_AGG_STAGE", mode)
            r = ctx.scan(prog, [(p, 1) for p in range(npacks)], nrows=[n] * npacks,
                         aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64), (F_ACCT, kb.UINT64)])
            got[mode] = (r["counts"].tolist(), [(g.count, int(bool(g.valid)), g.sum_bits, g.min_bits, g.max_bits) for g in r["aggs"]])
        prog.close()
        assert got["never"] == got["always"] == got["3"] == got["auto"], sel
        # and the oracle: integer aggregates bit for bit, float sum within the north-star tolerance of the sequential sum
        agg_i, agg_f, agg_a = ko.Agg(), ko.Agg(), ko.Agg()
        for p in range(npacks):
            cols = base[p % 2]
            bits = ko.Container(ko.I64, encs[p % 2][F_TS]).match(ko.RG, ko.scalar_u64(ko.I64, t_lo), ko.scalar_u64(ko.I64, t_hi))
            agg_i = ko.reduce(ko.I64, cols["amount"], bits, agg_i)
            agg_f = ko.reduce(ko.F64, cols["famount"], bits, agg_f)
            agg_a = ko.reduce(ko.U64, cols["acct"], bits, agg_a)
        gi, gf, ga = got["auto"][1]
        assert gi == (agg_i.count, 1, agg_i.sum_bits, agg_i.min_bits, agg_i.max_bits)
        assert ga == (agg_a.count, 1, agg_a.sum_bits, agg_a.min_bits, agg_a.max_bits)
        want = np.uint64(agg_f.sum_bits).view(np.float64)
        assert abs(np.uint64(gf[2]).view(np.float64) - want) <= 1e-12 * abs(want)
        assert gf[3:] == (agg_f.min_bits, agg_f.max_bits)
    monkeypatch.delenv("KX_AGG_STAGE", raising=False)
    for p in range(npacks):
        for f in (F_TS, F_AMT, F_FAMT, F_ACCT):
            ctx.block_drop(p, 1, f)


@pytest.mark.parametrize("ordered", [True, False])
def test_time_bucketed_reduce(ctx, ordered):
    """series query shape (pkg/series/series.go:192-256): filter, then count / sum / min / max per time window.
    Window starts follow TimeUnit.Next (first window starts at From, the following ones at aligned multiples of the
    step); integer results bit-exact, float64 sums within 1e-12 of the oracle's sequential per-window sum."""
    import knoxdb_b200 as kb
    nrows = [50_000, 70_001, 1, 33_333, 2048]
    packs, accts = make_table(RNG, len(nrows), nrows)
    if not ordered:   # journal-like packs: rows not in time order
        for cols in packs:
            perm = RNG.permutation(cols["ts"].size)
            for k in cols:
                cols[k] = cols[k][perm]
    blobs = put_table(ctx, packs, "dict")
    all_ts = np.concatenate([c["ts"] for c in packs])
    t_from, t_to = int(np.sort(all_ts)[all_ts.size // 20]) + 7, int(np.sort(all_ts)[all_ts.size * 9 // 10])
    refs = [(p, 1) for p in range(len(nrows))]
    for step in (60, 3600, 7 * 3600):
        edges = ko.window_edges(t_from, t_to, step)
        assert edges[0] == t_from and edges[-1] >= t_to and (np.diff(edges) > 0).all()
        setv = RNG.choice(accts, 200, replace=False)
        prog = kb.Program(ctx, [kb.Leaf(F_TS, kb.INT64, kb.RANGE, t_from, t_to - 1), kb.Leaf(F_ACCT, kb.UINT64, kb.IN, values=setv)])
        res = ctx.scan_buckets(prog, refs, F_TS, kb.INT64, edges, aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64), (F_ACCT, kb.UINT64)])
        nb = edges.size - 1
        st_i = st_f = st_a = None
        counts = []
        for cols, enc in zip(packs, blobs):
            l0 = ko.Container(ko.I64, enc[F_TS]).match(ko.RG, ko.scalar_u64(ko.I64, t_from), ko.scalar_u64(ko.I64, t_to - 1))
            l1 = ko.Container(ko.U64, enc[F_ACCT]).match_set(setv)
            bits = ko.tree_eval([0, 1, 0xFE], [l0, l1], cols["ts"].size)
            counts.append(int(np.unpackbits(bits).sum()))
            st_i = ko.bucket_reduce(ko.I64, cols["amount"], ko.I64, cols["ts"], bits, edges, st_i)
            st_f = ko.bucket_reduce(ko.F64, cols["famount"], ko.I64, cols["ts"], bits, edges, st_f)
            st_a = ko.bucket_reduce(ko.U64, cols["acct"], ko.I64, cols["ts"], bits, edges, st_a)
        assert res["counts"].tolist() == counts
        assert res["bucket_counts"].tolist() == [st_i[k].count for k in range(nb)]
        assert sum(counts) == int(res["bucket_counts"].sum())   # the range leaf keeps every match inside [From, To)
        for k in range(nb):
            gi, gf, ga = res["aggs"][0][k], res["aggs"][1][k], res["aggs"][2][k]
            assert (gi.count, bool(gi.valid)) == (st_i[k].count, bool(st_i[k].valid)), (step, k)
            if not st_i[k].valid:
                continue
            assert (gi.sum_bits, gi.min_bits, gi.max_bits) == (st_i[k].sum_bits, st_i[k].min_bits, st_i[k].max_bits), (step, k)
            assert (ga.sum_bits, ga.min_bits, ga.max_bits) == (st_a[k].sum_bits, st_a[k].min_bits, st_a[k].max_bits), (step, k)
            want = float(np.uint64(st_f[k].sum_bits).view(np.float64))
            assert abs(gf.value("sum", kb.FLOAT64) - want) <= 1e-12 * abs(want), (step, k)
            assert (gf.min_bits, gf.max_bits) == (st_f[k].min_bits, st_f[k].max_bits), (step, k)
        prog.close()
    # windows that do not cover every match: rows outside [edges[0], edges[-1]) belong to no window
    prog = kb.Program(ctx, [kb.Leaf(F_AMT, kb.INT64, kb.GT, 0)])
    edges = ko.window_edges(t_from, t_from + 5000, 600)
    res = ctx.scan_buckets(prog, refs, F_TS, kb.INT64, edges, aggs=[(F_AMT, kb.INT64)])
    st = None
    for cols, enc in zip(packs, blobs):
        bits = ko.Container(ko.I64, enc[F_AMT]).match(ko.GT, ko.scalar_u64(ko.I64, 0), 0)
        st = ko.bucket_reduce(ko.I64, cols["amount"], ko.I64, cols["ts"], bits, edges, st)
    assert [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in res["aggs"][0]] == [(s.count, s.sum_bits, s.min_bits, s.max_bits) for s in st]
    assert int(res["bucket_counts"].sum()) < int(res["counts"].sum())
    # the kernel has one instantiation per number of value columns (1, 2, 4): two columns and none must agree with the above
    res2 = ctx.scan_buckets(prog, refs, F_TS, kb.INT64, edges, aggs=[(F_AMT, kb.INT64), (F_ACCT, kb.UINT64)])
    res0 = ctx.scan_buckets(prog, refs, F_TS, kb.INT64, edges)
    assert res2["bucket_counts"].tolist() == res["bucket_counts"].tolist() == res0["bucket_counts"].tolist()
    assert [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in res2["aggs"][0]] == [(s.count, s.sum_bits, s.min_bits, s.max_bits) for s in st]
    sta = None
    for cols, enc in zip(packs, blobs):
        bits = ko.Container(ko.I64, enc[F_AMT]).match(ko.GT, ko.scalar_u64(ko.I64, 0), 0)
        sta = ko.bucket_reduce(ko.U64, cols["acct"], ko.I64, cols["ts"], bits, edges, sta)
    assert [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in res2["aggs"][1]] == [(s.count, s.sum_bits, s.min_bits, s.max_bits) for s in sta]
    prog.close()
    for p in range(len(nrows)):
        for f in (F_TS, F_ACCT, F_AMT, F_FAMT):
            ctx.block_drop(p, 1, f)


def _string_rows(rng, n, shape):
    if shape == "fixed20":      # config 4's address column: random 20-byte strings
        return [bytes(r) for r in rng.integers(0, 256, (n, 20), dtype=np.uint8)]
    if shape == "dups":         # few distinct values of different lengths (→ dictionary)
        vocab = [b"", b"a", b"ab", b"abc", b"abd", b"tz1VSUr8wwNhLAzempoch5d6hLRiTh8Cjcjb", b"zzzz", bytes(range(256))]
        return [vocab[i] for i in rng.integers(0, len(vocab), n)]
    if shape == "ragged":       # variable lengths 0..40, small alphabet (many shared prefixes)
        return [bytes(rng.integers(97, 100, int(k), dtype=np.uint8)) for k in rng.integers(0, 41, n)]
    return [b"same value"] * n  # constant


@pytest.mark.parametrize("shape", ["fixed20", "dups", "ragged", "const"])
def test_string_in_sets(ctx, shape):
    """bytesInSetMatcher / bytesNotInSetMatcher (internal/operator/filter/match_bytes.go:392-520: a row matches iff its bytes
    are a member of the de-duplicated set): sets of 1 … 300 strings with members, non-members, duplicates, the empty string
    and proper prefixes, on every string container; the truth is Python's own bytes membership on the oracle's rows."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(12)
    for n in (1, 33, 1000, 20_011):
        rows = _string_rows(rng, n, shape)
        kinds = [k for k in (ko.STR_CONST, ko.STR_FIXED, ko.STR_COMPACT, ko.STR_DICT) if ko.store_str(k, rows) is not None]
        if n > 5000:
            kinds = [k for k in kinds if k != ko.STR_DICT or shape in ("dups", "const")]
        sets = [[rows[0]], [b""], [rows[n // 2], rows[n // 2], b"zz-not-there"], [rows[-1][:-1], rows[-1] + b"\x00"],
                [rows[i] for i in rng.integers(0, n, 40)] + [b"", b"ab"],
                [bytes(rng.integers(0, 256, int(k), dtype=np.uint8)) for k in rng.integers(0, 30, 300)] + [rows[n // 3]]]
        for kind in kinds:
            blob = ko.store_str(kind, rows)
            oc = ko.StrContainer(blob)
            assert ctx.block_put(320, 1, 9, kb.BYTES, blob) == n
            got_rows = [oc.get(i) for i in range(min(n, 50))]
            assert got_rows == rows[: len(got_rows)]
            for members in sets:
                ms = set(members)
                truth = np.fromiter((r in ms for r in rows), dtype=bool, count=n)
                for mode, want_bool in ((kb.IN, truth), (kb.NIN, ~truth)):
                    prog = kb.Program(ctx, [kb.Leaf(9, kb.BYTES, mode, values=members)])
                    res = ctx.scan(prog, [(320, 1)], nrows=[n], want_bitsets=True)
                    assert (res["bitsets"][0] == kt.pack_bits(want_bool)).all(), (shape, n, kind, mode, len(members))
                    assert int(res["counts"][0]) == int(want_bool.sum())
                    prog.close()
            ctx.block_drop(320, 1, 9)


@pytest.mark.parametrize("shape", ["fixed20", "dups", "ragged", "const"])
def test_string_blocks_match_row_by_row(ctx, shape):
    """types.StringMatcher on the string containers (internal/encode/string_{const,fixed,compact,dict}.go, matchers
    string_match.go:13-188): the seven modes, operands that are members, non-members, prefixes and the empty string,
    every container that can hold the rows, sizes around the 32-row word and tile boundaries; then a string leaf
    ANDed with an integer leaf (config 4's `height BETWEEN … AND address = X` at row level)."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(11)
    for n in (1, 31, 33, 1000, 70_003):
        rows = _string_rows(rng, n, shape)
        kinds = [k for k in (ko.STR_CONST, ko.STR_FIXED, ko.STR_COMPACT, ko.STR_DICT) if ko.store_str(k, rows) is not None]
        if n > 5000:
            kinds = [k for k in kinds if k != ko.STR_DICT or shape in ("dups", "const")]   # the oracle's dictionary build is quadratic
        assert ko.STR_COMPACT in kinds
        operands = [rows[0], rows[n // 2], b"", b"ab", rows[-1] + b"\x00", b"\xff" * 3]
        for kind in kinds:
            blob = ko.store_str(kind, rows)
            oc = ko.StrContainer(blob)
            assert oc.n == n and oc.get(n // 2) == rows[n // 2]
            assert ctx.block_put(300, 1, 9, kb.BYTES, blob) == n
            for a in operands:
                for mode, kom in ((kb.EQ, ko.EQ), (kb.NE, ko.NE), (kb.LT, ko.LT), (kb.LE, ko.LE), (kb.GT, ko.GT), (kb.GE, ko.GE)):
                    prog = kb.Program(ctx, [kb.Leaf(9, kb.BYTES, mode, a)])
                    res = ctx.scan(prog, [(300, 1)], nrows=[n], want_bitsets=True)
                    want = oc.match(kom, a)
                    assert (res["bitsets"][0] == want).all(), (shape, n, kind, mode, a)
                    assert int(res["counts"][0]) == int(np.unpackbits(want).sum())
                    prog.close()
            for lo, hi in ((b"ab", b"abd"), (rows[0], rows[0]), (b"", b"\xff"), (b"b", b"a")):
                prog = kb.Program(ctx, [kb.Leaf(9, kb.BYTES, kb.RANGE, lo, hi)])
                res = ctx.scan(prog, [(300, 1)], nrows=[n], want_bitsets=True)
                assert (res["bitsets"][0] == oc.match(ko.RG, lo, hi)).all(), (shape, n, kind, "range", lo, hi)
                prog.close()
            ctx.block_drop(300, 1, 9)
    # row-level `height BETWEEN lo AND hi AND address = X` over several packs, aggregate over a value column
    nrows = [5000, 70_001, 64]
    packs = []
    for p, n in enumerate(nrows):
        rows = _string_rows(rng, n, shape)
        height = (1000 * p + np.arange(n)).astype(np.int64)
        amount = rng.integers(-10**6, 10**6, n).astype(np.int64)
        blob = ko.store_str(ko.STR_COMPACT if shape != "fixed20" else ko.STR_FIXED, rows)
        ctx.block_put(310 + p, 1, 1, kb.INT64, ko.store("best", ko.I64, height))
        ctx.block_put(310 + p, 1, 2, kb.BYTES, blob)
        ctx.block_put(310 + p, 1, 3, kb.INT64, ko.store("best", ko.I64, amount))
        packs.append((rows, height, amount, blob))
    x = packs[1][0][777]
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, 500, 60_000), kb.Leaf(2, kb.BYTES, kb.EQ, x)])
    res = ctx.scan(prog, [(310 + p, 1) for p in range(3)], nrows=nrows, want_bitsets=True, aggs=[(3, kb.INT64)])
    st = None
    for p, (rows, height, amount, blob) in enumerate(packs):
        l0 = kt.pack_bits((height >= 500) & (height <= 60_000))
        want = ko.tree_eval([0, 1, 0xFE], [l0, ko.StrContainer(blob).match(ko.EQ, x)], nrows[p])
        assert (res["bitsets"][p] == want).all()
        st = ko.reduce(ko.I64, amount, want, st)
    g = res["aggs"][0]
    assert (g.count, g.sum_bits, g.min_bits, g.max_bits) == (st.count, st.sum_bits, st.min_bits, st.max_bits) and st.count >= 1
    prog.close()
    for p in range(3):
        for f in (1, 2, 3):
            ctx.block_drop(310 + p, 1, f)


@pytest.mark.parametrize("shape", ["fixed20", "dups", "ragged", "const"])
def test_gather_bytes_returns_the_selected_rows(ctx, shape):
    """StringContainer.AppendTo(dst, sel) over a batch of packs (kx_gather_bytes): a string predicate picks the rows
    (kx_scan_select), the result column is another string block of the same packs; the bytes that come back are the
    oracle's rows at those ids, for every string container, with empty selections, empty strings and the size query."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(17)
    nrows = [1, 33, 5000, 20_011]
    kinds_all = (ko.STR_CONST, ko.STR_FIXED, ko.STR_COMPACT, ko.STR_DICT)
    tables = []
    for p, n in enumerate(nrows):
        rows = _string_rows(rng, n, shape)
        kinds = [k for k in kinds_all if ko.store_str(k, rows) is not None and (k != ko.STR_DICT or n <= 5000 or shape in ("dups", "const"))]
        kind = kinds[p % len(kinds)]                       # every pack another container
        key = rng.integers(0, 50, n).astype(np.int64)
        assert ctx.block_put(970 + p, 1, 1, kb.INT64, ko.store("best", ko.I64, key)) == n
        assert ctx.block_put(970 + p, 1, 2, kb.BYTES, ko.store_str(kind, rows)) == n
        tables.append((rows, key))
    packs = [(970 + p, 1) for p in range(len(nrows))]
    for lo, hi in ((7, 7), (0, 49), (60, 70), (3, 20)):
        prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, lo, hi)])
        r = ctx.scan_select(prog, packs)
        sel, off = r["sel"], r["sel_off"]
        prog.close()
        want = [rows[i] for (rows, key) in tables for i in np.nonzero((key >= lo) & (key <= hi))[0]]
        assert int(off[-1]) == len(want)
        got = ctx.gather_bytes(packs, 2, sel, off)
        assert got == want, (shape, lo, hi)
        if want:
            exact = sum(len(w) for w in want)
            assert ctx.gather_bytes(packs, 2, sel, off, capacity=exact) == want          # exactly as large as needed
            if exact:
                with pytest.raises(kb.KnoxError):
                    ctx.gather_bytes(packs, 2, sel, off, capacity=exact - 1)             # one byte short: refused, nothing written
    with pytest.raises(kb.KnoxError):
        ctx.gather_bytes(packs, 1, np.zeros(1, np.uint32), np.array([0, 1, 1, 1, 1], np.uint64))   # not a string block
    with pytest.raises(kb.KnoxError):
        ctx.gather_bytes(packs, 2, np.array([5], np.uint32), np.array([0, 1, 1, 1, 1], np.uint64))  # row id outside pack 0 (1 row)
    for p in range(len(nrows)):
        ctx.block_drop(970 + p, 1, 1); ctx.block_drop(970 + p, 1, 2)


@pytest.mark.parametrize("shape", ["fixed20", "dups", "ragged", "const"])
def test_scan_host_takes_string_blocks(ctx, shape):
    """kx_scan_host (cold device cache: blocks still in host memory) with a byte-string column: `height BETWEEN … AND
    address <op> X` + sum(amount) over packs whose address blocks use every string container; the result equals the
    resident scan's and the oracle's row by row."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(19)
    nrows = [64, 5000, 20_011, 1]
    kinds_all = (ko.STR_CONST, ko.STR_FIXED, ko.STR_COMPACT, ko.STR_DICT)
    hb, packs = [], []
    for p, n in enumerate(nrows):
        rows = _string_rows(rng, n, shape)
        kinds = [k for k in kinds_all if ko.store_str(k, rows) is not None and (k != ko.STR_DICT or n <= 5000 or shape in ("dups", "const"))]
        blob = ko.store_str(kinds[p % len(kinds)], rows)
        height = (100 * p + np.arange(n)).astype(np.int64)
        amount = rng.integers(-10**6, 10**6, n).astype(np.int64)
        hb.append([np.frombuffer(ko.store("best", ko.I64, height), np.uint8), np.frombuffer(blob, np.uint8), np.frombuffer(ko.store("best", ko.I64, amount), np.uint8)])
        packs.append((rows, height, amount, blob))
    fields = [(1, kb.INT64), (2, kb.BYTES), (3, kb.INT64)]
    x = packs[1][0][777]
    for leaf, kom, args in ((kb.Leaf(2, kb.BYTES, kb.EQ, x), ko.EQ, (x,)), (kb.Leaf(2, kb.BYTES, kb.GE, x), ko.GE, (x,)),
                            (kb.Leaf(2, kb.BYTES, kb.RANGE, b"ab", x + b"z"), ko.RG, (b"ab", x + b"z"))):
        prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, 10, 15_000), leaf])
        res = ctx.scan_host(prog, fields, hb, nrows=nrows, want_bitsets=True, aggs=[(3, kb.INT64)])
        st = None
        for p, (rows, height, amount, blob) in enumerate(packs):
            want = ko.tree_eval([0, 1, 0xFE], [kt.pack_bits((height >= 10) & (height <= 15_000)), ko.StrContainer(blob).match(kom, *args)], nrows[p])
            assert (res["bitsets"][p] == want).all(), (shape, p, kom)
            assert int(res["counts"][p]) == int(np.unpackbits(want).sum())
            st = ko.reduce(ko.I64, amount, want, st)
        g = res["aggs"][0]
        assert (g.count, g.sum_bits) == (st.count, st.sum_bits)
        if st.count:
            assert (g.min_bits, g.max_bits) == (st.min_bits, st.max_bits)
        prog.close()
    # IN set over the host blocks
    members = [packs[2][0][5], packs[0][0][0], b"not there"]
    prog = kb.Program(ctx, [kb.Leaf(2, kb.BYTES, kb.IN, values=members)])
    res = ctx.scan_host(prog, fields, hb, nrows=nrows, want_bitsets=True)
    for p, (rows, *_rest) in enumerate(packs):
        ms = set(members)
        assert (res["bitsets"][p] == kt.pack_bits(np.fromiter((r in ms for r in rows), dtype=bool, count=nrows[p]))).all()
    prog.close()


def test_repeated_scans_are_bit_reproducible(ctx):
    """The general kernel shares match words between warps, recycles ring stages and lets the producer pick the way
    value columns arrive from timing-dependent feedback: a race or an order dependence would show as run-to-run
    differences.  40 repetitions of a three-leaf scan with three aggregates must agree bit for bit (float sum included)."""
    import knoxdb_b200 as kb
    n, npacks = 200_000, 64
    base, accts = make_table(RNG, 2, [n, n])
    encs = []
    for cols in base:
        encs.append({F_TS: ko.store("best", ko.I64, cols["ts"]), F_AMT: ko.store("best", ko.I64, cols["amount"]),
                     F_FAMT: ko.store("raw", ko.F64, cols["famount"]), F_ACCT: ko.store("dict", ko.U64, cols["acct"])})
    for p in range(npacks):
        for f, t in ((F_TS, ko.I64), (F_AMT, ko.I64), (F_FAMT, ko.F64), (F_ACCT, ko.U64)):
            ctx.block_put(p, 1, f, t, encs[p % 2][f])
    ts = base[0]["ts"]
    prog = kb.Program(ctx, [kb.Leaf(F_TS, kb.INT64, kb.RANGE, int(ts[n // 10]), int(ts[n * 6 // 10])),
                            kb.Leaf(F_ACCT, kb.UINT64, kb.IN, values=RNG.choice(accts, 250, replace=False)),
                            kb.Leaf(F_AMT, kb.INT64, kb.GT, -5 * 10**8)], [0, 1, kb.OP_AND, 2, kb.OP_OR])
    refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
    first = None
    for rep in range(40):
        r = ctx.scan(prog, refs, nrows=[n] * npacks, want_bitsets=True, aggs=[(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64), (F_ACCT, kb.UINT64)])
        got = (r["counts"].tobytes(), b"".join(b.tobytes() for b in r["bitsets"]),
               [(g.count, g.sum_bits, g.sum_err, g.min_bits, g.max_bits) for g in r["aggs"]])
        if first is None:
            first = got
        assert got == first, rep
    prog.close()
    for p in range(npacks):
        for f in (F_TS, F_AMT, F_FAMT, F_ACCT):
            ctx.block_drop(p, 1, f)


def test_full_size_pack_properties(ctx):
    """BASELINE config 2 at full size: 4M-row bit-packed packs; size-independent checks
    (count == popcount(bitset) == numpy truth; NE is the complement of EQ; LT ∪ GE covers all rows)"""
    import knoxdb_b200 as kb
    n = 4 * 1024 * 1024
    for w in (8, 20, 33):
        vals = kt.rnd_bits(RNG, n, w) + np.uint64(1000)
        blob = ko.store("bitpack", ko.U64, vals)
        assert ctx.block_put(0, 1, 9, kb.UINT64, blob) == n
        med = int(np.median(vals))
        res = {}
        for mode in (kb.EQ, kb.NE, kb.LT, kb.GE):
            prog = kb.Program(ctx, [kb.Leaf(9, kb.UINT64, mode, int(vals[12345]) if mode in (kb.EQ, kb.NE) else med)])
            r = ctx.scan(prog, [(0, 1)], nrows=[n], want_bitsets=True)
            res[mode] = (int(r["counts"][0]), r["bitsets"][0].copy())
            prog.close()
        assert (res[kb.EQ][1] == kt.pack_bits(vals == vals[12345])).all()
        assert (res[kb.LT][1] == kt.pack_bits(vals < np.uint64(med))).all()
        for mode in res:
            assert res[mode][0] == int(np.unpackbits(res[mode][1]).sum())
        assert ((res[kb.EQ][1] ^ res[kb.NE][1]) == 0xFF).all() and res[kb.EQ][0] + res[kb.NE][0] == n
        assert ((res[kb.LT][1] | res[kb.GE][1]) == 0xFF).all() and res[kb.LT][0] + res[kb.GE][0] == n
        ctx.block_drop(0, 1, 9)


def test_prune_zone_maps_and_bloom(ctx):
    """stats.matchVector: zone maps via MatchRangeVectors semantics + bloom probes (config 4 shape)"""
    import knoxdb_b200 as kb
    L = ko.lib()
    npacks = 1000
    heights = np.arange(npacks, dtype=np.int64) * 100
    mins = np.stack([heights, RNG.integers(0, 2**40, npacks).astype(np.int64)], axis=1)
    maxs = np.stack([heights + 99, mins[:, 1] + RNG.integers(0, 2**40, npacks)], axis=1)
    # bloom per pack over 64 random "addresses" hashed with XXH3 of 20-byte strings
    m_bits = 64 * 16
    blooms, members = [], []
    for p in range(npacks):
        buf = np.zeros(L.ko_bloom_bytes(m_bits), dtype=np.uint8)
        L.ko_bloom_init(ko._p(buf), m_bits)
        addrs = RNG.integers(0, 256, (64, 20), dtype=np.uint8)
        hs = [L.ko_xxh3_bytes(ko._p(a), 20) for a in addrs]
        for h in hs:
            L.ko_bloom_add(ko._p(buf), buf.size, h)
        blooms.append([None, buf])
        members.append((addrs, hs))
    probe_addr = members[417][0][3]
    probe_hash = kb.lib().kx_hash_bytes(probe_addr.ctypes.data, 20)
    assert probe_hash == members[417][1][3]            # product hash == oracle hash
    lo, hi = 30000, 60000
    probe_val = int(mins[417, 1] + 5)
    prog = kb.Program(ctx, [kb.Leaf(0, kb.INT64, kb.RANGE, lo, hi), kb.Leaf(1, kb.INT64, kb.EQ, probe_val)])
    bits, nsurv = ctx.prune(prog, mins.view(np.uint64), maxs.view(np.uint64), blooms, [[], [probe_hash]])
    want = np.zeros(npacks, dtype=bool)
    for p in range(npacks):
        z = L.ko_match_range(ko.I64, ko.RG, ko.scalar_u64(ko.I64, lo), ko.scalar_u64(ko.I64, hi), int(mins[p, 0]) & (2**64 - 1), int(maxs[p, 0]) & (2**64 - 1))
        z2 = L.ko_match_range(ko.I64, ko.EQ, ko.scalar_u64(ko.I64, probe_val), 0, int(mins[p, 1]) & (2**64 - 1), int(maxs[p, 1]) & (2**64 - 1))
        bl = L.ko_bloom_contains(ko._p(blooms[p][1]), blooms[p][1].size, probe_hash)
        want[p] = bool(z and z2 and bl)
    assert (kt.unpack_bits(bits, npacks) == want).all()
    assert nsurv == int(want.sum()) and want[417]
    # without blooms: zone maps only
    bits2, n2 = ctx.prune(prog, mins.view(np.uint64), maxs.view(np.uint64))
    assert n2 >= nsurv and kt.unpack_bits(bits2, npacks)[417]
    prog.close()


def test_string_zone_maps_are_scans_over_the_min_max_blocks(ctx):
    """Zone maps of byte-string columns.  The reference keeps a stats pack's per-pack minima and maxima as two string BLOCKS
    and prunes by running ordinary string matchers over them (`bytes*Matcher.MatchRangeVectors`,
    internal/operator/filter/match_bytes.go:97-110 EQ, :165-175 GT / GE on maxs, :215-225 LT / LE on mins, :282-296 RANGE,
    :424-429 IN via the set's min / max): candidate packs = `mins LE hi AND maxs GE lo`.  The same call sequence through the
    C ABI: the two blocks are registered like any string block (row = data pack) and ONE two-leaf scan returns the
    candidate bitset; checked against Python's bytes ordering (= bytes.Compare) and against the packs' real rows (a zone
    map may keep too many packs, never too few)."""
    import knoxdb_b200 as kb
    rng = np.random.default_rng(21)
    npacks = 3000
    # sorted "address" table cut into packs of 1..40 rows: adjacent, occasionally overlapping [min, max] ranges
    rows = sorted(bytes(rng.integers(97, 123, int(k), dtype=np.uint8)) for k in rng.integers(1, 24, 40 * npacks))
    cuts = np.sort(rng.choice(np.arange(1, len(rows)), npacks - 1, replace=False))
    packs = [rows[a:b] for a, b in zip(np.r_[0, cuts], np.r_[cuts, len(rows)])]
    mins, maxs = [p[0] for p in packs], [p[-1] for p in packs]
    for kind in (ko.STR_COMPACT, ko.STR_DICT):
        bmin, bmax = ko.store_str(kind, mins), ko.store_str(kind, maxs)
        if bmin is None or bmax is None:
            continue
        assert ctx.block_put(900, 1, 1, kb.BYTES, bmin) == npacks and ctx.block_put(900, 1, 2, kb.BYTES, bmax) == npacks

        def candidates(lo, hi):
            prog = kb.Program(ctx, [kb.Leaf(1, kb.BYTES, kb.LE, hi), kb.Leaf(2, kb.BYTES, kb.GE, lo)])
            res = ctx.scan(prog, [(900, 1)], nrows=[npacks], want_bitsets=True)
            prog.close()
            return kt.unpack_bits(res["bitsets"][0], npacks), int(res["counts"][0])

        probes = [packs[1234][len(packs[1234]) // 2], packs[0][0], packs[-1][-1], b"", b"zzzzzzzzzzzzzzzzzzzzzzzzzzzz", packs[77][0] + b"\x00"]
        for v in probes:                                    # bytesEqualMatcher.MatchRangeVectors
            got, cnt = candidates(v, v)
            want = np.fromiter((mn <= v <= mx for mn, mx in zip(mins, maxs)), dtype=bool, count=npacks)
            assert (got == want).all() and cnt == int(want.sum())
            holds = np.fromiter((v in p for p in packs), dtype=bool, count=npacks)
            assert not (holds & ~got).any()                 # no pack that holds the value is pruned
        for lo, hi in ((b"d", b"f"), (b"kx", b"kxzz"), (b"", b"a"), (b"q", b"b")):   # bytesRangeMatcher.MatchRangeVectors
            got, cnt = candidates(lo, hi)
            want = np.fromiter((mn <= hi and mx >= lo for mn, mx in zip(mins, maxs)), dtype=bool, count=npacks)
            assert (got == want).all() and cnt == int(want.sum())
        members = [packs[5][0], packs[2500][-1], b"mmm"]    # bytesInSetMatcher.MatchRangeVectors: the set's min / max as a range
        got, _ = candidates(min(members), max(members))
        for m in members:
            assert not (np.fromiter((m in p for p in packs), dtype=bool, count=npacks) & ~got).any()
        ctx.block_drop(900, 1, 1); ctx.block_drop(900, 1, 2)


def test_agg_combine_matches_single_scan(ctx):
    """fixed-order combine of per-shard partials (multi-GPU epilogue) == one scan over all packs"""
    import ctypes as C
    import knoxdb_b200 as kb
    from knoxdb_b200.lib import AggOut
    nrows = [50000, 50000, 50000, 50000]
    packs, accts = make_table(RNG, 4, nrows)
    blobs = put_table(ctx, packs, "bitpack")
    prog = kb.Program(ctx, [kb.Leaf(F_AMT, kb.INT64, kb.GT, 0)])
    aggs = [(F_AMT, kb.INT64), (F_FAMT, kb.FLOAT64)]
    whole = ctx.scan(prog, [(p, 1) for p in range(4)], aggs=aggs)["aggs"]
    halves = [ctx.scan(prog, [(p, 1) for p in r], aggs=aggs)["aggs"] for r in ((0, 1), (2, 3))]
    for j, (f, t) in enumerate(aggs):
        parts = (AggOut * 2)(halves[0][j], halves[1][j])
        out = AggOut()
        assert kb.lib().kx_agg_combine(t, parts, 2, C.byref(out)) == 0
        assert out.count == whole[j].count and out.min_bits == whole[j].min_bits and out.max_bits == whole[j].max_bits
        if t == kb.INT64:
            assert out.sum_bits == whole[j].sum_bits
        else:
            a, b = np.uint64(out.sum_bits).view(np.float64), np.uint64(whole[j].sum_bits).view(np.float64)
            assert abs(a - b) <= 1e-14 * abs(b)
    for p in range(4):
        for f in (F_TS, F_ACCT, F_AMT, F_FAMT):
            ctx.block_drop(p, 1, f)
    prog.close()


def test_errors_are_loud(ctx):
    import knoxdb_b200 as kb
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.EQ, 5)])
    with pytest.raises(kb.KnoxError):
        ctx.scan(prog, [(999, 1)])                       # block not resident
    with pytest.raises(kb.KnoxError):
        ctx.block_put(0, 0, 0, kb.INT64, b"\x63\x00")    # unknown container id
    with pytest.raises(kb.KnoxError):
        kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.EQ, 5)], postfix=[0, kb.OP_AND])
    prog.close()
    # string blocks and leaves
    with pytest.raises(kb.KnoxError):
        ctx.block_put(0, 0, 0, kb.BYTES, b"\x63\x00\x00")    # unknown string container id
    with pytest.raises(kb.KnoxError):
        ctx.block_put(0, 0, 0, kb.BYTES, ko.store_str(ko.STR_FIXED, [b"abc", b"abd"])[:-2])   # truncated buffer
    vals = np.arange(100, dtype=np.int64)
    ctx.block_put(900, 1, 1, kb.INT64, ko.store("raw", ko.I64, vals))
    ctx.block_put(900, 1, 2, kb.BYTES, ko.store_str(ko.STR_COMPACT, [b"x%d" % i for i in range(100)]))
    p_str_on_int = kb.Program(ctx, [kb.Leaf(1, kb.BYTES, kb.EQ, b"x")])
    p_int_on_str = kb.Program(ctx, [kb.Leaf(2, kb.INT64, kb.EQ, 5)])
    p_prune_only = kb.Program(ctx, [kb.Leaf(2, kb.BYTES, kb.EQ)])
    for bad in (p_str_on_int, p_int_on_str, p_prune_only):
        with pytest.raises(kb.KnoxError):
            ctx.scan(bad, [(900, 1)])
        bad.close()
    # time-bucketed scans: edges must ascend, the window column must be an integer block, blocks must be resident
    ok = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.GE, 0)])
    r = ctx.scan_buckets(ok, [(900, 1)], 1, kb.INT64, [0, 50, 100], aggs=[(1, kb.INT64)])
    assert r["bucket_counts"].tolist() == [50, 50] and [g.sum_bits for g in r["aggs"][0]] == [int(vals[:50].sum()), int(vals[50:].sum())]
    with pytest.raises(kb.KnoxError):
        ctx.scan_buckets(ok, [(900, 1)], 1, kb.INT64, [0, 100, 50])
    with pytest.raises(kb.KnoxError):
        ctx.scan_buckets(ok, [(900, 1)], 2, kb.INT64, [0, 50, 100])          # window column is a string block
    with pytest.raises(kb.KnoxError):
        ctx.scan_buckets(ok, [(901, 1)], 1, kb.INT64, [0, 50, 100])          # pack not resident
    ok.close()
    ctx.block_drop(900, 1, 1); ctx.block_drop(900, 1, 2)


def test_resident_stats_index_bloom_build_and_prune(ctx):
    """kx_stats: blooms BUILT on the device are bit-identical to stats.BuildBloomFilter's buffers (oracle), and
    pruning over the resident index equals zone-map + bloom evaluation pack by pack (stats/match.go:92-195)."""
    import knoxdb_b200 as kb
    L = ko.lib()
    npacks, per_pack = 300, 500
    heights = np.arange(npacks, dtype=np.int64) * per_pack
    vals_u64 = [RNG.integers(0, 2**50, per_pack, dtype=np.uint64) for _ in range(npacks)]
    addr = [RNG.integers(0, 256, (per_pack, 20), dtype=np.uint8) for _ in range(npacks)]
    offs = (np.arange(per_pack + 1, dtype=np.uint32) * 20)
    fields = [(1, kb.INT64), (2, kb.UINT64), (3, kb.BYTES)]
    mins = np.stack([heights.view(np.uint64), np.array([v.min() for v in vals_u64], dtype=np.uint64), np.zeros(npacks, dtype=np.uint64)])
    maxs = np.stack([(heights + per_pack - 1).view(np.uint64), np.array([v.max() for v in vals_u64], dtype=np.uint64), np.zeros(npacks, dtype=np.uint64)])
    st = kb.Stats(ctx, fields, mins, maxs)
    want_u64, want_addr = [], []
    for p in range(npacks):
        st.build_bloom(1, p, kb.UINT64, vals_u64[p], per_pack, 2)
        want_u64.append(ko.bloom_build(vals_u64[p], per_pack, 2))
        if p % 3:   # every third pack has no address filter: it must survive bloom pruning
            st.build_bloom(2, p, kb.BYTES, addr[p].reshape(-1), per_pack, 3, offsets=offs)
            want_addr.append(ko.bloom_build(addr[p].reshape(-1), per_pack, 3, offsets=offs))
        else:
            want_addr.append(None)
    for p in (0, 1, 2, 77, 299):
        assert (st.get_bloom(1, p) == want_u64[p]).all()
        got = st.get_bloom(2, p)
        assert (got is None) == (want_addr[p] is None) and (got is None or (got == want_addr[p]).all())
    # narrow types and a stored filter attached verbatim
    for t, ktp in ((np.int32, kb.INT32), (np.uint16, kb.UINT16), (np.uint8, kb.UINT8), (np.float64, kb.FLOAT64)):
        v = RNG.integers(0, 200, 333).astype(t)
        st2 = kb.Stats(ctx, [(9, ktp)], np.zeros((1, 1), np.uint64), np.zeros((1, 1), np.uint64))
        st2.build_bloom(0, 0, ktp, v, 200, 4)
        assert (st2.get_bloom(0, 0) == ko.bloom_build(v, 200, 4)).all(), t
        st2.put_bloom(0, 0, want_u64[5])
        assert (st2.get_bloom(0, 0) == want_u64[5]).all()
        st2.close()
    # query: height range AND (value = x OR address = y)
    x, y = int(vals_u64[123][7]), addr[200][11]
    hy = kb.lib().kx_hash_bytes(y.ctypes.data, 20)
    hx = kb.lib().kx_hash_value(kb.UINT64, x)
    assert hy == L.ko_xxh3_bytes(ko._p(y), 20) and hx == L.ko_xxh3_u64(x)
    lo, hi = 20 * per_pack + 3, 250 * per_pack
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, lo, hi), kb.Leaf(2, kb.UINT64, kb.EQ, x), kb.Leaf(3, kb.BYTES, kb.EQ)],
                      postfix=[0, 1, 2, kb.OP_OR, kb.OP_AND])
    bits, nsurv = st.prune(prog, [[], [hx], [hy]])
    want = np.zeros(npacks, dtype=bool)
    for p in range(npacks):
        z0 = L.ko_match_range(ko.I64, ko.RG, lo, hi, int(mins[0, p]), int(maxs[0, p]))
        z1 = L.ko_match_range(ko.U64, ko.EQ, x, 0, int(mins[1, p]), int(maxs[1, p])) and L.ko_bloom_contains(ko._p(want_u64[p]), want_u64[p].size, hx)
        z2 = True if want_addr[p] is None else bool(L.ko_bloom_contains(ko._p(want_addr[p]), want_addr[p].size, hy))
        want[p] = bool(z0 and (z1 or z2))
    assert (kt.unpack_bits(bits, npacks) == want).all()
    assert nsurv == int(want.sum()) and want[123] and want[200]
    # library-side hashing of numeric operands (hashes = NULL): the address leaf then probes nothing
    prog2 = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, lo, hi), kb.Leaf(2, kb.UINT64, kb.IN, values=np.array([x, 12345], dtype=np.uint64))])
    bits2, n2 = st.prune(prog2)
    want2 = np.zeros(npacks, dtype=bool)
    h12345 = L.ko_xxh3_u64(12345)
    for p in range(npacks):
        z0 = L.ko_match_range(ko.I64, ko.RG, lo, hi, int(mins[0, p]), int(maxs[0, p]))
        inr = any(int(mins[1, p]) <= v <= int(maxs[1, p]) for v in (x, 12345))
        bl = any(L.ko_bloom_contains(ko._p(want_u64[p]), want_u64[p].size, h) for h in (hx, h12345))
        want2[p] = bool(z0 and inr and bl)
    assert (kt.unpack_bits(bits2, npacks) == want2).all() and n2 == int(want2.sum())
    with pytest.raises(kb.KnoxError):
        ctx.scan(prog, [(0, 1)], nrows=[1])      # byte-string programs are prune-only
    prog.close(); prog2.close(); st.close()


def test_alp_float_blocks(ctx):
    """FloatAlpContainer on the device: Match* through the encoded integer domain + patch
    correction stream, AppendTo (decode) and the fused reduce over an ALP value column — all against the oracle's
    restatement of float_alp.go, bit for bit (bitsets, counts, min/max) / 1e-12 (float sums)."""
    import knoxdb_b200 as kb
    import test_host_translation as tht
    t = ko.F64
    for n in (1, 33, 1000, 8192 + 77, 70_000):
        for name, vals in tht.alp_columns(RNG, n).items():
            blob = ko.store("alp", t, vals)
            oc = ko.Container(t, blob)
            got = ctx.container_decode(kb.FLOAT64, blob, n)
            assert (got.view(np.uint64) == oc.decode()).all(), (n, name)
            for a in tht.alp_operands(vals)[:: 2 if n > 10_000 else 1]:
                for b in (a + 2.5, a - 1.0):
                    for op in kt.OPS:
                        ua, ub = ko.scalar_u64(t, a), ko.scalar_u64(t, b)
                        want = oc.match(op, ua, ub)
                        bits, cnt = ctx.container_match(kb.FLOAT64, blob, op, a, b, nrows=n)
                        assert (bits == want).all(), (n, name, op, a, b)
                        assert cnt == int(np.unpackbits(want).sum())
    # scan with two ALP columns: price < p AND qty >= q, sum/min/max(price) and an int column
    n = 200_003
    price = np.round(RNG.uniform(0, 500, n), 2); price[::53] = RNG.uniform(0, 1, price[::53].size)
    qty = np.round(RNG.uniform(0, 50, n), 1); qty[::101] = RNG.uniform(0, 1, qty[::101].size) * np.pi
    ident = RNG.integers(0, 10**9, n).astype(np.int64)
    blobs = {1: (kb.FLOAT64, ko.store("alp", t, price)), 2: (kb.FLOAT64, ko.store("alp", t, qty)), 3: (kb.INT64, ko.store("best", ko.I64, ident))}
    for f, (bt, b) in blobs.items():
        assert ctx.block_put(77, 1, f, bt, b) == n
    prog = kb.Program(ctx, [kb.Leaf(1, kb.FLOAT64, kb.LT, 250.0), kb.Leaf(2, kb.FLOAT64, kb.GE, 12.3)])
    res = ctx.scan(prog, [(77, 1)], nrows=[n], want_bitsets=True, aggs=[(1, kb.FLOAT64), (3, kb.INT64)])
    os.environ["KX_AGG_STAGE"] = "always"   # the ALP value column staged through the ring: same bits
    try:
        res2 = ctx.scan(prog, [(77, 1)], nrows=[n], want_bitsets=True, aggs=[(1, kb.FLOAT64), (3, kb.INT64)])
    finally:
        del os.environ["KX_AGG_STAGE"]
    assert [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in res2["aggs"]] == [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in res["aggs"]]
    l0 = ko.Container(t, blobs[1][1]).match(ko.LT, ko.scalar_u64(t, 250.0))
    l1 = ko.Container(t, blobs[2][1]).match(ko.GE, ko.scalar_u64(t, 12.3))
    want = ko.tree_eval([0, 1, 0xFE], [l0, l1], n)
    assert (res["bitsets"][0] == want).all() and int(res["counts"][0]) == int(np.unpackbits(want).sum())
    ap = ko.reduce(t, ko.Container(t, blobs[1][1]).decode().view(np.float64), want)
    ai = ko.reduce(ko.I64, ident, want)
    g0, g1 = res["aggs"]
    assert g0.count == ap.count and (g0.min_bits, g0.max_bits) == (ap.min_bits, ap.max_bits)
    sa, sb = g0.value("sum", kb.FLOAT64), float(np.uint64(ap.sum_bits).view(np.float64))
    assert abs(sa - sb) <= 1e-12 * abs(sb)
    assert (g1.count, g1.sum_bits, g1.min_bits, g1.max_bits) == (ai.count, ai.sum_bits, ai.min_bits, ai.max_bits)
    prog.close()
    for f in blobs:
        ctx.block_drop(77, 1, f)


def test_scan_select_and_gather(ctx):
    """kx_scan_select = filter.Match + bits.Indexes per pack (reader.go:432-436), kx_gather = AppendTo(dst, sel) on the
    result columns (query/result.go:196-264): ids and gathered values equal the oracle's bitset → indexes → take."""
    import knoxdb_b200 as kb
    L = ko.lib()
    sizes = [70_001, 1, 8192, 33_333, 64, 100_000]
    packs, cols = [], {}
    for p, n in enumerate(sizes):
        ts = (1_700_000_000 + np.cumsum(RNG.integers(0, 3, n))).astype(np.int64)
        acct = RNG.choice(RNG.integers(0, 2**40, 50), n).astype(np.uint64)
        price = np.round(RNG.uniform(0, 100, n), 2); price[::41] = RNG.uniform(0, 1, price[::41].size)
        small = RNG.integers(-100, 100, n).astype(np.int16)
        enc = {1: (kb.INT64, ko.I64, ko.store("best", ko.I64, ts), ts), 2: (kb.UINT64, ko.U64, ko.store("dict" if n > 1 else "raw", ko.U64, acct), acct),
               3: (kb.FLOAT64, ko.F64, ko.store("alp", ko.F64, price), price), 4: (kb.INT16, ko.I16, ko.store("bitpack", ko.I16, small), small)}
        for f, (kbt, _, blob, _) in enc.items():
            assert ctx.block_put(500 + p, 3, f, kbt, blob) == n
        packs.append((500 + p, 3)); cols[p] = enc
    setv = np.unique(np.concatenate([cols[p][2][3][:3] for p in cols]))
    t_lo, t_hi = 1_700_000_000 + 500, 1_700_000_000 + 60_000
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(2, kb.UINT64, kb.IN, values=setv)])
    r = ctx.scan_select(prog, packs, cap=16)        # too small on purpose: the wrapper retries with the reported size
    want_ids, want_off = [], [0]
    for p, n in enumerate(sizes):
        l0 = ko.Container(ko.I64, cols[p][1][2]).match(ko.RG, ko.scalar_u64(ko.I64, t_lo), ko.scalar_u64(ko.I64, t_hi))
        l1 = ko.Container(ko.U64, cols[p][2][2]).match_set(setv)
        bits = ko.tree_eval([0, 1, 0xFE], [l0, l1], n)
        ids = np.zeros(n + 8, dtype=np.uint32)
        k = L.ko_bitset_indexes(ko._p(bits), n, ko._p(ids))
        want_ids.append(ids[:k]); want_off.append(want_off[-1] + k)
    assert r["sel_off"].tolist() == want_off
    assert (r["sel"] == np.concatenate(want_ids)).all()
    assert r["counts"].tolist() == [len(x) for x in want_ids] and want_off[-1] > 100
    for f in (1, 2, 3, 4):
        kbt = cols[0][f][0]
        got = ctx.gather(packs, f, kbt, r["sel"], r["sel_off"])
        want = np.concatenate([cols[p][f][3][want_ids[p]] for p in range(len(sizes))])
        if f == 3:   # ALP decodes -0.0 as 0.0 like the reference; compare through the oracle's decode
            want = np.concatenate([ko.Container(ko.F64, cols[p][3][2]).decode().view(np.float64)[want_ids[p]] for p in range(len(sizes))])
            assert (got.view(np.uint64) == want.view(np.uint64)).all()
        else:
            assert (got == want).all(), f
    # empty selection
    prog2 = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.LT, 5)])
    r2 = ctx.scan_select(prog2, packs)
    assert r2["sel"].size == 0 and r2["sel_off"].tolist() == [0] * (len(sizes) + 1)
    prog.close(); prog2.close()
    for p in range(len(sizes)):
        for f in (1, 2, 3, 4):
            ctx.block_drop(500 + p, 3, f)
