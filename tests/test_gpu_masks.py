"""Row masks through kx_scan_ex / kx_scan_buckets: what the reference's reader does between filter.Match and
bits.Indexes / aggregation — tombstoned rids and rows the snapshot may not see are cleared from the match bitset
(internal/pack/table/reader.go:347-413, engine.TableReader.WithMask engine/interface.go:96-106).  Expected results are
the oracle's filter bitsets ANDed with the mask on the CPU, then Indexes / reducers of the oracle."""
import numpy as np
import pytest

import knoxdb_b200 as kb
import oracle as ko

pytestmark = pytest.mark.gpu
RNG = np.random.default_rng(4242)


@pytest.fixture(scope="module")
def ctx():
    c = kb.Context(0)
    yield c
    c.close()


def _table(ctx, sizes, base):
    packs = []
    for p, n in enumerate(sizes):
        rid = (np.arange(n, dtype=np.uint64) + np.uint64(1_000_000 * p + 1))
        ts = (1_700_000_000 + 10_000 * p + np.cumsum(RNG.integers(0, 3, n))).astype(np.int64)
        acct = RNG.choice(RNG.integers(0, 2**40, 37), n).astype(np.uint64)
        amt = RNG.integers(-10**9, 10**9, n).astype(np.int64)
        val = RNG.integers(0, 2**40, n).astype(np.float64) / 100.0
        blobs = {1: (kb.INT64, ko.I64, ko.store("best", ko.I64, ts)), 2: (kb.UINT64, ko.U64, ko.store("dict" if n > 1 else "raw", ko.U64, acct)),
                 3: (kb.INT64, ko.I64, ko.store("raw", ko.I64, amt)), 4: (kb.FLOAT64, ko.F64, ko.store("raw", ko.F64, val)),
                 5: (kb.UINT64, ko.U64, ko.store("best", ko.U64, rid))}
        for f, (kbt, _, blob) in blobs.items():
            assert ctx.block_put(base + p, 1, f, kbt, blob) == n
        packs.append(dict(n=n, rid=rid, ts=ts, acct=acct, amt=amt, val=val, blobs=blobs))
    return packs


def _mask_from_tombstones(pk, tomb, xmax_dead):
    """eligible rows = not tombstoned and not deleted before the snapshot (a stand-in for the xmin / xmax test)"""
    alive = ~np.isin(pk["rid"], tomb) & ~xmax_dead
    return np.packbits(alive.astype(np.uint8), bitorder="little"), alive


@pytest.mark.parametrize("postfix", [[0, 1, kb.OP_AND], [0, 1, kb.OP_OR], [0]])
def test_row_masks_clear_tombstoned_and_invisible_rows(ctx, postfix):
    sizes = [70_001, 1, 8192, 33_333, 64, 16_384 + 5, 100_000]
    base = {kb.OP_AND: 7000, kb.OP_OR: 7100, 0: 7200}[postfix[-1]]
    packs = _table(ctx, sizes, base)
    refs = [(base + p, 1) for p in range(len(sizes))]
    all_ts = np.concatenate([pk["ts"] for pk in packs])
    t_lo, t_hi = int(np.quantile(all_ts, 0.2)), int(np.quantile(all_ts, 0.7))
    setv = np.unique(np.concatenate([pk["acct"][:4] for pk in packs]))
    leaves = [kb.Leaf(1, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(2, kb.UINT64, kb.IN, values=setv)][: (2 if len(postfix) > 1 else 1)]
    prog = kb.Program(ctx, leaves, postfix)
    masks, alive = [], []
    for p, pk in enumerate(packs):
        if p == 2:                      # a pack without a mask: every row stays eligible
            masks.append(None); alive.append(np.ones(pk["n"], dtype=bool)); continue
        tomb = RNG.choice(pk["rid"], max(1, pk["n"] // 9), replace=False)
        dead = RNG.random(pk["n"]) < (0.02 if p != 4 else 1.0)      # pack 4: nothing visible at all
        m, a = _mask_from_tombstones(pk, tomb, dead)
        masks.append(m); alive.append(a)
    aggs = [(3, kb.INT64), (4, kb.FLOAT64)]
    res = ctx.scan_ex(prog, refs, nrows=sizes, masks=masks, want_bitsets=True, aggs=aggs)
    sel = ctx.scan_ex(prog, refs, nrows=sizes, masks=masks, want_sel=True, sel_cap=8)
    st_i = st_f = None
    want_ids, want_off = [], [0]
    for p, pk in enumerate(packs):
        n = pk["n"]
        l0 = ko.Container(ko.I64, pk["blobs"][1][2]).match(ko.RG, ko.scalar_u64(ko.I64, t_lo), ko.scalar_u64(ko.I64, t_hi))
        lb = [l0]
        if len(leaves) > 1:
            lb.append(ko.Container(ko.U64, pk["blobs"][2][2]).match_set(setv))
        bits = ko.tree_eval(postfix, lb, n)
        want = bits & np.packbits(alive[p].astype(np.uint8), bitorder="little")
        assert (res["bitsets"][p] == want).all(), (p, postfix)
        cnt = int(np.unpackbits(want, bitorder="little")[:n].sum())
        assert int(res["counts"][p]) == cnt == int(sel["counts"][p])
        ids = np.flatnonzero(np.unpackbits(want, bitorder="little")[:n]).astype(np.uint32)
        want_ids.append(ids); want_off.append(want_off[-1] + ids.size)
        st_i = ko.reduce(ko.I64, pk["amt"], want, st_i)
        st_f = ko.reduce(ko.F64, pk["val"], want, st_f)
    assert sel["sel_off"].tolist() == want_off and (sel["sel"] == np.concatenate(want_ids)).all()
    gi, gf = res["aggs"]
    assert (gi.count, gi.sum_bits, gi.min_bits, gi.max_bits) == (st_i.count, st_i.sum_bits, st_i.min_bits, st_i.max_bits)
    assert (gf.count, gf.min_bits, gf.max_bits) == (st_f.count, st_f.min_bits, st_f.max_bits)
    a, b = float(np.uint64(gf.sum_bits).view(np.float64)), float(np.uint64(st_f.sum_bits).view(np.float64))
    assert abs(a - b) <= 1e-12 * abs(b)
    # the same through the $rid NIN {tombstones} recipe (no mask): equal to a mask that only carries the tombstones
    pk = packs[0]
    tomb = RNG.choice(pk["rid"], 500, replace=False)
    prog2 = kb.Program(ctx, leaves + [kb.Leaf(5, kb.UINT64, kb.NIN, values=tomb)], postfix + [len(leaves), kb.OP_AND])
    r_leaf = ctx.scan(prog2, refs[:1], nrows=sizes[:1], want_bitsets=True, aggs=aggs)
    m, _ = _mask_from_tombstones(pk, tomb, np.zeros(pk["n"], dtype=bool))
    r_mask = ctx.scan_ex(prog, refs[:1], nrows=sizes[:1], masks=[m], want_bitsets=True, aggs=aggs)
    assert (r_leaf["bitsets"][0] == r_mask["bitsets"][0]).all() and r_leaf["counts"].tolist() == r_mask["counts"].tolist()
    assert [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in r_leaf["aggs"]] == [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in r_mask["aggs"]]
    prog.close(); prog2.close()
    for r in refs:
        for f in range(1, 6):
            ctx.block_drop(r[0], 1, f)


def test_row_masks_reach_the_time_bucketed_reduce(ctx):
    sizes = [50_000, 8192, 20_001]
    base = 7900
    packs = _table(ctx, sizes, base)
    refs = [(base + p, 1) for p in range(len(sizes))]
    t0 = int(min(pk["ts"][0] for pk in packs)); t1 = int(max(pk["ts"][-1] for pk in packs)) + 1
    edges = ko.window_edges(t0, t1, 3600)
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t0, t1 - 1)])
    masks, alive = [], []
    for pk in packs:
        a = RNG.random(pk["n"]) < 0.6
        alive.append(a); masks.append(np.packbits(a.astype(np.uint8), bitorder="little"))
    res = ctx.scan_buckets(prog, refs, 1, kb.INT64, edges, aggs=[(3, kb.INT64)], masks=masks)
    nb = edges.size - 1
    want_cnt = np.zeros(nb, dtype=np.int64); want_sum = np.zeros(nb, dtype=np.uint64)
    for p, pk in enumerate(packs):
        bits = ko.Container(ko.I64, pk["blobs"][1][2]).match(ko.RG, ko.scalar_u64(ko.I64, t0), ko.scalar_u64(ko.I64, t1 - 1)) & masks[p]
        st = ko.bucket_reduce(ko.I64, pk["amt"], ko.I64, pk["ts"], bits, edges)
        for k in range(nb):
            want_cnt[k] += st[k].count
            want_sum[k] = np.uint64((int(want_sum[k]) + st[k].sum_bits) & 0xFFFFFFFFFFFFFFFF)
    assert res["bucket_counts"].tolist() == want_cnt.tolist()
    assert [g.sum_bits for g in res["aggs"][0]] == want_sum.tolist()
    assert int(res["counts"].sum()) == int(sum(a.sum() for a in alive))
    prog.close()
    for r in refs:
        for f in range(1, 6):
            ctx.block_drop(r[0], 1, f)
