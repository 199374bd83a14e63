#!/usr/bin/env python3
"""sweep_configs.py — device-time sweep over BASELINE.json's configs 1-3 (+ container variants).

Every case registers a few distinct encoded blocks under many pack ids (so one launch reads far more
than the 126 MB L2), checks ONE pack bit for bit against the oracle, then reports the CUDA-event time
of the scan kernels (kx_last_scan_stats) as rows/s and algorithmic GB/s against MEASURED_PEAKS.json.
Algorithmic bytes per row follow SURVEY.md §8(d): encoded bytes of every filter column (+ 1/8 B when
bitsets are materialised, + value-column bytes for aggregates).

usage (GPU box):  python profiles/sweep_configs.py --out gpurun_out/sweep.json [--only c1,c2,...]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import knoxdb_b200 as kb   # noqa: E402
import oracle as ko        # noqa: E402  (checker only)

PEAK = 6450.0
M1, M4 = 1 << 20, 1 << 22
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def uv(x):
    if x <= 240:
        return bytes([x])
    if x <= 2287:
        y = x - 240
        return bytes([241 + (y >> 8), y & 0xFF])
    if x <= 67823:
        y = x - 2288
        return bytes([249, y >> 8, y & 0xFF])
    nb = max(3, (x.bit_length() + 7) // 8)
    return bytes([247 + nb]) + x.to_bytes(nb, "big")


def bitpack_block(rng, n, w, base=1000):
    """random payload = uniform w-bit fields (any bit string is a valid stream)"""
    nbytes = (n * w + 63) // 64 * 8
    payload = rng.integers(0, 256, nbytes, dtype=np.uint8)
    if (n * w) % 64:   # zero the padding bits of the last word like bitpack.Encode
        words = payload.view(np.uint64)
        words[-1] &= np.uint64((1 << ((n * w) % 64)) - 1)
    hdr = bytes([4]) + uv(base) + uv(w) + uv(n)
    return np.concatenate([np.frombuffer(hdr, dtype=np.uint8), payload])


def raw_block(vals, is_float=False):
    hdr = bytes([15 if is_float else 7]) + uv(vals.size)
    return np.concatenate([np.frombuffer(hdr, dtype=np.uint8), vals.view(np.uint8)])


class Case:
    def __init__(self, name, nrows, npacks, fields, leaves, postfix=None, aggs=(), bitsets=False, bytes_per_row=0.0, note=""):
        self.name, self.nrows, self.npacks = name, nrows, npacks
        self.fields = fields          # {field: (kb type, [distinct encoded blocks], ko type)}
        self.leaves, self.postfix, self.aggs, self.bitsets = leaves, postfix, list(aggs), bitsets
        self.bytes_per_row, self.note = bytes_per_row, note


def oracle_check(case, res, pack_idx):
    """bit-exact check of pack `pack_idx` of the result against the oracle"""
    n = case.nrows
    leaf_bits = []
    for lf in case.leaves:
        kbt, blocks, kot = case.fields[lf.field]
        c = ko.Container(kot, blocks[pack_idx % len(blocks)].tobytes())
        if lf.mode in (kb.IN, kb.NIN):
            leaf_bits.append(c.match_set(lf.set, negate=(lf.mode == kb.NIN)))
        else:
            leaf_bits.append(c.match(lf.mode, lf.a, lf.b))
    pf = case.postfix if case.postfix is not None else [0] + [x for i in range(1, len(case.leaves)) for x in (i, 0xFE)]
    want = ko.tree_eval(pf, leaf_bits, n)
    cnt = int(np.unpackbits(want).sum())
    assert int(res["counts"][pack_idx]) == cnt, (case.name, int(res["counts"][pack_idx]), cnt)
    if case.bitsets:
        assert (res["bitsets"][pack_idx] == want).all(), case.name + ": bitset mismatch"
    return want, cnt


def run_case(ctx, case, reps=5):
    t_setup = time.time()
    for f, (kbt, blocks, _) in case.fields.items():
        pinned = []
        for b in blocks:
            h = ctx.host_array(b.size); h[:] = b; pinned.append(h)
        for p in range(case.npacks):
            assert ctx.block_put(p, 1, f, kbt, pinned[p % len(pinned)]) == case.nrows
    prog = kb.Program(ctx, case.leaves, case.postfix)
    packs = ctx.pack_refs([(p, 1) for p in range(case.npacks)])
    nrows = [case.nrows] * case.npacks
    kw = dict(nrows=nrows, want_bitsets=case.bitsets, aggs=case.aggs)
    if case.bitsets:
        _, total = ctx.bitset_layout(nrows)
        kw["bitset_buf"] = ctx.host_array(total)
    res = ctx.scan(prog, packs, **kw)
    ndistinct = max(len(v[1]) for v in case.fields.values())
    matches = 0
    for pi in range(min(ndistinct, case.npacks)):
        want, cnt = oracle_check(case, res, pi)
        matches += cnt
    # aggregates: oracle reduce over all packs (distinct packs repeat)
    agg_ok = None
    f64_dev = None
    if case.aggs:
        agg_ok = True
        for j, (f, kbt) in enumerate(case.aggs):
            _, blocks, kot = case.fields[f]
            st = None
            wants = {}
            for p in range(case.npacks):
                d = p % ndistinct
                if d not in wants:
                    leaf_bits = []
                    for lf in case.leaves:
                        _, lb, lkot = case.fields[lf.field]
                        c = ko.Container(lkot, lb[d % len(lb)].tobytes())
                        leaf_bits.append(c.match_set(lf.set, negate=(lf.mode == kb.NIN)) if lf.mode in (kb.IN, kb.NIN) else c.match(lf.mode, lf.a, lf.b))
                    pf = case.postfix if case.postfix is not None else [0] + [x for i in range(1, len(case.leaves)) for x in (i, 0xFE)]
                    vals = ko.Container(kot, blocks[d % len(blocks)].tobytes()).decode()
                    vals = vals.view(np.float64) if kot == ko.F64 else (vals.view(np.int64) if kot == ko.I64 else vals)
                    wants[d] = (ko.tree_eval(pf, leaf_bits, case.nrows), vals)
                st = ko.reduce(kot, wants[d][1], wants[d][0], st)
            g = res["aggs"][j]
            assert g.count == st.count, (case.name, g.count, st.count)
            if kot == ko.F64:
                # The reference sums sequentially in float64 (reducer.go:173-178); over 10^7..10^8 addends that
                # running sum itself drifts ~1e-12 from the exact sum, so the compensated device sum is checked
                # against the EXACT sum (math.fsum) and its distance to the sequential oracle is recorded.
                import math
                per = {d: math.fsum(w[1][np.unpackbits(w[0], bitorder="little")[:case.nrows].astype(bool)].tolist()) for d, w in wants.items()}
                exact = math.fsum(per[p % ndistinct] for p in range(case.npacks))
                a = float(np.uint64(g.sum_bits).view(np.float64)); b = float(np.uint64(st.sum_bits).view(np.float64))
                assert abs(a - exact) <= 1e-14 * max(abs(exact), 1e-300), (case.name, a, exact)
                f64_dev = abs(a - b) / max(abs(b), 1e-300)
                assert f64_dev <= 1e-10, (case.name, a, b)
                assert (g.min_bits, g.max_bits) == (st.min_bits, st.max_bits)
            else:
                assert (g.sum_bits, g.min_bits, g.max_bits) == (st.sum_bits, st.min_bits, st.max_bits), case.name
    for _ in range(2):
        ctx.scan(prog, packs, **kw)
    k_ms, t_ms = [], []
    for _ in range(reps):
        ctx.scan(prog, packs, **kw)
        st = ctx.last_scan_stats()
        k_ms.append(st["kernel_ms"]); t_ms.append(st["total_ms"])
    prog.close()
    for f in case.fields:
        for p in range(case.npacks):
            ctx.block_drop(p, 1, f)
    ctx.free_host_arrays()
    rows = case.nrows * case.npacks
    km = float(np.median(k_ms))
    gbs = rows * case.bytes_per_row / (km * 1e-3) / 1e9
    out = {"case": case.name, "rows_per_launch": rows, "npacks": case.npacks, "pack_rows": case.nrows, "kernel_ms": km,
           "kernel_ms_min": float(min(k_ms)), "total_ms": float(np.median(t_ms)), "rows_per_s": rows / (km * 1e-3),
           "bytes_per_row": case.bytes_per_row, "algorithmic_GBps": gbs, "frac_of_measured_peak": gbs / PEAK,
           "selectivity": matches / (case.nrows * min(ndistinct, case.npacks)), "parity": "bit-exact vs oracle",
           "note": case.note, "setup_s": time.time() - t_setup}
    if agg_ok:
        out["aggregates"] = "ints bit-exact vs oracle; f64 sum within 1e-14 of the exact sum"
        if f64_dev is not None:
            out["f64_sum_rel_dev_vs_sequential_oracle"] = f64_dev
    return out


def build_cases(rng, only):
    cases = []

    def want(tag):
        return not only or tag in only

    # ---- config 1: raw uint64, 1 Mi rows, Between -> bitset + popcount
    if want("c1"):
        raws = [raw_block(rng.integers(0, 2**60 - 1, M1, dtype=np.uint64)) for _ in range(4)]
        f = {1: (kb.UINT64, raws, ko.U64)}
        for nm, a, b in (("bw(5,127)", 5, 127), ("bw(2^58,3*2^58)", 1 << 58, 3 << 58)):
            cases.append(Case(f"c1 raw u64 1Mi {nm} bitset+count", M1, 256, f, [kb.Leaf(1, kb.UINT64, kb.RANGE, a, b)], bitsets=True, bytes_per_row=8.125))
            cases.append(Case(f"c1 raw u64 1Mi {nm} count", M1, 256, f, [kb.Leaf(1, kb.UINT64, kb.RANGE, a, b)], bytes_per_row=8.0))
        cases.append(Case("c1 raw u64 1Mi eq count", M1, 256, f, [kb.Leaf(1, kb.UINT64, kb.EQ, int(raws[0].view(np.uint8)[16:24].view(np.uint64)[0]))], bytes_per_row=8.0))
        fl = [raw_block((rng.integers(0, 2**52, M1, dtype=np.uint64).astype(np.float64) / 100.0), True) for _ in range(4)]
        ff = {1: (kb.FLOAT64, fl, ko.F64)}
        cases.append(Case("c1 raw f64 1Mi lt(median) count", M1, 256, ff, [kb.Leaf(1, kb.FLOAT64, kb.LT, 2**51 / 100.0)], bytes_per_row=8.0))
        cases.append(Case("c1 raw f64 1Mi range bitset+count", M1, 256, ff, [kb.Leaf(1, kb.FLOAT64, kb.RANGE, 2**50 / 100.0, 2**51 / 100.0)], bitsets=True, bytes_per_row=8.125))

    # ---- config 2: bit-packed uint64 4 Mi rows, all interesting widths
    if want("c2"):
        for w in (1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 13, 15, 16, 17, 20, 24, 27, 31, 32, 33, 40, 48, 56, 63):
            blocks = [bitpack_block(rng, M4, w) for _ in range(2 if w > 32 else 4)]
            npacks = int(min(1024, max(32, 1.4e9 // (M4 * w / 8))))
            f = {1: (kb.UINT64, blocks, ko.U64)}
            thr = 1000 + (1 << (w - 1))
            cases.append(Case(f"c2 bitpack w={w} lt(median) count", M4, npacks, f, [kb.Leaf(1, kb.UINT64, kb.LT, thr)], bytes_per_row=w / 8))
            if w in (8, 20, 32, 48):
                eqv = 1000 + int(ko.Container(ko.U64, blocks[0].tobytes()).decode()[12345]) - 1000
                cases.append(Case(f"c2 bitpack w={w} eq count", M4, npacks, f, [kb.Leaf(1, kb.UINT64, kb.EQ, eqv)], bytes_per_row=w / 8))
                cases.append(Case(f"c2 bitpack w={w} lt(median) bitset+count", M4, npacks, f, [kb.Leaf(1, kb.UINT64, kb.LT, thr)], bitsets=True, bytes_per_row=w / 8 + 0.125))
                cases.append(Case(f"c2 bitpack w={w} range bitset+count", M4, npacks, f, [kb.Leaf(1, kb.UINT64, kb.RANGE, 1000 + (1 << (w - 2)), 1000 + (1 << (w - 1)))], bitsets=True, bytes_per_row=w / 8 + 0.125))

    # ---- config 2 shapes: dups -> dictionary (15-bit codes), runs -> run-end, seq -> affine delta
    if want("c2s"):
        uniq = np.unique(rng.integers(0, 2**40, 40000, dtype=np.uint64))[:32768]
        dups = [uniq[rng.integers(0, uniq.size, M4)] for _ in range(2)]
        dblocks = [np.frombuffer(ko.store("dict", ko.U64, d), dtype=np.uint8) for d in dups]
        f = {1: (kb.UINT64, dblocks, ko.U64)}
        dict_bpr = (dblocks[0].size) / M4
        cases.append(Case("c2s dict(32768 uniques, 15-bit codes) eq count", M4, 128, f, [kb.Leaf(1, kb.UINT64, kb.EQ, int(dups[0][777]))], bytes_per_row=dict_bpr))
        cases.append(Case("c2s dict lt(median) bitset+count", M4, 128, f, [kb.Leaf(1, kb.UINT64, kb.LT, int(uniq[uniq.size // 2]))], bitsets=True, bytes_per_row=dict_bpr + 0.125))
        cases.append(Case("c2s dict in{64} count", M4, 128, f, [kb.Leaf(1, kb.UINT64, kb.IN, values=uniq[::512])], bytes_per_row=dict_bpr))
        runs = [np.repeat(rng.integers(0, 2**24, M4 // 10 + 1, dtype=np.uint64), 10)[:M4] for _ in range(2)]
        rblocks = [np.frombuffer(ko.store("runend", ko.U64, r), dtype=np.uint8) for r in runs]
        f = {1: (kb.UINT64, rblocks, ko.U64)}
        cases.append(Case("c2s runend(run=10) lt(median) bitset+count", M4, 128, f, [kb.Leaf(1, kb.UINT64, kb.LT, 1 << 23)], bitsets=True, bytes_per_row=rblocks[0].size / M4 + 0.125,
                          note="bytes/row = encoded run values+ends; bitset write dominates"))
        dl = [np.frombuffer(ko.store("delta", ko.U64, base=5000, delta=1, n=M4), dtype=np.uint8)]
        f = {1: (kb.UINT64, dl, ko.U64)}
        cases.append(Case("c2s delta(seq) lt bitset+count", M4, 256, f, [kb.Leaf(1, kb.UINT64, kb.LT, 5000 + M4 // 2)], bitsets=True, bytes_per_row=0.125, note="no column bytes; bitset write only"))

    # ---- config 3: ts RANGE AND acct IN{..} with sum/min/max over amount
    if want("c3"):
        nd = 2
        ts = [(1_700_000_000 + np.cumsum(rng.integers(0, 3, M1))).astype(np.int64) for _ in range(nd)]
        uniq = np.unique(rng.integers(0, 2**40, 40000, dtype=np.uint64))[:32768]
        acct = [uniq[rng.integers(0, uniq.size, M1)] for _ in range(nd)]
        amt_i = [rng.integers(-10**9, 10**9, M1).astype(np.int64) for _ in range(nd)]
        amt_f = [(rng.integers(0, 2**40, M1).astype(np.float64) / 100.0) for _ in range(nd)]
        b_ts = [np.frombuffer(ko.store("bitpack", ko.I64, t), dtype=np.uint8) for t in ts]
        b_ad = [np.frombuffer(ko.store("dict", ko.U64, a), dtype=np.uint8) for a in acct]
        b_ab = [np.frombuffer(ko.store("bitpack", ko.U64, a), dtype=np.uint8) for a in acct]
        b_ai = [raw_block(a.view(np.uint64)) for a in amt_i]
        b_af = [raw_block(a.view(np.uint64), True) for a in amt_f]
        e_ts, e_ad, e_ab = b_ts[0].size / M1, b_ad[0].size / M1, b_ab[0].size / M1
        tmin, tmax = int(min(t[0] for t in ts)), int(max(t[-1] for t in ts))
        span = tmax - tmin

        def rng_leaf(frac):
            return kb.Leaf(1, kb.INT64, kb.RANGE, tmin, tmin + int(span * frac))
        for acct_kind, b_a, e_a in (("dict15", b_ad, e_ad), ("bitpack40", b_ab, e_ab)):
            for nset in (64, 4096):
                for amt_kind, b_m, kbt, kot in (("i64", b_ai, kb.INT64, ko.I64), ("f64", b_af, kb.FLOAT64, ko.F64)):
                    f = {1: (kb.INT64, b_ts, ko.I64), 2: (kb.UINT64, b_a, ko.U64), 3: (kbt, b_m, kot)}
                    cases.append(Case(f"c3 ts range(50%) AND acct({acct_kind}) in{{{nset}}} sum/min/max {amt_kind}", M1, 256, f,
                                      [rng_leaf(0.5), kb.Leaf(2, kb.UINT64, kb.IN, values=uniq[:: uniq.size // nset][:nset])], aggs=[(3, kbt)],
                                      bytes_per_row=e_ts + e_a + 8.0, note="bytes/row counts the full value column (8 B) as SURVEY 8(d) does"))
        for frac in (0.001, 0.1, 0.25, 0.5, 0.9):
            for amt_kind, b_m, kbt, kot in (("i64", b_ai, kb.INT64, ko.I64), ("f64", b_af, kb.FLOAT64, ko.F64)):
                f = {1: (kb.INT64, b_ts, ko.I64), 3: (kbt, b_m, kot)}
                cases.append(Case(f"c3 ts range({frac * 100:g}%) sum/min/max {amt_kind}", M1, 256, f, [rng_leaf(frac)], aggs=[(3, kbt)], bytes_per_row=e_ts + 8.0))
        # two-leaf AND without aggregates, bitsets out
        f = {1: (kb.INT64, b_ts, ko.I64), 2: (kb.UINT64, b_ab, ko.U64)}
        cases.append(Case("c3 ts range(50%) AND acct(bitpack40) lt(median) bitset+count", M1, 256, f,
                          [rng_leaf(0.5), kb.Leaf(2, kb.UINT64, kb.LT, int(uniq[uniq.size // 2]))], bitsets=True, bytes_per_row=e_ts + e_ab + 0.125))
    return cases


def run_buckets(ctx, rng, reps=5):
    """config 3 as a series query: ts RANGE filter, then count / sum / min / max of the amount columns per time window
    (kx_scan_buckets).  256 packs x 1 Mi rows; every distinct pack is checked window by window against the oracle."""
    nd, npacks = 2, 256
    ts = [(1_700_000_000 + np.cumsum(rng.integers(0, 3, M1))).astype(np.int64) for _ in range(nd)]
    amt_i = [rng.integers(-10**9, 10**9, M1).astype(np.int64) for _ in range(nd)]
    amt_f = [(rng.integers(0, 2**40, M1).astype(np.float64) / 100.0) for _ in range(nd)]
    b_ts = [np.frombuffer(ko.store("bitpack", ko.I64, t), dtype=np.uint8) for t in ts]
    b_ai = [raw_block(a.view(np.uint64)) for a in amt_i]
    b_af = [raw_block(a.view(np.uint64), True) for a in amt_f]
    for f, kbt, blocks in ((1, kb.INT64, b_ts), (3, kb.INT64, b_ai), (4, kb.FLOAT64, b_af)):
        pinned = []
        for b in blocks:
            h = ctx.host_array(b.size); h[:] = b; pinned.append(h)
        for p in range(npacks):
            assert ctx.block_put(p, 1, f, kbt, pinned[p % nd]) == M1
    tmin, tmax = int(min(t[0] for t in ts)), int(max(t[-1] for t in ts))
    refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
    out = []
    for frac, step, name in ((0.5, 3600, "hourly"), (0.9, 60, "per-minute")):
        t_from, t_to = tmin + 17, tmin + int((tmax - tmin) * frac)
        edges = ko.window_edges(t_from, t_to, step)
        nb = edges.size - 1
        prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t_from, t_to - 1)])
        res = ctx.scan_buckets(prog, refs, 1, kb.INT64, edges, aggs=[(3, kb.INT64), (4, kb.FLOAT64)])
        # oracle: per distinct pack, scaled by its multiplicity (integer sums wrap mod 2^64 either way)
        mult = [len(range(d, npacks, nd)) for d in range(nd)]
        want_cnt = np.zeros(nb, dtype=np.int64)
        want_sum = np.zeros(nb, dtype=np.uint64)
        want_f = np.zeros(nb)
        for d in range(nd):
            bits = ko.Container(ko.I64, b_ts[d].tobytes()).match(ko.RG, ko.scalar_u64(ko.I64, t_from), ko.scalar_u64(ko.I64, t_to - 1))
            si = ko.bucket_reduce(ko.I64, amt_i[d], ko.I64, ts[d], bits, edges)
            sf = ko.bucket_reduce(ko.F64, amt_f[d], ko.I64, ts[d], bits, edges)
            for k in range(nb):
                want_cnt[k] += si[k].count * mult[d]
                want_sum[k] += np.uint64((si[k].sum_bits * mult[d]) & 0xFFFFFFFFFFFFFFFF)
                want_f[k] += float(np.uint64(sf[k].sum_bits).view(np.float64)) * mult[d]
        assert res["bucket_counts"].tolist() == want_cnt.tolist(), name
        assert [g.sum_bits for g in res["aggs"][0]] == want_sum.tolist(), name
        for k in range(nb):
            if want_cnt[k]:
                assert abs(res["aggs"][1][k].value("sum", kb.FLOAT64) - want_f[k]) <= 1e-11 * abs(want_f[k]), (name, k)
        ks = []
        for _ in range(reps):
            ctx.scan_buckets(prog, refs, 1, kb.INT64, edges, aggs=[(3, kb.INT64), (4, kb.FLOAT64)])
            ks.append(ctx.last_scan_stats()["kernel_ms"])
        km = float(np.median(ks))
        rows = npacks * M1
        bpr = b_ts[0].size / M1 + 16.0
        r = {"case": f"c3 series: ts range({frac * 100:g}%) -> {nb} {name} windows, count/sum/min/max of i64 + f64", "rows_per_launch": rows, "npacks": npacks,
             "pack_rows": M1, "kernel_ms": km, "rows_per_s": rows / (km * 1e-3), "bytes_per_row": bpr, "algorithmic_GBps": rows * bpr / (km * 1e-3) / 1e9,
             "frac_of_measured_peak": rows * bpr / (km * 1e-3) / 1e9 / PEAK, "selectivity": float(want_cnt.sum()) / rows, "windows": nb,
             "parity": "window counts and integer sums bit-exact vs oracle; float64 window sums within 1e-11",
             "note": "scan kernel + bucket kernel; bytes/row counts ts + both full value columns as SURVEY 8(d) does"}
        out.append(r)
        print(f"{r['case']:<78s} {km:8.3f} ms {r['rows_per_s'] / 1e9:9.1f} Grows/s {r['algorithmic_GBps']:8.1f} GB/s {100 * r['frac_of_measured_peak']:5.1f}% sel={r['selectivity']:.4f}", flush=True)
        prog.close()
    for p in range(npacks):
        for f in (1, 3, 4):
            ctx.block_drop(p, 1, f)
    return out


def run_strings(ctx, rng, reps=5):
    """config 4 at row level: `address = X` on FieldTypeBytes columns (20-byte addresses) of packs that survived pruning.
    1024 packs x 65 536 rows; fixed-size container (what the encoder picks for equal-length strings) and dictionary
    container (256 distinct addresses per pack); every distinct block is checked against the oracle."""
    n, npacks, nd = 65536, 1024, 4
    out = []
    for name, kind, make in (("fixed", ko.STR_FIXED, lambda: [bytes(r) for r in rng.integers(0, 256, (n, 20), dtype=np.uint8)]),
                             ("dict(256 uniques)", ko.STR_DICT, None)):
        blocks, rows_d = [], []
        for d in range(nd):
            if make:
                rows = make()
            else:
                vocab = [bytes(r) for r in rng.integers(0, 256, (256, 20), dtype=np.uint8)]
                rows = [vocab[i] for i in rng.integers(0, 256, n)]
            rows_d.append(rows)
            blocks.append(np.frombuffer(ko.store_str(kind, rows), dtype=np.uint8))
        for p in range(npacks):
            assert ctx.block_put(p, 1, 2, kb.BYTES, blocks[p % nd]) == n
        x = rows_d[1][4242]
        prog = kb.Program(ctx, [kb.Leaf(2, kb.BYTES, kb.EQ, x)])
        refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
        res = ctx.scan(prog, refs, nrows=[n] * npacks, want_bitsets=True)
        for d in range(nd):
            want = ko.StrContainer(blocks[d].tobytes()).match(ko.EQ, x)
            assert (res["bitsets"][d] == want).all() and int(res["counts"][d]) == int(np.unpackbits(want).sum()), name
        ks = []
        for _ in range(reps):
            ctx.scan(prog, refs, nrows=[n] * npacks)
            ks.append(ctx.last_scan_stats()["kernel_ms"])
        km = float(np.median(ks))
        rows = npacks * n
        bpr = blocks[0].size / n
        r = {"case": f"c4 rows: address = X on 20-byte strings, {name} container, count", "rows_per_launch": rows, "npacks": npacks, "pack_rows": n,
             "kernel_ms": km, "rows_per_s": rows / (km * 1e-3), "bytes_per_row": bpr, "algorithmic_GBps": rows * bpr / (km * 1e-3) / 1e9,
             "frac_of_measured_peak": rows * bpr / (km * 1e-3) / 1e9 / PEAK, "selectivity": float(res["counts"].sum()) / rows,
             "parity": "bit-exact vs oracle", "note": "strmatch pre-pass + scan; bytes/row = encoded block bytes"}
        out.append(r)
        print(f"{r['case']:<78s} {km:8.3f} ms {r['rows_per_s'] / 1e9:9.1f} Grows/s {r['algorithmic_GBps']:8.1f} GB/s {100 * r['frac_of_measured_peak']:5.1f}% sel={r['selectivity']:.4f}", flush=True)
        prog.close()
        for p in range(npacks):
            ctx.block_drop(p, 1, 2)
    return out


def run_c4(ctx, rng, reps=20):
    """config 4: zone-map + bloom pruning over a 1 B-row block table (15 259 packs of 65 536 rows): per pack the
    zone map of `height` (int64) and a bloom filter over 65 536 20-byte `address` strings (FilterTypeBloom2b:
    m = pow2(65536*2*8) = 2^20 bits = 128 KiB, 1.9 GiB in total), resident on the device (kx_stats).  The filters are
    BUILT on the device from the strings (stats.BuildBloomFilter) and sampled ones are compared bit for bit with the
    oracle's; every query's surviving pack set is compared with the oracle's evaluation pack by pack."""
    L = ko.lib()
    npacks, per_pack, nd = 15259, 65536, 64
    sets = [rng.integers(0, 256, (per_pack, 20), dtype=np.uint8) for _ in range(nd)]   # pack p holds set p % nd
    offs = np.arange(per_pack + 1, dtype=np.uint32) * 20
    heights = np.arange(npacks, dtype=np.int64) * per_pack
    mins = np.stack([heights.view(np.uint64), np.zeros(npacks, dtype=np.uint64)])
    maxs = np.stack([(heights + per_pack - 1).view(np.uint64), np.zeros(npacks, dtype=np.uint64)])
    st = kb.Stats(ctx, [(1, kb.INT64), (2, kb.BYTES)], mins, maxs)
    flat = [x.reshape(-1) for x in sets]
    t0 = time.time()
    for p in range(npacks):
        st.build_bloom(1, p, kb.BYTES, flat[p % nd], per_pack, 2, offsets=offs)
    t_build = time.time() - t0
    oracle_blooms = [ko.bloom_build(flat[d], per_pack, 2, offsets=offs) for d in range(nd)]
    for p in (0, 1, 63, 64, 5000, npacks - 1):
        assert (st.get_bloom(1, p) == oracle_blooms[p % nd]).all(), f"device-built bloom of pack {p} differs from the oracle's"
    out = [{"case": "c4 bloom build on device (XXH3 of 20-byte strings + 4 bit sets per value)", "packs": npacks, "values": npacks * per_pack,
            "seconds_incl_h2d_of_values": t_build, "values_per_s": npacks * per_pack / t_build, "filter_bytes_total": npacks * (1 << 17),
            "parity": "6 sampled filters bit-identical to the oracle's BuildBloomFilter"}]
    x = sets[7][1234]
    hx = kb.lib().kx_hash_bytes(x.ctypes.data, 20)
    in16 = [sets[d][99] for d in range(16)]
    h16 = [kb.lib().kx_hash_bytes(v.ctypes.data, 20) for v in in16]
    lo, hi = 1000 * per_pack + 17, 2500 * per_pack
    queries = [
        ("c4 height between (10% of packs) AND address = X", [kb.Leaf(1, kb.INT64, kb.RANGE, lo, hi), kb.Leaf(2, kb.BYTES, kb.EQ)], [[], [hx]], (lo, hi)),
        ("c4 address = X (all packs probed)", [kb.Leaf(2, kb.BYTES, kb.EQ)], [[hx]], None),
        ("c4 address IN {16} (all packs probed)", [kb.Leaf(2, kb.BYTES, kb.IN)], [h16], None),
    ]
    for name, leaves, hashes, rg in queries:
        prog = kb.Program(ctx, leaves)
        bits, n = st.prune(prog, hashes)
        hl = hashes[-1]
        hit_d = [any(L.ko_bloom_contains(ko._p(oracle_blooms[d]), oracle_blooms[d].size, h) for h in hl) for d in range(nd)]
        want = np.array([hit_d[p % nd] for p in range(npacks)])
        if rg:
            zone = (heights <= rg[1]) & (heights + per_pack - 1 >= rg[0])
            want &= zone
            probed = int(zone.sum())
        else:
            probed = npacks
        got = np.unpackbits(bits, bitorder="little")[:npacks].astype(bool)
        assert (got == want).all() and n == int(want.sum()), name
        ks, ts = [], []
        for _ in range(reps):
            st.prune(prog, hashes)
            s_ = ctx.last_scan_stats()
            ks.append(s_["kernel_ms"]); ts.append(s_["total_ms"])
        km = float(np.median(ks))
        nprobe = probed * len(hl) * 4
        algb = npacks * 16 * (1 if rg else 0) + nprobe * 32
        out.append({"case": name, "packs": npacks, "kernel_ms": km, "total_ms": float(np.median(ts)), "packs_per_s": npacks / (km * 1e-3),
                    "bit_probes": nprobe, "bit_probes_per_s": nprobe / (km * 1e-3), "algorithmic_bytes": algb, "algorithmic_GBps": algb / (km * 1e-3) / 1e9,
                    "survivors": int(n), "rows_represented": npacks * per_pack, "parity": "surviving pack set identical to the oracle's",
                    "bound": "launch latency / random 32 B sectors (not HBM bandwidth)"})
        prog.close()
    st.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/sweep.json")
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    only = set(x for x in args.only.split(",") if x)
    rng = np.random.default_rng(1)
    ctx = kb.Context(0)
    results = []

    def canary(tag):
        """a fixed HBM-bound scan (raw uint64, 256 x 1 Mi rows, range -> count: 8 B/row) timed before and after the
        sweep: the box is healthy when it runs at >= ~100 % of the measured copy peak; a slow or throttled box shows here"""
        crng = np.random.default_rng(99)
        raws = [raw_block(crng.integers(0, 2**60 - 1, M1, dtype=np.uint64)) for _ in range(2)]
        c = Case(f"canary ({tag}) raw u64 1Mi range count", M1, 256, {1: (kb.UINT64, raws, ko.U64)}, [kb.Leaf(1, kb.UINT64, kb.RANGE, 1 << 58, 3 << 58)], bytes_per_row=8.0)
        r = run_case(ctx, c, args.reps)
        r["canary"] = True
        results.append(r)
        print(f"{r['case']:<78s} {r['kernel_ms']:8.3f} ms {r['rows_per_s'] / 1e9:9.1f} Grows/s {r['algorithmic_GBps']:8.1f} GB/s {100 * r['frac_of_measured_peak']:5.1f}%", flush=True)
        return r["frac_of_measured_peak"]

    health = [canary("before")]
    for case in build_cases(rng, only):
        try:
            r = run_case(ctx, case, args.reps)
        except Exception as e:   # keep sweeping; a failed case is reported, not hidden
            r = {"case": case.name, "error": repr(e)}
        results.append(r)
        if "error" in r:
            print(f"{r['case']:<78s} ERROR {r['error']}", flush=True)
        else:
            print(f"{r['case']:<78s} {r['kernel_ms']:8.3f} ms {r['rows_per_s'] / 1e9:9.1f} Grows/s {r['algorithmic_GBps']:8.1f} GB/s {100 * r['frac_of_measured_peak']:5.1f}% sel={r['selectivity']:.4f}", flush=True)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        json.dump({"peak_GBps": PEAK, "results": results}, open(args.out, "w"), indent=1)
    if not only or "c3b" in only or "c3" in only:
        try:
            rs = run_buckets(ctx, rng, args.reps)
        except Exception as e:
            rs = [{"case": "c3 series", "error": repr(e)}]
            print("c3 series ERROR", repr(e), flush=True)
        results.extend(rs)
    if not only or "c4s" in only or "c4" in only:
        try:
            rs = run_strings(ctx, rng, args.reps)
        except Exception as e:
            rs = [{"case": "c4 rows", "error": repr(e)}]
            print("c4 rows ERROR", repr(e), flush=True)
        results.extend(rs)
    health.append(canary("after"))
    json.dump({"peak_GBps": PEAK, "box_health": {"canary_frac_of_peak": health, "healthy": min(health) >= 0.95}, "results": results}, open(args.out, "w"), indent=1)
    if min(health) < 0.95:
        print(f"WARNING: the canary scan ran at {100 * min(health):.0f} % of the measured copy peak: this box is slow or throttled, treat the numbers with care", flush=True)
    if not only or "c4" in only:
        try:
            rs = run_c4(ctx, rng)
        except Exception as e:
            rs = [{"case": "c4", "error": repr(e)}]
        for r in rs:
            print(json.dumps(r), flush=True)
        results.extend(rs)
        json.dump({"peak_GBps": PEAK, "box_health": {"canary_frac_of_peak": health, "healthy": min(health) >= 0.95}, "results": results}, open(args.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
