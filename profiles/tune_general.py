#!/usr/bin/env python3
"""tune_general.py — sweep the warp-autonomous (multi-leaf filter + fused reduce) kernel's knobs over BASELINE config 3 shapes:
tile geometry (KX_WARP_GEOMETRY = wd,stages,warps), scheduling chunk (KX_SCHED_CHUNK) and the dense-tile threshold
(KX_AGG_STAGE).  Every combination must reproduce the default run's counts and integer
aggregates bit for bit (float sums: bit for bit as well — the lane/row assignment does not depend on the knobs'
staging decision).  Used to choose the defaults in kx_api.cu (run_scan); not part of the product path.

usage (GPU box): python profiles/tune_general.py --out gpurun_out/tune_general.json [--cases dict64,ts0.1,...]
"""
import argparse
import itertools
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "profiles"))

import knoxdb_b200 as kb            # noqa: E402
import oracle as ko                 # noqa: E402  (encoder of the synthetic blocks only)
from sweep_configs import raw_block   # noqa: E402

M1 = 1 << 20
KNOBS = ("KX_SCAN_GEOMETRY", "KX_SCHED_CHUNK", "KX_AGG_STAGE", "KX_WARP_GEOMETRY")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tune_general.json")
    ap.add_argument("--cases", default="dict64,hash64,ts0.1,ts10,ts50,ts90,ts90f")
    ap.add_argument("--npacks", type=int, default=256)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--minimal", action="store_true", help="default, dense-prefetch modes and the v2 kernel only")
    ap.add_argument("--chunks-only", action="store_true", help="default (chunk chosen by run_scan) against pinned KX_SCHED_CHUNK values")
    args = ap.parse_args()
    rng = np.random.default_rng(1)
    ctx = kb.Context(0)
    nd, npacks = 2, args.npacks
    ts = [(1_700_000_000 + np.cumsum(rng.integers(0, 3, M1))).astype(np.int64) for _ in range(nd)]
    uniq = np.unique(rng.integers(0, 2**40, 40000, dtype=np.uint64))[:32768]
    acct = [uniq[rng.integers(0, uniq.size, M1)] for _ in range(nd)]
    amt_i = [rng.integers(-10**9, 10**9, M1).astype(np.int64) for _ in range(nd)]
    amt_f = [(rng.integers(0, 2**40, M1).astype(np.float64) / 100.0) for _ in range(nd)]
    cols = {1: (kb.INT64, [np.frombuffer(ko.store("bitpack", ko.I64, t), dtype=np.uint8) for t in ts]),
            2: (kb.UINT64, [np.frombuffer(ko.store("dict", ko.U64, a), dtype=np.uint8) for a in acct]),
            4: (kb.UINT64, [np.frombuffer(ko.store("bitpack", ko.U64, a), dtype=np.uint8) for a in acct]),
            3: (kb.INT64, [raw_block(a.view(np.uint64)) for a in amt_i]),
            5: (kb.FLOAT64, [raw_block(a.view(np.uint64), True) for a in amt_f])}
    for f, (kbt, blocks) in cols.items():
        pinned = []
        for b in blocks:
            h = ctx.host_array(b.size); h[:] = b; pinned.append(h)
        for p in range(npacks):
            assert ctx.block_put(p, 1, f, kbt, pinned[p % nd]) == M1
    tmin, tmax = int(min(t[0] for t in ts)), int(max(t[-1] for t in ts))

    def rg(frac):
        return kb.Leaf(1, kb.INT64, kb.RANGE, tmin, tmin + int((tmax - tmin) * frac))
    in64 = uniq[:: uniq.size // 64][:64]
    cases = {
        "dict64": ([rg(0.5), kb.Leaf(2, kb.UINT64, kb.IN, values=in64)], [(3, kb.INT64)]),
        "hash64": ([rg(0.5), kb.Leaf(4, kb.UINT64, kb.IN, values=in64)], [(3, kb.INT64)]),
        "ts0.1": ([rg(0.001)], [(3, kb.INT64)]),
        "ts10": ([rg(0.1)], [(3, kb.INT64)]),
        "ts50": ([rg(0.5)], [(3, kb.INT64)]),
        "ts90": ([rg(0.9)], [(3, kb.INT64)]),
        "ts90f": ([rg(0.9)], [(5, kb.FLOAT64)]),
        "ts90if": ([rg(0.9)], [(3, kb.INT64), (5, kb.FLOAT64)]),
    }
    geos = [None, "2,4,32", "2,3,32", "2,2,64", "2,2,32", "1,4,64", "1,6,64", "1,3,64"]
    chunks = ["1", "2", "4", "8", "16"]
    if args.quick:
        geos, chunks = [None, "2,4,32", "2,2,64", "1,4,64"], ["1", "4"]
    refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
    nrows = [M1] * npacks
    results = []

    def run(prog, aggs):
        ks = []
        for i in range(5):
            r = ctx.scan(prog, refs, nrows=nrows, aggs=aggs)
            if i:
                ks.append(ctx.last_scan_stats()["kernel_ms"])
        sig = (int(r["counts"].sum()), [(g.count, g.sum_bits, g.min_bits, g.max_bits) for g in r["aggs"]])
        return float(np.median(ks)), sig

    for name in [c for c in args.cases.split(",") if c]:
        leaves, aggs = cases[name]
        prog = kb.Program(ctx, leaves)
        for k in KNOBS:
            os.environ.pop(k, None)
        base_ms, base_sig = run(prog, aggs)
        print(f"{name:8s} default                                  {base_ms:8.4f} ms  {npacks * M1 / base_ms / 1e6:8.1f} Grows/s", flush=True)
        results.append({"case": name, "knobs": {}, "kernel_ms": base_ms})
        if True:
            wgeos = [None, "1,2,16", "1,3,16", "1,4,16", "2,2,16", "2,3,16", "2,2,12", "2,3,12", "4,2,16", "4,2,8", "2,2,8"]
            wchunks = ["1", "2", "4", "8", "16"]
            if args.quick:
                wgeos, wchunks = [None, "1,2,16", "2,2,16", "2,3,12", "4,2,8"], ["4", "8", "16"]
            combos = [{"KX_WARP_GEOMETRY": g, "KX_SCHED_CHUNK": c} for g, c in itertools.product(wgeos, wchunks)]
            combos += [{"KX_AGG_STAGE": a} for a in ("never", "always", "2", "5", "8")]
            if args.minimal:
                combos = [{"KX_AGG_STAGE": "never"}, {"KX_AGG_STAGE": "5"}] + [{"KX_WARP_GEOMETRY": g} for g in ("2,2,11", "2,2,10", "2,2,8", "1,2,16", "1,3,16", "2,3,8")]
        if args.chunks_only:
            combos = [{"KX_SCHED_CHUNK": c} for c in ("1", "2", "4", "8", "16")]
        for combo in combos:
            for k in KNOBS:
                os.environ.pop(k, None)
            for k, v in combo.items():
                if v is not None:
                    os.environ[k] = v
            try:
                ms, sig = run(prog, aggs)
            except Exception as e:
                print(f"{name:8s} {combo} ERROR {e!r}", flush=True)
                continue
            ok = sig == base_sig
            results.append({"case": name, "knobs": combo, "kernel_ms": ms, "same_result": ok})
            print(f"{name:8s} {str({k: v for k, v in combo.items() if v}):40s} {ms:8.4f} ms  {npacks * M1 / ms / 1e6:8.1f} Grows/s{'' if ok else '  RESULT DIFFERS'}", flush=True)
        prog.close()
        json.dump(results, open(args.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
