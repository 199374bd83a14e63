#!/usr/bin/env python3
"""tune_geometry.py — sweep the scan kernel's tile geometry (CTAs per SM, ring depth, groups per warp)
per column width through the KX_SCAN_GEOMETRY hook and print the device time of each combination.
Used to choose the defaults in kx_api.cu (run_scan); not part of the product path.

usage (GPU box): python profiles/tune_geometry.py --out gpurun_out/tune.json [--widths 8,20,64]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))

import knoxdb_b200 as kb            # noqa: E402
from sweep_configs import bitpack_block, raw_block   # noqa: E402

M4 = 1 << 22
MAXDYN = 200 * 1024


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="gpurun_out/tune.json")
    ap.add_argument("--widths", default="4,8,12,16,20,24,32,40,64")
    ap.add_argument("--bitsets", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(3)
    ctx = kb.Context(0)
    results = []
    for w in [int(x) for x in args.widths.split(",")]:
        if w == 64:
            blocks = [raw_block(rng.integers(0, 2**60, M4, dtype=np.uint64)) for _ in range(2)]
            leaf = kb.Leaf(1, kb.UINT64, kb.RANGE, 1 << 58, 3 << 58)
        else:
            blocks = [bitpack_block(rng, M4, w) for _ in range(2)]
            leaf = kb.Leaf(1, kb.UINT64, kb.LT, 1000 + (1 << (w - 1)))
        npacks = int(min(512, max(32, 1.4e9 // (M4 * w / 8))))
        pinned = []
        for b in blocks:
            h = ctx.host_array(b.size); h[:] = b; pinned.append(h)
        for p in range(npacks):
            ctx.block_put(p, 1, 1, kb.UINT64, pinned[p % 2])
        prog = kb.Program(ctx, [leaf])
        refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
        nrows = [M4] * npacks
        kw = dict(nrows=nrows, want_bitsets=args.bitsets)
        if args.bitsets:
            kw["bitset_buf"] = ctx.host_array(ctx.bitset_layout(nrows)[1])
        combos = [None]
        for c in (1, 2, 3):
            for st in (2, 3, 4, 6, 8):
                for R in (8, 16, 32, 64, 96, 128, 256):
                    stage = (32 * R * w + 32 + 127) // 128 * 128
                    if c * (128 + st * stage) <= MAXDYN and st * stage * c >= 96 * 1024:
                        combos.append((c, st, R))
        base = None
        for combo in combos:
            if combo is None:
                os.environ.pop("KX_SCAN_GEOMETRY", None)
            else:
                os.environ["KX_SCAN_GEOMETRY"] = "%d,%d,%d" % combo
            want = None
            ks = []
            for i in range(6):
                r = ctx.scan(prog, refs, **kw)
                tot = int(r["counts"].sum())
                if base is None:
                    base = tot
                assert tot == base, (w, combo, tot, base)
                if i:
                    ks.append(ctx.last_scan_stats()["kernel_ms"])
            km = float(np.median(ks))
            gbs = npacks * M4 * (w + (1 if args.bitsets else 0)) / 8 / (km * 1e-3) / 1e9
            results.append({"w": w, "geometry": combo, "kernel_ms": km, "GBps": gbs, "Grows_per_s": npacks * M4 / km / 1e6})
            print(f"w={w:2d} geo={str(combo):>14s} {km:8.4f} ms {gbs:8.1f} GB/s {npacks * M4 / km / 1e6:9.1f} Grows/s", flush=True)
        os.environ.pop("KX_SCAN_GEOMETRY", None)
        prog.close()
        for p in range(npacks):
            ctx.block_drop(p, 1, 1)
        ctx.free_host_arrays()
        json.dump(results, open(args.out, "w"), indent=1)
    ctx.close()


if __name__ == "__main__":
    main()
