#!/usr/bin/env python3
"""ncu_cases.py — tiny driver that runs ONE sweep case a few times so that ncu can capture its kernels.

usage (GPU box):
  ncu --set full --import-source on --clock-control none -k regex:scan_kernel --launch-skip 2 -c 1 \\
      -o gpurun_out/prof_<name> python profiles/ncu_cases.py --match "<substring of the case name>" --only c2
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "profiles"))

import knoxdb_b200 as kb                      # noqa: E402
import sweep_configs as sc                    # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--match", required=True)
    ap.add_argument("--only", default="")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--packs", type=int, default=0, help="override the case's pack count (smaller = faster ncu replay)")
    args = ap.parse_args()
    rng = np.random.default_rng(1)
    cases = [c for c in sc.build_cases(rng, set(x for x in args.only.split(",") if x)) if args.match in c.name]
    assert cases, "no case matches"
    case = cases[0]
    if args.packs:
        case.npacks = args.packs
    ctx = kb.Context(0)
    for f, (kbt, blocks, _) in case.fields.items():
        pinned = []
        for b in blocks:
            h = ctx.host_array(b.size); h[:] = b; pinned.append(h)
        for p in range(case.npacks):
            ctx.block_put(p, 1, f, kbt, pinned[p % len(pinned)])
    prog = kb.Program(ctx, case.leaves, case.postfix)
    refs = ctx.pack_refs([(p, 1) for p in range(case.npacks)])
    nrows = [case.nrows] * case.npacks
    kw = dict(nrows=nrows, want_bitsets=case.bitsets, aggs=case.aggs)
    if case.bitsets:
        kw["bitset_buf"] = ctx.host_array(ctx.bitset_layout(nrows)[1])
    for _ in range(args.reps):
        ctx.scan(prog, refs, **kw)
    st = ctx.last_scan_stats()
    print(case.name, "kernel_ms", st["kernel_ms"], "launches", st["launches"])
    prog.close()
    ctx.close()


if __name__ == "__main__":
    main()
