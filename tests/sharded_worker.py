"""One rank of the 2+ GPU sharded-scan test (tests/test_gpu_sharded.py): registers its shard of the packs on its own
GPU, joins the library's NCCL communicator and runs kx_scan_sharded; prints one JSON line with what it got."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import knoxdb_b200 as kb   # noqa: E402
import oracle as ko        # noqa: E402  (encoder of the synthetic blocks)
from knoxdb_b200 import shard   # noqa: E402


def make_packs(npacks, n, seed=5):
    rng = np.random.default_rng(seed)
    packs = []
    for p in range(npacks):
        ts = (1_700_000_000 + 1000 * p + np.cumsum(rng.integers(0, 3, n))).astype(np.int64)
        acct = rng.integers(0, 50, n).astype(np.uint64)
        amt = rng.integers(-10**6, 10**6, n).astype(np.int64)
        val = (rng.integers(0, 2**30, n).astype(np.float64) / 64.0)
        packs.append((ts, acct, amt, val))
    return packs


def main():
    rank, nranks, idfile, npacks, n = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    ctx = kb.Context(rank)
    if rank == 0:
        cid = kb.Context.comm_unique_id()
        tmp = idfile + ".tmp"
        cid.tofile(tmp)
        os.replace(tmp, idfile)
    else:
        t0 = time.time()
        while not os.path.exists(idfile):
            if time.time() - t0 > 120:
                raise SystemExit("no communicator id")
            time.sleep(0.05)
        cid = np.fromfile(idfile, dtype=np.uint8)
    ctx.comm_init(nranks, rank, cid)
    packs = make_packs(npacks, n)
    lo, hi = shard.shard_range(npacks, rank, nranks)
    for p in range(lo, hi):
        ts, acct, amt, val = packs[p]
        ctx.block_put(p, 1, 1, kb.INT64, ko.store("best", ko.I64, ts))
        ctx.block_put(p, 1, 2, kb.UINT64, ko.store("best", ko.U64, acct))
        ctx.block_put(p, 1, 3, kb.INT64, ko.store("best", ko.I64, amt))
        ctx.block_put(p, 1, 4, kb.FLOAT64, ko.store("raw", ko.F64, val))
    t_lo, t_hi = int(packs[0][0][n // 3]), int(packs[-1][0][n // 2])
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(2, kb.UINT64, kb.IN, values=np.arange(0, 50, 3, dtype=np.uint64))])
    out = []
    for rep in range(3):
        r = ctx.scan_sharded(prog, [(p, 1) for p in range(lo, hi)], aggs=[(3, kb.INT64), (4, kb.FLOAT64)])
        out.append({"total": r["total_count"], "local": int(r["counts"].sum()) if hi > lo else 0,
                    "aggs": [[g.count, g.sum_bits, g.min_bits, g.max_bits, g.valid] for g in r["aggs"]]})
    g = ctx.comm_allgather(np.asarray([rank * 7 + 1, 99], dtype=np.uint64))
    info = ctx.comm_info()
    print(json.dumps({"rank": rank, "runs": out, "gather": g.view(np.uint64).reshape(nranks, 2).tolist(), "info": info, "range": [lo, hi]}), flush=True)
    prog.close()
    ctx.close()


if __name__ == "__main__":
    main()
