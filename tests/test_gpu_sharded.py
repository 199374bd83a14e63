"""Pack-sharded scans through the library's own communicator (kx_comm_init / kx_scan_sharded): single-rank semantics on
one GPU, and — when the box has at least two GPUs — real NCCL ranks in separate processes."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import knoxdb_b200 as kb
import oracle as ko

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _setup(ctx, npacks=5, n=70_001, seed=3):
    rng = np.random.default_rng(seed)
    truth = []
    for p in range(npacks):
        k = rng.integers(0, 1000, n).astype(np.uint64)
        v = rng.integers(-10**9, 10**9, n).astype(np.int64)
        f = rng.integers(0, 2**40, n).astype(np.float64) / 100.0
        ctx.block_put(p, 1, 1, kb.UINT64, ko.store("best", ko.U64, k))
        ctx.block_put(p, 1, 2, kb.INT64, ko.store("raw", ko.I64, v))
        ctx.block_put(p, 1, 3, kb.FLOAT64, ko.store("raw", ko.F64, f))
        truth.append((k, v, f))
    return truth


def test_scan_sharded_on_one_rank_equals_scan_and_reports_query_stats():
    ctx = kb.Context(0)
    ctx.comm_init(1, 0)
    assert ctx.comm_info()["nranks"] == 1
    truth = _setup(ctx)
    prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.LT, 300)])
    refs = [(p, 1) for p in range(len(truth))]
    aggs = [(2, kb.INT64), (3, kb.FLOAT64)]
    a = ctx.scan(prog, refs, nrows=[t[0].size for t in truth], aggs=aggs)
    q = ctx.last_query_stats()
    b = ctx.scan_sharded(prog, refs, aggs=aggs)
    assert b["counts"].tolist() == a["counts"].tolist()
    assert b["total_count"] == int(a["counts"].sum()) == sum(int((t[0] < 300).sum()) for t in truth)
    for x, y in zip(a["aggs"], b["aggs"]):
        assert (x.count, x.sum_bits, x.min_bits, x.max_bits, x.valid) == (y.count, y.sum_bits, y.min_bits, y.max_bits, y.valid)
    assert q["rows_scanned"] == sum(t[0].size for t in truth) and q["packs_scanned"] == len(truth) and q["rows_matched"] == b["total_count"]
    assert q["kernel_launches"] >= 1 and q["scan_time_ns"] > 0
    # an empty shard still answers (and takes part in the collective)
    e = ctx.scan_sharded(prog, [], aggs=aggs)
    assert e["total_count"] == 0 and all(g.valid == 0 and g.count == 0 for g in e["aggs"])
    g = ctx.comm_allgather(np.arange(5, dtype=np.uint32))
    assert g.shape == (1, 20) and g.view(np.uint32).tolist() == [[0, 1, 2, 3, 4]]
    prog.close()
    ctx.close()


@pytest.mark.skipif(kb.lib().kx_device_count() < 2, reason="needs two GPUs")
def test_scan_sharded_over_nccl_ranks_matches_one_gpu():
    nranks = min(kb.lib().kx_device_count(), 4)
    npacks, n = 7, 50_001
    with tempfile.TemporaryDirectory() as d:
        idfile = os.path.join(d, "comm_id")
        procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "sharded_worker.py"), str(r), str(nranks), idfile, str(npacks), str(n)],
                                  stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for r in range(nranks)]
        outs = []
        for p in procs:
            so, se = p.communicate(timeout=600)
            assert p.returncode == 0, se[-2000:]
            outs.append(json.loads(so.strip().splitlines()[-1]))
    # the same query over all packs on one GPU
    sys.path.insert(0, HERE)
    from sharded_worker import make_packs
    packs = make_packs(npacks, n)
    ctx = kb.Context(0)
    for p, (ts, acct, amt, val) in enumerate(packs):
        ctx.block_put(p, 1, 1, kb.INT64, ko.store("best", ko.I64, ts))
        ctx.block_put(p, 1, 2, kb.UINT64, ko.store("best", ko.U64, acct))
        ctx.block_put(p, 1, 3, kb.INT64, ko.store("best", ko.I64, amt))
        ctx.block_put(p, 1, 4, kb.FLOAT64, ko.store("raw", ko.F64, val))
    t_lo, t_hi = int(packs[0][0][n // 3]), int(packs[-1][0][n // 2])
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(2, kb.UINT64, kb.IN, values=np.arange(0, 50, 3, dtype=np.uint64))])
    one = ctx.scan(prog, [(p, 1) for p in range(npacks)], nrows=[n] * npacks, aggs=[(3, kb.INT64), (4, kb.FLOAT64)])
    total = int(one["counts"].sum())
    assert total > 0
    for o in outs:
        assert o["info"]["nranks"] == nranks and o["info"]["nccl_version"] > 0
        assert o["gather"] == [[r * 7 + 1, 99] for r in range(nranks)]
        for run in o["runs"]:
            assert run["total"] == total
            assert run["aggs"] == outs[0]["runs"][0]["aggs"], "ranks disagree / runs are not reproducible"
    assert sum(o["runs"][0]["local"] for o in outs) == total
    gi, gf = outs[0]["runs"][0]["aggs"]
    oi, of = one["aggs"]
    assert gi == [oi.count, oi.sum_bits, oi.min_bits, oi.max_bits, oi.valid]            # integers: bit-exact however the packs are sharded
    assert (gf[0], gf[2], gf[3]) == (of.count, of.min_bits, of.max_bits)
    a, b = float(np.uint64(gf[1]).view(np.float64)), float(np.uint64(of.sum_bits).view(np.float64))
    assert abs(a - b) <= 1e-12 * abs(b)
    prog.close()
    ctx.close()
