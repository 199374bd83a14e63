"""The plan cache behind kx_scan / kx_scan_sharded: a repeated query over an unchanged store re-uses its translated
leaves and descriptors; any change of the store (kx_block_put / kx_block_drop), of the pack list, the outputs or the
program must be seen."""
import numpy as np
import pytest

import knoxdb_b200 as kb
import oracle as ko

pytestmark = pytest.mark.gpu


def test_repeated_scans_hit_the_cache_and_store_changes_invalidate_it():
    rng = np.random.default_rng(8)
    ctx = kb.Context(0)
    n, npacks = 20_000, 6
    vals = [rng.integers(0, 1000, n).astype(np.uint64) for _ in range(npacks)]
    amt = [rng.integers(-10**6, 10**6, n).astype(np.int64) for _ in range(npacks)]
    for p in range(npacks):
        ctx.block_put(p, 1, 1, kb.UINT64, ko.store("best", ko.U64, vals[p]))
        ctx.block_put(p, 1, 2, kb.INT64, ko.store("raw", ko.I64, amt[p]))
    prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.LT, 500)])
    refs = ctx.pack_refs([(p, 1) for p in range(npacks)])
    nrows = [n] * npacks

    def truth(vs, am):
        return [int((v < 500).sum()) for v in vs], sum(int(a[v < 500].sum()) for v, a in zip(vs, am)) & (2**64 - 1)

    want_counts, want_sum = truth(vals, amt)
    for _ in range(4):                                     # first call fills the cache, the others are served from it
        r = ctx.scan(prog, refs, nrows=nrows, aggs=[(2, kb.INT64)])
        assert r["counts"].tolist() == want_counts and r["aggs"][0].sum_bits == want_sum
    for _ in range(2):                                     # another output shape = another plan, not a stale one
        r = ctx.scan(prog, refs, nrows=nrows, want_bitsets=True)
        assert r["counts"].tolist() == want_counts
        assert all((b == np.packbits((v < 500).astype(np.uint8), bitorder="little")).all() for b, v in zip(r["bitsets"], vals))
    r = ctx.scan(prog, ctx.pack_refs([(p, 1) for p in range(npacks - 1, -1, -1)]), nrows=nrows)   # another pack order
    assert r["counts"].tolist() == want_counts[::-1]
    # replace a block: the cached plan must not be used any more
    vals[2] = rng.integers(0, 1000, n).astype(np.uint64)
    ctx.block_put(2, 1, 1, kb.UINT64, ko.store("best", ko.U64, vals[2]))
    want_counts, want_sum = truth(vals, amt)
    for _ in range(2):
        r = ctx.scan(prog, refs, nrows=nrows, aggs=[(2, kb.INT64)])
        assert r["counts"].tolist() == want_counts and r["aggs"][0].sum_bits == want_sum
    # a program compiled later (possibly at the same address) is another program
    prog.close()
    prog = kb.Program(ctx, [kb.Leaf(1, kb.UINT64, kb.GE, 500)])
    r = ctx.scan(prog, refs, nrows=nrows, aggs=[(2, kb.INT64)])
    assert r["counts"].tolist() == [n - c for c in want_counts]
    # the sharded entry point shares the cache
    ctx.comm_init(1, 0)
    for _ in range(3):
        s = ctx.scan_sharded(prog, refs, aggs=[(2, kb.INT64)])
        assert s["total_count"] == sum(n - c for c in want_counts) and s["counts"].tolist() == [n - c for c in want_counts]
    ctx.block_drop(0, 1, 1)
    with pytest.raises(kb.KnoxError):
        ctx.scan(prog, refs, nrows=nrows)
    prog.close()
    ctx.close()
