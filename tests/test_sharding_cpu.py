"""CPU tier: the N>1 path (pack sharding + one all-gather of 64 B partials + fixed-order combine)
with world_size 2 over gloo.  Per-rank partials come from the oracle here (no GPU in this tier); the
combine goes through the product's kx_agg_combine (a host-side function of the C ABI)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def make_packs(npacks):
    rng = np.random.default_rng(99)
    packs = []
    for p in range(npacks):
        n = 1000 + 37 * p
        packs.append((rng.integers(-10**12, 10**12, n).astype(np.int64), (rng.integers(0, 2**50, n) / 100.0).astype(np.float64),
                      np.packbits((rng.random(n) < 0.3).astype(np.uint8), bitorder="little")))
    return packs


def to_aggout(st, block_type):
    import knoxdb_b200 as kb
    from knoxdb_b200.lib import AggOut
    a = AggOut()
    a.count, a.sum_bits, a.min_bits, a.max_bits, a.valid = st.count, st.sum_bits, st.min_bits, st.max_bits, st.valid
    return a


def worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import knoxdb_b200 as kb
    import oracle as ko
    from knoxdb_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    packs = make_packs(11)
    lo, hi = shard.shard_range(len(packs), rank, world)
    si, sf = ko.Agg(), ko.Agg()
    for ints, flts, bits in packs[lo:hi]:
        si = ko.reduce(ko.I64, ints, bits, si)
        sf = ko.reduce(ko.F64, flts, bits, sf)
    out = shard.allgather_partials([to_aggout(si, kb.INT64), to_aggout(sf, kb.FLOAT64)], [kb.INT64, kb.FLOAT64], dist)
    if rank == 0:
        q.put([(o.count, o.sum_bits, o.min_bits, o.max_bits, o.valid) for o in out] + [(lo, hi)])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partial_exchange_matches_single_scan():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as ko
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    packs = make_packs(11)
    si, sf = ko.Agg(), ko.Agg()
    for ints, flts, bits in packs:
        si = ko.reduce(ko.I64, ints, bits, si)
        sf = ko.reduce(ko.F64, flts, bits, sf)
    gi, gf, rng0 = res
    assert rng0 == (0, 5)
    assert gi == (si.count, si.sum_bits, si.min_bits, si.max_bits, 1)           # integers: bit-exact
    assert gf[0] == sf.count and gf[2] == sf.min_bits and gf[3] == sf.max_bits
    a, b = np.uint64(gf[1]).view(np.float64), np.uint64(sf.sum_bits).view(np.float64)
    assert abs(a - b) <= 1e-12 * abs(b)                                         # float sum: north-star tolerance


def window_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import knoxdb_b200 as kb
    import oracle as ko
    from knoxdb_b200 import shard
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    packs, ts, edges = make_series_packs(9)
    lo, hi = shard.shard_range(len(packs), rank, world)
    si = sf = None
    nb = len(edges) - 1
    si, sf = (ko.Agg * nb)(), (ko.Agg * nb)()
    for (ints, flts, bits), t in zip(packs[lo:hi], ts[lo:hi]):
        si = ko.bucket_reduce(ko.I64, ints, ko.I64, t, bits, edges, si)
        sf = ko.bucket_reduce(ko.F64, flts, ko.I64, t, bits, edges, sf)
    out = shard.allgather_window_partials([[to_aggout(s, kb.INT64) for s in si], [to_aggout(s, kb.FLOAT64) for s in sf]], [kb.INT64, kb.FLOAT64], dist)
    if rank == 0:
        q.put([[(o.count, o.sum_bits, o.min_bits, o.max_bits, o.valid) for o in col] for col in out])
    dist.barrier()
    dist.destroy_process_group()


def make_series_packs(npacks):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as ko
    packs = make_packs(npacks)
    rng = np.random.default_rng(5)
    ts, t0 = [], 1_700_000_000
    for ints, _, _ in packs:
        t = (t0 + np.cumsum(rng.integers(0, 30, ints.size))).astype(np.int64)
        t0 = int(t[-1])
        ts.append(t)
    edges = ko.window_edges(int(ts[0][0]) + 11, int(ts[-1][-1]), 3600)
    return packs, ts, edges


def test_two_rank_window_partials_match_single_scan():
    """sharded series query: per-rank window tables (oracle) → one all-gather → per-window fixed-order combine"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle as ko
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=window_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gi, gf = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    packs, ts, edges = make_series_packs(9)
    nb = len(edges) - 1
    si, sf = (ko.Agg * nb)(), (ko.Agg * nb)()
    for (ints, flts, bits), t in zip(packs, ts):
        si = ko.bucket_reduce(ko.I64, ints, ko.I64, t, bits, edges, si)
        sf = ko.bucket_reduce(ko.F64, flts, ko.I64, t, bits, edges, sf)
    assert nb >= 3 and sum(s.count for s in si) > 0
    for k in range(nb):
        if not si[k].valid:
            assert gi[k][0] == 0 and gi[k][4] == 0
            continue
        assert gi[k] == (si[k].count, si[k].sum_bits, si[k].min_bits, si[k].max_bits, 1)
        assert gf[k][0] == sf[k].count and gf[k][2] == sf[k].min_bits and gf[k][3] == sf[k].max_bits
        a, b = np.uint64(gf[k][1]).view(np.float64), np.uint64(sf[k].sum_bits).view(np.float64)
        assert abs(a - b) <= 1e-12 * abs(b)


def test_shard_ranges_cover_all_packs():
    from knoxdb_b200 import shard
    for npacks in (0, 1, 7, 8, 1000, 7630):
        for world in (1, 2, 4, 8):
            rs = [shard.shard_range(npacks, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == npacks
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in rs) - min(h - l for l, h in rs) <= 1
