"""The operator-pipeline contract of the reference (PhysicalPipeline.Execute, internal/operator/pipeline.go:103-161),
mirrored in tests/pipeline_harness.py and driven through the C ABI: the batching source (gpu.BatchScan) and the
per-pack push filter (gpu.Filter) of go/internal/gpu/operator.go deliver every matching pack with the selection vector
PhysicalFilter would attach (internal/operator/filter.go:29-37), for every batch size; a PushOperator that holds packs
back — the adapter this replaces — demonstrably loses packs under the same pipeline."""
import numpy as np
import pytest

import knoxdb_b200 as kb
import oracle as ko
import pipeline_harness as ph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def table():
    rng = np.random.default_rng(99)
    ctx = kb.Context(0)
    sizes = [4096, 70_001, 1, 8192, 333, 16_384, 50_000, 64, 9_999, 20_000, 31, 65_536, 5_000]
    packs, want = [], {}
    t_lo, t_hi = 1_700_000_500, 1_700_040_000
    for i, n in enumerate(sizes):
        key = 9000 + i
        ts = (1_700_000_000 + 7_000 * i + np.cumsum(rng.integers(0, 2, n))).astype(np.int64)
        k = rng.integers(0, 100, n).astype(np.uint64)
        if i == 4:
            ts[:] = 1_700_001_000; k[:] = 3        # every row matches: WithSelection(nil)
        b_ts, b_k = ko.store("best", ko.I64, ts), ko.store("best", ko.U64, k)
        ctx.block_put(key, 1, 1, kb.INT64, b_ts); ctx.block_put(key, 1, 2, kb.UINT64, b_k)
        l0 = ko.Container(ko.I64, b_ts).match(ko.RG, ko.scalar_u64(ko.I64, t_lo), ko.scalar_u64(ko.I64, t_hi))
        l1 = ko.Container(ko.U64, b_k).match(ko.LT, 10, 0)
        bits = np.unpackbits(ko.tree_eval([0, 1, 0xFE], [l0, l1], n), bitorder="little")[:n]
        want[key] = np.flatnonzero(bits).astype(np.uint32)
        packs.append((key, 1, n))
    prog = kb.Program(ctx, [kb.Leaf(1, kb.INT64, kb.RANGE, t_lo, t_hi), kb.Leaf(2, kb.UINT64, kb.LT, 10)])
    yield ctx, prog, packs, want
    prog.close()
    ctx.close()


def _check(got, packs, want, expect_empty_packs):
    keys = [p.key for p in got]
    expected = [k for k, _, n in packs if expect_empty_packs or want[k].size]
    assert keys == expected, (keys, expected)
    for p in got:
        w = want[p.key]
        if w.size == p.n:
            assert p.selection is None                      # bits.All() → WithSelection(nil)
        else:
            assert isinstance(p.selection, np.ndarray) and (p.selection == w).all(), p.key


@pytest.mark.parametrize("batch", [1, 2, 3, 5, 13, 64])
def test_batching_source_delivers_every_matching_pack(table, batch):
    ctx, prog, packs, want = table
    assert any(w.size == 0 for w in want.values()) and any(w.size == n for (k, _, n), w in zip(packs, want.values()))
    src = ph.TableSource([ph.Pack(*p) for p in packs])
    scan = ph.BatchScan(ctx, prog, src, batch)
    sink = ph.CollectSink()
    ph.Pipeline(scan, [], sink).run()
    _check(sink.got, packs, want, expect_empty_packs=False)
    assert sink.finalized and scan.calls == -(-len(packs) // batch)      # one device call per batch
    assert all(p.released for p in src.packs if want[p.key].size == 0)   # packs without a match are released, like the reader does


def test_limit_in_the_sink_finalizes_and_releases_what_is_left(table):
    ctx, prog, packs, want = table
    src = ph.TableSource([ph.Pack(*p) for p in packs])
    scan = ph.BatchScan(ctx, prog, src, 8)
    sink = ph.CollectSink(limit=2)
    ph.Pipeline(scan, [], sink).run()
    assert len(sink.got) == 2 and sink.finalized
    scan.close()
    assert src.closed and all(p.released for p in src.packs[:8] if p not in sink.got)


def test_per_pack_push_filter_keeps_the_reference_contract(table):
    ctx, prog, packs, want = table
    sink = ph.CollectSink()
    ph.Pipeline(ph.TableSource([ph.Pack(*p) for p in packs]), [ph.PushFilter(ctx, prog)], sink).run()
    got = sink.got
    assert [p.key for p in got] == [k for k, _, _ in packs]        # PhysicalFilter passes every pack on, also without a match
    for p in got:
        w = want[p.key]
        assert (p.selection is None) if w.size == p.n else (p.selection == w).all()


def test_a_push_operator_that_holds_packs_back_loses_them(table):
    """why the round-1 adapter was replaced: under Execute's ResultMore semantics only a fraction of the packs arrives"""
    ctx, prog, packs, want = table
    sink = ph.CollectSink()
    ph.Pipeline(ph.TableSource([ph.Pack(*p) for p in packs]), [ph.HoldingFilter(ctx, prog, 4)], sink).run()
    assert len(sink.got) < len(packs) // 2
