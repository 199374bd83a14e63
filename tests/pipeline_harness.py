"""A Python mirror of the reference's operator pipeline contract, used to TEST the call sequence the Go adapters of
go/internal/gpu/operator.go make through the C ABI (Go is not installed here, so the adapters themselves are reviewed,
not compiled; this harness runs the same state machines against the same library).

  * Pipeline.execute  = PhysicalPipeline.Execute  (internal/operator/pipeline.go:103-161), including what it does with
                        ResultMore from an operator (drop out: the pack is gone) and that finalize only reaches the sink
  * TableSource       = PhysicalTableScan.Next     (internal/operator/table_scan.go:28-38)
  * BatchScan         = gpu.BatchScan              (go/internal/gpu/operator.go): a batching PullOperator
  * PushFilter        = gpu.Filter                 (per-pack PushOperator with PhysicalFilter's contract)
  * HoldingFilter     = the round-1 adapter that held packs back inside a PushOperator — kept to show that the
                        pipeline loses packs with it (the test asserts the loss, i.e. why it was replaced)
"""
import numpy as np

OK, MORE, DONE, ERROR = range(4)   # operator.Result (internal/operator/operator.go:13-21)


class Pack:
    def __init__(self, key, version, n):
        self.key, self.version, self.n = key, version, n
        self.selection = "unset"      # None = all rows (WithSelection(nil)); array = selected row ids
        self.released = False

    def with_selection(self, sel):
        self.selection = sel
        return self

    def release(self):
        self.released = True


class TableSource:
    def __init__(self, packs):
        self.packs, self.i, self.closed = list(packs), 0, False

    def next(self):
        if self.i >= len(self.packs):
            return None, DONE
        p = self.packs[self.i]
        self.i += 1
        return p, OK

    def close(self):
        self.closed = True


class BatchScan:
    """PullOperator: pulls up to batch_size packs from upstream, ONE kx_scan_ex for all of them, hands them out one by one"""

    def __init__(self, ctx, prog, src, batch_size, mask_fn=None):
        self.ctx, self.prog, self.src, self.batch_size, self.mask_fn = ctx, prog, src, max(1, batch_size), mask_fn
        self.ready, self.src_done, self.calls = [], False, 0

    def next(self):
        while not self.ready:
            if self.src_done:
                return None, DONE
            self._fill()
        return self.ready.pop(0), OK

    def _fill(self):
        batch = []
        while len(batch) < self.batch_size and not self.src_done:
            p, res = self.src.next()
            if res == DONE:
                self.src_done = True
            elif res == OK:
                assert p is not None
                batch.append(p)
            else:
                raise RuntimeError("source error")
        if not batch:
            return
        masks = None if self.mask_fn is None else [self.mask_fn(p) for p in batch]
        r = self.ctx.scan_ex(self.prog, [(p.key, p.version) for p in batch], nrows=[p.n for p in batch], masks=masks, want_sel=True,
                             sel_cap=sum(p.n for p in batch))
        self.calls += 1
        for i, p in enumerate(batch):
            cnt = int(r["counts"][i])
            if cnt == 0:
                p.release()
            elif cnt == p.n:
                self.ready.append(p.with_selection(None))
            else:
                self.ready.append(p.with_selection(r["sel"][int(r["sel_off"][i]):int(r["sel_off"][i + 1])].copy()))

    def close(self):
        for p in self.ready:
            p.release()
        self.ready = []
        self.src.close()


class PushFilter:
    """PushOperator with PhysicalFilter's contract: one pack in, the same pack out, ResultOK"""

    def __init__(self, ctx, prog):
        self.ctx, self.prog = ctx, prog

    def process(self, p):
        r = self.ctx.scan(self.prog, [(p.key, p.version)], nrows=[p.n], want_bitsets=True)
        bits = np.unpackbits(r["bitsets"][0], bitorder="little")[:p.n]
        p.with_selection(None if bits.all() else np.flatnonzero(bits).astype(np.uint32))
        return p, OK

    def finalize(self):
        return None


class HoldingFilter:
    """the round-1 shape: a PushOperator that collects a batch and answers ResultMore meanwhile"""

    def __init__(self, ctx, prog, batch_size):
        self.ctx, self.prog, self.batch_size, self.batch, self.ready = ctx, prog, batch_size, [], []

    def process(self, p):
        self.batch.append(p)
        if len(self.batch) >= self.batch_size:
            self._flush()
        if not self.ready:
            return None, MORE
        q = self.ready.pop(0)
        return q, (MORE if self.ready else OK)

    def _flush(self):
        if self.batch:
            r = self.ctx.scan_ex(self.prog, [(p.key, p.version) for p in self.batch], nrows=[p.n for p in self.batch], want_sel=True,
                                 sel_cap=sum(p.n for p in self.batch))
            for i, p in enumerate(self.batch):
                self.ready.append(p.with_selection(r["sel"][int(r["sel_off"][i]):int(r["sel_off"][i + 1])].copy()))
            self.batch = []

    def finalize(self):   # never called by the pipeline: finalize only reaches the sink (pipeline.go:169-175)
        self._flush()


class CollectSink:
    def __init__(self, limit=None):
        self.got, self.limit, self.finalized = [], limit, False

    def process(self, p):
        self.got.append(p)
        if self.limit is not None and len(self.got) >= self.limit:
            return None, DONE
        return None, OK

    def finalize(self):
        self.finalized = True


class Pipeline:
    """PhysicalPipeline.Execute, branch for branch (internal/operator/pipeline.go:103-161)"""

    def __init__(self, src, ops, sink):
        self.src, self.ops, self.sink, self.done = src, ops, sink, False

    def execute(self):
        if self.done:
            return
        pkg, res = self.src.next()
        if res == DONE:
            return self._finalize()
        assert res == OK and pkg is not None, "unexpected source result"
        for op in self.ops:
            pkg, res = op.process(pkg)
            if res == MORE:
                return                       # "operator needs more data to output a pack": the pack in hand is dropped
            if res == DONE:
                return self._finalize()
            assert res == OK and pkg is not None
        _, res = self.sink.process(pkg)
        if res == DONE:
            return self._finalize()

    def _finalize(self):
        self.sink.finalize()                  # only the sink: operators in the middle are never finalized
        self.done = True

    def run(self, max_steps=100000):
        steps = 0
        while not self.done:
            self.execute()
            steps += 1
            assert steps < max_steps
        return steps
