//go:build knoxgpu

package gpu

import (
	"context"

	"blockwatch.cc/knoxdb/internal/bitset"
	"blockwatch.cc/knoxdb/internal/engine"
	"blockwatch.cc/knoxdb/internal/operator"
	"blockwatch.cc/knoxdb/internal/operator/filter"
	"blockwatch.cc/knoxdb/internal/pack"
	"blockwatch.cc/knoxdb/internal/types"
)

// BatchScan is a batching SOURCE: an operator.PullOperator (internal/operator/operator.go:31-35) that wraps the
// upstream source of a pipeline (PhysicalTableScan without a filter, internal/operator/table_scan.go:15-47) and
// replaces the pair "source → PhysicalFilter" (internal/operator/filter.go:29-37).
//
// Why a source and not a PushOperator: PhysicalPipeline.Execute (internal/operator/pipeline.go:103-161) pulls ONE
// pack per call, treats ResultMore from an operator as "no output yet" and drops out of the loop, and finalize
// only ever calls the sink.  An operator in the middle therefore cannot hold packs back and hand them out later:
// they would be lost.  A source can: Next pulls up to BatchSize packs from upstream, evaluates the filter for all
// of them with ONE kx_scan_ex call (selection vectors come back from the device), attaches the selections exactly
// like PhysicalFilter does (WithSelection(nil) when every row matches, else the ids), and then hands the packs out
// one per Next call.  Packs without a match are released and skipped — what Reader.nextQueryMatch does
// (internal/pack/table/reader.go:336-345).  After upstream reports ResultDone the remaining packs are drained
// before BatchScan itself reports ResultDone, so nothing is ever dropped.  BatchSize = 1 degenerates to the
// reference's per-pack behaviour.
//
// MaskFn (optional) supplies the reader's exclusion step for a pack (tombstones ∪ invisible rows, as "eligible"
// bits): reader.go:347-413.  Packs must have been registered with Context.PutBlock when they were loaded.
type BatchScan struct {
	ctx       *Context
	prog      *Program
	src       operator.PullOperator
	BatchSize int
	MaskFn    func(*pack.Package) []byte
	ready     []*pack.Package
	srcDone   bool
	err       error
	stats     QueryStats
}

var _ operator.PullOperator = (*BatchScan)(nil)

func NewBatchScan(ctx *Context, src operator.PullOperator, node *filter.Node, batch int) (*BatchScan, error) {
	prog, err := ctx.Compile(node)
	if err != nil {
		return nil, err
	}
	if batch < 1 {
		batch = 1
	}
	return &BatchScan{ctx: ctx, prog: prog, src: src, BatchSize: batch}, nil
}

func (op *BatchScan) Next(ctx context.Context) (*pack.Package, operator.Result) {
	for len(op.ready) == 0 {
		if op.srcDone {
			return nil, operator.ResultDone
		}
		if err := op.fill(ctx); err != nil {
			op.err = err
			return nil, operator.ResultError // errors are stored and fetched with Err(), like every operator
		}
	}
	p := op.ready[0]
	op.ready = op.ready[1:]
	return p, operator.ResultOK
}

// fill pulls the next batch from upstream and filters it on the device.
func (op *BatchScan) fill(ctx context.Context) error {
	batch := make([]*pack.Package, 0, op.BatchSize)
	for len(batch) < op.BatchSize && !op.srcDone {
		p, res := op.src.Next(ctx)
		switch res {
		case operator.ResultError:
			return op.src.Err()
		case operator.ResultDone:
			op.srcDone = true
		case operator.ResultOK:
			if p == nil {
				return operator.ErrNilPack
			}
			batch = append(batch, p)
		default:
			return operator.ErrTodo // a source never answers 'more' (pipeline.go:119-120)
		}
	}
	n := len(batch)
	if n == 0 {
		return nil
	}
	a := &ScanArgs{Keys: make([]uint32, n), Versions: make([]uint32, n), Counts: make([]int64, n), SelOff: make([]uint64, n+1)}
	rows := 0
	for i, p := range batch {
		a.Keys[i], a.Versions[i] = p.Key(), p.Version()
		rows += p.Len()
	}
	if op.MaskFn != nil {
		a.Masks = make([][]byte, n)
		for i, p := range batch {
			a.Masks[i] = op.MaskFn(p)
		}
	}
	a.Sel = make([]uint32, rows) // worst case: every row matches
	if _, _, err := op.ctx.ScanEx(op.prog, a); err != nil {
		return err
	}
	st := op.ctx.LastQueryStats()
	op.stats.RowsScanned += st.RowsScanned
	op.stats.PacksScanned += st.PacksScanned
	op.stats.RowsMatched += st.RowsMatched
	op.stats.ScanTimeNs += st.ScanTimeNs
	for i, p := range batch {
		switch cnt := int(a.Counts[i]); {
		case cnt == 0:
			p.Release() // no match: the reader skips such packs (reader.go:336-345)
		case cnt == p.Len():
			op.ready = append(op.ready, p.WithSelection(nil))
		default:
			op.ready = append(op.ready, p.WithSelection(a.Sel[a.SelOff[i]:a.SelOff[i+1]:a.SelOff[i+1]]))
		}
	}
	return nil
}

// Stats returns what the reference's reader feeds into query.QueryStats (internal/query/stats.go:15-60).
func (op *BatchScan) Stats() QueryStats { return op.stats }
func (op *BatchScan) Err() error       { return op.err }
func (op *BatchScan) Close() {
	for _, p := range op.ready {
		p.Release()
	}
	op.ready = nil
	op.src.Close()
	op.prog.Close()
}

// Filter is the per-pack PushOperator with the exact contract of PhysicalFilter (one pack in, the same pack out,
// ResultOK): a 1:1 replacement for pipelines that cannot change their source.  It never returns ResultMore and
// never holds a pack back, so PhysicalPipeline.Execute loses nothing; batching needs BatchScan.
type Filter struct {
	ctx  *Context
	prog *Program
	bits []byte
	err  error
}

var _ operator.PushOperator = (*Filter)(nil)

func NewFilter(ctx *Context, node *filter.Node) (*Filter, error) {
	prog, err := ctx.Compile(node)
	if err != nil {
		return nil, err
	}
	return &Filter{ctx: ctx, prog: prog}, nil
}

func (op *Filter) Process(_ context.Context, src *pack.Package) (*pack.Package, operator.Result) {
	if src == nil {
		op.err = operator.ErrNilPack
		return nil, operator.ResultError
	}
	nb := (((src.Len() + 7) >> 3) + 7) &^ 7
	if cap(op.bits) < nb {
		op.bits = make([]byte, nb)
	}
	counts := []int64{0}
	if _, err := op.ctx.Scan(op.prog, []uint32{src.Key()}, []uint32{src.Version()}, op.bits[:nb], []uint64{0}, counts, nil, nil); err != nil {
		op.err = err
		return nil, operator.ResultError
	}
	b := bitset.NewFromBytes(op.bits[:nb], src.Len()).ResetCount(int(counts[0]))
	if b.All() {
		src.WithSelection(nil)
	} else {
		src.WithSelection(b.Indexes(nil))
	}
	return src, operator.ResultOK
}

func (op *Filter) Finalize(_ context.Context) error { return nil }
func (op *Filter) Err() error                       { return op.err }
func (op *Filter) Close()                           { op.prog.Close() }

// AggSink is an engine.QueryResultConsumer (internal/engine/interface.go:154-157) that replaces
// StreamResult.Append + Bucket.Push + Reducer.Reduce (internal/query/result.go:96-152,
// internal/reducer/reducer.go:138-314) for the un-bucketed case: it collects the pack keys the reader
// yields and reduces the matching rows of the value columns on the device in Finish.
type AggSink struct {
	ctx    *Context
	prog   *Program
	fields []uint16
	typs   []types.BlockType
	keys   []uint32
	vers   []uint32
	masks  [][]byte
	MaskFn func(*engine.Package) []byte // optional: eligible-row bits of a pack (reader.go:347-413)
	rows   int
}

var _ engine.QueryResultConsumer = (*AggSink)(nil)

func NewAggSink(ctx *Context, prog *Program, fields []uint16, typs []types.BlockType) *AggSink {
	return &AggSink{ctx: ctx, prog: prog, fields: fields, typs: typs}
}

func (s *AggSink) Append(_ engine.Context, p *engine.Package) error {
	s.keys, s.vers = append(s.keys, p.Key()), append(s.vers, p.Version())
	if s.MaskFn != nil {
		s.masks = append(s.masks, s.MaskFn(p))
	}
	s.rows += p.Len()
	return nil
}

func (s *AggSink) Len() int { return s.rows }

// Finish runs the fused filter + reduce over every appended pack and returns one AggOut per value column.
// sharded = true combines the result over all ranks of the communicator (every rank must call Finish).
func (s *AggSink) Finish(sharded bool) ([]AggOut, int64, error) {
	return s.ctx.ScanEx(s.prog, &ScanArgs{Keys: s.keys, Versions: s.vers, Masks: s.masks, AggFields: s.fields, AggTypes: s.typs, Sharded: sharded})
}
