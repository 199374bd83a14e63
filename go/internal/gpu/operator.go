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

// Filter is a drop-in operator.PushOperator (internal/operator/operator.go:31-42) for PhysicalFilter
// (internal/operator/filter.go:29-37).  PhysicalFilter calls filter.Match per pack; this operator batches
// BatchSize packs per kx_scan call so that one kernel launch covers the whole batch, then applies the same
// post-processing (All → WithSelection(nil), else WithSelection(bits.Indexes(nil))) and hands the packs
// downstream one by one.  Packs must have been registered with Context.PutBlock when they were loaded.
type Filter struct {
	ctx       *Context
	prog      *Program
	BatchSize int
	batch     []*pack.Package
	ready     []*pack.Package
	err       error
}

var _ operator.PushOperator = (*Filter)(nil)

func NewFilter(ctx *Context, node *filter.Node, batch int) (*Filter, error) {
	prog, err := ctx.Compile(node)
	if err != nil {
		return nil, err
	}
	return &Filter{ctx: ctx, prog: prog, BatchSize: batch}, nil
}

func (op *Filter) Process(_ context.Context, src *pack.Package) (*pack.Package, operator.Result) {
	if src == nil {
		op.err = operator.ErrNilPack
		return nil, operator.ResultError
	}
	op.batch = append(op.batch, src)
	if len(op.batch) >= op.BatchSize {
		if err := op.flush(); err != nil {
			op.err = err
			return nil, operator.ResultError // errors are stored and fetched with Err(), like every operator
		}
	}
	return op.pop()
}

// pop hands out one finished pack; ResultMore tells the pipeline to call again without new input.
func (op *Filter) pop() (*pack.Package, operator.Result) {
	if len(op.ready) == 0 {
		return nil, operator.ResultMore
	}
	p := op.ready[0]
	op.ready = op.ready[1:]
	if len(op.ready) > 0 {
		return p, operator.ResultMore
	}
	return p, operator.ResultOK
}

func (op *Filter) flush() error {
	n := len(op.batch)
	if n == 0 {
		return nil
	}
	keys, vers := make([]uint32, n), make([]uint32, n)
	offs, counts := make([]uint64, n), make([]int64, n)
	total := 0
	for i, p := range op.batch {
		keys[i], vers[i] = p.Key(), p.Version()
		offs[i] = uint64(total)
		total += (((p.Len() + 7) >> 3) + 7) &^ 7 // offsets must be multiples of 8
	}
	bits := make([]byte, total)
	if _, err := op.ctx.Scan(op.prog, keys, vers, bits, offs, counts, nil, nil); err != nil {
		return err
	}
	for i, p := range op.batch {
		b := bitset.NewFromBytes(bits[offs[i]:], p.Len()).ResetCount(int(counts[i]))
		if b.All() {
			p.WithSelection(nil)
		} else {
			p.WithSelection(b.Indexes(nil))
		}
		op.ready = append(op.ready, p)
	}
	op.batch = op.batch[:0]
	return nil
}

func (op *Filter) Finalize(_ context.Context) error { return op.flush() }
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
	rows   int
}

var _ engine.QueryResultConsumer = (*AggSink)(nil)

func NewAggSink(ctx *Context, prog *Program, fields []uint16, typs []types.BlockType) *AggSink {
	return &AggSink{ctx: ctx, prog: prog, fields: fields, typs: typs}
}

func (s *AggSink) Append(_ engine.Context, p *engine.Package) error {
	s.keys, s.vers = append(s.keys, p.Key()), append(s.vers, p.Version())
	s.rows += p.Len()
	return nil
}

func (s *AggSink) Len() int { return s.rows }

// Finish runs the fused filter + reduce over every appended pack and returns one AggOut per value column.
func (s *AggSink) Finish() ([]AggOut, error) {
	return s.ctx.Scan(s.prog, s.keys, s.vers, nil, nil, nil, s.fields, s.typs)
}
