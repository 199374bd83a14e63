//go:build knoxgpu

// Package gpu binds libknoxgpu.so (include/knoxgpu.h), the B200 implementation of the pack scan
// path, behind KnoxDB's own interfaces.  There is no CPU fallback: Open fails with KX_ENODEV when
// no CUDA device is present and callers keep using the stock Go path in that case.
package gpu

/*
#cgo CFLAGS:  -I${SRCDIR}/../../third_party/knoxgpu/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/knoxgpu/lib -lknoxgpu
#include <stdlib.h>
#include "knoxgpu.h"
*/
import "C"

import (
	"errors"
	"math"
	"unsafe"

	"blockwatch.cc/knoxdb/internal/operator/filter"
	"blockwatch.cc/knoxdb/internal/types"
	"blockwatch.cc/knoxdb/internal/xroar"
)

// Context owns one device: a CUDA stream, the resident block store and scratch buffers.
// Calls on one Context are serialised inside the library; use one Context per GPU.
type Context struct{ h *C.kx_ctx }

// Open creates a context on CUDA device `device`; hbmBudget = 0 lets the library use 80% of free HBM.
func Open(device int, hbmBudget uint64) (*Context, error) {
	var h *C.kx_ctx
	if rc := C.kx_ctx_create(C.int(device), C.size_t(hbmBudget), &h); rc != 0 {
		return nil, errors.New(C.GoString(C.kx_last_error(nil)))
	}
	return &Context{h}, nil
}

func (c *Context) Close()     { C.kx_ctx_destroy(c.h); c.h = nil }
func (c *Context) err() error { return errors.New(C.GoString(C.kx_last_error(c.h))) }

// PutBlock registers the encoded bytes of one column block (what Container.Store wrote, i.e.
// block.Encode output minus the leading compression byte; internal/block/encode.go:194-226).
// Call it where Package.LoadFromDisk decodes a block (internal/pack/storage.go:128-190).
// The library copies the bytes; Go memory is not retained (cgo pointer rule).
func (c *Context) PutBlock(packKey, version uint32, fieldID uint16, t types.BlockType, enc []byte) (int, error) {
	var n C.uint32_t
	rc := C.kx_block_put(c.h, C.uint32_t(packKey), C.uint32_t(version), C.uint16_t(fieldID), C.uint8_t(t),
		unsafe.Pointer(unsafe.SliceData(enc)), C.size_t(len(enc)), &n)
	if rc != 0 {
		return 0, c.err()
	}
	return int(n), nil
}

func (c *Context) DropBlock(packKey, version uint32, fieldID uint16) error {
	if rc := C.kx_block_drop(c.h, C.uint32_t(packKey), C.uint32_t(version), C.uint16_t(fieldID)); rc != 0 {
		return c.err()
	}
	return nil
}

// Program is a compiled filter tree (filter.Node → postfix AND/OR program over typed leaves).
type Program struct {
	h   *C.kx_prog
	ctx *Context
}

func (p *Program) Close() { C.kx_prog_free(p.h); p.h = nil }

// pattern returns the 64-bit operand pattern the C ABI expects: sign-extended integers, IEEE bits
// for floats (include/knoxgpu.h, kx_leaf).
func pattern(v any) uint64 {
	switch x := v.(type) {
	case int64:
		return uint64(x)
	case int32:
		return uint64(int64(x))
	case int16:
		return uint64(int64(x))
	case int8:
		return uint64(int64(x))
	case uint64:
		return x
	case uint32:
		return uint64(x)
	case uint16:
		return uint64(x)
	case uint8:
		return uint64(x)
	case float64:
		return math.Float64bits(x)
	case float32:
		return uint64(math.Float32bits(x))
	}
	return 0
}

// flatten lists the members of an IN/NIN set.  Set members are already `uint64(v)` of the column
// value (internal/encode/int_raw.go:339-357).
func flatten(s *xroar.Bitmap) []uint64 {
	out := make([]uint64, 0, s.Count())
	it := s.NewIterator()
	for v, ok := it.Next(); ok; v, ok = it.Next() {
		out = append(out, v)
	}
	return out
}

// Compile walks a filter tree (internal/operator/filter/node.go:29-37) into the library's postfix form.
// Leaves use the field's storage id (Filter.Id) so that they address blocks registered with PutBlock.
func (c *Context) Compile(root *filter.Node) (*Program, error) {
	var (
		leaves []C.kx_leaf
		post   []C.uint8_t
		cbufs  []unsafe.Pointer // C copies of IN sets (kx_leaf must not hold Go pointers)
	)
	defer func() {
		for _, p := range cbufs {
			C.free(p)
		}
	}()
	var walk func(n *filter.Node) error
	walk = func(n *filter.Node) error {
		if n.IsLeaf() {
			f := n.Filter
			if len(leaves) == C.KX_MAX_LEAVES {
				return errors.New("knoxgpu: more than 8 filter leaves")
			}
			l := C.kx_leaf{field: C.uint16_t(f.Id), block_type: C.uint8_t(f.Type), mode: C.uint8_t(f.Mode)}
			if f.Type == types.BlockBytes && f.Mode != types.FilterModeIn && f.Mode != types.FilterModeNotIn {
				// byte-string leaf (bytesMatcher, internal/operator/filter/match_bytes.go): operand bytes travel in a C
				// copy behind `set`, nset = 1, a / b = lengths (RANGE: lower bound followed by the upper bound)
				var lo, hi []byte
				if f.Mode == types.FilterModeRange {
					rg := f.Value.([2]any)
					lo, hi = rg[0].([]byte), rg[1].([]byte)
				} else {
					lo = f.Value.([]byte)
				}
				p := C.malloc(C.size_t(len(lo) + len(hi) + 1))
				buf := unsafe.Slice((*byte)(p), len(lo)+len(hi)+1)
				copy(buf, lo)
				copy(buf[len(lo):], hi)
				cbufs = append(cbufs, p)
				l.set, l.nset, l.a, l.b = (*C.uint64_t)(p), 1, C.uint64_t(len(lo)), C.uint64_t(len(hi))
				post = append(post, C.uint8_t(len(leaves)))
				leaves = append(leaves, l)
				return nil
			}
			switch f.Mode {
			case types.FilterModeRange:
				rg := f.Value.([2]any)
				l.a, l.b = C.uint64_t(pattern(rg[0])), C.uint64_t(pattern(rg[1]))
			case types.FilterModeIn, types.FilterModeNotIn:
				set := flatten(f.Matcher.Value().(*xroar.Bitmap))
				if len(set) > 0 {
					p := C.malloc(C.size_t(8 * len(set)))
					copy(unsafe.Slice((*uint64)(p), len(set)), set)
					cbufs = append(cbufs, p)
					l.set, l.nset = (*C.uint64_t)(p), C.uint32_t(len(set))
				}
			default:
				l.a = C.uint64_t(pattern(f.Value))
			}
			post = append(post, C.uint8_t(len(leaves)))
			leaves = append(leaves, l)
			return nil
		}
		for i, ch := range n.Children {
			if err := walk(ch); err != nil {
				return err
			}
			if i > 0 {
				if n.OrKind {
					post = append(post, C.KX_OP_OR)
				} else {
					post = append(post, C.KX_OP_AND)
				}
			}
		}
		return nil
	}
	if err := walk(root); err != nil {
		return nil, err
	}
	var h *C.kx_prog
	if rc := C.kx_prog_compile(c.h, &leaves[0], C.int(len(leaves)), &post[0], C.int(len(post)), &h); rc != 0 {
		return nil, c.err()
	}
	return &Program{h: h, ctx: c}, nil
}

// AggOut mirrors kx_agg_out: the un-bucketed Count/Sum/Min/Max reducers of
// internal/reducer/reducer.go:138-314 over the matching rows of one value column.
type AggOut struct {
	Count            int64
	SumBits          uint64  // ints wrap in T like SumReducer[T]; float64: IEEE bits of the compensated sum
	SumErr           float64 // float64 only: residual of the compensated sum
	MinBits, MaxBits uint64
	Valid            bool // reducer Value() ok flag
}

// Scan evaluates prog over a batch of resident packs: per-pack bitsets (LSB-first,
// internal/bitset/bitset.go:23-29) at bits[offs[i]:], per-pack match counts and aggregates over the
// whole batch.  Replaces the per-pack loop of Reader.nextQueryMatch (internal/pack/table/reader.go:288-450).
func (c *Context) Scan(prog *Program, keys, versions []uint32, bits []byte, offs []uint64, counts []int64,
	aggFields []uint16, aggTypes []types.BlockType) ([]AggOut, error) {
	n := len(keys)
	if n == 0 {
		return nil, nil
	}
	refs := make([]C.kx_packref, n)
	for i := range refs {
		refs[i] = C.kx_packref{pack: C.uint32_t(keys[i]), version: C.uint32_t(versions[i])}
	}
	var (
		bp *C.uint8_t
		op *C.size_t
		cp *C.int64_t
		ar *C.kx_agg_req
		ao *C.kx_agg_out
	)
	if bits != nil {
		bp = (*C.uint8_t)(unsafe.SliceData(bits))
		op = (*C.size_t)(unsafe.Pointer(unsafe.SliceData(offs)))
	}
	if counts != nil {
		cp = (*C.int64_t)(unsafe.Pointer(unsafe.SliceData(counts)))
	}
	reqs := make([]C.kx_agg_req, len(aggFields))
	outs := make([]C.kx_agg_out, len(aggFields))
	for i := range reqs {
		reqs[i] = C.kx_agg_req{field: C.uint16_t(aggFields[i]), block_type: C.uint8_t(aggTypes[i])}
	}
	if len(reqs) > 0 {
		ar, ao = &reqs[0], &outs[0]
	}
	if rc := C.kx_scan(c.h, prog.h, &refs[0], C.int(n), bp, op, cp, ar, C.int(len(reqs)), ao); rc != 0 {
		return nil, c.err()
	}
	res := make([]AggOut, len(outs))
	for i, o := range outs {
		res[i] = AggOut{int64(o.count), uint64(o.sum_bits), float64(o.sum_err), uint64(o.min_bits), uint64(o.max_bits), o.valid != 0}
	}
	return res, nil
}

// CombineAgg merges per-shard partial aggregates in a fixed order (multi-GPU: one NCCL
// all-gather of the 64-byte partials, then this call on every rank; SURVEY §8e).
func CombineAgg(t types.BlockType, parts []AggOut) AggOut {
	in := make([]C.kx_agg_out, len(parts))
	for i, p := range parts {
		in[i] = C.kx_agg_out{count: C.int64_t(p.Count), sum_bits: C.uint64_t(p.SumBits), sum_err: C.double(p.SumErr),
			min_bits: C.uint64_t(p.MinBits), max_bits: C.uint64_t(p.MaxBits)}
		if p.Valid {
			in[i].valid = 1
		}
	}
	var o C.kx_agg_out
	if len(in) > 0 {
		C.kx_agg_combine(C.uint8_t(t), &in[0], C.int(len(in)), &o)
	}
	return AggOut{int64(o.count), uint64(o.sum_bits), float64(o.sum_err), uint64(o.min_bits), uint64(o.max_bits), o.valid != 0}
}
