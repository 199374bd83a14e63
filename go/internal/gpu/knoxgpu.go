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
	"encoding/binary"
	"errors"
	"fmt"
	"math"
	"runtime"
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
// for floats (include/knoxgpu.h, kx_leaf).  Operand types the library does not scan (int, uint, bool,
// time.Time, named types, 128/256-bit integers) are an error: a filter must never run with a zero operand
// because its value could not be translated.
func pattern(v any) (uint64, error) {
	switch x := v.(type) {
	case int64:
		return uint64(x), nil
	case int32:
		return uint64(int64(x)), nil
	case int16:
		return uint64(int64(x)), nil
	case int8:
		return uint64(int64(x)), nil
	case uint64:
		return x, nil
	case uint32:
		return uint64(x), nil
	case uint16:
		return uint64(x), nil
	case uint8:
		return uint64(x), nil
	case float64:
		return math.Float64bits(x), nil
	case float32:
		return uint64(math.Float32bits(x)), nil
	}
	return 0, fmt.Errorf("knoxgpu: unsupported filter operand type %T", v)
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
				var ok bool
				if f.Mode == types.FilterModeRange {
					rg, isRange := f.Value.([2]any)
					if !isRange {
						return fmt.Errorf("knoxgpu: range filter carries %T", f.Value)
					}
					lo, ok = rg[0].([]byte)
					if ok {
						hi, ok = rg[1].([]byte)
					}
				} else {
					lo, ok = f.Value.([]byte)
				}
				if !ok {
					return fmt.Errorf("knoxgpu: byte-string filter carries %T", f.Value)
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
			if f.Type == types.BlockBytes {
				// IN / NOT IN on byte strings (bytesInSetMatcher / bytesNotInSetMatcher, match_bytes.go:392-520): the set
				// travels as nset little-endian uint32 lengths followed by the concatenated bytes, a = buffer size
				vals, ok := f.Value.([][]byte)
				if !ok || len(vals) == 0 {
					return fmt.Errorf("knoxgpu: byte-string set filter carries %T", f.Value)
				}
				size := 4 * len(vals)
				for _, v := range vals {
					size += len(v)
				}
				p := C.malloc(C.size_t(size + 1))
				buf := unsafe.Slice((*byte)(p), size+1)
				pos := 4 * len(vals)
				for i, v := range vals {
					binary.LittleEndian.PutUint32(buf[4*i:], uint32(len(v)))
					pos += copy(buf[pos:], v)
				}
				cbufs = append(cbufs, p)
				l.set, l.nset, l.a = (*C.uint64_t)(p), C.uint32_t(len(vals)), C.uint64_t(size)
				post = append(post, C.uint8_t(len(leaves)))
				leaves = append(leaves, l)
				return nil
			}
			switch f.Mode {
			case types.FilterModeRange:
				rg, ok := f.Value.([2]any)
				if !ok {
					return fmt.Errorf("knoxgpu: range filter carries %T", f.Value)
				}
				lo, err := pattern(rg[0])
				if err != nil {
					return err
				}
				hi, err := pattern(rg[1])
				if err != nil {
					return err
				}
				l.a, l.b = C.uint64_t(lo), C.uint64_t(hi)
			case types.FilterModeIn, types.FilterModeNotIn:
				bm, ok := f.Matcher.Value().(*xroar.Bitmap)
				if !ok {
					return fmt.Errorf("knoxgpu: set filter carries %T", f.Matcher.Value())
				}
				set := flatten(bm)
				if len(set) > 0 {
					p := C.malloc(C.size_t(8 * len(set)))
					copy(unsafe.Slice((*uint64)(p), len(set)), set)
					cbufs = append(cbufs, p)
					l.set, l.nset = (*C.uint64_t)(p), C.uint32_t(len(set))
				}
			default:
				a, err := pattern(f.Value)
				if err != nil {
					return err
				}
				l.a = C.uint64_t(a)
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
	if root == nil {
		return nil, errors.New("knoxgpu: empty filter tree")
	}
	if err := walk(root); err != nil {
		return nil, err
	}
	if len(leaves) == 0 || len(post) == 0 {
		return nil, errors.New("knoxgpu: filter tree without leaves") // (a match-all query needs no scan)
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

// ScanArgs collects the optional inputs / outputs of ScanEx (kx_scan_ex).
type ScanArgs struct {
	Keys, Versions []uint32
	// Masks[i] (may be nil) holds one bit per row of pack i, bit set = row stays eligible: the complement of the
	// reader's exclusion step — tombstoned rids from the journal and rows whose $xmin / $xmax fail the snapshot
	// (internal/pack/table/reader.go:347-413; engine.TableReader.WithMask, internal/engine/interface.go:96-106).
	Masks [][]byte
	// outputs (all optional)
	Bits      []byte   // per-pack bitsets at Offs[i] (multiples of 8)
	Offs      []uint64 // …
	Counts    []int64
	Sel       []uint32 // selection vectors (not together with Bits); SelOff has len(Keys)+1 entries
	SelOff    []uint64
	AggFields []uint16
	AggTypes  []types.BlockType
	Sharded   bool // combine aggregates and TotalCount over all ranks of the communicator (collective)
}

// ScanEx is Scan with row masks, selection vectors and the cross-rank combine in one call.  Returns the aggregates
// and, for sharded scans, the match count over all ranks.
func (c *Context) ScanEx(prog *Program, a *ScanArgs) ([]AggOut, int64, error) {
	n := len(a.Keys)
	refs := make([]C.kx_packref, n+1)
	for i := 0; i < n; i++ {
		refs[i] = C.kx_packref{pack: C.uint32_t(a.Keys[i]), version: C.uint32_t(a.Versions[i])}
	}
	reqs := make([]C.kx_agg_req, len(a.AggFields)+1)
	outs := make([]C.kx_agg_out, len(a.AggFields)+1)
	for i := range a.AggFields {
		reqs[i] = C.kx_agg_req{field: C.uint16_t(a.AggFields[i]), block_type: C.uint8_t(a.AggTypes[i])}
	}
	var total C.int64_t
	args := C.kx_scan_args{struct_size: C.uint32_t(unsafe.Sizeof(C.kx_scan_args{})), packs: &refs[0], npacks: C.int32_t(n),
		naggs: C.int32_t(len(a.AggFields)), aggs: &reqs[0], agg_out: &outs[0], total_count: &total}
	if a.Sharded {
		args.flags = C.KX_SCAN_SHARDED
	}
	// cgo: a C struct must not carry Go pointers to Go pointers, so the mask table lives in C memory and the masks
	// are pinned for the duration of the call
	var pin runtime.Pinner
	defer pin.Unpin()
	if len(a.Masks) == n && n > 0 {
		tab := (*[1 << 28]*C.uint8_t)(C.malloc(C.size_t(n) * C.size_t(unsafe.Sizeof(uintptr(0)))))[:n:n]
		defer C.free(unsafe.Pointer(&tab[0]))
		for i, m := range a.Masks {
			tab[i] = nil
			if m != nil {
				pin.Pin(unsafe.SliceData(m))
				tab[i] = (*C.uint8_t)(unsafe.SliceData(m))
			}
		}
		args.row_masks = (**C.uint8_t)(unsafe.Pointer(&tab[0]))
	}
	if a.Bits != nil {
		pin.Pin(unsafe.SliceData(a.Bits)); pin.Pin(unsafe.SliceData(a.Offs))
		args.bitsets = (*C.uint8_t)(unsafe.SliceData(a.Bits))
		args.bitset_off = (*C.size_t)(unsafe.Pointer(unsafe.SliceData(a.Offs)))
	}
	if a.Counts != nil {
		pin.Pin(unsafe.SliceData(a.Counts))
		args.counts = (*C.int64_t)(unsafe.Pointer(unsafe.SliceData(a.Counts)))
	}
	if a.SelOff != nil {
		pin.Pin(unsafe.SliceData(a.SelOff))
		args.sel_off = (*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(a.SelOff)))
		if len(a.Sel) > 0 {
			pin.Pin(unsafe.SliceData(a.Sel))
			args.sel, args.sel_cap = (*C.uint32_t)(unsafe.SliceData(a.Sel)), C.size_t(len(a.Sel))
		}
	}
	pin.Pin(&refs[0]); pin.Pin(&reqs[0]); pin.Pin(&outs[0]); pin.Pin(&total)
	if rc := C.kx_scan_ex(c.h, prog.h, &args); rc != 0 {
		return nil, 0, c.err()
	}
	res := make([]AggOut, len(a.AggFields))
	for i := range res {
		o := outs[i]
		res[i] = AggOut{int64(o.count), uint64(o.sum_bits), float64(o.sum_err), uint64(o.min_bits), uint64(o.max_bits), o.valid != 0}
	}
	return res, int64(total), nil
}

// QueryStats mirrors kx_query_stats: the counters the reference reports through query.QueryStats
// (internal/query/stats.go:15-60) for the last scan on this context.
type QueryStats struct {
	RowsScanned, PacksScanned, RowsMatched uint64
	ScanTimeNs, TotalTimeNs                uint64
	KernelLaunches                         uint32
}

func (c *Context) LastQueryStats() QueryStats {
	var q C.kx_query_stats
	C.kx_last_query_stats(c.h, &q)
	return QueryStats{uint64(q.rows_scanned), uint64(q.packs_scanned), uint64(q.rows_matched), uint64(q.scan_time_ns), uint64(q.total_time_ns),
		uint32(q.kernel_launches)}
}

// ---- multi-GPU: one Context per GPU, packs sharded by key, one NCCL all-gather per query inside the library

// CommID creates the communicator id on rank 0 (ncclGetUniqueId); hand the bytes to the other ranks.
func CommID() ([]byte, error) {
	id := make([]byte, C.KX_COMM_ID_BYTES)
	if rc := C.kx_comm_unique_id(unsafe.Pointer(&id[0]), C.size_t(len(id))); rc != 0 {
		return nil, errors.New(C.GoString(C.kx_last_error(nil)))
	}
	return id, nil
}

// CommInit joins the communicator (collective: returns when all ranks called it).  One goroutine per GPU.
func (c *Context) CommInit(nranks, rank int, id []byte) error {
	var p unsafe.Pointer
	if len(id) > 0 {
		p = unsafe.Pointer(&id[0])
	}
	if rc := C.kx_comm_init(c.h, C.int(nranks), C.int(rank), p, C.size_t(len(id))); rc != 0 {
		return c.err()
	}
	return nil
}
