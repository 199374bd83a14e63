//go:build knoxgpu

package gpu

/*
#include "knoxgpu.h"
*/
import "C"

import (
	"unsafe"

	"blockwatch.cc/knoxdb/internal/bitset"
	"blockwatch.cc/knoxdb/internal/types"
	"blockwatch.cc/knoxdb/internal/xroar"
)

// Matcher implements types.NumberMatcher[T] (internal/types/number.go:37-47) for one encoded block, the
// interface block.GetMatcher[T](b) hands to the leaf matchers (internal/block/access.go:42-44,
// internal/operator/filter/match_num.go:327-817).  `bits` is caller-allocated, block-length and pre-zeroed;
// the callee sets bits and calls bits.ResetCount(n), exactly the contract of RawContainer.MatchEqual
// (internal/encode/int_raw.go:122-149).  The `mask` argument is a pure optimisation in the reference
// (every numeric kernel ignores it, internal/cmp/matcher.go:80-113) and is ignored here.
//
// This per-block seam uploads the block on every call; it exists for parity and for journal segments.
// The fast path is the batched Scan (knoxgpu.go) behind the operator in operator.go.
type Matcher[T types.Number] struct {
	Ctx *Context
	Typ types.BlockType
	Enc []byte // container bytes: [type id][uvarint header…][payload]
}

var _ types.NumberMatcher[int64] = Matcher[int64]{}

func (m Matcher[T]) match(mode types.FilterMode, a, b uint64, set []uint64, bits *bitset.Bitset) {
	var sp *C.uint64_t
	if len(set) > 0 {
		sp = (*C.uint64_t)(unsafe.SliceData(set))
	}
	n := C.kx_container_match(m.Ctx.h, C.uint8_t(m.Typ), unsafe.Pointer(unsafe.SliceData(m.Enc)), C.size_t(len(m.Enc)),
		C.uint8_t(mode), C.uint64_t(a), C.uint64_t(b), sp, C.uint32_t(len(set)), (*C.uint8_t)(unsafe.SliceData(bits.Bytes())))
	if n < 0 {
		panic(m.Ctx.err()) // the interface has no error return; the reference panics on corrupt containers too
	}
	bits.ResetCount(int(n))
}

func (m Matcher[T]) MatchEqual(v T, bits, _ *bitset.Bitset)        { m.match(types.FilterModeEqual, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchNotEqual(v T, bits, _ *bitset.Bitset)     { m.match(types.FilterModeNotEqual, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchLess(v T, bits, _ *bitset.Bitset)         { m.match(types.FilterModeLt, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchLessEqual(v T, bits, _ *bitset.Bitset)    { m.match(types.FilterModeLe, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchGreater(v T, bits, _ *bitset.Bitset)      { m.match(types.FilterModeGt, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchGreaterEqual(v T, bits, _ *bitset.Bitset) { m.match(types.FilterModeGe, pattern(v), 0, nil, bits) }
func (m Matcher[T]) MatchBetween(a, b T, bits, _ *bitset.Bitset) {
	m.match(types.FilterModeRange, pattern(a), pattern(b), nil, bits)
}
func (m Matcher[T]) MatchInSet(s any, bits, _ *bitset.Bitset) {
	m.match(types.FilterModeIn, 0, 0, flatten(s.(*xroar.Bitmap)), bits)
}
func (m Matcher[T]) MatchNotInSet(s any, bits, _ *bitset.Bitset) {
	m.match(types.FilterModeNotIn, 0, 0, flatten(s.(*xroar.Bitmap)), bits)
}
