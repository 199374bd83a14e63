//go:build knoxgpu

package gpu

/*
#include "knoxgpu.h"
*/
import "C"

import (
	"unsafe"

	"blockwatch.cc/knoxdb/internal/types"
)

// HashValue / HashBytes are hash.HashT / hash.Hash (XXH3-64, seed 0; internal/hash/hash.go:26,67-92):
// the probe hashes of EQ / IN operands, computed once per query.
func HashValue(t types.BlockType, v any) uint64 {
	return uint64(C.kx_hash_value(C.uint8_t(t), C.uint64_t(pattern(v))))
}
func HashBytes(b []byte) uint64 {
	return uint64(C.kx_hash_bytes(unsafe.Pointer(unsafe.SliceData(b)), C.size_t(len(b))))
}

// Prune evaluates prog over per-pack zone maps and bloom filters: the vectorised part of
// stats.matchVector → Matcher.MatchRangeVectors + bloom.Filter.Contains per candidate
// (internal/pack/stats/match.go:92-195).  mins/maxs are npacks×nleaves operand patterns of the leaf
// columns' statistics (the min/max columns of a statistics pack), blooms the buffers stored under
// encodeFilterKey (nil where a pack has none).  out has ceil(npacks/8) bytes; a set bit keeps the pack.
func (c *Context) Prune(prog *Program, npacks int, mins, maxs []uint64, blooms [][]byte, hashes []uint64, hashOff []uint32, out []byte) (int64, error) {
	var (
		bp *unsafe.Pointer
		bl *C.size_t
		hp *C.uint64_t
		ho *C.uint32_t
	)
	var pin []unsafe.Pointer
	var lens []C.size_t
	if blooms != nil {
		// C arrays of C copies: the pointer table itself must not contain Go pointers
		pin = make([]unsafe.Pointer, len(blooms))
		lens = make([]C.size_t, len(blooms))
		tab := (*[1 << 28]unsafe.Pointer)(C.malloc(C.size_t(len(blooms)) * C.size_t(unsafe.Sizeof(uintptr(0)))))
		defer C.free(unsafe.Pointer(tab))
		for i, b := range blooms {
			if len(b) > 0 {
				pin[i] = C.CBytes(b)
				lens[i] = C.size_t(len(b))
			}
			tab[i] = pin[i]
		}
		defer func() {
			for _, p := range pin {
				if p != nil {
					C.free(p)
				}
			}
		}()
		bp, bl = (*unsafe.Pointer)(unsafe.Pointer(tab)), &lens[0]
		hp, ho = (*C.uint64_t)(unsafe.SliceData(hashes)), (*C.uint32_t)(unsafe.SliceData(hashOff))
	}
	n := C.kx_prune(c.h, prog.h, C.int(npacks), (*C.uint64_t)(unsafe.SliceData(mins)), (*C.uint64_t)(unsafe.SliceData(maxs)),
		bp, bl, hp, ho, (*C.uint8_t)(unsafe.SliceData(out)))
	if n < 0 {
		return 0, c.err()
	}
	return int64(n), nil
}
