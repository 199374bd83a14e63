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

// Stats is a statistics index resident on the device: per data pack and column the zone map (the min/max columns of a
// statistics pack, internal/pack/stats/index.go) and optionally a bloom filter (the buffers stored under
// encodeFilterKey, internal/pack/stats/filter.go:26-32).  It replaces the per-statistics-pack loop of
// stats.matchVector / matchFilterVector (internal/pack/stats/match.go:92-195) with one kernel launch over all packs
// and removes the per-candidate KV read of the filter.
type Stats struct {
	h   *C.kx_stats
	ctx *Context
}

// NewStats uploads zone maps; mins/maxs are column-major [field][pack] operand patterns (see pattern()).
func (c *Context) NewStats(npacks int, fields []uint16, typs []types.BlockType, mins, maxs []uint64) (*Stats, error) {
	ft := make([]C.uint8_t, len(typs))
	for i, t := range typs {
		ft[i] = C.uint8_t(t)
	}
	var h *C.kx_stats
	rc := C.kx_stats_create(c.h, C.int(npacks), (*C.uint16_t)(unsafe.SliceData(fields)), &ft[0], C.int(len(fields)),
		(*C.uint64_t)(unsafe.SliceData(mins)), (*C.uint64_t)(unsafe.SliceData(maxs)), &h)
	if rc != 0 {
		return nil, c.err()
	}
	return &Stats{h, c}, nil
}

func (s *Stats) Close() { C.kx_stats_free(s.h); s.h = nil }

// PutBloom attaches a stored filter ([k][m/8 bytes], bloom.Filter.Bytes()).
func (s *Stats) PutBloom(field, pack int, buf []byte) error {
	if rc := C.kx_stats_put_bloom(s.h, C.int(field), C.int(pack), unsafe.Pointer(unsafe.SliceData(buf)), C.size_t(len(buf))); rc != 0 {
		return s.ctx.err()
	}
	return nil
}

// BuildBloom is stats.BuildBloomFilter (internal/pack/stats/filter.go:296-367) on the device: values is the
// materialised column slice of the pack (little-endian, as block.Block holds it); the resulting buffer is
// bit-identical to bloom.NewFilter(cardinality*factor*8) + Add(hash.Vec64(values)...).
func (s *Stats) BuildBloom(field, pack int, t types.BlockType, values unsafe.Pointer, n, cardinality, factor int) error {
	rc := C.kx_stats_build_bloom(s.h, C.int(field), C.int(pack), C.uint8_t(t), values, nil, C.size_t(n), C.int(cardinality), C.int(factor))
	if rc != 0 {
		return s.ctx.err()
	}
	return nil
}

// Bloom reads a filter back as bloom.NewFilterBuffer expects it, to persist it the way the reference does.
func (s *Stats) Bloom(field, pack int) ([]byte, error) {
	var n C.size_t
	if rc := C.kx_stats_get_bloom(s.h, C.int(field), C.int(pack), nil, 0, &n); rc != 0 {
		return nil, s.ctx.err()
	}
	if n == 0 {
		return nil, nil
	}
	buf := make([]byte, int(n))
	if rc := C.kx_stats_get_bloom(s.h, C.int(field), C.int(pack), unsafe.Pointer(&buf[0]), n, &n); rc != 0 {
		return nil, s.ctx.err()
	}
	return buf, nil
}

// Query evaluates prog over the resident index; out has ceil(npacks/8) bytes, a set bit keeps the pack.
// hashes/hashOff carry the probe hashes of byte-string EQ/IN leaves (hash.Hash); nil lets the library hash the
// numeric operands itself (hash.HashT).
func (s *Stats) Query(prog *Program, hashes []uint64, hashOff []uint32, out []byte) (int64, error) {
	var hp *C.uint64_t
	var ho *C.uint32_t
	if hashOff != nil {
		if len(hashes) > 0 {
			hp = (*C.uint64_t)(unsafe.SliceData(hashes))
		} else {
			var zero C.uint64_t
			hp = &zero
		}
		ho = (*C.uint32_t)(unsafe.SliceData(hashOff))
	}
	n := C.kx_prune_stats(s.ctx.h, prog.h, s.h, hp, ho, (*C.uint8_t)(unsafe.SliceData(out)))
	if n < 0 {
		return 0, s.ctx.err()
	}
	return int64(n), nil
}
