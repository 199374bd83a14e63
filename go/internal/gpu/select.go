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

// ScanSelect is Scan with selection vectors instead of bitsets: what Reader.nextQueryMatch and PhysicalFilter do
// right after filter.Match — sel := bits.Indexes(hits); pack.WithSelection(sel)
// (internal/pack/table/reader.go:432-436, internal/operator/filter.go:29-37).  Pack i owns sel[off[i]:off[i+1]].
func (c *Context) ScanSelect(prog *Program, keys, versions []uint32, capacity int) (sel []uint32, off []uint64, err error) {
	n := len(keys)
	refs := make([]C.kx_packref, n)
	for i := range refs {
		refs[i] = C.kx_packref{pack: C.uint32_t(keys[i]), version: C.uint32_t(versions[i])}
	}
	off = make([]uint64, n+1)
	for {
		sel = make([]uint32, max(capacity, 1))
		rc := C.kx_scan_select(c.h, prog.h, &refs[0], C.int(n), (*C.uint32_t)(unsafe.SliceData(sel)), C.size_t(capacity),
			(*C.uint64_t)(unsafe.SliceData(off)), nil, nil, 0, nil)
		if rc == C.KX_ENOMEM && int(off[n]) > capacity {
			capacity = int(off[n]) // the library reports the required size
			continue
		}
		if rc != 0 {
			return nil, nil, c.err()
		}
		return sel[:off[n]], off, nil
	}
}

// Gather is NumberContainer.AppendTo(dst, sel) for a batch of packs (what Result.Append copies for the selected
// rows, internal/query/result.go:196-264): dst receives len(sel) elements of the column's type.
func (c *Context) Gather(keys, versions []uint32, field uint16, t types.BlockType, sel []uint32, off []uint64, dst unsafe.Pointer) error {
	refs := make([]C.kx_packref, len(keys))
	for i := range refs {
		refs[i] = C.kx_packref{pack: C.uint32_t(keys[i]), version: C.uint32_t(versions[i])}
	}
	rc := C.kx_gather(c.h, &refs[0], C.int(len(refs)), C.uint16_t(field), C.uint8_t(t), (*C.uint32_t)(unsafe.SliceData(sel)),
		(*C.uint64_t)(unsafe.SliceData(off)), dst)
	if rc != 0 {
		return c.err()
	}
	return nil
}

// GatherBytes is StringContainer.AppendTo(dst, sel) for a batch of packs: the byte-string twin of Gather
// (internal/encode/string_{const,fixed,compact,dict}.go AppendTo with a selection).  Row i of the selection is
// buf[offs[i]:offs[i+1]]; the first call asks the library for the size (KX_ENOMEM convention of kx_scan_select).
func (c *Context) GatherBytes(keys, versions []uint32, field uint16, sel []uint32, off []uint64) (buf []byte, offs []uint64, err error) {
	refs := make([]C.kx_packref, len(keys))
	for i := range refs {
		refs[i] = C.kx_packref{pack: C.uint32_t(keys[i]), version: C.uint32_t(versions[i])}
	}
	offs = make([]uint64, len(sel)+1)
	if len(sel) == 0 || len(refs) == 0 {
		return nil, offs, nil
	}
	capacity := 0
	for {
		buf = make([]byte, max(capacity, 1))
		rc := C.kx_gather_bytes(c.h, &refs[0], C.int(len(refs)), C.uint16_t(field), (*C.uint32_t)(unsafe.SliceData(sel)),
			(*C.uint64_t)(unsafe.SliceData(off)), (*C.uint64_t)(unsafe.SliceData(offs)), (*C.uint8_t)(unsafe.SliceData(buf)), C.size_t(capacity))
		if rc == C.KX_ENOMEM && int(offs[len(sel)]) > capacity {
			capacity = int(offs[len(sel)])
			continue
		}
		if rc != 0 {
			return nil, nil, c.err()
		}
		return buf[:offs[len(sel)]], offs, nil
	}
}
