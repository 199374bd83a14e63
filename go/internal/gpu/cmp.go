//go:build knoxgpu

package gpu

/*
#include "knoxgpu.h"
*/
import "C"

import (
	"unsafe"

	"blockwatch.cc/knoxdb/internal/cmp"
	"blockwatch.cc/knoxdb/internal/types"
)

// Cmp is kx_cmp: cmp.<Type><Op>(src []T, val T, bits []byte) int64 and <Type>Between(src, a, b, bits)
// (internal/cmp/cmp.go:6-114): writes ceil(n/8) LSB-first bytes into bits, returns the popcount.
func (c *Context) Cmp(t types.BlockType, mode types.FilterMode, src unsafe.Pointer, n int, a, b uint64, bits []byte) int64 {
	return int64(C.kx_cmp(c.h, C.uint8_t(t), C.uint8_t(mode), src, C.size_t(n), C.uint64_t(a), C.uint64_t(b),
		(*C.uint8_t)(unsafe.SliceData(bits))))
}

// InstallCmpKernels repoints the assignable compare kernels exactly the way the AVX2 / AVX-512 builds do at
// init() (internal/cmp/cmp_amd64.go:14-218).  Shown for uint64 and int64; the other element types follow the
// same two lines per operator.
func InstallCmpKernels(c *Context) {
	u64 := func(mode types.FilterMode) func([]uint64, uint64, []byte) int64 {
		return func(src []uint64, val uint64, bits []byte) int64 {
			return c.Cmp(types.BlockUint64, mode, unsafe.Pointer(unsafe.SliceData(src)), len(src), val, 0, bits)
		}
	}
	cmp.Uint64Equal, cmp.Uint64NotEqual = u64(types.FilterModeEqual), u64(types.FilterModeNotEqual)
	cmp.Uint64Less, cmp.Uint64LessEqual = u64(types.FilterModeLt), u64(types.FilterModeLe)
	cmp.Uint64Greater, cmp.Uint64GreaterEqual = u64(types.FilterModeGt), u64(types.FilterModeGe)
	cmp.Uint64Between = func(src []uint64, a, b uint64, bits []byte) int64 {
		return c.Cmp(types.BlockUint64, types.FilterModeRange, unsafe.Pointer(unsafe.SliceData(src)), len(src), a, b, bits)
	}
	i64 := func(mode types.FilterMode) func([]int64, int64, []byte) int64 {
		return func(src []int64, val int64, bits []byte) int64 {
			return c.Cmp(types.BlockInt64, mode, unsafe.Pointer(unsafe.SliceData(src)), len(src), uint64(val), 0, bits)
		}
	}
	cmp.Int64Equal, cmp.Int64NotEqual = i64(types.FilterModeEqual), i64(types.FilterModeNotEqual)
	cmp.Int64Less, cmp.Int64LessEqual = i64(types.FilterModeLt), i64(types.FilterModeLe)
	cmp.Int64Greater, cmp.Int64GreaterEqual = i64(types.FilterModeGt), i64(types.FilterModeGe)
	cmp.Int64Between = func(src []int64, a, b int64, bits []byte) int64 {
		return c.Cmp(types.BlockInt64, types.FilterModeRange, unsafe.Pointer(unsafe.SliceData(src)), len(src), uint64(a), uint64(b), bits)
	}
}
