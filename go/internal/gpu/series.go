//go:build knoxgpu

package gpu

/*
#include "knoxgpu.h"
*/
import "C"

import (
	"time"
	"unsafe"

	"blockwatch.cc/knoxdb/internal/types"
	"blockwatch.cc/knoxdb/pkg/util"
)

// WindowEdges lists the window starts a series query walks: the first window starts at r.From, every following one
// at unit.Next of its predecessor (what TimeUnit.TruncateRelative steps through for every streamed row,
// pkg/util/timeunit.go:234-263), up to the first start at or after r.To.  Calendar units (week, month, quarter, year)
// give irregular windows, which is why the device takes explicit edges instead of an interval.
func WindowEdges(r util.TimeRange, unit util.TimeUnit) []int64 {
	edges := []int64{r.From.UnixNano()}
	for t := r.From; t.Before(r.To); {
		t = unit.Next(t, 1)
		edges = append(edges, t.UnixNano())
	}
	return edges
}

// ScanBuckets is the series query's row loop on the device (pkg/series/series.go:192-256): evaluate prog over the
// packs, map every matching row to its time window and reduce the value columns per window with the
// order-independent reducers count / sum / min / max (internal/reducer/reducer.go:138-297).  res[j][k] is value
// column j in window k = [edges[k], edges[k+1]); counts[k] is the window's match count.  The caller turns the cells
// into reducer.Bucket output (mean = sum / count; fill modes stay on the host, internal/reducer/fill.go).
func (c *Context) ScanBuckets(prog *Program, keys, versions []uint32, tsField uint16, edges []int64,
	aggFields []uint16, aggTypes []types.BlockType) (res [][]AggOut, counts []int64, err error) {
	n, nb := len(keys), len(edges)-1
	if n == 0 || nb < 1 {
		return nil, nil, nil
	}
	refs := make([]C.kx_packref, n)
	for i := range refs {
		refs[i] = C.kx_packref{pack: C.uint32_t(keys[i]), version: C.uint32_t(versions[i])}
	}
	reqs := make([]C.kx_agg_req, max(len(aggFields), 1))
	for i := range aggFields {
		reqs[i] = C.kx_agg_req{field: C.uint16_t(aggFields[i]), block_type: C.uint8_t(aggTypes[i])}
	}
	outs := make([]C.kx_agg_out, max(len(aggFields)*nb, 1))
	counts = make([]int64, nb)
	rc := C.kx_scan_buckets(c.h, prog.h, &refs[0], C.int(n), C.uint16_t(tsField), C.uint8_t(types.BlockInt64),
		(*C.uint64_t)(unsafe.Pointer(unsafe.SliceData(edges))), C.int(nb), &reqs[0], C.int(len(aggFields)),
		(*C.int64_t)(unsafe.Pointer(unsafe.SliceData(counts))), &outs[0], nil)
	if rc != 0 {
		return nil, nil, c.err()
	}
	res = make([][]AggOut, len(aggFields))
	for j := range res {
		res[j] = make([]AggOut, nb)
		for k := 0; k < nb; k++ {
			o := outs[j*nb+k]
			res[j][k] = AggOut{int64(o.count), uint64(o.sum_bits), float64(o.sum_err), uint64(o.min_bits), uint64(o.max_bits), o.valid != 0}
		}
	}
	return res, counts, nil
}

var _ = time.Second
