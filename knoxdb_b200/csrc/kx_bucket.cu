// kx_bucket.cu — time-bucketed reduce (group by time window) over the match bitsets of a scan.
//
// Replaces the row loop of the reference's series query (pkg/series/series.go:192-256: for every streamed row
// t = Interval.TruncateRelative(ts, Range.From), then Bucket.Push → the window's Reducer.Reduce) and the window
// bookkeeping of reducer.NativeBucket[T].Push (internal/reducer/bucket_native.go:104-167) for the order-independent
// reducers count / sum / min / max (reducer.go:138-297).  Window edges are computed by the caller (TimeUnit.Next walks
// the calendar, pkg/util/timeunit.go:234-263); window k is [edge[k], edge[k+1]).
//
// One warp owns a run of consecutive 32-row groups of a pack; lane l takes row 32 g + l of every group that has
// matches (the bitset word is read once per group).  A lane decodes the timestamp of its matching row at index,
// finds the window (the previous window is tried first: time-ordered packs stay in one window for thousands of rows)
// and folds the row's values into lane-local accumulators, which are flushed to the window table in HBM with atomics
// only when the lane's window changes.  Integer sums wrap mod 2^64 like the reference's `r.v += v`; min / max use an
// order-preserving unsigned key, so every integer result is exact and independent of the flush order.  float64 sums
// are Neumaier-compensated inside a lane and added with atomicAdd(double): the order of those few adds is not fixed.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"
#include "kx_kernels.h"
#include "kx_decode.cuh"

namespace kx {

__device__ __forceinline__ uint64_t decode_value_at(const ColView& v, uint32_t row) { return decode_value(v, row, nullptr, 0); }

__device__ __forceinline__ uint64_t f64_key(uint64_t b) { return (b >> 63) ? ~b : (b | 0x8000000000000000ull); }

struct LaneAcc { uint64_t sum, err, mn, mx; };

__device__ __forceinline__ void lane_reset(LaneAcc& a) { a.sum = 0; a.err = 0; a.mn = ~0ull; a.mx = 0; }

// NA = value columns the instantiation carries (1, 2 or 4): fewer columns → fewer registers → more resident warps
template <int NA>
__global__ void __launch_bounds__(BUCKET_THREADS) bucket_kernel(const BucketParams P) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t ncell = P.naggs + 1u;
    for (uint32_t job = warp; job < P.njobs; job += nwarps) {
        uint32_t lo = 0, hi = P.npacks;   // last pack with job0 <= job
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (P.job0[m] <= job) lo = m; else hi = m; }
        const uint32_t pack = lo;
        const PackInfo pi = P.packs[pack];
        const ColView tsv = P.views[(size_t)pack * ncell];
        const uint32_t g_begin = (job - P.job0[pack]) * BUCKET_JOB_GROUPS, ngroups = (pi.n + 31u) >> 5;
        const uint32_t g_end = min(g_begin + BUCKET_JOB_GROUPS, ngroups);
        const uint32_t* words = reinterpret_cast<const uint32_t*>(P.bits + pi.bitset_off);

        uint32_t cur = 0xffffffffu, cnt = 0;   // window of the rows held in the lane accumulators
        uint64_t cur_lo = 1, cur_hi = 0;       // its key range [cur_lo, cur_hi)
        LaneAcc acc[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) lane_reset(acc[j]);

        auto flush = [&]() {
            if (cur == 0xffffffffu || cnt == 0) return;
            BucketCell* cell = P.table + (size_t)cur * ncell;
            atomicAdd(reinterpret_cast<unsigned long long*>(&cell[0].count), (unsigned long long)cnt);
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                if ((uint32_t)j >= P.naggs) break;
                BucketCell* c = cell + 1 + j;
                atomicAdd(reinterpret_cast<unsigned long long*>(&c->count), (unsigned long long)cnt);
                if (P.agg_type[j] == 9) {
                    const double s = __longlong_as_double((long long)acc[j].sum) + __longlong_as_double((long long)acc[j].err);
                    atomicAdd(reinterpret_cast<double*>(&c->sum), s);
                } else {
                    atomicAdd(reinterpret_cast<unsigned long long*>(&c->sum), (unsigned long long)acc[j].sum);
                }
                atomicMin(reinterpret_cast<unsigned long long*>(&c->mn), (unsigned long long)acc[j].mn);
                atomicMax(reinterpret_cast<unsigned long long*>(&c->mx), (unsigned long long)acc[j].mx);
                lane_reset(acc[j]);
            }
            cnt = 0;
        };

        // one matching row: find its window (the lane's current window first), fold the values in
        auto consume = [&](uint64_t key, const uint64_t (&v)[NA]) {
            if (key < cur_lo || key >= cur_hi) {
                flush();
                // time-ordered packs walk the windows in order: try the NEXT window before searching (one load instead of
                // log2(nbuckets) dependent ones)
                if (cur != 0xffffffffu && cur + 1u < P.nbuckets && key >= cur_hi) {
                    const uint64_t nh = __ldg(P.edges + cur + 2);
                    if (key < nh) { ++cur; cur_lo = cur_hi; cur_hi = nh; goto found; }
                }
                {
                    // window k with edge[k] <= key < edge[k + 1]; rows outside [edge[0], edge[nbuckets]) belong to no window
                    uint32_t a = 0, b = P.nbuckets + 1u;   // first edge > key
                    while (a < b) { uint32_t m = (a + b) >> 1; if (__ldg(P.edges + m) <= key) a = m + 1; else b = m; }
                    if (a == 0 || a > P.nbuckets) { cur = 0xffffffffu; cur_lo = 1; cur_hi = 0; return; }
                    cur = a - 1u; cur_lo = __ldg(P.edges + cur); cur_hi = __ldg(P.edges + cur + 1);
                }
            found:;
            }
            ++cnt;
#pragma unroll
            for (int j = 0; j < NA; ++j) {
                if ((uint32_t)j >= P.naggs) break;
                const uint64_t bits = v[j];
                LaneAcc& A = acc[j];
                if (P.agg_type[j] == 9) {
                    const double x = __longlong_as_double((long long)bits), sum = __longlong_as_double((long long)A.sum);
                    double err = __longlong_as_double((long long)A.err);
                    const double t = sum + x;
                    err += (fabs(sum) >= fabs(x)) ? ((sum - t) + x) : ((x - t) + sum);
                    A.sum = (uint64_t)__double_as_longlong(t); A.err = (uint64_t)__double_as_longlong(err);
                    const uint64_t k = f64_key(bits);
                    if (k < A.mn) A.mn = k;
                    if (k > A.mx) A.mx = k;
                } else {
                    A.sum += bits;
                    const uint64_t k = bits ^ (type_is_signed(P.agg_type[j]) ? 0x8000000000000000ull : 0ull);
                    if (k < A.mn) A.mn = k;
                    if (k > A.mx) A.mx = k;
                }
            }
        };

        // Fast path (the common series shape): bit-packed / raw timestamp column, raw 64-bit value columns.  U groups at a
        // time: the raw words of every matching row (timestamp words + values) are requested back to back and only
        // then decoded and consumed in row order, so a warp pays one memory round trip per U groups, not three per group.
        bool fast = tsv.kind == CK_BITS && tsv.width != 0;
        const unsigned long long* vptr[NA];
        uint64_t vbase[NA];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            vptr[j] = nullptr; vbase[j] = 0;
            if ((uint32_t)j < P.naggs) {
                const ColView& av = P.views[(size_t)pack * ncell + 1 + j];
                if (av.kind == CK_BITS && av.width == 64) { vptr[j] = reinterpret_cast<const unsigned long long*>(av.data); vbase[j] = type_is_float(av.type) ? 0ull : av.base; }
                else fast = false;
            }
        }
        constexpr int U = 4;
        // Short windows (a job spans many of them): with lane = row mod 32 every lane would change its window every few
        // rows and flush each time.  The job's first and last timestamp tell how many windows it spans; when a window is
        // shorter than ~256 rows each lane takes a run of CONSECUTIVE rows instead (rows_per_lane of them: few window
        // changes per lane; the strided loads stay in L1 because a lane walks its cache lines to the end).
        bool consecutive = false;
        if (fast) {
            const uint32_t r_first = g_begin * 32u, r_last = min(g_end * 32u, pi.n) - 1u;
            const uint64_t k0 = decode_value_at(tsv, r_first) ^ P.ts_flip, k1 = decode_value_at(tsv, r_last) ^ P.ts_flip;
            uint32_t a0 = 0, b0 = P.nbuckets + 1u, a1 = 0, b1 = P.nbuckets + 1u;
            while (a0 < b0) { uint32_t m = (a0 + b0) >> 1; if (__ldg(P.edges + m) <= k0) a0 = m + 1; else b0 = m; }
            while (a1 < b1) { uint32_t m = (a1 + b1) >> 1; if (__ldg(P.edges + m) <= k1) a1 = m + 1; else b1 = m; }
            const uint32_t span = a1 > a0 ? a1 - a0 + 1u : a0 - a1 + 1u;
            consecutive = (uint64_t)span * 256u > (uint64_t)(r_last - r_first + 1u);
        }
        if (fast && consecutive) {
            const uint32_t tw = tsv.width;
            const uint32_t* tsw = reinterpret_cast<const uint32_t*>(tsv.data);
            const uint64_t tmask = width_mask((int)tw);
            const uint32_t r_begin = g_begin * 32u, r_end = min(g_end * 32u, pi.n);
            const uint32_t per_lane = (r_end - r_begin + 31u) / 32u;
            const uint32_t my0 = r_begin + lane * per_lane, my1 = min(my0 + per_lane, r_end);
            for (uint32_t r0 = my0; r0 < my1; r0 += U) {
                bool on[U];
                uint32_t t0[U], t1[U], t2[U];
                uint64_t val[U][NA];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t row = r0 + u;
                    on[u] = row < my1 && ((__ldg(words + (row >> 5)) >> (row & 31u)) & 1u);
                    const uint64_t bit = (uint64_t)row * tw;
                    const uint32_t* wp = tsw + (bit >> 5);
                    t0[u] = on[u] ? __ldg(wp) : 0u;
                    t1[u] = on[u] ? __ldg(wp + 1) : 0u;
                    t2[u] = (on[u] && tw > 32u) ? __ldg(wp + 2) : 0u;
#pragma unroll
                    for (int j = 0; j < NA; ++j) val[u][j] = (on[u] && vptr[j]) ? __ldg(vptr[j] + row) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!on[u]) continue;
                    const uint32_t sh = (uint32_t)(((uint64_t)(r0 + u) * tw) & 31u);
                    const uint64_t f = (((uint64_t)__funnelshift_r(t1[u], t2[u], sh) << 32) | __funnelshift_r(t0[u], t1[u], sh)) & tmask;
                    uint64_t v[NA];
#pragma unroll
                    for (int j = 0; j < NA; ++j) v[j] = val[u][j] + vbase[j];
                    consume(type_ext(tsv.type, f + tsv.base) ^ P.ts_flip, v);
                }
            }
        } else if (fast) {
            const uint32_t tw = tsv.width;
            const uint32_t* tsw = reinterpret_cast<const uint32_t*>(tsv.data);
            const uint64_t tmask = width_mask((int)tw);
            for (uint32_t g0 = g_begin; g0 < g_end; g0 += U) {
                uint32_t wd[U], anyw = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    wd[u] = g0 + u < g_end ? __ldg(words + g0 + u) : 0u;
                    anyw |= wd[u];
                }
                if (anyw == 0) continue;                   // warp-uniform: groups without a match cost one load each
                uint32_t t0[U], t1[U], t2[U];
                uint64_t val[U][NA];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool on = (wd[u] >> lane) & 1u;
                    const uint32_t row = (g0 + u) * 32u + lane;
                    const uint64_t bit = (uint64_t)row * tw;
                    const uint32_t* wp = tsw + (bit >> 5);
                    t0[u] = on ? __ldg(wp) : 0u;
                    t1[u] = on ? __ldg(wp + 1) : 0u;
                    t2[u] = (on && tw > 32u) ? __ldg(wp + 2) : 0u;
#pragma unroll
                    for (int j = 0; j < NA; ++j) val[u][j] = (on && vptr[j]) ? __ldg(vptr[j] + row) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (!((wd[u] >> lane) & 1u)) continue;
                    const uint32_t sh = (uint32_t)(((uint64_t)((g0 + u) * 32u + lane) * tw) & 31u);
                    const uint64_t f = (((uint64_t)__funnelshift_r(t1[u], t2[u], sh) << 32) | __funnelshift_r(t0[u], t1[u], sh)) & tmask;
                    uint64_t v[NA];
#pragma unroll
                    for (int j = 0; j < NA; ++j) v[j] = val[u][j] + vbase[j];
                    consume(type_ext(tsv.type, f + tsv.base) ^ P.ts_flip, v);
                }
            }
        } else {
            // generic path: any container kind, decoded at index row by row
            for (uint32_t g = g_begin; g < g_end; ++g) {
                const uint32_t word = __ldg(words + g);
                if (!((word >> lane) & 1u)) continue;
                const uint32_t row = g * 32u + lane;
                uint64_t v[NA];
#pragma unroll
                for (int j = 0; j < NA; ++j) v[j] = (uint32_t)j < P.naggs ? decode_value_at(P.views[(size_t)pack * ncell + 1 + j], row) : 0ull;
                consume(decode_value_at(tsv, row) ^ P.ts_flip, v);
            }
        }
        // end of the job: when every lane that holds rows holds them for the SAME window (time-ordered packs, windows
        // longer than a job), the warp combines its 32 accumulators with shuffles and flushes once instead of 32 times
        {
            const uint32_t holders = __ballot_sync(0xffffffffu, cur != 0xffffffffu && cnt != 0);
            if (holders) {
                const uint32_t c0 = __shfl_sync(0xffffffffu, cur, __ffs((int)holders) - 1);
                if (__all_sync(0xffffffffu, cnt == 0 || cur == 0xffffffffu || cur == c0)) {
                    if (cur != c0) cnt = 0;   // lanes without rows (their accumulators are at the identity)
                    for (int off = 16; off > 0; off >>= 1) {
                        cnt += __shfl_down_sync(0xffffffffu, cnt, off);
#pragma unroll
                        for (int j = 0; j < NA; ++j) {
                            if ((uint32_t)j >= P.naggs) break;
                            const uint64_t s2 = __shfl_down_sync(0xffffffffu, acc[j].sum, off), e2 = __shfl_down_sync(0xffffffffu, acc[j].err, off);
                            const uint64_t mn2 = __shfl_down_sync(0xffffffffu, acc[j].mn, off), mx2 = __shfl_down_sync(0xffffffffu, acc[j].mx, off);
                            if (P.agg_type[j] == 9) {   // double-double style merge of two compensated sums
                                const double a = __longlong_as_double((long long)acc[j].sum), b = __longlong_as_double((long long)s2);
                                const double t = a + b, c = (fabs(a) >= fabs(b)) ? ((a - t) + b) : ((b - t) + a);
                                acc[j].sum = (uint64_t)__double_as_longlong(t);
                                acc[j].err = (uint64_t)__double_as_longlong(__longlong_as_double((long long)acc[j].err) + __longlong_as_double((long long)e2) + c);
                            } else {
                                acc[j].sum += s2;
                            }
                            if (mn2 < acc[j].mn) acc[j].mn = mn2;
                            if (mx2 > acc[j].mx) acc[j].mx = mx2;
                        }
                    }
                    cur = c0;
                    if (lane != 0) cnt = 0;
                }
            }
            flush();
        }
    }
}

cudaError_t launch_bucket(const BucketParams& P, int num_sms, cudaStream_t stream) {
    if (P.njobs == 0) return cudaSuccess;
    const uint32_t warps_per_block = BUCKET_THREADS / 32;
    uint32_t grid = (P.njobs + warps_per_block - 1) / warps_per_block;
    const uint32_t cap = (uint32_t)num_sms * 8u;
    if (grid > cap) grid = cap;
    if (P.naggs <= 1) bucket_kernel<1><<<grid, BUCKET_THREADS, 0, stream>>>(P);
    else if (P.naggs <= 2) bucket_kernel<2><<<grid, BUCKET_THREADS, 0, stream>>>(P);
    else bucket_kernel<4><<<grid, BUCKET_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
