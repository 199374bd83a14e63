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
        LaneAcc acc[MAX_AGGS];
#pragma unroll
        for (int j = 0; j < MAX_AGGS; ++j) lane_reset(acc[j]);

        auto flush = [&]() {
            if (cur == 0xffffffffu || cnt == 0) return;
            BucketCell* cell = P.table + (size_t)cur * ncell;
            atomicAdd(reinterpret_cast<unsigned long long*>(&cell[0].count), (unsigned long long)cnt);
#pragma unroll
            for (int j = 0; j < MAX_AGGS; ++j) {
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

        for (uint32_t g = g_begin; g < g_end; ++g) {
            const uint32_t word = __ldg(words + g);
            if (word == 0) continue;                       // warp-uniform: groups without a match cost one load
            if (!((word >> lane) & 1u)) continue;
            const uint32_t row = g * 32u + lane;
            const uint64_t key = decode_value_at(tsv, row) ^ P.ts_flip;
            if (key < cur_lo || key >= cur_hi) {
                flush();
                // window k with edge[k] <= key < edge[k + 1]; rows outside [edge[0], edge[nbuckets]) belong to no window
                uint32_t a = 0, b = P.nbuckets + 1u;       // first edge > key
                while (a < b) { uint32_t m = (a + b) >> 1; if (__ldg(P.edges + m) <= key) a = m + 1; else b = m; }
                if (a == 0 || a > P.nbuckets) { cur = 0xffffffffu; cur_lo = 1; cur_hi = 0; continue; }
                cur = a - 1u; cur_lo = __ldg(P.edges + cur); cur_hi = __ldg(P.edges + cur + 1);
            }
            ++cnt;
#pragma unroll
            for (int j = 0; j < MAX_AGGS; ++j) {
                if ((uint32_t)j >= P.naggs) break;
                const uint64_t bits = decode_value_at(P.views[(size_t)pack * ncell + 1 + j], row);
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
        }
        flush();
    }
}

cudaError_t launch_bucket(const BucketParams& P, int num_sms, cudaStream_t stream) {
    if (P.njobs == 0) return cudaSuccess;
    const uint32_t warps_per_block = BUCKET_THREADS / 32;
    uint32_t grid = (P.njobs + warps_per_block - 1) / warps_per_block;
    const uint32_t cap = (uint32_t)num_sms * 8u;
    if (grid > cap) grid = cap;
    bucket_kernel<<<grid, BUCKET_THREADS, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
