// kx_scan.cu — the fused decode + filter + reduce kernel of libknoxgpu (sm_100a).
//
// One persistent, warp-specialised kernel scans a whole batch of packs:
//   * a producer warp streams each tile's packed column bytes HBM → shared memory with
//     TMA bulk copies (cp.async.bulk … mbarrier::complete_tx) through a 4-stage ring;
//   * eight consumer warps unpack fields straight out of shared memory with funnel shifts,
//     evaluate every filter leaf as one wrap-around range test, build LSB-first bitset
//     words with __ballot_sync, combine leaves with the AND/OR program in registers,
//     popcount, and (optionally) reduce the matching rows of the value columns;
//   * decoded column vectors never exist in HBM — only bitset words, per-pack counts and
//     per-CTA partial aggregates are written.
//
// Bitset word layout: a little-endian 32-bit word of KnoxDB's bitset (row 8k+i ↔ bit i of
// byte k, internal/bitset/bitset.go:23-29) is exactly the __ballot_sync mask of 32
// consecutive rows.
//
// Replaces (reference, CPU): internal/encode/bitpack/cmp.go:20-130 + cmp_{eq,lt,le,bw}.go,
// internal/cmp/number.go:13-243, float.go:13-242, internal/bitset/generic/bitset.go
// (And/Or/PopCount), internal/operator/filter/match_core.go:44-215 and the reducers of
// internal/reducer/reducer.go:138-314.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"
#include "kx_kernels.h"

namespace kx {

// ------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra KX_DONE;\n"
        "bra KX_WAIT;\n"
        "KX_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global → shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// ------------------------------------------------------------------------------ field access
// `w`-bit field starting at bit offset `bit` of a little-endian bit string (32-bit words).
__device__ __forceinline__ uint64_t load_field(const uint32_t* __restrict__ words, uint64_t bit, uint32_t w) {
    uint64_t idx = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t w0 = words[idx], w1 = words[idx + 1];
    uint32_t lo = __funnelshift_r(w0, w1, sh);
    if (w <= 32) return lo & (uint32_t)width_mask((int)w);
    uint32_t w2 = words[idx + 2];
    uint32_t hi = __funnelshift_r(w1, w2, sh);
    return (((uint64_t)hi << 32) | lo) & width_mask((int)w);
}

__device__ __forceinline__ uint32_t run_of_row(const uint32_t* __restrict__ ends, uint32_t nruns, uint32_t row) {
    uint32_t lo = 0, hi = nruns;   // first run with ends[k] >= row (ends are inclusive)
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (__ldg(ends + m) >= row) hi = m; else lo = m + 1;
    }
    return lo;
}

// value of row `row` of a block as the sign-/zero-extended 64-bit pattern of T (IEEE bits for
// floats).  `staged`: the tile's bit stream in shared memory (row index relative to the tile),
// or nullptr to read the block's stream from global memory.
__device__ __forceinline__ uint64_t decode_value(const ColView& v, uint32_t row, const uint32_t* staged, uint32_t row_in_tile) {
    switch (v.kind) {
    case CK_CONST: return v.base;
    case CK_DELTA: return type_ext(v.type, (uint64_t)row * v.delta + v.base);
    case CK_BITS: {
        uint64_t f;
        if (staged) f = load_field(staged, (uint64_t)row_in_tile * v.width, v.width);
        else if (v.width == 64) f = __ldg(reinterpret_cast<const unsigned long long*>(v.data) + row);
        else f = load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width);
        if (type_is_float(v.type)) return f;
        return type_ext(v.type, f + v.base);
    }
    case CK_DICT: {
        uint64_t code = staged ? load_field(staged, (uint64_t)row_in_tile * v.width, v.width)
                               : load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width);
        code += v.delta;
        return __ldg(reinterpret_cast<const unsigned long long*>(v.aux) + code);
    }
    case CK_RUNEND: {
        uint32_t k = run_of_row(reinterpret_cast<const uint32_t*>(v.aux), v.naux, row);
        return __ldg(reinterpret_cast<const unsigned long long*>(v.data) + k);
    }
    }
    return 0;
}

__device__ __forceinline__ bool set_has(const uint64_t* __restrict__ s, uint32_t n, uint64_t v) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (__ldg(s + m) < v) lo = m + 1; else hi = m;
    }
    return lo < n && __ldg(s + lo) == v;
}

// ------------------------------------------------------------------------------ leaf kernels
// Every function evaluates one leaf for the warp's chunk of 32*R rows and returns the chunk's
// bitset in "word per lane" form: lane j (< R) holds the bitset word of rows [32j, 32j+32) of
// the chunk.

// ---- fast path, width W <= 32 (compile time): each lane owns 32 CONSECUTIVE rows = exactly W
// 32-bit words of the stream.  After full unrolling every field position is a constant, so a
// row costs one shift that brings the field to the TOP of a register (low garbage bits are
// harmless for the compare), an optional subtract, one compare and one predicated OR — no
// ballot, no mask, and the W words arrive with 128/64/32-bit shared-memory loads.
// The compare ((f - a) mod 2^W) <= d becomes (t - (a << K)) <= ((d << K) | (2^K - 1)), K = 32 - W.
template <int W, bool SUB>
__device__ __forceinline__ uint32_t leaf_b32(const uint32_t* __restrict__ seg, uint32_t a_top, uint32_t lim) {
    uint32_t x[W + 1];
    if constexpr (W % 4 == 0) {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            uint4 v = reinterpret_cast<const uint4*>(seg)[i];
            x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
    } else if constexpr (W % 2 == 0) {
#pragma unroll
        for (int i = 0; i < W / 2; ++i) {
            uint2 v = reinterpret_cast<const uint2*>(seg)[i];
            x[2 * i] = v.x; x[2 * i + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < W; ++i) x[i] = seg[i];
    }
    x[W] = 0;
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * W, wi = bit >> 5, sh = bit & 31;
        uint32_t t;
        if (sh + W <= 32) t = x[wi] << (32 - sh - W);
        else t = __funnelshift_l(x[wi], x[wi + 1], 64 - sh - W);
        if (SUB) t -= a_top;
        if (t <= lim) word |= (1u << j);
    }
    return word;
}

template <bool SUB>
__device__ __noinline__ uint32_t leaf_b32_dispatch(const uint32_t* __restrict__ sw, uint32_t group, uint32_t w, uint32_t a_top, uint32_t lim) {
    const uint32_t* seg = sw + group * w;
    switch (w) {
#define KX_CASE(W) case W: return leaf_b32<W, SUB>(seg, a_top, lim);
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
        KX_CASE(17) KX_CASE(18) KX_CASE(19) KX_CASE(20) KX_CASE(21) KX_CASE(22) KX_CASE(23) KX_CASE(24)
        KX_CASE(25) KX_CASE(26) KX_CASE(27) KX_CASE(28) KX_CASE(29) KX_CASE(30) KX_CASE(31) KX_CASE(32)
#undef KX_CASE
    }
    return 0;
}

// one LM_RANGE32 leaf for the warp chunk; lanes >= R own no rows
__device__ __forceinline__ uint32_t leaf_range32(const uint32_t* __restrict__ sw, uint32_t w, uint32_t row0, uint32_t R, uint32_t lane,
                                                 uint32_t a, uint32_t d) {
    if (lane >= R) return 0;
    const uint32_t k = 32u - w;
    const uint32_t a_top = a << k, lim = (d << k) | ((1u << k) - 1u);   // k == 0: a, d
    const uint32_t group = (row0 >> 5) + lane;
    return a ? leaf_b32_dispatch<true>(sw, group, w, a_top, lim) : leaf_b32_dispatch<false>(sw, group, w, 0u, lim);
}

// ---- general path (33..64-bit fields): lane l handles rows l, l+32, l+64 … so its bit offset
// advances by exactly w 32-bit words per iteration and its shift stays constant; bitset words
// are built with __ballot_sync.
__device__ __forceinline__ uint32_t leaf_range64(const uint32_t* __restrict__ sw, uint32_t w, uint32_t row0, uint32_t R, uint32_t lane,
                                                 uint64_t a, uint64_t d, uint64_t wm) {
    uint32_t bit = (row0 + lane) * w;
    uint32_t idx = bit >> 5, sh = bit & 31u;
    uint64_t fm = width_mask((int)w);
    uint32_t word = 0;
    if (w <= 32) {
#pragma unroll 4
        for (uint32_t it = 0; it < R; ++it) {
            uint64_t f = __funnelshift_r(sw[idx], sw[idx + 1], sh) & (uint32_t)fm;
            uint32_t b = __ballot_sync(0xffffffffu, ((f - a) & wm) <= d);
            if (lane == it) word = b;
            idx += w;
        }
    } else {
#pragma unroll 4
        for (uint32_t it = 0; it < R; ++it) {
            uint32_t w0 = sw[idx], w1 = sw[idx + 1], w2 = sw[idx + 2];
            uint64_t f = (((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh)) & fm;
            uint32_t b = __ballot_sync(0xffffffffu, ((f - a) & wm) <= d);
            if (lane == it) word = b;
            idx += w;
        }
    }
    return word;
}

template <typename F>
__device__ __forceinline__ bool float_pred(uint32_t op, F x, F a, F b) {
    switch (op) {   // IEEE ordered-quiet compares, != true on NaN (internal/cmp/float.go:13-242)
    case 1: return x == a;
    case 2: return x != a;
    case 3: return x > a;
    case 4: return x >= a;
    case 5: return x < a;
    case 6: return x <= a;
    case 9: return a <= x && x <= b;
    }
    return false;
}

__device__ __forceinline__ uint32_t leaf_float(const uint32_t* __restrict__ sw, uint32_t w, uint32_t row0, uint32_t R, uint32_t lane,
                                               uint32_t op, uint64_t a, uint64_t b) {
    uint32_t word = 0;
    if (w == 64) {
        const double* sd = reinterpret_cast<const double*>(sw);
        double da = __longlong_as_double((long long)a), db = __longlong_as_double((long long)b);
#pragma unroll 4
        for (uint32_t it = 0; it < R; ++it) {
            uint32_t bal = __ballot_sync(0xffffffffu, float_pred<double>(op, sd[row0 + it * 32 + lane], da, db));
            if (lane == it) word = bal;
        }
    } else {
        const float* sf = reinterpret_cast<const float*>(sw);
        float fa = __uint_as_float((uint32_t)a), fb = __uint_as_float((uint32_t)b);
#pragma unroll 4
        for (uint32_t it = 0; it < R; ++it) {
            uint32_t bal = __ballot_sync(0xffffffffu, float_pred<float>(op, sf[row0 + it * 32 + lane], fa, fb));
            if (lane == it) word = bal;
        }
    }
    return word;
}

// generic per-row leaves (IN / NIN sets, run-end blocks): value decode + test
__device__ __forceinline__ uint32_t leaf_generic(const PackLeaf& L, const ColView& v, const uint32_t* staged, uint32_t pack_row0,
                                                 uint32_t row0, uint32_t R, uint32_t lane, uint32_t nrows, const uint64_t* __restrict__ sets) {
    uint32_t word = 0;
    for (uint32_t it = 0; it < R; ++it) {
        uint32_t rt = row0 + it * 32 + lane;        // row within tile
        uint32_t row = pack_row0 + rt;              // row within pack
        bool p = false;
        if (row < nrows) {
            uint64_t val = decode_value(v, row, staged, rt);
            if (L.mode == LM_SET) p = set_has(sets + L.a, (uint32_t)L.d, val);
            else p = ((val ^ L.wm) - L.a) <= L.d;   // LM_VALRANGE
        }
        uint32_t b = __ballot_sync(0xffffffffu, p);
        if (lane == it) word = b;
    }
    return word;
}

// ------------------------------------------------------------------------------ aggregates
// Per-thread accumulator of one value column: four 64-bit slots, meaning depends on the type
//   integers: s0 = sum mod 2^64, s1 = min, s2 = max (order-preserving unsigned domain)
//   float64 : s0 = running sum, s1 = Neumaier compensation, s2 = min, s3 = max (IEEE bits)
// The match count / validity is shared by all value columns of a thread.
struct AggAcc { uint64_t s[4]; };

__device__ __forceinline__ double as_f64(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t as_u64(double d) { return (uint64_t)__double_as_longlong(d); }

__device__ __forceinline__ void agg_add(AggAcc& A, int type, uint64_t bits, bool first) {
    if (type == 9) {   // float64: compensated running sum (deterministic per thread)
        double x = as_f64(bits), sum = as_f64(A.s[0]), err = as_f64(A.s[1]);
        double t = sum + x;
        err += (fabs(sum) >= fabs(x)) ? ((sum - t) + x) : ((x - t) + sum);
        A.s[0] = as_u64(t); A.s[1] = as_u64(err);
        if (first || x < as_f64(A.s[2])) A.s[2] = bits;
        if (first || x > as_f64(A.s[3])) A.s[3] = bits;
    } else {
        A.s[0] += bits;   // wraps mod 2^64; narrower T is truncated on the host
        uint64_t k = type_is_signed(type) ? bits ^ 0x8000000000000000ull : bits;
        if (first || k < A.s[1]) A.s[1] = k;
        if (first || k > A.s[2]) A.s[2] = k;
    }
}

// double-double style merge of two compensated sums
__device__ __forceinline__ void fsum_merge(double& s, double& e, double s2, double e2) {
    double t = s + s2;
    double c = (fabs(s) >= fabs(s2)) ? ((s - t) + s2) : ((s2 - t) + s);
    s = t;
    e += e2 + c;
}

// merge B into A; both non-empty
__device__ __forceinline__ void agg_merge(AggAcc& A, const AggAcc& B, int type) {
    if (type == 9) {
        double s = as_f64(A.s[0]), e = as_f64(A.s[1]);
        fsum_merge(s, e, as_f64(B.s[0]), as_f64(B.s[1]));
        A.s[0] = as_u64(s); A.s[1] = as_u64(e);
        if (as_f64(B.s[2]) < as_f64(A.s[2])) A.s[2] = B.s[2];
        if (as_f64(B.s[3]) > as_f64(A.s[3])) A.s[3] = B.s[3];
    } else {
        A.s[0] += B.s[0];
        if (B.s[1] < A.s[1]) A.s[1] = B.s[1];
        if (B.s[2] > A.s[2]) A.s[2] = B.s[2];
    }
}

// ------------------------------------------------------------------------------ the kernel
// first pack whose tile range contains tile t (packs with zero tiles are skipped)
__device__ __forceinline__ uint32_t pack_of_tile(const PackInfo* __restrict__ packs, uint32_t npacks, uint32_t t) {
    uint32_t lo = 0, hi = npacks;   // last pack with tile0 <= t
    while (hi - lo > 1) {
        uint32_t m = (lo + hi) >> 1;
        if (packs[m].tile0 <= t) lo = m; else hi = m;
    }
    return lo;
}

// SIMPLE = one leaf, no aggregates: the hot configuration (fused decode + compare + popcount)
template <bool SIMPLE>
__global__ void __launch_bounds__(SCAN_THREADS, 2) scan_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + STAGES;
    uint8_t* stage_base = smem + 128;
    __shared__ AggAcc warp_acc[CONSUMER_WARPS];
    __shared__ unsigned long long warp_cnt[CONSUMER_WARPS];

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t R = P.R, tile_rows = R * 32u * CONSUMER_WARPS;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // contiguous tile range of this CTA: consecutive tiles stay inside one pack (descriptor reuse,
    // sequential DRAM pages) and the split is static, so results are reproducible run to run
    const uint32_t t_begin = (uint32_t)(((uint64_t)blockIdx.x * P.ntiles) / gridDim.x);
    const uint32_t t_end = (uint32_t)(((uint64_t)(blockIdx.x + 1) * P.ntiles) / gridDim.x);
    if (t_begin >= t_end) return;

    uint32_t pack = pack_of_tile(P.packs, P.npacks, t_begin);
    PackInfo pi = P.packs[pack];
    uint32_t pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
    uint32_t chunk = t_begin - pi.tile0;
    auto next_tile = [&]() {   // advance (pack, chunk) to the following tile
        if (++chunk >= pack_tiles) {
            do { ++pack; pi = P.packs[pack]; } while (pi.n == 0);
            pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
            chunk = 0;
        }
    };

    if (warp == CONSUMER_WARPS) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            uint32_t k = 0;
            for (uint32_t t = t_begin; t < t_end; ++t, ++k) {
                uint32_t s = k % STAGES, ph = (k / STAGES) & 1u;
                mbar_wait(&empty_bar[s], ph ^ 1u);          // slot released by all consumer warps
                uint32_t rows = min(tile_rows, pi.n - chunk * tile_rows);
                const PackLeaf* L = P.leaves + (size_t)pack * P.nleaves;
                uint32_t total = 0;
                for (uint32_t l = 0; l < P.nleaves; ++l)
                    if (L[l].data) total += (((rows * (uint32_t)L[l].width + 7u) >> 3) + 15u) & ~15u;
                mbar_expect_tx(&full_bar[s], total);        // arrive (count 1) + expected bytes
                uint8_t* dst = stage_base + (size_t)s * P.stage_bytes;
                for (uint32_t l = 0; l < P.nleaves; ++l) {
                    if (!L[l].data) continue;
                    uint32_t w = L[l].width;
                    uint32_t bytes = (((rows * w + 7u) >> 3) + 15u) & ~15u;
                    const uint8_t* src = L[l].data + (size_t)chunk * (tile_rows / 8u) * w;
                    tma_load_1d(dst, src, bytes, &full_bar[s]);
                    dst += (tile_rows / 8u) * w + 16u;      // slot = full-tile bytes + over-read pad
                }
                if (t + 1 < t_end) next_tile();
            }
        }
        return;
    }

    // ===================== consumers: unpack + filter + reduce =====================
    AggAcc acc[SIMPLE ? 1 : MAX_AGGS];
#pragma unroll
    for (int j = 0; j < (SIMPLE ? 1 : MAX_AGGS); ++j) acc[j] = AggAcc{};
    unsigned long long nmatch = 0;   // rows this thread reduced
    uint32_t lane_cnt = 0;           // matches of the current pack seen by this lane
    const uint32_t row0 = warp * R * 32u;   // first row of the warp chunk within a tile

    auto flush_count = [&](uint32_t pk) {
        uint32_t c = __reduce_add_sync(0xffffffffu, lane_cnt);
        if (P.counts && lane == 0 && c) atomicAdd(P.counts + pk, (unsigned long long)c);
        lane_cnt = 0;
    };

    uint32_t k = 0;
    for (uint32_t t = t_begin; t < t_end; ++t, ++k) {
        const uint32_t s = k % STAGES, ph = (k / STAGES) & 1u;
        const PackLeaf* L = P.leaves + (size_t)pack * P.nleaves;
        const uint32_t pack_row0 = chunk * tile_rows;          // first row of the tile within the pack
        const uint8_t* stage = stage_base + (size_t)s * P.stage_bytes;

        mbar_wait(&full_bar[s], ph);                           // TMA bytes have landed

        auto eval_leaf = [&](const PackLeaf& lf, const uint32_t* sw) -> uint32_t {
            uint32_t word;
            switch (lf.mode) {
            case LM_NONE: word = 0; break;
            case LM_ALL: word = 0xffffffffu; break;
            case LM_RANGE32: word = leaf_range32(sw, lf.width, row0, R, lane, (uint32_t)lf.a, (uint32_t)lf.d); break;
            case LM_RANGE64: word = leaf_range64(sw, lf.width, row0, R, lane, lf.a, lf.d, lf.wm); break;
            case LM_FLOAT: word = leaf_float(sw, lf.width, row0, R, lane, lf.fop, lf.a, lf.d); break;
            case LM_ROWRANGE: {
                // rows [a, a+d] of the pack → bits of this lane's word
                uint64_t r = (uint64_t)pack_row0 + row0 + lane * 32u;     // first row of the word
                uint64_t lo = lf.a, hi = lf.a + lf.d;
                word = 0;
                if (hi >= r && lo < r + 32u) {
                    uint32_t b0 = lo > r ? (uint32_t)(lo - r) : 0u;
                    uint32_t b1 = hi < r + 31u ? (uint32_t)(hi - r) : 31u;
                    word = (0xffffffffu >> (31u - b1)) & (0xffffffffu << b0);
                }
                break;
            }
            default:
                word = leaf_generic(lf, P.views[lf.view], lf.data ? sw : nullptr, pack_row0, row0, R, lane, pi.n, P.set_vals);
                break;
            }
            return lf.neg ? ~word : word;
        };

        uint32_t word;
        if (SIMPLE) {
            word = eval_leaf(L[0], reinterpret_cast<const uint32_t*>(stage));
        } else {
            // evaluate the leaves and the AND/OR program on word-per-lane bitsets
            uint32_t stack[MAX_LEAVES];
            uint32_t leaf_off[MAX_LEAVES];
            int sp = 0;
            uint32_t off = 0;
            for (uint32_t l = 0; l < P.nleaves; ++l) {
                leaf_off[l] = off;
                if (L[l].data) off += (tile_rows / 8u) * L[l].width + 16u;
            }
            for (uint32_t i = 0; i < P.npost; ++i) {
                uint32_t op = P.postfix[i];
                if (op < 0x80u) {
                    stack[sp++] = eval_leaf(L[op], reinterpret_cast<const uint32_t*>(stage + leaf_off[op]));
                } else {
                    uint32_t y = stack[--sp];
                    stack[sp - 1] = (op == 0xFEu) ? (stack[sp - 1] & y) : (stack[sp - 1] | y);
                }
            }
            word = stack[0];
        }

        // all shared-memory reads of this stage are done: hand the slot back to the producer early
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);

        // mask rows past the end of the pack (tail bits must be zero) and lanes >= R
        const uint64_t wr = (uint64_t)pack_row0 + row0 + lane * 32u;   // first row of this lane's word
        {
            uint32_t valid = 0;
            if (lane < R && wr < pi.n) {
                uint32_t left = pi.n - (uint32_t)wr;
                valid = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
            }
            word &= valid;
        }

        // outputs: bitset words (coalesced 128 B per warp), per-pack match count
        if (P.bitsets && lane < R && wr < pi.n)
            *reinterpret_cast<uint32_t*>(P.bitsets + pi.bitset_off + (wr >> 3)) = word;
        lane_cnt += __popc(word);

        // fused reduce over the matching rows of the value columns (read on demand)
        if (!SIMPLE && P.naggs && __any_sync(0xffffffffu, word != 0)) {
            for (uint32_t it = 0; it < R; ++it) {
                uint32_t wd = __shfl_sync(0xffffffffu, word, it);
                if (wd == 0) continue;
                if ((wd >> lane) & 1u) {
                    uint32_t row = pack_row0 + row0 + it * 32u + lane;
#pragma unroll
                    for (int j = 0; j < MAX_AGGS; ++j) {
                        if (j < (int)P.naggs) {
                            const ColView& v = P.views[P.agg_view0 + (size_t)pack * P.naggs + j];
                            agg_add(acc[SIMPLE ? 0 : j], P.agg_type[j], decode_value(v, row, nullptr, 0), nmatch == 0);
                        }
                    }
                    ++nmatch;
                }
            }
        }

        if (t + 1 < t_end) {
            const uint32_t prev = pack;
            next_tile();
            if (pack != prev) flush_count(prev);
        }
    }
    flush_count(pack);

    if (SIMPLE) return;

    // ---- per-CTA partial aggregates: fixed-order tree inside the warp, then across warps
    for (uint32_t j = 0; j < P.naggs; ++j) {
        const int type = P.agg_type[j];
        AggAcc a = acc[SIMPLE ? 0 : j];
        unsigned long long c = nmatch;
        for (int off = 16; off > 0; off >>= 1) {
            AggAcc b;
#pragma unroll
            for (int q = 0; q < 4; ++q) b.s[q] = __shfl_down_sync(0xffffffffu, a.s[q], off);
            unsigned long long cb = __shfl_down_sync(0xffffffffu, c, off);
            if (cb) { if (c) agg_merge(a, b, type); else a = b; }
            c += cb;
        }
        if (lane == 0) { warp_acc[warp] = a; warp_cnt[warp] = c; }
        // consumer-only barrier (the producer warp has exited)
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
        if (threadIdx.x == 0) {
            AggAcc r = warp_acc[0];
            unsigned long long rc = warp_cnt[0];
            for (int q = 1; q < CONSUMER_WARPS; ++q) {
                if (warp_cnt[q]) { if (rc) agg_merge(r, warp_acc[q], type); else r = warp_acc[q]; }
                rc += warp_cnt[q];
            }
            AggPartial o;
            o.count = rc; o.valid = rc != 0; o.pad = 0;
            if (type == 9) { o.sum = r.s[0]; o.err = as_f64(r.s[1]); o.mn = r.s[2]; o.mx = r.s[3]; }
            else { o.sum = r.s[0]; o.err = 0.0; o.mn = r.s[1]; o.mx = r.s[2]; }
            P.partials[(size_t)blockIdx.x * P.naggs + j] = o;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
    }
}

// Combines the per-CTA partials in CTA order (fixed topology → bit-reproducible results).
__global__ void finalize_kernel(const AggPartial* parts, uint32_t nparts, uint32_t naggs, const uint8_t* agg_type4, AggPartial* out) {
    uint32_t j = threadIdx.x;
    if (j >= naggs) return;
    int type = agg_type4[j];
    AggPartial r{};
    for (uint32_t i = 0; i < nparts; ++i) {
        const AggPartial p = parts[(size_t)i * naggs + j];
        if (!p.valid) continue;
        if (!r.valid) { r = p; continue; }
        r.count += p.count;
        if (type == 9) {
            double s = __longlong_as_double((long long)r.sum), e = r.err;
            fsum_merge(s, e, __longlong_as_double((long long)p.sum), p.err);
            r.sum = (uint64_t)__double_as_longlong(s); r.err = e;
            if (__longlong_as_double((long long)p.mn) < __longlong_as_double((long long)r.mn)) r.mn = p.mn;
            if (__longlong_as_double((long long)p.mx) > __longlong_as_double((long long)r.mx)) r.mx = p.mx;
        } else {
            r.sum += p.sum;
            if (p.mn < r.mn) r.mn = p.mn;
            if (p.mx > r.mx) r.mx = p.mx;
        }
    }
    out[j] = r;
}

// ------------------------------------------------------------------------------ small kernels

// bitset.{And,AndNot,Or,Xor} with any/all flags: internal/bitset/generic/bitset.go:13-295
__global__ void bitset_op_kernel(uint32_t* dst, const uint32_t* src, uint64_t nbits, int op, unsigned int* flags /*[0]=any,[1]=notall*/) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t any = 0, notall = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t d = dst[i], s = src[i], r;
        switch (op) {
        case 0: r = d & s; break;
        case 1: r = d & ~s; break;
        case 2: r = d | s; break;
        default: r = d ^ s; break;
        }
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        r &= mask;
        dst[i] = r;
        any |= r;
        notall |= (r ^ mask);
    }
    if (__any_sync(0xffffffffu, any != 0) && (threadIdx.x & 31) == 0) atomicOr(&flags[0], 1u);
    if (__any_sync(0xffffffffu, notall != 0) && (threadIdx.x & 31) == 0) atomicOr(&flags[1], 1u);
}

__global__ void bitset_neg_kernel(uint32_t* buf, uint64_t nbits) {
    uint64_t nwords = (nbits + 31) >> 5;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        buf[i] = ~buf[i] & mask;
    }
}

__global__ void bitset_popcount_kernel(const uint32_t* buf, uint64_t nbits, unsigned long long* out) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        c += __popc(buf[i] & mask);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// Bitset.Indexes (internal/bitset/iterator.go:269-290): pass 1 = per-block popcounts,
// pass 2 (after an exclusive scan on the host side of the stream) = ordered scatter.
__global__ void bitset_block_counts_kernel(const uint32_t* buf, uint64_t nbits, uint32_t* block_counts) {
    // one block handles 256 consecutive words
    uint64_t nwords = (nbits + 31) >> 5;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t c = 0;
    if (i < nwords) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        c = __popc(buf[i] & mask);
    }
    __shared__ uint32_t ws[8];
    uint32_t s = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; ++q) t += ws[q]; block_counts[blockIdx.x] = t; }
}

__global__ void exclusive_scan_kernel(uint32_t* v, uint32_t n, unsigned long long* total) {
    // single-block scan (n = #256-word blocks; ≤ 2^32/8192 entries)
    __shared__ uint32_t carry;
    __shared__ uint32_t ws[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n ? v[i] : 0, incl = x;
        for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0, wi = w;
            for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, wi, off); if (threadIdx.x >= (uint32_t)off) wi += y; }
            ws[threadIdx.x] = wi - w;
        }
        __syncthreads();
        uint32_t excl = carry + ws[threadIdx.x >> 5] + incl - x;
        if (i < n) v[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void bitset_scatter_kernel(const uint32_t* buf, uint64_t nbits, const uint32_t* block_offs, uint32_t* dst) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t wv = 0;
    if (i < nwords) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        wv = buf[i] & mask;
    }
    uint32_t c = __popc(wv), incl = c;
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t q = 0; q < (threadIdx.x >> 5); ++q) woff += ws[q];
    uint32_t pos = block_offs[blockIdx.x] + woff + incl - c;
    uint32_t base = (uint32_t)(i << 5);
    while (wv) { uint32_t b = __ffs(wv) - 1; dst[pos++] = base + b; wv &= wv - 1; }
}

// NumberContainer.AppendTo(dst, nil) / bitpack.Decode: one thread per row
__global__ void decode_kernel(ColView v, uint8_t* dst) {
    int nb = type_bits(v.type) / 8;
    for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < v.n; row += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = decode_value(v, (uint32_t)row, nullptr, 0);
        switch (nb) {
        case 8: reinterpret_cast<uint64_t*>(dst)[row] = x; break;
        case 4: reinterpret_cast<uint32_t*>(dst)[row] = (uint32_t)x; break;
        case 2: reinterpret_cast<uint16_t*>(dst)[row] = (uint16_t)x; break;
        default: dst[row] = (uint8_t)x; break;
        }
    }
}

// Zone-map + bloom pruning: one thread per pack (stats.matchVector, internal/pack/stats/match.go:92-195)
__global__ void prune_kernel(PruneParams P) {
    uint32_t pack = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (pack < P.npacks) {
        uint32_t stack = 0; int sp = 0;   // bit stack (depth <= 32)
        for (uint32_t i = 0; i < P.npost; ++i) {
            uint32_t op = P.postfix[i];
            if (op < 0x80u) {
                const PruneLeaf& L = P.leaves[op];
                uint64_t mn = P.mins[(size_t)pack * P.nleaves + op], mx = P.maxs[(size_t)pack * P.nleaves + op];
                uint64_t f = L.flip;
                uint64_t kmn = mn ^ f, kmx = mx ^ f, ka = L.a ^ f, kb = L.b ^ f;
                bool m = true;
                if (L.is_float) {
                    double dmn = __longlong_as_double((long long)mn), dmx = __longlong_as_double((long long)mx);
                    double da = __longlong_as_double((long long)L.a), db = __longlong_as_double((long long)L.b);
                    switch (L.mode) {
                    case 1: m = dmn <= da && dmx >= da; break;
                    case 3: m = dmx > da; break;
                    case 4: m = dmx >= da; break;
                    case 5: m = dmn < da; break;
                    case 6: m = dmn <= da; break;
                    case 9: m = dmn <= db && dmx >= da; break;
                    default: m = true;
                    }
                } else {
                    switch (L.mode) {   // MatchRangeVectors, internal/operator/filter/match_num.go
                    case 1: m = kmn <= ka && kmx >= ka; break;            // EQ  :357-371
                    case 3: m = kmx > ka; break;                           // GT  :429-434
                    case 4: m = kmx >= ka; break;                          // GE  :460-465
                    case 5: m = kmn < ka; break;                           // LT  :491-496
                    case 6: m = kmn <= ka; break;                          // LE  :522-527
                    case 9: m = kmn <= kb && kmx >= ka; break;             // RG  :573-588
                    case 7: {                                              // IN  :693-735 set.ContainsRange(min,max) in uint64 order
                        uint64_t lo = mn, hi = mx;
                        if (lo > hi) { uint64_t t2 = lo; lo = hi; hi = t2; }
                        const uint64_t* s = P.set_vals + L.set_off;
                        uint32_t l2 = 0, h2 = L.nset;
                        while (l2 < h2) { uint32_t mid = (l2 + h2) >> 1; if (s[mid] < lo) l2 = mid + 1; else h2 = mid; }
                        m = l2 < L.nset && s[l2] <= hi;
                        break;
                    }
                    default: m = true;                                     // NE :394-401, NIN :810-817 (undecided)
                    }
                }
                // bloom probe for EQ / IN when the pack carries a filter (match.go:141-192)
                if (m && P.blooms && (L.mode == 1 || L.mode == 7)) {
                    const uint8_t* bf = P.blooms[(size_t)pack * P.nleaves + op];
                    if (bf) {
                        uint32_t mbits = (uint32_t)((P.bloom_len[(size_t)pack * P.nleaves + op] - 1) * 8);
                        uint32_t mask = mbits - 1u, kk = bf[0];
                        const uint8_t* bits = bf + 1;
                        bool anyhit = false;
                        for (uint32_t h = P.hash_off[op]; h < P.hash_off[op + 1] && !anyhit; ++h) {
                            uint64_t hv = P.hashes[h];
                            uint32_t h0 = (uint32_t)hv, h1 = (uint32_t)(hv >> 32);
                            bool hit = true;
                            for (uint32_t q = 0; q < kk && hit; ++q) {      // bloom.go:136-150
                                hit = (bits[(h0 & mask) >> 3] >> (h0 & 7u)) & 1u;
                                h0 += h1;
                            }
                            anyhit = hit;
                        }
                        if (P.hash_off[op + 1] > P.hash_off[op]) m = anyhit;
                    }
                }
                stack = (stack << 1) | (m ? 1u : 0u); ++sp;
            } else {
                uint32_t y = stack & 1u; stack >>= 1; --sp;
                uint32_t x = stack & 1u;
                stack = (stack & ~1u) | (op == 0xFEu ? (x & y) : (x | y));
            }
        }
        alive = (stack & 1u) != 0;
    }
    uint32_t b = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && pack < P.npacks) P.out[pack >> 5] = b;
    uint32_t c = __popc(b);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(P.count, (unsigned long long)c);
}

// ------------------------------------------------------------------------------ launchers
static int grid_for(uint64_t items, uint64_t cap) { uint64_t g = (items + 255) / 256; return (int)(g < cap ? g : cap); }

cudaError_t launch_scan(const ScanParams& P, int grid, size_t smem_bytes, cudaStream_t stream) {
    const bool simple = P.nleaves == 1 && P.naggs == 0;
    auto kern = simple ? scan_kernel<true> : scan_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    kern<<<grid, SCAN_THREADS, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_finalize(const AggPartial* parts, uint32_t nparts, uint32_t naggs, const uint8_t* agg_type_dev, AggPartial* out, cudaStream_t stream) {
    finalize_kernel<<<1, 32, 0, stream>>>(parts, nparts, naggs, agg_type_dev, out);
    return cudaGetLastError();
}

cudaError_t launch_bitset_op(uint32_t* dst, const uint32_t* src, uint64_t nbits, int op, unsigned int* flags, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_op_kernel<<<grid ? grid : 1, 256, 0, stream>>>(dst, src, nbits, op, flags);
    return cudaGetLastError();
}
cudaError_t launch_bitset_neg(uint32_t* buf, uint64_t nbits, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_neg_kernel<<<grid ? grid : 1, 256, 0, stream>>>(buf, nbits);
    return cudaGetLastError();
}
cudaError_t launch_bitset_popcount(const uint32_t* buf, uint64_t nbits, unsigned long long* out, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_popcount_kernel<<<grid ? grid : 1, 256, 0, stream>>>(buf, nbits, out);
    return cudaGetLastError();
}
cudaError_t launch_bitset_indexes(const uint32_t* buf, uint64_t nbits, uint32_t* block_tmp, unsigned long long* total, uint32_t* dst, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t nblocks = (uint32_t)((nwords + 255) / 256);
    if (nblocks == 0) return cudaMemsetAsync(total, 0, 8, stream);
    bitset_block_counts_kernel<<<nblocks, 256, 0, stream>>>(buf, nbits, block_tmp);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(block_tmp, nblocks, total);
    bitset_scatter_kernel<<<nblocks, 256, 0, stream>>>(buf, nbits, block_tmp, dst);
    return cudaGetLastError();
}
cudaError_t launch_decode(const ColView& v, void* dst, cudaStream_t stream) {
    int grid = grid_for(v.n, 148 * 16);
    decode_kernel<<<grid ? grid : 1, 256, 0, stream>>>(v, reinterpret_cast<uint8_t*>(dst));
    return cudaGetLastError();
}
cudaError_t launch_prune(const PruneParams& P, cudaStream_t stream) {
    int grid = (int)((P.npacks + 255) / 256);
    prune_kernel<<<grid ? grid : 1, 256, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
