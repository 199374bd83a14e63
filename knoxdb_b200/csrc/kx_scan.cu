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
#include "kx_decode.cuh"

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
// producer-side wait: the producer only has to notice a released slot "soon"; sleeping between polls keeps its
// spin loop from stealing issue slots (and power) from the eight consumer warps of the CTA
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}\n" : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(128);
    }
}
// TMA 1-D bulk copy global → shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
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
// A tile is 256*R rows: every consumer warp owns R consecutive 32-row groups of it and walks them
// in passes of up to 32 groups.  Every leaf function evaluates one leaf for one pass and returns
// the pass's bitset in "word per lane" form: lane j (< Rp) holds the bitset word of group g0 + j
// of the tile (rows [32 (g0+j), 32 (g0+j) + 32)).

// Shared-memory bank conflicts of the fast path: lane j reads the W words of its own group, i.e.
// the lanes of a quarter warp are W words apart.  That is conflict-free for every width except
// W = 8, 16, 24, 32, where the 128-bit chunks of neighbouring lanes fall onto the same banks.  For
// those widths lane j starts `rot` chunks into its group (wrapping around); because the chunks of
// these widths hold whole rows the lane simply computes a ROTATED bitset word and rotates it back.
template <int W> __device__ __forceinline__ int rot_chunks(uint32_t lane) {
    if constexpr (W == 32) return (int)(lane & 7u);             // 8 chunks of 4 rows
    else if constexpr (W == 16) return (int)((lane >> 1) & 3u); // 4 chunks of 8 rows
    else if constexpr (W == 8) return (int)((lane >> 2) & 1u);  // 2 chunks of 16 rows
    else if constexpr (W == 24) return (int)((lane >> 2) & 1u) * 3;   // 6 chunks, 3 chunks = 16 rows
    else return 0;
}

// ---- fast path, width W <= 32 (compile time): each lane owns 32 CONSECUTIVE rows = exactly W
// 32-bit words of the stream.  After full unrolling every field position is a constant, so a
// row costs one shift that brings the field to the TOP of a register (low garbage bits are
// harmless for the compare), an optional subtract, one compare and one predicated OR — no
// ballot, no mask, and the W words arrive with 128/64/32-bit shared-memory loads.
// The compare ((f - a) mod 2^W) <= d becomes (t - (a << K)) <= ((d << K) | (2^K - 1)), K = 32 - W.
template <int W, bool SUB>
__device__ __forceinline__ uint32_t leaf_b32(const uint32_t* __restrict__ seg, uint32_t lane, uint32_t a_top, uint32_t lim) {
    __builtin_assume(__isShared(seg));   // the staged stream lives in shared memory: LDS, not generic loads
    uint32_t x[W + 1];
    int rot_rows = 0;
    if constexpr (W % 4 == 0) {
        constexpr int NC = W / 4;
        const int rc = rot_chunks<W>(lane);
        rot_rows = (rc * 128) / W;
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            int c = i + rc;
            if (c >= NC) c -= NC;
            uint4 v = reinterpret_cast<const uint4*>(seg)[c];
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
    uint32_t wq[4] = {0, 0, 0, 0};   // four independent accumulators: short dependency chains, one predicated OR per row
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * W, wi = bit >> 5, sh = bit & 31;
        uint32_t t;
        if (sh + W <= 32) t = x[wi] << (32 - sh - W);
        else t = __funnelshift_l(x[wi], x[wi + 1], 64 - sh - W);
        if (SUB) t -= a_top;
        if (t <= lim) wq[j & 3] |= (1u << j);
    }
    uint32_t word = (wq[0] | wq[1]) | (wq[2] | wq[3]);
    if constexpr (W == 8 || W == 16 || W == 24 || W == 32) word = __funnelshift_l(word, word, rot_rows);
    return word;
}

template <bool SUB>
__device__ __noinline__ uint32_t leaf_b32_dispatch(const uint32_t* __restrict__ seg, uint32_t lane, uint32_t w, uint32_t a_top, uint32_t lim) {
    switch (w) {
#define KX_CASE(W) case W: return leaf_b32<W, SUB>(seg, lane, a_top, lim);
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
        KX_CASE(17) KX_CASE(18) KX_CASE(19) KX_CASE(20) KX_CASE(21) KX_CASE(22) KX_CASE(23) KX_CASE(24)
        KX_CASE(25) KX_CASE(26) KX_CASE(27) KX_CASE(28) KX_CASE(29) KX_CASE(30) KX_CASE(31) KX_CASE(32)
#undef KX_CASE
    }
    return 0;
}

// one LM_RANGE32 leaf for one pass; lanes >= Rp own no group
__device__ __forceinline__ uint32_t leaf_range32(const uint32_t* __restrict__ sw, uint32_t w, uint32_t g0, uint32_t Rp, uint32_t lane,
                                                 uint32_t a, uint32_t d) {
    if (lane >= Rp) return 0;
    const uint32_t k = 32u - w;
    const uint32_t a_top = a << k, lim = (d << k) | ((1u << k) - 1u);   // k == 0: a, d
    const uint32_t* seg = sw + (size_t)(g0 + lane) * w;
    return a ? leaf_b32_dispatch<true>(seg, lane, w, a_top, lim) : leaf_b32_dispatch<false>(seg, lane, w, 0u, lim);
}

// ---- fast path for 33..63-bit fields (compile-time width): same lane-owns-32-consecutive-rows layout,
// 64-bit top-aligned arithmetic: T = field << (64 - W) (low garbage bits harmless), (T - a_top) <= lim.
template <int W, bool SUB>
__device__ __forceinline__ uint32_t leaf_b64(const uint32_t* __restrict__ seg, uint64_t a_top, uint64_t lim) {
    __builtin_assume(__isShared(seg));
    uint32_t x[W + 2];
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
    x[W] = 0; x[W + 1] = 0;
    uint32_t word = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * W, wi = bit >> 5, e = (bit & 31) + W;   // field = bits [e - W, e) of x[wi], x[wi+1], x[wi+2]
        uint32_t hi, lo;
        if (e <= 64) { hi = __funnelshift_l(x[wi], x[wi + 1], 64 - e); lo = x[wi] << (64 - e); }
        else { hi = __funnelshift_l(x[wi + 1], x[wi + 2], 96 - e); lo = __funnelshift_l(x[wi], x[wi + 1], 96 - e); }
        uint64_t t = ((uint64_t)hi << 32) | lo;
        if (SUB) t -= a_top;
        if (t <= lim) word |= (1u << j);
    }
    return word;
}

template <bool SUB>
__device__ __noinline__ uint32_t leaf_b64_dispatch(const uint32_t* __restrict__ seg, uint32_t w, uint64_t a_top, uint64_t lim) {
    switch (w) {
#define KX_CASE(W) case W: return leaf_b64<W, SUB>(seg, a_top, lim);
        KX_CASE(33) KX_CASE(34) KX_CASE(35) KX_CASE(36) KX_CASE(37) KX_CASE(38) KX_CASE(39) KX_CASE(40)
        KX_CASE(41) KX_CASE(42) KX_CASE(43) KX_CASE(44) KX_CASE(45) KX_CASE(46) KX_CASE(47) KX_CASE(48)
        KX_CASE(49) KX_CASE(50) KX_CASE(51) KX_CASE(52) KX_CASE(53) KX_CASE(54) KX_CASE(55) KX_CASE(56)
        KX_CASE(57) KX_CASE(58) KX_CASE(59) KX_CASE(60) KX_CASE(61) KX_CASE(62) KX_CASE(63)
#undef KX_CASE
    }
    return 0;
}

// LM_RANGE64 leaf for one pass.  33..63-bit fields take the compile-time-width path above; 64-bit
// streams (raw uint64/int64, full-width bit-packing) are lane-strided — lane l handles rows l, l+32, …
// with one LDS.64 per row — and build bitset words with __ballot_sync.
__device__ __forceinline__ uint32_t leaf_range64(const uint32_t* __restrict__ sw, uint32_t w, uint32_t g0, uint32_t Rp, uint32_t lane,
                                                 uint64_t a, uint64_t d, uint64_t wm) {
    __builtin_assume(__isShared(sw));
    uint32_t word = 0;
    if (w == 64) {
        const unsigned long long* s64 = reinterpret_cast<const unsigned long long*>(sw) + (size_t)g0 * 32u + lane;
#pragma unroll 8
        for (uint32_t it = 0; it < Rp; ++it) {
            uint32_t b = __ballot_sync(0xffffffffu, (s64[it * 32u] - a) <= d);
            if (lane == it) word = b;
        }
        return word;
    }
    if (w > 32) {
        if (lane >= Rp) return 0;
        const uint32_t k = 64u - w;
        const uint64_t a_top = a << k, lim = (d << k) | ((1ull << k) - 1ull);
        const uint32_t* seg = sw + (size_t)(g0 + lane) * w;
        return a ? leaf_b64_dispatch<true>(seg, w, a_top, lim) : leaf_b64_dispatch<false>(seg, w, 0ull, lim);
    }
    // <= 32-bit fields evaluated in 64-bit arithmetic (not produced by the host translation; kept for completeness)
    uint32_t bit = (g0 * 32u + lane) * w;
    uint32_t idx = bit >> 5, sh = bit & 31u;
    uint64_t fm = width_mask((int)w);
#pragma unroll 4
    for (uint32_t it = 0; it < Rp; ++it) {
        uint64_t f = __funnelshift_r(sw[idx], sw[idx + 1], sh) & (uint32_t)fm;
        uint32_t b = __ballot_sync(0xffffffffu, ((f - a) & wm) <= d);
        if (lane == it) word = b;
        idx += w;
    }
    return word;
}

// IEEE ordered-quiet compares, != true on NaN (internal/cmp/float.go:13-242); OP = types.FilterMode
template <int OP, typename F>
__device__ __forceinline__ bool float_pred(F x, F a, F b) {
    if constexpr (OP == 1) return x == a;
    else if constexpr (OP == 2) return x != a;
    else if constexpr (OP == 3) return x > a;
    else if constexpr (OP == 4) return x >= a;
    else if constexpr (OP == 5) return x < a;
    else if constexpr (OP == 6) return x <= a;
    else return a <= x && x <= b;
}

template <int OP, typename F>
__device__ __forceinline__ uint32_t leaf_float_op(const F* __restrict__ sf, uint32_t Rp, uint32_t lane, F a, F b) {
    __builtin_assume(__isShared(sf));
    uint32_t word = 0;
#pragma unroll 8
    for (uint32_t it = 0; it < Rp; ++it) {
        uint32_t bal = __ballot_sync(0xffffffffu, float_pred<OP, F>(sf[it * 32u], a, b));
        if (lane == it) word = bal;
    }
    return word;
}

template <typename F>
__device__ __forceinline__ uint32_t leaf_float_t(const F* __restrict__ sf, uint32_t Rp, uint32_t lane, uint32_t op, F a, F b) {
    switch (op) {
    case 1: return leaf_float_op<1, F>(sf, Rp, lane, a, b);
    case 2: return leaf_float_op<2, F>(sf, Rp, lane, a, b);
    case 3: return leaf_float_op<3, F>(sf, Rp, lane, a, b);
    case 4: return leaf_float_op<4, F>(sf, Rp, lane, a, b);
    case 5: return leaf_float_op<5, F>(sf, Rp, lane, a, b);
    case 6: return leaf_float_op<6, F>(sf, Rp, lane, a, b);
    case 9: return leaf_float_op<9, F>(sf, Rp, lane, a, b);
    }
    return 0;
}

__device__ __forceinline__ uint32_t leaf_float(const uint32_t* __restrict__ sw, uint32_t w, uint32_t g0, uint32_t Rp, uint32_t lane,
                                               uint32_t op, uint64_t a, uint64_t b) {
    if (w == 64)
        return leaf_float_t<double>(reinterpret_cast<const double*>(sw) + (size_t)g0 * 32u + lane, Rp, lane, op,
                                    __longlong_as_double((long long)a), __longlong_as_double((long long)b));
    return leaf_float_t<float>(reinterpret_cast<const float*>(sw) + (size_t)g0 * 32u + lane, Rp, lane, op,
                               __uint_as_float((uint32_t)a), __uint_as_float((uint32_t)b));
}

// ---- IN / NOT IN on a dictionary block (DictionaryContainer.MatchInSet, int_dict.go:361-398): the set
// was translated into a bitmap over the pack's codes (translateSet :400) by codeset_kernel; the consumers
// copy the current pack's bitmap (<= 8 KB) into shared memory and each lane tests the 32 codes of its own
// group.  The bitmap covers every code a W-bit field can produce (the host sizes and zeroes it), so there
// is no bounds check; bits are shifted in row by row.
template <int W>
__device__ __forceinline__ uint32_t leaf_code32(const uint32_t* __restrict__ seg, uint32_t code_base, const uint32_t* __restrict__ bm) {
    __builtin_assume(__isShared(seg));
    __builtin_assume(__isShared(bm));   // the pack's code bitmap is cached in shared memory (one LDS per row)
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
        uint32_t f = (sh + W <= 32) ? (x[wi] >> sh) : __funnelshift_r(x[wi], x[wi + 1], sh);
        uint32_t code = (f & ((1u << W) - 1u)) + code_base;
        uint32_t wv = bm[code >> 5];
        word = __funnelshift_r(word, __funnelshift_r(wv, 0u, code), 1);   // shift bit (code & 31) of wv in from the top
    }
    return word;   // after 32 steps row j sits at bit j
}

__device__ __noinline__ uint32_t leaf_code32_dispatch(const uint32_t* __restrict__ seg, uint32_t w, uint32_t code_base, const uint32_t* __restrict__ bm) {
    switch (w) {
#define KX_CASE(W) case W: return leaf_code32<W>(seg, code_base, bm);
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
#undef KX_CASE
    }
    return 0;
}

__device__ __forceinline__ uint32_t leaf_codeset(const uint32_t* __restrict__ sw, uint32_t w, uint32_t g0, uint32_t Rp, uint32_t lane,
                                                 uint32_t code_base, const uint32_t* __restrict__ bm) {
    if (lane >= Rp) return 0;
    return leaf_code32_dispatch(sw + (size_t)(g0 + lane) * w, w, code_base, bm);   // dictionary codes are uint16: w <= 16
}

// ---- IN / NOT IN on a bit-packed / raw integer block (int_bitpack.go:249-291, int_raw.go:339-380).
// Phase 1 walks the pass lane-strided (lane l takes row 32 it + l: consecutive fields, conflict-free shared-memory
// reads), hashes the decoded value T(field + For) with two multiply-adds and tests ONE bit of the leaf's prefilter
// bitmap in shared memory; the ballots of the rows that pass become candidate words (word per lane).  Phase 2: every
// lane verifies the candidates of its own group against the exact set — a bucketised hash table (4 keys per 32 B
// bucket, built by the host at kx_prog_compile; empty slots hold keys of other buckets, so a plain compare of the
// four slots is exact), in shared memory when it is small, else in global memory.
// field of WIDE ? 33..64 : 1..32 bits at bit offset `bit` of a shared-memory stream, as the 64-bit pattern of T
// (EXT: T is narrower than 64 bits — truncate and sign-/zero-extend, `sh` = 64 - bits(T))
template <bool WIDE, bool EXT>
__device__ __forceinline__ uint64_t hs_value(const uint32_t* __restrict__ sw, uint32_t bit, uint32_t w, uint64_t base, uint32_t sh, bool sgn) {
    const uint32_t idx = bit >> 5, s = bit & 31u;
    const uint32_t w0 = sw[idx], w1 = sw[idx + 1];
    uint64_t f;
    if (WIDE) {
        const uint32_t w2 = sw[idx + 2];
        f = (((uint64_t)__funnelshift_r(w1, w2, s) << 32) | __funnelshift_r(w0, w1, s)) & (w >= 64u ? ~0ull : ((1ull << w) - 1ull));
    } else {
        f = __funnelshift_r(w0, w1, s) & (w >= 32u ? 0xffffffffu : ((1u << w) - 1u));
    }
    uint64_t val = f + base;
    if (EXT) val = sgn ? (uint64_t)((int64_t)(val << sh) >> sh) : ((val << sh) >> sh);
    return val;
}

template <bool WIDE, bool EXT>
__device__ __forceinline__ uint32_t leaf_hashset_t(const uint32_t* __restrict__ sw, uint32_t w, uint64_t base, uint32_t sh, bool sgn, uint32_t g0,
                                                   uint32_t Rp, uint32_t lane, const uint32_t* __restrict__ pre, uint32_t pre_log2,
                                                   const ulonglong2* __restrict__ tab, uint32_t tab_log2, uint32_t keep) {
    __builtin_assume(__isShared(sw));
    __builtin_assume(__isShared(pre));
    const uint32_t pre_shift = 32u - pre_log2;
    uint32_t cand = 0;
    uint32_t bit = (g0 * 32u + lane) * w;
#pragma unroll 4
    for (uint32_t it = 0; it < Rp; ++it, bit += 32u * w) {
        const uint32_t idx = set_hash32(hs_value<WIDE, EXT>(sw, bit, w, base, sh, sgn)) >> pre_shift;
        const uint32_t b = __ballot_sync(0xffffffffu, (pre[idx >> 5] >> (idx & 31u)) & 1u);
        if (lane == it) cand = b;
    }
    cand &= keep;   // rows the enclosing AND has already ruled out need no verification
    uint32_t word = 0;
    const uint32_t gbit = (g0 + lane) * 32u * w, tab_shift = 32u - tab_log2;
    while (cand) {
        const uint32_t j = (uint32_t)__ffs((int)cand) - 1u;
        cand &= cand - 1u;
        const uint64_t val = hs_value<WIDE, EXT>(sw, gbit + j * w, w, base, sh, sgn);
        const ulonglong2* b = tab + 2u * (size_t)(set_hash32(val) >> tab_shift);
        const ulonglong2 p = b[0], q = b[1];
        word |= (uint32_t)((p.x == val) | (p.y == val) | (q.x == val) | (q.y == val)) << j;
    }
    return word;
}

__device__ __forceinline__ uint32_t leaf_hashset(const uint32_t* __restrict__ sw, uint32_t w, int type, uint64_t base, uint32_t g0, uint32_t Rp, uint32_t lane,
                                              const uint32_t* __restrict__ pre, uint32_t pre_log2, const ulonglong2* __restrict__ tab, uint32_t tab_log2,
                                              uint32_t keep) {
    const uint32_t sh = 64u - (uint32_t)type_bits(type);
    const bool sgn = type_is_signed(type);
    if (w > 32u) return sh ? leaf_hashset_t<true, true>(sw, w, base, sh, sgn, g0, Rp, lane, pre, pre_log2, tab, tab_log2, keep)
                           : leaf_hashset_t<true, false>(sw, w, base, 0u, false, g0, Rp, lane, pre, pre_log2, tab, tab_log2, keep);
    return sh ? leaf_hashset_t<false, true>(sw, w, base, sh, sgn, g0, Rp, lane, pre, pre_log2, tab, tab_log2, keep)
              : leaf_hashset_t<false, false>(sw, w, base, 0u, false, g0, Rp, lane, pre, pre_log2, tab, tab_log2, keep);
}

// ---- run-end blocks (RunEndContainer.Match* + applyMatch, int_runend.go:224-318): the predicate is
// evaluated on run VALUES; each lane finds the run of its group's first row once and walks forward.
__device__ __forceinline__ uint32_t leaf_runend(const PackLeaf& L, const ColView& v, uint32_t grow0, uint32_t nrows, bool active,
                                                const uint64_t* __restrict__ sets) {
    if (!active || grow0 >= nrows) return 0;
    const uint32_t* ends = reinterpret_cast<const uint32_t*>(v.aux);
    const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(v.data);
    const uint32_t rend = min(grow0 + 32u, nrows);
    uint32_t k = run_of_row(ends, v.naux, grow0);
    uint32_t word = 0, r = grow0;
    while (r < rend && k < v.naux) {
        uint32_t hi = min(__ldg(ends + k), rend - 1u);     // inclusive
        uint64_t val = __ldg(vals + k);
        bool p = (L.mode == LM_SET) ? set_has(sets + L.a, (uint32_t)L.d, val) : ((val ^ L.wm) - L.a) <= L.d;
        if (p) word |= (0xffffffffu >> (31u - (hi - grow0))) & (0xffffffffu << (r - grow0));
        r = hi + 1u; ++k;
    }
    return word;
}

// generic per-row fallback (sets on affine blocks, …): value decode + test
__device__ __forceinline__ uint32_t leaf_generic(const PackLeaf& L, const ColView& v, const uint32_t* staged, uint32_t pack_row0,
                                                 uint32_t g0, uint32_t Rp, uint32_t lane, uint32_t nrows, const uint64_t* __restrict__ sets) {
    uint32_t word = 0;
    for (uint32_t it = 0; it < Rp; ++it) {
        uint32_t rt = (g0 + it) * 32u + lane;       // row within tile
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

// Dictionary-set translation (DictionaryContainer.translateSet, int_dict.go:400-440) on the device: one thread per
// SET value binary-searches the pack's dictionary (sorted, unique, in T order: `flip` maps it to unsigned order) and
// sets the bit of the code it finds.  Runs on the scan stream right before scan_kernel; blockIdx.y = (pack, leaf) job.
__global__ void codeset_kernel(const CodesetJob* __restrict__ jobs, const uint64_t* __restrict__ set_vals, uint32_t* __restrict__ out) {
    const CodesetJob J = jobs[blockIdx.y];
    const unsigned long long* dict = reinterpret_cast<const unsigned long long*>(J.dict);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < J.nset; i += gridDim.x * blockDim.x) {
        const uint64_t val = __ldg(set_vals + J.set_off + i), key = val ^ J.flip;
        uint32_t lo = 0, hi = J.ndict;   // first dictionary entry >= val
        while (lo < hi) {
            uint32_t m = (lo + hi) >> 1;
            if ((__ldg(dict + m) ^ J.flip) < key) lo = m + 1; else hi = m;
        }
        if (lo < J.ndict && __ldg(dict + lo) == val) atomicOr(out + J.out_off + (lo >> 5), 1u << (lo & 31u));
    }
}

// ------------------------------------------------------------------------------ aggregates
// Per-thread accumulator of one value column: four 64-bit slots, meaning depends on the type
//   integers: s0 = sum mod 2^64, s1 = min, s2 = max (order-preserving unsigned domain)
//   float64 : s0 = running sum, s1 = Neumaier compensation, s2 = min, s3 = max (IEEE bits)
// Accumulators start at the identity (min = +max, max = -max); the match count decides validity.
struct AggAcc { uint64_t s[4]; };

__device__ __forceinline__ double as_f64(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t as_u64(double d) { return (uint64_t)__double_as_longlong(d); }

__device__ __forceinline__ AggAcc agg_identity(int type) {
    AggAcc A;
    if (type == 9) { A.s[0] = 0; A.s[1] = 0; A.s[2] = 0x7ff0000000000000ull; A.s[3] = 0xfff0000000000000ull; }
    else { A.s[0] = 0; A.s[1] = ~0ull; A.s[2] = 0; A.s[3] = 0; }
    return A;
}

__device__ __forceinline__ void agg_add(AggAcc& A, int type, uint64_t bits) {
    if (type == 9) {   // float64: compensated running sum (deterministic per thread)
        double x = as_f64(bits), sum = as_f64(A.s[0]), err = as_f64(A.s[1]);
        double t = sum + x;
        err += (fabs(sum) >= fabs(x)) ? ((sum - t) + x) : ((x - t) + sum);
        A.s[0] = as_u64(t); A.s[1] = as_u64(err);
        if (x < as_f64(A.s[2])) A.s[2] = bits;
        if (x > as_f64(A.s[3])) A.s[3] = bits;
    } else {
        A.s[0] += bits;   // wraps mod 2^64; narrower T is truncated on the host
        uint64_t k = type_is_signed(type) ? bits ^ 0x8000000000000000ull : bits;
        if (k < A.s[1]) A.s[1] = k;
        if (k > A.s[2]) A.s[2] = k;
    }
}

// double-double style merge of two compensated sums
__device__ __forceinline__ void fsum_merge(double& s, double& e, double s2, double e2) {
    double t = s + s2;
    double c = (fabs(s) >= fabs(s2)) ? ((s - t) + s2) : ((s2 - t) + s);
    s = t;
    e += e2 + c;
}

// merge B into A (identities merge as no-ops)
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

// The fused reduce for ONE value column over `ng` consecutive 32-row groups whose match words sit in shared
// memory (`fw`).  Lane l first looks at the word of group l: one ballot tells the warp which groups have matches at
// all, and only those are visited (a sparse tile costs a handful of instructions).  For a visited group lane l
// reduces row l, so every load instruction reads 32 consecutive values (coalesced); the loads of up to B groups are
// issued back to back before they are consumed (memory-level parallelism).
// SMEM: `vp` points into a ring stage (the producer staged the tile's slice of the column); otherwise the matching
// rows are read on demand from global memory.  Groups are visited in ascending order and the (lane, row) assignment
// is the same in both variants: a tile gives bit-identical partial sums whichever way its values arrive.
template <int BD, int BS, typename Body>
__device__ __forceinline__ void for_matching_groups(const uint32_t* __restrict__ fw, uint32_t ng, uint32_t lane, Body&& body) {
    __builtin_assume(__isShared(fw));
    for (uint32_t blk = 0; blk < ng; blk += 32u) {
        const uint32_t myw = blk + lane < ng ? fw[blk + lane] : 0u;
        uint32_t mask = __ballot_sync(0xffffffffu, myw != 0u);
        const uint32_t nb = min(32u, ng - blk);
        if (__popc(mask) * 2u >= nb) {
            // most groups have matches: walk them all, BD at a time (no bit scans; words read as shared-memory broadcasts)
            for (uint32_t it0 = 0; it0 < nb; it0 += BD) {
                uint32_t gi[BD], wd[BD];
#pragma unroll
                for (int u = 0; u < BD; ++u) {
                    gi[u] = blk + it0 + u;
                    wd[u] = it0 + u < nb ? fw[gi[u]] : 0u;
                }
                body(gi, wd);
            }
            continue;
        }
        while (mask) {   // few groups have matches: visit only those, BS at a time
            uint32_t gi[BS], wd[BS];
#pragma unroll
            for (int u = 0; u < BS; ++u) {
                gi[u] = mask ? (uint32_t)__ffs((int)mask) - 1u : 0u;
                wd[u] = __shfl_sync(0xffffffffu, myw, gi[u]);
                if (!mask) wd[u] = 0u;
                mask &= mask - 1u;
                gi[u] += blk;
            }
            body(gi, wd);
        }
    }
}

template <bool F64, bool SMEM>
__device__ __forceinline__ void agg_groups_raw64(AggAcc& A, const unsigned long long* __restrict__ vp, const uint32_t* __restrict__ fw, uint32_t ng,
                                                 uint32_t lane, uint64_t base, uint64_t flip) {
    if (SMEM) __builtin_assume(__isShared(vp));
    for_matching_groups<8, 2>(fw, ng, lane, [&](const auto& gi, const auto& wd) {
        constexpr int B = (int)(sizeof(gi) / sizeof(gi[0]));
        uint64_t val[B];
#pragma unroll
        for (int u = 0; u < B; ++u) val[u] = ((wd[u] >> lane) & 1u) ? (SMEM ? vp[(size_t)gi[u] * 32u] : __ldg(vp + (size_t)gi[u] * 32u)) : 0ull;
#pragma unroll
        for (int u = 0; u < B; ++u) {
            if ((wd[u] >> lane) & 1u) {
                if (F64) {
                    double x = as_f64(val[u]), sum = as_f64(A.s[0]), err = as_f64(A.s[1]);
                    double t = sum + x;
                    err += (fabs(sum) >= fabs(x)) ? ((sum - t) + x) : ((x - t) + sum);
                    A.s[0] = as_u64(t); A.s[1] = as_u64(err);
                    if (x < as_f64(A.s[2])) A.s[2] = val[u];
                    if (x > as_f64(A.s[3])) A.s[3] = val[u];
                } else {
                    uint64_t v = val[u] + base, k = v ^ flip;
                    A.s[0] += v;
                    if (k < A.s[1]) A.s[1] = k;
                    if (k > A.s[2]) A.s[2] = k;
                }
            }
        }
    });
}

// any other value column layout (bit-packed, dictionary, affine, run-end, narrow types, ALP): decode per row.
// `row0` = pack row of (first group, this lane); `staged` = the column's bit stream from pack row `srow0` on (or nullptr).
__device__ __forceinline__ void agg_groups_generic(AggAcc& A, const ColView& v, int type, uint32_t row0, const uint32_t* __restrict__ fw, uint32_t ng,
                                                   uint32_t lane, const uint32_t* staged, uint32_t srow0) {
    for_matching_groups<4, 2>(fw, ng, lane, [&](const auto& gi, const auto& wd) {
        constexpr int B = (int)(sizeof(gi) / sizeof(gi[0]));
        uint64_t val[B];
#pragma unroll
        for (int u = 0; u < B; ++u)
            if ((wd[u] >> lane) & 1u) val[u] = decode_value(v, row0 + gi[u] * 32u, staged, row0 + gi[u] * 32u - srow0);
#pragma unroll
        for (int u = 0; u < B; ++u)
            if ((wd[u] >> lane) & 1u) agg_add(A, type, val[u]);
    });
}

// ALP blocks: rows that are patches carry their true value outside the encoded stream.  One thread per patch
// evaluates the float predicate on it (the loops over `vals, pos` of float_alp.go:238-495) and sets the bit of
// its row in the leaf's correction stream (zeroed before the launch), which the scan ORs in / ANDs out.
__global__ void alpfix_kernel(const AlpFixJob* __restrict__ jobs, uint8_t* __restrict__ out_base) {
    const AlpFixJob J = jobs[blockIdx.y];
    const uint32_t* pos = reinterpret_cast<const uint32_t*>(J.blob);
    const double* vals = reinterpret_cast<const double*>(J.blob + alp_vals_off(J.np));
    uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
    const double a = __longlong_as_double((long long)J.a), b = __longlong_as_double((long long)J.b);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < J.np; k += gridDim.x * blockDim.x) {
        const double x = vals[k];
        bool p;
        switch (J.mode) {
        case 1: p = (a != a) ? (x != x) : (x == a); break;   // MatchEqual: a NaN operand matches NaN patches
        case 3: p = x > a; break;
        case 4: p = x >= a; break;
        case 5: p = x < a; break;
        case 6: p = x <= a; break;
        default: p = x >= a && x <= b; break;                  // MatchBetween
        }
        if (p != (J.invert != 0)) { uint32_t r = pos[k]; atomicOr(out + (r >> 5), 1u << (r & 31u)); }
    }
}

// Run-end blocks: the predicate is evaluated once per RUN by this pre-pass (RunEndContainer.Match* +
// applyMatch, internal/encode/int_runend.go:224-318: match the run values, SetRange(start, end) per
// matching run); the scan kernel then streams the resulting per-leaf bitset like a 1-bit column.
__global__ void runfill_kernel(const RunFillJob* __restrict__ jobs, const uint64_t* __restrict__ set_vals, uint8_t* __restrict__ out_base) {
    const RunFillJob J = jobs[blockIdx.y];
    const uint32_t* ends = reinterpret_cast<const uint32_t*>(J.ends);
    const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(J.vals);
    uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < J.nruns; k += gridDim.x * blockDim.x) {
        uint64_t val = __ldg(vals + k);
        bool p = J.is_set ? set_has(set_vals + J.a, (uint32_t)J.d, val) : ((val ^ J.wm) - J.a) <= J.d;
        if (!p) continue;
        uint32_t start = k ? __ldg(ends + k - 1) + 1u : 0u, end = __ldg(ends + k);   // inclusive
        if (end >= J.nrows) end = J.nrows - 1u;
        if (start > end) continue;
        uint32_t w0 = start >> 5, w1 = end >> 5;
        uint32_t m0 = 0xffffffffu << (start & 31u), m1 = 0xffffffffu >> (31u - (end & 31u));
        if (w0 == w1) { atomicOr(out + w0, m0 & m1); continue; }
        atomicOr(out + w0, m0);
        for (uint32_t w = w0 + 1; w < w1; ++w) out[w] = 0xffffffffu;   // words owned by this run alone
        atomicOr(out + w1, m1);
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

// SIMPLE = one leaf, no aggregates: the hot configuration (fused decode + compare + popcount).
// ONLY32 (with SIMPLE) = every pack's leaf is a <= 32-bit packed range test (or all / none): a lean
// instantiation without the other leaf paths (small code footprint, fewer registers), MINB CTAs per SM.
// AGG (general kernels only) = the launch reduces value columns; the filter-only instantiation carries none of that code.
template <bool SIMPLE, bool ONLY32, int MINB, bool AGG>
__global__ void __launch_bounds__(SCAN_THREADS, MINB) scan_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint8_t* stage_base = smem + 128;
    __shared__ AggAcc warp_acc[CONSUMER_WARPS];
    __shared__ unsigned long long warp_cnt[CONSUMER_WARPS];
    __shared__ unsigned int sm_match, sm_wtiles;   // matches / (warp, tile) pairs finished so far: selectivity feedback for the producer
    __shared__ uint32_t stage_flag[MAX_STAGES];    // per ring slot: does the stage carry a value-column chunk (1) or nothing (0)?

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t R = P.R, tile_rows = R * 32u * CONSUMER_WARPS, nstages = P.stages;
    uint32_t* code_smem = reinterpret_cast<uint32_t*>(stage_base + (size_t)nstages * P.stage_bytes);   // LM_CODESET bitmaps of the current pack
    // staged value columns: a tile's slice of a column = agg_chunks ring stages of chunk_rows rows (whole warps)
    const uint32_t agg_chunks = AGG ? P.agg_chunks : 0u, chunk_rows = agg_chunks ? tile_rows / agg_chunks : 0u;

    if (threadIdx.x == 0) {
        sm_match = 0; sm_wtiles = 0;
        for (uint32_t s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
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

    // Ring protocol.  SIMPLE kernels: one stage per TILE (the single leaf's stream; tiles may hold several
    // passes).  General kernels: one stage per (tile, staged leaf) in postfix order — the stage only has to hold
    // the widest column of 8192 rows, so multi-predicate programs keep full tiles and two CTAs per SM; the
    // AND/OR stack lives in registers while the ring advances from one leaf column to the next.
    if (warp == CONSUMER_WARPS) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            auto load_stream = [&](const uint8_t* data, size_t off, uint32_t w, uint32_t rows, uint32_t flag = 0u) {
                mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);  // slot released by all consumer warps
                if (!data) w = 0;
                data += off;
                uint32_t bytes = w ? ((((rows * w + 7u) >> 3) + 15u) & ~15u) : 0u;
                stage_flag[s] = flag;                       // published by the arrive below (release) / the consumers' wait (acquire)
                mbar_expect_tx(&full_bar[s], bytes);        // arrive (count 1) + expected bytes
                if (bytes) tma_load_1d(stage_base + (size_t)s * P.stage_bytes, data, bytes, &full_bar[s]);
                if (++s == nstages) { s = 0; ph ^= 1u; }
            };
            uint32_t pf_m0 = 0, pf_d0 = 0;   // selectivity feedback snapshot
            bool dense = P.agg_dense_thr == 0;
            for (uint32_t t = t_begin; t < t_end; ++t) {
                const uint32_t rows = min(tile_rows, pi.n - chunk * tile_rows);
                const PackLeaf* L = P.leaves + (size_t)pack * P.nleaves;
                const size_t tile_byte0 = (size_t)chunk * (tile_rows / 8u);   // * width = first byte of the tile in a stream
                if constexpr (SIMPLE) load_stream(L[0].data, tile_byte0 * L[0].width, L[0].width, rows);   // (an unstaged leaf still cycles its stage: lockstep)
                else {
                    // Value columns: tiles that match densely get their slice of every stageable value column streamed through
                    // the ring in agg_chunks chunks (full-bandwidth bulk copies); in sparse tiles the matching rows are read on
                    // demand from global memory.  The decision travels with the tile's first ring stage (stage_flag).
                    if (agg_chunks && P.agg_dense_thr != 0 && P.agg_dense_thr != 0xffffffffu) {
                        uint32_t m = *(volatile unsigned int*)&sm_match, d = *(volatile unsigned int*)&sm_wtiles;
                        if (d - pf_d0 >= CONSUMER_WARPS) {
                            dense = (uint64_t)(m - pf_m0) * P.agg_dense_thr * CONSUMER_WARPS > (uint64_t)(d - pf_d0) * tile_rows;
                            pf_m0 = m; pf_d0 = d;
                        }
                    }
                    const uint32_t flag = (agg_chunks && dense) ? 1u : 0u;
                    bool told = false;
                    for (uint32_t i = 0; i < P.npost; ++i) {
                        uint32_t op = P.postfix[i];
                        if (op >= 0x80u) continue;
                        if (L[op].data) { load_stream(L[op].data, tile_byte0 * L[op].width, L[op].width, rows, flag); told = true; }
                        if (L[op].fixmode) { load_stream(L[op].fix, tile_byte0, 1u, rows, flag); told = true; }      // ALP patch correction stream
                    }
                    if constexpr (AGG) {
                        if (!told) load_stream(nullptr, 0, 0u, 0u, flag);   // no leaf column is staged: an empty stage carries the decision
                        if (dense) {
                            for (uint32_t j = 0; j < P.naggs; ++j) {
                                const ColView& v = P.views[P.agg_view0 + (size_t)pack * P.naggs + j];
                                if (!agg_stageable(v)) continue;
                                for (uint32_t k = 0; k < agg_chunks; ++k) {
                                    const uint32_t r0 = k * chunk_rows, nr = rows > r0 ? min(chunk_rows, rows - r0) : 0u;
                                    load_stream(v.data, (tile_byte0 + r0 / 8u) * v.width, v.width, nr, 1u);
                                }
                            }
                        }
                    }
                }
                if (t + 1 < t_end) next_tile();
            }
        }
        return;
    }

    // ===================== consumers: unpack + filter + reduce =====================
    AggAcc acc[AGG ? MAX_AGGS : 1];
#pragma unroll
    for (int j = 0; j < (AGG ? MAX_AGGS : 1); ++j) acc[j] = agg_identity(AGG ? P.agg_type[j] : 0);
    unsigned long long nmatch = 0;   // matches this thread accounted for (per-CTA totals only)
    uint32_t lane_cnt = 0;           // matches of the current pack seen by this lane
    uint32_t bm_pack = 0xffffffffu;  // pack whose code bitmaps are cached in shared memory
    const uint32_t passes = (R + 31u) >> 5, Rp = min(R, 32u);
    // general kernels (shared memory behind the code bitmaps): the AND/OR stack of this warp, one word per (slot, pass,
    // lane), and the tile's final match words (CTA-shared, double-buffered by tile parity) for the fused reduce
    const uint32_t tile_groups = R * CONSUMER_WARPS;
    uint32_t* stk = code_smem + P.stack_off_words + warp * (P.stack_depth * passes * 32u);
    uint32_t* fin_base = code_smem + P.stack_off_words + CONSUMER_WARPS * P.stack_depth * passes * 32u;
    uint32_t fin_sel = 0;

    auto flush_count = [&](uint32_t pk) {
        uint32_t c = __reduce_add_sync(0xffffffffu, lane_cnt);
        if (P.counts && lane == 0 && c) atomicAdd(P.counts + pk, (unsigned long long)c);
        lane_cnt = 0;
    };

    if constexpr (!ONLY32) {
        // hash-set leaves: prefilter bitmaps (and small exact tables) are the same for every pack — copy them into shared
        // memory once per CTA
        bool any = false;
        for (uint32_t l = 0; l < P.nleaves; ++l) {
            if (!P.pre_log2[l]) continue;
            any = true;
            const uint32_t npre = (1u << P.pre_log2[l]) >> 5;
            for (uint32_t i = threadIdx.x; i < npre; i += CONSUMER_WARPS * 32u) code_smem[P.hs_smem_off[l] + i] = __ldg(P.set_pre + P.pre_off[l] + i);
            if (P.hs_tab_smem_off[l] != 0xffffffffu) {
                const uint32_t nt = 8u << P.tab_log2[l];
                const uint32_t* src = reinterpret_cast<const uint32_t*>(P.set_tabs + P.tab_off[l]);
                for (uint32_t i = threadIdx.x; i < nt; i += CONSUMER_WARPS * 32u) code_smem[P.hs_tab_smem_off[l] + i] = __ldg(src + i);
            }
        }
        if (any) asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
    }

    uint32_t s = 0, ph = 0;
    for (uint32_t t = t_begin; t < t_end; ++t) {
        const PackLeaf* L = P.leaves + (size_t)pack * P.nleaves;
        const uint32_t pack_row0 = chunk * tile_rows;          // first row of the tile within the pack

        if (!ONLY32 && P.code_bitmap_words && pack != bm_pack) {
            // new pack: the consumers copy its code bitmaps (built by codeset_kernel) into shared memory
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));   // everybody is done with the previous pack's bitmaps
            for (uint32_t l = 0; l < P.nleaves; ++l) {
                if (L[l].mode != LM_CODESET) continue;
                const uint32_t nw = (((1u << L[l].width) + (uint32_t)L[l].wm + 31u) >> 5) + 1u;
                const uint32_t* src = P.code_bits + L[l].a;
                uint32_t* dst = code_smem + P.code_smem_off[l];
                for (uint32_t i = threadIdx.x; i < nw; i += CONSUMER_WARPS * 32u) dst[i] = __ldg(src + i);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
            bm_pack = pack;
        }

        // one leaf for one pass: sw = the leaf's staged stream (or nullptr), g0 = first group of the pass
        // `keep`: rows (bits of this lane's word) whose result matters — the other operand of an enclosing AND
        auto eval_leaf = [&](uint32_t li, const uint32_t* sw, uint32_t g0, uint64_t wr, uint32_t keep = 0xffffffffu) -> uint32_t {
            const PackLeaf& lf = L[li];
            uint32_t word;
            if constexpr (ONLY32) {
                if (lf.mode == LM_RANGE32) word = leaf_range32(sw, lf.width, g0, Rp, lane, (uint32_t)lf.a, (uint32_t)lf.d);
                else word = lf.mode == LM_ALL ? 0xffffffffu : 0u;
                return lf.neg ? ~word : word;
            }
            switch (lf.mode) {
            case LM_NONE: word = 0; break;
            case LM_ALL: word = 0xffffffffu; break;
            case LM_RANGE32: word = leaf_range32(sw, lf.width, g0, Rp, lane, (uint32_t)lf.a, (uint32_t)lf.d); break;
            case LM_RANGE64: word = leaf_range64(sw, lf.width, g0, Rp, lane, lf.a, lf.d, lf.wm); break;
            case LM_FLOAT: word = leaf_float(sw, lf.width, g0, Rp, lane, lf.fop, lf.a, lf.d); break;
            case LM_ROWRANGE: {
                // rows [a, a+d] of the pack → bits of this lane's word
                uint64_t lo = lf.a, hi = lf.a + lf.d;
                word = 0;
                if (hi >= wr && lo < wr + 32u) {
                    uint32_t b0 = lo > wr ? (uint32_t)(lo - wr) : 0u;
                    uint32_t b1 = hi < wr + 31u ? (uint32_t)(hi - wr) : 31u;
                    word = (0xffffffffu >> (31u - b1)) & (0xffffffffu << b0);
                }
                break;
            }
            case LM_BITS: __builtin_assume(__isShared(sw)); word = lane < Rp ? sw[g0 + lane] : 0u; break;   // precomputed leaf bitset (run-end pre-pass)
            case LM_CODESET:
                word = leaf_codeset(sw, lf.width, g0, Rp, lane, (uint32_t)lf.wm, code_smem + P.code_smem_off[li]);
                break;
            case LM_HASHSET: {
                const uint32_t to = P.hs_tab_smem_off[li];
                const ulonglong2* tab = to != 0xffffffffu ? reinterpret_cast<const ulonglong2*>(code_smem + to)
                                                          : reinterpret_cast<const ulonglong2*>(P.set_tabs + P.tab_off[li]);
                const ColView& hv = P.views[lf.view];
                word = leaf_hashset(sw, hv.width, hv.type, hv.base, g0, Rp, lane, code_smem + P.hs_smem_off[li], P.pre_log2[li], tab, P.tab_log2[li], keep);
                break;
            }
            default: {
                const ColView& v = P.views[lf.view];
                if (v.kind == CK_RUNEND) word = leaf_runend(lf, v, (uint32_t)wr, pi.n, lane < Rp, P.set_vals);
                else word = leaf_generic(lf, v, lf.data ? sw : nullptr, pack_row0, g0, Rp, lane, pi.n, P.set_vals);
                break;
            }
            }
            return lf.neg ? ~word : word;
        };

        // tail masking, bitset store and popcount of one pass; returns the masked word
        auto emit = [&](uint32_t word, uint64_t wr) -> uint32_t {
            // mask rows past the end of the pack (tail bits must be zero) and lanes >= Rp
            uint32_t valid = 0;
            if (lane < Rp && wr < pi.n) {
                uint32_t left = pi.n - (uint32_t)wr;
                valid = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
            }
            word &= valid;
            // outputs: bitset words (coalesced 128 B per warp), per-pack match count
            if (P.bitsets && lane < Rp && wr < pi.n)
                *reinterpret_cast<uint32_t*>(P.bitsets + pi.bitset_off + (wr >> 3)) = word;
            lane_cnt += __popc(word);
            return word;
        };
        auto release = [&]() {   // this warp is done with ring stage s
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[s]);
            if (++s == nstages) { s = 0; ph ^= 1u; }
        };

        if constexpr (SIMPLE) {
            const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
            mbar_wait(&full_bar[s], ph);                           // TMA bytes have landed
            for (uint32_t pass = 0; pass < passes; ++pass) {
                const uint32_t g0 = warp * R + pass * 32u;         // first group (of the tile) of this pass
                const uint64_t wr = (uint64_t)pack_row0 + (uint64_t)(g0 + lane) * 32u;   // first pack row of this lane's word
                uint32_t word = eval_leaf(0, sw, g0, wr);
                if (L[0].neg2) word = ~word;   // float-level NOT of an ALP leaf without patches (with patches: general kernel)
                if (pass + 1 == passes) {   // all shared-memory reads of this stage are done: release it early
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty_bar[s]);
                }
                emit(word, wr);
            }
            if (++s == nstages) { s = 0; ph ^= 1u; }
        } else {
            // ---- leaves and the AND/OR program.  Every staged leaf column is one ring stage holding the column's slice of
            // the whole tile; a leaf is evaluated for ALL passes of the warp before the next one is touched (its unrolled
            // body stays hot in the instruction cache), the per-pass words wait on the warp's stack in shared memory.
            const uint32_t gw0 = warp * R, pstride = passes * 32u;
            uint32_t sp = 0;
            bool told = false, dense = false;   // the producer's staging decision arrives with the tile's first ring stage
            for (uint32_t i = 0; i < P.npost; ++i) {
                const uint32_t op = P.postfix[i];
                if (op < 0x80u) {
                    const PackLeaf& lf = L[op];
                    uint32_t* dst = stk + sp * pstride + lane;
                    const uint32_t* sw = nullptr;
                    const bool staged_leaf = lf.data != nullptr;
                    if (staged_leaf) {
                        mbar_wait(&full_bar[s], ph);
                        if (!told) { dense = stage_flag[s] != 0; told = true; }
                        sw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
                    }
                    const bool inv = lf.neg2 && !lf.fixmode;
                    // MatchAnd's early-out (match_core.go:44-130), per warp and pass: when this leaf is ANDed with the word
                    // on top of the stack next, rows that word has ruled out need no work — a pass whose 1024 rows are all
                    // ruled out skips the leaf altogether (time-range filters on ordered packs rule out whole tiles)
                    // (MatchOr's early-out, :132-215, is the mirror image: rows the other operand already matched.)
                    const bool and_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFEu;
                    const bool or_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFFu;
                    const uint32_t* prev = stk + (sp ? sp - 1u : 0u) * pstride + lane;
                    for (uint32_t pass = 0; pass < passes; ++pass) {
                        const uint32_t g0 = gw0 + pass * 32u;
                        const uint32_t keep = and_next ? prev[pass * 32u] : (or_next ? ~prev[pass * 32u] : 0xffffffffu);
                        uint32_t word = 0;
                        if (__any_sync(0xffffffffu, keep != 0u)) {
                            word = eval_leaf(op, sw, g0, (uint64_t)pack_row0 + (uint64_t)(g0 + lane) * 32u, keep);
                            if (inv) word = ~word;
                        }
                        dst[pass * 32u] = word;
                    }
                    if (staged_leaf) release();
                    if (lf.fixmode) {   // ALP: correct the rows that are patches (1-bit stream in the next stage)
                        mbar_wait(&full_bar[s], ph);
                        if (!told) { dense = stage_flag[s] != 0; told = true; }
                        const uint32_t* fw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
                        __builtin_assume(__isShared(fw));
                        for (uint32_t pass = 0; pass < passes; ++pass) {
                            const uint32_t fx = lane < Rp ? fw[gw0 + pass * 32u + lane] : 0u;
                            uint32_t word = dst[pass * 32u];
                            word = lf.fixmode == FIX_OR_PRED ? (word | fx) : (word & ~fx);
                            dst[pass * 32u] = lf.neg2 ? ~word : word;
                        }
                        release();
                    }
                    ++sp;
                } else {
                    --sp;
                    uint32_t* x = stk + (sp - 1u) * pstride + lane;
                    const uint32_t* y = stk + sp * pstride + lane;
                    for (uint32_t pass = 0; pass < passes; ++pass)
                        x[pass * 32u] = (op == 0xFEu) ? (x[pass * 32u] & y[pass * 32u]) : (x[pass * 32u] | y[pass * 32u]);
                }
            }
            // ---- outputs of the tile
            uint32_t* fin = fin_base + fin_sel * tile_groups;
            uint32_t tile_cnt = 0;
            for (uint32_t pass = 0; pass < passes; ++pass) {
                const uint32_t g0 = gw0 + pass * 32u;
                const uint32_t word = emit(stk[pass * 32u + lane], (uint64_t)pack_row0 + (uint64_t)(g0 + lane) * 32u);
                tile_cnt += __popc(word);
                if (AGG && lane < Rp) fin[g0 + lane] = word;
                if (AGG && word && __popc(word) <= 2) {
                    // sparse matches: start pulling their value rows towards L2 now; the reduce below (after the CTA barrier)
                    // then pays an L2 hit instead of a DRAM round trip per visited group
                    for (uint32_t j = 0; j < P.naggs; ++j) {
                        const ColView& v = P.views[P.agg_view0 + (size_t)pack * P.naggs + j];
                        if (v.kind != CK_BITS || v.width != 64) continue;
                        const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(v.data) + pack_row0 + (g0 + lane) * 32u;
                        uint32_t b0 = (uint32_t)__ffs((int)word) - 1u, b1 = 31u - (uint32_t)__clz((int)word);
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + b0));
                        if (b1 != b0) asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + b1));
                    }
                }
            }
            // ---- fused reduce over the matching rows of the value columns.  The tile's match words are shared by the
            // CTA: chunk k (groups [k G, (k + 1) G), G = tile groups / chunks) is reduced by ALL warps, warp w taking
            // G / 8 consecutive groups of it — the same assignment whether the chunk was staged through the ring by the
            // producer (dense tiles: full-bandwidth bulk copies) or its matching rows are read on demand from global
            // memory (sparse tiles: groups without a match are never touched; the column cycles ONE empty stage).
            if constexpr (AGG) {
                nmatch += tile_cnt;   // per-CTA totals only: any partition of the matches over threads will do
                const uint32_t c = __reduce_add_sync(0xffffffffu, tile_cnt);
                if (lane == 0) { atomicAdd(&sm_match, c); atomicAdd(&sm_wtiles, 1u); }   // selectivity feedback for the producer
                asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));              // the tile's words are complete
                fin_sel ^= 1u;
                const uint32_t K = P.agg_chunks, G = tile_groups / K, gpw = G / CONSUMER_WARPS;
                if (!told) {   // no leaf column was staged: the decision sits in an empty stage
                    mbar_wait(&full_bar[s], ph);
                    dense = stage_flag[s] != 0;
                    release();
                }
                for (uint32_t j = 0; j < P.naggs; ++j) {
                    const ColView& v = P.views[P.agg_view0 + (size_t)pack * P.naggs + j];
                    const int type = P.agg_type[j];
                    const bool raw64 = v.kind == CK_BITS && v.width == 64;
                    const uint64_t flip = type_is_signed(type) ? 0x8000000000000000ull : 0ull;
                    AggAcc a = acc[j];
                    const bool staged = dense && agg_stageable(v);
                    for (uint32_t k = 0; k < K; ++k) {
                        if (staged) mbar_wait(&full_bar[s], ph);
                        const uint32_t gk = k * G + warp * gpw;                 // this warp's first group of chunk k
                        const uint32_t row0 = pack_row0 + gk * 32u + lane;      // pack row of (group gk, this lane)
                        const uint32_t* fw = fin + gk;
                        const uint8_t* stg = stage_base + (size_t)s * P.stage_bytes;
                        if (raw64) {
                            if (staged) {
                                const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(stg) + (warp * gpw * 32u + lane);
                                if (type == 9) agg_groups_raw64<true, true>(a, vp, fw, gpw, lane, 0ull, 0ull);
                                else agg_groups_raw64<false, true>(a, vp, fw, gpw, lane, v.base, flip);
                            } else {
                                const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(v.data) + row0;
                                if (type == 9) agg_groups_raw64<true, false>(a, vp, fw, gpw, lane, 0ull, 0ull);
                                else agg_groups_raw64<false, false>(a, vp, fw, gpw, lane, v.base, flip);
                            }
                        } else {
                            agg_groups_generic(a, v, type, row0, fw, gpw, lane, staged ? reinterpret_cast<const uint32_t*>(stg) : nullptr,
                                               pack_row0 + k * G * 32u);
                        }
                        if (staged) release();
                    }
                    acc[j] = a;
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

    if constexpr (AGG) {

    // ---- per-CTA partial aggregates: fixed-order tree inside the warp, then across warps
    for (uint32_t j = 0; j < P.naggs; ++j) {
        const int type = P.agg_type[j];
        AggAcc a = acc[j];
        unsigned long long c = nmatch;
        for (int off = 16; off > 0; off >>= 1) {
            AggAcc b;
#pragma unroll
            for (int q = 0; q < 4; ++q) b.s[q] = __shfl_down_sync(0xffffffffu, a.s[q], off);
            unsigned long long cb = __shfl_down_sync(0xffffffffu, c, off);
            agg_merge(a, b, type);
            c += cb;
        }
        if (lane == 0) { warp_acc[warp] = a; warp_cnt[warp] = c; }
        // consumer-only barrier (the producer warp has exited)
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
        if (threadIdx.x == 0) {
            AggAcc r = warp_acc[0];
            unsigned long long rc = warp_cnt[0];
            for (int q = 1; q < CONSUMER_WARPS; ++q) { agg_merge(r, warp_acc[q], type); rc += warp_cnt[q]; }
            AggPartial o;
            o.count = rc; o.valid = rc != 0; o.pad = 0;
            if (type == 9) { o.sum = r.s[0]; o.err = as_f64(r.s[1]); o.mn = r.s[2]; o.mx = r.s[3]; }
            else { o.sum = r.s[0]; o.err = 0.0; o.mn = r.s[1]; o.mx = r.s[2]; }
            P.partials[(size_t)blockIdx.x * P.naggs + j] = o;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
    }
    }   // AGG
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

// ---- Bitset.Indexes for a whole batch of packs (reader.go:432-436: sel := bits.Indexes(hits) per pack).
// The packs' bitsets sit at ascending offsets of one device buffer; a thread owns one 32-bit word of that
// buffer, finds its pack by binary search over the offsets and ignores words that lie in the gaps.
__device__ __forceinline__ uint32_t sel_word(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t w,
                                             uint64_t total_words, uint32_t* local_word) {
    if (w >= total_words) return 0;
    const uint64_t byte = w * 4;
    uint32_t lo = 0, hi = npacks;   // last pack with bitset_off <= byte
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (packs[m].bitset_off <= byte) lo = m; else hi = m; }
    const uint64_t lw = (byte - packs[lo].bitset_off) >> 2;
    if (lw >= ((uint64_t)packs[lo].n + 31) >> 5) return 0;   // gap between two packs
    *local_word = (uint32_t)lw;
    return *reinterpret_cast<const uint32_t*>(bits + byte);   // tail bits past n are already zero
}

__global__ void select_counts_kernel(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t total_words,
                                     uint32_t* __restrict__ block_counts) {
    uint32_t lw;
    uint32_t c = __popc(sel_word(packs, npacks, bits, (uint64_t)blockIdx.x * 256 + threadIdx.x, total_words, &lw));
    __shared__ uint32_t ws[8];
    uint32_t s = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; ++q) t += ws[q]; block_counts[blockIdx.x] = t; }
}

__global__ void select_scatter_kernel(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t total_words,
                                      const uint32_t* __restrict__ block_offs, uint32_t* __restrict__ dst) {
    uint32_t lw = 0;
    uint32_t wv = sel_word(packs, npacks, bits, (uint64_t)blockIdx.x * 256 + threadIdx.x, total_words, &lw);
    uint32_t c = __popc(wv), incl = c;
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t q = 0; q < (threadIdx.x >> 5); ++q) woff += ws[q];
    uint32_t pos = block_offs[blockIdx.x] + woff + incl - c;
    const uint32_t base = lw << 5;   // row id relative to the pack
    while (wv) { uint32_t b = __ffs(wv) - 1; dst[pos++] = base + b; wv &= wv - 1; }
}

// NumberContainer.AppendTo(dst, sel) for a batch of packs (internal/encode/int_*.go AppendTo with a selection;
// query/result.go:196-264 copies the selected rows): one thread per selected row, decode at index.
__global__ void gather_kernel(const ColView* __restrict__ views, const unsigned long long* __restrict__ sel_off, uint32_t npacks,
                              const uint32_t* __restrict__ sel, uint64_t total, int elem_bytes, uint8_t* __restrict__ dst) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = npacks;   // last pack with sel_off <= i
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sel_off[m] <= i) lo = m; else hi = m; }
        uint64_t x = decode_value(views[lo], sel[i], nullptr, 0);
        switch (elem_bytes) {
        case 8: reinterpret_cast<uint64_t*>(dst)[i] = x; break;
        case 4: reinterpret_cast<uint32_t*>(dst)[i] = (uint32_t)x; break;
        case 2: reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)x; break;
        default: dst[i] = (uint8_t)x; break;
        }
    }
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

cudaError_t launch_scan(const ScanParams& P, int grid, size_t smem_bytes, bool simple, bool only32, int ctas_per_sm, cudaStream_t stream) {
    int variant = !simple ? (P.naggs ? 0 : 4) : (!only32 ? 1 : (ctas_per_sm >= 3 ? 3 : 2));
    void (*kern)(const ScanParams) = scan_kernel<false, false, 2, true>;
    if (variant == 1) kern = scan_kernel<true, false, 2, false>;
    if (variant == 2) kern = scan_kernel<true, true, 2, false>;
    if (variant == 3) kern = scan_kernel<true, true, 3, false>;
    if (variant == 4) kern = scan_kernel<false, false, 2, false>;
    // function attributes are per device and sticky: set them once per (device, variant)
    static bool configured[64][5] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev][variant]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_MAX_DYN_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev][variant] = true;
    }
    kern<<<grid, SCAN_THREADS, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_alpfix(const AlpFixJob* jobs, uint32_t njobs, uint32_t max_patches, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_patches == 0) return cudaSuccess;
    uint32_t gx = (max_patches + 255u) / 256u;
    if (gx > 148u) gx = 148u;
    alpfix_kernel<<<dim3(gx, njobs), 256, 0, stream>>>(jobs, out_base);
    return cudaGetLastError();
}

cudaError_t launch_runfill(const RunFillJob* jobs, uint32_t njobs, uint32_t max_runs, const uint64_t* set_vals, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_runs == 0) return cudaSuccess;
    uint32_t gx = (max_runs + 255u) / 256u;
    if (gx > 148u * 4u) gx = 148u * 4u;
    runfill_kernel<<<dim3(gx, njobs), 256, 0, stream>>>(jobs, set_vals, out_base);
    return cudaGetLastError();
}

cudaError_t launch_codeset(const CodesetJob* jobs, uint32_t njobs, uint32_t max_set, const uint64_t* set_vals, uint32_t* out, cudaStream_t stream) {
    if (njobs == 0 || max_set == 0) return cudaSuccess;
    uint32_t gx = (max_set + 127u) / 128u;
    if (gx > 32u) gx = 32u;
    codeset_kernel<<<dim3(gx, njobs), 128, 0, stream>>>(jobs, set_vals, out);
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
cudaError_t launch_select(const PackInfo* packs, uint32_t npacks, const uint8_t* bits, uint64_t total_words, uint32_t* block_tmp,
                          unsigned long long* total, uint32_t* dst, cudaStream_t stream) {
    uint32_t nblocks = (uint32_t)((total_words + 255) / 256);
    if (nblocks == 0) return cudaMemsetAsync(total, 0, 8, stream);
    select_counts_kernel<<<nblocks, 256, 0, stream>>>(packs, npacks, bits, total_words, block_tmp);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(block_tmp, nblocks, total);
    select_scatter_kernel<<<nblocks, 256, 0, stream>>>(packs, npacks, bits, total_words, block_tmp, dst);
    return cudaGetLastError();
}
cudaError_t launch_gather(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                          int elem_bytes, void* dst, cudaStream_t stream) {
    if (total == 0) return cudaSuccess;
    int grid = grid_for(total, 148 * 16);
    gather_kernel<<<grid ? grid : 1, 256, 0, stream>>>(views, sel_off, npacks, sel, total, elem_bytes, reinterpret_cast<uint8_t*>(dst));
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
