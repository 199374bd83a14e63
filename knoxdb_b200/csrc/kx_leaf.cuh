// kx_leaf.cuh — device building blocks shared by the scan kernels (kx_scan.cu: single-leaf kernels, kx_general.cu:
// multi-leaf filter + fused reduce): PTX helpers for the TMA ring, the per-leaf evaluation functions (one pass of
// 32 groups of 32 rows per call, "bitset word per lane"), and the aggregate accumulators.
#pragma once
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
    // try_wait suspends the thread in hardware until the phase completes or the time hint (ns) runs out: the loop spins
    // a handful of times per wait instead of burning issue slots the other warps of the SM need
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra KX_DONE;\n"
        "bra KX_WAIT;\n"
        "KX_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity), "r"(20000u)
        : "memory");
}
// the same on a precomputed 32-bit shared-memory address (no generic → shared conversion per call)
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "KX_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n"
        "@P1 bra KX_DONE;\n"
        "bra KX_WAIT;\n"
        "KX_DONE:\n"
        "}\n" ::"r"(bar_addr), "r"(parity), "r"(20000u)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
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
// Pipe balance of the row code (KX_IMAD_ROWS, chosen per translation unit).  IADD3 / LOP3 / SHF / ISETP issue on the ALU pipe (one
// warp instruction per 2 cycles and SMSP), IMAD on the FMA pipe (same rate, in parallel).
//   0: shift, subtract, compare, predicated OR — 3-4 ALU operations per row (the single-leaf kernel: measured equal or
//      better there, its narrow widths are bound by issue slots, not by a pipe);
//   1: the shift AND the subtract are ONE multiply-add — field << k == word * 2^k (mod 2^32), the multiplier read from a
//      constant bank so that ptxas cannot strength-reduce it back to a shift — and every other row's bit is accumulated with
//      a predicated multiply-add: 1.5 ALU + 1.5 FMA operations per row;
//   2: multiply-add as in 1, then a carry chain: acc = 2 acc + carry(lim - t) is one IADD3 with carry-out and one
//      add-with-carry (ptxas issues it as IMAD.X): 1 ALU + 2 FMA per row, no ISETP, no predicated instruction (the
//      warp-autonomous kernel: 3-5 % faster on the config-3 cases, profiles/r2_tune_warp.txt §6).
// Fields that straddle two words keep the funnel shift in every form.
#ifndef KX_IMAD_ROWS
#define KX_IMAD_ROWS 0
#endif
static __constant__ uint32_t kx_pow2[32] = {   // 2^k from a constant bank: a multiply ptxas cannot turn back into a shift

    0x1u, 0x2u, 0x4u, 0x8u, 0x10u, 0x20u, 0x40u, 0x80u, 0x100u, 0x200u, 0x400u, 0x800u, 0x1000u, 0x2000u, 0x4000u, 0x8000u,
    0x10000u, 0x20000u, 0x40000u, 0x80000u, 0x100000u, 0x200000u, 0x400000u, 0x800000u, 0x1000000u, 0x2000000u, 0x4000000u, 0x8000000u,
    0x10000000u, 0x20000000u, 0x40000000u, 0x80000000u};

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
    uint32_t wq[4] = {0, 0, 0, 0};   // four independent accumulators: short dependency chains
#if KX_IMAD_ROWS == 2
    // carry-chain form: accumulator q collects rows 8q … 8q+7, highest row first: acc = 2 acc + carry(lim - t), so the
    // compare and the bit insert are one IADD3 with carry-out and one add-with-carry — no ISETP, no predicated instruction.
    // (sub.cc leaves the hardware carry of lim + ~t + 1: set exactly when t <= lim, i.e. when the row matches.)
    const uint32_t na_top = 0u - a_top, one = kx_pow2[0];
#pragma unroll
    for (int jj = 0; jj < 32; ++jj) {
        const int j = (jj & 3) * 8 + 7 - (jj >> 2);   // interleave the four chains
        const int bit = j * W, wi = bit >> 5, sh = bit & 31;
        uint32_t t;
        if (sh + W <= 32) {
            const int k = 32 - sh - W;
            if (SUB) t = x[wi] * kx_pow2[k] + na_top;
            else t = k ? x[wi] * kx_pow2[k] : x[wi];
        } else {
            t = __funnelshift_l(x[wi], x[wi + 1], 64 - sh - W);
            if (SUB) t = t * one + na_top;
        }
        uint32_t dummy;
        asm("sub.cc.u32 %1, %2, %3;\n\taddc.u32 %0, %0, %0;" : "+r"(wq[j >> 3]), "=r"(dummy) : "r"(lim), "r"(t));
    }
    uint32_t word = (wq[0] | (wq[1] << 8)) | ((wq[2] << 16) | (wq[3] << 24));
    if constexpr (W == 8 || W == 16 || W == 24 || W == 32) word = __funnelshift_l(word, word, rot_rows);
    return word;
#else
#if KX_IMAD_ROWS
    const uint32_t na_top = 0u - a_top, one = kx_pow2[0];
#endif
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * W, wi = bit >> 5, sh = bit & 31;
        uint32_t t;
#if KX_IMAD_ROWS
        if (sh + W <= 32) {
            const int k = 32 - sh - W;
            if (SUB) t = x[wi] * kx_pow2[k] + na_top;          // IMAD: shift and subtract in one
            else t = k ? x[wi] * kx_pow2[k] : x[wi];
        } else {
            t = __funnelshift_l(x[wi], x[wi + 1], 64 - sh - W);
            if (SUB) t = t * one + na_top;                      // the subtract alone, still off the ALU pipe
        }
        if (j & 1) {
            asm("{\n.reg .pred p;\nsetp.le.u32 p, %1, %2;\n@p mad.lo.u32 %0, %3, %4, %0;\n}" : "+r"(wq[j & 3]) : "r"(t), "r"(lim), "r"(one), "r"(1u << j));
        } else {
            if (t <= lim) wq[j & 3] |= (1u << j);
        }
#else
        if (sh + W <= 32) t = x[wi] << (32 - sh - W);
        else t = __funnelshift_l(x[wi], x[wi + 1], 64 - sh - W);
        if (SUB) t -= a_top;
        if (t <= lim) wq[j & 3] |= (1u << j);
#endif
    }
    uint32_t word = (wq[0] | wq[1]) | (wq[2] | wq[3]);
    if constexpr (W == 8 || W == 16 || W == 24 || W == 32) word = __funnelshift_l(word, word, rot_rows);
    return word;
#endif
}

template <bool SUB>
static __device__ __noinline__ uint32_t leaf_b32_dispatch(const uint32_t* __restrict__ seg, uint32_t lane, uint32_t w, uint32_t a_top, uint32_t lim) {
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

// every pass of one tile (single-leaf kernel): the width dispatch sits OUTSIDE the pass loop, so a tile pays for it once and
// the unrolled row code of one width runs `passes` times back to back; emit(pass, word) consumes the lane's bitset word.
// seg0 = the lane's group of pass 0 (pass p: 32 groups = 32 W words further); lanes that own no group (`own` false) emit 0.
template <bool SUB, class F>
__device__ __forceinline__ void leaf_b32_passes(const uint32_t* __restrict__ seg0, uint32_t lane, uint32_t w, uint32_t a_top, uint32_t lim,
                                                uint32_t passes, bool own, F emit) {
    switch (w) {
#define KX_CASE(W)                                                                                                          \
    case W:                                                                                                                 \
        _Pragma("unroll 1") for (uint32_t p = 0; p < passes; ++p)                                                           \
            emit(p, own ? leaf_b32<W, SUB>(seg0 + (size_t)p * (32u * W), lane, a_top, lim) : 0u);                           \
        break;
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
        KX_CASE(17) KX_CASE(18) KX_CASE(19) KX_CASE(20) KX_CASE(21) KX_CASE(22) KX_CASE(23) KX_CASE(24)
        KX_CASE(25) KX_CASE(26) KX_CASE(27) KX_CASE(28) KX_CASE(29) KX_CASE(30) KX_CASE(31) KX_CASE(32)
#undef KX_CASE
    }
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
static __device__ __noinline__ uint32_t leaf_b64_dispatch(const uint32_t* __restrict__ seg, uint32_t w, uint64_t a_top, uint64_t lim) {
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
// GBM: the bitmap is read from global memory (L1-resident: <= 8 KB per pack and leaf) instead of a shared-memory copy
template <int W, bool GBM>
__device__ __forceinline__ uint32_t leaf_code32(const uint32_t* __restrict__ seg, uint32_t code_base, const uint32_t* __restrict__ bm) {
    __builtin_assume(__isShared(seg));
    if constexpr (!GBM) __builtin_assume(__isShared(bm));   // the pack's code bitmap is cached in shared memory (one LDS per row)
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
    // The bitmap is indexed by the FIELD (codeset_kernel subtracts the code stream's base), so a row needs no mask and no add:
    // the word index comes from the top-aligned field (one IMAD on the FMA pipe + one shift), the bit position is the low
    // five bits of the right-aligned field (funnel shifts ignore the garbage above): 4 ALU + 2 FMA-pipe operations per
    // row instead of 8-9 ALU operations.
    uint32_t word = 0;
    if constexpr (GBM) {
        // bitmap read in place through L1 (warp-autonomous kernel): the plain form — mask, index, load, two funnel shifts —
        // measured 7 % faster there than the multiply-add form below (the longer address chain in front of a global load)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const int bit = j * W, wi = bit >> 5, sh = bit & 31;
            const uint32_t f = (sh + W <= 32) ? (x[wi] >> sh) : __funnelshift_r(x[wi], x[wi + 1], sh);
            const uint32_t code = f & ((1u << W) - 1u);
            const uint32_t wv = __ldg(bm + (code >> 5));
            word = __funnelshift_r(word, __funnelshift_r(wv, 0u, code), 1);
        }
        return word;
    }
    const uint32_t w0 = (W <= 5) ? (GBM ? __ldg(bm) : bm[0]) : 0u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const int bit = j * W, wi = bit >> 5, sh = bit & 31;
        uint32_t f, wv;
        if (sh + W <= 32) f = sh ? (x[wi] >> sh) : x[wi];
        else f = __funnelshift_r(x[wi], x[wi + 1], sh);
        if constexpr (W < 5) f &= (1u << W) - 1u;   // (narrower than the bit index: the garbage above the field must go)
        if constexpr (W <= 5) {
            wv = w0;
        } else {
            const uint32_t t = (sh + W <= 32) ? x[wi] * kx_pow2[32 - sh - W] : f * kx_pow2[32 - W];   // field << (32 - W)
            const uint32_t idx = t >> (32 - W + 5);
            wv = GBM ? __ldg(bm + idx) : bm[idx];
        }
        word = __funnelshift_r(word, __funnelshift_r(wv, 0u, f), 1);   // shift bit (field & 31) of wv in from the top
    }
    return word;   // after 32 steps row j sits at bit j
}

template <bool GBM>
static __device__ __noinline__ uint32_t leaf_code32_dispatch(const uint32_t* __restrict__ seg, uint32_t w, uint32_t code_base, const uint32_t* __restrict__ bm) {
    switch (w) {
#define KX_CASE(W) case W: return leaf_code32<W, GBM>(seg, code_base, bm);
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
#undef KX_CASE
    }
    return 0;
}

// every pass of one tile for a dictionary IN / NOT IN leaf (single-leaf kernel, bitmap in shared memory): one width dispatch
// per tile, as leaf_b32_passes
template <class F>
__device__ __forceinline__ void leaf_code32_passes(const uint32_t* __restrict__ seg0, uint32_t w, const uint32_t* __restrict__ bm, uint32_t passes, bool own, F emit) {
    switch (w) {
#define KX_CASE(W)                                                                                                          \
    case W:                                                                                                                 \
        _Pragma("unroll 1") for (uint32_t p = 0; p < passes; ++p)                                                           \
            emit(p, own ? leaf_code32<W, false>(seg0 + (size_t)p * (32u * W), 0u, bm) : 0u);                                \
        break;
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
#undef KX_CASE
    }
}

template <bool GBM>
__device__ __forceinline__ uint32_t leaf_codeset(const uint32_t* __restrict__ sw, uint32_t w, uint32_t g0, uint32_t Rp, uint32_t lane,
                                                 uint32_t code_base, const uint32_t* __restrict__ bm) {
    if (lane >= Rp) return 0;
    return leaf_code32_dispatch<GBM>(sw + (size_t)(g0 + lane) * w, w, code_base, bm);   // dictionary codes are uint16: w <= 16
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
        bool p = (L.mode == LM_SET) ? set_has(sets + L.a, (uint32_t)L.d, val) : (L.mode == LM_RUNRANGE) ? ((uint64_t)k - L.a) <= L.d : ((val ^ L.wm) - L.a) <= L.d;
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


// ------------------------------------------------------------------------------ one leaf, one pass
// first pack whose tile range contains tile t (packs with zero tiles are skipped)
__device__ __forceinline__ uint32_t pack_of_tile(const PackInfo* __restrict__ packs, uint32_t npacks, uint32_t t) {
    uint32_t lo = 0, hi = npacks;   // last pack with tile0 <= t
    while (hi - lo > 1) {
        uint32_t m = (lo + hi) >> 1;
        if (packs[m].tile0 <= t) lo = m; else hi = m;
    }
    return lo;
}

struct LeafEnv {
    const ScanParams& P;
    const uint32_t* code_smem;   // shared memory behind the ring: code bitmaps, prefilters, small hash tables
    uint32_t nrows;              // rows of the pack
    uint32_t pack_row0;          // first pack row of the tile
};

// One leaf (index `li` of the program) for one pass of 32 groups: sw = the leaf's staged stream (or nullptr), g0 = first
// group of the pass within the tile, wr = first pack row of this lane's word.  `keep`: rows (bits of this lane's word)
// whose result matters — the other operand of an enclosing AND / OR.  Returns the word with the leaf's `neg` applied.
// GBM: dictionary-code bitmaps are read in place (global memory) instead of from the CTA's shared-memory copy.
template <bool GBM = false>
__device__ __forceinline__ uint32_t eval_leaf(const LeafEnv& E, const PackLeaf& lf, uint32_t li, const uint32_t* sw, uint32_t g0, uint32_t Rp,
                                              uint32_t lane, uint64_t wr, uint32_t keep) {
    const ScanParams& P = E.P;
    uint32_t word;
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
    case LM_BITS: __builtin_assume(__isShared(sw)); word = lane < Rp ? sw[g0 + lane] : 0u; break;   // precomputed leaf bitset (pre-pass kernels, row masks)
    case LM_CODESET:
        word = leaf_codeset<GBM>(sw, lf.width, g0, Rp, lane, (uint32_t)lf.wm, GBM ? P.code_bits + lf.a : E.code_smem + P.code_smem_off[li]);
        break;
    case LM_HASHSET: {   // a = For of the block, fop = its element type
        const uint32_t to = P.hs_tab_smem_off[li];
        const ulonglong2* tab = to != 0xffffffffu ? reinterpret_cast<const ulonglong2*>(E.code_smem + to)
                                                  : reinterpret_cast<const ulonglong2*>(P.set_tabs + P.tab_off[li]);
        word = leaf_hashset(sw, lf.width, lf.fop, lf.a, g0, Rp, lane, E.code_smem + P.hs_smem_off[li], P.pre_log2[li], tab, P.tab_log2[li], keep);
        break;
    }
    default: {
        const ColView& v = P.views[lf.view];
        if (v.kind == CK_RUNEND) word = leaf_runend(lf, v, (uint32_t)wr, E.nrows, lane < Rp, P.set_vals);
        else word = leaf_generic(lf, v, lf.data ? sw : nullptr, E.pack_row0, g0, Rp, lane, E.nrows, P.set_vals);
        break;
    }
    }
    return lf.neg ? ~word : word;
}

// ------------------------------------------------------------------------------ one leaf, all words of a warp's tile
// (warp-autonomous kernel, kx_warp.cu).  A tile is wd x 1024 rows; word j of lane l covers rows [32 (32 j + l), + 32).
// The loop over the words sits INSIDE the per-width / per-kind code, so the leaf is dispatched once per tile, its
// loop-invariant operands are prepared once, and the unrolled row code of one width runs wd times back to back.
struct WordsIO {
    uint32_t* dst;          // lane-private column (+ 32 words per step) the results go to
    const uint32_t* opnd;   // lane-private column of the other operand of the enclosing AND / OR (nullptr: none)
    uint32_t keep_inv;      // 0: rows whose operand bit is 1 matter (AND); ~0: rows whose operand bit is 0 (OR)
    uint32_t comb;          // 0: dst = word (0 when the whole word was skipped), 1: dst = opnd & word, 2: dst = opnd | word
    uint32_t flip;          // XORed onto the leaf's word (NE / GT / NIN … and the float-level NOT)
    uint32_t wd;
};

template <class F>
__device__ __forceinline__ void words_loop(const WordsIO& io, F f) {
#pragma unroll 1
    for (uint32_t j = 0; j < io.wd; ++j) {
        const uint32_t opnd = io.opnd ? io.opnd[j * 32u] : 0u;
        const uint32_t keep = io.opnd ? (opnd ^ io.keep_inv) : 0xffffffffu;
        if (__any_sync(0xffffffffu, keep != 0u)) {   // MatchAnd / MatchOr early-out per warp and word
            const uint32_t word = f(j, keep) ^ io.flip;
            io.dst[j * 32u] = io.comb == 0u ? word : (io.comb == 1u ? (opnd & word) : (opnd | word));
        } else if (io.comb == 0u) {
            io.dst[j * 32u] = 0u;
        }
    }
}

template <bool SUB>
static __device__ __noinline__ void leaf_b32_words(const uint32_t* __restrict__ seg0, uint32_t lane, uint32_t w, uint32_t a_top, uint32_t lim, WordsIO io) {
    switch (w) {
#define KX_CASE(W) case W: words_loop(io, [&](uint32_t j, uint32_t) { return leaf_b32<W, SUB>(seg0 + j * (32u * W), lane, a_top, lim); }); break;
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
        KX_CASE(17) KX_CASE(18) KX_CASE(19) KX_CASE(20) KX_CASE(21) KX_CASE(22) KX_CASE(23) KX_CASE(24)
        KX_CASE(25) KX_CASE(26) KX_CASE(27) KX_CASE(28) KX_CASE(29) KX_CASE(30) KX_CASE(31) KX_CASE(32)
#undef KX_CASE
    }
}

template <bool SUB>
static __device__ __noinline__ void leaf_b64_words(const uint32_t* __restrict__ seg0, uint32_t w, uint64_t a_top, uint64_t lim, WordsIO io) {
    switch (w) {
#define KX_CASE(W) case W: words_loop(io, [&](uint32_t j, uint32_t) { return leaf_b64<W, SUB>(seg0 + j * (32u * W), a_top, lim); }); break;
        KX_CASE(33) KX_CASE(34) KX_CASE(35) KX_CASE(36) KX_CASE(37) KX_CASE(38) KX_CASE(39) KX_CASE(40)
        KX_CASE(41) KX_CASE(42) KX_CASE(43) KX_CASE(44) KX_CASE(45) KX_CASE(46) KX_CASE(47) KX_CASE(48)
        KX_CASE(49) KX_CASE(50) KX_CASE(51) KX_CASE(52) KX_CASE(53) KX_CASE(54) KX_CASE(55) KX_CASE(56)
        KX_CASE(57) KX_CASE(58) KX_CASE(59) KX_CASE(60) KX_CASE(61) KX_CASE(62) KX_CASE(63)
#undef KX_CASE
    }
}

template <bool GBM>
static __device__ __noinline__ void leaf_code32_words(const uint32_t* __restrict__ seg0, uint32_t w, uint32_t code_base, const uint32_t* __restrict__ bm, WordsIO io) {
    switch (w) {
#define KX_CASE(W) case W: words_loop(io, [&](uint32_t j, uint32_t) { return leaf_code32<W, GBM>(seg0 + j * (32u * W), code_base, bm); }); break;
        KX_CASE(1) KX_CASE(2) KX_CASE(3) KX_CASE(4) KX_CASE(5) KX_CASE(6) KX_CASE(7) KX_CASE(8)
        KX_CASE(9) KX_CASE(10) KX_CASE(11) KX_CASE(12) KX_CASE(13) KX_CASE(14) KX_CASE(15) KX_CASE(16)
#undef KX_CASE
    }
}

// sw = the leaf's staged slice of the tile (or nullptr), wr0 = first pack row of this lane's word 0; io.flip carries the
// inversions the caller wants on top of the leaf's own `neg`
template <bool GBM>
__device__ __forceinline__ void eval_leaf_words(const LeafEnv& E, const PackLeaf& lf, uint32_t li, const uint32_t* sw, uint32_t lane, uint32_t wr0, WordsIO io) {
    const ScanParams& P = E.P;
    if (lf.neg) io.flip = ~io.flip;
    const uint32_t w = lf.width;
    switch (lf.mode) {
    case LM_NONE: words_loop(io, [](uint32_t, uint32_t) { return 0u; }); break;
    case LM_ALL: words_loop(io, [](uint32_t, uint32_t) { return 0xffffffffu; }); break;
    case LM_RANGE32: {
        const uint32_t k = 32u - w, a = (uint32_t)lf.a, d = (uint32_t)lf.d;
        const uint32_t a_top = a << k, lim = (d << k) | ((1u << k) - 1u);   // k == 0: a, d
        const uint32_t* seg0 = sw + (size_t)lane * w;
        if (a) leaf_b32_words<true>(seg0, lane, w, a_top, lim, io);
        else leaf_b32_words<false>(seg0, lane, w, 0u, lim, io);
        break;
    }
    case LM_RANGE64: {
        if (w > 32u && w < 64u) {
            const uint32_t k = 64u - w;
            const uint64_t a_top = lf.a << k, lim = (lf.d << k) | ((1ull << k) - 1ull);
            const uint32_t* seg0 = sw + (size_t)lane * w;
            if (lf.a) leaf_b64_words<true>(seg0, w, a_top, lim, io);
            else leaf_b64_words<false>(seg0, w, 0ull, lim, io);
        } else {
            const uint64_t a = lf.a, d = lf.d, wm = lf.wm;
            words_loop(io, [&](uint32_t j, uint32_t) { return leaf_range64(sw, w, j * 32u, 32u, lane, a, d, wm); });
        }
        break;
    }
    case LM_FLOAT: {
        const uint32_t fop = lf.fop;
        const uint64_t a = lf.a, d = lf.d;
        words_loop(io, [&](uint32_t j, uint32_t) { return leaf_float(sw, w, j * 32u, 32u, lane, fop, a, d); });
        break;
    }
    case LM_ROWRANGE: {
        // rows [a, a+d] of the pack → bits of this lane's words
        const uint64_t lo = lf.a, hi = lf.a + lf.d;
        words_loop(io, [&](uint32_t j, uint32_t) {
            const uint64_t wr = (uint64_t)wr0 + j * 1024u;
            uint32_t word = 0;
            if (hi >= wr && lo < wr + 32u) {
                const uint32_t b0 = lo > wr ? (uint32_t)(lo - wr) : 0u;
                const uint32_t b1 = hi < wr + 31u ? (uint32_t)(hi - wr) : 31u;
                word = (0xffffffffu >> (31u - b1)) & (0xffffffffu << b0);
            }
            return word;
        });
        break;
    }
    case LM_BITS: {   // precomputed leaf bitset (pre-pass kernels, row masks)
        __builtin_assume(__isShared(sw));
        words_loop(io, [&](uint32_t j, uint32_t) { return sw[j * 32u + lane]; });
        break;
    }
    case LM_CODESET:
        leaf_code32_words<GBM>(sw + (size_t)lane * w, w, (uint32_t)lf.wm, GBM ? P.code_bits + lf.a : E.code_smem + P.code_smem_off[li], io);
        break;
    case LM_HASHSET: {   // a = For of the block, fop = its element type
        const uint32_t to = P.hs_tab_smem_off[li];
        const ulonglong2* tab = to != 0xffffffffu ? reinterpret_cast<const ulonglong2*>(E.code_smem + to)
                                                  : reinterpret_cast<const ulonglong2*>(P.set_tabs + P.tab_off[li]);
        const uint32_t* pre = E.code_smem + P.hs_smem_off[li];
        const uint32_t pre_log2 = P.pre_log2[li], tab_log2 = P.tab_log2[li];
        const int type = lf.fop;
        const uint64_t base = lf.a;
        words_loop(io, [&](uint32_t j, uint32_t keep) { return leaf_hashset(sw, w, type, base, j * 32u, 32u, lane, pre, pre_log2, tab, tab_log2, keep); });
        break;
    }
    default: {
        const ColView& v = P.views[lf.view];
        if (v.kind == CK_RUNEND) words_loop(io, [&](uint32_t j, uint32_t) { return leaf_runend(lf, v, wr0 + j * 1024u, E.nrows, true, P.set_vals); });
        else words_loop(io, [&](uint32_t j, uint32_t) { return leaf_generic(lf, v, lf.data ? sw : nullptr, E.pack_row0, j * 32u, 32u, lane, E.nrows, P.set_vals); });
        break;
    }
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
    if (type == 9 || type == 10) { A.s[0] = 0; A.s[1] = 0; A.s[2] = 0x7ff0000000000000ull; A.s[3] = 0xfff0000000000000ull; }
    else { A.s[0] = 0; A.s[1] = ~0ull; A.s[2] = 0; A.s[3] = 0; }
    return A;
}

__device__ __forceinline__ void agg_add(AggAcc& A, int type, uint64_t bits) {
    if (type == 10) { bits = as_u64((double)__uint_as_float((uint32_t)bits)); type = 9; }   // float32 columns accumulate in float64 (exact widening)
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
    if (type == 9 || type == 10) {
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

// ------------------------------------------------------------------------------ reduce helpers shared by the fused kernels
template <bool F64>
__device__ __forceinline__ void acc_raw64(AggAcc& A, uint64_t raw, uint64_t base, uint64_t flip) {
    if (F64) {
        double x = as_f64(raw), sum = as_f64(A.s[0]), err = as_f64(A.s[1]);
        double t = sum + x;
        err += (fabs(sum) >= fabs(x)) ? ((sum - t) + x) : ((x - t) + sum);
        A.s[0] = as_u64(t); A.s[1] = as_u64(err);
        if (x < as_f64(A.s[2])) A.s[2] = raw;
        if (x > as_f64(A.s[3])) A.s[3] = raw;
    } else {
        uint64_t v = raw + base, k = v ^ flip;
        A.s[0] += v;
        if (k < A.s[1]) A.s[1] = k;
        if (k > A.s[2]) A.s[2] = k;
    }
}

// the rows of one reduce chunk this lane owns, as a bit mask rotated by `rot`: bit s ↔ row (s + rot) mod G of the lane's
// G-row range
__device__ __forceinline__ uint32_t lane_rows(uint32_t word, uint32_t sub, uint32_t G, uint32_t rot) {
    if (G == 32u) return __funnelshift_r(word, word, rot);
    const uint32_t m = (1u << G) - 1u, bits = (word >> (sub * G)) & m;
    return ((bits >> rot) | (bits << (G - rot))) & m;
}

// raw 64-bit value column, staged chunk in shared memory: positional walk (all lanes at the same step: conflict free)
template <bool F64>
__device__ __forceinline__ void reduce_staged_raw64(AggAcc& A, const unsigned long long* __restrict__ vp, uint32_t r, uint32_t G, uint32_t rot,
                                                    uint64_t base, uint64_t flip) {
    __builtin_assume(__isShared(vp));
    if (!__any_sync(0xffffffffu, r != 0u)) return;
#pragma unroll 4
    for (uint32_t s = 0; s < G; ++s) {
        if ((r >> s) & 1u) acc_raw64<F64>(A, vp[(s + rot) & (G - 1u)], base, flip);
    }
}

// raw 64-bit value column read on demand from global memory: only matching rows, four loads in flight per lane
template <bool F64>
__device__ __forceinline__ void reduce_global_raw64(AggAcc& A, const unsigned long long* __restrict__ gp, uint32_t r, uint32_t G, uint32_t rot,
                                                    uint64_t base, uint64_t flip) {
    while (r) {
        uint64_t val[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ok[u] = r != 0u;
            const uint32_t s = ok[u] ? (uint32_t)__ffs((int)r) - 1u : 0u;
            r &= r - 1u;
            val[u] = ok[u] ? __ldg(gp + ((s + rot) & (G - 1u))) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (ok[u]) acc_raw64<F64>(A, val[u], base, flip);
    }
}

// any other value column layout (bit-packed, dictionary, affine, run-end, narrow types, ALP): decode per row.
// row0 = pack row of the lane's range, srow0 = the same row relative to the staged slice (`staged` may be nullptr)
__device__ __forceinline__ void reduce_generic(AggAcc& A, const ColView& v, int type, uint32_t row0, const uint32_t* staged, uint32_t srow0, uint32_t r,
                                               uint32_t G, uint32_t rot) {
    while (r) {
        const uint32_t s = (uint32_t)__ffs((int)r) - 1u;
        r &= r - 1u;
        const uint32_t b = (s + rot) & (G - 1u);
        agg_add(A, type, decode_value(v, row0 + b, staged, srow0 + b));
    }
}

// r ⊕= p for two partial aggregates (p follows r in CTA order); invalid partials (no match) are neutral
__device__ __forceinline__ void partial_merge(AggPartial& r, const AggPartial& p, int type) {
    if (!p.valid) return;
    if (!r.valid) { r = p; return; }
    r.count += p.count;
    if (type == 9 || type == 10) {
        double s = as_f64(r.sum), e = r.err;
        fsum_merge(s, e, as_f64(p.sum), p.err);
        r.sum = as_u64(s); r.err = e;
        if (as_f64(p.mn) < as_f64(r.mn)) r.mn = p.mn;
        if (as_f64(p.mx) > as_f64(r.mx)) r.mx = p.mx;
    } else {
        r.sum += p.sum;
        if (p.mn < r.mn) r.mn = p.mn;
        if (p.mx > r.mx) r.mx = p.mx;
    }
}


}  // namespace kx
