/*
 * ko_simd.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h): an AVX-512 (VBMI) version of the fused
 * bit-pack compare, used ONLY as the CPU baseline of bench.py.
 *
 * The reference's fused compare is generated scalar Go (internal/encode/bitpack/cmp_lt.go:1343+ and siblings,
 * 2.5–2.8 values/ns/core on an i9-12900K, bench.md:72-74); its SIMD lives in internal/cmp (raw vectors) only.
 * This kernel computes the SAME bitset words (ko_bitpack_cmp is the checker, tests/test_oracle_props.py) with
 * vector instructions, so that the GPU is compared against the best this host can do rather than against a
 * scalar port: 16 rows per 512-bit vector (vpermb gathers each row's bytes, vpsrlvd aligns, vpcmpud compares),
 * widths 1..25; 8 rows per vector with 64-bit lanes for widths 26..57; other widths fall back to the scalar port.
 */
#include "knox_oracle.h"
#include <immintrin.h>
#include <string.h>

int ko_simd_available(void) {
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi") &&
           __builtin_cpu_supports("avx512vl");
}

/* kind: 0 eq, 1 lt, 2 le, 3 between (v - a <= c2) — ko_bitpack_cmp's kinds */
__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl,popcnt")))
static void cmp_w32(int kind, int neg, const uint8_t* src, int w, uint32_t a, uint32_t c2, size_t groups, uint8_t* bits) {
    /* rows r = 0..15 of a vector: bytes src[(r*w)>>3 .. +3], shift (r*w)&7; the next 16 rows start 2w bytes further */
    uint8_t idx[64]; uint32_t sh[16];
    for (int r = 0; r < 16; r++) {
        int bit = r * w;
        for (int k = 0; k < 4; k++) idx[4 * r + k] = (uint8_t)((bit >> 3) + k);
        sh[r] = (uint32_t)(bit & 7);
    }
    const __m512i vidx = _mm512_loadu_si512(idx), vsh = _mm512_loadu_si512(sh);
    const __m512i vmask = _mm512_set1_epi32((int)(w >= 32 ? 0xffffffffu : ((1u << w) - 1u)));
    const __m512i va = _mm512_set1_epi32((int)a), vc = _mm512_set1_epi32((int)c2);
    const uint64_t flip = neg ? ~0ull : 0ull;
    for (size_t g = 0; g < groups; g++) {
        const uint8_t* p = src + g * (size_t)(8 * w);
        uint64_t word = 0;
        for (int q = 0; q < 4; q++) {
            __m512i raw = _mm512_loadu_si512(p + q * 2 * w);          /* (reads up to 63 bytes past the 16 rows: callers pad) */
            __m512i v = _mm512_and_si512(_mm512_srlv_epi32(_mm512_permutexvar_epi8(vidx, raw), vsh), vmask);
            __mmask16 m;
            switch (kind) {
            case 0: m = _mm512_cmpeq_epu32_mask(v, va); break;
            case 1: m = _mm512_cmplt_epu32_mask(v, va); break;
            case 2: m = _mm512_cmple_epu32_mask(v, va); break;
            default: m = _mm512_cmple_epu32_mask(_mm512_sub_epi32(v, va), vc); break;
            }
            word |= (uint64_t)m << (16 * q);
        }
        word ^= flip;
        memcpy(bits + g * 8, &word, 8);
    }
}

__attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vl,popcnt")))
static void cmp_w64(int kind, int neg, const uint8_t* src, int w, uint64_t a, uint64_t c2, size_t groups, uint8_t* bits) {
    uint8_t idx[64]; uint64_t sh[8];
    for (int r = 0; r < 8; r++) {
        int bit = r * w;
        for (int k = 0; k < 8; k++) idx[8 * r + k] = (uint8_t)((bit >> 3) + k);
        sh[r] = (uint64_t)(bit & 7);
    }
    const __m512i vidx = _mm512_loadu_si512(idx), vsh = _mm512_loadu_si512(sh);
    const __m512i vmask = _mm512_set1_epi64((long long)((1ull << w) - 1ull));
    const __m512i va = _mm512_set1_epi64((long long)a), vc = _mm512_set1_epi64((long long)c2);
    const uint64_t flip = neg ? ~0ull : 0ull;
    for (size_t g = 0; g < groups; g++) {
        const uint8_t* p = src + g * (size_t)(8 * w);
        uint64_t word = 0;
        for (int q = 0; q < 8; q++) {
            __m512i raw = _mm512_loadu_si512(p + q * w);
            __m512i v = _mm512_and_si512(_mm512_srlv_epi64(_mm512_permutexvar_epi8(vidx, raw), vsh), vmask);
            __mmask8 m;
            switch (kind) {
            case 0: m = _mm512_cmpeq_epu64_mask(v, va); break;
            case 1: m = _mm512_cmplt_epu64_mask(v, va); break;
            case 2: m = _mm512_cmple_epu64_mask(v, va); break;
            default: m = _mm512_cmple_epu64_mask(_mm512_sub_epi64(v, va), vc); break;
            }
            word |= (uint64_t)m << (8 * q);
        }
        word ^= flip;
        memcpy(bits + g * 8, &word, 8);
    }
}

/* same contract as ko_bitpack_cmp; `src` must be readable 64 bytes past the packed stream (the bench pads its packs).
 * Returns 1 when the vector path ran, 0 when the caller must use the scalar port (no AVX-512 VBMI, width 0 or > 57). */
int ko_bitpack_cmp_simd(int op, const uint64_t* src, int log2, uint64_t a, uint64_t b, size_t n, uint8_t* bits) {
    if (log2 < 1 || log2 > 57 || !ko_simd_available()) return 0;
    int kind, neg = 0;
    switch (op) {
    case KO_EQ: kind = 0; break;
    case KO_NE: kind = 0; neg = 1; break;
    case KO_LT: kind = 1; break;
    case KO_LE: kind = 2; break;
    case KO_GT: kind = 2; neg = 1; break;
    case KO_GE: kind = 1; neg = 1; break;
    case KO_RG: kind = 3; break;
    default: return 0;
    }
    const uint64_t mask = (1ull << log2) - 1ull, a0 = a;
    uint64_t c2 = b - a;
    /* operands beyond the field range: the 64-bit compares of the reference still decide them; fold them into the
     * field domain so that narrow lanes give the same answer */
    if (kind != 3 && a > mask) { if (kind == 0) { memset(bits, neg ? 0xff : 0, (n / 64) * 8); goto tail; } a = mask; if (kind == 1) kind = 2; }
    if (kind == 3 && (a > mask || c2 > mask)) return 0;   /* rare shapes: scalar port */
    {
        const size_t groups = n / 64;
        if (log2 <= 25) cmp_w32(kind, neg, (const uint8_t*)src, log2, (uint32_t)a, (uint32_t)c2, groups, bits);
        else cmp_w64(kind, neg, (const uint8_t*)src, log2, a, c2, groups, bits);
    }
tail:
    if (n & 63) {   /* the tail like the scalar port: decoded row by row */
        size_t k = n & ~(size_t)63, rem = n & 63;
        const uint64_t* p = src + (k / 64) * (size_t)log2;
        size_t tail_words = ko_bitpack_size(log2, rem) / 8;
        for (size_t i = 0; i < rem; i++) {
            uint64_t v = ko_bitpack_value(p, tail_words, i, log2, 0);
            int r;
            switch (op) {
            case KO_EQ: r = v == a0; break; case KO_NE: r = v != a0; break;
            case KO_LT: r = v < a0; break; case KO_LE: r = v <= a0; break;
            case KO_GT: r = v > a0; break; case KO_GE: r = v >= a0; break;
            default: r = (v - a0) <= (b - a0); break;
            }
            if (r) bits[(k + i) >> 3] |= (uint8_t)(1u << ((k + i) & 7));
        }
    }
    return 1;
}
