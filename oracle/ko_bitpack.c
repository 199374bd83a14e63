/*
 * ko_bitpack.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h):
 * horizontal bit-packing with 64-bit code words + fused compare on packed words.
 */
#include "knox_oracle.h"
#include <string.h>

/* internal/encode/bitpack/bitpack.go:9-11: (log2*n + 63) &^ 63 / 8 */
size_t ko_bitpack_size(int log2, size_t n) {
    return (((size_t)log2 * n + 63) & ~(size_t)63) / 8;
}

/* internal/types/number.go:163-169: bits.Len64(max - min) (signed: in int64 space) */
int ko_log2range(int type, uint64_t minv, uint64_t maxv) {
    (void)type; /* values arrive sign-/zero-extended to 64 bit, so the u64 difference is the range */
    uint64_t d = maxv - minv;
    return d ? 64 - __builtin_clzll(d) : 0;
}

static inline uint64_t wmask(int log2) { return log2 >= 64 ? ~0ull : ((1ull << log2) - 1); }

/* internal/encode/bitpack/encode.go:216-246 (generic `encode`): one contiguous LSB-first
 * bit string in 64-bit LE words, value i at bits [i*log2, (i+1)*log2), last word zero
 * padded.  The unrolled pack_u64[w] kernels (uint64.go) produce the same layout.
 * Returns bytes written (multiple of 8). */
size_t ko_bitpack_encode(uint64_t* dst, const uint64_t* vals, size_t n, int log2, uint64_t minv) {
    if (log2 == 0) return 0;
    uint64_t word = 0, mask = wmask(log2);
    int offset = 0;
    size_t k = 0;
    for (size_t i = 0; i < n; i++) {
        uint64_t v = (vals[i] - minv) & mask;
        word |= v << offset;
        offset += log2;
        if (offset >= 64) {
            dst[k++] = word;
            offset -= 64;
            word = offset > 0 ? v >> (log2 - offset) : 0;
        }
    }
    if (offset > 0) dst[k++] = word;
    return k * 8;
}

/* internal/encode/bitpack/decode.go:56-74 (DecodeValue) */
uint64_t ko_bitpack_value(const uint64_t* src, size_t nwords, size_t i, int log2, uint64_t minv) {
    if (log2 == 0) return minv;
    size_t idx = i * (size_t)log2;
    size_t pos = idx >> 6;
    int shift = (int)(idx & 63);
    uint64_t word = src[pos] >> shift;
    int diff = 64 - shift;
    if (diff < log2 && pos + 1 < nwords) word |= src[pos + 1] << diff;
    return (word & wmask(log2)) + minv;
}

/* internal/encode/bitpack/decode.go:131-208 (Decode / decode): out[i] = field + minv */
void ko_bitpack_decode(uint64_t* dst, const uint64_t* src, size_t n, int log2, uint64_t minv) {
    size_t nwords = ko_bitpack_size(log2, n) / 8;
    for (size_t i = 0; i < n; i++) dst[i] = ko_bitpack_value(src, nwords, i, log2, minv);
}

/* one 64-row group → one bitset word; follows the generated cmp_<w>_{eq,lt,le,bw}
 * (cmp_eq.go:25-30 w=0 special case: eq → all ones iff val == 0; cmp_bw.go:25-30:
 * bw → all ones iff val1 == 0; lt w=0: 0 < val; le w=0: always). */
static inline __attribute__((always_inline)) uint64_t group_cmp(int kind /*0 eq,1 lt,2 le,3 bw*/, const uint64_t* p, int log2, uint64_t a, uint64_t b) {
    uint64_t out = 0, mask = wmask(log2), c2 = b - a;
    if (log2 == 0 && kind == 3) return a == 0 ? ~0ull : 0; /* cmp_bw.go:25-30 (differs from the formula only for a > b) */
    for (int i = 0; i < 64; i++) {
        uint64_t v;
        if (log2 == 0) v = 0;
        else {
            size_t bit = (size_t)i * log2; size_t pos = bit >> 6; int sh = (int)(bit & 63);
            v = p[pos] >> sh;
            if (64 - sh < log2) v |= p[pos + 1] << (64 - sh);
            v &= mask;
        }
        int r;
        switch (kind) {
        case 0: r = v == a; break;
        case 1: r = v < a; break;
        case 2: r = v <= a; break;
        default: r = (v - a) <= c2; break;
        }
        out |= (uint64_t)r << i;
    }
    return out;
}

static void put64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

/* internal/encode/bitpack/cmp.go:20-130.  NE/GT/GE are the bitwise NOT of EQ/LE/LT
 * (cmp.go:24-46); full 64-row groups OVERWRITE bitset words, the tail (<64 rows) is
 * decoded and its bits are ORed in one by one (cmp.go:55-86).  a/b are already in the
 * min-FOR domain (the container applies the `val < For` pre-checks). */
void ko_bitpack_cmp(int op, const uint64_t* src, int log2, uint64_t a, uint64_t b, size_t n, uint8_t* bits) {
    int kind, neg = 0;
    switch (op) {
    case KO_EQ: kind = 0; break;
    case KO_NE: kind = 0; neg = 1; break;
    case KO_LT: kind = 1; break;
    case KO_LE: kind = 2; break;
    case KO_GT: kind = 2; neg = 1; break;
    case KO_GE: kind = 1; neg = 1; break;
    case KO_RG: kind = 3; break;
    default: return;
    }
    size_t groups = n / 64;
    const uint64_t* p = src;
    /* one specialised loop per (kind, width) — the C analogue of the reference's generated
     * cmp_<w>_<op> tables (cmp_eq.go:13-23): constant width → constant shifts after unrolling */
#define KO_W(K) case K: for (size_t g = 0; g < groups; g++) { uint64_t w = group_cmp(KIND, p, K, a, b); put64(bits + g * 8, neg ? ~w : w); p += K; } break;
#define KO_W8(B) KO_W(B) KO_W(B + 1) KO_W(B + 2) KO_W(B + 3) KO_W(B + 4) KO_W(B + 5) KO_W(B + 6) KO_W(B + 7)
#define KO_ALLW switch (log2) { KO_W8(0) KO_W8(8) KO_W8(16) KO_W8(24) KO_W8(32) KO_W8(40) KO_W8(48) KO_W8(56) KO_W(64) }
    switch (kind) {
#define KIND 0
    case 0: KO_ALLW break;
#undef KIND
#define KIND 1
    case 1: KO_ALLW break;
#undef KIND
#define KIND 2
    case 2: KO_ALLW break;
#undef KIND
#define KIND 3
    default: KO_ALLW break;
#undef KIND
    }
    size_t rem = n & 63;
    if (rem) {
        size_t k = n & ~(size_t)63;
        uint64_t c2 = b - a;
        size_t tail_words = ko_bitpack_size(log2, rem) / 8;
        for (size_t i = 0; i < rem; i++) {
            uint64_t v = ko_bitpack_value(p, tail_words, i, log2, 0);
            int r;
            switch (kind) {
            case 0: r = v == a; break;
            case 1: r = v < a; break;
            case 2: r = v <= a; break;
            default: r = (v - a) <= c2; break;
            }
            if (neg) r = !r;
            if (r) bits[(k + i) >> 3] |= (uint8_t)(1u << ((k + i) & 7));
        }
    }
}
