/*
 * ko_hash_reduce.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h):
 * XXH3-64, bloom filter, reducers, filter-tree combination, zone-map range match and
 * the multi-threaded CPU baseline driver used by bench.py.
 */
#include "knox_oracle.h"
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ XXH3
 * Integer specialisations: internal/hash/xxh3.go:22-58 (closed forms of XXH3_64bits,
 * seed 0, for 8/4/2/1-byte little-endian inputs). */
static inline uint64_t rol64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
#define KEY64_008 0x1cad21f72c81017cull
#define KEY64_016 0xdb979083e96dd4deull
#define KEY32_000 0xbe4ba423ull
#define KEY32_004 0x396cfeb8ull
#define PRIME32_1 0x9E3779B1ull
#define PRIME32_2 0x85EBCA77ull
#define PRIME32_3 0xC2B2AE3Dull
#define PRIME64_1 0x9E3779B185EBCA87ull
#define PRIME64_2 0xC2B2AE3D27D4EB4Full
#define PRIME64_3 0x165667B19E3779F9ull
#define PRIME64_4 0x85EBCA77C2B2AE63ull
#define PRIME64_5 0x27D4EB2F165667C5ull

static uint64_t rrmxmx(uint64_t h, uint64_t len) {
    h ^= rol64(h, 49) ^ rol64(h, 24);
    h *= 0x9fb21c651e98df25ull;
    h ^= (h >> 35) + len;
    h *= 0x9fb21c651e98df25ull;
    h ^= h >> 28;
    return h;
}
/* xxh3.go:247-254 (xxhAvalancheSmall == XXH64_avalanche) */
static uint64_t xxh64_avalanche(uint64_t x) {
    x ^= x >> 33; x *= PRIME64_2; x ^= x >> 29; x *= PRIME64_3; x ^= x >> 32;
    return x;
}
uint64_t ko_xxh3_u64(uint64_t v) { return rrmxmx(((v >> 32) + (v << 32)) ^ (KEY64_008 ^ KEY64_016), 8); }   /* xxh3.go:22-33 */
uint64_t ko_xxh3_u32(uint32_t v) { return rrmxmx(((uint64_t)v + ((uint64_t)v << 32)) ^ (KEY64_008 ^ KEY64_016), 4); } /* :35-46 */
uint64_t ko_xxh3_u16(uint16_t v) { /* :48-52 */
    uint64_t h = (((uint64_t)v * ((1u << 24) + 1)) >> 8) + (2u << 8);
    h ^= KEY32_000 ^ KEY32_004;
    return xxh64_avalanche(h);
}
uint64_t ko_xxh3_u8(uint8_t v) { /* :54-58 */
    uint64_t h = (uint64_t)v * ((1u << 24) + (1u << 16) + 1) + (1u << 8);
    h ^= KEY32_000 ^ KEY32_004;
    return xxh64_avalanche(h);
}

/* Byte strings: internal/hash/hash.go:26 `Hash = xxh3.Hash` → github.com/zeebo/xxh3 v1.1.0
 * (go.mod:20, NOT vendored).  Restated from the published XXH3_64bits algorithm
 * (Cyan4973/xxHash v0.8, default secret, seed 0); cross-checked against python-xxhash. */
static const uint8_t kSecret[192] = {
    0xb8, 0xfe, 0x6c, 0x39, 0x23, 0xa4, 0x4b, 0xbe, 0x7c, 0x01, 0x81, 0x2c, 0xf7, 0x21, 0xad, 0x1c, 0xde, 0xd4, 0x6d, 0xe9, 0x83, 0x90, 0x97, 0xdb,
    0x72, 0x40, 0xa4, 0xa4, 0xb7, 0xb3, 0x67, 0x1f, 0xcb, 0x79, 0xe6, 0x4e, 0xcc, 0xc0, 0xe5, 0x78, 0x82, 0x5a, 0xd0, 0x7d, 0xcc, 0xff, 0x72, 0x21,
    0xb8, 0x08, 0x46, 0x74, 0xf7, 0x43, 0x24, 0x8e, 0xe0, 0x35, 0x90, 0xe6, 0x81, 0x3a, 0x26, 0x4c, 0x3c, 0x28, 0x52, 0xbb, 0x91, 0xc3, 0x00, 0xcb,
    0x88, 0xd0, 0x65, 0x8b, 0x1b, 0x53, 0x2e, 0xa3, 0x71, 0x64, 0x48, 0x97, 0xa2, 0x0d, 0xf9, 0x4e, 0x38, 0x19, 0xef, 0x46, 0xa9, 0xde, 0xac, 0xd8,
    0xa8, 0xfa, 0x76, 0x3f, 0xe3, 0x9c, 0x34, 0x3f, 0xf9, 0xdc, 0xbb, 0xc7, 0xc7, 0x0b, 0x4f, 0x1d, 0x8a, 0x51, 0xe0, 0x4b, 0xcd, 0xb4, 0x59, 0x31,
    0xc8, 0x9f, 0x7e, 0xc9, 0xd9, 0x78, 0x73, 0x64, 0xea, 0xc5, 0xac, 0x83, 0x34, 0xd3, 0xeb, 0xc3, 0xc5, 0x81, 0xa0, 0xff, 0xfa, 0x13, 0x63, 0xeb,
    0x17, 0x0d, 0xdd, 0x51, 0xb7, 0xf0, 0xda, 0x49, 0xd3, 0x16, 0x55, 0x26, 0x29, 0xd4, 0x68, 0x9e, 0x2b, 0x16, 0xbe, 0x58, 0x7d, 0x47, 0xa1, 0xfc,
    0x8f, 0xf8, 0xb8, 0xd1, 0x7a, 0xd0, 0x31, 0xce, 0x45, 0xcb, 0x3a, 0x8f, 0x95, 0x16, 0x04, 0x28, 0xaf, 0xd7, 0xfb, 0xca, 0xbb, 0x4b, 0x40, 0x7e,
};
static inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t mul_fold(uint64_t a, uint64_t b) { __uint128_t m = (__uint128_t)a * b; return (uint64_t)m ^ (uint64_t)(m >> 64); }
static inline uint64_t xxh3_avalanche(uint64_t h) { h ^= h >> 37; h *= 0x165667919E3779F9ull; h ^= h >> 32; return h; }
static inline uint64_t mix16(const uint8_t* in, const uint8_t* sec) { return mul_fold(rd64(in) ^ rd64(sec), rd64(in + 8) ^ rd64(sec + 8)); }

static void acc512(uint64_t* acc, const uint8_t* in, const uint8_t* sec) {
    for (int i = 0; i < 8; i++) {
        uint64_t dv = rd64(in + 8 * i), dk = dv ^ rd64(sec + 8 * i);
        acc[i ^ 1] += dv;
        acc[i] += (uint64_t)(uint32_t)dk * (dk >> 32);
    }
}

uint64_t ko_xxh3_bytes(const uint8_t* in, size_t len) {
    const uint8_t* s = kSecret;
    if (len == 0) return xxh64_avalanche(rd64(s + 56) ^ rd64(s + 64));
    if (len <= 3) {
        uint32_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
        uint32_t comb = (c1 << 16) | (c2 << 24) | c3 | ((uint32_t)len << 8);
        return xxh64_avalanche((uint64_t)comb ^ (uint64_t)(rd32(s) ^ rd32(s + 4)));
    }
    if (len <= 8) {
        uint64_t in64 = (uint64_t)rd32(in + len - 4) + ((uint64_t)rd32(in) << 32);
        return rrmxmx(in64 ^ (rd64(s + 8) ^ rd64(s + 16)), len);
    }
    if (len <= 16) {
        uint64_t lo = rd64(in) ^ (rd64(s + 24) ^ rd64(s + 32));
        uint64_t hi = rd64(in + len - 8) ^ (rd64(s + 40) ^ rd64(s + 48));
        uint64_t acc = len + __builtin_bswap64(lo) + hi + mul_fold(lo, hi);
        return xxh3_avalanche(acc);
    }
    if (len <= 128) {
        uint64_t acc = len * PRIME64_1;
        if (len > 32) {
            if (len > 64) {
                if (len > 96) { acc += mix16(in + 48, s + 96); acc += mix16(in + len - 64, s + 112); }
                acc += mix16(in + 32, s + 64); acc += mix16(in + len - 48, s + 80);
            }
            acc += mix16(in + 16, s + 32); acc += mix16(in + len - 32, s + 48);
        }
        acc += mix16(in, s); acc += mix16(in + len - 16, s + 16);
        return xxh3_avalanche(acc);
    }
    if (len <= 240) {
        uint64_t acc = len * PRIME64_1;
        size_t rounds = len / 16;
        for (size_t i = 0; i < 8; i++) acc += mix16(in + 16 * i, s + 16 * i);
        acc = xxh3_avalanche(acc);
        for (size_t i = 8; i < rounds; i++) acc += mix16(in + 16 * i, s + 16 * (i - 8) + 3);
        acc += mix16(in + len - 16, s + 136 - 17);
        return xxh3_avalanche(acc);
    }
    uint64_t acc[8] = {PRIME32_3, PRIME64_1, PRIME64_2, PRIME64_3, PRIME64_4, PRIME32_2, PRIME64_5, PRIME32_1};
    const size_t nstripes = (192 - 64) / 8, block = nstripes * 64;
    size_t nblocks = (len - 1) / block;
    for (size_t b = 0; b < nblocks; b++) {
        for (size_t n = 0; n < nstripes; n++) acc512(acc, in + b * block + n * 64, s + n * 8);
        for (int i = 0; i < 8; i++) { uint64_t a = acc[i]; a ^= a >> 47; a ^= rd64(s + 192 - 64 + 8 * i); a *= PRIME32_1; acc[i] = a; }
    }
    size_t rem = ((len - 1) - block * nblocks) / 64;
    for (size_t n = 0; n < rem; n++) acc512(acc, in + nblocks * block + n * 64, s + n * 8);
    acc512(acc, in + len - 64, s + 192 - 64 - 7);
    uint64_t r = len * PRIME64_1;
    for (int i = 0; i < 4; i++) r += mul_fold(acc[2 * i] ^ rd64(s + 11 + 16 * i), acc[2 * i + 1] ^ rd64(s + 11 + 16 * i + 8));
    return xxh3_avalanche(r);
}

/* ------------------------------------------------------------------ bloom
 * internal/filter/bloom/bloom.go:48-60 (NewFilter: m → pow2, buf[0] = k = 4),
 * :136-150 (containsUnroll4), :152-160 (addUnroll4), :182-184 (Contains: h0 = lo32, h1 = hi32) */
static size_t pow2_ceil(size_t m) { size_t p = 1; while (p < m) p <<= 1; return p; }
size_t ko_bloom_bytes(size_t m_bits) { size_t m = pow2_ceil(m_bits < 8 ? 8 : m_bits); return 1 + (m >> 3); }
void ko_bloom_init(uint8_t* buf, size_t m_bits) { size_t l = ko_bloom_bytes(m_bits); memset(buf, 0, l); buf[0] = 4; }
void ko_bloom_add(uint8_t* buf, size_t buflen, uint64_t h) {
    uint32_t mask = (uint32_t)((buflen - 1) * 8 - 1), h0 = (uint32_t)h, h1 = (uint32_t)(h >> 32);
    uint8_t* bits = buf + 1;
    for (uint32_t i = 0; i < buf[0]; i++) { bits[(h0 & mask) >> 3] |= (uint8_t)(1u << (h0 & 7)); h0 += h1; }
}
int ko_bloom_contains(const uint8_t* buf, size_t buflen, uint64_t h) {
    uint32_t mask = (uint32_t)((buflen - 1) * 8 - 1), h0 = (uint32_t)h, h1 = (uint32_t)(h >> 32);
    const uint8_t* bits = buf + 1;
    for (uint32_t i = 0; i < buf[0]; i++) { if (!(bits[(h0 & mask) >> 3] & (1u << (h0 & 7)))) return 0; h0 += h1; }
    return 1;
}

size_t ko_bloom_build(uint8_t* buf, size_t cap, int elem_bytes, const uint8_t* values, const uint32_t* offsets, size_t n,
                      int cardinality, int factor) {
    if (cardinality <= 0 || factor <= 0) return 0;            /* filter.go:297-299 */
    size_t m = 8;                                              /* bloom.pow2: smallest power of two >= v, at least 8 */
    while (m < (size_t)cardinality * (size_t)factor * 8) m <<= 1;
    size_t len = 1 + (m >> 3);
    if (len > cap) return 0;
    memset(buf, 0, len); buf[0] = 4;
    for (size_t i = 0; i < n; i++) {
        uint64_t h, v = 0;
        if (elem_bytes) memcpy(&v, values + i * (size_t)elem_bytes, (size_t)elem_bytes);
        switch (elem_bytes) {
        case 8: h = ko_xxh3_u64(v); break;
        case 4: h = ko_xxh3_u32((uint32_t)v); break;
        case 2: h = ko_xxh3_u16((uint16_t)v); break;
        case 1: h = ko_xxh3_u8((uint8_t)v); break;
        default: h = ko_xxh3_bytes(values + offsets[i], offsets[i + 1] - offsets[i]); break;
        }
        ko_bloom_add(buf, len, h);
    }
    return len;
}

/* ------------------------------------------------------------------ reducers
 * internal/reducer/reducer.go:138-149 (Count), :168-179 (Sum: r.v += v in T),
 * :256-267 (Max: first || r.v < v), :286-297 (Min: first || r.v > v), fed row by row
 * in ascending row order by StreamResult.Append (internal/query/result.go:96-152). */
static void reduce_one(int type, uint64_t v, ko_agg* st) {
    st->count++;
    if (type == KO_F64) {
        double d, s, mn, mx; memcpy(&d, &v, 8);
        memcpy(&s, &st->sum_bits, 8); s += d; memcpy(&st->sum_bits, &s, 8);
        memcpy(&mn, &st->min_bits, 8); memcpy(&mx, &st->max_bits, 8);
        if (!st->valid || mx < d) memcpy(&st->max_bits, &d, 8);
        if (!st->valid || mn > d) memcpy(&st->min_bits, &d, 8);
    } else if (type == KO_F32) {   /* SumReducer[float32]: the running sum is a float32 */
        uint32_t u = (uint32_t)v, w; float d, s, mn, mx; memcpy(&d, &u, 4);
        w = (uint32_t)st->sum_bits; memcpy(&s, &w, 4); s += d; memcpy(&w, &s, 4); st->sum_bits = w;
        w = (uint32_t)st->min_bits; memcpy(&mn, &w, 4); w = (uint32_t)st->max_bits; memcpy(&mx, &w, 4);
        if (!st->valid || mx < d) st->max_bits = u;
        if (!st->valid || mn > d) st->min_bits = u;
    } else if (type == KO_I64 || type == KO_I32 || type == KO_I16 || type == KO_I8) {
        st->sum_bits += v; /* wraps mod 2^64 like int64 `+=` */
        if (!st->valid || (int64_t)st->max_bits < (int64_t)v) st->max_bits = v;
        if (!st->valid || (int64_t)st->min_bits > (int64_t)v) st->min_bits = v;
    } else {
        st->sum_bits += v;
        if (!st->valid || st->max_bits < v) st->max_bits = v;
        if (!st->valid || st->min_bits > v) st->min_bits = v;
    }
    st->valid = 1;
}

void ko_reduce(int type, const uint64_t* vals, size_t n, const uint8_t* bits, ko_agg* st) {
    for (size_t i = 0; i < n; i++) {
        if (bits && !((bits[i >> 3] >> (i & 7)) & 1)) continue;
        reduce_one(type, vals[i], st);
    }
}

/* ------------------------------------------------------------------ time-bucketed reduce
 * The series query (pkg/series/series.go:192-256) maps every streamed row to the start of its window,
 * t = Interval.TruncateRelative(ts, Range.From) — a walk `last, next = base, Next(base)` that advances while
 * !ts.Before(next) (pkg/util/timeunit.go:234-243) — and NativeBucket.Push (internal/reducer/bucket_native.go:104-167)
 * feeds the row to that window's Reducer (reducer.go:138-297).  `edges` are the window starts produced by that walk
 * (edges[0] = From, edges[k+1] = Next(edges[k])); rows before edges[0] or at/after edges[nbuckets] are outside the
 * query's time range.  states: one ko_agg per window. */
static int ts_before(int ts_type, uint64_t a, uint64_t b) {
    if (ts_type == KO_I64 || ts_type == KO_I32 || ts_type == KO_I16 || ts_type == KO_I8) return (int64_t)a < (int64_t)b;
    return a < b;
}
void ko_bucket_reduce(int type, const uint64_t* vals, int ts_type, const uint64_t* ts, size_t n, const uint8_t* bits,
                      const uint64_t* edges, int nbuckets, ko_agg* states) {
    for (size_t i = 0; i < n; i++) {
        if (bits && !((bits[i >> 3] >> (i & 7)) & 1)) continue;
        if (ts_before(ts_type, ts[i], edges[0]) || !ts_before(ts_type, ts[i], edges[nbuckets])) continue;
        int k = 0;
        while (!ts_before(ts_type, ts[i], edges[k + 1])) k++;   /* TruncateRelative's walk */
        reduce_one(type, vals ? vals[i] : 0, &states[k]);
    }
}

/* Window starts of a fixed-duration unit (minutes, hours: `step` in the timestamp's resolution, dividing one day):
 * TimeUnit.Next(t, 1) = Truncate(t) + Duration (timeunit.go:245-249) with Truncate = time.Truncate(Duration), i.e.
 * rounding DOWN to a multiple of the step (:196-199; multiples counted from Go's zero time, which is a whole number of
 * days before the Unix epoch).  out[0] = from, then aligned starts until one is >= to; returns the number of edges. */
int ko_window_edges(int64_t from, int64_t to, int64_t step, int64_t* out, int cap) {
    int n = 0;
    int64_t t = from;
    while (n < cap) {
        out[n++] = t;
        if (t >= to) break;
        int64_t m = t % step; if (m < 0) m += step;
        t = (t - m) + step;
    }
    return n;
}

/* ------------------------------------------------------------------ filter tree
 * internal/operator/filter/match_core.go:44-130 (MatchAnd: bits.One(), AndFlag per child),
 * :132-215 (MatchOr: zero, Or per child).  The zone-map "always true" skips and the
 * early-outs do not change the result, so the combination is plain AND/OR with the tail
 * kept zero. */
int ko_tree_eval(const uint8_t* postfix, int npost, uint8_t* const* leaf_bits, int nleaves, size_t n, uint8_t* out) {
    size_t l = (n + 7) / 8;
    uint8_t* stack[64]; int sp = 0, rc = 0;
    for (int i = 0; i < npost; i++) {
        uint8_t op = postfix[i];
        if (op < 0x80) {
            if (op >= nleaves || sp >= 64) { rc = -1; break; }
            uint8_t* b = (uint8_t*)malloc(l + 8); memcpy(b, leaf_bits[op], l); stack[sp++] = b;
        } else {
            if (sp < 2) { rc = -1; break; }
            uint8_t* r = stack[--sp]; uint8_t* d = stack[sp - 1];
            int any, all;
            if (op == 0xFE) ko_bitset_and_flag(d, r, n, &any, &all); else ko_bitset_or(d, r, n);
            free(r);
        }
    }
    if (rc == 0 && sp == 1) memcpy(out, stack[0], l); else rc = -1;
    while (sp > 0) free(stack[--sp]);
    return rc;
}

/* ------------------------------------------------------------------ zone maps
 * MatchRangeVectors semantics, internal/operator/filter/match_num.go:357-371 (EQ:
 * min <= v && max >= v), :394-401 (NE: undecided → true), :429-434 (GT: max > v),
 * :460-465 (GE), :491-496 (LT: min < v), :522-527 (LE), :573-588 (RG: min <= to &&
 * max >= from), :810-817 (NIN: true).  IN is handled by the caller with the set. */
int ko_match_range(int type, int op, uint64_t a, uint64_t b, uint64_t minv, uint64_t maxv) {
    int sg = type >= KO_I64 && type <= KO_I8;
#define LT(x, y) (sg ? (int64_t)(x) < (int64_t)(y) : (x) < (y))
#define LE(x, y) (!LT(y, x))
    switch (op) {
    case KO_EQ: return LE(minv, a) && LE(a, maxv);
    case KO_NE: case KO_NI: return 1;
    case KO_GT: return LT(a, maxv);
    case KO_GE: return LE(a, maxv);
    case KO_LT: return LT(minv, a);
    case KO_LE: return LE(minv, a);
    case KO_RG: return LE(minv, b) && LE(a, maxv);
    }
#undef LT
#undef LE
    return 1;
}

/* ------------------------------------------------------------------ CPU baseline
 * The reference scans packs one after another on one goroutine
 * (internal/pack/table/reader.go:299-449); for a generous baseline packs are statically
 * partitioned over nthreads.  Each pack runs the fused bitpack compare (a8) + popcount. */
typedef struct {
    const uint64_t* const* packs; const size_t* nrows; size_t lo, hi;
    int log2, op; uint64_t a, b; uint8_t* const* bitsets; int64_t total; int simd;
} bl_arg;

static void* bl_worker(void* p) {
    bl_arg* g = (bl_arg*)p;
    int64_t tot = 0;
    for (size_t i = g->lo; i < g->hi; i++) {
        size_t n = g->nrows[i];
        memset(g->bitsets[i], 0, (n + 7) / 8);
        if (!(g->simd && ko_bitpack_cmp_simd(g->op, g->packs[i], g->log2, g->a, g->b, n, g->bitsets[i])))
            ko_bitpack_cmp(g->op, g->packs[i], g->log2, g->a, g->b, n, g->bitsets[i]);
        tot += ko_bitset_popcount(g->bitsets[i], n);
    }
    g->total = tot;
    return NULL;
}

/* nthreads < 0: |nthreads| threads running the scalar port only (the shape of the reference's generated Go code);
 * nthreads > 0: the AVX-512 kernel of ko_simd.c where the host has VBMI (packs must be padded by 64 readable bytes) */
int64_t ko_baseline_bitpack_scan(const uint64_t* const* packs, const size_t* nrows, size_t npacks,
                                 int log2, int op, uint64_t a, uint64_t b,
                                 uint8_t* const* bitsets, int nthreads) {
    int simd = nthreads > 0;
    if (nthreads < 0) nthreads = -nthreads;
    if (nthreads < 1) nthreads = 1;
    if ((size_t)nthreads > npacks) nthreads = (int)(npacks ? npacks : 1);
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    bl_arg* args = (bl_arg*)calloc((size_t)nthreads, sizeof(bl_arg));
    for (int t = 0; t < nthreads; t++) {
        args[t] = (bl_arg){packs, nrows, npacks * (size_t)t / (size_t)nthreads, npacks * (size_t)(t + 1) / (size_t)nthreads,
                           log2, op, a, b, bitsets, 0, simd};
        pthread_create(&th[t], NULL, bl_worker, &args[t]);
    }
    int64_t tot = 0;
    for (int t = 0; t < nthreads; t++) { pthread_join(th[t], NULL); tot += args[t].total; }
    free(th); free(args);
    return tot;
}
