/*
 * knox_oracle.h — CPU restatement ("oracle") of KnoxDB's pack-engine scan path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.  The product
 * path (libknoxgpu.so) never links, loads or calls it.
 *
 * The reference (blockwatch-cc/knoxdb, Go 1.26 + Plan-9 asm) cannot be built in
 * this image (no Go toolchain), so this is a plain-C restatement.  Every function
 * cites the reference file:line it follows (paths relative to /root/reference).
 *
 * Parity pinning (tests/test_oracle_golden.py, runs on CPU):
 *   - cmp kernels      : pinned by the 1476 golden cases of internal/cmp/tests/<type>.go
 *   - bitset pop/index : pinned by internal/bitset/tests/{pop,run}.go
 *   - xxh3 u32/u64     : pinned by internal/hash/xxh3_test.go:14-31
 *   - bitpack / containers / bloom : the reference holds NO golden bytes for these;
 *     pinned the way the reference pins them — round trip + agreement with the
 *     scalar predicate on the original values (bitpack/tests/tests.go:60-298,
 *     encode/tests/tests.go:140-205, bloom_test.go:18-150)
 *   - reducers         : the reference has no tests at all → PARITY UNPINNED beyond
 *     the 6-line function bodies (internal/reducer/reducer.go:138-314)
 *   - xxh3 byte strings: third-party github.com/zeebo/xxh3 v1.1.0 (go.mod:20),
 *     canonical XXH3_64bits(seed 0); cross-checked against python-xxhash when present.
 */
#ifndef KNOX_ORACLE_H
#define KNOX_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* element types: values equal types.BlockType (internal/types/block.go:20-36) */
enum {
    KO_I64 = 1, KO_I32 = 2, KO_I16 = 3, KO_I8 = 4,
    KO_U64 = 5, KO_U32 = 6, KO_U16 = 7, KO_U8 = 8,
    KO_F64 = 9, KO_F32 = 10,
};

/* compare ops: values equal types.FilterMode (internal/types/mode.go:14-23) */
enum {
    KO_EQ = 1, KO_NE = 2, KO_GT = 3, KO_GE = 4, KO_LT = 5, KO_LE = 6,
    KO_IN = 7, KO_NI = 8, KO_RG = 9,
};

/* container ids (internal/encode/container.go:20-55) */
enum {
    KO_TCONST = 1, KO_TDELTA = 2, KO_TRUNEND = 3, KO_TBITPACK = 4, KO_TDICT = 5,
    KO_TS8B = 6, KO_TRAW = 7, KO_TFLOATRAW = 15,
    KO_TSTRCONST = 16, KO_TSTRFIXED = 17, KO_TSTRCOMPACT = 18, KO_TSTRDICT = 19,   /* container.go:43-46 */
};

int ko_type_size(int type);

/* ---- varint: pkg/num/varint.go:85-192 (SQLite4 style) ---- */
int ko_put_uvarint(uint8_t* b, uint64_t x);
int ko_uvarint(const uint8_t* b, uint64_t* x);

/* ---- compare kernels: internal/cmp/number.go:13-243, float.go:13-242 ----
 * a/b are passed as raw 64-bit patterns of the type (sign-extended ints, IEEE bits
 * for floats; f32 bits in the low 32).  bits must hold ceil(n/8) bytes, pre-zeroed. */
int64_t ko_cmp(int type, int op, const void* src, size_t n, uint64_t a, uint64_t b, uint8_t* bits);

/* ---- bitset: internal/bitset/generic/bitset.go:13-396, utils.go:12-18 ---- */
uint8_t ko_bytemask(size_t size);
void    ko_bitset_and(uint8_t* dst, const uint8_t* src, size_t size);
void    ko_bitset_and_flag(uint8_t* dst, const uint8_t* src, size_t size, int* any, int* all);
void    ko_bitset_andnot(uint8_t* dst, const uint8_t* src, size_t size);
void    ko_bitset_or(uint8_t* dst, const uint8_t* src, size_t size);
void    ko_bitset_or_flag(uint8_t* dst, const uint8_t* src, size_t size, int* any, int* all);
void    ko_bitset_xor(uint8_t* dst, const uint8_t* src, size_t size);
void    ko_bitset_neg(uint8_t* buf, size_t size);
void    ko_bitset_one(uint8_t* buf, size_t size);
void    ko_bitset_set_range(uint8_t* buf, size_t size, int64_t start, int64_t end /*inclusive*/);
int64_t ko_bitset_popcount(const uint8_t* buf, size_t size);
size_t  ko_bitset_indexes(const uint8_t* buf, size_t size, uint32_t* dst);

/* ---- bitpack: internal/encode/bitpack/{bitpack,encode,decode,cmp}.go ---- */
size_t  ko_bitpack_size(int log2, size_t n);                                  /* bitpack.go:9-11 */
int     ko_log2range(int type, uint64_t minv, uint64_t maxv);                 /* types/number.go:163-169 */
size_t  ko_bitpack_encode(uint64_t* dst, const uint64_t* vals, size_t n, int log2, uint64_t minv); /* encode.go:216-246 */
void    ko_bitpack_decode(uint64_t* dst, const uint64_t* src, size_t n, int log2, uint64_t minv);
uint64_t ko_bitpack_value(const uint64_t* src, size_t nwords, size_t i, int log2, uint64_t minv); /* decode.go:56-74 */
void    ko_bitpack_cmp(int op, const uint64_t* src, int log2, uint64_t a, uint64_t b, size_t n, uint8_t* bits); /* cmp.go:20-130 */

/* ---- containers: internal/encode/int*.go, float_raw.go ---- */
typedef struct ko_container {
    int      ctype;      /* KO_T* */
    int      type;       /* element type KO_* */
    size_t   n;          /* logical length */
    uint64_t val;        /* const: Val ; delta/bitpack/s8b: For */
    uint64_t delta;      /* delta: Delta */
    int      log2;       /* bitpack width */
    const uint8_t* payload; /* raw values / packed words / s8b words */
    size_t   payload_len;
    struct ko_container* child[3]; /* dict: {Dict, Codes}; runend: {Values, Ends}; alp: {Values, Patches, Positions} */
    int      alp_e, alp_f, alp_flags; /* FloatAlp: Exponent, Factor, flags (1 = patched, 2 = safe int) */
} ko_container;

/* parses one container at buf (no outer compression byte); returns bytes consumed or <0 */
long    ko_container_load(int type, const uint8_t* buf, size_t len, ko_container** out);
void    ko_container_free(ko_container* c);
uint64_t ko_container_get(const ko_container* c, size_t i);          /* value as sign/zero-extended 64-bit (IEEE bits for floats) */
void    ko_container_decode(const ko_container* c, uint64_t* dst);   /* AppendTo(dst, nil) */
/* Match<Op>(a[,b], bits, nil): bits pre-zeroed, ceil(n/8) bytes */
void    ko_container_match(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits);
/* MatchInSet / MatchNotInSet with the roaring set flattened to a sorted unique u64 array */
void    ko_container_match_set(const ko_container* c, int negate, const uint64_t* set, size_t nset, uint8_t* bits);

/* ---- ALP float64: internal/encode/float_alp.go, internal/encode/alp/{constants,encoder,decoder}.go ---- */
#define KO_TFLOATALP 13
#define KO_TFLOATALPRD 14
/* ALP-RD (ko_alprd.c): left / right split of the IEEE bits; the cut (Shift) is kept in ko_container.log2 */
size_t  ko_store_alprd(uint8_t* dst, int type, const uint64_t* vals, size_t n, int shift /* < 0: analyse */);
long    ko_alprd_load(ko_container* c, const uint8_t* buf, size_t len);
uint64_t ko_alprd_get(const ko_container* c, size_t i);
void    ko_alprd_decode_all(const ko_container* c, uint64_t* dst);
void    ko_alprd_match(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits);
int64_t ko_alp_encode_single(double v, int e, int f, int* ok);   /* Encoder.EncodeSingle */
int64_t ko_alp_encode_above(double v, int e, int f);
int64_t ko_alp_encode_below(double v, int e, int f);
double  ko_alp_decode(int64_t enc, int e, int f);                 /* Decoder.decode */
size_t  ko_store_alp(uint8_t* dst, const uint64_t* vals, size_t n, int e, int f);   /* e < 0: choose exponents */
long    ko_alp_load(ko_container* c, const uint8_t* buf, size_t len);
uint64_t ko_alp_get(const ko_container* c, size_t i);
void    ko_alp_decode_all(const ko_container* c, uint64_t* dst);
void    ko_alp_match(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits);

/* Store(): writers used to synthesise packs (return bytes written into dst) */
size_t  ko_store_const(uint8_t* dst, uint64_t val, size_t n);
size_t  ko_store_delta(uint8_t* dst, uint64_t for_, uint64_t delta, size_t n);
size_t  ko_store_raw(uint8_t* dst, int type, const uint64_t* vals, size_t n);
size_t  ko_store_bitpack(uint8_t* dst, int type, const uint64_t* vals, size_t n);   /* computes min/max/log2 */
size_t  ko_store_best(uint8_t* dst, int type, const uint64_t* vals, size_t n, int lvl); /* scheme choice like context.go:257-293 */
size_t  ko_store_dict(uint8_t* dst, int type, const uint64_t* vals, size_t n);
size_t  ko_store_runend(uint8_t* dst, int type, const uint64_t* vals, size_t n);
size_t  ko_store_s8b(uint8_t* dst, int type, const uint64_t* vals, size_t n);
size_t  ko_store_bound(int type, size_t n);  /* upper bound on Store size for any scheme */

/* ---- simple8b: internal/encode/s8b/generic/{encode,decode,cmp}.go ---- */
size_t  ko_s8b_encode(uint64_t* dst, const uint64_t* vals, size_t n, uint64_t minv);
size_t  ko_s8b_decode(uint64_t* dst, size_t cap, const uint64_t* words, size_t nwords, uint64_t minv);

/* ---- hashing + bloom: internal/hash/xxh3.go:22-58, filter/bloom/bloom.go ---- */
uint64_t ko_xxh3_u64(uint64_t v);
uint64_t ko_xxh3_u32(uint32_t v);
uint64_t ko_xxh3_u16(uint16_t v);
uint64_t ko_xxh3_u8(uint8_t v);
uint64_t ko_xxh3_bytes(const uint8_t* p, size_t len);     /* XXH3_64bits seed 0, len <= 240 */
size_t  ko_bloom_bytes(size_t m_bits);                    /* 1 + pow2(m)/8 */
void    ko_bloom_init(uint8_t* buf, size_t m_bits);       /* k = 4 */
void    ko_bloom_add(uint8_t* buf, size_t buflen, uint64_t h);
int     ko_bloom_contains(const uint8_t* buf, size_t buflen, uint64_t h);
/* stats.BuildBloomFilter (internal/pack/stats/filter.go:296-367): m = pow2(cardinality*factor*8) bits, k = 4;
 * fixed-width values (elem_bytes 8/4/2/1) are hashed with hash.Vec64/32/16/8, byte strings (elem_bytes 0,
 * offsets[n+1]) with hash.Hash.  Returns the buffer length written ([k][m/8 bytes]), 0 if no filter. */
size_t  ko_bloom_build(uint8_t* buf, size_t cap, int elem_bytes, const uint8_t* values, const uint32_t* offsets, size_t n,
                       int cardinality, int factor);

/* ---- reducers over selected rows: internal/reducer/reducer.go:138-314 ----
 * sequential, in type T (i64/u64 wrap, f64 naive left-to-right).  bits may be NULL
 * (all rows).  Carry-in state lets callers chain packs in order like the reference. */
typedef struct ko_agg {
    int64_t  count;
    uint64_t sum_bits;   /* i64/u64: wrapped sum ; f64: IEEE bits of running sum */
    uint64_t min_bits, max_bits;
    int      valid;      /* 0 until the first row was reduced (r.t.IsZero()) */
} ko_agg;
void ko_reduce(int type, const uint64_t* vals, size_t n, const uint8_t* bits, ko_agg* state);
void ko_bucket_reduce(int type, const uint64_t* vals, int ts_type, const uint64_t* ts, size_t n, const uint8_t* bits,
                      const uint64_t* edges, int nbuckets, ko_agg* states);
int ko_window_edges(int64_t from, int64_t to, int64_t step, int64_t* out, int cap);

/* ---- byte-string containers: internal/encode/string_{const,fixed,compact,dict,match}.go (ko_string.c) ---- */
typedef struct ko_str ko_str;
size_t  ko_store_str(int kind, const uint8_t* bytes, const uint32_t* offs, size_t n, uint8_t* dst);
long    ko_str_load(const uint8_t* enc, size_t len, ko_str** out);
void    ko_str_free(ko_str* s);
size_t  ko_str_len(const ko_str* s);
const uint8_t* ko_str_get(const ko_str* s, size_t i, size_t* len);
void    ko_str_match(const ko_str* s, int op, const uint8_t* a, size_t al, const uint8_t* b, size_t bl, uint8_t* bits);

/* ---- filter tree: internal/operator/filter/match_core.go:14-215 ----
 * postfix program over leaf bitsets: byte < 0x80 → push leaf id, 0xFE = AND, 0xFF = OR
 * (binary).  leaf_bits[i] are ceil(n/8)-byte bitsets already matched per leaf. */
int  ko_tree_eval(const uint8_t* postfix, int npost, uint8_t* const* leaf_bits, int nleaves,
                  size_t n, uint8_t* out);

/* ---- zone-map + bloom pruning: internal/pack/stats/match.go:112-195,
 *      operator/filter/match_num.go MatchRange ---- */
int  ko_match_range(int type, int op, uint64_t a, uint64_t b, uint64_t minv, uint64_t maxv);

/* ---- CPU baseline driver (bench.py only): fused bitpack compare over many packs with
 *      nthreads pthreads; returns total matches. ---- */
/* AVX-512 (VBMI) version of ko_bitpack_cmp for the CPU baseline (ko_simd.c); 0 = not taken, use the scalar port */
int ko_simd_available(void);
int ko_bitpack_cmp_simd(int op, const uint64_t* src, int log2, uint64_t a, uint64_t b, size_t n, uint8_t* bits);
int64_t ko_baseline_bitpack_scan(const uint64_t* const* packs, const size_t* nrows, size_t npacks,
                                 int log2, int op, uint64_t a, uint64_t b,
                                 uint8_t* const* bitsets, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
