/*
 * ko_cmp_bitset.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h):
 * compare kernels, bitset ops, varint.
 */
#include "knox_oracle.h"
#include <string.h>

int ko_type_size(int type) {
    switch (type) {
    case KO_I64: case KO_U64: case KO_F64: return 8;
    case KO_I32: case KO_U32: case KO_F32: return 4;
    case KO_I16: case KO_U16: return 2;
    case KO_I8: case KO_U8: return 1;
    }
    return 0;
}

/* ------------------------------------------------------------------ varint
 * pkg/num/varint.go:85-155 (PutUvarint) and :157-192 (Uvarint) */
int ko_put_uvarint(uint8_t* b, uint64_t x) {
    if (x <= 240) { b[0] = (uint8_t)x; return 1; }
    if (x <= 2287) { uint64_t y = x - 240; b[0] = (uint8_t)((y >> 8) + 241); b[1] = (uint8_t)y; return 2; }
    if (x <= 67823) { uint64_t y = x - 2288; b[0] = 249; b[1] = (uint8_t)(y >> 8); b[2] = (uint8_t)y; return 3; }
    int nbytes;
    if (x <= 0xffffffull) nbytes = 3;
    else if (x <= 0xffffffffull) nbytes = 4;
    else if (x <= 0xffffffffffull) nbytes = 5;
    else if (x <= 0xffffffffffffull) nbytes = 6;
    else if (x <= 0xffffffffffffffull) nbytes = 7;
    else nbytes = 8;
    b[0] = (uint8_t)(247 + nbytes);
    for (int i = 0; i < nbytes; i++) b[1 + i] = (uint8_t)(x >> (8 * (nbytes - 1 - i)));
    return nbytes + 1;
}

int ko_uvarint(const uint8_t* b, uint64_t* x) {
    uint8_t b0 = b[0];
    if (b0 <= 240) { *x = b0; return 1; }
    if (b0 <= 248) { *x = 240 + ((uint64_t)(b0 - 241) << 8) + b[1]; return 2; }
    if (b0 == 249) { *x = 2288 + ((uint64_t)b[1] << 8) + b[2]; return 3; }
    int nbytes = b0 - 247; /* 250 → 3 … 255 → 8 */
    uint64_t v = 0;
    for (int i = 0; i < nbytes; i++) v = (v << 8) | b[1 + i];
    *x = v;
    return nbytes + 1;
}

/* ------------------------------------------------------------------ cmp
 * internal/cmp/number.go:13-209 (eq/ne/lt/le/gt/ge), :211-243 (bw as U(v-a) <= U(b-a)),
 * internal/cmp/float.go:13-242 (IEEE compares; bw = a <= v && v <= b).
 * Structure kept: whole output bytes are OVERWRITTEN for full groups of 8, the tail
 * (<8 values) ORs bits into res[n] (number.go:33-41). */
#define KO_CMP_LOOP(T, EXPR)                                                     \
    do {                                                                         \
        const T* s = (const T*)src;                                              \
        size_t nb = n / 8, idx = 0;                                              \
        for (size_t i = 0; i < nb; i++) {                                        \
            uint8_t x = 0;                                                       \
            for (int k = 0; k < 8; k++) { T v = s[idx + k]; x |= (uint8_t)((EXPR) ? 1u << k : 0); } \
            bits[i] = x;                                                         \
            cnt += __builtin_popcount(x);                                        \
            idx += 8;                                                            \
        }                                                                        \
        for (size_t k = 0; idx + k < n; k++) {                                   \
            T v = s[idx + k];                                                    \
            if (EXPR) { bits[nb] |= (uint8_t)(1u << k); cnt++; }                 \
        }                                                                        \
    } while (0)

#define KO_CMP_INT(T, U)                                                         \
    static int64_t cmp_##T(int op, const void* src, size_t n, uint64_t a64, uint64_t b64, uint8_t* bits) { \
        int64_t cnt = 0;                                                         \
        T a = (T)a64, b = (T)b64;                                                \
        U diff = (U)((U)b - (U)a);                                               \
        switch (op) {                                                            \
        case KO_EQ: KO_CMP_LOOP(T, v == a); break;                               \
        case KO_NE: KO_CMP_LOOP(T, v != a); break;                               \
        case KO_LT: KO_CMP_LOOP(T, v < a); break;                                \
        case KO_LE: KO_CMP_LOOP(T, v <= a); break;                               \
        case KO_GT: KO_CMP_LOOP(T, v > a); break;                                \
        case KO_GE: KO_CMP_LOOP(T, v >= a); break;                               \
        case KO_RG: KO_CMP_LOOP(T, (U)((U)v - (U)a) <= diff); break;             \
        default: return -1;                                                      \
        }                                                                        \
        return cnt;                                                              \
    }

#define KO_CMP_FLT(T)                                                            \
    static int64_t cmp_##T(int op, const void* src, size_t n, T a, T b, uint8_t* bits) { \
        int64_t cnt = 0;                                                         \
        switch (op) {                                                            \
        case KO_EQ: KO_CMP_LOOP(T, v == a); break;                               \
        case KO_NE: KO_CMP_LOOP(T, v != a); break;                               \
        case KO_LT: KO_CMP_LOOP(T, v < a); break;                                \
        case KO_LE: KO_CMP_LOOP(T, v <= a); break;                               \
        case KO_GT: KO_CMP_LOOP(T, v > a); break;                                \
        case KO_GE: KO_CMP_LOOP(T, v >= a); break;                               \
        case KO_RG: KO_CMP_LOOP(T, a <= v && v <= b); break;                     \
        default: return -1;                                                      \
        }                                                                        \
        return cnt;                                                              \
    }

KO_CMP_INT(int64_t, uint64_t)
KO_CMP_INT(int32_t, uint32_t)
KO_CMP_INT(int16_t, uint16_t)
KO_CMP_INT(int8_t, uint8_t)
KO_CMP_INT(uint64_t, uint64_t)
KO_CMP_INT(uint32_t, uint32_t)
KO_CMP_INT(uint16_t, uint16_t)
KO_CMP_INT(uint8_t, uint8_t)
KO_CMP_FLT(double)
KO_CMP_FLT(float)

int64_t ko_cmp(int type, int op, const void* src, size_t n, uint64_t a, uint64_t b, uint8_t* bits) {
    switch (type) {
    case KO_I64: return cmp_int64_t(op, src, n, a, b, bits);
    case KO_I32: return cmp_int32_t(op, src, n, a, b, bits);
    case KO_I16: return cmp_int16_t(op, src, n, a, b, bits);
    case KO_I8: return cmp_int8_t(op, src, n, a, b, bits);
    case KO_U64: return cmp_uint64_t(op, src, n, a, b, bits);
    case KO_U32: return cmp_uint32_t(op, src, n, a, b, bits);
    case KO_U16: return cmp_uint16_t(op, src, n, a, b, bits);
    case KO_U8: return cmp_uint8_t(op, src, n, a, b, bits);
    case KO_F64: { double x, y; memcpy(&x, &a, 8); memcpy(&y, &b, 8); return cmp_double(op, src, n, x, y, bits); }
    case KO_F32: { uint32_t ua = (uint32_t)a, ub = (uint32_t)b; float x, y; memcpy(&x, &ua, 4); memcpy(&y, &ub, 4);
                   return cmp_float(op, src, n, x, y, bits); }
    }
    return -1;
}

/* ------------------------------------------------------------------ bitset
 * internal/bitset/generic/utils.go:12-18 */
uint8_t ko_bytemask(size_t size) { return (uint8_t)(0xff >> (7 - ((size - 1) & 7))); }

static size_t blen(size_t size) { return (size + 7) >> 3; }

/* generic/bitset.go:13-48 */
void ko_bitset_and(uint8_t* dst, const uint8_t* src, size_t size) {
    size_t l = blen(size); if (!l) return;
    for (size_t i = 0; i < l; i++) dst[i] &= src[i];
    dst[l - 1] &= ko_bytemask(size);
}
/* generic/bitset.go:50-118 */
void ko_bitset_and_flag(uint8_t* dst, const uint8_t* src, size_t size, int* any, int* all) {
    size_t l = size >> 3; uint8_t a = 0, f = 0xff;
    for (size_t i = 0; i < l; i++) { dst[i] &= src[i]; a |= dst[i]; f &= dst[i]; }
    if (size & 7) {
        dst[l] &= src[l]; dst[l] &= ko_bytemask(size);
        a |= dst[l]; f &= (uint8_t)(dst[l] | ~ko_bytemask(size));
    }
    *any = a != 0; *all = f == 0xff;
}
/* generic/bitset.go:120-151 */
void ko_bitset_andnot(uint8_t* dst, const uint8_t* src, size_t size) {
    size_t l = blen(size); if (!l) return;
    for (size_t i = 0; i < l; i++) dst[i] &= (uint8_t)~src[i];
    dst[l - 1] &= ko_bytemask(size);
}
/* generic/bitset.go:153-188 */
void ko_bitset_or(uint8_t* dst, const uint8_t* src, size_t size) {
    size_t l = blen(size); if (!l) return;
    for (size_t i = 0; i < l; i++) dst[i] |= src[i];
    dst[l - 1] &= ko_bytemask(size);
}
/* generic/bitset.go:190-258 */
void ko_bitset_or_flag(uint8_t* dst, const uint8_t* src, size_t size, int* any, int* all) {
    size_t l = size >> 3; uint8_t a = 0, f = 0xff;
    for (size_t i = 0; i < l; i++) { dst[i] |= src[i]; a |= dst[i]; f &= dst[i]; }
    if (size & 7) {
        dst[l] |= src[l]; dst[l] &= ko_bytemask(size);
        a |= dst[l]; f &= (uint8_t)(dst[l] | ~ko_bytemask(size));
    }
    *any = a != 0; *all = f == 0xff;
}
/* generic/bitset.go:260-295 */
void ko_bitset_xor(uint8_t* dst, const uint8_t* src, size_t size) {
    size_t l = blen(size); if (!l) return;
    for (size_t i = 0; i < l; i++) dst[i] ^= src[i];
    dst[l - 1] &= ko_bytemask(size);
}
/* generic/bitset.go:297-330 */
void ko_bitset_neg(uint8_t* buf, size_t size) {
    size_t l = blen(size); if (!l) return;
    for (size_t i = 0; i < l; i++) buf[i] = (uint8_t)~buf[i];
    buf[l - 1] &= ko_bytemask(size);
}
/* internal/bitset/bitset.go:520-531 (One): all ones, tail masked */
void ko_bitset_one(uint8_t* buf, size_t size) {
    size_t l = blen(size); if (!l) return;
    memset(buf, 0xff, l);
    buf[l - 1] &= ko_bytemask(size);
}
/* internal/bitset/bitset.go:156-200 (SetRange): sets bits [start, end] inclusive,
 * clamped to the bitset; no-op when start > end after clamping. */
void ko_bitset_set_range(uint8_t* buf, size_t size, int64_t start, int64_t end) {
    if (size == 0) return;
    if (start < 0) start = 0;
    if (end >= (int64_t)size) end = (int64_t)size - 1;
    for (int64_t i = start; i <= end; i++) buf[i >> 3] |= (uint8_t)(1u << (i & 7));
}
/* generic/bitset.go:332-353: popcount with the last byte masked to size */
int64_t ko_bitset_popcount(const uint8_t* buf, size_t size) {
    size_t l = blen(size); if (!l) return 0;
    int64_t cnt = 0;
    for (size_t i = 0; i + 1 < l; i++) cnt += __builtin_popcount(buf[i]);
    cnt += __builtin_popcount(buf[l - 1] & ko_bytemask(size));
    return cnt;
}
/* generic/bitset.go:355-396: ascending row ids of set bits (tail masked first) */
size_t ko_bitset_indexes(const uint8_t* buf, size_t size, uint32_t* dst) {
    size_t l = blen(size), j = 0;
    for (size_t i = 0; i < l; i++) {
        uint8_t b = buf[i];
        if (i == l - 1) b &= ko_bytemask(size);
        while (b) { int k = __builtin_ctz(b); dst[j++] = (uint32_t)(i * 8 + k); b &= (uint8_t)(b - 1); }
    }
    return j;
}
