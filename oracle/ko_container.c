/*
 * ko_container.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h):
 * integer / float-raw column containers: Load, Store, Get, decode, fused matchers.
 *
 * Value convention: every integer value travels as a uint64 holding the sign- or
 * zero-extension of T; floats travel as IEEE bit patterns.
 */
#define _GNU_SOURCE
#include "knox_oracle.h"
#include <stdlib.h>
#include <string.h>

static int is_signed(int t) { return t >= KO_I64 && t <= KO_I8; }
static int is_float(int t) { return t == KO_F64 || t == KO_F32; }

/* T(x): truncate to the width of T, then extend back to 64 bit */
static uint64_t ext(int t, uint64_t x) {
    switch (t) {
    case KO_I32: return (uint64_t)(int64_t)(int32_t)x;
    case KO_I16: return (uint64_t)(int64_t)(int16_t)x;
    case KO_I8:  return (uint64_t)(int64_t)(int8_t)x;
    case KO_U32: case KO_F32: return (uint32_t)x;
    case KO_U16: return (uint16_t)x;
    case KO_U8:  return (uint8_t)x;
    }
    return x;
}
static int t_lt(int t, uint64_t a, uint64_t b) { return is_signed(t) ? (int64_t)a < (int64_t)b : a < b; }
static int t_gt(int t, uint64_t a, uint64_t b) { return t_lt(t, b, a); }
static int t_le(int t, uint64_t a, uint64_t b) { return !t_lt(t, b, a); }
static int t_ge(int t, uint64_t a, uint64_t b) { return !t_lt(t, a, b); }

static inline void setbit(uint8_t* bits, size_t i) { bits[i >> 3] |= (uint8_t)(1u << (i & 7)); }
static inline void clrbit(uint8_t* bits, size_t i) { bits[i >> 3] &= (uint8_t)~(1u << (i & 7)); }

/* ------------------------------------------------------------------ Load
 * internal/encode/int.go:14-33,109-115 (LoadInt: byte 0 = ContainerType) and the
 * per-container Load methods cited below. */
long ko_container_load(int type, const uint8_t* buf, size_t len, ko_container** out) {
    if (len == 0) return -1;
    ko_container* c = (ko_container*)calloc(1, sizeof(*c));
    c->ctype = buf[0];
    c->type = type;
    const uint8_t* p = buf + 1;
    uint64_t v;
    switch (c->ctype) {
    case KO_TCONST: /* int_const.go:73-84: uv(Val) uv(N) */
        p += ko_uvarint(p, &v); c->val = ext(type, v);
        p += ko_uvarint(p, &v); c->n = v;
        break;
    case KO_TDELTA: /* int_delta.go:77-91: uv(For) uv(Delta) uv(N) */
        p += ko_uvarint(p, &v); c->val = ext(type, v);
        p += ko_uvarint(p, &v); c->delta = ext(type, v);
        p += ko_uvarint(p, &v); c->n = v;
        break;
    case KO_TBITPACK: /* int_bitpack.go:100-124: uv(For) uv(Log2) uv(N) packed[EstimateSize] */
        p += ko_uvarint(p, &v); c->val = ext(type, v);
        p += ko_uvarint(p, &v); c->log2 = (int)v;
        p += ko_uvarint(p, &v); c->n = v;
        c->payload = p; c->payload_len = ko_bitpack_size(c->log2, c->n);
        p += c->payload_len;
        break;
    case KO_TRAW:      /* int_raw.go:82-94: uv(N) N*sizeof(T) */
    case KO_TFLOATRAW: /* float_raw.go:73-89 */
        p += ko_uvarint(p, &v); c->n = v;
        c->payload = p; c->payload_len = c->n * ko_type_size(type);
        p += c->payload_len;
        break;
    case KO_TS8B: /* int_s8b.go:96-115: uv(For) uv(N) uv(len) words */
        p += ko_uvarint(p, &v); c->val = ext(type, v);
        p += ko_uvarint(p, &v); c->n = v;
        p += ko_uvarint(p, &v); c->payload = p; c->payload_len = v;
        p += v;
        break;
    case KO_TDICT: { /* int_dict.go:82-99: <Dict of T> <Codes of uint16> */
        long k = ko_container_load(type, p, len - (size_t)(p - buf), &c->child[0]);
        if (k < 0) { ko_container_free(c); return -1; }
        p += k;
        k = ko_container_load(KO_U16, p, len - (size_t)(p - buf), &c->child[1]);
        if (k < 0) { ko_container_free(c); return -1; }
        p += k;
        c->n = c->child[1]->n;
        break;
    }
    case KO_TRUNEND: { /* int_runend.go:90-112: <Values of T> <Ends of uint32>; n = last(Ends)+1 */
        long k = ko_container_load(type, p, len - (size_t)(p - buf), &c->child[0]);
        if (k < 0) { ko_container_free(c); return -1; }
        p += k;
        k = ko_container_load(KO_U32, p, len - (size_t)(p - buf), &c->child[1]);
        if (k < 0) { ko_container_free(c); return -1; }
        p += k;
        c->n = c->child[1]->n ? (size_t)ko_container_get(c->child[1], c->child[1]->n - 1) + 1 : 0;
        break;
    }
    case KO_TFLOATALP: { /* float_alp.go:122-165 */
        long k = ko_alp_load(c, buf, len);
        if (k < 0) { ko_container_free(c); return -1; }
        p = buf + k;
        break;
    }
    case KO_TFLOATALPRD: { /* float_alprd.go:86-107 */
        long k = ko_alprd_load(c, buf, len);
        if (k < 0) { ko_container_free(c); return -1; }
        p = buf + k;
        break;
    }
    default:
        free(c);
        return -1;
    }
    if ((size_t)(p - buf) > len) { ko_container_free(c); return -1; }
    *out = c;
    return (long)(p - buf);
}

void ko_container_free(ko_container* c) {
    if (!c) return;
    ko_container_free(c->child[0]);
    ko_container_free(c->child[1]);
    ko_container_free(c->child[2]);
    free(c);
}

static uint64_t raw_get(const ko_container* c, size_t i) {
    const uint8_t* p = c->payload + i * ko_type_size(c->type);
    switch (c->type) {
    case KO_I64: case KO_U64: case KO_F64: { uint64_t v; memcpy(&v, p, 8); return v; }
    case KO_I32: { int32_t v; memcpy(&v, p, 4); return (uint64_t)(int64_t)v; }
    case KO_U32: case KO_F32: { uint32_t v; memcpy(&v, p, 4); return v; }
    case KO_I16: { int16_t v; memcpy(&v, p, 2); return (uint64_t)(int64_t)v; }
    case KO_U16: { uint16_t v; memcpy(&v, p, 2); return v; }
    case KO_I8: return (uint64_t)(int64_t)(int8_t)p[0];
    case KO_U8: return p[0];
    }
    return 0;
}

/* run index of row i: first k with Ends[k] >= i (Ends are inclusive, int_runend.go:179-200) */
static size_t run_of(const ko_container* ends, size_t i) {
    size_t lo = 0, hi = ends->n;
    while (lo < hi) { size_t m = (lo + hi) / 2; if (ko_container_get(ends, m) >= i) hi = m; else lo = m + 1; }
    return lo;
}

/* Container.Get(i) — int_const.go:86, int_delta.go:93-95, int_bitpack.go:126-131,
 * int_raw.go:96, int_dict.go:101-103, int_runend.go:114-120 */
uint64_t ko_container_get(const ko_container* c, size_t i) {
    switch (c->ctype) {
    case KO_TFLOATALP: return ko_alp_get(c, i);
    case KO_TFLOATALPRD: return ko_alprd_get(c, i);
    case KO_TCONST: return c->val;
    case KO_TDELTA: return ext(c->type, (uint64_t)i * c->delta + c->val);
    case KO_TBITPACK: {
        uint64_t v = ko_bitpack_value((const uint64_t*)c->payload, c->payload_len / 8, i, c->log2, 0);
        return ext(c->type, ext(c->type, v) + c->val); /* T(word&mask) + minv */
    }
    case KO_TRAW: case KO_TFLOATRAW: return raw_get(c, i);
    case KO_TDICT: return ko_container_get(c->child[0], (size_t)ko_container_get(c->child[1], i));
    case KO_TRUNEND: return ko_container_get(c->child[0], run_of(c->child[1], i));
    case KO_TS8B: {
        uint64_t* tmp = (uint64_t*)malloc((c->n + 128) * 8);
        ko_s8b_decode(tmp, c->n + 128, (const uint64_t*)c->payload, c->payload_len / 8, 0);
        uint64_t v = ext(c->type, ext(c->type, tmp[i]) + c->val);
        free(tmp);
        return v;
    }
    }
    return 0;
}

/* AppendTo(dst, nil) */
void ko_container_decode(const ko_container* c, uint64_t* dst) {
    switch (c->ctype) {
    case KO_TFLOATALP: ko_alp_decode_all(c, dst); return;
    case KO_TFLOATALPRD: ko_alprd_decode_all(c, dst); return;
    case KO_TS8B: {
        uint64_t* tmp = (uint64_t*)malloc((c->n + 128) * 8);
        ko_s8b_decode(tmp, c->n + 128, (const uint64_t*)c->payload, c->payload_len / 8, 0);
        for (size_t i = 0; i < c->n; i++) dst[i] = ext(c->type, ext(c->type, tmp[i]) + c->val);
        free(tmp);
        return;
    }
    case KO_TDICT: {
        size_t dl = c->child[0]->n;
        uint64_t* d = (uint64_t*)malloc((dl ? dl : 1) * 8);
        ko_container_decode(c->child[0], d);
        uint64_t* codes = (uint64_t*)malloc((c->n ? c->n : 1) * 8);
        ko_container_decode(c->child[1], codes);
        for (size_t i = 0; i < c->n; i++) dst[i] = d[codes[i]];
        free(d); free(codes);
        return;
    }
    case KO_TRUNEND: {
        size_t nr = c->child[1]->n;
        uint64_t* vals = (uint64_t*)malloc((nr ? nr : 1) * 8);
        uint64_t* ends = (uint64_t*)malloc((nr ? nr : 1) * 8);
        ko_container_decode(c->child[0], vals);
        ko_container_decode(c->child[1], ends);
        size_t i = 0;
        for (size_t r = 0; r < nr; r++) for (; i <= ends[r]; i++) dst[i] = vals[r];
        free(vals); free(ends);
        return;
    }
    default:
        for (size_t i = 0; i < c->n; i++) dst[i] = ko_container_get(c, i);
    }
}

/* -------------------------------------------------------------- matchers */

/* raw block bytes → cmp kernels: int_raw.go:122-337, float_raw.go:116-207 */
static void match_raw(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    ko_cmp(c->type, op, c->payload, c->n, a, b, bits);
}

/* int_const.go:133-173 */
static void match_const(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    int t = c->type, r = 0;
    switch (op) {
    case KO_EQ: r = c->val == a; break;
    case KO_NE: r = c->val != a; break;
    case KO_LT: r = t_lt(t, c->val, a); break;
    case KO_LE: r = t_le(t, c->val, a); break;
    case KO_GT: r = t_gt(t, c->val, a); break;
    case KO_GE: r = t_ge(t, c->val, a); break;
    case KO_RG: r = t_ge(t, c->val, a) && t_le(t, c->val, b); break;
    }
    if (r) ko_bitset_one(bits, c->n);
}

/* int_bitpack.go:163-247: `val < For` pre-checks, then compare in the min-FOR domain */
static void match_bitpack(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    int t = c->type;
    const uint64_t* packed = (const uint64_t*)c->payload;
    if (op == KO_RG) {
        if (t_lt(t, b, c->val)) return;
        if (t_lt(t, a, c->val)) a = c->val;
        a = ext(t, a - c->val); b = ext(t, b - c->val);
        ko_bitpack_cmp(KO_RG, packed, c->log2, a, b, c->n, bits);
        return;
    }
    if (t_lt(t, a, c->val)) {
        if (op == KO_NE || op == KO_GT || op == KO_GE) ko_bitset_one(bits, c->n);
        return;
    }
    a = ext(t, a - c->val);
    ko_bitpack_cmp(op, packed, c->log2, a, 0, c->n, bits);
    /* NE/GT/GE write ^cmp over full 64-row groups only, so the tail is already clean */
}

/* Go semantics helpers: truncated int64 division (guarding the one trapping case) */
static int64_t sdiv(int64_t a, int64_t b) { return (b == -1) ? (int64_t)(0 - (uint64_t)a) : a / b; }
static int64_t smod(int64_t a, int64_t b) { return (b == -1) ? 0 : a % b; }

/* int_delta.go:149-449 — closed-form index arithmetic + SetRange.  Restated branch by
 * branch, including the int64-space arithmetic (`v64 := int64(val) - int64(c.For)`). */
static void match_delta(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    int t = c->type;
    int64_t N = (int64_t)c->n;
    if (N == 0) return;
    int dpos = is_signed(t) ? (int64_t)c->delta > 0 : c->delta > 0;
    int64_t d64 = (int64_t)c->delta;
    int64_t v64 = (int64_t)a - (int64_t)c->val;

    if (op == KO_EQ || op == KO_NE) { /* :149-197 */
        if (dpos ? t_lt(t, a, c->val) : t_gt(t, a, c->val)) {
            if (op == KO_NE) ko_bitset_one(bits, c->n);
            return;
        }
        uint64_t val = ext(t, a - c->val); /* may wrap */
        if (op == KO_NE) ko_bitset_one(bits, c->n);
        int divisible; int64_t q;
        if (is_signed(t)) { divisible = smod((int64_t)val, d64) == 0; q = sdiv((int64_t)val, d64); }
        else { divisible = val % c->delta == 0; q = (int64_t)(val / c->delta); }
        if (op == KO_EQ) {
            if (divisible && q >= 0 && q < N) setbit(bits, (size_t)q);
        } else {
            if ((c->delta == 1 || divisible) && q >= 0 && q < N) clrbit(bits, (size_t)q);
        }
        return;
    }

    switch (op) {
    case KO_LT: /* :199-249 */
        if (dpos) {
            if (t_lt(t, a, c->val)) return;
            if (d64 * (N - 1) < v64) { ko_bitset_one(bits, c->n); return; }
            int64_t n = sdiv(v64, d64);
            if (smod(v64, d64) == 0) n--;
            ko_bitset_set_range(bits, c->n, 0, n);
        } else {
            if (t_gt(t, a, c->val)) { ko_bitset_one(bits, c->n); return; }
            if (d64 * (N - 1) >= v64) return;
            int64_t n = sdiv(v64, d64) + 1;
            ko_bitset_set_range(bits, c->n, n, N - 1);
        }
        return;
    case KO_LE: /* :251-296 */
        if (dpos) {
            if (t_lt(t, a, c->val)) return;
            if (d64 * (N - 1) < v64) { ko_bitset_one(bits, c->n); return; }
            ko_bitset_set_range(bits, c->n, 0, sdiv(v64, d64));
        } else {
            if (t_ge(t, a, c->val)) { ko_bitset_one(bits, c->n); return; }
            if (d64 * (N - 1) > v64) return;
            int64_t n = sdiv(v64, d64);
            if (smod(v64, d64) != 0) n++;
            ko_bitset_set_range(bits, c->n, n, N - 1);
        }
        return;
    case KO_GT: /* :298-349 */
        if (dpos) {
            if (t_lt(t, a, c->val)) { ko_bitset_one(bits, c->n); return; }
            if (d64 * (N - 1) < v64) return;
            ko_bitset_set_range(bits, c->n, sdiv(v64, d64) + 1, N - 1);
        } else {
            if (t_gt(t, a, c->val)) return;
            if (d64 * (N - 1) > v64) { ko_bitset_one(bits, c->n); return; }
            int64_t n = sdiv(v64, d64);
            if (smod(v64, d64) == 0) n--;
            ko_bitset_set_range(bits, c->n, 0, n);
        }
        return;
    case KO_GE: /* :351-398 */
        if (dpos) {
            if (t_le(t, a, c->val)) { ko_bitset_one(bits, c->n); return; }
            if (d64 * (N - 1) < v64) return;
            int64_t n = sdiv(v64, d64);
            if (smod(v64, d64) > 0) n++;
            ko_bitset_set_range(bits, c->n, n, N - 1);
        } else {
            if (t_gt(t, a, c->val)) return;
            if (d64 * (N - 1) > v64) { ko_bitset_one(bits, c->n); return; }
            ko_bitset_set_range(bits, c->n, 0, sdiv(v64, d64));
        }
        return;
    case KO_RG: { /* :400-449 */
        int64_t a64 = (int64_t)a - (int64_t)c->val, b64 = (int64_t)b - (int64_t)c->val;
        if (dpos) {
            if (t_lt(t, b, c->val) || a64 > d64 * (N - 1)) return;
            int64_t na = sdiv(a64, d64), nb = sdiv(b64, d64);
            if (smod(a64, d64) != 0) na++;
            if (nb > N - 1) nb = N - 1;
            ko_bitset_set_range(bits, c->n, na, nb);
        } else {
            if (t_gt(t, a, c->val) || b64 < d64 * (N - 1)) return;
            int64_t na = sdiv(a64, d64), nb = sdiv(b64, d64);
            if (smod(b64, d64) != 0) nb++;
            if (nb > N - 1) nb = N - 1;
            ko_bitset_set_range(bits, c->n, nb, na);
        }
        return;
    }
    }
}

/* sort.Search(l, func(i) bool { return Dict.Get(i) >= val }) resp. > val */
static size_t dict_search(const uint64_t* d, size_t l, int t, uint64_t val, int strict) {
    size_t lo = 0, hi = l;
    while (lo < hi) {
        size_t m = (lo + hi) / 2;
        int ok = strict ? t_gt(t, d[m], val) : t_ge(t, d[m], val);
        if (ok) hi = m; else lo = m + 1;
    }
    return lo;
}

/* int_dict.go:181-359: translate the value predicate into a code predicate */
static void match_dict(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    int t = c->type;
    size_t l = c->child[0]->n;
    if (l == 0) return;
    uint64_t* d = (uint64_t*)malloc(l * 8);
    ko_container_decode(c->child[0], d);
    uint64_t first = d[0], last = d[l - 1];
    const ko_container* codes = c->child[1];
    size_t idx;
    switch (op) {
    case KO_EQ:
        if (t_lt(t, a, first) || t_gt(t, a, last)) break;
        idx = dict_search(d, l, t, a, 0);
        if (idx == l || d[idx] != a) break;
        ko_container_match(codes, KO_EQ, idx, 0, bits);
        break;
    case KO_NE:
        if (t_lt(t, a, first) || t_gt(t, a, last)) { ko_bitset_one(bits, c->n); break; }
        idx = dict_search(d, l, t, a, 0);
        if (idx == l || d[idx] != a) { ko_bitset_one(bits, c->n); break; }
        ko_container_match(codes, KO_NE, idx, 0, bits);
        break;
    case KO_LT:
        if (t_lt(t, a, first)) break;
        if (t_gt(t, a, last)) { ko_bitset_one(bits, c->n); break; }
        idx = dict_search(d, l, t, a, 0);
        if (idx == l) idx--;
        ko_container_match(codes, KO_LT, idx, 0, bits);
        break;
    case KO_LE:
        if (t_lt(t, a, first)) break;
        if (t_ge(t, a, last)) { ko_bitset_one(bits, c->n); break; }
        idx = dict_search(d, l, t, a, 0);
        if (idx == l || t_lt(t, a, d[idx])) idx--;
        ko_container_match(codes, KO_LE, idx, 0, bits);
        break;
    case KO_GT:
        if (t_lt(t, a, first)) { ko_bitset_one(bits, c->n); break; }
        if (t_ge(t, a, last)) break;
        idx = dict_search(d, l, t, a, 1);
        ko_container_match(codes, KO_GE, idx, 0, bits);
        break;
    case KO_GE:
        if (t_lt(t, a, first)) { ko_bitset_one(bits, c->n); break; }
        if (t_gt(t, a, last)) break;
        idx = dict_search(d, l, t, a, 0);
        ko_container_match(codes, KO_GE, idx, 0, bits);
        break;
    case KO_RG: {
        if (t_lt(t, b, first) || t_gt(t, a, last)) break;
        if (t_le(t, a, first) && t_ge(t, b, last)) { ko_bitset_one(bits, c->n); break; }
        size_t ai = dict_search(d, l, t, a, 0), bi = dict_search(d, l, t, b, 0);
        uint64_t v = d[ai];
        if (ai == bi && v != a && v != b) break;
        if (bi == l || d[bi] != b) bi--;
        ko_container_match(codes, KO_RG, ai, bi, bits);
        break;
    }
    }
    free(d);
}

/* int_runend.go:296-318 (applyMatch) */
static void runend_apply(const ko_container* c, const uint8_t* vbits, uint8_t* bits) {
    size_t nr = c->child[0]->n;
    int64_t cnt = ko_bitset_popcount(vbits, nr);
    if (cnt == 0) return;
    if ((size_t)cnt == nr) { ko_bitset_one(bits, c->n); return; }
    uint64_t* ends = (uint64_t*)malloc(nr * 8);
    ko_container_decode(c->child[1], ends);
    for (size_t k = 0; k < nr; k++) {
        if (!(vbits[k >> 3] >> (k & 7) & 1)) continue;
        int64_t start = k > 0 ? (int64_t)ends[k - 1] + 1 : 0;
        ko_bitset_set_range(bits, c->n, start, (int64_t)ends[k]);
    }
    free(ends);
}

/* s8b: int_s8b.go:162-235 pre-checks; the fused s8b kernels (s8b/generic/cmp.go:13-120)
 * produce exactly the scalar predicate over the decoded min-FOR values */
static void match_s8b(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    int t = c->type;
    if (op == KO_RG) {
        if (t_lt(t, b, c->val)) return;
        if (t_lt(t, a, c->val)) a = c->val;
        a -= c->val; b -= c->val;
    } else if (t_lt(t, a, c->val)) {
        if (op == KO_NE || op == KO_GT) { ko_bitset_one(bits, c->n); return; }
        if (op == KO_GE) a = c->val; else return;
        a -= c->val;
    } else a -= c->val;
    uint64_t* tmp = (uint64_t*)malloc((c->n + 128) * 8);
    ko_s8b_decode(tmp, c->n + 128, (const uint64_t*)c->payload, c->payload_len / 8, 0);
    ko_cmp(KO_U64, op, tmp, c->n, a, b, bits);
    free(tmp);
}

void ko_container_match(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    if (!is_float(c->type)) { a = ext(c->type, a); b = ext(c->type, b); }
    switch (c->ctype) {
    case KO_TCONST: match_const(c, op, a, b, bits); return;
    case KO_TDELTA: match_delta(c, op, a, b, bits); return;
    case KO_TBITPACK: match_bitpack(c, op, a, b, bits); return;
    case KO_TRAW: case KO_TFLOATRAW: match_raw(c, op, a, b, bits); return;
    case KO_TDICT: match_dict(c, op, a, b, bits); return;
    case KO_TS8B: match_s8b(c, op, a, b, bits); return;
    case KO_TFLOATALP: ko_alp_match(c, op, a, b, bits); return;
    case KO_TFLOATALPRD: ko_alprd_match(c, op, a, b, bits); return;
    case KO_TRUNEND: { /* int_runend.go:224-294 */
        size_t nr = c->child[0]->n;
        uint8_t* vbits = (uint8_t*)calloc((nr + 7) / 8 + 8, 1);
        ko_container_match(c->child[0], op, a, b, vbits);
        runend_apply(c, vbits, bits);
        free(vbits);
        return;
    }
    }
}

static int set_contains(const uint64_t* set, size_t n, uint64_t v) {
    size_t lo = 0, hi = n;
    while (lo < hi) { size_t m = (lo + hi) / 2; if (set[m] < v) lo = m + 1; else hi = m; }
    return lo < n && set[lo] == v;
}

/* MatchInSet / MatchNotInSet with mask == nil: int_const.go:175-187, int_delta.go:451-500,
 * int_raw.go:339-380, int_bitpack.go:249-291, int_dict.go:361-398 (+translateSet :400-470),
 * int_runend.go:282-294.  `set.Contains(uint64(v))`: v sign-extended. */
void ko_container_match_set(const ko_container* c, int negate, const uint64_t* set, size_t nset, uint8_t* bits) {
    switch (c->ctype) {
    case KO_TCONST:
        if (set_contains(set, nset, c->val) != negate) ko_bitset_one(bits, c->n);
        return;
    case KO_TDICT: {
        size_t l = c->child[0]->n;
        uint64_t* d = (uint64_t*)malloc((l ? l : 1) * 8);
        ko_container_decode(c->child[0], d);
        uint64_t* cset = (uint64_t*)malloc((l ? l : 1) * 8);
        size_t nc = 0;
        for (size_t i = 0; i < l; i++) if (set_contains(set, nset, d[i])) cset[nc++] = i; /* ascending codes */
        if (nc == 0) { if (negate) ko_bitset_one(bits, c->n); }
        else if (nc == 1) ko_container_match(c->child[1], negate ? KO_NE : KO_EQ, cset[0], 0, bits);
        else ko_container_match_set(c->child[1], negate, cset, nc, bits);
        free(d); free(cset);
        return;
    }
    case KO_TRUNEND: {
        size_t nr = c->child[0]->n;
        uint8_t* vbits = (uint8_t*)calloc((nr + 7) / 8 + 8, 1);
        ko_container_match_set(c->child[0], negate, set, nset, vbits);
        runend_apply(c, vbits, bits);
        free(vbits);
        return;
    }
    default: {
        uint64_t* vals = (uint64_t*)malloc((c->n ? c->n : 1) * 8);
        ko_container_decode(c, vals);
        for (size_t i = 0; i < c->n; i++)
            if (set_contains(set, nset, vals[i]) != negate) setbit(bits, i);
        free(vals);
    }
    }
}

/* ------------------------------------------------------------------ Store */

size_t ko_store_bound(int type, size_t n) { (void)type; return 64 + n * 8 + (n / 4 + 2) * 8 * 3; }

/* int_const.go:67-71 */
size_t ko_store_const(uint8_t* dst, uint64_t val, size_t n) {
    uint8_t* p = dst; *p++ = KO_TCONST;
    p += ko_put_uvarint(p, val); p += ko_put_uvarint(p, n);
    return (size_t)(p - dst);
}
/* int_delta.go:70-75 */
size_t ko_store_delta(uint8_t* dst, uint64_t for_, uint64_t delta, size_t n) {
    uint8_t* p = dst; *p++ = KO_TDELTA;
    p += ko_put_uvarint(p, for_); p += ko_put_uvarint(p, delta); p += ko_put_uvarint(p, n);
    return (size_t)(p - dst);
}
/* int_raw.go:73-80, float_raw.go:64-71 */
size_t ko_store_raw(uint8_t* dst, int type, const uint64_t* vals, size_t n) {
    uint8_t* p = dst; *p++ = is_float(type) ? KO_TFLOATRAW : KO_TRAW;
    p += ko_put_uvarint(p, n);
    int sz = ko_type_size(type);
    for (size_t i = 0; i < n; i++) { memcpy(p, &vals[i], (size_t)sz); p += sz; } /* little endian host */
    return (size_t)(p - dst);
}

static void minmax(int type, const uint64_t* vals, size_t n, uint64_t* mn, uint64_t* mx) {
    *mn = *mx = n ? vals[0] : 0;
    for (size_t i = 1; i < n; i++) {
        if (t_lt(type, vals[i], *mn)) *mn = vals[i];
        if (t_gt(type, vals[i], *mx)) *mx = vals[i];
    }
}

/* int_bitpack.go:92-98 + Encode :147-158 (For = min, Log2 = Log2Range(min,max)) */
size_t ko_store_bitpack(uint8_t* dst, int type, const uint64_t* vals, size_t n) {
    uint64_t mn, mx; minmax(type, vals, n, &mn, &mx);
    int log2 = ko_log2range(type, mn, mx);
    uint8_t* p = dst; *p++ = KO_TBITPACK;
    p += ko_put_uvarint(p, mn); p += ko_put_uvarint(p, (uint64_t)log2); p += ko_put_uvarint(p, n);
    size_t sz = ko_bitpack_size(log2, n);
    uint64_t* tmp = (uint64_t*)calloc(sz / 8 + 1, 8);
    ko_bitpack_encode(tmp, vals, n, log2, mn);
    memcpy(p, tmp, sz); free(tmp);
    return (size_t)(p - dst) + sz;
}

static int cmp_u64_t(const void* a, const void* b, void* tp) {
    int t = *(int*)tp; uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return t_lt(t, x, y) ? -1 : (t_lt(t, y, x) ? 1 : 0);
}

/* int_dict.go:76-80 + Encode: Dict = sorted unique values, Codes = uint16 positions,
 * both children encoded recursively (EncodeInt at lvl-1) */
size_t ko_store_dict(uint8_t* dst, int type, const uint64_t* vals, size_t n) {
    uint64_t* d = (uint64_t*)malloc((n ? n : 1) * 8);
    memcpy(d, vals, n * 8);
    qsort_r(d, n, 8, cmp_u64_t, &type);
    size_t l = 0;
    for (size_t i = 0; i < n; i++) if (i == 0 || d[i] != d[l - 1]) d[l++] = d[i];
    uint64_t* codes = (uint64_t*)malloc((n ? n : 1) * 8);
    for (size_t i = 0; i < n; i++) codes[i] = dict_search(d, l, type, vals[i], 0);
    uint8_t* p = dst; *p++ = KO_TDICT;
    p += ko_store_best(p, type, d, l, 1);
    p += ko_store_best(p, KO_U16, codes, n, 1);
    free(d); free(codes);
    return (size_t)(p - dst);
}

/* int_runend.go:84-88 + Encode :140-177 (values[p], ends[p] = inclusive last index) */
size_t ko_store_runend(uint8_t* dst, int type, const uint64_t* vals, size_t n) {
    uint64_t* rv = (uint64_t*)malloc((n ? n : 1) * 8);
    uint64_t* re = (uint64_t*)malloc((n ? n : 1) * 8);
    size_t r = 0;
    for (size_t i = 0; i < n; i++) {
        if (i == 0 || vals[i] != vals[i - 1]) { rv[r] = vals[i]; r++; }
        re[r - 1] = i;
    }
    uint8_t* p = dst; *p++ = KO_TRUNEND;
    p += ko_store_best(p, type, rv, r, 1);
    p += ko_store_best(p, KO_U32, re, r, 1);
    free(rv); free(re);
    return (size_t)(p - dst);
}

/* int_s8b.go:87-94 */
size_t ko_store_s8b(uint8_t* dst, int type, const uint64_t* vals, size_t n) {
    uint64_t mn, mx; minmax(type, vals, n, &mn, &mx);
    uint64_t* words = (uint64_t*)malloc((n + 1) * 8);
    size_t nb = ko_s8b_encode(words, vals, n, mn);
    uint8_t* p = dst; *p++ = KO_TS8B;
    p += ko_put_uvarint(p, mn); p += ko_put_uvarint(p, n); p += ko_put_uvarint(p, nb);
    memcpy(p, words, nb); free(words);
    return (size_t)(p - dst) + nb;
}

/* Scheme choice in the spirit of context.go:257-293 (EligibleIntSchemes) with the cost
 * model reduced to its ordering: const (1 run) → delta (constant positive step, n > 2)
 * → run-end (avg run >= 4, lvl > 0) → dict (<= 32768 uniques and cheaper than bitpack,
 * lvl > 0) → bitpack (width shrinks) → raw.  lvl mirrors MAX_LEVEL nesting (context.go:29). */
size_t ko_store_best(uint8_t* dst, int type, const uint64_t* vals, size_t n, int lvl) {
    if (is_float(type)) return ko_store_raw(dst, type, vals, n);
    if (n == 0) return ko_store_raw(dst, type, vals, n);
    size_t runs = 1; int const_delta = n > 1; uint64_t delta = n > 1 ? vals[1] - vals[0] : 0;
    for (size_t i = 1; i < n; i++) {
        if (vals[i] != vals[i - 1]) runs++;
        if (vals[i] - vals[i - 1] != delta) const_delta = 0;
    }
    if (runs == 1) return ko_store_const(dst, vals[0], n);
    int dpos = is_signed(type) ? (int64_t)ext(type, delta) > 0 : (ext(type, delta) > 0 && vals[1] > vals[0]);
    if (const_delta && dpos && n > 2) return ko_store_delta(dst, vals[0], ext(type, delta), n);
    uint64_t mn, mx; minmax(type, vals, n, &mn, &mx);
    int usebits = ko_log2range(type, mn, mx), phybits = ko_type_size(type) * 8;
    if (lvl > 0 && n / runs >= 4) return ko_store_runend(dst, type, vals, n);
    if (lvl > 0 && n >= 64) {
        /* cardinality via sort (test helper; sizes are small) */
        uint64_t* d = (uint64_t*)malloc(n * 8); memcpy(d, vals, n * 8);
        qsort_r(d, n, 8, cmp_u64_t, &type);
        size_t uniq = 1; for (size_t i = 1; i < n; i++) if (d[i] != d[i - 1]) uniq++;
        free(d);
        if (uniq <= 32768) {
            int cbits = uniq > 1 ? 64 - __builtin_clzll((uint64_t)uniq - 1) : 0;
            size_t dict_cost = uniq * (size_t)phybits + n * (size_t)cbits;
            size_t bp_cost = n * (size_t)usebits;
            if (dict_cost < bp_cost) return ko_store_dict(dst, type, vals, n);
        }
    }
    if (usebits < phybits) return ko_store_bitpack(dst, type, vals, n);
    return ko_store_raw(dst, type, vals, n);
}

/* ------------------------------------------------------------------ simple8b
 * internal/encode/s8b/generic/encode.go:61 (maxNPerSelector), :63-115 tables */
static const int S8_N[16] = {128, 128, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
static const int S8_BITS[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

static int s8_maxn_for_bits(int b) {
    static const int t[61] = {60, 60, 30, 20, 15, 12, 10, 8, 7, 6, 6, 5, 5, 4, 4, 4, 3, 3, 3, 3, 3,
                              2, 2, 2, 2, 2, 2, 2, 2, 2, 2,
                              1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};
    return t[b];
}
static int s8_code_for_bits(int b) {
    if (b <= 1) return 2; if (b <= 8) return b + 1; if (b <= 10) return 10; if (b <= 12) return 11;
    if (b <= 15) return 12; if (b <= 20) return 13; if (b <= 30) return 14; return 15;
}

/* encode.go:117-210 (Encode): incremental packing, every code word is full; selectors
 * 0/1 = 128 zeros / 128 ones (post min-FOR).  Returns bytes written. */
size_t ko_s8b_encode(uint64_t* dst, const uint64_t* vals, size_t n, uint64_t minv) {
    uint64_t mx = 0;
    for (size_t i = 0; i < n; i++) if (vals[i] - minv > mx) mx = vals[i] - minv;
    int maxlog2 = mx ? 64 - __builtin_clzll(mx) : 0;
    if (maxlog2 > 60) return 0;
    size_t i = 0, j = 0;
    while (i < n) {
        size_t nleft = n - i;
        if (nleft >= 128 && vals[i] - minv <= 1) {
            int zero = 1, one = 1;
            for (int k = 0; k < 128; k++) { uint64_t d = vals[i + k] - minv; if (d != 0) zero = 0; if (d != 1) one = 0; }
            if (zero) { dst[j++] = 0; i += 128; continue; }
            if (one) { dst[j++] = 1ull << 60; i += 128; continue; }
        }
        int maxN = 60, used = 1, full = 0; size_t k = 0; uint64_t maxSeen = 1;
        while (k < nleft) {
            uint64_t v = vals[i + k] - minv;
            if (v > maxSeen) {
                maxSeen = v; used = 64 - __builtin_clzll(v); maxN = s8_maxn_for_bits(used);
                if ((int)k > maxN) break;
                if (maxN > 5 && used == maxlog2) { k = (size_t)maxN < nleft ? (size_t)maxN : nleft; full = (size_t)maxN <= nleft; break; }
            }
            k++;
            if ((int)k == maxN) { full = 1; break; }
        }
        int sel = s8_code_for_bits(used);
        if (!full) {
            while (sel < 15 && (int)k < S8_N[sel]) sel++;
            if ((int)k > S8_N[sel]) k = (size_t)S8_N[sel];
        }
        /* note: when the incremental loop broke on an oversize value, repack what fits */
        {
            int cnt = S8_N[sel], bits = S8_BITS[sel];
            /* all cnt values must fit `bits` (guaranteed by the selector escalation) */
            uint64_t w = (uint64_t)sel << 60;
            for (int q = 0; q < cnt; q++) w |= (vals[i + q] - minv) << (q * bits);
            dst[j++] = w;
            i += (size_t)cnt;
        }
    }
    return j * 8;
}

/* decode.go:15-81 (Decode): selector in the top 4 bits, values LSB-first */
size_t ko_s8b_decode(uint64_t* dst, size_t cap, const uint64_t* words, size_t nwords, uint64_t minv) {
    size_t j = 0;
    for (size_t i = 0; i < nwords; i++) {
        uint64_t w = words[i]; int sel = (int)(w >> 60) & 0xf;
        int cnt = S8_N[sel], bits = S8_BITS[sel];
        if (j + (size_t)cnt > cap) return j;
        if (sel == 0) for (int q = 0; q < cnt; q++) dst[j++] = minv;
        else if (sel == 1) for (int q = 0; q < cnt; q++) dst[j++] = 1 + minv;
        else {
            uint64_t m = bits == 60 ? ((1ull << 60) - 1) : ((1ull << bits) - 1);
            for (int q = 0; q < cnt; q++) dst[j++] = ((w >> (q * bits)) & m) + minv;
        }
    }
    return j;
}
