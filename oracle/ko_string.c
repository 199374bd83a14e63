/* ko_string.c — CPU restatement of KnoxDB's byte-string containers and their matchers (TEST INFRASTRUCTURE: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it; the product never loads it).
 *
 * Follows internal/encode/string_const.go:49-69 (Store/Load: [16] uv(N) uv(len) val), string_fixed.go:59-87
 * ([17] uv(N) uv(sz) N*sz bytes), string_compact.go:64-105 ([18] <Ofs int container of uint32> <Len container>
 * uv(len(buf)) buf; Get(i) = buf[ofs[i] : ofs[i]+len[i]]), string_dict.go:70-119 ([19] <Ofs> <Len> <Code>
 * uv(len(dict)) dict; Get(i): ptr = code[i], dict[ofs[ptr] : ofs[ptr]+len[ptr]]) and the row-by-row matchers of
 * string_match.go:13-188 (bytes.Equal / bytes.Compare; Between = !(v < from) && !(v > to), equal operands → Equal).
 * The reference has no golden bytes for these containers (string_test.go round-trips only): parity unpinned beyond
 * the Store/Load/Match function bodies. */
#include <stdlib.h>
#include <string.h>

#include "knox_oracle.h"

struct ko_str {
    int ctype;
    size_t n, sz, nbuf;
    const uint8_t* buf;             /* constant value / fixed rows / compact buffer / dictionary */
    ko_container *ofs, *len, *code; /* nested uint32 containers */
};

static size_t put_uv(uint8_t* dst, uint64_t v) { return (size_t)ko_put_uvarint(dst, v); }

/* bytes.Compare */
static int bytes_cmp(const uint8_t* a, size_t al, const uint8_t* b, size_t bl) {
    size_t m = al < bl ? al : bl;
    int c = m ? memcmp(a, b, m) : 0;
    if (c) return c < 0 ? -1 : 1;
    return al < bl ? -1 : (al > bl ? 1 : 0);
}

/* Store: rows i = bytes[offs[i] .. offs[i+1]).  kind = container id 16..19.  Returns bytes written (0: the rows do
 * not fit the scheme — constant needs equal rows, fixed needs equal lengths). */
size_t ko_store_str(int kind, const uint8_t* bytes, const uint32_t* offs, size_t n, uint8_t* dst) {
    uint8_t* p = dst;
    *p++ = (uint8_t)kind;
    if (kind == KO_TSTRCONST) {
        size_t l0 = n ? offs[1] - offs[0] : 0;
        for (size_t i = 1; i < n; i++)
            if (offs[i + 1] - offs[i] != l0 || memcmp(bytes + offs[i], bytes + offs[0], l0)) return 0;
        p += put_uv(p, n); p += put_uv(p, l0);
        memcpy(p, bytes + (n ? offs[0] : 0), l0); p += l0;
        return (size_t)(p - dst);
    }
    if (kind == KO_TSTRFIXED) {
        size_t sz = n ? offs[1] - offs[0] : 0;
        for (size_t i = 1; i < n; i++) if (offs[i + 1] - offs[i] != sz) return 0;
        p += put_uv(p, n); p += put_uv(p, sz);
        for (size_t i = 0; i < n; i++) { memcpy(p, bytes + offs[i], sz); p += sz; }
        return (size_t)(p - dst);
    }
    uint64_t* tmp = (uint64_t*)calloc((n + 1) * 3, sizeof(uint64_t));
    uint64_t *o = tmp, *l = tmp + n + 1, *c = tmp + 2 * (n + 1);
    if (kind == KO_TSTRCOMPACT) {
        for (size_t i = 0; i < n; i++) { o[i] = offs[i] - offs[0]; l[i] = offs[i + 1] - offs[i]; }
        p += ko_store_best(p, KO_U32, o, n, 3);
        p += ko_store_best(p, KO_U32, l, n, 3);
        size_t nb = n ? offs[n] - offs[0] : 0;
        p += put_uv(p, nb);
        memcpy(p, bytes + (n ? offs[0] : 0), nb); p += nb;
        free(tmp);
        return (size_t)(p - dst);
    }
    /* dictionary: unique rows in order of first occurrence */
    size_t m = 0, nb = 0;
    uint8_t* dict = (uint8_t*)malloc((n ? offs[n] - offs[0] : 0) + 1);
    for (size_t i = 0; i < n; i++) {
        size_t li = offs[i + 1] - offs[i], k;
        for (k = 0; k < m; k++)
            if (l[k] == li && !memcmp(dict + o[k], bytes + offs[i], li)) break;
        if (k == m) { o[m] = nb; l[m] = li; memcpy(dict + nb, bytes + offs[i], li); nb += li; m++; }
        c[i] = k;
    }
    p += ko_store_best(p, KO_U32, o, m, 3);
    p += ko_store_best(p, KO_U32, l, m, 3);
    p += ko_store_best(p, KO_U32, c, n, 3);
    p += put_uv(p, nb);
    memcpy(p, dict, nb); p += nb;
    free(dict); free(tmp);
    return (size_t)(p - dst);
}

long ko_str_load(const uint8_t* enc, size_t len, ko_str** out) {
    if (len == 0) return -1;
    ko_str* s = (ko_str*)calloc(1, sizeof(*s));
    const uint8_t* p = enc + 1;
    uint64_t v;
    long used;
    s->ctype = enc[0];
    switch (s->ctype) {
    case KO_TSTRCONST:
        p += ko_uvarint(p, &v); s->n = v;
        p += ko_uvarint(p, &v); s->sz = v; s->buf = p; p += v;
        break;
    case KO_TSTRFIXED:
        p += ko_uvarint(p, &v); s->n = v;
        p += ko_uvarint(p, &v); s->sz = v; s->buf = p; p += s->n * s->sz;
        break;
    case KO_TSTRCOMPACT:
        if ((used = ko_container_load(KO_U32, p, len - (size_t)(p - enc), &s->ofs)) < 0) goto bad;
        p += used; s->n = s->ofs->n;
        if ((used = ko_container_load(KO_U32, p, len - (size_t)(p - enc), &s->len)) < 0) goto bad;
        p += used;
        p += ko_uvarint(p, &v); s->nbuf = v; s->buf = p; p += v;
        break;
    case KO_TSTRDICT:
        if ((used = ko_container_load(KO_U32, p, len - (size_t)(p - enc), &s->ofs)) < 0) goto bad;
        p += used;
        if ((used = ko_container_load(KO_U32, p, len - (size_t)(p - enc), &s->len)) < 0) goto bad;
        p += used;
        if ((used = ko_container_load(KO_U32, p, len - (size_t)(p - enc), &s->code)) < 0) goto bad;
        p += used; s->n = s->code->n;
        p += ko_uvarint(p, &v); s->nbuf = v; s->buf = p; p += v;
        break;
    default: goto bad;
    }
    *out = s;
    return (long)(p - enc);
bad:
    ko_str_free(s);
    return -1;
}

void ko_str_free(ko_str* s) {
    if (!s) return;
    if (s->ofs) ko_container_free(s->ofs);
    if (s->len) ko_container_free(s->len);
    if (s->code) ko_container_free(s->code);
    free(s);
}

size_t ko_str_len(const ko_str* s) { return s->n; }

/* Get(i) */
const uint8_t* ko_str_get(const ko_str* s, size_t i, size_t* len) {
    switch (s->ctype) {
    case KO_TSTRCONST: *len = s->sz; return s->buf;
    case KO_TSTRFIXED: *len = s->sz; return s->buf + i * s->sz;
    case KO_TSTRCOMPACT: *len = (size_t)ko_container_get(s->len, i); return s->buf + ko_container_get(s->ofs, i);
    default: {
        size_t ptr = (size_t)ko_container_get(s->code, i);
        *len = (size_t)ko_container_get(s->len, ptr);
        return s->buf + ko_container_get(s->ofs, ptr);
    }
    }
}

/* Match<Op>(val[, to], bits, nil): bits pre-zeroed, ceil(n/8) bytes; op = types.FilterMode (KO_EQ … KO_RG) */
void ko_str_match(const ko_str* s, int op, const uint8_t* a, size_t al, const uint8_t* b, size_t bl, uint8_t* bits) {
    if (op == KO_RG && al == bl && !memcmp(a, b, al)) op = KO_EQ;   /* matchStringBetween: from == to → Equal */
    for (size_t i = 0; i < s->n; i++) {
        size_t vl;
        const uint8_t* v = ko_str_get(s, i, &vl);
        int hit;
        switch (op) {
        case KO_EQ: hit = vl == al && !memcmp(v, a, al); break;
        case KO_NE: hit = !(vl == al && !memcmp(v, a, al)); break;
        case KO_LT: hit = bytes_cmp(v, vl, a, al) < 0; break;
        case KO_LE: hit = bytes_cmp(v, vl, a, al) <= 0; break;
        case KO_GT: hit = bytes_cmp(v, vl, a, al) > 0; break;
        case KO_GE: hit = bytes_cmp(v, vl, a, al) >= 0; break;
        default: hit = !(bytes_cmp(v, vl, a, al) < 0) && !(bytes_cmp(v, vl, b, bl) > 0); break;
        }
        if (hit) bits[i >> 3] |= (uint8_t)(1u << (i & 7));
    }
}
