/*
 * ko_alprd.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h): the ALP-RD float container ("real doubles":
 * values that decimal ALP cannot represent are cut into a LEFT part — the top bits, few distinct values — and a RIGHT
 * part — the low `Shift` bits).
 *
 * Restates internal/encode/float_alprd.go (FloatAlpRdContainer[T, E]: Store :78-84 = [14]<Left: int container of
 * uint16><Right: int container of E>uv(Shift); Load :86-107; Get :109-111; AppendTo :113-131; Encode :133-175 — left
 * side a dictionary when it has at most RD_MAX_DICT_SIZE = 8 distinct values, else bit-packed; right side always
 * bit-packed; Match* :181-211 = decode chunks + the generic float compare kernels) and internal/encode/alp/rd.go
 * (split :45-76 / :125-150: left = uint16(bits >> shift), right = bits & (1<<shift - 1); DecoderRD.DecodeValue
 * :212-219: bits = uint64(left) << shift | right) for float64 (E = uint64) and float32 (E = uint32).
 */
#include "knox_oracle.h"
#include <stdlib.h>
#include <string.h>

#define RD_MAX_DICT_SIZE 8   /* alp/analyze.go:43 */

static int width_of(int type) { return type == KO_F32 ? 32 : 64; }
static int right_type(int type) { return type == KO_F32 ? KO_U32 : KO_U64; }

/* analyzeRD (alp/analyze.go:164-228) on the whole vector instead of a sample: the cut with the smallest estimated size;
 * any cut that keeps the left part within 16 bits loads everywhere */
static int pick_shift(int type, const uint64_t* vals, size_t n) {
    const int w = width_of(type);
    int best_shift = w - 16; double best = 1e300;
    for (int i = 1; i <= 16; i++) {
        const int shift = w - i;
        const uint64_t mask = (1ull << shift) - 1;
        uint64_t lmin = ~0ull, lmax = 0, rmin = ~0ull, rmax = 0;
        for (size_t k = 0; k < n; k++) {
            uint64_t v = type == KO_F32 ? (uint32_t)vals[k] : vals[k], l = v >> shift, r = v & mask;
            if (l < lmin) lmin = l; if (l > lmax) lmax = l;
            if (r < rmin) rmin = r; if (r > rmax) rmax = r;
        }
        if (!n) { lmin = lmax = rmin = rmax = 0; }
        uint8_t seen[1 << 13] = {0}; size_t uniq = 0;
        for (size_t k = 0; k < n && uniq <= RD_MAX_DICT_SIZE; k++) {
            uint64_t l = ((type == KO_F32 ? (uint32_t)vals[k] : vals[k]) >> shift) - lmin;
            if (!(seen[l >> 3] & (1u << (l & 7)))) { seen[l >> 3] |= (uint8_t)(1u << (l & 7)); uniq++; }
        }
        int lbits = lmax > lmin ? 64 - __builtin_clzll(lmax - lmin) : 0, rbits = rmax > rmin ? 64 - __builtin_clzll(rmax - rmin) : 0;
        int dbits = uniq > 1 ? 64 - __builtin_clzll(uniq - 1) : 0;
        double lcost = uniq <= RD_MAX_DICT_SIZE ? (double)n * dbits + 16.0 * uniq : (double)n * lbits;
        double cost = lcost + (double)n * rbits;
        if (cost <= best) { best = cost; best_shift = shift; }
    }
    return best_shift;
}

size_t ko_store_alprd(uint8_t* dst, int type, const uint64_t* vals, size_t n, int shift) {
    const int w = width_of(type);
    if (shift < w - 16 || shift >= w) shift = pick_shift(type, vals, n);
    uint64_t* left = (uint64_t*)malloc((n ? n : 1) * 8);
    uint64_t* right = (uint64_t*)malloc((n ? n : 1) * 8);
    const uint64_t mask = (1ull << shift) - 1;
    for (size_t k = 0; k < n; k++) {
        uint64_t v = type == KO_F32 ? (uint32_t)vals[k] : vals[k];
        left[k] = (uint16_t)(v >> shift); right[k] = v & mask;
    }
    /* distinct left values (<= 8: dictionary) */
    uint64_t seen[RD_MAX_DICT_SIZE + 1]; size_t uniq = 0;
    for (size_t k = 0; k < n && uniq <= RD_MAX_DICT_SIZE; k++) {
        size_t j = 0;
        while (j < uniq && seen[j] != left[k]) j++;
        if (j == uniq) seen[uniq++] = left[k];
    }
    uint8_t* p = dst; *p++ = KO_TFLOATALPRD;
    p += (uniq <= RD_MAX_DICT_SIZE && n >= 2) ? ko_store_dict(p, KO_U16, left, n) : ko_store_bitpack(p, KO_U16, left, n);
    p += ko_store_bitpack(p, right_type(type), right, n);
    p += ko_put_uvarint(p, (uint64_t)shift);
    free(left); free(right);
    return (size_t)(p - dst);
}

/* Load (float_alprd.go:86-107); buf points at the type byte; the cut is kept in c->log2 */
long ko_alprd_load(ko_container* c, const uint8_t* buf, size_t len) {
    const uint8_t* p = buf + 1;
    long k = ko_container_load(KO_U16, p, len - (size_t)(p - buf), &c->child[0]);
    if (k < 0) return -1;
    p += k;
    k = ko_container_load(right_type(c->type), p, len - (size_t)(p - buf), &c->child[1]);
    if (k < 0) return -1;
    p += k;
    uint64_t v; p += ko_uvarint(p, &v);
    if (v >= (uint64_t)width_of(c->type) || c->child[0]->n != c->child[1]->n) return -1;
    c->log2 = (int)v;
    c->n = c->child[0]->n;
    return (long)(p - buf);
}

/* Get → DecoderRD.DecodeValue (alp/rd.go:212-219) */
uint64_t ko_alprd_get(const ko_container* c, size_t i) {
    uint64_t l = ko_container_get(c->child[0], i), r = ko_container_get(c->child[1], i);
    if (c->type == KO_F32) return (uint32_t)(((uint32_t)l << c->log2) | (uint32_t)r);
    return (l << c->log2) | r;
}

void ko_alprd_decode_all(const ko_container* c, uint64_t* dst) {
    for (size_t i = 0; i < c->n; i++) dst[i] = ko_alprd_get(c, i);
}

/* Match* (float_alprd.go:181-211): matchIt / matchRangeIt decode chunk by chunk and run the float compare kernels */
void ko_alprd_match(const ko_container* c, int op, uint64_t a, uint64_t b, uint8_t* bits) {
    if (!c->n) return;
    if (c->type == KO_F32) {
        uint32_t* tmp = (uint32_t*)malloc(c->n * 4);
        for (size_t i = 0; i < c->n; i++) tmp[i] = (uint32_t)ko_alprd_get(c, i);
        ko_cmp(KO_F32, op, tmp, c->n, a, b, bits);
        free(tmp);
    } else {
        uint64_t* tmp = (uint64_t*)malloc(c->n * 8);
        ko_alprd_decode_all(c, tmp);
        ko_cmp(KO_F64, op, tmp, c->n, a, b, bits);
        free(tmp);
    }
}
