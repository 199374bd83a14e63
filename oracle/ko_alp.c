/*
 * ko_alp.c — oracle (TEST INFRASTRUCTURE ONLY, see knox_oracle.h): the ALP float64 container.
 *
 * Restates internal/encode/float_alp.go (FloatAlpContainer[float64,int64]: Store :109-120, Load :122-165,
 * Get :167-170, AppendTo :172-206, Match* :238-495) and internal/encode/alp/{constants,encoder,decoder}.go.
 * Compiled with -ffp-contract=off: Go on amd64 never fuses `v*F10[e]*IF10[f] + SWEET`.
 * Float → int conversions follow Go on amd64 (CVTTSD2SQ: NaN / out of range → 0x8000000000000000); the
 * reference's matchers depend on it (e.g. MatchLess(+Inf) matches no encoded value, float_alp.go:295-330).
 */
#define _GNU_SOURCE
#include "knox_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* alp/constants.go:88-150 */
static const double F10[24] = {
    1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0, 10000000.0, 100000000.0, 1000000000.0, 10000000000.0,
    100000000000.0, 1000000000000.0, 10000000000000.0, 100000000000000.0, 1000000000000000.0, 10000000000000000.0,
    100000000000000000.0, 1000000000000000000.0, 10000000000000000000.0, 100000000000000000000.0,
    1000000000000000000000.0, 10000000000000000000000.0, 100000000000000000000000.0};
static const double IF10[21] = {
    1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001, 0.00000001, 0.000000001, 0.0000000001, 0.00000000001,
    0.000000000001, 0.0000000000001, 0.00000000000001, 0.000000000000001, 0.0000000000000001, 0.00000000000000001,
    0.000000000000000001, 0.0000000000000000001, 0.00000000000000000001};
static const double SWEET = 6755399441055744.0; /* 1<<52 + 1<<51 */

static double f64_of(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static uint64_t bits_of(double d) { uint64_t b; memcpy(&b, &d, 8); return b; }
static int64_t go_f2i(double x) {
    if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)x;
}

/* alp/encoder.go:112-125 */
int64_t ko_alp_encode_single(double v, int e, int f, int* ok) {
    int64_t enc = go_f2i((v * F10[e] * IF10[f] + SWEET) - SWEET);
    double dec = (double)enc * F10[f] * IF10[e];
    *ok = v == dec;
    return enc;
}
int64_t ko_alp_encode_above(double v, int e, int f) { return go_f2i(ceil((v * F10[e] * IF10[f] + SWEET) - SWEET)); }
int64_t ko_alp_encode_below(double v, int e, int f) { return go_f2i(floor((v * F10[e] * IF10[f] + SWEET) - SWEET)); }
/* alp/decoder.go:122-124 */
double ko_alp_decode(int64_t enc, int e, int f) { return (double)enc * F10[f] * IF10[e]; }

/* Store: Encoder.Encode (alp/encoder.go:72-110) with the given exponents (e < 0: pick the pair with the fewest
 * exceptions over a sample — the reference samples too, alp/analyze.go; any valid pair loads everywhere). */
size_t ko_store_alp(uint8_t* dst, const uint64_t* vals, size_t n, int e, int f) {
    if (e < 0) {
        /* estimated size in bits over a sample: exceptions cost value + position, the rest the FOR bit width */
        double best = 1e300; int be = 0, bf = 0;
        size_t step = n > 512 ? n / 512 : 1;
        for (int ce = 0; ce <= 18; ce++) for (int cf = 0; cf <= ce; cf++) {
            size_t bad = 0, cnt = 0; int64_t lo = INT64_MAX, hi = INT64_MIN;
            for (size_t i = 0; i < n; i += step, cnt++) {
                int ok; int64_t x = ko_alp_encode_single(f64_of(vals[i]), ce, cf, &ok);
                if (!ok) { bad++; continue; }
                if (x < lo) lo = x;
                if (x > hi) hi = x;
            }
            int w = (hi > lo) ? 64 - __builtin_clzll((uint64_t)hi - (uint64_t)lo) : 0;
            double cost = (double)bad * 96.0 + (double)(cnt - bad) * w + (bad == cnt ? 1e9 : 0);
            if (cost < best) { best = cost; be = ce; bf = cf; }
        }
        e = be; f = bf;
    }
    int64_t* enc = (int64_t*)malloc((n ? n : 1) * 8);
    uint32_t* pos = (uint32_t*)malloc((n ? n : 1) * 4);
    uint64_t* pv = (uint64_t*)malloc((n ? n : 1) * 8);
    size_t np = 0;
    int64_t mn = INT64_MAX, mx = 0;
    for (size_t i = 0; i < n; i++) {
        int ok; int64_t x = ko_alp_encode_single(f64_of(vals[i]), e, f, &ok);
        if (ok) { enc[i] = x; if (x < mn) mn = x; if (x > mx) mx = x; }
        else pos[np++] = (uint32_t)i;
    }
    for (size_t k = 0; k < np; k++) { enc[pos[k]] = mn; pv[k] = vals[pos[k]]; }   /* exceptions hold the minimum */
    uint8_t flags = (np ? 1 : 0) | ((mn > -(int64_t)(1ll << 51) && mx < (int64_t)(1ll << 51)) ? 2 : 0);
    uint8_t* p = dst; *p++ = KO_TFLOATALP;
    p += ko_put_uvarint(p, (uint64_t)e); p += ko_put_uvarint(p, (uint64_t)f); *p++ = flags;
    p += ko_store_best(p, KO_I64, (const uint64_t*)enc, n, 2);
    if (np) {
        p += ko_store_raw(p, KO_F64, pv, np);
        uint64_t* p64 = (uint64_t*)malloc(np * 8);
        for (size_t k = 0; k < np; k++) p64[k] = pos[k];
        p += ko_store_best(p, KO_U32, p64, np, 1);
        free(p64);
    }
    free(enc); free(pos); free(pv);
    return (size_t)(p - dst);
}

/* Load (float_alp.go:122-165); buf points at the type byte */
long ko_alp_load(ko_container* c, const uint8_t* buf, size_t len) {
    const uint8_t* p = buf + 1; uint64_t v;
    p += ko_uvarint(p, &v); c->alp_e = (int)v;
    p += ko_uvarint(p, &v); c->alp_f = (int)v;
    c->alp_flags = *p++;
    if (c->alp_e > 20 || c->alp_f > 23) return -1;
    long k = ko_container_load(KO_I64, p, len - (size_t)(p - buf), &c->child[0]);
    if (k < 0) return -1;
    p += k;
    if (c->alp_flags & 1) {
        k = ko_container_load(KO_F64, p, len - (size_t)(p - buf), &c->child[1]);
        if (k < 0) return -1;
        p += k;
        k = ko_container_load(KO_U32, p, len - (size_t)(p - buf), &c->child[2]);
        if (k < 0) return -1;
        p += k;
    }
    c->n = c->child[0]->n;
    return (long)(p - buf);
}

static size_t npatches(const ko_container* c) { return (c->alp_flags & 1) ? c->child[2]->n : 0; }

/* Get (float_alp.go:167-170 → Decoder.DecodeValue, alp/decoder.go:97-109) */
uint64_t ko_alp_get(const ko_container* c, size_t i) {
    size_t np = npatches(c), lo = 0, hi = np;
    while (lo < hi) { size_t m = (lo + hi) / 2; if (ko_container_get(c->child[2], m) < i) lo = m + 1; else hi = m; }
    if (lo < np && ko_container_get(c->child[2], lo) == i) return ko_container_get(c->child[1], lo);
    return bits_of(ko_alp_decode((int64_t)ko_container_get(c->child[0], i), c->alp_e, c->alp_f));
}

/* AppendTo(dst, nil) (float_alp.go:172-206 → Decoder.Decode, alp/decoder.go:127-157) */
void ko_alp_decode_all(const ko_container* c, uint64_t* dst) {
    ko_container_decode(c->child[0], dst);
    for (size_t i = 0; i < c->n; i++) dst[i] = bits_of(ko_alp_decode((int64_t)dst[i], c->alp_e, c->alp_f));
    size_t np = npatches(c);
    for (size_t k = 0; k < np; k++) dst[ko_container_get(c->child[2], k)] = ko_container_get(c->child[1], k);
}

static void setb(uint8_t* bits, size_t i, int on) {
    if (on) bits[i >> 3] |= (uint8_t)(1u << (i & 7)); else bits[i >> 3] &= (uint8_t)~(1u << (i & 7));
}

/* Match* (float_alp.go:238-495).  a, b: IEEE bit patterns. */
void ko_alp_match(const ko_container* c, int op, uint64_t ua, uint64_t ub, uint8_t* bits) {
    const double a = f64_of(ua), b = f64_of(ub);
    const int e = c->alp_e, f = c->alp_f;
    const size_t np = npatches(c);
    const ko_container *V = c->child[0], *PV = c->child[1], *PP = c->child[2];
    int ok; int64_t av, bv;
    switch (op) {
    case KO_EQ: case KO_NE: {   /* MatchEqual :238-290, MatchNotEqual :292-295 = Equal + Neg */
        int isnan = a != a;
        if (!isnan) { av = ko_alp_encode_single(a, e, f, &ok); if (ok) ko_container_match(V, KO_EQ, (uint64_t)av, 0, bits); }
        if (np) {
            size_t p0 = (size_t)ko_container_get(PP, 0);
            if (bits[p0 >> 3] & (1u << (p0 & 7))) {   /* av == replacement: undo every patch position */
                for (size_t k = 0; k < np; k++) setb(bits, (size_t)ko_container_get(PP, k), 0);
            } else {
                for (size_t k = 0; k < np; k++) {
                    double pv = f64_of(ko_container_get(PV, k));
                    if (isnan ? (pv != pv) : (pv == a)) setb(bits, (size_t)ko_container_get(PP, k), 1);
                }
            }
        }
        if (op == KO_NE) ko_bitset_neg(bits, c->n);
        return;
    }
    case KO_LT:   /* :297-330 */
        if (a != a || (isinf(a) && a < 0)) return;
        av = ko_alp_encode_single(a, e, f, &ok);
        if (ok) ko_container_match(V, KO_LT, (uint64_t)av, 0, bits);
        else ko_container_match(V, KO_LE, (uint64_t)ko_alp_encode_below(a, e, f), 0, bits);
        for (size_t k = 0; k < np; k++) setb(bits, (size_t)ko_container_get(PP, k), f64_of(ko_container_get(PV, k)) < a);
        return;
    case KO_LE:   /* :332-371 */
        if (a != a) return;
        if (isinf(a) && a > 0) { ko_bitset_one(bits, c->n); return; }
        av = ko_alp_encode_single(a, e, f, &ok);
        if (!ok) av = ko_alp_encode_below(a, e, f);
        ko_container_match(V, KO_LE, (uint64_t)av, 0, bits);
        for (size_t k = 0; k < np; k++) setb(bits, (size_t)ko_container_get(PP, k), f64_of(ko_container_get(PV, k)) <= a);
        return;
    case KO_GT:   /* :373-407 */
        if (a != a || (isinf(a) && a > 0)) return;
        av = ko_alp_encode_single(a, e, f, &ok);
        if (ok) ko_container_match(V, KO_GT, (uint64_t)av, 0, bits);
        else ko_container_match(V, KO_GE, (uint64_t)ko_alp_encode_above(a, e, f), 0, bits);
        for (size_t k = 0; k < np; k++) setb(bits, (size_t)ko_container_get(PP, k), f64_of(ko_container_get(PV, k)) > a);
        return;
    case KO_GE:   /* :409-448 */
        if (a != a) return;
        if (isinf(a) && a < 0) { ko_bitset_one(bits, c->n); return; }
        av = ko_alp_encode_single(a, e, f, &ok);
        if (!ok) av = ko_alp_encode_above(a, e, f);
        ko_container_match(V, KO_GE, (uint64_t)av, 0, bits);
        for (size_t k = 0; k < np; k++) setb(bits, (size_t)ko_container_get(PP, k), f64_of(ko_container_get(PV, k)) >= a);
        return;
    case KO_RG:   /* MatchBetween :450-490 */
        if (a != a || b != b) return;
        av = ko_alp_encode_single(a, e, f, &ok); if (!ok) av = ko_alp_encode_above(a, e, f);
        bv = ko_alp_encode_single(b, e, f, &ok); if (!ok) bv = ko_alp_encode_below(b, e, f);
        ko_container_match(V, KO_RG, (uint64_t)av, (uint64_t)bv, bits);
        for (size_t k = 0; k < np; k++) { double pv = f64_of(ko_container_get(PV, k)); setb(bits, (size_t)ko_container_get(PP, k), pv >= a && pv <= b); }
        return;
    }
}
