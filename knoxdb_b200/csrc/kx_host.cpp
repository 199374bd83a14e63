// kx_host.cpp — container parsing, block normalisation and leaf translation (host side).
//
// Reference behaviour followed (paths relative to the knoxdb repository):
//   wire format        internal/encode/container.go:20-55, int.go:109-115, pkg/num/varint.go:85-192
//   container headers  internal/encode/int_{const,delta,raw,bitpack,dict,runend,s8b}.go (Load), float_raw.go
//   leaf translation   internal/encode/int_bitpack.go:163-247, int_delta.go:149-449, int_dict.go:181-359,
//                      int_const.go:133-173, internal/cmp/number.go:13-243
#include "kx_host.h"
#include "kx_xxh3.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace kx {

// ---------------------------------------------------------------------------------- varint
// SQLite4-style varint of pkg/num/varint.go: 0..240 in one byte, two/three byte forms with
// an offset, then a length byte 250..255 followed by 3..8 big-endian bytes.
int put_uvarint(uint8_t* b, uint64_t x) {
    if (x <= 240) { b[0] = uint8_t(x); return 1; }
    if (x <= 2287) { x -= 240; b[0] = uint8_t(241 + (x >> 8)); b[1] = uint8_t(x); return 2; }
    if (x <= 67823) { x -= 2288; b[0] = 249; b[1] = uint8_t(x >> 8); b[2] = uint8_t(x); return 3; }
    int nb = 3;
    while (nb < 8 && (x >> (8 * nb)) != 0) nb++;
    b[0] = uint8_t(247 + nb);
    for (int i = 0; i < nb; i++) b[1 + i] = uint8_t(x >> (8 * (nb - 1 - i)));
    return nb + 1;
}

int get_uvarint(const uint8_t* b, size_t avail, uint64_t* x) {
    if (avail < 1) return 0;
    unsigned b0 = b[0];
    if (b0 <= 240) { *x = b0; return 1; }
    if (b0 <= 248) { if (avail < 2) return 0; *x = 240 + (uint64_t(b0 - 241) << 8) + b[1]; return 2; }
    if (b0 == 249) { if (avail < 3) return 0; *x = 2288 + (uint64_t(b[1]) << 8) + b[2]; return 3; }
    int nb = int(b0) - 247;
    if (avail < size_t(nb) + 1) return 0;
    uint64_t v = 0;
    for (int i = 0; i < nb; i++) v = (v << 8) | b[1 + i];
    *x = v;
    return nb + 1;
}

size_t bitpack_bytes(int log2, size_t n) { return ((size_t(log2) * n + 63) & ~size_t(63)) / 8; }

// ---------------------------------------------------------------------------------- ALP arithmetic
// internal/encode/alp/constants.go:88-150, encoder.go:112-125, decoder.go:122-124 (float64 / int64).  Built with
// -ffp-contract=off: Go on amd64 never fuses v*F10[e]*IF10[f] + SWEET.  Float → int conversions follow Go on
// amd64 (CVTTSD2SQ: NaN / out of range → 0x8000000000000000), which the reference's matchers rely on.
namespace {
const double ALP_F10[24] = {
    1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0, 10000000.0, 100000000.0, 1000000000.0, 10000000000.0,
    100000000000.0, 1000000000000.0, 10000000000000.0, 100000000000000.0, 1000000000000000.0, 10000000000000000.0,
    100000000000000000.0, 1000000000000000000.0, 10000000000000000000.0, 100000000000000000000.0,
    1000000000000000000000.0, 10000000000000000000000.0, 100000000000000000000000.0};
const double ALP_IF10[21] = {
    1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001, 0.00000001, 0.000000001, 0.0000000001, 0.00000000001,
    0.000000000001, 0.0000000000001, 0.00000000000001, 0.000000000000001, 0.0000000000000001, 0.00000000000000001,
    0.000000000000000001, 0.0000000000000000001, 0.00000000000000000001};
const double ALP_SWEET = 6755399441055744.0;   // 1<<52 + 1<<51
int64_t go_f2i(double x) {
    if (!(x >= -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return int64_t(x);
}
}  // namespace
int64_t alp_encode_single(double v, int e, int f, bool* ok) {
    int64_t enc = go_f2i((v * ALP_F10[e] * ALP_IF10[f] + ALP_SWEET) - ALP_SWEET);
    double dec = double(enc) * ALP_F10[f] * ALP_IF10[e];
    *ok = v == dec;
    return enc;
}
int64_t alp_encode_above(double v, int e, int f) { return go_f2i(std::ceil((v * ALP_F10[e] * ALP_IF10[f] + ALP_SWEET) - ALP_SWEET)); }
int64_t alp_encode_below(double v, int e, int f) { return go_f2i(std::floor((v * ALP_F10[e] * ALP_IF10[f] + ALP_SWEET) - ALP_SWEET)); }
double alp_decode(int64_t enc, int e, int f) { return double(enc) * ALP_F10[f] * ALP_IF10[e]; }

// ---------------------------------------------------------------------------------- parsing
namespace {

struct Reader {
    const uint8_t* p; size_t left; bool ok = true;
    uint64_t uv() {
        uint64_t v = 0; int k = get_uvarint(p, left, &v);
        if (!k) { ok = false; return 0; }
        p += k; left -= size_t(k); return v;
    }
    const uint8_t* take(size_t nbytes) {
        if (nbytes > left) { ok = false; return nullptr; }
        const uint8_t* q = p; p += nbytes; left -= nbytes; return q;
    }
};

const int S8_COUNT[16] = {128, 128, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
const int S8_WIDTH[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

uint64_t load_le(const uint8_t* p, int nbytes) { uint64_t v = 0; std::memcpy(&v, p, size_t(nbytes)); return v; }

uint64_t bit_field(const uint8_t* stream, size_t stream_len, size_t row, int w) {
    if (w == 0) return 0;
    size_t bit = row * size_t(w), byte = bit >> 3;
    int sh = int(bit & 7);
    // gather up to 9 bytes
    unsigned __int128 acc = 0;
    for (int i = 0; i < 9 && byte + size_t(i) < stream_len; i++) acc |= (unsigned __int128)stream[byte + size_t(i)] << (8 * i);
    return uint64_t(acc >> sh) & width_mask(w);
}

}  // namespace

long parse_container(int type, const uint8_t* buf, size_t len, std::unique_ptr<Container>& out, std::string& err, int depth) {
    if (len < 1) { err = "empty container"; return -1; }
    // the reference nests at most three levels (dictionary / run-end / ALP over a leaf container): a deeper chain is corrupt
    if (depth > 3) { err = "containers nested too deeply"; return -1; }
    auto c = std::make_unique<Container>();
    c->ctype = buf[0];
    c->type = type;
    Reader r{buf + 1, len - 1};
    switch (c->ctype) {
    case T_CONST:
        c->val = type_ext(type, r.uv()); c->n = r.uv();
        break;
    case T_DELTA:
        c->val = type_ext(type, r.uv()); c->delta = type_ext(type, r.uv()); c->n = r.uv();
        // the matchers divide by Delta (int_delta.go:149-449: Go panics on a zero divisor); the encoder only emits Delta != 0
        if (r.ok && c->delta == 0) { err = "delta container with zero delta"; return -1; }
        break;
    case T_BITPACK:
        c->val = type_ext(type, r.uv()); c->log2 = int(r.uv()); c->n = r.uv();
        if (c->log2 < 0 || c->log2 > 64) { err = "bitpack: bad width"; return -1; }
        c->payload_len = bitpack_bytes(c->log2, c->n);
        c->payload = r.take(c->payload_len);
        break;
    case T_RAW:
    case T_FLOATRAW:
        if ((c->ctype == T_FLOATRAW) != type_is_float(type)) { err = "raw container / block type mismatch"; return -1; }
        c->n = r.uv();
        c->payload_len = c->n * size_t(type_bits(type) / 8);
        c->payload = r.take(c->payload_len);
        break;
    case T_S8B: {
        c->val = type_ext(type, r.uv()); c->n = r.uv();
        c->payload_len = r.uv();
        c->payload = r.take(c->payload_len);
        break;
    }
    case T_DICT:
    case T_RUNEND: {
        long k = parse_container(type, r.p, r.left, c->child[0], err, depth + 1);
        if (k < 0) return -1;
        r.take(size_t(k));
        k = parse_container(c->ctype == T_DICT ? 7 /*uint16*/ : 6 /*uint32*/, r.p, r.left, c->child[1], err, depth + 1);
        if (k < 0) return -1;
        r.take(size_t(k));
        if (c->ctype == T_DICT) c->n = c->child[1]->n;
        else {
            // N = last(Ends) + 1 (int_runend.go:108-111)
            std::vector<uint64_t> ends;
            if (!decode_container(*c->child[1], ends, err)) return -1;
            c->n = ends.empty() ? 0 : size_t(ends.back()) + 1;
            if (c->child[0]->n != c->child[1]->n) { err = "runend: values/ends length mismatch"; return -1; }
        }
        break;
    }
    case T_FLOATALP: {   // float_alp.go:109-165: uv(Exponent) uv(Factor) u8 flags <Values int64> [<Patches T> <Positions uint32>]
        if (type != 9) { err = "ALP: only float64 blocks are supported"; return -1; }
        c->alp_e = int(r.uv()); c->alp_f = int(r.uv());
        const uint8_t* fl = r.take(1);
        if (!r.ok || !fl) { err = "truncated container"; return -1; }
        c->alp_flags = *fl;
        if (c->alp_e > 20 || c->alp_f > 23) { err = "ALP: exponent out of range"; return -1; }
        long k = parse_container(1 /*int64*/, r.p, r.left, c->child[0], err, depth + 1);
        if (k < 0) return -1;
        r.take(size_t(k));
        if (c->alp_flags & 1) {
            k = parse_container(type, r.p, r.left, c->child[1], err, depth + 1);
            if (k < 0) return -1;
            r.take(size_t(k));
            k = parse_container(6 /*uint32*/, r.p, r.left, c->child[2], err, depth + 1);
            if (k < 0) return -1;
            r.take(size_t(k));
            if (c->child[1]->n != c->child[2]->n) { err = "ALP: patches/positions length mismatch"; return -1; }
        }
        c->n = c->child[0]->n;
        break;
    }
    case T_FLOATALPRD: {   // float_alprd.go:78-107: <Left int container of uint16> <Right int container of uint64 / uint32> uv(Shift)
        if (!type_is_float(type)) { err = "ALP-RD container in a non-float block"; return -1; }
        long k = parse_container(7 /*uint16*/, r.p, r.left, c->child[0], err, depth + 1);
        if (k < 0) return -1;
        r.take(size_t(k));
        k = parse_container(type == 9 ? 5 /*uint64*/ : 6 /*uint32*/, r.p, r.left, c->child[1], err, depth + 1);
        if (k < 0) return -1;
        r.take(size_t(k));
        c->log2 = int(r.uv());   // the cut: right = low `Shift` bits, left = the (at most 16) bits above
        if (!r.ok) { err = "truncated container"; return -1; }
        if (c->log2 < 0 || c->log2 >= type_bits(type) || type_bits(type) - c->log2 > 16) { err = "ALP-RD: bad shift"; return -1; }
        if (c->child[0]->n != c->child[1]->n) { err = "ALP-RD: left/right length mismatch"; return -1; }
        c->n = c->child[0]->n;
        break;
    }
    default:
        err = "unsupported container type " + std::to_string(c->ctype);
        return -1;
    }
    if (!r.ok) { err = "truncated container"; return -1; }
    if (c->n > 0xffffffffull) { err = "block longer than 2^32 rows"; return -1; }
    out = std::move(c);
    return long(len - r.left);
}

bool decode_container(const Container& c, std::vector<uint64_t>& out, std::string& err) {
    // constant / affine containers encode any length in a few bytes: refuse lengths no pack can have before materialising
    if (c.n > MAX_MATERIALIZED_ROWS) { err = "container claims " + std::to_string(c.n) + " rows"; return false; }
    out.resize(c.n);
    switch (c.ctype) {
    case T_CONST:
        std::fill(out.begin(), out.end(), c.val);
        return true;
    case T_DELTA:
        for (size_t i = 0; i < c.n; i++) out[i] = type_ext(c.type, uint64_t(i) * c.delta + c.val);
        return true;
    case T_BITPACK:
        for (size_t i = 0; i < c.n; i++) out[i] = type_ext(c.type, bit_field(c.payload, c.payload_len, i, c.log2) + c.val);
        return true;
    case T_RAW:
    case T_FLOATRAW: {
        int nb = type_bits(c.type) / 8;
        for (size_t i = 0; i < c.n; i++) {
            uint64_t v = load_le(c.payload + i * size_t(nb), nb);
            out[i] = type_is_float(c.type) ? v : type_ext(c.type, v);
        }
        return true;
    }
    case T_S8B: {
        size_t j = 0;
        for (size_t wi = 0; wi + 8 <= c.payload_len && j < c.n; wi += 8) {
            uint64_t w = load_le(c.payload + wi, 8);
            int sel = int(w >> 60), cnt = S8_COUNT[sel], bits = S8_WIDTH[sel];
            for (int q = 0; q < cnt && j < c.n; q++) {
                uint64_t f = sel == 0 ? 0 : sel == 1 ? 1 : (w >> (q * bits)) & width_mask(bits);
                out[j++] = type_ext(c.type, f + c.val);
            }
        }
        if (j != c.n) { err = "simple8b: short stream"; return false; }
        return true;
    }
    case T_DICT: {
        std::vector<uint64_t> dict, codes;
        if (!decode_container(*c.child[0], dict, err) || !decode_container(*c.child[1], codes, err)) return false;
        for (size_t i = 0; i < c.n; i++) {
            if (codes[i] >= dict.size()) { err = "dict: code out of range"; return false; }
            out[i] = dict[codes[i]];
        }
        return true;
    }
    case T_FLOATALP: {
        if (!decode_container(*c.child[0], out, err)) return false;
        for (size_t i = 0; i < c.n; i++) { double d = alp_decode(int64_t(out[i]), c.alp_e, c.alp_f); std::memcpy(&out[i], &d, 8); }
        if (c.alp_flags & 1) {
            std::vector<uint64_t> pv, pp;
            if (!decode_container(*c.child[1], pv, err) || !decode_container(*c.child[2], pp, err)) return false;
            for (size_t k = 0; k < pp.size(); k++) { if (pp[k] >= c.n) { err = "ALP: patch position out of range"; return false; } out[pp[k]] = pv[k]; }
        }
        return true;
    }
    case T_FLOATALPRD: {
        std::vector<uint64_t> l, r;
        if (!decode_container(*c.child[0], l, err) || !decode_container(*c.child[1], r, err)) return false;
        for (size_t i = 0; i < c.n; i++) { uint64_t b = ((l[i] & 0xffff) << c.log2) | r[i]; out[i] = c.type == 10 ? uint64_t(uint32_t(b)) : b; }
        return true;
    }
    case T_RUNEND: {
        std::vector<uint64_t> vals, ends;
        if (!decode_container(*c.child[0], vals, err) || !decode_container(*c.child[1], ends, err)) return false;
        size_t i = 0;
        for (size_t r = 0; r < ends.size(); r++) {
            if (ends[r] >= c.n) { err = "runend: end out of range"; return false; }
            for (; i <= ends[r]; i++) out[i] = vals[r];
        }
        return true;
    }
    }
    err = "decode: unsupported container";
    return false;
}

// ---------------------------------------------------------------------------------- normalise
namespace {

// host bit-packer for blocks that are transcoded at registration (simple8b, exotic children)
void pack_stream(const std::vector<uint64_t>& fields, int w, std::vector<uint8_t>& out) {
    out.assign(bitpack_bytes(w, fields.size()), 0);
    uint64_t* words = reinterpret_cast<uint64_t*>(out.data());
    size_t bit = 0;
    for (uint64_t f : fields) {
        size_t wi = bit >> 6; int sh = int(bit & 63);
        words[wi] |= f << sh;
        if (sh + w > 64) words[wi + 1] |= f >> (64 - sh);
        bit += size_t(w);
    }
}

// fills view/stream for a leaf-level stream container (bitpack / raw); false if not stream-like
bool stream_view(const Container& c, BlockLayout& out) {
    ColView& v = out.view;
    if (c.ctype == T_BITPACK) {
        v.kind = c.log2 == 0 ? CK_CONST : CK_BITS;
        v.base = c.val; v.width = uint8_t(c.log2); v.is_raw = 0;
        out.stream = c.payload; out.stream_len = c.payload_len;
        return true;
    }
    if (c.ctype == T_RAW || c.ctype == T_FLOATRAW) {
        v.kind = CK_BITS; v.base = 0; v.width = uint8_t(type_bits(c.type)); v.is_raw = 1;
        out.stream = c.payload; out.stream_len = c.payload_len;
        return true;
    }
    return false;
}

int log2_range(uint64_t lo, uint64_t hi) { uint64_t d = hi - lo; return d ? 64 - __builtin_clzll(d) : 0; }

}  // namespace

int normalize_block(int type, const uint8_t* enc, size_t len, BlockLayout& out, std::string& err) {
    std::unique_ptr<Container> c;
    if (type_bits(type) == 0) { err = "unsupported block type"; return -6; }
    long used = parse_container(type, enc, len, c, err);
    if (used < 0) return -6;
    ColView& v = out.view;
    v = ColView{};
    v.type = uint8_t(type);
    v.n = uint32_t(c->n);
    switch (c->ctype) {
    case T_CONST:
        v.kind = CK_CONST; v.base = c->val;
        return 0;
    case T_DELTA:
        v.kind = CK_DELTA; v.base = c->val; v.delta = c->delta;
        return 0;
    case T_BITPACK:
    case T_RAW:
    case T_FLOATRAW:
        stream_view(*c, out);
        return 0;
    case T_FLOATALP: {
        // the encoded int64 values stay a min-FOR bit stream; patches (few) become a small resident blob
        const Container& vc = *c->child[0];
        v.kind = CK_ALP; v.delta = (uint64_t(c->alp_e) << 8) | uint64_t(c->alp_f); v.is_raw = 0;
        BlockLayout tmp;
        if (vc.ctype == T_BITPACK) {
            v.base = vc.val; v.width = uint8_t(vc.log2);
            out.stream = vc.payload; out.stream_len = vc.payload_len;
        } else if (vc.ctype == T_CONST) {
            v.base = vc.val; v.width = 0;
        } else {   // raw / delta / dict / run-end children: re-pack once at registration
            std::vector<uint64_t> vals;
            if (!decode_container(vc, vals, err)) return -6;
            int64_t mn = INT64_MAX, mx = INT64_MIN;
            for (uint64_t x : vals) { mn = std::min(mn, int64_t(x)); mx = std::max(mx, int64_t(x)); }
            if (vals.empty()) mn = mx = 0;
            int w = log2_range(uint64_t(mn), uint64_t(mx));
            for (auto& x : vals) x -= uint64_t(mn);
            v.base = uint64_t(mn); v.width = uint8_t(w);
            if (w) pack_stream(vals, w, out.owned);
        }
        if (c->alp_flags & 1) {
            std::vector<uint64_t> pv, pp;
            if (!decode_container(*c->child[1], pv, err) || !decode_container(*c->child[2], pp, err)) return -6;
            const uint32_t np = uint32_t(pp.size());
            v.naux = np;
            if (np) {
                size_t mask_words = (size_t(v.n) + 31) / 32;
                out.blob.assign(alp_mask_off(np) + mask_words * 4 + 64, 0);
                uint32_t* bp = reinterpret_cast<uint32_t*>(out.blob.data());
                uint64_t* bv = reinterpret_cast<uint64_t*>(out.blob.data() + alp_vals_off(np));
                uint32_t* bm = reinterpret_cast<uint32_t*>(out.blob.data() + alp_mask_off(np));
                for (uint32_t k = 0; k < np; k++) {
                    if (pp[k] >= v.n || (k && pp[k] <= pp[k - 1])) { err = "ALP: patch positions must ascend inside the block"; return -6; }
                    bp[k] = uint32_t(pp[k]); bv[k] = pv[k];
                    bm[pp[k] >> 5] |= 1u << (pp[k] & 31);
                }
                // the encoded value stored at patch positions (Encoder.Encode writes the vector minimum there)
                const uint8_t* st = out.owned.empty() ? out.stream : out.owned.data();
                size_t sl = out.owned.empty() ? out.stream_len : out.owned.size();
                v.extra = v.base + (v.width ? bit_field(st, sl, pp[0], v.width) : 0);
            }
        }
        return 0;
    }
    case T_FLOATALPRD: {
        // both halves stay bit streams on the device: the right one verbatim (always bit-packed by the encoder), the left one
        // verbatim too — bit-packed left values or the bit-packed codes of a dictionary of at most eight uint16 entries
        // (float_alprd.go:148-160).  Any other child shape (never written by the encoder) is re-packed once here.
        const Container &lc = *c->child[0], &rc = *c->child[1];
        v.kind = CK_ALPRD; v.is_raw = 0;
        uint32_t lw = 0, isdict = 0; uint64_t lfor = 0;
        auto repack = [&](const Container& x, std::vector<uint8_t>& dst, uint32_t& w, uint64_t& base) -> bool {
            std::vector<uint64_t> vals;
            if (!decode_container(x, vals, err)) return false;
            uint64_t mn = ~0ull, mx = 0;
            for (uint64_t q : vals) { mn = std::min(mn, q); mx = std::max(mx, q); }
            if (vals.empty()) mn = mx = 0;
            w = uint32_t(log2_range(mn, mx)); base = mn;
            for (auto& q : vals) q -= mn;
            if (w) pack_stream(vals, int(w), dst);
            return true;
        };
        if (rc.ctype == T_BITPACK) { v.width = uint8_t(rc.log2); v.base = rc.val; out.stream = rc.payload; out.stream_len = rc.payload_len; }
        else if (rc.ctype == T_CONST) { v.width = 0; v.base = rc.val; }
        else { uint32_t w = 0; uint64_t b = 0; if (!repack(rc, out.owned, w, b)) return -6; v.width = uint8_t(w); v.base = b; }
        bool done = false;
        if (lc.ctype == T_BITPACK) {
            lw = uint32_t(lc.log2); lfor = lc.val & 0xffff;
            out.blob.assign(lc.payload, lc.payload + lc.payload_len);
            done = true;
        } else if (lc.ctype == T_CONST) {
            lw = 0; lfor = lc.val & 0xffff; done = true;
        } else if (lc.ctype == T_DICT && lc.child[1]->ctype == T_BITPACK && lc.child[0]->n <= 8) {
            std::vector<uint64_t> dv;
            if (!decode_container(*lc.child[0], dv, err)) return -6;
            for (size_t k = 0; k < dv.size(); k++) (k < 4 ? v.delta : v.extra) |= (dv[k] & 0xffff) << (16 * (k & 3));
            const Container& codes = *lc.child[1];
            if (codes.val + ((codes.log2 >= 16) ? 0xffffull : ((1ull << codes.log2) - 1)) >= 8 && codes.log2) {
                // codes could exceed the eight slots only in a corrupt block: check them
                std::vector<uint64_t> cv;
                if (!decode_container(codes, cv, err)) return -6;
                for (uint64_t q : cv) if (q >= dv.size()) { err = "ALP-RD: left code outside its dictionary"; return -6; }
            } else if (codes.val >= dv.size() && codes.n) { err = "ALP-RD: left code outside its dictionary"; return -6; }
            lw = uint32_t(codes.log2); lfor = codes.val & 0xffff; isdict = 1;
            out.blob.assign(codes.payload, codes.payload + codes.payload_len);
            done = true;
        }
        if (!done) { uint32_t w = 0; uint64_t b = 0; if (!repack(lc, out.blob, w, b)) return -6; lw = w; lfor = b & 0xffff; }
        out.blob.resize(out.blob.size() + 64, 0);   // readable slack behind the left stream
        v.naux = lw | (isdict << 8) | (uint32_t(c->log2) << 16);
        v.pad = uint32_t(lfor);
        return 0;
    }
    case T_S8B: {
        // legacy scheme (never selected for new data, context.go:275-282): transcoded ON THE DEVICE to a min-FOR bit stream
        // once at registration (s8b_count_kernel / s8b_pack_kernel), so the scan kernels see one stream format
        if (c->payload_len % 8) { err = "simple8b: stream is not a whole number of words"; return -6; }
        v.kind = CK_BITS; v.base = c->val; v.width = 0; v.is_raw = 0;
        out.stream = c->payload; out.stream_len = c->payload_len;
        out.s8b = true;
        return 0;
    }
    case T_DICT: {
        if (!decode_container(*c->child[0], out.aux64, err)) return -6;
        v.kind = CK_DICT; v.naux = uint32_t(out.aux64.size());
        const Container& codes = *c->child[1];
        BlockLayout tmp;
        if (stream_view(codes, tmp) && tmp.view.kind == CK_BITS) {
            v.width = tmp.view.width; v.delta = tmp.view.base; v.is_raw = tmp.view.is_raw;
            out.stream = tmp.stream; out.stream_len = tmp.stream_len;
        } else {
            std::vector<uint64_t> cv;
            if (!decode_container(codes, cv, err)) return -6;
            v.width = 16; v.delta = 0; v.is_raw = 1;
            pack_stream(cv, 16, out.owned);
        }
        return 0;
    }
    case T_RUNEND: {
        std::vector<uint64_t> ends;
        if (!decode_container(*c->child[0], out.aux64, err) || !decode_container(*c->child[1], ends, err)) return -6;
        out.aux32.assign(ends.begin(), ends.end());
        v.kind = CK_RUNEND; v.naux = uint32_t(ends.size());
        // an affine Values child keeps its (For, Delta): the reference matches it with DeltaContainer's index arithmetic
        // (int_runend.go:224-283 → int_delta.go:149-449), rounding quirk included — compile_leaf does the same over the runs
        if (c->child[0]->ctype == T_DELTA && c->child[0]->n == ends.size()) { v.is_raw = 1; v.base = c->child[0]->val; v.delta = c->child[0]->delta; }
        return 0;
    }
    }
    err = "normalize: unsupported container";
    return -7;
}

// ---------------------------------------------------------------------------------- predicates
namespace {

inline bool t_lt(int t, uint64_t a, uint64_t b) { return type_is_signed(t) ? int64_t(a) < int64_t(b) : a < b; }
inline bool t_le(int t, uint64_t a, uint64_t b) { return !t_lt(t, b, a); }
inline uint64_t t_min(int t) { return type_is_signed(t) ? uint64_t(-(int64_t(1) << (type_bits(t) - 1))) : 0; }
inline uint64_t t_max(int t) {
    int w = type_bits(t);
    return type_is_signed(t) ? (uint64_t(1) << (w - 1)) - 1 : width_mask(w);
}

void set_const(PackLeaf& o, bool all) { o.mode = all ? LM_ALL : LM_NONE; o.data = nullptr; o.width = 0; o.neg = 0; }

// closed interval [lo, hi] of T values (lo <= hi in T order) on a raw stream of `w` bits
void raw_interval(PackLeaf& o, int w, uint64_t lo, uint64_t hi, bool neg) {
    uint64_t wm = width_mask(w);
    o.mode = w <= 32 ? LM_RANGE32 : LM_RANGE64;
    o.a = lo & wm; o.d = (hi - lo) & wm; o.wm = wm; o.neg = neg;
}

// ((field - a) <=u64 d) on a min-FOR field of width w, evaluated by the reference in 64-bit
// arithmetic (bitpack/cmp.go, cmp_bw.go).  Fields are < 2^w, so the match set is rewritten
// as a w-bit MODULAR range ((field - a') mod 2^w) <= d' with a', d' < 2^w — the one form the
// kernels evaluate (top-aligned 32-bit arithmetic for w <= 32):
//   * a + d does not wrap 2^64 (every sane query): matches [a, a+d] ∩ [0, 2^w)
//   * a + d wraps (inverted range a > b): matches [a, 2^w) ∪ [0, e], e = a + d - 2^64
void packed_range(PackLeaf& o, int w, uint64_t a, uint64_t d, bool neg) {
    const uint64_t mask = width_mask(w);
    uint64_t e = a + d;                       // mod 2^64
    bool wraps = e < a;
    if (!wraps) {
        if (a > mask) { set_const(o, neg); return; }
        d = std::min(d, mask - a);
        if (a == 0 && d == mask) { set_const(o, !neg); return; }
    } else {
        if (a > mask) { a = 0; d = std::min(e, mask); }                 // only the low part [0, e]
        else if (e >= mask || e + 1 >= a) { set_const(o, !neg); return; }  // the two parts cover the whole field
        else d = (e - a) & mask;                                        // wrap-around range in w-bit arithmetic
        if (a == 0 && d == mask) { set_const(o, !neg); return; }
    }
    o.mode = w <= 32 ? LM_RANGE32 : LM_RANGE64;
    o.a = a; o.d = d; o.wm = mask; o.neg = neg;
}

// leaf on a CK_BITS stream of integer type t
void compile_bits_int(PackLeaf& o, int t, int w, uint64_t base, bool is_raw, int mode, uint64_t a, uint64_t b) {
    a = type_ext(t, a); b = type_ext(t, b);
    if (is_raw) {
        // cmp.<T><Op> on the reinterpreted slice: internal/encode/int_raw.go:122-337
        uint64_t mn = t_min(t), mx = t_max(t);
        switch (mode) {
        case M_EQ: raw_interval(o, w, a, a, false); return;
        case M_NE: raw_interval(o, w, a, a, true); return;
        case M_LT: if (a == mn) set_const(o, false); else raw_interval(o, w, mn, a - 1, false); return;
        case M_LE: raw_interval(o, w, mn, a, false); return;
        case M_GT: if (a == mx) set_const(o, false); else raw_interval(o, w, a + 1, mx, false); return;
        case M_GE: raw_interval(o, w, a, mx, false); return;
        case M_RANGE: {  // U(v - a) <= U(b - a): internal/cmp/number.go:211-243
            uint64_t wm = width_mask(w);
            o.mode = w <= 32 ? LM_RANGE32 : LM_RANGE64;
            o.a = a & wm; o.d = (b - a) & wm; o.wm = wm; o.neg = 0;
            return;
        }
        }
        set_const(o, false);
        return;
    }
    // bit-packed: `val < For` pre-checks, then compare in the min-FOR domain (int_bitpack.go:163-247)
    if (mode == M_RANGE) {
        if (t_lt(t, b, base)) { set_const(o, false); return; }
        if (t_lt(t, a, base)) a = base;
        uint64_t fa = type_ext(t, a - base), fb = type_ext(t, b - base);
        packed_range(o, w, fa, fb - fa, false);
        return;
    }
    if (t_lt(t, a, base)) { set_const(o, mode == M_NE || mode == M_GT || mode == M_GE); return; }
    uint64_t fa = type_ext(t, a - base);
    switch (mode) {
    case M_EQ: packed_range(o, w, fa, 0, false); return;
    case M_NE: packed_range(o, w, fa, 0, true); return;
    case M_LT: if (fa == 0) set_const(o, false); else packed_range(o, w, 0, fa - 1, false); return;   // field < fa
    case M_LE: packed_range(o, w, 0, fa, false); return;                                                 // field <= fa
    case M_GT: packed_range(o, w, 0, fa, true); return;                                                  // ^(field <= fa)
    case M_GE: if (fa == 0) set_const(o, true); else packed_range(o, w, 0, fa - 1, true); return;       // ^(field < fa)
    }
    set_const(o, false);
}

int64_t go_div(int64_t a, int64_t b) { return b == -1 ? int64_t(0 - uint64_t(a)) : a / b; }
int64_t go_mod(int64_t a, int64_t b) { return b == -1 ? 0 : a % b; }

// bits.SetRange(start, end) clamped like internal/bitset/bitset.go:156-200
void row_range(PackLeaf& o, int64_t n, int64_t start, int64_t end, bool neg) {
    if (start > n) { set_const(o, neg); return; }
    start = std::max<int64_t>(0, start);
    end = std::min<int64_t>(n - 1, end);
    if (start > end) { set_const(o, neg); return; }
    if (start == 0 && end == n - 1) { set_const(o, !neg); return; }
    o.mode = LM_ROWRANGE; o.a = uint64_t(start); o.d = uint64_t(end - start); o.neg = neg; o.data = nullptr; o.width = 0;
}

// DeltaContainer.Match*: closed-form index arithmetic in int64 space, int_delta.go:149-449
void compile_delta(PackLeaf& o, const ColView& v, int mode, uint64_t a, uint64_t b) {
    int t = v.type;
    a = type_ext(t, a); b = type_ext(t, b);
    int64_t N = int64_t(v.n);
    if (N == 0) { set_const(o, false); return; }
    uint64_t For = v.base;
    bool dpos = type_is_signed(t) ? int64_t(v.delta) > 0 : v.delta > 0;
    int64_t d64 = int64_t(v.delta), v64 = int64_t(a) - int64_t(For);
    auto lt = [&](uint64_t x, uint64_t y) { return t_lt(t, x, y); };
    auto gt = [&](uint64_t x, uint64_t y) { return t_lt(t, y, x); };

    if (mode == M_EQ || mode == M_NE) {
        bool ne = mode == M_NE;
        if (dpos ? lt(a, For) : gt(a, For)) { set_const(o, ne); return; }
        uint64_t val = type_ext(t, a - For);
        bool divisible; int64_t q;
        if (type_is_signed(t)) { divisible = go_mod(int64_t(val), d64) == 0; q = go_div(int64_t(val), d64); }
        else { divisible = v.delta != 0 && val % v.delta == 0; q = v.delta ? int64_t(val / v.delta) : -1; }
        bool hit = (ne ? (v.delta == 1 || divisible) : divisible) && q >= 0 && q < N;
        if (!hit) { set_const(o, ne); return; }
        if (N == 1) { set_const(o, !ne); return; }
        o.mode = LM_ROWRANGE; o.a = uint64_t(q); o.d = 0; o.neg = ne; o.data = nullptr; o.width = 0;
        return;
    }
    int64_t last = d64 * (N - 1);
    switch (mode) {
    case M_LT:
        if (dpos) {
            if (lt(a, For)) { set_const(o, false); return; }
            if (last < v64) { set_const(o, true); return; }
            int64_t n = go_div(v64, d64); if (go_mod(v64, d64) == 0) n--;
            row_range(o, N, 0, n, false);
        } else {
            if (gt(a, For)) { set_const(o, true); return; }
            if (last >= v64) { set_const(o, false); return; }
            row_range(o, N, go_div(v64, d64) + 1, N - 1, false);
        }
        return;
    case M_LE:
        if (dpos) {
            if (lt(a, For)) { set_const(o, false); return; }
            if (last < v64) { set_const(o, true); return; }
            row_range(o, N, 0, go_div(v64, d64), false);
        } else {
            if (!lt(a, For)) { set_const(o, true); return; }
            if (last > v64) { set_const(o, false); return; }
            int64_t n = go_div(v64, d64); if (go_mod(v64, d64) != 0) n++;
            row_range(o, N, n, N - 1, false);
        }
        return;
    case M_GT:
        if (dpos) {
            if (lt(a, For)) { set_const(o, true); return; }
            if (last < v64) { set_const(o, false); return; }
            row_range(o, N, go_div(v64, d64) + 1, N - 1, false);
        } else {
            if (gt(a, For)) { set_const(o, false); return; }
            if (last > v64) { set_const(o, true); return; }
            int64_t n = go_div(v64, d64); if (go_mod(v64, d64) == 0) n--;
            row_range(o, N, 0, n, false);
        }
        return;
    case M_GE:
        if (dpos) {
            if (!gt(a, For)) { set_const(o, true); return; }
            if (last < v64) { set_const(o, false); return; }
            int64_t n = go_div(v64, d64); if (go_mod(v64, d64) > 0) n++;
            row_range(o, N, n, N - 1, false);
        } else {
            if (gt(a, For)) { set_const(o, false); return; }
            if (last > v64) { set_const(o, true); return; }
            row_range(o, N, 0, go_div(v64, d64), false);
        }
        return;
    case M_RANGE: {
        int64_t a64 = int64_t(a) - int64_t(For), b64 = int64_t(b) - int64_t(For);
        if (dpos) {
            if (lt(b, For) || a64 > last) { set_const(o, false); return; }
            int64_t na = go_div(a64, d64), nb = go_div(b64, d64);
            if (go_mod(a64, d64) != 0) na++;      // also for a below For: reference behaviour, see DESIGN.md
            nb = std::min(nb, N - 1);
            row_range(o, N, na, nb, false);
        } else {
            if (gt(a, For) || b64 < last) { set_const(o, false); return; }
            int64_t na = go_div(a64, d64), nb = go_div(b64, d64);
            if (go_mod(b64, d64) != 0) nb++;
            nb = std::min(nb, N - 1);
            row_range(o, N, nb, na, false);
        }
        return;
    }
    }
    set_const(o, false);
}

// first index with dict[i] >= val (strict: > val), dict sorted ascending in T order
size_t dict_lower(const uint64_t* d, size_t l, int t, uint64_t val, bool strict) {
    size_t lo = 0, hi = l;
    while (lo < hi) {
        size_t m = (lo + hi) / 2;
        bool ok = strict ? t_lt(t, val, d[m]) : !t_lt(t, d[m], val);
        if (ok) hi = m; else lo = m + 1;
    }
    return lo;
}

// DictionaryContainer.Match*: value predicate → code predicate, int_dict.go:181-359
void compile_dict(PackLeaf& o, const ColView& v, const uint64_t* d, int mode, uint64_t a, uint64_t b) {
    int t = v.type; size_t l = v.naux;
    a = type_ext(t, a); b = type_ext(t, b);
    if (l == 0 || d == nullptr) { set_const(o, false); return; }
    uint64_t first = d[0], last = d[l - 1];
    auto codes = [&](int cmode, uint64_t ca, uint64_t cb) {
        compile_bits_int(o, 7 /*uint16*/, v.width, v.delta, v.is_raw != 0, cmode, ca, cb);
    };
    size_t idx;
    switch (mode) {
    case M_EQ:
        if (t_lt(t, a, first) || t_lt(t, last, a)) { set_const(o, false); return; }
        idx = dict_lower(d, l, t, a, false);
        if (idx == l || d[idx] != a) { set_const(o, false); return; }
        codes(M_EQ, idx, 0); return;
    case M_NE:
        if (t_lt(t, a, first) || t_lt(t, last, a)) { set_const(o, true); return; }
        idx = dict_lower(d, l, t, a, false);
        if (idx == l || d[idx] != a) { set_const(o, true); return; }
        codes(M_NE, idx, 0); return;
    case M_LT:
        if (t_lt(t, a, first)) { set_const(o, false); return; }
        if (t_lt(t, last, a)) { set_const(o, true); return; }
        idx = dict_lower(d, l, t, a, false); if (idx == l) idx--;
        codes(M_LT, idx, 0); return;
    case M_LE:
        if (t_lt(t, a, first)) { set_const(o, false); return; }
        if (!t_lt(t, a, last)) { set_const(o, true); return; }
        idx = dict_lower(d, l, t, a, false);
        if (idx == l || t_lt(t, a, d[idx])) idx--;
        codes(M_LE, idx, 0); return;
    case M_GT:
        if (t_lt(t, a, first)) { set_const(o, true); return; }
        if (!t_lt(t, a, last)) { set_const(o, false); return; }
        idx = dict_lower(d, l, t, a, true);
        codes(M_GE, idx, 0); return;
    case M_GE:
        if (t_lt(t, a, first)) { set_const(o, true); return; }
        if (t_lt(t, last, a)) { set_const(o, false); return; }
        idx = dict_lower(d, l, t, a, false);
        codes(M_GE, idx, 0); return;
    case M_RANGE: {
        if (t_lt(t, b, first) || t_lt(t, last, a)) { set_const(o, false); return; }
        if (t_le(t, a, first) && t_le(t, last, b)) { set_const(o, true); return; }
        size_t ai = dict_lower(d, l, t, a, false), bi = dict_lower(d, l, t, b, false);
        uint64_t x = d[ai];
        if (ai == bi && x != a && x != b) { set_const(o, false); return; }   // range inside a dictionary gap
        if (bi == l || d[bi] != b) bi--;
        codes(M_RANGE, ai, bi); return;
    }
    }
    set_const(o, false);
}

// scalar predicate on a decoded value → ((v ^ flip) - A) <= D
void compile_valrange(PackLeaf& o, int t, int mode, uint64_t a, uint64_t b) {
    a = type_ext(t, a); b = type_ext(t, b);
    uint64_t flip = type_is_signed(t) ? 0x8000000000000000ull : 0;
    uint64_t mn = t_min(t), mx = t_max(t);
    auto iv = [&](uint64_t lo, uint64_t hi, bool neg) {
        o.mode = LM_VALRANGE; o.a = lo ^ flip; o.d = (hi ^ flip) - (lo ^ flip); o.wm = flip; o.neg = neg;
    };
    switch (mode) {
    case M_EQ: iv(a, a, false); return;
    case M_NE: iv(a, a, true); return;
    case M_LT: if (a == mn) set_const(o, false); else iv(mn, a - 1, false); return;
    case M_LE: iv(mn, a, false); return;
    case M_GT: if (a == mx) set_const(o, false); else iv(a + 1, mx, false); return;
    case M_GE: iv(a, mx, false); return;
    case M_RANGE: o.mode = LM_VALRANGE; o.a = a ^ flip; o.d = b - a; o.wm = flip; o.neg = 0; return;
    }
    set_const(o, false);
}

// does the translated integer predicate accept the encoded value x?  (device semantics of LM_RANGE*/ALL/NONE)
bool packleaf_accepts(const PackLeaf& o, const ColView& vv, uint64_t x) {
    if (o.mode == LM_ALL) return true;
    if (o.mode == LM_NONE) return false;
    uint64_t field = (x - vv.base) & width_mask(vv.width);
    bool r = ((field - o.a) & o.wm) <= o.d;
    return r != (o.neg != 0);
}

}  // namespace

// FloatAlpContainer.Match* (float_alp.go:238-495): translate the float predicate into the encoded integer domain
// (EncodeSingle / EncodeAbove / EncodeBelow), run the integer matcher of the Values child, then correct the rows
// that are patches.  All patch slots hold the same replacement value, so the correction is one of: OR in the
// patches that satisfy the float predicate (replacement did not match) or AND out those that do not (it did).
static void compile_alp(PackLeaf& o, const ColView& v, int mode, uint64_t ua, uint64_t ub, uint32_t view_index) {
    double a, b; std::memcpy(&a, &ua, 8); std::memcpy(&b, &ub, 8);
    const int e = int(v.delta >> 8), f = int(v.delta & 0xff);
    const bool patched = v.naux > 0;
    ColView vv = v;   // the encoded values as an int64 min-FOR column
    vv.kind = v.width ? CK_BITS : CK_CONST; vv.type = 1; vv.is_raw = 0; vv.aux = nullptr; vv.naux = 0;
    auto inner = [&](int imode, int64_t x, int64_t y) {
        LeafSpec ls; ls.type = 1; ls.mode = uint8_t(imode); ls.a = uint64_t(x); ls.b = uint64_t(y);
        compile_leaf(vv, nullptr, ls, view_index, o);
    };
    auto fix_by_pred = [&]() {
        if (patched) o.fixmode = packleaf_accepts(o, vv, v.extra) ? FIX_ANDNOT_NPRED : FIX_OR_PRED;
    };
    const bool nan_a = a != a, nan_b = b != b;
    bool ok = false; int64_t av, bv;
    switch (mode) {
    case M_EQ: case M_NE: {   // :238-295
        bool ran = false;
        if (!nan_a) { av = alp_encode_single(a, e, f, &ok); if (ok) { inner(M_EQ, av, 0); ran = true; } }
        if (!ran) { o = PackLeaf{}; o.view = view_index; o.mode = LM_NONE; }
        if (patched) o.fixmode = (ran && packleaf_accepts(o, vv, v.extra)) ? FIX_ANDNOT_ALL : FIX_OR_PRED;
        o.neg2 = mode == M_NE;
        return;
    }
    case M_LT:   // :297-330
        if (nan_a || (std::isinf(a) && a < 0)) { o.mode = LM_NONE; return; }
        av = alp_encode_single(a, e, f, &ok);
        if (ok) inner(M_LT, av, 0); else inner(M_LE, alp_encode_below(a, e, f), 0);
        fix_by_pred();
        return;
    case M_LE:   // :332-371
        if (nan_a) { o.mode = LM_NONE; return; }
        if (std::isinf(a) && a > 0) { o.mode = LM_ALL; return; }
        av = alp_encode_single(a, e, f, &ok);
        inner(M_LE, ok ? av : alp_encode_below(a, e, f), 0);
        fix_by_pred();
        return;
    case M_GT:   // :373-407
        if (nan_a || (std::isinf(a) && a > 0)) { o.mode = LM_NONE; return; }
        av = alp_encode_single(a, e, f, &ok);
        if (ok) inner(M_GT, av, 0); else inner(M_GE, alp_encode_above(a, e, f), 0);
        fix_by_pred();
        return;
    case M_GE:   // :409-448
        if (nan_a) { o.mode = LM_NONE; return; }
        if (std::isinf(a) && a < 0) { o.mode = LM_ALL; return; }
        av = alp_encode_single(a, e, f, &ok);
        inner(M_GE, ok ? av : alp_encode_above(a, e, f), 0);
        fix_by_pred();
        return;
    case M_RANGE:   // :450-490
        if (nan_a || nan_b) { o.mode = LM_NONE; return; }
        av = alp_encode_single(a, e, f, &ok); if (!ok) av = alp_encode_above(a, e, f);
        bv = alp_encode_single(b, e, f, &ok); if (!ok) bv = alp_encode_below(b, e, f);
        inner(M_RANGE, av, bv);
        fix_by_pred();
        return;
    }
    o.mode = LM_NONE;
}

bool set_contains(const std::vector<uint64_t>& s, uint64_t v) { return std::binary_search(s.begin(), s.end(), v); }

bool build_set_table(const std::vector<uint64_t>& set, std::vector<uint64_t>& slots, int& log2nb) {
    slots.clear(); log2nb = 0;
    if (set.empty()) return false;
    int lg = 1;
    while ((size_t(4) << lg) < set.size() * 2) ++lg;          // start at a load factor <= 0.5
    const int lg_max = std::min(24, lg + 6);
    std::vector<uint8_t> fill;
    for (; lg <= lg_max; ++lg) {
        const size_t nb = size_t(1) << lg;
        slots.assign(nb * 4, 0); fill.assign(nb, 0);
        bool ok = true;
        for (uint64_t v : set) {
            uint32_t b = set_table_bucket(v, lg);
            if (fill[b] == 4) { ok = false; break; }
            slots[size_t(b) * 4 + fill[b]++] = v;
        }
        if (!ok) continue;
        for (size_t b = 0; b < nb; ++b) {
            uint64_t pad;
            if (fill[b]) pad = slots[b * 4];
            else { pad = 0; while (set_table_bucket(pad, lg) == b) ++pad; }   // a key that can never probe bucket b
            for (int k = fill[b]; k < 4; ++k) slots[b * 4 + size_t(k)] = pad;
        }
        log2nb = lg;
        return true;
    }
    slots.clear();
    return false;
}

// ---- byte-string containers
static bool load_u32_child(const uint8_t*& p, const uint8_t* end, std::vector<uint32_t>& out, size_t at, std::string& err) {
    std::unique_ptr<Container> c;
    long used = parse_container(6 /* uint32 */, p, size_t(end - p), c, err);
    if (used <= 0) { if (err.empty()) err = "bad nested container of a string block"; return false; }
    // constant / affine children encode any length in a few bytes: refuse lengths no pack can have before materialising
    if (c->n > (size_t(1) << 26)) { err = "string block: nested container claims " + std::to_string(c->n) + " rows"; return false; }
    std::vector<uint64_t> v;
    if (!decode_container(*c, v, err)) return false;
    if (out.size() < at + v.size()) out.resize(at + v.size());
    for (size_t i = 0; i < v.size(); ++i) out[at + i] = uint32_t(v[i]);
    p += used;
    return true;
}

int normalize_string_block(const uint8_t* enc, size_t len, StrLayout& out, std::string& err) {
    if (len < 2) { err = "string block too short"; return -6; }
    const uint8_t *p = enc + 1, *end = enc + len;
    ColView& v = out.view;
    v = ColView{};
    v.kind = CK_STR; v.type = 12;
    uint64_t x = 0;
    auto uv = [&](uint64_t& dst) { int k = get_uvarint(p, size_t(end - p), &dst); if (!k) { err = "truncated string block"; return false; } p += k; return true; };
    switch (enc[0]) {
    case T_STRCONST:   // string_const.go:49-69: uv(N) uv(len) val
        if (!uv(x)) return -6;
        if (x > (uint64_t(1) << 26)) { err = "string block claims " + std::to_string(x) + " rows"; return -6; }
        v.n = uint32_t(x);
        if (!uv(x) || size_t(end - p) < x) { err = "truncated string block"; return -6; }
        v.is_raw = STR_CONST; v.delta = x; out.bytes = p; out.nbytes = size_t(x);
        break;
    case T_STRFIXED:   // string_fixed.go:59-80: uv(N) uv(sz) N*sz bytes
        if (!uv(x)) return -6;
        if (x > (uint64_t(1) << 26)) { err = "string block claims " + std::to_string(x) + " rows"; return -6; }
        v.n = uint32_t(x);
        if (!uv(x) || (v.n && x > size_t(end - p) / v.n)) { err = "truncated string block"; return -6; }
        v.is_raw = STR_FIXED; v.delta = x; out.bytes = p; out.nbytes = size_t(x) * v.n;
        break;
    case T_STRCOMPACT: {   // string_compact.go:64-96: <Ofs> <Len> uv(len(buf)) buf
        if (!load_u32_child(p, end, out.idx, 0, err)) return -6;
        const size_t n = out.idx.size();
        if (!load_u32_child(p, end, out.idx, n, err) || out.idx.size() != 2 * n) { if (err.empty()) err = "string block: offsets / lengths differ in size"; return -6; }
        if (!uv(x) || size_t(end - p) < x) { err = "truncated string block"; return -6; }
        v.n = uint32_t(n); v.is_raw = STR_COMPACT; out.bytes = p; out.nbytes = size_t(x);
        for (size_t i = 0; i < n; ++i)
            if (uint64_t(out.idx[i]) + out.idx[n + i] > x) { err = "string block: row outside the buffer"; return -6; }
        break;
    }
    case T_STRDICT: {   // string_dict.go:70-109: <Ofs> <Len> <Code> uv(len(dict)) dict
        std::vector<uint32_t> ofs, ln, code;
        if (!load_u32_child(p, end, ofs, 0, err) || !load_u32_child(p, end, ln, 0, err) || !load_u32_child(p, end, code, 0, err)) return -6;
        if (ofs.size() != ln.size()) { err = "string block: offsets / lengths differ in size"; return -6; }
        if (!uv(x) || size_t(end - p) < x) { err = "truncated string block"; return -6; }
        const size_t n = code.size(), m = ofs.size();
        for (size_t i = 0; i < n; ++i) if (code[i] >= m) { err = "string block: code outside the dictionary"; return -6; }
        for (size_t k = 0; k < m; ++k) if (uint64_t(ofs[k]) + ln[k] > x) { err = "string block: entry outside the dictionary"; return -6; }
        out.idx.reserve(n + 2 * m);
        out.idx = code; out.idx.insert(out.idx.end(), ofs.begin(), ofs.end()); out.idx.insert(out.idx.end(), ln.begin(), ln.end());
        v.n = uint32_t(n); v.naux = uint32_t(m); v.is_raw = STR_DICT; out.bytes = p; out.nbytes = size_t(x);
        break;
    }
    default:
        err = "unknown string container id " + std::to_string(int(enc[0]));
        return -6;
    }
    v.base = out.nbytes;
    return 0;
}

static int bytes_compare(const uint8_t* a, size_t al, const uint8_t* b, size_t bl) {
    const size_t m = std::min(al, bl);
    const int c = m ? std::memcmp(a, b, m) : 0;
    if (c) return c < 0 ? -1 : 1;
    return al < bl ? -1 : (al > bl ? 1 : 0);
}
bool string_pred(int mode, const uint8_t* v, size_t vl, const uint8_t* a, size_t al, const uint8_t* b, size_t bl) {
    switch (mode) {
    case M_EQ: return vl == al && (al == 0 || !std::memcmp(v, a, al));
    case M_NE: return !(vl == al && (al == 0 || !std::memcmp(v, a, al)));
    case M_LT: return bytes_compare(v, vl, a, al) < 0;
    case M_LE: return bytes_compare(v, vl, a, al) <= 0;
    case M_GT: return bytes_compare(v, vl, a, al) > 0;
    case M_GE: return bytes_compare(v, vl, a, al) >= 0;
    case M_RANGE: return bytes_compare(v, vl, a, al) >= 0 && bytes_compare(v, vl, b, bl) <= 0;
    }
    return false;
}

// membership in a sorted, de-duplicated set laid out as [count x (offset u32, length u32)] + bytes (offsets from the table start)
bool string_set_pred(int mode, const uint8_t* v, size_t vl, const uint8_t* table, uint32_t count) {
    uint32_t lo = 0, hi = count;
    bool found = false;
    while (lo < hi) {
        const uint32_t m = (lo + hi) / 2;
        uint32_t off, len;
        std::memcpy(&off, table + size_t(m) * 8, 4); std::memcpy(&len, table + size_t(m) * 8 + 4, 4);
        const int c = bytes_compare(v, vl, table + off, len);
        if (c == 0) { found = true; break; }
        if (c < 0) hi = m; else lo = m + 1;
    }
    return mode == M_IN ? found : !found;
}

void build_set_prefilter(const std::vector<uint64_t>& set, std::vector<uint32_t>& words, int& log2bits) {
    int lg = 10;
    while (lg < 17 && (size_t(1) << lg) < set.size() * 256) ++lg;
    log2bits = lg;
    words.assign((size_t(1) << lg) / 32, 0u);
    for (uint64_t v : set) {
        uint32_t idx = set_hash32(v) >> (32 - lg);
        words[idx >> 5] |= 1u << (idx & 31u);
    }
}

bool scalar_match(int t, int mode, uint64_t v, uint64_t a, uint64_t b) {
    if (type_is_float(t)) {
        double x, y, z;
        if (t == 10) {
            float fx, fy, fz; uint32_t u;
            u = uint32_t(v); std::memcpy(&fx, &u, 4); u = uint32_t(a); std::memcpy(&fy, &u, 4); u = uint32_t(b); std::memcpy(&fz, &u, 4);
            x = fx; y = fy; z = fz;
        } else { std::memcpy(&x, &v, 8); std::memcpy(&y, &a, 8); std::memcpy(&z, &b, 8); }
        switch (mode) {
        case M_EQ: return x == y; case M_NE: return x != y; case M_LT: return x < y; case M_LE: return x <= y;
        case M_GT: return x > y; case M_GE: return x >= y; case M_RANGE: return y <= x && x <= z;
        }
        return false;
    }
    a = type_ext(t, a); b = type_ext(t, b);
    switch (mode) {
    case M_EQ: return v == a;
    case M_NE: return v != a;
    case M_LT: return t_lt(t, v, a);
    case M_LE: return t_le(t, v, a);
    case M_GT: return t_lt(t, a, v);
    case M_GE: return t_le(t, a, v);
    case M_RANGE: return t_le(t, a, v) && t_le(t, v, b);
    }
    return false;
}

void compile_leaf(const ColView& v, const uint64_t* dict_host, const LeafSpec& leaf, uint32_t view_index, PackLeaf& o) {
    o = PackLeaf{};
    o.view = view_index;
    int mode = leaf.mode, t = v.type;
    bool is_set = mode == M_IN || mode == M_NIN;
    if (v.n == 0) { set_const(o, false); return; }

    if (is_set) {
        bool neg = mode == M_NIN;
        if (leaf.set.empty()) { set_const(o, neg); return; }
        if (v.kind == CK_CONST) { set_const(o, set_contains(leaf.set, v.base) != neg); return; }
        o.neg = neg;
        if (v.kind == CK_DICT) {
            // DictionaryContainer.MatchInSet (int_dict.go:361-398): the set is translated into a bitmap
            // over this pack's codes on the device (codeset_kernel); the launch assigns o.a
            o.mode = LM_CODESET; o.wm = v.delta; o.d = v.naux; o.a = 0;
            o.data = v.data; o.width = v.width;
            return;
        }
        if (v.kind == CK_BITS && leaf.has_table && !type_is_float(t)) {
            // int_bitpack.go:249-291 / int_raw.go:339-380: decoded value looked up in the leaf's hash table
            o.mode = LM_HASHSET; o.data = v.data; o.width = v.width; o.a = v.base; o.fop = v.type;   // For and element type travel with the leaf
            return;
        }
        // decoded value ∈ sorted set, binary search per row (affine blocks) / per run (run-end blocks)
        o.mode = LM_SET;
        o.a = leaf.set_off; o.d = leaf.set.size();
        if (v.kind == CK_BITS) { o.data = v.data; o.width = v.width; }
        return;
    }

    switch (v.kind) {
    case CK_ALP:
        compile_alp(o, v, mode, leaf.a, leaf.b, view_index);
        return;
    case CK_CONST:   // ConstContainer.Match*: all-or-nothing, int_const.go:133-173
        set_const(o, scalar_match(t, mode, v.base, leaf.a, leaf.b));
        return;
    case CK_DELTA:
        compile_delta(o, v, mode, leaf.a, leaf.b);
        return;
    case CK_BITS:
        if (type_is_float(t)) {   // FloatRawContainer.Match* → cmp.Float64<Op>, float_raw.go:116-207
            o.mode = LM_FLOAT; o.fop = uint8_t(mode); o.a = leaf.a; o.d = leaf.b;
        } else {
            compile_bits_int(o, t, v.width, v.base, v.is_raw != 0, mode, leaf.a, leaf.b);
        }
        if (o.mode != LM_NONE && o.mode != LM_ALL) { o.data = v.data; o.width = v.width; }
        return;
    case CK_DICT:
        compile_dict(o, v, dict_host, mode, leaf.a, leaf.b);
        if (o.mode != LM_NONE && o.mode != LM_ALL) { o.data = v.data; o.width = v.width; }
        return;
    case CK_RUNEND:  // RunEndContainer.Match*: predicate on the run values, int_runend.go:224-318
        if (v.is_raw) {
            // Values is a DeltaContainer: its closed-form matchers pick a range of RUN indexes (LM_ROWRANGE over naux runs);
            // the run pre-pass turns the runs of that range into rows (applyMatch, int_runend.go:296-318)
            ColView dv{};
            dv.kind = CK_DELTA; dv.type = v.type; dv.base = v.base; dv.delta = v.delta; dv.n = v.naux;
            compile_delta(o, dv, mode, leaf.a, leaf.b);
            if (o.mode == LM_ROWRANGE) o.mode = LM_RUNRANGE;
            o.view = view_index;
            return;
        }
        compile_valrange(o, t, mode, leaf.a, leaf.b);
        return;
    }
    set_const(o, false);
}

// ---------------------------------------------------------------------------------- XXH3-64
// one implementation for host and device: kx_xxh3.h
uint64_t xxh3_bytes(const uint8_t* in, size_t len) { return xxh3::bytes(in, len); }
uint64_t xxh3_u64(uint64_t v) { return xxh3::u64(v); }
uint64_t xxh3_u32(uint32_t v) { return xxh3::u32(v); }
uint64_t xxh3_u16(uint16_t v) { return xxh3::u16(v); }
uint64_t xxh3_u8(uint8_t v) { return xxh3::u8(v); }

}  // namespace kx
