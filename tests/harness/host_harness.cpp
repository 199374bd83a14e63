// host_harness.cpp — TEST INFRASTRUCTURE ONLY (never part of libknoxgpu.so).
//
// Links the product's host-side translation code (knoxdb_b200/csrc/kx_host.cpp: container
// parsing, block normalisation, leaf → PackLeaf translation) into a CPU-only shared library
// and interprets the resulting PackLeaf records with a scalar emulator of the device
// semantics documented in kx_types.h.  This lets the CPU test tier check the translation
// logic (min-FOR pre-checks, dictionary code ranges, affine row ranges …) against the oracle
// without a GPU.  The CUDA kernels themselves are only exercised by the `-m gpu` tests.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../knoxdb_b200/csrc/kx_host.h"

using namespace kx;

namespace {

uint64_t field_at(const uint8_t* s, size_t len, uint64_t row, int w) {
    if (w == 0) return 0;
    uint64_t bit = row * uint64_t(w); size_t byte = bit >> 3; int sh = int(bit & 7);
    unsigned __int128 acc = 0;
    for (int i = 0; i < 9 && byte + i < len; i++) acc |= (unsigned __int128)s[byte + i] << (8 * i);
    return uint64_t(acc >> sh) & width_mask(w);
}

// scalar emulation of the device-side Simple8b transcode (s8b_count_kernel / s8b_pack_kernel in kx_scan.cu): codewords →
// fixed-width LSB-first stream of (value - For), width = widest value
bool s8b_transcode(BlockLayout& lay) {
    static const int CNT[16] = {128, 128, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
    static const int BITS[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};
    std::vector<uint64_t> f;
    for (size_t wi = 0; wi + 8 <= lay.stream_len; wi += 8) {
        uint64_t w; std::memcpy(&w, lay.stream + wi, 8);
        const int sel = int(w >> 60);
        for (int q = 0; q < CNT[sel]; q++) f.push_back(sel == 0 ? 0 : sel == 1 ? 1 : (w >> (q * BITS[sel])) & width_mask(BITS[sel]));
    }
    if (f.size() < lay.view.n) return false;
    uint64_t any = 0;
    for (uint64_t x : f) any |= x;
    const int w = any ? 64 - __builtin_clzll(any) : 0;
    lay.view.width = uint8_t(w);
    lay.view.kind = w ? CK_BITS : CK_CONST;
    lay.owned.assign((size_t(lay.view.n) * size_t(w) + 7) / 8 + 16, 0);
    for (uint32_t r = 0; r < lay.view.n && w; r++) {
        const uint64_t bit = uint64_t(r) * uint64_t(w);
        for (int k = 0; k < w; k++)
            if ((f[r] >> k) & 1) lay.owned[(bit + k) >> 3] |= uint8_t(1u << ((bit + k) & 7));
    }
    lay.s8b = false;
    return true;
}

int normalize(int block_type, const uint8_t* enc, size_t len, BlockLayout& lay, std::string& err) {
    int rc = normalize_block(block_type, enc, len, lay, err);
    if (!rc && lay.s8b && !s8b_transcode(lay)) { err = "simple8b: short stream"; return -6; }
    return rc;
}

struct HostBlock {
    BlockLayout lay;
    const uint8_t* stream() const { return lay.owned.empty() ? lay.stream : lay.owned.data(); }
    size_t stream_len() const { return lay.owned.empty() ? lay.stream_len : lay.owned.size(); }
    uint64_t value(uint32_t row) const {
        const ColView& v = lay.view;
        switch (v.kind) {
        case CK_CONST: return v.base;
        case CK_DELTA: return type_ext(v.type, uint64_t(row) * v.delta + v.base);
        case CK_BITS: {
            uint64_t f = field_at(stream(), stream_len(), row, v.width);
            return type_is_float(v.type) ? f : type_ext(v.type, f + v.base);
        }
        case CK_DICT: return lay.aux64[field_at(stream(), stream_len(), row, v.width) + v.delta];
        case CK_ALP: {
            const uint32_t np = v.naux;
            if (np) {
                const uint32_t* pos = reinterpret_cast<const uint32_t*>(lay.blob.data());
                const uint64_t* pv = reinterpret_cast<const uint64_t*>(lay.blob.data() + alp_vals_off(np));
                const uint32_t* it = std::lower_bound(pos, pos + np, row);
                if (it != pos + np && *it == row) return pv[it - pos];
            }
            uint64_t f = field_at(stream(), stream_len(), row, v.width);
            double d = alp_decode(int64_t(f + v.base), int(v.delta >> 8), int(v.delta & 0xff));
            uint64_t u; std::memcpy(&u, &d, 8); return u;
        }
        case CK_RUNEND: {
            size_t lo = 0, hi = lay.aux32.size();
            while (lo < hi) { size_t m = (lo + hi) / 2; if (lay.aux32[m] >= row) hi = m; else lo = m + 1; }
            return lay.aux64[lo];
        }
        }
        return 0;
    }
};

bool fpred(int op, double x, double a, double b) {
    switch (op) {
    case 1: return x == a; case 2: return x != a; case 3: return x > a; case 4: return x >= a;
    case 5: return x < a; case 6: return x <= a; case 9: return a <= x && x <= b;
    }
    return false;
}

}  // namespace

extern "C" {

// returns rows, or <0 on error; bits: ceil(n/8) bytes pre-zeroed; mode_out: PackLeaf.mode chosen
long kxh_match(int block_type, const uint8_t* enc, size_t len, int mode, uint64_t a, uint64_t b,
               const uint64_t* set, uint32_t nset, uint8_t* bits, int* mode_out) {
    HostBlock hb; std::string err;
    if (normalize(block_type, enc, len, hb.lay, err)) return -1;
    const ColView& v = hb.lay.view;
    LeafSpec leaf; leaf.type = uint8_t(block_type); leaf.mode = uint8_t(mode); leaf.a = a; leaf.b = b;
    std::vector<uint64_t> tab; int tab_log2 = 0;
    std::vector<uint32_t> pre; int pre_log2 = 0;   // one-hash prefilter bitmap, tested before the table like leaf_hashset does
    if (set) {
        leaf.set.assign(set, set + nset); std::sort(leaf.set.begin(), leaf.set.end()); leaf.set.erase(std::unique(leaf.set.begin(), leaf.set.end()), leaf.set.end());
        leaf.has_table = build_set_table(leaf.set, tab, tab_log2);
        if (leaf.has_table) build_set_prefilter(leaf.set, pre, pre_log2);
    }
    ColView dv = v;
    dv.data = hb.stream();   // non-null marks "has a stream" for compile_leaf
    PackLeaf L;
    compile_leaf(dv, hb.lay.aux64.empty() ? nullptr : hb.lay.aux64.data(), leaf, 0, L);
    if (mode_out) *mode_out = L.mode;
    for (uint32_t row = 0; row < v.n; row++) {
        bool p = false;
        switch (L.mode) {
        case LM_NONE: p = false; break;
        case LM_ALL: p = true; break;
        case LM_RANGE32: {
            uint32_t f = uint32_t(field_at(hb.stream(), hb.stream_len(), row, L.width));
            p = ((f - uint32_t(L.a)) & uint32_t(L.wm)) <= uint32_t(L.d);
            break;
        }
        case LM_RANGE64: {
            uint64_t f = field_at(hb.stream(), hb.stream_len(), row, L.width);
            p = ((f - L.a) & L.wm) <= L.d;
            break;
        }
        case LM_FLOAT: {
            uint64_t f = field_at(hb.stream(), hb.stream_len(), row, L.width);
            double x, da, db;
            if (L.width == 32) {
                float fx, fa, fb; uint32_t u = uint32_t(f); std::memcpy(&fx, &u, 4);
                u = uint32_t(L.a); std::memcpy(&fa, &u, 4); u = uint32_t(L.d); std::memcpy(&fb, &u, 4);
                x = fx; da = fa; db = fb;
            } else { std::memcpy(&x, &f, 8); std::memcpy(&da, &L.a, 8); std::memcpy(&db, &L.d, 8); }
            p = fpred(L.fop, x, da, db);
            break;
        }
        case LM_ROWRANGE: p = (uint64_t(row) - L.a) <= L.d; break;
        case LM_SET: p = set_contains(leaf.set, hb.value(row)); break;
        case LM_CODESET: {   // bit (field + code base) of the pack's code bitmap = dict[code] ∈ set (codeset_kernel)
            uint64_t code = field_at(hb.stream(), hb.stream_len(), row, L.width) + L.wm;
            p = code < L.d && set_contains(leaf.set, hb.lay.aux64[code]);
            break;
        }
        case LM_HASHSET: {   // leaf_hashset: compare with the four slots of the home bucket
            uint64_t val = type_ext(L.fop, field_at(hb.stream(), hb.stream_len(), row, L.width) + L.a);   // For / element type travel with the leaf
            const uint32_t idx = set_hash32(val) >> (32 - pre_log2);
            const bool cand = (pre[idx >> 5] >> (idx & 31u)) & 1u;   // phase 1: prefilter (no false negatives allowed)
            const uint64_t* b = tab.data() + size_t(set_table_bucket(val, tab_log2)) * 4;
            p = cand && (b[0] == val || b[1] == val || b[2] == val || b[3] == val);
            break;
        }
        case LM_VALRANGE: p = ((hb.value(row) ^ L.wm) - L.a) <= L.d; break;
        case LM_RUNRANGE: {   // runfill_kernel: the run that holds the row lies in the range of runs the closed form chose
            size_t lo = 0, hi = hb.lay.aux32.size();
            while (lo < hi) { size_t m = (lo + hi) / 2; if (hb.lay.aux32[m] >= row) hi = m; else lo = m + 1; }
            p = (uint64_t(lo) - L.a) <= L.d;
            break;
        }
        }
        if (L.neg && L.mode != LM_NONE && L.mode != LM_ALL) p = !p;
        if (L.fixmode) {   // ALP patch correction (alpfix_kernel + the fix stage of scan_kernel)
            const uint32_t np = v.naux;
            const uint32_t* pos = reinterpret_cast<const uint32_t*>(hb.lay.blob.data());
            const uint64_t* pvs = reinterpret_cast<const uint64_t*>(hb.lay.blob.data() + alp_vals_off(np));
            const uint32_t* it = std::lower_bound(pos, pos + np, row);
            if (it != pos + np && *it == row) {
                double x, fa, fb; std::memcpy(&x, &pvs[it - pos], 8); std::memcpy(&fa, &a, 8); std::memcpy(&fb, &b, 8);
                int m = mode == 2 ? 1 : mode;
                bool pred = m == 1 ? ((fa != fa) ? (x != x) : (x == fa)) : fpred(m, x, fa, fb);
                if (L.fixmode == FIX_OR_PRED) p = p || pred;
                else if (L.fixmode == FIX_ANDNOT_NPRED) p = p && pred;
                else p = false;
            }
        }
        if (L.neg2) p = !p;
        if (p) bits[row >> 3] |= uint8_t(1u << (row & 7));
    }
    return long(v.n);
}

long kxh_decode(int block_type, const uint8_t* enc, size_t len, uint64_t* dst, size_t cap) {
    HostBlock hb; std::string err;
    if (normalize(block_type, enc, len, hb.lay, err)) return -1;
    if (hb.lay.view.n > cap) return -2;
    for (uint32_t r = 0; r < hb.lay.view.n; r++) dst[r] = hb.value(r);
    return long(hb.lay.view.n);
}

int kxh_view_kind(int block_type, const uint8_t* enc, size_t len) {
    BlockLayout lay; std::string err;
    if (normalize(block_type, enc, len, lay, err)) return -1;
    return lay.view.kind;
}

uint64_t kxh_xxh3_bytes(const uint8_t* p, size_t len) { return xxh3_bytes(p, len); }
uint64_t kxh_xxh3_fixed(int nbytes, uint64_t v) {
    switch (nbytes) { case 8: return xxh3_u64(v); case 4: return xxh3_u32(uint32_t(v)); case 2: return xxh3_u16(uint16_t(v)); case 1: return xxh3_u8(uint8_t(v)); }
    return 0;
}

// Byte-string blocks: the product's normalisation (normalize_string_block: container ids 16..19 → byte buffer + flat u32
// index array) interpreted row by row with the layout rules of kx_types.h (STR_*) and the product's string_pred — the
// scalar twin of strmatch_kernel.  Returns rows, or < 0 on a parse error.
long kxh_str_match(const uint8_t* enc, size_t len, int mode, const uint8_t* a, size_t al, const uint8_t* b, size_t bl, uint8_t* bits) {
    StrLayout lay; std::string err;
    if (normalize_string_block(enc, len, lay, err)) return -1;
    const ColView& v = lay.view;
    const uint32_t n = v.n, m = v.naux;
    for (uint32_t i = 0; i < n; ++i) {
        size_t ofs = 0, ln = 0;
        switch (v.is_raw) {
        case STR_CONST: ofs = 0; ln = size_t(v.delta); break;
        case STR_FIXED: ln = size_t(v.delta); ofs = size_t(i) * ln; break;
        case STR_COMPACT: ofs = lay.idx[i]; ln = lay.idx[n + i]; break;
        default: { uint32_t c = lay.idx[i]; ofs = lay.idx[n + c]; ln = lay.idx[n + m + c]; break; }
        }
        if (ofs + ln > lay.nbytes) return -2;
        if (string_pred(mode, lay.bytes + ofs, ln, a, al, b, bl)) bits[i >> 3] |= uint8_t(1u << (i & 7));
    }
    return long(n);
}

}  // extern "C"
