// kx_types.h — structures shared by the host runtime and the CUDA kernels of libknoxgpu.
//
// The device never sees KnoxDB's Go object graph.  It sees, per (pack, field), a ColView:
// a flat description of one encoded column block resident in HBM, and per (pack, leaf) a
// PackLeaf: a filter leaf already translated by the host into the block's own domain
// (min-FOR domain for bit-packed blocks, code domain for dictionaries, row-index ranges for
// affine "delta" blocks).
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define KX_HD
#else
#define KX_HD __host__ __device__
#endif

namespace kx {

// how a column block is laid out on the device
enum ColKind : uint8_t {
    CK_NONE = 0,
    CK_CONST = 1,   // every row == base                         (IntConstant, bitpack w = 0)
    CK_DELTA = 2,   // row i == base + i * delta (in T)          (IntDelta: affine sequence)
    CK_BITS = 3,    // LSB-first bit stream, `width` bits per row, value = field + base
                    // (IntBitpacked verbatim; IntRaw / FloatRaw are the width = 8*sizeof(T) case)
    CK_DICT = 4,    // CK_BITS stream of codes (+ delta as code base); aux = dict values (u64)
    CK_RUNEND = 5,  // data = run values (u64), aux = inclusive run ends (u32)
    CK_ALP = 6,     // FloatAlp (float64): data = min-FOR bit stream of the encoded int64 values (width, base),
                    // value = double(field + base) * F10[factor] * IF10[exponent]; delta = exponent << 8 | factor;
                    // aux = patch blob [positions u32 x naux | values u64 x naux | patch bitmap u32 x ceil(n/32)],
                    // extra = the encoded replacement value stored at patch positions
    CK_ALPRD = 8,   // FloatAlpRd (float64 / float32): IEEE bits = (left << shift) | right.  data = RIGHT bit stream (`width`
                    // bits per row, value = field + base); aux = LEFT bit stream (lw bits per row, value = field + pad);
                    // naux = lw | is_dict << 8 | shift << 16; left values that are dictionary codes index the <= 8
                    // uint16 entries packed into delta (entries 0..3) and extra (4..7)
    CK_STR = 7,     // byte strings (BlockBytes): data = byte buffer, aux = u32 index array, is_raw = STR_* layout,
                    // delta = row size (STR_FIXED / STR_CONST), naux = dictionary entries (STR_DICT), base = bytes in data
};
// layouts of a CK_STR block: row i = data[ofs .. ofs + len)
enum : uint8_t {
    STR_CONST = 0,     // every row = data[0 .. delta)                                   (string_const.go)
    STR_FIXED = 1,     // ofs = i * delta, len = delta                                    (string_fixed.go)
    STR_COMPACT = 2,   // ofs = aux[i], len = aux[n + i]                                  (string_compact.go)
    STR_DICT = 3,      // c = aux[i]; ofs = aux[n + c], len = aux[n + naux + c]           (string_dict.go)
};

struct ColView {
    const uint8_t* data;
    const uint8_t* aux;
    uint64_t base;
    uint64_t delta;
    uint32_t n;
    uint32_t naux;
    uint8_t kind;
    uint8_t width;
    uint8_t type;    // types.BlockType
    uint8_t is_raw;  // CK_BITS: raw block (compare in T, width-bit modular arithmetic)
    uint32_t pad;
    uint64_t extra;  // CK_ALP: replacement value of the patch slots
};
// CK_ALP patch blob offsets (bytes) for naux patches
KX_HD inline size_t alp_vals_off(uint32_t np) { return ((size_t)np * 4 + 15) & ~(size_t)15; }
KX_HD inline size_t alp_mask_off(uint32_t np) { return (alp_vals_off(np) + (size_t)np * 8 + 15) & ~(size_t)15; }   // 16 B aligned: TMA source

// how a leaf predicate is evaluated for one pack
enum LeafMode : uint8_t {
    LM_NONE = 0,      // no row matches
    LM_ALL = 1,       // every row matches
    LM_RANGE32 = 2,   // staged stream, width <= 32: ((field - a) & wm) <= d   (32-bit)
    LM_RANGE64 = 3,   // staged stream, any width:   ((field - a) & wm) <= d   (64-bit)
    LM_FLOAT = 4,     // staged raw float stream: IEEE compare fop(a, b)
    LM_ROWRANGE = 5,  // row index in [a, d] (affine blocks: closed-form index arithmetic)
    LM_SET = 6,       // decoded value ∈ sorted set (binary search)
    LM_VALRANGE = 7,  // decoded value v: ((v ^ flip) - a) <= d  (run-end / generic fallback)
    LM_CODESET = 8,   // staged dictionary codes (code = field + wm): bit `field` of the pack's code bitmap at code_bits + a
                      // (d = number of codes); the bitmap is built on the device per (pack, leaf) and query
    LM_HASHSET = 9,   // staged integer stream: T(field + base) looked up in the leaf's bucketised hash table
    LM_BITS = 10,     // staged stream IS the leaf's bitset, 1 bit per row (run-end blocks: filled per run by
                      // runfill_kernel right before the scan)
    LM_RUNRANGE = 11, // run-end block whose run values are affine: RUN index in [a, a + d] (host-side only: becomes a
                      // runfill_kernel job and then LM_BITS)
};

struct PackLeaf {
    const uint8_t* data;   // staged stream (LM_RANGE*, LM_FLOAT, LM_SET on CK_BITS/CK_DICT)
    uint64_t a, d;         // range / operands (float: IEEE bits of a, b)
    uint64_t wm;           // modular mask (LM_RANGE*) or sign flip (LM_VALRANGE)
    uint8_t mode;
    uint8_t width;         // bits per row of the staged stream (0: nothing staged)
    uint8_t neg;           // invert the result (NE, GT, GE, NIN)
    uint8_t fop;           // LM_FLOAT: FilterMode ; float width via `width`
    uint32_t view;         // LM_SET / LM_VALRANGE: index into the ColView table of this leaf's block
    // ALP blocks: the integer predicate above runs on the encoded values; rows that are PATCHES (values the
    // encoding could not represent) are then corrected with a per-query 1-bit stream built by alpfix_kernel
    // (or the block's resident patch bitmap), and the float-level NOT (MatchNotEqual) is applied last.
    const uint8_t* fix;    // staged after the leaf's own stream when fixmode != 0
    uint8_t fixmode;       // FIX_*
    uint8_t neg2;          // invert after the fix
    uint8_t pad[6];
};
enum : uint8_t {
    FIX_NONE = 0,
    FIX_OR_PRED = 1,       // word |= {patch rows whose true value satisfies the float predicate}
    FIX_ANDNOT_NPRED = 2,  // word &= ~{patch rows whose true value does NOT satisfy it}
    FIX_ANDNOT_ALL = 3,    // word &= ~{all patch rows}   (fix = resident patch bitmap)
};

struct PackInfo {
    uint32_t n;           // rows
    uint32_t tile0;       // first tile index of this pack
    uint64_t bitset_off;  // byte offset of the pack's bitset in the output buffer
};

// per-CTA partial aggregate (combined in fixed order by finalize kernel)
struct AggPartial {
    uint64_t count;
    uint64_t sum;      // integer sum mod 2^64, or IEEE bits of compensated sum (hi)
    double   err;      // float: compensation (lo)
    uint64_t mn, mx;   // order-preserving unsigned domain (ints: ^signflip; floats: IEEE bits)
    uint32_t valid;
    uint32_t pad;
};

constexpr int MAX_LEAVES = 8;
constexpr int MAX_AGGS = 4;
constexpr int MAX_SCAN_LEAVES = MAX_LEAVES + 1;   // the leaves of a program + the row mask of a masked scan
constexpr int MAX_POSTFIX = 2 * MAX_SCAN_LEAVES;

constexpr int CONSUMER_WARPS = 8;
constexpr int SCAN_THREADS = (CONSUMER_WARPS + 1) * 32;   // + 1 TMA producer warp
constexpr int MAX_STAGES = 8;               // ring depth is chosen per launch (2..8)

struct ScanParams {
    const PackInfo* packs;     // [npacks]
    const PackLeaf* leaves;    // [npacks][nleaves]
    const ColView*  views;     // ColView table (leaf blocks first, then agg blocks)
    const uint32_t* tile_pack; // [ntiles] pack index of each tile (nullptr: uniform packs)
    const uint64_t* set_vals;  // concatenated sorted sets
    const uint64_t* set_tabs;  // bucketised hash tables of the IN/NIN leaves (4 x u64 per bucket)
    const uint32_t* code_bits; // per (pack, leaf) dictionary-code bitmaps of this launch (LM_CODESET)
    uint8_t*  bitsets;         // nullable
    unsigned long long* counts; // [npacks] nullable
    AggPartial* partials;      // [gridDim.x][naggs]
    AggPartial* agg_out;       // [naggs] combined by the CTA that finishes last
    uint32_t* done;            // CTAs finished so far (zeroed before the launch)
    uint32_t npacks, ntiles;
    uint32_t nleaves, npost, naggs;
    uint32_t R;                // 32-row groups per warp per tile; tile rows = 256 * R; R > 32 => multiple of 32
    uint32_t stages;           // depth of the TMA ring (<= MAX_STAGES)
    uint32_t tiles_per_pack;   // uniform case
    uint32_t stage_bytes;
    uint32_t set_off[MAX_SCAN_LEAVES + 1];
    uint32_t tab_off[MAX_SCAN_LEAVES];     // LM_HASHSET: first u64 of the leaf's table in set_tabs
    uint8_t  tab_log2[MAX_SCAN_LEAVES];    //             log2(#buckets) >= 1
    // LM_HASHSET: a one-hash prefilter bitmap of 2^pre_log2 bits per leaf (bit set_hash32(v) >> (32 - pre_log2)), copied
    // into shared memory at hs_smem_off[l] (word offset behind the code bitmaps); the exact table is copied to
    // hs_tab_smem_off[l] as well when it is small (0xffffffff: looked up in global memory)
    const uint32_t* set_pre;          // concatenated prefilter bitmaps
    uint32_t pre_off[MAX_SCAN_LEAVES];     // first word of the leaf's bitmap in set_pre
    uint8_t  pre_log2[MAX_SCAN_LEAVES];    // 0: leaf has no table
    uint32_t hs_smem_off[MAX_SCAN_LEAVES];
    uint32_t hs_tab_smem_off[MAX_SCAN_LEAVES];
    uint32_t agg_view0;        // views[agg_view0 + pack * naggs + j]
    uint32_t leaf_view0;       // views[leaf_view0 + pack * nleaves + l]
    uint8_t  postfix[MAX_POSTFIX];
    uint8_t  agg_type[MAX_AGGS];
    // LM_CODESET leaves: the current pack's code bitmap of leaf l is cached in shared memory (after the ring) at word
    // code_smem_off[l]; code_smem_words = size of that area (0: the program has no such leaf)
    uint32_t code_smem_off[MAX_SCAN_LEAVES];
    uint32_t code_smem_words;
    uint32_t code_bitmap_words;       // words of that area holding per-pack code bitmaps (0: nothing to reload per pack)
    uint32_t stack_depth;      // warp kernel: slots of the per-warp AND/OR stack (trees that are not a pure AND / OR chain)
    uint32_t flat_op;          // 1: the program is a pure AND of its leaves, 2: a pure OR, 0: general tree
    uint32_t sched_chunk;      // warp kernel: tiles per scheduling chunk (spread evenly over one pack; chunks go round-robin to the warps)
    // fused reduce: a tile is read with the dense walk when matches * thr > rows; 0 = always, 0xffffffff = never
    uint32_t agg_dense_thr;
    // warp-autonomous kernel (kx_warp.cu): every warp runs its own TMA ring over tiles of 1024 w_wd rows; a ring stage holds
    // one slot per staged column of the program (leaf streams in postfix order, ALP patch-correction streams behind their leaf)
    uint32_t w_wd;             // bitset words per lane and tile (1, 2 or 4)
    uint32_t w_warps;          // warps of a CTA that take tiles
    uint32_t w_warp_bytes;     // shared memory of one warp (barriers, stage queue, tile records, match words, stack, descriptors, ring)
    uint32_t w_stage_off;      // first ring stage inside a warp's area (bytes)
    uint32_t w_ncols;          // slots per stage (>= 1; lane i of the warp issues the copy of slot i)
    uint32_t w_mw_slots;       // match-word slots / pending-tile records / value-view sets per warp (stages + 2)
    uint32_t w_chunk_rows;     // rows of a raw 64-bit value chunk streamed through a ring stage (multiple of 32)
    uint8_t  w_slot[32];       // slot -> leaf index, | 0x80: the leaf's fix stream; 0xff: nothing to copy
    uint16_t w_slot_off[32];   // slot -> offset inside a stage, in 16-byte units
    uint16_t w_col_off[MAX_SCAN_LEAVES];   // leaf -> offset of its stream inside a stage (16-byte units)
    uint16_t w_fix_off[MAX_SCAN_LEAVES];   // leaf -> offset of its fix stream
};

// can the value column be streamed through the ring (bit stream of fixed width per row)?
KX_HD inline bool agg_stageable(const ColView& v) {
    return (v.kind == CK_BITS || v.kind == CK_DICT || v.kind == CK_ALP) && v.width != 0;
}

KX_HD inline bool type_is_signed(int t) { return t >= 1 && t <= 4; }
KX_HD inline bool type_is_float(int t) { return t == 9 || t == 10; }
KX_HD inline int type_bits(int t) {
    switch (t) {
    case 1: case 5: case 9: return 64;
    case 2: case 6: case 10: return 32;
    case 3: case 7: return 16;
    case 4: case 8: return 8;
    }
    return 0;
}
// T(x): truncate to the width of T and extend back to 64 bit
KX_HD inline uint64_t type_ext(int t, uint64_t x) {
    switch (t) {
    case 2: return (uint64_t)(int64_t)(int32_t)x;
    case 3: return (uint64_t)(int64_t)(int16_t)x;
    case 4: return (uint64_t)(int64_t)(int8_t)x;
    case 6: case 10: return (uint32_t)x;
    case 7: return (uint16_t)x;
    case 8: return (uint8_t)x;
    }
    return x;
}
KX_HD inline uint64_t width_mask(int w) { return w >= 64 ? ~0ull : ((1ull << w) - 1ull); }
// 32-bit hash of a 64-bit value for the IN/NIN set structures (two multiply-adds; top bits are used)
KX_HD inline uint32_t set_hash32(uint64_t v) { return uint32_t(v) * 0x9E3779B1u + uint32_t(v >> 32) * 0x85EBCA77u; }

}  // namespace kx
