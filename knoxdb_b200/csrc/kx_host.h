// kx_host.h — host-side runtime pieces of libknoxgpu: parsing of KnoxDB's encoded column
// containers, normalisation into device ColViews and translation of filter leaves into
// per-pack PackLeaf records.  Pure C++ (no CUDA), so it is unit-testable on a CPU box.
#pragma once
#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "kx_types.h"

namespace kx {

// container ids, internal/encode/container.go:20-55
enum : int { T_CONST = 1, T_DELTA = 2, T_RUNEND = 3, T_BITPACK = 4, T_DICT = 5, T_S8B = 6, T_RAW = 7, T_FLOATALP = 13, T_FLOATALPRD = 14, T_FLOATRAW = 15,
             T_STRCONST = 16, T_STRFIXED = 17, T_STRCOMPACT = 18, T_STRDICT = 19 };
// types.FilterMode, internal/types/mode.go:14-23
enum : int { M_EQ = 1, M_NE = 2, M_GT = 3, M_GE = 4, M_LT = 5, M_LE = 6, M_IN = 7, M_NIN = 8, M_RANGE = 9 };

int  put_uvarint(uint8_t* b, uint64_t x);
int  get_uvarint(const uint8_t* b, size_t avail, uint64_t* x);   // 0 on truncated input

// parsed (not yet normalised) container tree; payload pointers alias the caller's buffer
struct Container {
    int ctype = 0, type = 0, log2 = 0;
    size_t n = 0;
    uint64_t val = 0, delta = 0;
    const uint8_t* payload = nullptr;
    size_t payload_len = 0;
    std::unique_ptr<Container> child[3];   // dict: {Dict, Codes}; runend: {Values, Ends}; alp: {Values, Patches, Positions}
    int alp_e = 0, alp_f = 0, alp_flags = 0;
};
// `depth`: nesting level (children are parsed with depth + 1; more than three levels is refused as corrupt)
long parse_container(int type, const uint8_t* buf, size_t len, std::unique_ptr<Container>& out, std::string& err, int depth = 0);
// the largest container decode_container materialises on the host (dictionaries, run values / ends, patches, string index
// arrays, legacy simple8b streams): 2^26 rows — far above any pack (<= 4 Mi rows); larger claims are refused as corrupt
constexpr size_t MAX_MATERIALIZED_ROWS = size_t(1) << 26;
bool decode_container(const Container& c, std::vector<uint64_t>& out, std::string& err);

// normalised block, ready for upload
struct BlockLayout {
    ColView view{};                     // pointers filled after upload
    const uint8_t* stream = nullptr;    // verbatim packed / raw bytes inside the caller's buffer …
    size_t stream_len = 0;
    std::vector<uint8_t> owned;         // … or bytes re-packed on the host (simple8b, exotic code children)
    std::vector<uint64_t> aux64;        // dict values (CK_DICT) or run values (CK_RUNEND → data)
    std::vector<uint32_t> aux32;        // run ends (CK_RUNEND → aux)
    std::vector<uint8_t> blob;          // CK_ALP patch blob (positions | values | patch bitmap), uploaded as aux
    // top-level Simple8b block: `stream` holds the 64-bit codewords; the device transcodes them into a fixed-width bit
    // stream at upload (width and kind of the view are filled in then)
    bool s8b = false;
};

// ALP float64 arithmetic of the reference (internal/encode/alp/{constants,encoder,decoder}.go), host side:
// used to translate float predicates into the encoded integer domain (float_alp.go:238-495)
int64_t alp_encode_single(double v, int e, int f, bool* ok);
int64_t alp_encode_above(double v, int e, int f);
int64_t alp_encode_below(double v, int e, int f);
double  alp_decode(int64_t enc, int e, int f);
int normalize_block(int type, const uint8_t* enc, size_t len, BlockLayout& out, std::string& err);

// Byte-string block (BlockBytes; containers 16..19, internal/encode/string_{const,fixed,compact,dict}.go) normalised for
// the device: the byte buffer stays verbatim, the nested uint32 containers (offsets, lengths, codes) are decoded once
// into one flat u32 index array (layout: kx_types.h STR_*).
struct StrLayout {
    ColView view{};                  // kind = CK_STR; pointers filled after upload
    const uint8_t* bytes = nullptr;  // inside the caller's buffer
    size_t nbytes = 0;
    std::vector<uint32_t> idx;
};
int normalize_string_block(const uint8_t* enc, size_t len, StrLayout& out, std::string& err);
// bytes.Compare / the seven string predicates of string_match.go on one value (host side: constant blocks)
bool string_pred(int mode, const uint8_t* v, size_t vl, const uint8_t* a, size_t al, const uint8_t* b, size_t bl);

struct LeafSpec {
    uint16_t field = 0;
    uint8_t type = 0, mode = 0;
    uint64_t a = 0, b = 0;
    std::vector<uint8_t> sa, sb; // byte-string operands (block type BYTES)
    uint32_t sa_off = 0, sb_off = 0;   // their offsets in the program's device byte pool
    std::vector<uint64_t> set;   // sorted unique
    uint32_t set_off = 0;        // offset into the program's concatenated device set array
    bool has_table = false;      // a bucketised hash table of the set exists on the device (LM_HASHSET)
    bool has_str = false;        // byte-string leaf with operand bytes (row-level predicate on CK_STR blocks)
};

// Bucketised hash table of an IN/NIN set for the device lookup (leaf_hashset in kx_scan.cu):
// 2^log2nb buckets of 4 keys; key v lives in bucket set_hash32(v) >> (32 - log2nb).
// Unused slots repeat a key of the same bucket, empty buckets hold a key of ANOTHER bucket, so
// comparing a probe with the four slots of its home bucket is exact.  Returns false (no table)
// when the set does not fit 4-key buckets at a sane size; the scan then uses the sorted array.
bool build_set_table(const std::vector<uint64_t>& set, std::vector<uint64_t>& slots, int& log2nb);
inline uint32_t set_table_bucket(uint64_t v, int log2nb) { return set_hash32(v) >> (32 - log2nb); }
// One-hash prefilter bitmap of the set (2^log2bits bits, ~256 bits per key, 1 Ki … 128 Ki bits): the scan tests it for
// every row out of shared memory and consults the exact table only for the rows that pass.
void build_set_prefilter(const std::vector<uint64_t>& set, std::vector<uint32_t>& words, int& log2bits);
// byte-string IN / NIN against a sorted set table ([count x (offset, length)] + bytes), bytes.Compare order
bool string_set_pred(int mode, const uint8_t* v, size_t vl, const uint8_t* table, uint32_t count);

// Translate one leaf for one block.  dict_host: host copy of the block's dictionary values
// (CK_DICT only).  view_index: index of the block's ColView in the launch's view table.
void compile_leaf(const ColView& v, const uint64_t* dict_host, const LeafSpec& leaf, uint32_t view_index, PackLeaf& out);

// bits per row the leaf needs staged in shared memory (0 = none)
inline int leaf_stage_width(const PackLeaf& l) { return l.data ? l.width : 0; }

// scalar predicate on a decoded value (used for constants); a/b/v sign-extended patterns
bool scalar_match(int type, int mode, uint64_t v, uint64_t a, uint64_t b);
bool set_contains(const std::vector<uint64_t>& s, uint64_t v);

// XXH3-64 (seed 0): internal/hash/xxh3.go:22-58 and hash.go:26
uint64_t xxh3_u64(uint64_t v);
uint64_t xxh3_u32(uint32_t v);
uint64_t xxh3_u16(uint16_t v);
uint64_t xxh3_u8(uint8_t v);
uint64_t xxh3_bytes(const uint8_t* p, size_t len);

size_t bitpack_bytes(int log2, size_t n);   // internal/encode/bitpack/bitpack.go:9-11

}  // namespace kx
