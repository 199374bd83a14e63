// kx_kernels.h — launch interface between the host runtime (kx_api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"

namespace kx {

struct PruneLeaf {
    uint64_t a, b;        // operands (64-bit pattern of T)
    uint64_t flip;        // sign flip making unsigned order == T order
    uint32_t set_off, nset;
    uint8_t mode, is_float, pad[6];
};

struct PruneParams {
    const uint64_t* mins;          // [npacks][nleaves]
    const uint64_t* maxs;
    const uint8_t* const* blooms;  // [npacks][nleaves] device pointers or nullptr (table may be nullptr)
    const uint64_t* bloom_len;     // bytes of each bloom buffer
    const uint64_t* hashes;        // concatenated probe hashes
    const uint64_t* set_vals;      // concatenated sorted IN sets
    uint32_t* out;                 // ceil(npacks/32) words
    unsigned long long* count;
    uint32_t npacks, nleaves, npost;
    uint32_t hash_off[MAX_LEAVES + 1];
    PruneLeaf leaves[MAX_LEAVES];
    uint8_t postfix[MAX_POSTFIX];
};

// one dictionary-set translation job of codeset_kernel: bit (c - base) of out[out_off ...] = dict[c] ∈ set
struct CodesetJob {
    const uint8_t* dict;   // u64 dictionary values on the device
    uint32_t ndict;
    uint32_t set_off, nset;   // sorted set inside the program's set_vals
    uint32_t out_off;         // first word of the bitmap
    uint32_t base;            // min-FOR base of the code stream: bit (code - base) is set, so that the scan indexes the bitmap with the raw field
    uint64_t flip;            // sign flip that maps the dictionary's T order to unsigned order
};

// one run-end leaf of runfill_kernel: runs whose value satisfies the predicate set rows [start, end] of the
// leaf bitset at out_base + out_off (zeroed before the launch)
struct RunFillJob {
    const uint8_t* vals;   // u64 run values
    const uint8_t* ends;   // u32 inclusive run ends
    uint64_t a, d, wm;     // LM_VALRANGE operands, (set offset, set size) for LM_SET, or (first run, runs - 1) for LM_RUNRANGE
    uint64_t out_off;      // byte offset of the leaf bitset
    uint32_t nruns, nrows;
    uint32_t is_set, pad;
};

// one byte-string leaf of strmatch_kernel: bit i of the leaf bitset at out_base + out_off = pred(row i)
struct StrJob {
    ColView view;          // CK_STR block
    uint64_t out_off;      // byte offset of the leaf bitset
    uint32_t a_off, a_len; // operand(s) inside the program's byte-string pool
    uint32_t b_off, b_len;
    uint32_t mode, pad;    // types.FilterMode (EQ, NE, GT, GE, LT, LE, RANGE)
};
cudaError_t launch_strmatch(const StrJob* jobs, uint32_t njobs, uint32_t max_rows, const uint8_t* pool, uint8_t* out_base, cudaStream_t stream);
cudaError_t launch_strgather_len(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                                 uint32_t* lens, unsigned long long* nbytes, cudaStream_t stream);
cudaError_t launch_strgather_copy(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                                  const uint32_t* offs, uint8_t* out, cudaStream_t stream);
cudaError_t launch_exclusive_scan(uint32_t* v, uint32_t n, unsigned long long* total, cudaStream_t stream);

// one leaf of valmatch_kernel: bit i of the leaf bitset at out_base + out_off = float predicate on the value decoded at row i
// (ALP-RD blocks: FloatAlpRdContainer.Match*, internal/encode/float_alprd.go:181-211)
struct ValJob {
    ColView view;
    uint64_t a, b;         // operands (IEEE bits of the block's float type)
    uint64_t out_off;      // byte offset of the leaf bitset
    uint32_t mode, pad;    // types.FilterMode (EQ, NE, GT, GE, LT, LE, RANGE)
};
cudaError_t launch_valmatch(const ValJob* jobs, uint32_t njobs, uint32_t max_rows, uint8_t* out_base, cudaStream_t stream);

// one ALP leaf of alpfix_kernel: bit pos[k] of the correction stream = pred(patch value k) (invert: !pred)
struct AlpFixJob {
    const uint8_t* blob;   // patch blob of the block (positions | values | bitmap)
    uint64_t a, b;         // float64 operands (IEEE bits)
    uint64_t out_off;      // byte offset of the correction stream
    uint32_t np, mode;     // types.FilterMode of the float predicate (EQ also serves NE)
    uint32_t invert, pad;
};

// time-bucketed reduce (kx_bucket.cu): one table cell per (window, 1 + value column); cell 0 of a window = match count
struct BucketCell {
    uint64_t count;
    uint64_t sum;      // integer sum mod 2^64, or IEEE bits of the float64 sum
    uint64_t mn, mx;   // order-preserving unsigned keys (ints: ^ sign flip; floats: f64_key); identities ~0 / 0
};
constexpr int BUCKET_THREADS = 128;
constexpr uint32_t BUCKET_JOB_GROUPS = 256;  // consecutive 32-row groups per warp job (8192 rows)
struct BucketParams {
    const PackInfo* packs;      // [npacks] (bitset_off = the pack's bitset inside `bits`)
    const uint8_t*  bits;       // device-resident match bitsets of the scan
    const ColView*  views;      // [npacks][1 + naggs]: timestamp column, then the value columns
    const uint64_t* edges;      // [nbuckets + 1] ascending window edges as order keys (value ^ ts_flip)
    const uint32_t* job0;       // [npacks + 1] first warp job of every pack
    BucketCell*     table;      // [nbuckets][1 + naggs]
    uint64_t ts_flip;
    uint32_t npacks, njobs, nbuckets, naggs;
    uint8_t  agg_type[MAX_AGGS];
};
cudaError_t launch_bucket(const BucketParams& P, int num_sms, cudaStream_t stream);

constexpr size_t SCAN_MAX_DYN_SMEM = 200 * 1024;   // dynamic shared memory the scan kernel may ask for
constexpr size_t WARP_MAX_DYN_SMEM = 224 * 1024;   // … and the warp-autonomous kernel (one CTA per SM)

// pruning over a device-resident statistics index (kx_stats): statistics are column-major [field][pack]
struct PruneStatsParams {
    const uint64_t* mins;
    const uint64_t* maxs;
    const uint64_t* bloom_ptr;     // [field][pack] device address of the filter's bit array, 0 = none
    const uint32_t* bloom_mask;    // m - 1
    const uint8_t*  bloom_k;
    const uint64_t* hashes;        // concatenated probe hashes
    const uint64_t* set_vals;      // concatenated sorted IN sets
    uint32_t* out;                 // ceil(npacks/32) words
    unsigned long long* count;
    uint32_t npacks, nleaves, npost;
    uint32_t hash_off[MAX_LEAVES + 1];
    uint8_t  leaf_field[MAX_LEAVES];   // index of the leaf's column in the index
    uint8_t  leaf_nozone[MAX_LEAVES];  // column carries no zone map (byte strings): filter only
    PruneLeaf leaves[MAX_LEAVES];
    uint8_t postfix[MAX_POSTFIX];
};
cudaError_t launch_prune_stats(const PruneStatsParams& P, cudaStream_t stream);
cudaError_t launch_bloom_build(const uint8_t* values, const uint32_t* offsets, uint64_t n, int elem_bytes, uint32_t* bits, uint32_t mask,
                               uint32_t k, cudaStream_t stream);

// single-leaf scans without aggregates (kx_scan.cu) / everything else (kx_general.cu)
cudaError_t launch_scan(const ScanParams& P, int grid, size_t smem_bytes, bool only32, int ctas_per_sm, cudaStream_t stream);
cudaError_t launch_scan_warp(const ScanParams& P, int grid, size_t smem_bytes, cudaStream_t stream);   // kx_warp.cu: one CTA per SM
cudaError_t launch_alpfix(const AlpFixJob* jobs, uint32_t njobs, uint32_t max_patches, uint8_t* out_base, cudaStream_t stream);
cudaError_t launch_runfill(const RunFillJob* jobs, uint32_t njobs, uint32_t max_runs, const uint64_t* set_vals, uint8_t* out_base, cudaStream_t stream);
cudaError_t launch_codeset(const CodesetJob* jobs, uint32_t njobs, uint32_t max_set, const uint64_t* set_vals, uint32_t* out, cudaStream_t stream);
cudaError_t launch_bitset_op(uint32_t* dst, const uint32_t* src, uint64_t nbits, int op, unsigned int* flags, cudaStream_t stream);
cudaError_t launch_bitset_neg(uint32_t* buf, uint64_t nbits, cudaStream_t stream);
cudaError_t launch_bitset_popcount(const uint32_t* buf, uint64_t nbits, unsigned long long* out, cudaStream_t stream);
cudaError_t launch_bitset_indexes(const uint32_t* buf, uint64_t nbits, uint32_t* block_tmp, unsigned long long* total,
                                  uint32_t* dst, cudaStream_t stream);
cudaError_t launch_select(const PackInfo* packs, uint32_t npacks, const uint8_t* bits, uint64_t total_words, uint32_t* block_tmp,
                          unsigned long long* total, uint32_t* dst, cudaStream_t stream);
cudaError_t launch_gather(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                          int elem_bytes, void* dst, cudaStream_t stream);
cudaError_t launch_decode(const ColView& v, void* dst, cudaStream_t stream);
// Simple8b transcode at registration: counts[i] = values of codeword i → exclusive row offsets (in place), *maxbits = widest
// value, *total = values in the stream; then every value is written into a zeroed `width`-bit LSB-first stream
cudaError_t launch_s8b_count(const void* words, uint32_t nwords, uint32_t* counts, uint32_t* maxbits, unsigned long long* total, cudaStream_t stream);
cudaError_t launch_s8b_pack(const void* words, uint32_t nwords, const uint32_t* offs, uint32_t nrows, uint32_t width, void* out, cudaStream_t stream);
cudaError_t launch_prune(const PruneParams& P, cudaStream_t stream);

}  // namespace kx
