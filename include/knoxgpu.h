/*
 * knoxgpu.h — C ABI of libknoxgpu.so, the B200 (sm_100a) implementation of KnoxDB's
 * pack-engine scan path: decode compressed column blocks → evaluate filter predicates
 * into LSB-first bitsets → reduce matching rows to count/sum/min/max.
 *
 * The reference has no FFI for this path; the seams a replacement sits behind are Go
 * interfaces.  Each entry point names the reference interface it
 * replaces (paths relative to the knoxdb repository root).  INTEGRATION.md shows the
 * cgo stubs that bind these symbols behind those interfaces.
 *
 * Conventions
 *   - plain pointers + sizes, no C++/torch types; all functions are re-entrant per ctx
 *   - return 0 on success, <0 (KX_E*) on error; kx_last_error(ctx) gives the message
 *   - host pointers are never retained after return (cgo rule); kx_block_put copies
 *   - bitsets: bit i of a pack ↔ byte i>>3, bit i&7 (internal/bitset/bitset.go:23-29),
 *     ceil(n/8) bytes, tail bits zero
 *   - element types are types.BlockType values, modes are types.FilterMode values
 *   - there is NO CPU fallback: every call fails with KX_ENODEV without a CUDA device
 */
#ifndef KNOXGPU_H
#define KNOXGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KX_ABI_VERSION 2

/* error codes */
enum {
    KX_OK = 0, KX_EINVAL = -1, KX_ENODEV = -2, KX_ENOMEM = -3, KX_ECUDA = -4,
    KX_ENOTFOUND = -5, KX_EFORMAT = -6, KX_EUNSUPPORTED = -7,
};

/* internal/types/block.go:20-36 (BlockType) */
enum {
    KX_INT64 = 1, KX_INT32 = 2, KX_INT16 = 3, KX_INT8 = 4,
    KX_UINT64 = 5, KX_UINT32 = 6, KX_UINT16 = 7, KX_UINT8 = 8,
    KX_FLOAT64 = 9, KX_FLOAT32 = 10,
    KX_BYTES = 12,   /* byte strings (BlockBytes): string containers 16..19, row-level predicates, bloom probes / build */
};

/* internal/types/mode.go:14-23 (FilterMode) */
enum {
    KX_MODE_EQ = 1, KX_MODE_NE = 2, KX_MODE_GT = 3, KX_MODE_GE = 4, KX_MODE_LT = 5,
    KX_MODE_LE = 6, KX_MODE_IN = 7, KX_MODE_NIN = 8, KX_MODE_RANGE = 9,
};

/* postfix program opcodes (filter tree, internal/operator/filter/node.go): a byte < 0x80
 * pushes the bitset of leaf <byte>; KX_OP_AND / KX_OP_OR combine the two topmost. */
#define KX_OP_AND 0xFE
#define KX_OP_OR  0xFF
#define KX_MAX_LEAVES 8
#define KX_MAX_AGGS 4

typedef struct kx_ctx kx_ctx;
typedef struct kx_prog kx_prog;

/* ---------------------------------------------------------------- lifecycle
 * New (no reference counterpart; closest: engine options, pkg/knox/interface.go:29-50).
 * device: CUDA ordinal; hbm_budget: max bytes of resident blocks (0 = 80% of free HBM). */
int  kx_abi_version(void);
int  kx_device_count(void);                /* usable CUDA devices (0 without a driver / device) */
int  kx_ctx_create(int device, size_t hbm_budget, kx_ctx** out);
void kx_ctx_destroy(kx_ctx* ctx);
const char* kx_last_error(kx_ctx* ctx);   /* ctx may be NULL: error of the last failed kx_ctx_create */

/* pinned host memory for encoded blocks / result buffers that cross PCIe at full speed
 * (replaces internal/arena allocations for buffers handed to the scan) */
void* kx_host_alloc(kx_ctx* ctx, size_t bytes);
void  kx_host_free(kx_ctx* ctx, void* p);

/* ---------------------------------------------------------------- device pack store
 * Immutable (pack key, version, field id) → encoded block, exactly the bytes produced by
 * Container.Store (internal/encode/int_*.go, float_raw.go) WITHOUT the outer compression
 * byte of block.Encode (internal/block/encode.go:194-226; outer s2/lz4/zstd is undone on the
 * host).  Replaces Package.LoadFromDisk + block.Decode → encode.LoadInt/LoadFloat
 * (internal/pack/storage.go:128-190, internal/encode/int.go:109-115) as the source of
 * column vectors for the scan.  Copies `len` bytes; returns the block's row count.
 * KX_BYTES blocks: the string containers of internal/encode/string_{const,fixed,compact,dict}.go (ids 16..19). */
int kx_block_put(kx_ctx* ctx, uint32_t pack, uint32_t version, uint16_t field, uint8_t block_type,
                 const void* enc, size_t len, uint32_t* nrows_out);
int kx_block_drop(kx_ctx* ctx, uint32_t pack, uint32_t version, uint16_t field);
/* device_bytes = bytes of the live blocks (256 B granules); slab_bytes = device memory the store really holds (capacity of
 * its slabs: freed extents are coalesced and reused, a slab returns to CUDA when its last block is dropped); the HBM
 * budget of kx_ctx_create bounds slab_bytes.  Any out pointer may be NULL. */
int kx_store_stats(kx_ctx* ctx, uint64_t* nblocks, uint64_t* encoded_bytes, uint64_t* device_bytes, uint64_t* slab_bytes);

/* ---------------------------------------------------------------- predicate program
 * One leaf = one filter.Filter (internal/operator/filter/filter.go) with its Matcher
 * (match_num.go:327-817): field, type, mode, operand(s).  a/b carry the operand as the
 * 64-bit pattern of the type (sign-extended ints, IEEE bits for floats); RANGE uses [a,b].
 * IN/NIN pass the set flattened to u64 values (xroar.Bitmap contents, `uint64(v)` of each
 * member, internal/encode/int_raw.go:339-357); order and duplicates do not matter.
 * Byte-string leaves (block_type KX_BYTES; types.StringMatcher, internal/encode/string_match.go:13-188: the seven
 * scalar modes, bytes.Equal / bytes.Compare row by row): nset = 1, `set` points to the operand BYTES, a = length of
 * the operand, and for RANGE b = length of the upper bound that follows it.  IN / NIN on byte strings
 * (bytesInSetMatcher / bytesNotInSetMatcher, internal/operator/filter/match_bytes.go:392-520: exact membership in the
 * de-duplicated set): nset = number of strings, `set` points to nset little-endian uint32 lengths followed by the
 * concatenated string bytes, a = size of that buffer in bytes.  A KX_BYTES leaf with nset = 0 carries
 * no operand and can only be used for pruning (bloom probes with caller-supplied hashes). */
typedef struct kx_leaf {
    uint16_t field;
    uint8_t  block_type;
    uint8_t  mode;
    uint32_t nset;
    uint64_t a, b;
    const uint64_t* set;
} kx_leaf;

/* Replaces the compiled filter tree that filter.Match walks (match_core.go:14-215). */
int  kx_prog_compile(kx_ctx* ctx, const kx_leaf* leaves, int nleaves,
                     const uint8_t* postfix, int npost, kx_prog** out);
void kx_prog_free(kx_prog* prog);

/* ---------------------------------------------------------------- scan
 * Replaces the per-pack hot loop Reader.nextQueryMatch → filter.Match → bits.Indexes /
 * CountResult.Append / StreamResult.Append + Reducer.Reduce
 * (internal/pack/table/reader.go:288-450, query/result.go:44-51,96-152,
 * reducer/reducer.go:138-314) for a whole batch of packs in one call.
 *
 * aggregates: count/sum/min/max over the matching rows of one value column each. */
typedef struct kx_packref { uint32_t pack, version; } kx_packref;
typedef struct kx_agg_req { uint16_t field; uint8_t block_type; uint8_t reserved; } kx_agg_req;
typedef struct kx_agg_out {
    int64_t  count;       /* CountReducer */
    uint64_t sum_bits;    /* SumReducer: ints wrap in T (pattern sign-extended to 64 bit); float64: IEEE bits */
    double   sum_err;     /* float64: compensation term (sum = hi + err), 0 for ints */
    uint64_t min_bits;    /* MinReducer / MaxReducer as 64-bit pattern of T */
    uint64_t max_bits;
    int32_t  valid;       /* 0 when no row matched (reducer Value() ok == false) */
    int32_t  reserved;
} kx_agg_out;

/* bitsets (nullable): caller buffer; pack i's bitset starts at bitset_off[i] (multiple of
 * 8) and spans ceil(n_i/8) bytes.  counts (nullable): npacks match counts.  agg_out
 * (naggs entries) is combined over all packs of the call in pack order. */
int kx_scan(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks,
            uint8_t* bitsets, const size_t* bitset_off, int64_t* counts,
            const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out);

/* kx_scan with every option in one argument block — and the one the reader binding uses, because it takes ROW MASKS:
 * between filter.Match and bits.Indexes / aggregation the reference's reader clears the bits of rows that a live
 * journal has tombstoned or that the transaction may not see (internal/pack/table/reader.go:347-413: `hits` is
 * AND-NOT'ed with the exclude mask from engine.TableReader.WithMask, engine/interface.go:96-106, and with the rows
 * whose $xmin / $xmax fail the snapshot).  row_masks (nullable) holds per pack one bitset of ceil(n/8) bytes in HOST
 * memory, or NULL for "all rows": bit i SET = row i stays eligible (the caller passes ¬(tombstones ∪ invisible)).  The
 * mask is ANDed into the filter result on the device, so counts, bitsets, selection vectors and the fused aggregates all
 * see it.  (Alternative without a mask: add `$rid NIN {tombstoned rids}` and range leaves on `$xmin` / `$xmax` to the
 * program — INTEGRATION.md §4.)
 * Outputs are all nullable: bitsets + bitset_off as in kx_scan; counts; sel / sel_cap / sel_off as in kx_scan_select (not
 * together with bitsets); aggs / agg_out; flags & KX_SCAN_SHARDED combines agg_out and *total_count over all ranks of
 * the communicator like kx_scan_sharded. */
#define KX_SCAN_SHARDED 1u
typedef struct kx_scan_args {
    uint32_t struct_size;               /* sizeof(kx_scan_args) */
    uint32_t flags;
    const kx_packref* packs;
    int32_t npacks, naggs;
    const uint8_t* const* row_masks;    /* [npacks], entries may be NULL */
    uint8_t* bitsets; const size_t* bitset_off;
    int64_t* counts;
    uint32_t* sel; size_t sel_cap; uint64_t* sel_off;
    const kx_agg_req* aggs; kx_agg_out* agg_out;
    int64_t* total_count;               /* KX_SCAN_SHARDED */
} kx_scan_args;
int kx_scan_ex(kx_ctx* ctx, const kx_prog* prog, const kx_scan_args* args);

/* kx_scan that hands back SELECTION VECTORS instead of bitsets: what the reader and PhysicalFilter do with a match
 * bitset — sel := bits.Indexes(hits); pack.WithSelection(sel) (internal/pack/table/reader.go:432-436,
 * internal/operator/filter.go:29-37, Bitset.Indexes internal/bitset/iterator.go:269-290).  The bitsets stay on the
 * device; sel receives the ascending row ids of every pack's matches, concatenated in pack order: pack i owns
 * sel[sel_off[i] .. sel_off[i+1]).  sel_off has npacks + 1 entries.  If more than sel_cap ids match, nothing is
 * written to sel, KX_ENOMEM is returned and sel_off[npacks] holds the required capacity. */
int kx_scan_select(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks,
                   uint32_t* sel, size_t sel_cap, uint64_t* sel_off, int64_t* counts,
                   const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out);

/* NumberContainer.AppendTo(dst, sel) for a batch of packs (internal/encode/int_*.go, float_raw.go, float_alp.go
 * AppendTo with a selection; query/result.go:196-264 copies the selected rows of the result columns): decode the
 * selected rows of `field` straight from the resident encoded blocks.  sel / sel_off as produced by
 * kx_scan_select; dst receives sel_off[npacks] elements of block_type. */
int kx_gather(kx_ctx* ctx, const kx_packref* packs, int npacks, uint16_t field, uint8_t block_type,
              const uint32_t* sel, const uint64_t* sel_off, void* dst);

/* StringContainer.AppendTo(dst, sel) for a batch of packs: the byte-string twin of kx_gather
 * (internal/encode/string_{const,fixed,compact,dict}.go AppendTo with a selection; query/result.go:196-264 copies the
 * selected rows of bytes columns like any other result column).  Row i of the selection (i < sel_off[npacks]) occupies
 * out[out_off[i] .. out_off[i+1]); out_off has sel_off[npacks] + 1 entries.  If the rows hold more than out_cap bytes
 * nothing is written to out, KX_ENOMEM is returned and out_off[sel_off[npacks]] holds the required capacity (the
 * kx_scan_select convention; out may be NULL with out_cap = 0 to ask for the size). */
int kx_gather_bytes(kx_ctx* ctx, const kx_packref* packs, int npacks, uint16_t field,
                    const uint32_t* sel, const uint64_t* sel_off, uint64_t* out_off, uint8_t* out, size_t out_cap);

/* Scan + TIME-BUCKETED reduce (group by time window): what a series query does with every streamed row —
 * t = Interval.TruncateRelative(ts, Range.From) picks the window, Bucket.Push feeds the window's reducer
 * (pkg/series/series.go:192-256, internal/reducer/bucket_native.go:104-167) — for the order-independent reducers
 * count / sum / min / max (internal/reducer/reducer.go:138-297; mean = sum / count).  The caller computes the window
 * edges with TimeUnit.Next (pkg/util/timeunit.go:234-263; calendar units give irregular windows): `edges` holds
 * nbuckets + 1 ascending values of the timestamp column's type as 64-bit patterns, window k = [edges[k], edges[k+1]);
 * matching rows outside [edges[0], edges[nbuckets]) belong to no window.  bucket_counts (nullable): nbuckets match
 * counts.  out: naggs x nbuckets results, out[j * nbuckets + k] = value column j in window k (integer results are
 * exact and order-independent; float64 sums are compensated per thread and combined with atomic adds).  counts
 * (nullable): npacks per-pack match counts of the filter, as in kx_scan. */
int kx_scan_buckets(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks,
                    uint16_t ts_field, uint8_t ts_type, const uint64_t* edges, int nbuckets,
                    const kx_agg_req* aggs, int naggs, int64_t* bucket_counts, kx_agg_out* out, int64_t* counts,
                    const uint8_t* const* row_masks /* nullable: see kx_scan_ex */);

/* Same scan over blocks that still live in HOST memory (cold device cache): the blocks of
 * all referenced fields are uploaded, scanned and dropped in pipelined batches.
 * blocks[i*nfields + f] / block_len[...] = encoded block of pack i, field fields[f] (any block type kx_block_put
 * takes, KX_BYTES included). */
int kx_scan_host(kx_ctx* ctx, const kx_prog* prog, int npacks,
                 const uint16_t* fields, const uint8_t* field_types, int nfields,
                 const void* const* blocks, const size_t* block_len,
                 uint8_t* bitsets, const size_t* bitset_off, int64_t* counts,
                 const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out);

/* Deterministic combine of per-shard partial aggregates (multi-GPU: the 64-byte partials
 * are exchanged with ONE NCCL all-gather and combined in rank order). */
int kx_agg_combine(uint8_t block_type, const kx_agg_out* parts, int nparts, kx_agg_out* out);

/* Timing of the last kx_scan / kx_scan_host on this ctx (CUDA events on the scan stream):
 * kernel_ms = device time of the scan kernels only, total_ms = first copy/launch → results
 * on host.  launches = kernels launched. */
int kx_last_scan_stats(kx_ctx* ctx, double* kernel_ms, double* total_ms, int* launches);

/* Test hook.  With KX_GUARD=1 in the environment (read when the library is loaded) every device scratch / result buffer of
 * the library is allocated exactly as large as needed and followed by 256 bytes of 0xFA (the reference poisons the slack of
 * its test outputs the same way, internal/cmp/tests/gen.go:13-44); this call returns how many of those zones were
 * overwritten (0 = no kernel or copy wrote past the end of a buffer), KX_EUNSUPPORTED when guard mode is off. */
int kx_debug_check_guards(kx_ctx* ctx);

/* The counters a query reports through QueryStats (internal/query/stats.go:15-60: rows_scanned, packs_scanned,
 * rows_matched, scan_time …) for the last kx_scan* call on this ctx, so that the Go adapter can feed
 * stats.Count / stats.Tick with what the device did (one call covers what the reference counts pack by pack in
 * reader.go:330-345). */
typedef struct kx_query_stats {
    uint64_t rows_scanned;     /* rows of all packs of the call */
    uint64_t packs_scanned;
    uint64_t rows_matched;     /* sum of the per-pack match counts */
    uint64_t scan_time_ns;     /* device time of the scan kernels (CUDA events) */
    uint64_t total_time_ns;    /* first copy / launch → results on the host */
    uint32_t kernel_launches;
    uint32_t reserved;
} kx_query_stats;
int kx_last_query_stats(kx_ctx* ctx, kx_query_stats* out);

/* ---------------------------------------------------------------- multi-GPU: pack-sharded scans
 * Packs are independent (internal/pack/table/reader.go:299-449 keeps no cross-pack state): a table is sharded by pack
 * key over the GPUs of one box — one kx_ctx per GPU, blocks registered on the owning GPU only, no data-path collective.
 * Per query the ranks exchange ONE 208-byte record each (match count + partial aggregates) with one NCCL all-gather
 * that the library enqueues on the scan stream behind the scan kernel, and combine the records in rank order on the
 * device: every rank returns the bit-identical total.  NCCL is bound at run time (libnccl.so.2; KX_NCCL_LIB overrides).
 *
 * kx_comm_unique_id: rank 0 creates the communicator id (ncclGetUniqueId) and hands its KX_COMM_ID_BYTES to the other
 *                    ranks by any means (the Go host: over the channel that starts its per-GPU workers).
 * kx_comm_init:      every rank, once per ctx (collective: returns when all ranks joined).  nranks = 1 needs no id.
 * kx_scan_sharded:   kx_scan over this rank's packs (possibly none); counts (nullable) = this rank's per-pack counts;
 *                    agg_out / total_count = combined over ALL ranks.  Collective: every rank of the communicator must
 *                    call it for every query, in the same order.
 * kx_comm_allgather: the same exchange for caller-defined partials (e.g. the window table of a sharded
 *                    kx_scan_buckets): bytes from every rank, concatenated in rank order into recv (host buffers). */
#define KX_COMM_ID_BYTES 128
int kx_comm_unique_id(void* id_out, size_t cap);
int kx_comm_init(kx_ctx* ctx, int nranks, int rank, const void* id, size_t id_len);
int kx_comm_info(kx_ctx* ctx, int* nranks, int* rank, int* nccl_version);
int kx_scan_sharded(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, int64_t* counts,
                    const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out, int64_t* total_count);
int kx_comm_allgather(kx_ctx* ctx, const void* send, void* recv, size_t bytes);

/* ---------------------------------------------------------------- narrow drop-ins
 * Host-pointer kernels with the signatures of the reference's leaf functions; used behind
 * the assignable function variables / NumberMatcher methods and by the parity tests.
 * Each call uploads, runs the same CUDA kernels as kx_scan, and downloads.
 *
 * kx_cmp: cmp.<Type><Op>(src []T, val T, bits []byte) int64 and <Type>Between(src, a, b, bits)
 *         (internal/cmp/cmp.go:6-114, number.go:13-243, float.go:13-242).  mode ∈ EQ..LE, RANGE. */
int64_t kx_cmp(kx_ctx* ctx, uint8_t block_type, uint8_t mode, const void* src, size_t n,
               uint64_t a, uint64_t b, uint8_t* bits);

/* kx_bitpack_cmp: bitpack.Equal/NotEqual/Less/LessEqual/Greater/GreaterEqual/Between
 *         (buf, log2, val[, val2], n, bits) (internal/encode/bitpack/cmp.go:20-46);
 *         operands are already in the min-FOR domain. Returns the match count. */
int64_t kx_bitpack_cmp(kx_ctx* ctx, uint8_t mode, const void* packed, int log2, uint64_t a, uint64_t b,
                       size_t n, uint8_t* bits);

/* kx_bitpack_decode: bitpack.Decode(dst, buf, log2, minv) (bitpack/decode.go:131-208) into
 *         n values of block_type. */
int kx_bitpack_decode(kx_ctx* ctx, uint8_t block_type, const void* packed, int log2, uint64_t minv,
                      size_t n, void* dst);

/* kx_container_match: types.NumberMatcher[T].Match{Equal..Between,InSet,NotInSet}(val, bits, mask)
 *         on an encoded container (internal/types/number.go:37-47, internal/block/access.go:42-44,
 *         internal/encode/int_*.go Match*).  bits: ceil(n/8) bytes, overwritten. Returns count. */
int64_t kx_container_match(kx_ctx* ctx, uint8_t block_type, const void* enc, size_t len, uint8_t mode,
                           uint64_t a, uint64_t b, const uint64_t* set, uint32_t nset, uint8_t* bits);

/* kx_container_decode: NumberContainer[T].AppendTo(dst, nil) (internal/encode/int_*.go) */
int kx_container_decode(kx_ctx* ctx, uint8_t block_type, const void* enc, size_t len, void* dst, size_t dst_cap_rows);

/* bitset.{And,AndFlag,AndNot,Or,OrFlag,Xor,Neg,PopCount} and Bitset.Indexes
 * (internal/bitset/bitset.go:400-531, generic/bitset.go:13-396, iterator.go:269-290) */
enum { KX_BIT_AND = 0, KX_BIT_ANDNOT = 1, KX_BIT_OR = 2, KX_BIT_XOR = 3 };
int     kx_bitset_op(kx_ctx* ctx, int op, uint8_t* dst, const uint8_t* src, size_t nbits, int* any, int* all);
int     kx_bitset_neg(kx_ctx* ctx, uint8_t* buf, size_t nbits);
int64_t kx_bitset_popcount(kx_ctx* ctx, const uint8_t* buf, size_t nbits);
int64_t kx_bitset_indexes(kx_ctx* ctx, const uint8_t* buf, size_t nbits, uint32_t* dst);

/* ---------------------------------------------------------------- pruning
 * Zone-map + bloom pruning of candidate packs, replacing stats.matchVector →
 * Matcher.MatchRangeVectors + bloom.Filter.Contains per candidate
 * (internal/pack/stats/match.go:92-195, operator/filter/match_num.go MatchRangeVectors,
 * internal/filter/bloom/bloom.go:136-150,182-184).
 * mins/maxs: npacks × nleaves 64-bit patterns (row-major: pack, leaf) of the leaf's column.
 * blooms (nullable): npacks × nleaves pointers to bloom buffers ([k][m/8 bytes]) or NULL;
 * bloom_len their byte lengths; hashes: per leaf the XXH3-64 of the probe value(s)
 * (EQ: 1 hash; IN: nset hashes, concatenated; hash_off[nleaves+1]).
 * out: ceil(npacks/8) bytes, bit set = pack may match. Returns number of surviving packs. */
int64_t kx_prune(kx_ctx* ctx, const kx_prog* prog, int npacks,
                 const uint64_t* mins, const uint64_t* maxs,
                 const void* const* blooms, const size_t* bloom_len,
                 const uint64_t* hashes, const uint32_t* hash_off, uint8_t* out);

/* ---------------------------------------------------------------- resident statistics index
 * The same pruning over a statistics index that LIVES on the device: per data pack and column the zone map
 * (min, max — the min/max columns of the reference's statistics packs, internal/pack/stats/index.go) and
 * optionally a bloom filter (the buffers stored under encodeFilterKey, internal/pack/stats/filter.go:26-32).
 * One kx_stats covers any number of data packs (a reference statistics pack holds <= 2048).
 * mins/maxs are COLUMN-major: [nfields][npacks] 64-bit patterns; byte-string columns (KX_BYTES) carry no
 * zone map here (their entries are ignored) and are pruned by their filters only. */
typedef struct kx_stats kx_stats;
int  kx_stats_create(kx_ctx* ctx, int npacks, const uint16_t* fields, const uint8_t* field_types, int nfields,
                     const uint64_t* mins, const uint64_t* maxs, kx_stats** out);
void kx_stats_free(kx_stats* stats);
/* attach a stored filter ([k][m/8 bytes], m a power of two: bloom.NewFilterBuffer, bloom.go:83-100); copies */
int  kx_stats_put_bloom(kx_stats* stats, int field_index, int pack_index, const void* bloom, size_t len);
/* stats.BuildBloomFilter (internal/pack/stats/filter.go:296-367) on the device: m = pow2(cardinality*factor*8)
 * bits, k = 4, every value hashed with XXH3-64 (hash.Vec64/Vec32/Vec16/Vec8 for fixed-width types, hash.Hash
 * for byte strings) and added (bloom.Filter.Add).  values: n elements of block_type in host memory; for
 * KX_BYTES the concatenated strings with offsets[n+1]. The result is bit-identical to the reference's buffer. */
int  kx_stats_build_bloom(kx_stats* stats, int field_index, int pack_index, uint8_t block_type, const void* values,
                          const uint32_t* offsets, size_t n, int cardinality, int factor);
/* read a filter back as [k][m/8 bytes] (to persist it like the reference does); out may be NULL to query len */
int  kx_stats_get_bloom(kx_stats* stats, int field_index, int pack_index, void* out, size_t cap, size_t* len);
/* kx_prune over the resident index: no per-pack host work, no uploads besides the probe hashes.  hashes /
 * hash_off as in kx_prune; pass NULL/NULL to let the library hash the numeric EQ / IN operands itself
 * (byte-string leaves then probe nothing).  out: ceil(npacks/8) bytes.  Returns the surviving packs. */
int64_t kx_prune_stats(kx_ctx* ctx, const kx_prog* prog, kx_stats* stats, const uint64_t* hashes, const uint32_t* hash_off,
                       uint8_t* out);

/* hash.Uint64/Uint32/Uint16/Uint8 and hash.Hash ([]byte) (internal/hash/hash.go:26,67-92,
 * xxh3.go:22-58): XXH3-64, seed 0.  Computed on the host side of the library (one hash per
 * probe value per query). */
uint64_t kx_hash_value(uint8_t block_type, uint64_t pattern);
uint64_t kx_hash_bytes(const void* p, size_t len);

#ifdef __cplusplus
}
#endif
#endif /* KNOXGPU_H */
