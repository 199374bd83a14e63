// kx_api.cu — C ABI of libknoxgpu.so (see include/knoxgpu.h): context, device pack store,
// predicate programs, batched scans and the narrow drop-in entry points.
//
// There is no CPU fallback anywhere in this file: every entry point needs a CUDA device and
// fails with KX_ENODEV / KX_ECUDA otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <tuple>
#include <unordered_map>
#include <vector>

#include "../../include/knoxgpu.h"
#include "kx_comm.h"
#include "kx_host.h"
#include "kx_kernels.h"
#include "kx_types.h"

using namespace kx;

namespace {

thread_local std::string g_create_error;

constexpr size_t STREAM_PAD = 64;      // slack after every bit stream (TMA 16 B rounding, idx+2 over-read)
constexpr size_t SLAB_BYTES = 256ull << 20;

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Guard mode (KX_GUARD=1 in the environment when the library is loaded; tests only — compute-sanitizer is not available on
// every pool): device scratch buffers are allocated EXACTLY as large as asked, followed by a 256 B zone of 0xFA, and
// kx_debug_check_guards() verifies the zones: a kernel or copy that writes past the end of a result / scratch buffer is caught
// (the reference poisons the slack of its own test outputs the same way, internal/cmp/tests/gen.go:13-44).
constexpr size_t GUARD_BYTES = 256;
bool guard_mode() { static const bool g = getenv("KX_GUARD") && atoi(getenv("KX_GUARD")) != 0; return g; }
struct DevBuf;
std::vector<DevBuf*>& guarded_bufs() { static std::vector<DevBuf*> v; return v; }
std::mutex& guarded_mu() { static std::mutex m; return m; }

// grow-only device / pinned buffers
struct DevBuf {
    void* p = nullptr; size_t cap = 0, asked = 0;
    cudaError_t reserve(size_t n) {
        if (guard_mode()) {
            if (p && n == asked) return cudaSuccess;
            if (p) cudaFree(p); else { std::lock_guard<std::mutex> lk(guarded_mu()); guarded_bufs().push_back(this); }
            p = nullptr; cap = 0; asked = n;
            cudaError_t e = cudaMalloc(&p, n + GUARD_BYTES);
            if (e == cudaSuccess) { cap = n; e = cudaMemset(p, 0, n); }
            if (e == cudaSuccess) e = cudaMemset(static_cast<uint8_t*>(p) + n, 0xFA, GUARD_BYTES);
            if (e == cudaSuccess) e = cudaDeviceSynchronize();
            return e;
        }
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = round_up(std::max(n, size_t(4096)) * 5 / 4, 4096);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) { cap = want; e = cudaMemset(p, 0, want); }
        // cudaMemset runs on the legacy default stream, asynchronously to the host, and the context's
        // stream is non-blocking: without this the zero-fill can land AFTER the first copies into the buffer
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        return e;
    }
    // guard mode: is the zone behind the buffer intact?
    bool guard_intact() const {
        if (!guard_mode() || !p) return true;
        uint8_t z[GUARD_BYTES];
        if (cudaMemcpy(z, static_cast<const uint8_t*>(p) + asked, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        for (uint8_t b : z) if (b != 0xFA) return false;
        return true;
    }
    ~DevBuf() {
        if (guard_mode()) { std::lock_guard<std::mutex> lk(guarded_mu()); auto& v = guarded_bufs(); v.erase(std::remove(v.begin(), v.end(), this), v.end()); }
        if (p) cudaFree(p);
    }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = round_up(std::max(n, size_t(4096)) * 5 / 4, 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    ~PinBuf() { if (p) cudaFreeHost(p); }
};

// slab allocator for resident blocks: 256 B aligned extents carved first-fit out of 256 MB slabs; freed extents are
// coalesced and reused (pack versions come and go: kx_block_put / kx_block_drop churn), a slab is returned to CUDA when
// its last extent is released.  The HBM budget is enforced on the slabs' CAPACITY (what is really taken from the device).
struct Slab {
    uint8_t* base = nullptr; size_t cap = 0, live = 0;
    std::map<size_t, size_t> free;   // offset -> bytes of every free extent (coalesced)
};
struct SlabAlloc { int slab = -1; size_t off = 0, bytes = 0; };

struct StoredBlock {
    ColView view{};
    std::vector<uint64_t> dict;     // host copy of dictionary values (leaf translation)
    std::vector<uint8_t> cstr;      // host copy of a constant string block's value (leaf translation)
    std::vector<SlabAlloc> allocs;
    size_t enc_len = 0;
};

struct BlockKey {
    uint32_t pack, ver; uint16_t field;
    bool operator==(const BlockKey& o) const { return pack == o.pack && ver == o.ver && field == o.field; }
};
struct BlockKeyHash {
    size_t operator()(const BlockKey& k) const {
        uint64_t x = (uint64_t(k.pack) << 32 | k.ver) * 0x9E3779B97F4A7C15ull;
        x ^= (uint64_t(k.field) + 0x632BE59BD9B4E019ull) * 0xD6E8FEB86659FD93ull;
        return size_t(x ^ (x >> 29));
    }
};

}  // namespace

namespace { struct LaunchArgs; }

struct kx_prog {
    kx_ctx* ctx = nullptr;
    uint64_t id = 0;                     // unique per compiled program (the plan cache must not trust a recycled pointer)
    std::vector<LeafSpec> leaves;
    std::vector<uint8_t> postfix;
    std::vector<uint64_t> sets;          // concatenated sorted sets
    uint32_t set_off[MAX_LEAVES + 1] = {0};
    uint64_t* dev_sets = nullptr;        // sets, then the hash tables (one allocation)
    std::vector<uint64_t> tabs;          // concatenated bucketised hash tables of the IN/NIN leaves
    uint32_t tab_off[MAX_LEAVES] = {0};
    uint8_t tab_log2[MAX_LEAVES] = {0};
    const uint64_t* dev_tabs = nullptr;
    std::vector<uint32_t> pres;          // concatenated prefilter bitmaps of the same leaves
    uint32_t pre_off[MAX_LEAVES] = {0};
    uint8_t pre_log2[MAX_LEAVES] = {0};
    const uint32_t* dev_pres = nullptr;
    std::vector<uint8_t> strs;           // operands of the byte-string leaves
    uint8_t* dev_strs = nullptr;
    bool prune_only = false;             // has a byte-string leaf: usable with kx_prune* only
    ~kx_prog() {   // (also runs when kx_prog_compile fails half way: no device buffer outlives the host struct)
        if (dev_sets) cudaFree(dev_sets);
        if (dev_strs) cudaFree(dev_strs);
    }
};

struct kx_ctx {
    int device = 0;
    int num_sms = 148;
    size_t budget = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_k0 = nullptr, ev_k1 = nullptr, ev_end = nullptr;
    std::mutex mu;
    std::string err;

    std::vector<Slab> slabs;
    std::unordered_map<BlockKey, StoredBlock, BlockKeyHash> store;   // one lookup per (pack, field) and query: O(1)
    size_t store_enc_bytes = 0, store_dev_bytes = 0, slab_bytes = 0;   // encoded bytes registered, live extent bytes, slab capacity

    // scratch (grow only)
    DevBuf d_packs, d_leaves, d_views, d_tilepack, d_counts, d_bitsets, d_partials, d_aggout, d_tmp, d_tmp2, d_misc, d_stage, d_stage2, d_codebits, d_leafbits;
    PinBuf h_desc, h_res, h_aux, h_aux2;
    cudaStream_t copy_stream = nullptr;            // kx_scan_host: uploads of batch b + 1 run beside the scan of batch b
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};

    double last_kernel_ms = 0, last_total_ms = 0;
    int last_launches = 0;
    // QueryStats counters of the last scan (internal/query/stats.go:15-60)
    uint64_t last_rows_scanned = 0, last_packs_scanned = 0, last_rows_matched = 0;

    // plan cache (see plan_cache_run): remembered launches of repeated queries; any kx_block_put / kx_block_drop bumps the epoch
    struct PlanEntry {
        bool valid = false; uint64_t epoch = 0; const kx_prog* prog = nullptr; uint64_t prog_id = 0, env_hash = 0;
        std::vector<kx_packref> packs; std::vector<kx_agg_req> aggs; bool dev_bits = false; std::vector<size_t> bitset_off;
        DevBuf desc; std::shared_ptr<LaunchArgs> A;
    };
    static constexpr int NPLANS = 4;
    PlanEntry plans[NPLANS];
    uint32_t plan_next = 0;
    uint64_t store_epoch = 1;

    // pack-sharded scans: communicator (run-time bound NCCL) and the exchange scratch [own | nranks gathered | combined]
    CommState* comm = nullptr;
    DevBuf d_xchg;
};

// device-resident statistics index (zone maps + bloom filters), see include/knoxgpu.h
struct kx_stats {
    kx_ctx* ctx = nullptr;
    int npacks = 0, nfields = 0;
    std::vector<uint16_t> fields;
    std::vector<uint8_t> types;
    uint64_t* d_mins = nullptr;            // [nfields][npacks]
    uint64_t* d_maxs = nullptr;
    std::vector<uint64_t> h_bloom_ptr;     // [nfields][npacks] device address of the bit array, 0 = none
    std::vector<uint32_t> h_bloom_mask;
    std::vector<uint8_t> h_bloom_k;
    uint8_t* d_tab = nullptr;              // ptr | mask | k tables on the device
    bool dirty = true;
    std::vector<SlabAlloc> cell_alloc;   // [nfields][npacks] slab extent of each bit array (slab = -1: none)
    ~kx_stats() {   // (also runs when kx_stats_create fails half way)
        if (d_mins) cudaFree(d_mins);
        if (d_maxs) cudaFree(d_maxs);
        if (d_tab) cudaFree(d_tab);
    }
};

namespace {

int fail(kx_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            return fail(ctx, KX_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

// Every C entry point runs inside this guard: host allocation failures and other C++ exceptions become error codes
// instead of terminating the caller's process (a corrupt block must never take a database down).
template <typename R, typename F>
R kx_guarded(kx_ctx* ctx, F&& f) {
    try { return f(); }
    catch (const std::bad_alloc&) { return R(fail(ctx, KX_ENOMEM, "out of host memory")); }
    catch (const std::exception& e) { return R(fail(ctx, KX_EINVAL, std::string("internal error: ") + e.what())); }
}

// ------------------------------------------------------------------ slab allocation
int slab_alloc(kx_ctx* ctx, size_t bytes, uint8_t** out, SlabAlloc* rec) {
    bytes = round_up(std::max<size_t>(bytes, 1), 256);
    for (size_t i = 0; i < ctx->slabs.size(); ++i) {
        Slab& s = ctx->slabs[i];
        if (!s.base) continue;
        for (auto it = s.free.begin(); it != s.free.end(); ++it) {
            if (it->second < bytes) continue;
            const size_t off = it->first, rest = it->second - bytes;
            s.free.erase(it);
            if (rest) s.free.emplace(off + bytes, rest);
            s.live += bytes; ctx->store_dev_bytes += bytes;
            *out = s.base + off; *rec = SlabAlloc{int(i), off, bytes};
            return KX_OK;
        }
    }
    const size_t cap = std::max(bytes, SLAB_BYTES);
    if (ctx->budget && ctx->slab_bytes + cap > ctx->budget) {
        // a last, smaller slab may still fit the budget
        if (ctx->slab_bytes + bytes > ctx->budget) return fail(ctx, KX_ENOMEM, "HBM budget exceeded");
    }
    const size_t take = (ctx->budget && ctx->slab_bytes + cap > ctx->budget) ? bytes : cap;
    uint8_t* p = nullptr;
    cudaError_t e = cudaMalloc(&p, take);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(ctx, KX_ENOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e)); }
    size_t idx = ctx->slabs.size();
    for (size_t i = 0; i < ctx->slabs.size(); ++i) if (!ctx->slabs[i].base) { idx = i; break; }
    if (idx == ctx->slabs.size()) ctx->slabs.emplace_back();
    Slab& s = ctx->slabs[idx];
    s.base = p; s.cap = take; s.live = bytes; s.free.clear();
    if (take > bytes) s.free.emplace(bytes, take - bytes);
    ctx->slab_bytes += take; ctx->store_dev_bytes += bytes;
    *out = p; *rec = SlabAlloc{int(idx), 0, bytes};
    return KX_OK;
}

void slab_release(kx_ctx* ctx, const SlabAlloc& a) {
    if (a.slab < 0) return;
    Slab& s = ctx->slabs[size_t(a.slab)];
    s.live -= a.bytes; ctx->store_dev_bytes -= a.bytes;
    if (s.live == 0 && s.base) {
        cudaFree(s.base);
        ctx->slab_bytes -= s.cap;
        s = Slab{};
        return;
    }
    size_t off = a.off, bytes = a.bytes;
    auto nx = s.free.lower_bound(off);
    if (nx != s.free.begin()) {   // merge with the extent that ends where this one starts
        auto pv = std::prev(nx);
        if (pv->first + pv->second == off) { off = pv->first; bytes += pv->second; s.free.erase(pv); }
    }
    if (nx != s.free.end() && off + bytes == nx->first) { bytes += nx->second; s.free.erase(nx); }
    s.free.emplace(off, bytes);
}

// upload one normalised block; `into` receives device pointers.  Synchronous w.r.t. the host
// buffers only if they are pageable (cudaMemcpyAsync semantics); pinned sources stay async.
int upload_block(kx_ctx* ctx, const BlockLayout& lay, StoredBlock& sb) {
    sb.view = lay.view;
    auto put = [&](const void* src, size_t len, const uint8_t** devp) -> int {
        uint8_t* d = nullptr; SlabAlloc rec;
        int rc = slab_alloc(ctx, len + STREAM_PAD, &d, &rec);
        if (rc) return rc;
        sb.allocs.push_back(rec);
        if (len) CK(cudaMemcpyAsync(d, src, len, cudaMemcpyHostToDevice, ctx->stream));
        // the pad only has to be addressable: rows past the end are masked, over-read bits are cut by the field mask
        *devp = d;
        return KX_OK;
    };
    int rc;
    const ColView& v = lay.view;
    if (lay.s8b) {
        // Simple8b: codewords → device scratch, selector counts + widest value, exclusive scan, then the fixed-width stream
        const uint32_t nwords = uint32_t(lay.stream_len / 8);
        const size_t off_counts = round_up(lay.stream_len, 256), off_meta = off_counts + round_up(size_t(nwords) * 4 + 4, 256);
        CK(ctx->d_tmp2.reserve(off_meta + 64));
        uint8_t* scratch = static_cast<uint8_t*>(ctx->d_tmp2.p);
        if (nwords) CK(cudaMemcpyAsync(scratch, lay.stream, lay.stream_len, cudaMemcpyHostToDevice, ctx->stream));
        uint32_t* counts = reinterpret_cast<uint32_t*>(scratch + off_counts);
        unsigned long long* total = reinterpret_cast<unsigned long long*>(scratch + off_meta);
        uint32_t* maxbits = reinterpret_cast<uint32_t*>(scratch + off_meta + 8);
        CK(launch_s8b_count(scratch, nwords, counts, maxbits, total, ctx->stream));
        struct { unsigned long long total; uint32_t maxbits, pad; } meta{};
        CK(cudaMemcpyAsync(&meta, total, sizeof(meta), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (meta.total < v.n) return fail(ctx, KX_EFORMAT, "kx_block_put: simple8b: short stream");
        sb.view.width = uint8_t(meta.maxbits);
        sb.view.kind = meta.maxbits ? CK_BITS : CK_CONST;
        if (meta.maxbits) {
            const size_t bytes = round_up((size_t(v.n) * meta.maxbits + 7) / 8, 8);
            uint8_t* d = nullptr; SlabAlloc rec;
            if ((rc = slab_alloc(ctx, bytes + STREAM_PAD, &d, &rec))) return rc;
            sb.allocs.push_back(rec);
            CK(cudaMemsetAsync(d, 0, bytes + STREAM_PAD, ctx->stream));
            CK(launch_s8b_pack(scratch, nwords, counts, v.n, meta.maxbits, d, ctx->stream));
            sb.view.data = d;
        }
        return KX_OK;
    }
    if (v.kind == CK_BITS || v.kind == CK_DICT || ((v.kind == CK_ALP || v.kind == CK_ALPRD) && v.width)) {
        const void* src = lay.owned.empty() ? (const void*)lay.stream : (const void*)lay.owned.data();
        size_t len = lay.owned.empty() ? lay.stream_len : lay.owned.size();
        if ((rc = put(src, len, &sb.view.data))) return rc;
    }
    if ((v.kind == CK_ALP || v.kind == CK_ALPRD) && !lay.blob.empty()) {   // ALP: patch blob; ALP-RD: the left bit stream
        if ((rc = put(lay.blob.data(), lay.blob.size(), &sb.view.aux))) return rc;
    }
    if (v.kind == CK_DICT) {
        if ((rc = put(lay.aux64.data(), lay.aux64.size() * 8, &sb.view.aux))) return rc;
        sb.dict = lay.aux64;
    }
    if (v.kind == CK_RUNEND) {
        if ((rc = put(lay.aux64.data(), lay.aux64.size() * 8, &sb.view.data))) return rc;
        if ((rc = put(lay.aux32.data(), lay.aux32.size() * 4, &sb.view.aux))) return rc;
    }
    if (!lay.owned.empty() || v.kind == CK_DICT || v.kind == CK_RUNEND || v.kind == CK_ALP || v.kind == CK_ALPRD) {
        // host-owned temporaries (lay.owned / aux vectors) die with `lay`: finish the copies now
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return KX_OK;
}

// byte-string block: byte buffer verbatim + flat u32 index array (kx_types.h STR_*)
int upload_string_block(kx_ctx* ctx, const StrLayout& lay, StoredBlock& sb) {
    sb.view = lay.view;
    auto put = [&](const void* src, size_t len, const uint8_t** devp) -> int {
        uint8_t* d = nullptr; SlabAlloc rec;
        int rc = slab_alloc(ctx, len + STREAM_PAD, &d, &rec);
        if (rc) return rc;
        sb.allocs.push_back(rec);
        if (len) CK(cudaMemcpyAsync(d, src, len, cudaMemcpyHostToDevice, ctx->stream));
        *devp = d;
        return KX_OK;
    };
    int rc;
    if ((rc = put(lay.bytes, lay.nbytes, &sb.view.data))) return rc;
    if (!lay.idx.empty() && (rc = put(lay.idx.data(), lay.idx.size() * 4, &sb.view.aux))) return rc;
    if (lay.view.is_raw == STR_CONST) sb.cstr.assign(lay.bytes, lay.bytes + lay.nbytes);
    CK(cudaStreamSynchronize(ctx->stream));   // lay.idx dies with `lay`
    return KX_OK;
}

void free_block(kx_ctx* ctx, StoredBlock& sb) {
    for (auto& a : sb.allocs) slab_release(ctx, a);
    sb.allocs.clear();
}

// ------------------------------------------------------------------ the scan driver
struct ScanJob {
    int npacks = 0;
    std::vector<uint32_t> nrows;               // [npacks]
    std::vector<ColView> leaf_views;           // [npacks][nleaves]
    std::vector<const uint64_t*> leaf_dicts;   // [npacks][nleaves] host dict copies (or null)
    std::vector<const std::vector<uint8_t>*> leaf_cstr;   // [npacks][nleaves] value of a constant string block (or null)
    std::vector<ColView> agg_views;            // [npacks][naggs]
};

}  // namespace
static int build_scan_job(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, const kx_agg_req* aggs, int naggs, ScanJob& job);
namespace {

// optional selection-vector output of a scan (kx_scan_select)
struct SelectOut {
    uint32_t* sel = nullptr; size_t cap = 0; uint64_t* off = nullptr;   // caller buffers: ids, capacity in ids, npacks + 1 offsets
    bool overflow = false;
};

// optional cross-rank combine of a scan's totals (kx_scan_sharded)
struct ShardOut { int64_t* total_count = nullptr; };

// enqueue on the scan stream: pack this rank's record (sum of the per-pack counts + combined aggregates), ONE all-gather,
// combine in rank order.  Afterwards the combined record sits at xchg_final(ctx).
const RankPartial* xchg_final(kx_ctx* ctx) { return static_cast<const RankPartial*>(ctx->d_xchg.p) + 1 + comm_nranks(ctx->comm); }
int enqueue_exchange(kx_ctx* ctx, uint32_t npacks, uint32_t naggs, const uint8_t* agg_type) {
    const int nranks = comm_nranks(ctx->comm);
    CK(ctx->d_xchg.reserve(sizeof(RankPartial) * size_t(nranks + 2)));
    RankPartial* xs = static_cast<RankPartial*>(ctx->d_xchg.p);
    CK(launch_xchg_pack(static_cast<const unsigned long long*>(ctx->d_counts.p), npacks, static_cast<const AggPartial*>(ctx->d_aggout.p), naggs, xs, ctx->stream));
    std::string err;
    if (comm_allgather(ctx->comm, xs, xs + 1, sizeof(RankPartial), ctx->stream, err)) return fail(ctx, KX_ECUDA, "kx_scan_sharded: " + err);
    CK(launch_xchg_combine(xs + 1, uint32_t(nranks), naggs, agg_type, xs + 1 + nranks, ctx->stream));
    ctx->last_launches += 2 + (nranks > 1 ? 1 : 0);
    return KX_OK;
}

// float32 zone maps / operands travel as float64 patterns on the device (exact widening; same ordering, NaN stays NaN)
uint64_t f32_to_f64_bits(uint64_t pat) {
    uint32_t u = uint32_t(pat); float f; std::memcpy(&f, &u, 4);
    double d = double(f); uint64_t o; std::memcpy(&o, &d, 8);
    return o;
}

// device partial → C ABI result (SumReducer wraps in T; float64 sums carry their compensation term)
kx_agg_out agg_result(const AggPartial& a, int t) {
    kx_agg_out o{};
    o.count = int64_t(a.count); o.valid = a.valid ? 1 : 0;
    if (a.valid) {
        if (t == KX_FLOAT64) {
            double hi, s; std::memcpy(&hi, &a.sum, 8);
            s = hi + a.err;
            std::memcpy(&o.sum_bits, &s, 8);
            o.sum_err = (hi - s) + a.err;
            o.min_bits = a.mn; o.max_bits = a.mx;
        } else if (t == KX_FLOAT32) {
            // float32 columns are reduced in float64 on the device (every float32 is exact in float64) and rounded ONCE here:
            // closer to the exact sum than SumReducer[float32]'s running float32 sum (reducer.go:173-178), equal for min / max
            double hi, mn, mx; std::memcpy(&hi, &a.sum, 8); std::memcpy(&mn, &a.mn, 8); std::memcpy(&mx, &a.mx, 8);
            const double s = hi + a.err;
            const float fs = float(s), fmn = float(mn), fmx = float(mx);
            uint32_t u; std::memcpy(&u, &fs, 4); o.sum_bits = u;
            std::memcpy(&u, &fmn, 4); o.min_bits = u;
            std::memcpy(&u, &fmx, 4); o.max_bits = u;
            o.sum_err = s - double(fs);
        } else {
            uint64_t flip = type_is_signed(t) ? 0x8000000000000000ull : 0;
            o.sum_bits = type_ext(t, a.sum);     // SumReducer wraps in T
            o.min_bits = a.mn ^ flip; o.max_bits = a.mx ^ flip;
        }
    }
    return o;
}

// shared-memory layout of the warp-autonomous kernel for one program (ScanParams::w_*)
struct WarpGeo {
    uint32_t wd, stages, warps, ncols, stage_bytes, stage_off, warp_bytes, mw_slots, chunk_rows;
    uint8_t slot[32];
    uint16_t slot_off[32], col_off[MAX_SCAN_LEAVES], fix_off[MAX_SCAN_LEAVES];
};

// everything the launch + result phase of a scan needs (filled by run_scan, or taken from the plan cache)
struct LaunchArgs {
    ScanParams P{};
    int grid = 1; size_t smem_bytes = 0; bool simple = true, only32 = true; int ctas = 2;
    int npacks = 0, naggs = 0; uint32_t ntiles = 0; uint64_t total_rows = 0;
    const uint8_t* hd = nullptr; uint8_t* dd = nullptr; size_t desc_bytes = 0;   // descriptor block: host staging → device (hd == nullptr: already resident)
    size_t leafbits_bytes = 0; uint32_t code_words = 0;
    const std::vector<std::pair<size_t, int>>* mask_jobs = nullptr; const uint8_t* const* row_masks = nullptr; const uint32_t* nrows = nullptr;
    const AlpFixJob* ajobs = nullptr; uint32_t najobs = 0, max_patches = 0;
    const StrJob* sjobs = nullptr; uint32_t nsjobs = 0, max_str_rows = 0;
    const ValJob* vjobs = nullptr; uint32_t nvjobs = 0, max_val_rows = 0;
    const RunFillJob* rjobs = nullptr; uint32_t nrjobs = 0, max_runs = 0;
    const CodesetJob* cjobs = nullptr; uint32_t ncjobs = 0, max_code_set = 0;
    const kx_prog* prog = nullptr;
    uint8_t* bitsets = nullptr; size_t bitset_total = 0; bool dev_bits = false;
    int64_t* counts = nullptr; const kx_agg_req* aggs = nullptr; kx_agg_out* agg_out = nullptr;
    SelectOut* so = nullptr; uint64_t sel_words = 0; ShardOut* sh = nullptr;
};

int launch_and_collect(kx_ctx* ctx, LaunchArgs& A) {
    const ScanParams& P = A.P;
    const int npacks = A.npacks, naggs = A.naggs;
    const uint32_t ntiles = A.ntiles;
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    if (A.hd) CK(cudaMemcpyAsync(A.dd, A.hd, A.desc_bytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_counts.p, 0, sizeof(unsigned long long) * size_t(npacks), ctx->stream));
    if (naggs) CK(cudaMemsetAsync(ctx->d_aggout.p, 0, sizeof(AggPartial) * MAX_AGGS + 16, ctx->stream));   // "no match" results, counter = 0
    CK(cudaEventRecord(ctx->ev_k0, ctx->stream));
    if (ntiles && A.leafbits_bytes) CK(cudaMemsetAsync(ctx->d_leafbits.p, 0, A.leafbits_bytes, ctx->stream));
    if (ntiles && A.mask_jobs) for (auto& mj : *A.mask_jobs)   // the caller's masks (host memory) land behind the zero fill, before any kernel reads them
        CK(cudaMemcpyAsync(static_cast<uint8_t*>(ctx->d_leafbits.p) + mj.first, A.row_masks[mj.second], (size_t(A.nrows[size_t(mj.second)]) + 7) / 8,
                           cudaMemcpyHostToDevice, ctx->stream));
    if (ntiles && A.najobs) {
        CK(launch_alpfix(A.ajobs, A.najobs, A.max_patches, static_cast<uint8_t*>(ctx->d_leafbits.p), ctx->stream));
        ctx->last_launches++;
    }
    if (ntiles && A.nsjobs) {
        CK(launch_strmatch(A.sjobs, A.nsjobs, A.max_str_rows, A.prog->dev_strs, static_cast<uint8_t*>(ctx->d_leafbits.p), ctx->stream));
        ctx->last_launches++;
    }
    if (ntiles && A.nvjobs) {
        CK(launch_valmatch(A.vjobs, A.nvjobs, A.max_val_rows, static_cast<uint8_t*>(ctx->d_leafbits.p), ctx->stream));
        ctx->last_launches++;
    }
    if (ntiles && A.nrjobs) {
        CK(launch_runfill(A.rjobs, A.nrjobs, A.max_runs, A.prog->dev_sets, static_cast<uint8_t*>(ctx->d_leafbits.p), ctx->stream));
        ctx->last_launches++;
    }
    if (ntiles && A.ncjobs) {
        CK(cudaMemsetAsync(ctx->d_codebits.p, 0, size_t(A.code_words) * 4, ctx->stream));
        CK(launch_codeset(A.cjobs, A.ncjobs, A.max_code_set, A.prog->dev_sets, static_cast<uint32_t*>(ctx->d_codebits.p), ctx->stream));
        ctx->last_launches++;
    }
    if (ntiles) {
        if (A.simple) CK(launch_scan(P, A.grid, A.smem_bytes, A.only32, A.ctas, ctx->stream));
        else CK(launch_scan_warp(P, A.grid, A.smem_bytes, ctx->stream));
        ctx->last_launches++;
    }
    if (A.sh) { int rc = enqueue_exchange(ctx, uint32_t(npacks), uint32_t(naggs), P.agg_type); if (rc) return rc; }
    CK(cudaEventRecord(ctx->ev_k1, ctx->stream));

    // ---- results back to the host
    SelectOut* so = A.so;
    uint8_t* hr = static_cast<uint8_t*>(ctx->h_res.p);
    size_t res_counts = sizeof(unsigned long long) * size_t(npacks);
    if (so && ntiles) {
        // Bitset.Indexes for every pack (reader.go:432-436): ids are written only if they fit the caller's buffer
        // (the per-pack counts decide that after the copy below)
        unsigned long long total_matches = 0;
        CK(cudaMemcpyAsync(hr, ctx->d_counts.p, res_counts, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int p = 0; p < npacks; ++p) total_matches += reinterpret_cast<unsigned long long*>(hr)[p];
        if (total_matches > so->cap) so->overflow = true;
        else if (total_matches) {
            CK(launch_select(P.packs, uint32_t(npacks), static_cast<const uint8_t*>(ctx->d_bitsets.p), A.sel_words,
                             reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->d_misc.p) + 64), static_cast<unsigned long long*>(ctx->d_misc.p),
                             static_cast<uint32_t*>(ctx->d_tmp2.p), ctx->stream));
            ctx->last_launches += 3;
            CK(cudaEventRecord(ctx->ev_k1, ctx->stream));
            CK(cudaMemcpyAsync(so->sel, ctx->d_tmp2.p, size_t(total_matches) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    CK(cudaMemcpyAsync(hr, ctx->d_counts.p, res_counts, cudaMemcpyDeviceToHost, ctx->stream));   // also feeds the QueryStats counters
    RankPartial* h_rp = reinterpret_cast<RankPartial*>(hr + round_up(res_counts, 64));
    if (A.sh) CK(cudaMemcpyAsync(h_rp, xchg_final(ctx), sizeof(RankPartial), cudaMemcpyDeviceToHost, ctx->stream));   // combined over all ranks
    else if (naggs) CK(cudaMemcpyAsync(h_rp->agg, ctx->d_aggout.p, sizeof(AggPartial) * size_t(naggs), cudaMemcpyDeviceToHost, ctx->stream));
    if (A.bitsets && A.bitset_total) CK(cudaMemcpyAsync(A.bitsets, ctx->d_bitsets.p, A.bitset_total, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1)); ctx->last_kernel_ms = ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end)); ctx->last_total_ms = ms;

    int64_t* counts = A.counts;
    std::vector<int64_t> tmp_counts;
    if (so && !counts) { tmp_counts.resize(size_t(npacks)); counts = tmp_counts.data(); }
    if (counts) for (int p = 0; p < npacks; ++p) counts[p] = int64_t(reinterpret_cast<unsigned long long*>(hr)[p]);
    if (so) for (int p = 0; p < npacks; ++p) so->off[p + 1] = so->off[p] + uint64_t(counts[p]);
    for (int j = 0; j < naggs; ++j) A.agg_out[j] = agg_result(h_rp->agg[j], A.aggs[j].block_type);
    if (A.sh && A.sh->total_count) *A.sh->total_count = int64_t(h_rp->total_count);
    ctx->last_rows_scanned = A.total_rows; ctx->last_packs_scanned = uint64_t(npacks);
    for (int p = 0; p < npacks; ++p) ctx->last_rows_matched += reinterpret_cast<unsigned long long*>(hr)[p];
    return KX_OK;
}

// ------------------------------------------------------------------ plan cache
// A query that is repeated over an unchanged store (dashboards, the benchmark's steps) re-uses its translated leaves and
// descriptors: the entry keeps its own device copy of the descriptor block and the launch parameters.
struct PlanKey {
    const kx_prog* prog = nullptr;
    const kx_packref* packs = nullptr; int npacks = 0;
    const kx_agg_req* aggs = nullptr; int naggs = 0;
    bool dev_bits = false; const size_t* bitset_off = nullptr;
    uint64_t env_hash = 0;
};
uint64_t env_knobs_hash() {   // the tuning hooks change the plan: they are part of the key
    uint64_t h = 1469598103934665603ull;
    for (const char* k : {"KX_SCAN_GEOMETRY", "KX_SCHED_CHUNK", "KX_AGG_STAGE", "KX_HASH_SMEM_KB", "KX_WARP_GEOMETRY"}) {
        const char* v = getenv(k);
        for (const char* c = v ? v : ""; *c; ++c) h = (h ^ uint8_t(*c)) * 1099511628211ull;
        h = (h ^ 0xff) * 1099511628211ull;
    }
    return h;
}
bool plan_matches(const kx_ctx::PlanEntry& e, const kx_ctx* ctx, const PlanKey& k) {
    if (!e.valid || e.epoch != ctx->store_epoch || e.prog != k.prog || e.prog_id != k.prog->id || e.env_hash != k.env_hash) return false;
    if (int(e.packs.size()) != k.npacks || int(e.aggs.size()) != k.naggs || e.dev_bits != k.dev_bits) return false;
    if (k.npacks && std::memcmp(e.packs.data(), k.packs, sizeof(kx_packref) * size_t(k.npacks))) return false;
    if (k.naggs && std::memcmp(e.aggs.data(), k.aggs, sizeof(kx_agg_req) * size_t(k.naggs))) return false;
    if (k.dev_bits && std::memcmp(e.bitset_off.data(), k.bitset_off, sizeof(size_t) * size_t(k.npacks))) return false;
    return true;
}
void plan_cache_store(kx_ctx* ctx, const PlanKey& k, const LaunchArgs& A, size_t desc_bytes, size_t off_leaves, size_t off_views, size_t off_tiles) {
    kx_ctx::PlanEntry& e = ctx->plans[ctx->plan_next++ % kx_ctx::NPLANS];
    e.valid = false;
    if (e.desc.reserve(desc_bytes) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMemcpyAsync(e.desc.p, A.dd, desc_bytes, cudaMemcpyDeviceToDevice, ctx->stream) != cudaSuccess) { cudaGetLastError(); return; }
    cudaStreamSynchronize(ctx->stream);
    e.epoch = ctx->store_epoch; e.prog = k.prog; e.prog_id = k.prog->id; e.env_hash = k.env_hash;
    e.packs.assign(k.packs, k.packs + k.npacks);
    e.aggs.assign(k.aggs, k.aggs + k.naggs);
    e.dev_bits = k.dev_bits;
    if (k.dev_bits) e.bitset_off.assign(k.bitset_off, k.bitset_off + k.npacks); else e.bitset_off.clear();
    e.A = std::make_shared<LaunchArgs>(A);
    LaunchArgs& C = *e.A;
    uint8_t* d = static_cast<uint8_t*>(e.desc.p);
    C.P.packs = reinterpret_cast<const PackInfo*>(d);
    C.P.leaves = reinterpret_cast<const PackLeaf*>(d + off_leaves);
    C.P.views = reinterpret_cast<const ColView*>(d + off_views);
    C.P.tile_pack = off_tiles ? reinterpret_cast<const uint32_t*>(d + off_tiles) : nullptr;
    C.hd = nullptr; C.dd = d; C.mask_jobs = nullptr; C.row_masks = nullptr; C.nrows = nullptr;
    C.bitsets = nullptr; C.counts = nullptr; C.aggs = nullptr; C.agg_out = nullptr; C.so = nullptr; C.sh = nullptr;
    e.valid = true;
}
// the fast path of kx_scan / kx_scan_sharded: KX_OK + *hit = true when a remembered plan served the call
int plan_cache_run(kx_ctx* ctx, const PlanKey& k, uint8_t* bitsets, int64_t* counts, kx_agg_out* agg_out, ShardOut* sh, bool* hit) {
    *hit = false;
    for (auto& e : ctx->plans) {
        if (!plan_matches(e, ctx, k)) continue;
        LaunchArgs A = *e.A;
        const int npacks = A.npacks, naggs = A.naggs;
        ctx->last_kernel_ms = ctx->last_total_ms = 0; ctx->last_launches = 0;
        ctx->last_rows_scanned = ctx->last_packs_scanned = ctx->last_rows_matched = 0;
        // scratch buffers may have been re-allocated by other calls since: reserve and refresh the pointers
        CK(ctx->d_counts.reserve(sizeof(unsigned long long) * size_t(npacks)));
        if (A.dev_bits) CK(ctx->d_bitsets.reserve(round_up(A.bitset_total, 8) + 64));
        if (naggs) {
            CK(ctx->d_partials.reserve(sizeof(AggPartial) * size_t(A.grid) * naggs));
            CK(ctx->d_aggout.reserve(sizeof(AggPartial) * MAX_AGGS + 16));
        }
        CK(ctx->h_res.reserve(sizeof(unsigned long long) * size_t(npacks) + sizeof(RankPartial) + 128));
        A.P.bitsets = A.dev_bits ? static_cast<uint8_t*>(ctx->d_bitsets.p) : nullptr;
        A.P.counts = static_cast<unsigned long long*>(ctx->d_counts.p);
        A.P.partials = static_cast<AggPartial*>(ctx->d_partials.p);
        A.P.agg_out = static_cast<AggPartial*>(ctx->d_aggout.p);
        A.P.done = naggs ? reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->d_aggout.p) + sizeof(AggPartial) * MAX_AGGS) : nullptr;
        A.prog = k.prog; A.bitsets = bitsets; A.counts = counts; A.aggs = k.aggs; A.agg_out = agg_out; A.sh = sh;
        *hit = true;
        return launch_and_collect(ctx, A);
    }
    return KX_OK;
}

// keep_layout (optional): leave the match bitsets on the device (ctx->d_bitsets, pack i at keep_layout[i]; PackInfo table
// at the start of ctx->d_packs) for a follow-up kernel on the same stream (kx_scan_buckets)
int run_scan(kx_ctx* ctx, const kx_prog* prog, const ScanJob& job, uint8_t* bitsets, const size_t* bitset_off,
             int64_t* counts, const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out, SelectOut* so = nullptr,
             std::vector<size_t>* keep_layout = nullptr, ShardOut* sh = nullptr, const uint8_t* const* row_masks = nullptr,
             const PlanKey* cache_key = nullptr) {
    const int npacks = job.npacks, nleaves = int(prog->leaves.size());
    // Row masks (TableReader.WithMask, engine/interface.go:96-106; the tombstone / visibility step of reader.go:347-413):
    // one more leaf, ANDed last — a 1-bit column per pack that the scan streams like a run-end pre-pass result
    const bool masked = row_masks != nullptr;
    const int nl = nleaves + (masked ? 1 : 0);
    std::vector<uint8_t> post(prog->postfix);
    if (masked) { post.push_back(uint8_t(nleaves)); post.push_back(KX_OP_AND); }
    // selection vectors are extracted from device-resident bitsets laid out back to back (8-byte aligned)
    std::vector<size_t> sel_layout;
    const bool dev_bits = bitsets != nullptr || so != nullptr || keep_layout != nullptr;
    if (so || keep_layout) {
        if (bitsets) return fail(ctx, KX_EINVAL, "selection vectors and host bitsets cannot be requested together");
        size_t o = 0;
        sel_layout.resize(size_t(npacks));
        for (int p = 0; p < npacks; ++p) { sel_layout[size_t(p)] = o; o += round_up((size_t(job.nrows[size_t(p)]) + 7) / 8, 8); }
        bitset_off = sel_layout.data();
        if (so) for (int p = 0; p <= npacks; ++p) so->off[p] = 0;
        if (keep_layout) *keep_layout = sel_layout;
    }
    ctx->last_kernel_ms = ctx->last_total_ms = 0; ctx->last_launches = 0;
    if (naggs < 0 || naggs > MAX_AGGS) return fail(ctx, KX_EINVAL, "too many aggregates");
    for (int j = 0; j < naggs; ++j) {
        int t = aggs[j].block_type;
        if (type_bits(t) == 0) return fail(ctx, KX_EUNSUPPORTED, "aggregate over unsupported block type");
    }
    ctx->last_rows_scanned = ctx->last_packs_scanned = ctx->last_rows_matched = 0;
    if (npacks == 0) {
        for (int j = 0; j < naggs; ++j) agg_out[j] = kx_agg_out{};
        if (!sh) return KX_OK;
        // a rank whose shard is empty still takes part in the query's collective
        uint8_t types[MAX_AGGS] = {};
        for (int j = 0; j < naggs; ++j) types[j] = aggs[j].block_type;
        CK(ctx->d_counts.reserve(8));
        CK(ctx->d_aggout.reserve(sizeof(AggPartial) * MAX_AGGS + 16));
        CK(ctx->h_res.reserve(sizeof(RankPartial) + 64));
        CK(cudaMemsetAsync(ctx->d_aggout.p, 0, sizeof(AggPartial) * MAX_AGGS + 16, ctx->stream));
        int rc = enqueue_exchange(ctx, 0, uint32_t(naggs), types);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->h_res.p, xchg_final(ctx), sizeof(RankPartial), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const RankPartial* rp = static_cast<const RankPartial*>(ctx->h_res.p);
        if (sh->total_count) *sh->total_count = int64_t(rp->total_count);
        for (int j = 0; j < naggs; ++j) agg_out[j] = agg_result(rp->agg[j], aggs[j].block_type);
        return KX_OK;
    }
    if (bitsets && !bitset_off) return fail(ctx, KX_EINVAL, "bitsets without bitset_off");
    std::vector<int64_t> sel_counts;
    if (so && !counts) { sel_counts.resize(size_t(npacks)); counts = sel_counts.data(); }

    // ---- per-pack / per-leaf descriptors (host), pinned staging
    size_t sz_packs = sizeof(PackInfo) * size_t(npacks), sz_leaves = sizeof(PackLeaf) * size_t(npacks) * size_t(nl);
    size_t nviews = size_t(npacks) * size_t(nleaves + naggs), sz_views = sizeof(ColView) * nviews;
    size_t off_leaves = round_up(sz_packs, 256), off_views = off_leaves + round_up(sz_leaves, 256);
    size_t off_tiles = off_views + round_up(sz_views, 256);

    // translate every leaf for every pack; the widest staged bits/row over all packs decides the tile
    std::vector<PackLeaf> pl(size_t(npacks) * size_t(nl));
    std::vector<std::pair<size_t, int>> mask_jobs;   // (byte offset in the leaf-bit scratch, pack) of every row mask
    std::vector<CodesetJob> cjobs;
    std::vector<RunFillJob> rjobs;
    std::vector<size_t> rjob_leaf;     // index into pl of each run-fill job
    std::vector<AlpFixJob> ajobs;
    std::vector<size_t> ajob_leaf;
    std::vector<StrJob> sjobs;          // byte-string leaves: predicate per row in a pre-pass (strmatch_kernel)
    std::vector<size_t> sjob_leaf;
    std::vector<ValJob> vjobs;          // ALP-RD leaves: decode + float compare per row in a pre-pass (valmatch_kernel)
    std::vector<size_t> vjob_leaf;
    uint32_t max_val_rows = 0;
    uint32_t max_str_rows = 0;
    uint32_t max_patches = 0;
    bool any_fix = false;
    size_t leafbits_bytes = 0;
    uint32_t max_runs = 0;
    uint32_t max_stage_bits = 0, code_words = 0, max_code_set = 0;
    uint32_t code_leaf_words[MAX_SCAN_LEAVES] = {};   // per leaf: largest code bitmap of any pack (cached in shared memory by the scan)
    bool hash_leaf[MAX_SCAN_LEAVES] = {};             // leaf is looked up in its hash set (LM_HASHSET) for some pack
    bool only32 = true;   // every leaf of every pack is a <= 32-bit packed range test (or all / none)
    uint64_t total_rows = 0;
    bool uniform = true;
    for (int p = 0; p < npacks; ++p) {
        uint32_t bits = 0;   // widest staged leaf column of the pack: one ring stage holds ONE column of a tile
        for (int l = 0; l < nleaves; ++l) {
            const ColView& v = job.leaf_views[size_t(p) * nleaves + l];
            if (v.n != job.nrows[size_t(p)]) return fail(ctx, KX_EINVAL, "blocks of one pack differ in length");
            PackLeaf& o = pl[size_t(p) * nl + l];
            if (prog->leaves[size_t(l)].type == KX_BYTES) {
                // byte-string leaf (StringMatcher.Match*, internal/encode/string_*.go): constant blocks are decided here
                // (string_const.go:113-153), every other layout by strmatch_kernel into a 1-bit column the scan streams
                const LeafSpec& ls = prog->leaves[size_t(l)];
                if (v.kind != CK_STR || !ls.has_str) return fail(ctx, KX_EINVAL, "byte-string leaf over a non-string block (or without operand bytes)");
                o = PackLeaf{};
                only32 = false;
                if (v.n == 0 || v.is_raw == STR_CONST) {
                    const std::vector<uint8_t>* cv = job.leaf_cstr.empty() ? nullptr : job.leaf_cstr[size_t(p) * nleaves + l];
                    static const std::vector<uint8_t> empty;
                    if (!cv) cv = &empty;
                    const bool is_set = ls.mode == KX_MODE_IN || ls.mode == KX_MODE_NIN;
                    const bool all = v.n != 0 && (is_set ? string_set_pred(ls.mode, cv->data(), cv->size(), ls.sa.data(), uint32_t(ls.a))
                                                          : string_pred(ls.mode, cv->data(), cv->size(), ls.sa.data(), ls.sa.size(), ls.sb.data(), ls.sb.size()));
                    o.mode = all ? LM_ALL : LM_NONE;
                } else {
                    const bool is_set = ls.mode == KX_MODE_IN || ls.mode == KX_MODE_NIN;   // a_len = number of strings of the set then
                    sjobs.push_back(StrJob{v, leafbits_bytes, ls.sa_off, is_set ? uint32_t(ls.a) : uint32_t(ls.sa.size()), ls.sb_off, uint32_t(ls.sb.size()), uint32_t(ls.mode), 0});
                    sjob_leaf.push_back(size_t(p) * nl + l);
                    leafbits_bytes += round_up((size_t(v.n) + 7) / 8 + STREAM_PAD, 256);
                    max_str_rows = std::max(max_str_rows, v.n);
                    o.mode = LM_BITS; o.width = 1;   // o.data is patched once the scratch buffer is reserved
                    bits = std::max(bits, 1u);
                }
                continue;
            }
            if (v.kind == CK_ALPRD) {
                // ALP-RD block: FloatAlpRdContainer.Match* decode and run the float compare kernels (float_alprd.go:181-211);
                // here a pre-pass decodes every row at index and writes the leaf's 1-bit column
                const LeafSpec& ls = prog->leaves[size_t(l)];
                o = PackLeaf{};
                only32 = false;
                if (v.n == 0 || ls.mode == KX_MODE_IN || ls.mode == KX_MODE_NIN) o.mode = LM_NONE;   // (float IN / NIN never compile)
                else {
                    vjobs.push_back(ValJob{v, ls.a, ls.b, leafbits_bytes, uint32_t(ls.mode), 0});
                    vjob_leaf.push_back(size_t(p) * nl + l);
                    leafbits_bytes += round_up((size_t(v.n) + 7) / 8 + STREAM_PAD, 256);
                    max_val_rows = std::max(max_val_rows, v.n);
                    o.mode = LM_BITS; o.width = 1;   // o.data is patched once the scratch buffer is reserved
                    bits = std::max(bits, 1u);
                }
                continue;
            }
            compile_leaf(v, job.leaf_dicts[size_t(p) * nleaves + l], prog->leaves[size_t(l)], uint32_t(size_t(p) * nleaves + l), o);
            if (o.mode == LM_CODESET) {   // dictionary-set translation runs on the device, one job per (pack, leaf)
                const LeafSpec& ls = prog->leaves[size_t(l)];
                cjobs.push_back(CodesetJob{v.aux, v.naux, ls.set_off, uint32_t(ls.set.size()), code_words, uint32_t(v.delta),
                                           type_is_signed(v.type) ? 0x8000000000000000ull : 0ull});
                max_code_set = std::max(max_code_set, uint32_t(ls.set.size()));
                o.a = code_words;
                // the bitmap spans every code a `width`-bit field can produce (leaf_code32 does not bounds-check)
                const uint32_t nw = uint32_t(((uint64_t(1) << v.width) + v.delta + 31u) / 32u) + 1u;
                code_words += nw;
                code_leaf_words[l] = std::max(code_leaf_words[l], nw);
            }
            if (o.mode == LM_HASHSET) hash_leaf[l] = true;
            if (v.kind == CK_RUNEND && (o.mode == LM_VALRANGE || o.mode == LM_SET || o.mode == LM_RUNRANGE)) {
                // run-end blocks: predicate per run in a pre-pass, the scan streams the resulting 1-bit column
                rjobs.push_back(RunFillJob{v.data, v.aux, o.a, o.d, o.wm, leafbits_bytes, v.naux, v.n, o.mode == LM_SET ? 1u : (o.mode == LM_RUNRANGE ? 2u : 0u), 0});
                rjob_leaf.push_back(size_t(p) * nl + l);
                leafbits_bytes += round_up((size_t(v.n) + 7) / 8 + STREAM_PAD, 256);
                max_runs = std::max(max_runs, v.naux);
                o.mode = LM_BITS; o.width = 1;   // o.data is patched once the scratch buffer is reserved
                bits = std::max(bits, 1u);
            }
            if (o.fixmode) {   // ALP block with patches: the correction is a 1-bit stream staged after the leaf's own
                any_fix = true;
                bits = std::max(bits, 1u);
                if (o.fixmode == FIX_ANDNOT_ALL) o.fix = v.aux + alp_mask_off(v.naux);   // resident patch bitmap
                else {
                    const LeafSpec& ls = prog->leaves[size_t(l)];
                    ajobs.push_back(AlpFixJob{v.aux, ls.a, ls.b, leafbits_bytes, v.naux, uint32_t(ls.mode == KX_MODE_NE ? uint8_t(KX_MODE_EQ) : ls.mode),
                                              o.fixmode == FIX_ANDNOT_NPRED ? 1u : 0u, 0});
                    ajob_leaf.push_back(size_t(p) * nl + l);
                    leafbits_bytes += round_up((size_t(v.n) + 7) / 8 + STREAM_PAD, 256);
                    max_patches = std::max(max_patches, v.naux);
                }
            }
            bits = std::max(bits, uint32_t(leaf_stage_width(o)));
            if (o.mode != LM_RANGE32 && o.mode != LM_NONE && o.mode != LM_ALL) only32 = false;
        }
        if (masked) {
            PackLeaf& o = pl[size_t(p) * nl + nleaves];
            o = PackLeaf{};
            only32 = false;
            if (!row_masks[p] || job.nrows[size_t(p)] == 0) o.mode = LM_ALL;   // no mask for this pack: every row stays eligible
            else {
                mask_jobs.push_back({leafbits_bytes, p});
                leafbits_bytes += round_up((size_t(job.nrows[size_t(p)]) + 7) / 8 + STREAM_PAD, 256);
                o.mode = LM_BITS; o.width = 1;   // o.data is patched once the scratch buffer is reserved
                bits = std::max(bits, 1u);
            }
        }
        for (int j = 0; j < naggs; ++j) {
            const ColView& av = job.agg_views[size_t(p) * naggs + j];
            if (av.n != job.nrows[size_t(p)]) return fail(ctx, KX_EINVAL, "value block length differs from pack");
        }
        max_stage_bits = std::max(max_stage_bits, bits);
        total_rows += job.nrows[size_t(p)];
        if (job.nrows[size_t(p)] != job.nrows[0]) uniform = false;
    }

    // ---- tile geometry.  A tile is 256 R rows (R 32-row groups per consumer warp); the lane-owns-a-group
    // paths want R >= 32 so that every lane has work.  Measured on B200 (profiles/r1_tune_geometry.txt): big
    // tiles in a 2-deep ring beat small tiles in a deep ring (per-tile barrier/dispatch cost, larger TMA
    // copies), narrow ALU-bound columns like 3 CTAs per SM, wide ones 2 (or 1 when 8192 rows need > 50 KB).
    struct Geo { int ctas, stages; size_t budget; };
    const bool simple = nl == 1 && naggs == 0 && !any_fix;   // patch corrections need the general ring protocol
    const bool simple32 = only32 && simple;
    // multi-leaf programs, patch corrections and fused reduces run the warp-autonomous kernel (kx_warp.cu)
    const bool use_warp = !simple;
    // dictionary-code bitmaps cached in shared memory behind the ring (the warp kernel reads them in place)
    uint32_t code_smem_off[MAX_SCAN_LEAVES] = {}, code_smem_words = 0;
    if (!use_warp) for (int l = 0; l < nleaves; ++l) { code_smem_off[l] = code_smem_words; code_smem_words += code_leaf_words[l]; }
    // hash-set leaves: prefilter bitmap (<= 16 KB) and, when it is small (<= 16 KB: sets up to ~500 keys), the exact table
    // in shared memory too; larger tables stay in global memory (L2) and only candidates are verified against them
    uint32_t hs_smem_off[MAX_SCAN_LEAVES] = {}, hs_tab_smem_off[MAX_SCAN_LEAVES] = {};
    uint32_t hs_tab_limit = 16u * 1024u;
    if (const char* e = getenv("KX_HASH_SMEM_KB")) hs_tab_limit = uint32_t(atoi(e)) * 1024u;   // tuning hook
    const uint32_t code_bitmap_words = code_smem_words;
    code_smem_words = uint32_t(round_up(code_smem_words, 8));   // tables are read with 128-bit loads
    for (int l = 0; l < nleaves; ++l) {
        hs_tab_smem_off[l] = 0xffffffffu;
        if (!hash_leaf[l]) continue;
        hs_smem_off[l] = code_smem_words;
        code_smem_words += (1u << prog->pre_log2[l]) / 32u;
        const uint32_t tab_words = (4u << prog->tab_log2[l]) * 2u;
        // (all bitmaps and tables together stay below 96 KB, so that a two-stage ring still fits beside them)
        if (tab_words * 4u <= hs_tab_limit && (size_t(code_smem_words) + tab_words) * 4u <= 96u * 1024u) { hs_tab_smem_off[l] = code_smem_words; code_smem_words += tab_words; }
    }
    const size_t code_smem_bytes = round_up(size_t(code_smem_words) * 4, 128);
    // fused reduce: a tile whose matches * thr exceed its rows reads the value rows of every lane with 128-bit loads (dense
    // walk), others read matching rows on demand (KX_AGG_STAGE = never | always | <thr>; tuning hook)
    uint32_t agg_dense_thr = 3;
    if (naggs) {
        const char* e = getenv("KX_AGG_STAGE");
        if (e && !strcmp(e, "never")) agg_dense_thr = 0xffffffffu;
        else if (e && !strcmp(e, "always")) agg_dense_thr = 0;
        else if (e && atoi(e) > 0) agg_dense_thr = uint32_t(atoi(e));
    }
    // pure AND / pure OR programs combine into one running word per lane; other trees (and ALP patch corrections) use a
    // per-warp AND/OR stack in shared memory
    uint32_t stack_depth = 0, flat_op = 0;
    if (!simple) {
        bool all_and = true, all_or = true;
        for (uint8_t op : post) { if (op == KX_OP_AND) all_or = false; else if (op == KX_OP_OR) all_and = false; }
        if (!any_fix) flat_op = all_and ? 1u : (all_or ? 2u : 0u);
        if (!flat_op) {
            uint32_t sp = 0;
            for (uint8_t op : post) { if (op < 0x80) ++sp; else --sp; stack_depth = std::max(stack_depth, sp); }
        }
    }
    auto desc_words_for = [&](uint32_t slots) { return uint32_t(nl * (sizeof(PackLeaf) / 4) + size_t(slots) * size_t(naggs) * (sizeof(ColView) / 4)); };
    const size_t stage_fixed = 32;
    auto stage_bytes_for = [&](uint32_t r) { return round_up(size_t(32) * r * max_stage_bits + stage_fixed, 128); };
    auto rmax_for = [&](const Geo& g) {
        const size_t fixed = code_smem_bytes / 2 + stage_fixed + 128;
        return g.budget > fixed ? (g.budget - fixed) / (size_t(32) * std::max<uint32_t>(max_stage_bits, 1)) : size_t(0);
    };
    const Geo g3{3, 2, 33 * 1024}, g2{2, 2, 50 * 1024}, g1{1, 2, 99 * 1024};
    Geo geo = g2;
    uint32_t R = 32;
    if (simple) {
        if (max_stage_bits) {
            if (simple32 && max_stage_bits <= 16) geo = g3;
            size_t rmax = rmax_for(geo);
            if (rmax < 32) { geo = g1; rmax = rmax_for(geo); }
            R = rmax >= 32 ? uint32_t(std::min<size_t>(rmax / 32 * 32, 256)) : uint32_t(std::max<size_t>(rmax, 1));
        }
    }
    if (const char* e = getenv("KX_SCAN_GEOMETRY")) {   // tuning hook: "ctas,stages,R"
        int c = 0, st = 0, r = 0;
        if (sscanf(e, "%d,%d,%d", &c, &st, &r) == 3 && c >= 1 && c <= 3 && st >= 2 && st <= MAX_STAGES && r >= 1 &&
            simple && (r <= 32 || r % 32 == 0) &&
            128 + size_t(st) * stage_bytes_for(uint32_t(r)) + code_smem_bytes <= SCAN_MAX_DYN_SMEM / size_t(c)) {
            geo.ctas = c; geo.stages = st; R = uint32_t(r);
        }
    }
    // small scans: prefer more, smaller tiles so that every CTA of the persistent grid gets work
    while (R > 32 && total_rows / (uint64_t(256) * R) < uint64_t(4) * ctx->num_sms * geo.ctas) R -= 32;
    // ---- warp-autonomous kernel: a tile is 1024 wd rows of ONE warp; a ring stage holds one slot per staged column
    // (leaf streams in postfix order, ALP correction streams behind their leaf); every warp of the single CTA per SM owns
    // `stages` stages plus its match words, AND/OR stack and descriptor cache
    WarpGeo wg{};
    if (use_warp) {
        uint32_t leaf_w[MAX_SCAN_LEAVES] = {};
        bool leaf_fixs[MAX_SCAN_LEAVES] = {};
        for (int p = 0; p < npacks; ++p)
            for (size_t l = 0; l < nl; ++l) {
                const PackLeaf& o = pl[size_t(p) * nl + l];
                const uint32_t w = o.data ? o.width : (o.mode == LM_BITS ? 1u : 0u);   // (pre-pass columns get their pointer below)
                leaf_w[l] = std::max(leaf_w[l], w);
                if (o.fixmode) leaf_fixs[l] = true;
            }
        const uint32_t max_warps = naggs > 2 ? 8u : 16u;
        auto layout = [&](uint32_t wd, uint32_t stages, WarpGeo& g) {
            g = WarpGeo{};
            g.wd = wd; g.stages = stages;
            size_t off = 0;
            uint32_t ns = 0;
            for (uint8_t op : post) {
                if (op >= 0x80) continue;
                if (leaf_w[op]) {
                    g.col_off[op] = uint16_t(off / 16); g.slot[ns] = op; g.slot_off[ns] = uint16_t(off / 16); ++ns;
                    off += round_up(size_t(128) * wd * leaf_w[op], 16) + 16;   // (+ 16: the hash-set walk reads two words past a field)
                }
                if (leaf_fixs[op]) {
                    g.fix_off[op] = uint16_t(off / 16); g.slot[ns] = uint8_t(op | 0x80); g.slot_off[ns] = uint16_t(off / 16); ++ns;
                    off += size_t(128) * wd + 16;
                }
            }
            if (ns == 0) { g.slot[0] = 0xff; g.slot_off[0] = 0; ns = 1; }   // nothing staged: the barrier still gets one (empty) arrival per tile
            g.ncols = ns;
            // (a fused reduce streams the raw 64-bit value columns of densely matching tiles through the same stages, in
            // chunks of stage_bytes / 8 rows: keep a stage at 4 KB or more then)
            g.stage_bytes = uint32_t(round_up(std::max<size_t>(off, naggs ? 4096 : 16), 128));
            g.mw_slots = stages + 2;
            g.chunk_rows = (g.stage_bytes / 8) & ~31u;
            const size_t fixed = 256 + size_t(g.mw_slots) * wd * 128 + size_t(stack_depth) * wd * 128 + size_t(desc_words_for(g.mw_slots)) * 4;
            g.stage_off = uint32_t(round_up(fixed, 128));
            g.warp_bytes = g.stage_off + stages * g.stage_bytes;
            const size_t room = WARP_MAX_DYN_SMEM > code_smem_bytes ? WARP_MAX_DYN_SMEM - code_smem_bytes : 0;
            g.warps = uint32_t(std::min<size_t>(max_warps, room / g.warp_bytes));
        };
        // preference (measured, profiles/r2_tune_warp.txt): 2048-row tiles beat 1024-row tiles even when only 10-12 warps
        // fit beside them (per-tile control code is amortised over twice the rows); all 16 warps with two stages beat 12 with three
        const uint32_t cand[][3] = {{2, 2, 16}, {2, 3, 12}, {2, 2, 10}, {1, 3, 16}, {1, 2, 1}};
        bool found = false;
        for (auto& c : cand) {
            layout(c[0], c[1], wg);
            if (wg.warps >= std::min(c[2], max_warps)) { found = true; break; }
        }
        if (!found) return fail(ctx, KX_EUNSUPPORTED, "scan program does not fit the shared memory of one SM");
        if (const char* e = getenv("KX_WARP_GEOMETRY")) {   // tuning hook: "wd,stages,warps"
            int a = 0, b = 0, c = 0;
            if (sscanf(e, "%d,%d,%d", &a, &b, &c) == 3 && (a == 1 || a == 2 || a == 4) && b >= 2 && b <= 4 && c >= 1 && c <= int(max_warps)) {
                WarpGeo t{};
                layout(uint32_t(a), uint32_t(b), t);
                if (t.warps >= 1) { wg = t; wg.warps = std::min<uint32_t>(wg.warps, uint32_t(c)); }
            }
        }
        // small scans: prefer more, smaller tiles so that every warp of the persistent grid gets work
        while (wg.wd > 1 && total_rows / (uint64_t(1024) * wg.wd) < uint64_t(4) * ctx->num_sms * wg.warps) {
            const uint32_t keep_warps = wg.warps;
            layout(wg.wd / 2, wg.stages, wg);
            wg.warps = std::min(wg.warps, keep_warps);
        }
        R = 4 * wg.wd;   // tile rows = 256 R = 1024 wd
        geo.ctas = 1; geo.stages = int(wg.stages);
    }
    const uint32_t tile_rows = 256 * R;
    const size_t stage_bytes = use_warp ? wg.stage_bytes : stage_bytes_for(R);
    const size_t smem_bytes = use_warp ? size_t(wg.warps) * wg.warp_bytes + code_smem_bytes : 128 + size_t(geo.stages) * stage_bytes + code_smem_bytes;

    // scheduling units.  Single-leaf kernel: a tile (contiguous tile range per CTA); warp kernel: a CHUNK of up to sched_chunk tiles of one pack spread evenly over the pack (tiles j, j + nch, j + 2 nch, …: every chunk
    // samples all regions of a time-ordered pack, so chunks cost the same whatever part of the pack a range predicate
    // selects, and a warp changes pack once per chunk) — `tile0` / `ntiles` / `tile_pack` count chunks there
    uint32_t sched_chunk = 8;
    if (const char* e = getenv("KX_SCHED_CHUNK")) sched_chunk = uint32_t(std::max(1, std::min(atoi(e), 64)));
    while (sched_chunk & (sched_chunk - 1)) sched_chunk &= sched_chunk - 1;   // a power of two (the kernel shifts)
    if (simple) sched_chunk = 1;
    if (use_warp) {   // small scans: smaller chunks first, so that every warp of the persistent grid gets work
        while (sched_chunk > 1 && total_rows / (uint64_t(tile_rows) * sched_chunk) < uint64_t(4) * ctx->num_sms * wg.warps) sched_chunk /= 2;
        // the chunks are dealt round-robin to num_sms x warps pipelines, so the scan takes ceil(chunks / pipelines) rounds:
        // with few rounds the last, partly filled one costs up to 1 / rounds of the run (128 packs x 1 Mi rows at 8 tiles per
        // chunk: 3.46 -> 4 rounds, 13 % idle).  Halve the chunk while that removes more than the ~3 % a smaller chunk costs
        // (one descriptor reload per chunk, profiles/r2_tune_warp.txt) — unless the caller pinned it
        if (!getenv("KX_SCHED_CHUNK")) {
            const uint64_t slots = uint64_t(ctx->num_sms) * wg.warps;
            auto fill = [&](uint32_t c) {
                uint64_t units = 0;
                for (int p = 0; p < npacks; ++p) {
                    const uint64_t tiles = (uint64_t(job.nrows[size_t(p)]) + tile_rows - 1) / tile_rows;
                    units += (tiles + c - 1) / c;
                }
                const uint64_t rounds = (units + slots - 1) / slots;
                return rounds ? double(units) / double(rounds * slots) : 1.0;
            };
            double f = fill(sched_chunk);
            while (sched_chunk > 1 && f < 0.95) {
                const double f2 = fill(sched_chunk / 2);
                if (f2 < f + 0.03) break;
                sched_chunk /= 2; f = f2;
            }
        }
    }
    uint64_t ntiles64 = 0;
    std::vector<uint32_t> tile0(size_t(npacks) + 1);
    for (int p = 0; p < npacks; ++p) {
        tile0[size_t(p)] = uint32_t(ntiles64);
        const uint64_t tiles = (uint64_t(job.nrows[size_t(p)]) + tile_rows - 1) / tile_rows;
        ntiles64 += use_warp ? (tiles + sched_chunk - 1) / sched_chunk : tiles;
    }
    if (ntiles64 > 0xfffffff0ull) return fail(ctx, KX_EINVAL, "too many tiles in one scan call");
    const uint32_t ntiles = uint32_t(ntiles64);
    size_t sz_tiles = uniform ? 0 : sizeof(uint32_t) * size_t(ntiles);
    size_t off_cjobs = off_tiles + round_up(sz_tiles, 256);
    size_t off_rjobs = off_cjobs + round_up(sizeof(CodesetJob) * cjobs.size(), 256);
    size_t off_ajobs = off_rjobs + round_up(sizeof(RunFillJob) * rjobs.size(), 256);
    size_t off_sjobs = off_ajobs + round_up(sizeof(AlpFixJob) * ajobs.size(), 256);
    size_t off_vjobs = off_sjobs + round_up(sizeof(StrJob) * sjobs.size(), 256);
    size_t desc_bytes = off_vjobs + round_up(sizeof(ValJob) * vjobs.size(), 256);
    if (leafbits_bytes) {
        CK(ctx->d_leafbits.reserve(leafbits_bytes));
        for (size_t i = 0; i < rjobs.size(); ++i) pl[rjob_leaf[i]].data = static_cast<const uint8_t*>(ctx->d_leafbits.p) + rjobs[i].out_off;
        for (size_t i = 0; i < ajobs.size(); ++i) pl[ajob_leaf[i]].fix = static_cast<const uint8_t*>(ctx->d_leafbits.p) + ajobs[i].out_off;
        for (size_t i = 0; i < sjobs.size(); ++i) pl[sjob_leaf[i]].data = static_cast<const uint8_t*>(ctx->d_leafbits.p) + sjobs[i].out_off;
        for (size_t i = 0; i < vjobs.size(); ++i) pl[vjob_leaf[i]].data = static_cast<const uint8_t*>(ctx->d_leafbits.p) + vjobs[i].out_off;
        for (auto& mj : mask_jobs) pl[size_t(mj.second) * nl + nleaves].data = static_cast<const uint8_t*>(ctx->d_leafbits.p) + mj.first;
    }

    CK(ctx->h_desc.reserve(desc_bytes));
    CK(ctx->d_packs.reserve(desc_bytes));
    uint8_t* hd = static_cast<uint8_t*>(ctx->h_desc.p);
    PackInfo* h_packs = reinterpret_cast<PackInfo*>(hd);
    size_t bitset_total = 0;
    for (int p = 0; p < npacks; ++p) {
        size_t off = dev_bits ? bitset_off[p] : 0;
        if (dev_bits && (off & 7)) return fail(ctx, KX_EINVAL, "bitset_off must be a multiple of 8");
        h_packs[p] = PackInfo{job.nrows[size_t(p)], tile0[size_t(p)], off};
        if (dev_bits) bitset_total = std::max(bitset_total, off + (size_t(job.nrows[size_t(p)]) + 7) / 8);
    }
    std::memcpy(hd + off_leaves, pl.data(), sz_leaves);
    ColView* h_views = reinterpret_cast<ColView*>(hd + off_views);
    if (nleaves) std::memcpy(h_views, job.leaf_views.data(), sizeof(ColView) * size_t(npacks) * nleaves);
    if (naggs) std::memcpy(h_views + size_t(npacks) * nleaves, job.agg_views.data(), sizeof(ColView) * size_t(npacks) * naggs);
    if (!uniform) {
        uint32_t* tp = reinterpret_cast<uint32_t*>(hd + off_tiles);
        for (int p = 0; p < npacks; ++p)
            for (uint32_t t = tile0[size_t(p)]; t < (p + 1 < npacks ? tile0[size_t(p) + 1] : ntiles); ++t) tp[t] = uint32_t(p);
    }

    if (!cjobs.empty()) std::memcpy(hd + off_cjobs, cjobs.data(), sizeof(CodesetJob) * cjobs.size());
    if (!rjobs.empty()) std::memcpy(hd + off_rjobs, rjobs.data(), sizeof(RunFillJob) * rjobs.size());
    if (!ajobs.empty()) std::memcpy(hd + off_ajobs, ajobs.data(), sizeof(AlpFixJob) * ajobs.size());
    if (!sjobs.empty()) std::memcpy(hd + off_sjobs, sjobs.data(), sizeof(StrJob) * sjobs.size());
    if (!vjobs.empty()) std::memcpy(hd + off_vjobs, vjobs.data(), sizeof(ValJob) * vjobs.size());

    // ---- launch geometry: persistent grid, static contiguous tile ranges
    const uint64_t max_grid = uint64_t(ctx->num_sms) * geo.ctas;
    int grid = int(std::min<uint64_t>(use_warp ? (uint64_t(ntiles) + wg.warps - 1) / std::max(wg.warps, 1u) : ntiles, max_grid));
    if (grid < 1) grid = 1;
    if (code_words) CK(ctx->d_codebits.reserve(size_t(code_words) * 4));

    CK(ctx->d_counts.reserve(sizeof(unsigned long long) * size_t(npacks)));
    if (dev_bits) CK(ctx->d_bitsets.reserve(round_up(bitset_total, 8) + 64));
    const uint64_t sel_words = so ? (round_up(bitset_total, 4) / 4) : 0;
    if (so) {
        CK(ctx->d_tmp2.reserve(size_t(so->cap) * 4 + 64));                       // selection ids
        CK(ctx->d_misc.reserve(64 + size_t((sel_words + 255) / 256) * 4));      // total + per-block counts
    }
    if (naggs) {
        CK(ctx->d_partials.reserve(sizeof(AggPartial) * size_t(grid) * naggs));
        CK(ctx->d_aggout.reserve(sizeof(AggPartial) * MAX_AGGS + 16));   // combined aggregates + the finished-CTA counter
    }
    CK(ctx->h_res.reserve(sizeof(unsigned long long) * size_t(npacks) + sizeof(RankPartial) + 128));

    uint8_t* dd = static_cast<uint8_t*>(ctx->d_packs.p);
    ScanParams P{};
    P.packs = reinterpret_cast<const PackInfo*>(dd);
    P.leaves = reinterpret_cast<const PackLeaf*>(dd + off_leaves);
    P.views = reinterpret_cast<const ColView*>(dd + off_views);
    P.tile_pack = uniform ? nullptr : reinterpret_cast<const uint32_t*>(dd + off_tiles);
    P.set_vals = prog->dev_sets;
    P.set_tabs = prog->dev_tabs;
    P.code_bits = static_cast<const uint32_t*>(ctx->d_codebits.p);
    std::memcpy(P.tab_off, prog->tab_off, sizeof(prog->tab_off));
    std::memcpy(P.tab_log2, prog->tab_log2, sizeof(prog->tab_log2));
    P.set_pre = prog->dev_pres;
    std::memcpy(P.pre_off, prog->pre_off, sizeof(prog->pre_off));
    for (int l = 0; l < MAX_LEAVES; ++l) P.pre_log2[l] = hash_leaf[l] ? prog->pre_log2[l] : 0;
    std::memcpy(P.hs_smem_off, hs_smem_off, sizeof(P.hs_smem_off));
    std::memcpy(P.hs_tab_smem_off, hs_tab_smem_off, sizeof(P.hs_tab_smem_off));
    P.stages = uint32_t(geo.stages);
    std::memcpy(P.code_smem_off, code_smem_off, sizeof(P.code_smem_off));
    P.code_smem_words = code_smem_words;
    P.code_bitmap_words = code_bitmap_words;
    P.stack_depth = stack_depth;
    P.flat_op = flat_op;
    P.sched_chunk = sched_chunk;
    P.agg_dense_thr = agg_dense_thr;
    if (use_warp) {
        P.w_wd = wg.wd; P.w_warps = wg.warps; P.w_warp_bytes = wg.warp_bytes; P.w_stage_off = wg.stage_off; P.w_ncols = wg.ncols;
        P.w_mw_slots = wg.mw_slots; P.w_chunk_rows = wg.chunk_rows;
        std::memcpy(P.w_slot, wg.slot, sizeof(P.w_slot));
        std::memcpy(P.w_slot_off, wg.slot_off, sizeof(P.w_slot_off));
        std::memcpy(P.w_col_off, wg.col_off, sizeof(P.w_col_off));
        std::memcpy(P.w_fix_off, wg.fix_off, sizeof(P.w_fix_off));
    }
    P.bitsets = dev_bits ? static_cast<uint8_t*>(ctx->d_bitsets.p) : nullptr;
    P.counts = static_cast<unsigned long long*>(ctx->d_counts.p);
    P.partials = static_cast<AggPartial*>(ctx->d_partials.p);
    P.agg_out = static_cast<AggPartial*>(ctx->d_aggout.p);
    P.done = naggs ? reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->d_aggout.p) + sizeof(AggPartial) * MAX_AGGS) : nullptr;
    P.npacks = uint32_t(npacks); P.ntiles = ntiles;
    P.nleaves = uint32_t(nl); P.npost = uint32_t(post.size()); P.naggs = uint32_t(naggs);
    P.R = R; P.tiles_per_pack = uniform ? (ntiles / uint32_t(npacks)) : 0; P.stage_bytes = uint32_t(stage_bytes);
    std::memcpy(P.set_off, prog->set_off, sizeof(prog->set_off));
    P.agg_view0 = uint32_t(size_t(npacks) * nleaves); P.leaf_view0 = 0;
    std::memcpy(P.postfix, post.data(), post.size());
    for (int j = 0; j < naggs; ++j) P.agg_type[j] = aggs[j].block_type;
    if (uniform && P.tiles_per_pack == 0) P.tiles_per_pack = 1;   // all packs empty

    LaunchArgs A;
    A.P = P; A.grid = grid; A.smem_bytes = smem_bytes; A.simple = simple; A.only32 = only32; A.ctas = geo.ctas;
    A.npacks = npacks; A.naggs = naggs; A.ntiles = ntiles; A.total_rows = total_rows;
    A.hd = hd; A.dd = dd; A.desc_bytes = desc_bytes;
    A.leafbits_bytes = leafbits_bytes; A.code_words = code_words;
    A.mask_jobs = &mask_jobs; A.row_masks = row_masks; A.nrows = job.nrows.data();
    A.ajobs = ajobs.empty() ? nullptr : reinterpret_cast<const AlpFixJob*>(dd + off_ajobs); A.najobs = uint32_t(ajobs.size()); A.max_patches = max_patches;
    A.sjobs = sjobs.empty() ? nullptr : reinterpret_cast<const StrJob*>(dd + off_sjobs); A.nsjobs = uint32_t(sjobs.size()); A.max_str_rows = max_str_rows;
    A.vjobs = vjobs.empty() ? nullptr : reinterpret_cast<const ValJob*>(dd + off_vjobs); A.nvjobs = uint32_t(vjobs.size()); A.max_val_rows = max_val_rows;
    A.rjobs = rjobs.empty() ? nullptr : reinterpret_cast<const RunFillJob*>(dd + off_rjobs); A.nrjobs = uint32_t(rjobs.size()); A.max_runs = max_runs;
    A.cjobs = cjobs.empty() ? nullptr : reinterpret_cast<const CodesetJob*>(dd + off_cjobs); A.ncjobs = uint32_t(cjobs.size()); A.max_code_set = max_code_set;
    A.prog = prog; A.bitsets = bitsets; A.bitset_total = bitset_total; A.dev_bits = dev_bits; A.counts = counts; A.aggs = aggs; A.agg_out = agg_out;
    A.so = so; A.sel_words = sel_words; A.sh = sh;
    int rc = launch_and_collect(ctx, A);
    if (rc) return rc;
    // a plan without pre-pass kernels, masks or selection output is remembered: the next call with the same program, pack
    // list and outputs skips block lookup, leaf translation and the descriptor upload (kx_scan / kx_scan_sharded)
    if (cache_key && ntiles && !masked && !so && !keep_layout && ajobs.empty() && sjobs.empty() && vjobs.empty() && rjobs.empty() && cjobs.empty())
        plan_cache_store(ctx, *cache_key, A, desc_bytes, off_leaves, off_views, uniform ? size_t(0) : off_tiles);
    return KX_OK;
}

int check_prog_job(kx_ctx* ctx, const kx_prog* prog, bool scan = false) {
    if (!prog || prog->ctx != ctx) return fail(ctx, KX_EINVAL, "program does not belong to this context");
    if (scan && prog->prune_only) return fail(ctx, KX_EUNSUPPORTED, "programs with byte-string leaves can only be used for pruning");
    return KX_OK;
}

// temporary single-block helpers for the narrow drop-ins
struct TempBlock {
    kx_ctx* ctx; StoredBlock sb;
    explicit TempBlock(kx_ctx* c) : ctx(c) {}
    ~TempBlock() { free_block(ctx, sb); }
};

int make_prog(kx_ctx* ctx, const kx_leaf* leaves, int nleaves, const uint8_t* postfix, int npost, kx_prog** out);

int single_leaf_scan(kx_ctx* ctx, const StoredBlock& sb, uint8_t block_type, uint8_t mode, uint64_t a, uint64_t b,
                     const uint64_t* set, uint32_t nset, uint8_t* bits, int64_t* count) {
    kx_leaf lf{}; lf.field = 0; lf.block_type = block_type; lf.mode = mode; lf.a = a; lf.b = b; lf.set = set; lf.nset = nset;
    uint8_t pf = 0;
    kx_prog* prog = nullptr;
    int rc = make_prog(ctx, &lf, 1, &pf, 1, &prog);
    if (rc) return rc;
    ScanJob job; job.npacks = 1; job.nrows = {sb.view.n}; job.leaf_views = {sb.view};
    job.leaf_dicts = {sb.dict.empty() ? nullptr : sb.dict.data()};
    size_t off = 0;
    rc = run_scan(ctx, prog, job, bits, &off, count, nullptr, 0, nullptr);
    delete prog;   // (its destructor frees the device copies of the sets)
    return rc;
}

int make_prog(kx_ctx* ctx, const kx_leaf* leaves, int nleaves, const uint8_t* postfix, int npost, kx_prog** out) {
    if (!leaves || nleaves < 1 || nleaves > MAX_LEAVES) return fail(ctx, KX_EINVAL, "program needs 1..8 leaves");
    if (!postfix || npost < 1 || npost > 2 * MAX_LEAVES) return fail(ctx, KX_EINVAL, "bad postfix length");
    int depth = 0;
    for (int i = 0; i < npost; ++i) {
        if (postfix[i] < 0x80) { if (postfix[i] >= nleaves) return fail(ctx, KX_EINVAL, "postfix references unknown leaf"); depth++; }
        else if (postfix[i] == KX_OP_AND || postfix[i] == KX_OP_OR) { if (depth < 2) return fail(ctx, KX_EINVAL, "postfix stack underflow"); depth--; }
        else return fail(ctx, KX_EINVAL, "bad postfix opcode");
    }
    if (depth != 1) return fail(ctx, KX_EINVAL, "postfix does not reduce to one bitset");
    auto p = std::make_unique<kx_prog>();
    static std::atomic<uint64_t> next_prog_id{1};
    p->ctx = ctx; p->id = next_prog_id++;
    p->postfix.assign(postfix, postfix + npost);
    for (int l = 0; l < nleaves; ++l) {
        const kx_leaf& in = leaves[l];
        LeafSpec s; s.field = in.field; s.type = in.block_type; s.mode = in.mode; s.a = in.a; s.b = in.b;
        if (s.type == KX_BYTES) {
            s.set_off = uint32_t(p->sets.size()); p->set_off[l] = s.set_off;
            if (in.nset == 1 && in.mode != KX_MODE_IN && in.mode != KX_MODE_NIN) {
                // row-level string predicate: operand bytes at `set`, a = length of the operand, b = length of the
                // upper bound that follows it (RANGE); StringMatcher has the seven scalar modes (types/strings.go)
                if (s.mode < KX_MODE_EQ || s.mode > KX_MODE_RANGE) return fail(ctx, KX_EINVAL, "leaf: unsupported filter mode");
                if ((in.a || in.b) && !in.set) return fail(ctx, KX_EINVAL, "leaf: operand bytes missing");
                if (in.a > (1u << 24) || in.b > (1u << 24)) return fail(ctx, KX_EINVAL, "leaf: operand too long");
                const uint8_t* ob = reinterpret_cast<const uint8_t*>(in.set);
                s.sa.assign(ob, ob + in.a);
                if (s.mode == KX_MODE_RANGE) s.sb.assign(ob + in.a, ob + in.a + in.b);
                s.sa_off = uint32_t(p->strs.size()); p->strs.insert(p->strs.end(), s.sa.begin(), s.sa.end());
                s.sb_off = uint32_t(p->strs.size()); p->strs.insert(p->strs.end(), s.sb.begin(), s.sb.end());
                s.a = s.b = 0;
                s.has_str = true;
            } else if ((in.mode == KX_MODE_IN || in.mode == KX_MODE_NIN) && in.nset >= 1 && in.set) {
                // row-level set membership (match_bytes.go:392-520): the strings are sorted (bytes.Compare order) and
                // de-duplicated like slicex.OrderedBytes.SetUnique; the device pool gets [count x (offset, length)] + bytes
                const uint8_t* ob = reinterpret_cast<const uint8_t*>(in.set);
                const uint64_t head = uint64_t(in.nset) * 4;
                if (in.nset > (1u << 20) || in.a < head || in.a > (1ull << 28)) return fail(ctx, KX_EINVAL, "leaf: bad byte-string set");
                std::vector<std::pair<const uint8_t*, uint32_t>> items;
                uint64_t pos = head;
                for (uint32_t i = 0; i < in.nset; ++i) {
                    uint32_t len; std::memcpy(&len, ob + size_t(i) * 4, 4);
                    if (pos + len > in.a) return fail(ctx, KX_EINVAL, "leaf: byte-string set overruns its buffer");
                    items.push_back({ob + pos, len});
                    pos += len;
                }
                auto less = [](const std::pair<const uint8_t*, uint32_t>& x, const std::pair<const uint8_t*, uint32_t>& y) {
                    const int c = std::memcmp(x.first, y.first, std::min(x.second, y.second));
                    return c < 0 || (c == 0 && x.second < y.second);
                };
                std::sort(items.begin(), items.end(), less);
                items.erase(std::unique(items.begin(), items.end(), [&](const auto& x, const auto& y) { return !less(x, y) && !less(y, x); }), items.end());
                while (p->strs.size() & 3) p->strs.push_back(0);   // the (offset, length) table is read as uint32
                s.sa_off = uint32_t(p->strs.size());
                const uint32_t cnt = uint32_t(items.size());
                s.sa.resize(size_t(cnt) * 8);                       // host copy of the table + bytes (constant blocks are decided on the host)
                uint32_t boff = cnt * 8;
                for (uint32_t i = 0; i < cnt; ++i) {
                    std::memcpy(s.sa.data() + size_t(i) * 8, &boff, 4);
                    std::memcpy(s.sa.data() + size_t(i) * 8 + 4, &items[i].second, 4);
                    boff += items[i].second;
                }
                for (auto& it : items) s.sa.insert(s.sa.end(), it.first, it.first + it.second);
                p->strs.insert(p->strs.end(), s.sa.begin(), s.sa.end());
                s.sb_off = s.sa_off;
                s.a = cnt; s.b = 0;                                 // a = number of distinct strings
                s.has_str = true;
            } else {
                // no operand: the leaf takes part in pruning only (bloom probes with caller-supplied hashes)
                if (s.mode != KX_MODE_EQ && s.mode != KX_MODE_IN && s.mode != KX_MODE_NE && s.mode != KX_MODE_NIN)
                    return fail(ctx, KX_EUNSUPPORTED, "leaf: byte-string leaves without operand bytes support EQ/NE/IN/NIN (pruning) only");
                p->prune_only = true;
            }
            p->leaves.push_back(std::move(s));
            continue;
        }
        if (type_bits(s.type) == 0) return fail(ctx, KX_EINVAL, "leaf: unsupported block type");
        if (s.mode < KX_MODE_EQ || s.mode > KX_MODE_RANGE) return fail(ctx, KX_EINVAL, "leaf: unsupported filter mode");
        if (s.mode == KX_MODE_IN || s.mode == KX_MODE_NIN) {
            // float IN/NIN are no-ops in the reference containers (float_raw.go:209-210) and are
            // matched by a separate matcher; out of scope here
            if (type_is_float(s.type)) return fail(ctx, KX_EUNSUPPORTED, "IN/NIN on float blocks is not supported");
            if (in.nset && !in.set) return fail(ctx, KX_EINVAL, "leaf: set pointer missing");
            s.set.assign(in.set, in.set + in.nset);
            std::sort(s.set.begin(), s.set.end());
            s.set.erase(std::unique(s.set.begin(), s.set.end()), s.set.end());
        }
        s.set_off = uint32_t(p->sets.size());
        p->set_off[l] = s.set_off;
        p->sets.insert(p->sets.end(), s.set.begin(), s.set.end());
        if (!s.set.empty()) {
            std::vector<uint64_t> slots; int lg = 0;
            if (build_set_table(s.set, slots, lg)) {
                s.has_table = true;
                p->tab_off[l] = uint32_t(p->tabs.size()); p->tab_log2[l] = uint8_t(lg);
                p->tabs.insert(p->tabs.end(), slots.begin(), slots.end());
                std::vector<uint32_t> pre; int plg = 0;
                build_set_prefilter(s.set, pre, plg);
                p->pre_off[l] = uint32_t(p->pres.size()); p->pre_log2[l] = uint8_t(plg);
                p->pres.insert(p->pres.end(), pre.begin(), pre.end());
            }
        }
        p->leaves.push_back(std::move(s));
    }
    for (int l = nleaves; l <= MAX_LEAVES; ++l) p->set_off[l] = uint32_t(p->sets.size());
    if (!p->sets.empty()) {
        // one allocation: sorted sets, then (32 B aligned) the hash tables
        size_t set_bytes = round_up(p->sets.size() * 8, 32), tab_bytes = p->tabs.size() * 8, pre_bytes = p->pres.size() * 4;
        CK(cudaMalloc(&p->dev_sets, set_bytes + tab_bytes + pre_bytes + 32));
        CK(cudaMemcpyAsync(p->dev_sets, p->sets.data(), p->sets.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (tab_bytes) {
            uint8_t* dt = reinterpret_cast<uint8_t*>(p->dev_sets) + set_bytes;
            CK(cudaMemcpyAsync(dt, p->tabs.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream));
            p->dev_tabs = reinterpret_cast<const uint64_t*>(dt);
            CK(cudaMemcpyAsync(dt + tab_bytes, p->pres.data(), pre_bytes, cudaMemcpyHostToDevice, ctx->stream));
            p->dev_pres = reinterpret_cast<const uint32_t*>(dt + tab_bytes);
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    if (!p->strs.empty()) {
        CK(cudaMalloc(&p->dev_strs, p->strs.size() + 16));
        CK(cudaMemcpyAsync(p->dev_strs, p->strs.data(), p->strs.size(), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    *out = p.release();
    return KX_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int kx_abi_version(void) { return KX_ABI_VERSION; }

int kx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int kx_ctx_create(int device, size_t hbm_budget, kx_ctx** out) {
    return kx_guarded<int>(nullptr, [&]() -> int {
    if (!out) return KX_EINVAL;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, KX_ENODEV, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if (device < 0 || device >= ndev) return fail(nullptr, KX_EINVAL, "bad device ordinal");
    kx_ctx* ctx = nullptr;   // for CK
    CK(cudaSetDevice(device));
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(nullptr, KX_ENODEV, "libknoxgpu holds sm_100a code only: it needs a compute capability 10.x (Blackwell B200) device");
    auto c = std::make_unique<kx_ctx>();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    c->budget = hbm_budget ? hbm_budget : free_b / 10 * 8;
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev_start)); CK(cudaEventCreate(&c->ev_k0)); CK(cudaEventCreate(&c->ev_k1)); CK(cudaEventCreate(&c->ev_end));
    CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_copy[0], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c->ev_copy[1], cudaEventDisableTiming));
    *out = c.release();
    return KX_OK;
    });
}

void kx_ctx_destroy(kx_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->slabs) if (s.base) cudaFree(s.base);
    cudaEventDestroy(ctx->ev_start); cudaEventDestroy(ctx->ev_k0); cudaEventDestroy(ctx->ev_k1); cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_copy[0]) cudaEventDestroy(ctx->ev_copy[0]);
    if (ctx->ev_copy[1]) cudaEventDestroy(ctx->ev_copy[1]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    comm_destroy(ctx->comm);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* kx_last_error(kx_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

void* kx_host_alloc(kx_ctx* ctx, size_t bytes) {
    if (!ctx) return nullptr;
    cudaSetDevice(ctx->device);
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); ctx->err = "cudaMallocHost failed"; return nullptr; }
    return p;
}
void kx_host_free(kx_ctx* ctx, void* p) { if (ctx && p) { cudaSetDevice(ctx->device); cudaFreeHost(p); } }

int kx_block_put(kx_ctx* ctx, uint32_t pack, uint32_t version, uint16_t field, uint8_t block_type, const void* enc, size_t len,
                 uint32_t* nrows_out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!enc || !len) return fail(ctx, KX_EINVAL, "empty block");
    CK(cudaSetDevice(ctx->device));
    BlockLayout lay; StrLayout slay; std::string err;
    int rc = block_type == KX_BYTES ? normalize_string_block(static_cast<const uint8_t*>(enc), len, slay, err)
                                    : normalize_block(block_type, static_cast<const uint8_t*>(enc), len, lay, err);
    if (rc) return fail(ctx, rc, "kx_block_put: " + err);
    BlockKey key{pack, version, field};
    // the new block is uploaded BEFORE a resident one with the same key is released: a failed put leaves the store as it was
    StoredBlock sb;
    rc = block_type == KX_BYTES ? upload_string_block(ctx, slay, sb) : upload_block(ctx, lay, sb);
    // the caller's buffer may be reused after return (cgo rule): finish the copy
    if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = fail(ctx, KX_ECUDA, "kx_block_put: copy failed");
    if (rc) { free_block(ctx, sb); return rc; }
    ctx->store_epoch++;
    auto it = ctx->store.find(key);
    if (it != ctx->store.end()) { ctx->store_enc_bytes -= it->second.enc_len; free_block(ctx, it->second); ctx->store.erase(it); }
    sb.enc_len = len;
    ctx->store_enc_bytes += len;
    if (nrows_out) *nrows_out = sb.view.n;
    ctx->store.emplace(key, std::move(sb));
    return KX_OK;
    });
}

int kx_block_drop(kx_ctx* ctx, uint32_t pack, uint32_t version, uint16_t field) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    auto it = ctx->store.find(BlockKey{pack, version, field});
    if (it == ctx->store.end()) return fail(ctx, KX_ENOTFOUND, "block not resident");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->store_epoch++;
    ctx->store_enc_bytes -= it->second.enc_len;
    free_block(ctx, it->second);
    ctx->store.erase(it);
    return KX_OK;
    });
}

int kx_store_stats(kx_ctx* ctx, uint64_t* nblocks, uint64_t* encoded_bytes, uint64_t* device_bytes, uint64_t* slab_bytes) {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nblocks) *nblocks = ctx->store.size();
    if (encoded_bytes) *encoded_bytes = ctx->store_enc_bytes;
    if (device_bytes) *device_bytes = ctx->store_dev_bytes;
    if (slab_bytes) *slab_bytes = ctx->slab_bytes;
    return KX_OK;
}

int kx_prog_compile(kx_ctx* ctx, const kx_leaf* leaves, int nleaves, const uint8_t* postfix, int npost, kx_prog** out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx || !out) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    CK(cudaSetDevice(ctx->device));
    return make_prog(ctx, leaves, nleaves, postfix, npost, out);
    });
}

void kx_prog_free(kx_prog* prog) {
    if (!prog) return;
    cudaSetDevice(prog->ctx->device);
    delete prog;
}

int kx_scan(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, uint8_t* bitsets, const size_t* bitset_off,
            int64_t* counts, const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (npacks < 0 || (npacks && !packs)) return fail(ctx, KX_EINVAL, "bad pack list");
    if (naggs && (!aggs || !agg_out)) return fail(ctx, KX_EINVAL, "aggregate buffers missing");
    if (bitsets && !bitset_off) return fail(ctx, KX_EINVAL, "bitsets without bitset_off");
    CK(cudaSetDevice(ctx->device));
    const PlanKey key{prog, packs, npacks, aggs, naggs, bitsets != nullptr, bitset_off, env_knobs_hash()};
    bool hit = false;
    rc = plan_cache_run(ctx, key, bitsets, counts, agg_out, nullptr, &hit);
    if (hit) return rc;
    ScanJob job;
    if ((rc = build_scan_job(ctx, prog, packs, npacks, aggs, naggs, job))) return rc;
    return run_scan(ctx, prog, job, bitsets, bitset_off, counts, aggs, naggs, agg_out, nullptr, nullptr, nullptr, nullptr, &key);
    });
}

int kx_scan_sharded(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, int64_t* counts,
                    const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out, int64_t* total_count) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (npacks < 0 || (npacks && !packs)) return fail(ctx, KX_EINVAL, "bad pack list");
    if (naggs < 0 || naggs > MAX_AGGS || (naggs && (!aggs || !agg_out))) return fail(ctx, KX_EINVAL, "aggregate buffers missing");
    CK(cudaSetDevice(ctx->device));
    ShardOut sh; sh.total_count = total_count;
    const PlanKey key{prog, packs, npacks, aggs, naggs, false, nullptr, env_knobs_hash()};
    bool hit = false;
    rc = plan_cache_run(ctx, key, nullptr, counts, agg_out, &sh, &hit);
    if (hit) return rc;
    ScanJob job;
    if ((rc = build_scan_job(ctx, prog, packs, npacks, aggs, naggs, job))) return rc;
    return run_scan(ctx, prog, job, nullptr, nullptr, counts, aggs, naggs, agg_out, nullptr, nullptr, &sh, nullptr, &key);
    });
}

int kx_scan_ex(kx_ctx* ctx, const kx_prog* prog, const kx_scan_args* a) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (!a || a->struct_size < sizeof(kx_scan_args)) return fail(ctx, KX_EINVAL, "kx_scan_ex: bad argument block");
    if (a->npacks < 0 || (a->npacks && !a->packs)) return fail(ctx, KX_EINVAL, "bad pack list");
    if (a->naggs < 0 || a->naggs > MAX_AGGS || (a->naggs && (!a->aggs || !a->agg_out))) return fail(ctx, KX_EINVAL, "aggregate buffers missing");
    const bool want_sel = a->sel_off != nullptr;
    if (want_sel && (a->bitsets || (a->sel_cap && !a->sel))) return fail(ctx, KX_EINVAL, "kx_scan_ex: selection vectors and host bitsets cannot be requested together");
    CK(cudaSetDevice(ctx->device));
    ScanJob job;
    if ((rc = build_scan_job(ctx, prog, a->packs, a->npacks, a->aggs, a->naggs, job))) return rc;
    SelectOut so; so.sel = a->sel; so.cap = a->sel_cap; so.off = a->sel_off;
    ShardOut sh; sh.total_count = a->total_count;
    rc = run_scan(ctx, prog, job, a->bitsets, a->bitset_off, a->counts, a->aggs, a->naggs, a->agg_out, want_sel ? &so : nullptr, nullptr,
                  (a->flags & KX_SCAN_SHARDED) ? &sh : nullptr, a->row_masks);
    if (rc) return rc;
    if (want_sel && so.overflow) return fail(ctx, KX_ENOMEM, "kx_scan_ex: selection buffer too small (sel_off[npacks] holds the required size)");
    return KX_OK;
    });
}

int kx_comm_unique_id(void* id_out, size_t cap) {
    if (!id_out || cap < KX_COMM_ID_BYTES) return KX_EINVAL;
    std::string err;
    if (comm_unique_id(id_out, err)) return fail(nullptr, KX_EUNSUPPORTED, "kx_comm_unique_id: " + err);
    return KX_OK;
}

int kx_comm_init(kx_ctx* ctx, int nranks, int rank, const void* id, size_t id_len) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nranks < 1 || rank < 0 || rank >= nranks || (nranks > 1 && (!id || id_len < KX_COMM_ID_BYTES))) return fail(ctx, KX_EINVAL, "kx_comm_init: bad arguments");
    if (ctx->comm) return fail(ctx, KX_EINVAL, "kx_comm_init: this context already has a communicator");
    CK(cudaSetDevice(ctx->device));
    std::string err;
    if (comm_create(nranks, rank, id, &ctx->comm, err)) return fail(ctx, KX_EUNSUPPORTED, "kx_comm_init: " + err);
    return KX_OK;
    });
}

int kx_comm_info(kx_ctx* ctx, int* nranks, int* rank, int* nccl_version) {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nranks) *nranks = comm_nranks(ctx->comm);
    if (rank) *rank = comm_rank(ctx->comm);
    if (nccl_version) *nccl_version = ctx->comm && comm_nranks(ctx->comm) > 1 ? comm_nccl_version() : 0;
    return KX_OK;
}

int kx_comm_allgather(kx_ctx* ctx, const void* send, void* recv, size_t bytes) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!bytes) return KX_OK;
    if (!send || !recv) return fail(ctx, KX_EINVAL, "kx_comm_allgather: null buffers");
    CK(cudaSetDevice(ctx->device));
    const size_t nranks = size_t(comm_nranks(ctx->comm)), padded = round_up(bytes, 16);
    CK(ctx->d_tmp.reserve(padded * (nranks + 1)));
    uint8_t* d = static_cast<uint8_t*>(ctx->d_tmp.p);
    CK(cudaMemcpyAsync(d, send, bytes, cudaMemcpyHostToDevice, ctx->stream));
    std::string err;
    if (comm_allgather(ctx->comm, d, d + padded, padded, ctx->stream, err)) return fail(ctx, KX_ECUDA, "kx_comm_allgather: " + err);
    for (size_t r = 0; r < nranks; ++r)
        CK(cudaMemcpyAsync(static_cast<uint8_t*>(recv) + r * bytes, d + padded * (r + 1), bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KX_OK;
    });
}

}  // extern "C"
// shared by kx_scan / kx_scan_select: resolve the blocks of a pack list
static int build_scan_job(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, const kx_agg_req* aggs, int naggs, ScanJob& job) {
    const int nleaves = int(prog->leaves.size());
    job.npacks = npacks;
    job.nrows.resize(size_t(npacks));
    job.leaf_views.resize(size_t(npacks) * nleaves);
    job.leaf_dicts.resize(size_t(npacks) * nleaves);
    job.leaf_cstr.assign(size_t(npacks) * nleaves, nullptr);
    job.agg_views.resize(size_t(npacks) * size_t(naggs));
    for (int p = 0; p < npacks; ++p) {
        for (int l = 0; l < nleaves; ++l) {
            auto it = ctx->store.find(BlockKey{packs[p].pack, packs[p].version, prog->leaves[size_t(l)].field});
            if (it == ctx->store.end()) return fail(ctx, KX_ENOTFOUND, "filter block not resident: pack " + std::to_string(packs[p].pack));
            if (it->second.view.type != prog->leaves[size_t(l)].type) return fail(ctx, KX_EINVAL, "leaf / block type mismatch");
            job.leaf_views[size_t(p) * nleaves + l] = it->second.view;
            job.leaf_dicts[size_t(p) * nleaves + l] = it->second.dict.empty() ? nullptr : it->second.dict.data();
            if (it->second.view.kind == CK_STR && it->second.view.is_raw == STR_CONST) job.leaf_cstr[size_t(p) * nleaves + l] = &it->second.cstr;
            if (l == 0) job.nrows[size_t(p)] = it->second.view.n;
        }
        for (int j = 0; j < naggs; ++j) {
            auto it = ctx->store.find(BlockKey{packs[p].pack, packs[p].version, aggs[j].field});
            if (it == ctx->store.end()) return fail(ctx, KX_ENOTFOUND, "value block not resident: pack " + std::to_string(packs[p].pack));
            if (it->second.view.type != aggs[j].block_type) return fail(ctx, KX_EINVAL, "aggregate / block type mismatch");
            job.agg_views[size_t(p) * naggs + j] = it->second.view;
        }
    }
    return KX_OK;
}

extern "C" {

int kx_scan_select(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, uint32_t* sel, size_t sel_cap, uint64_t* sel_off,
                   int64_t* counts, const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (npacks < 0 || (npacks && !packs) || !sel_off || (sel_cap && !sel)) return fail(ctx, KX_EINVAL, "kx_scan_select: bad arguments");
    if (naggs && (!aggs || !agg_out)) return fail(ctx, KX_EINVAL, "aggregate buffers missing");
    CK(cudaSetDevice(ctx->device));
    ScanJob job;
    if ((rc = build_scan_job(ctx, prog, packs, npacks, aggs, naggs, job))) return rc;
    SelectOut so; so.sel = sel; so.cap = sel_cap; so.off = sel_off;
    rc = run_scan(ctx, prog, job, nullptr, nullptr, counts, aggs, naggs, agg_out, &so);
    if (rc) return rc;
    if (so.overflow) return fail(ctx, KX_ENOMEM, "kx_scan_select: selection buffer too small (sel_off[npacks] holds the required size)");
    return KX_OK;
    });
}

int kx_scan_buckets(kx_ctx* ctx, const kx_prog* prog, const kx_packref* packs, int npacks, uint16_t ts_field, uint8_t ts_type,
                    const uint64_t* edges, int nbuckets, const kx_agg_req* aggs, int naggs, int64_t* bucket_counts, kx_agg_out* out,
                    int64_t* counts, const uint8_t* const* row_masks) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (npacks < 0 || (npacks && !packs) || nbuckets < 1 || !edges || naggs < 0 || naggs > MAX_AGGS || (naggs && (!aggs || !out)))
        return fail(ctx, KX_EINVAL, "kx_scan_buckets: bad arguments");
    if (type_bits(ts_type) == 0 || type_is_float(ts_type)) return fail(ctx, KX_EUNSUPPORTED, "window column must be an integer / timestamp block");
    const uint64_t ts_flip = type_is_signed(ts_type) ? 0x8000000000000000ull : 0ull;
    for (int k = 0; k < nbuckets; ++k)
        if ((edges[k] ^ ts_flip) > (edges[k + 1] ^ ts_flip)) return fail(ctx, KX_EINVAL, "window edges must ascend");
    for (int j = 0; j < naggs; ++j) {
        int t = aggs[j].block_type;
        if (type_bits(t) == 0 || t == KX_FLOAT32) return fail(ctx, KX_EUNSUPPORTED, "kx_scan_buckets: float32 value columns are not supported (widen them to float64)");
    }
    // the scan proper: filter -> match bitsets that stay on the device; the window column and the value columns ride
    // along as "aggregate" views of the job (looked up, not reduced by the scan)
    std::vector<kx_agg_req> cols(size_t(naggs) + 1);
    cols[0] = kx_agg_req{ts_field, ts_type, 0};
    for (int j = 0; j < naggs; ++j) cols[size_t(j) + 1] = aggs[j];
    ScanJob job;
    rc = build_scan_job(ctx, prog, packs, npacks, cols.data(), naggs + 1, job);
    if (rc) return rc;
    std::vector<ColView> views;
    views.swap(job.agg_views);   // [npacks][1 + naggs]
    for (int p = 0; p < npacks; ++p)
        for (int c = 0; c <= naggs; ++c)
            if (views[size_t(p) * (naggs + 1) + c].n != job.nrows[size_t(p)]) return fail(ctx, KX_EINVAL, "blocks of one pack differ in length");
    std::vector<size_t> layout;
    rc = run_scan(ctx, prog, job, nullptr, nullptr, counts, nullptr, 0, nullptr, nullptr, &layout, nullptr, row_masks);
    if (rc) return rc;
    const double scan_kernel_ms = ctx->last_kernel_ms, scan_total_ms = ctx->last_total_ms;

    // ---- bucket table + descriptors
    const size_t ncell = size_t(naggs) + 1, ncells = size_t(nbuckets) * ncell;
    std::vector<uint32_t> job0(size_t(npacks) + 1);
    uint64_t njobs = 0;
    for (int p = 0; p < npacks; ++p) {
        job0[size_t(p)] = uint32_t(njobs);
        njobs += ((uint64_t(job.nrows[size_t(p)]) + 31) / 32 + BUCKET_JOB_GROUPS - 1) / BUCKET_JOB_GROUPS;
    }
    job0[size_t(npacks)] = uint32_t(njobs);
    if (njobs > 0xfffffff0ull) return fail(ctx, KX_EINVAL, "too many rows in one kx_scan_buckets call");
    const size_t off_views = 0, off_edges = round_up(sizeof(ColView) * views.size(), 256);
    const size_t off_job0 = off_edges + round_up(8 * (size_t(nbuckets) + 1), 256);
    const size_t off_table = off_job0 + round_up(4 * job0.size(), 256);
    const size_t total = off_table + sizeof(BucketCell) * ncells;
    CK(ctx->h_aux.reserve(total));
    CK(ctx->d_tmp.reserve(total));
    uint8_t* h = static_cast<uint8_t*>(ctx->h_aux.p);
    if (!views.empty()) std::memcpy(h + off_views, views.data(), sizeof(ColView) * views.size());
    uint64_t* hk = reinterpret_cast<uint64_t*>(h + off_edges);
    for (int k = 0; k <= nbuckets; ++k) hk[k] = edges[k] ^ ts_flip;
    std::memcpy(h + off_job0, job0.data(), 4 * job0.size());
    BucketCell* ht = reinterpret_cast<BucketCell*>(h + off_table);
    for (size_t i = 0; i < ncells; ++i) ht[i] = BucketCell{0, 0, ~0ull, 0};
    uint8_t* d = static_cast<uint8_t*>(ctx->d_tmp.p);
    BucketParams P{};
    P.packs = static_cast<const PackInfo*>(ctx->d_packs.p);
    P.bits = static_cast<const uint8_t*>(ctx->d_bitsets.p);
    P.views = reinterpret_cast<const ColView*>(d + off_views);
    P.edges = reinterpret_cast<const uint64_t*>(d + off_edges);
    P.job0 = reinterpret_cast<const uint32_t*>(d + off_job0);
    P.table = reinterpret_cast<BucketCell*>(d + off_table);
    P.ts_flip = ts_flip;
    P.npacks = uint32_t(npacks); P.njobs = uint32_t(njobs); P.nbuckets = uint32_t(nbuckets); P.naggs = uint32_t(naggs);
    for (int j = 0; j < naggs; ++j) P.agg_type[j] = aggs[j].block_type;
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    CK(cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaEventRecord(ctx->ev_k0, ctx->stream));
    if (npacks) { CK(launch_bucket(P, ctx->num_sms, ctx->stream)); ctx->last_launches++; }
    CK(cudaEventRecord(ctx->ev_k1, ctx->stream));
    CK(cudaMemcpyAsync(ht, d + off_table, sizeof(BucketCell) * ncells, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1)); ctx->last_kernel_ms = scan_kernel_ms + ms;
    CK(cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end)); ctx->last_total_ms = scan_total_ms + ms;

    for (int k = 0; k < nbuckets; ++k) {
        if (bucket_counts) bucket_counts[k] = int64_t(ht[size_t(k) * ncell].count);
        for (int j = 0; j < naggs; ++j) {
            const BucketCell& c = ht[size_t(k) * ncell + 1 + size_t(j)];
            kx_agg_out o{};
            const int t = aggs[j].block_type;
            o.count = int64_t(c.count); o.valid = c.count ? 1 : 0;
            if (c.count) {
                if (t == KX_FLOAT64) {
                    o.sum_bits = c.sum;
                    o.min_bits = (c.mn >> 63) ? (c.mn & 0x7fffffffffffffffull) : ~c.mn;   // inverse of the order key
                    o.max_bits = (c.mx >> 63) ? (c.mx & 0x7fffffffffffffffull) : ~c.mx;
                } else {
                    const uint64_t flip = type_is_signed(t) ? 0x8000000000000000ull : 0;
                    o.sum_bits = type_ext(t, c.sum);
                    o.min_bits = c.mn ^ flip; o.max_bits = c.mx ^ flip;
                }
            }
            out[size_t(j) * size_t(nbuckets) + size_t(k)] = o;
        }
    }
    return KX_OK;
    });
}

int kx_gather(kx_ctx* ctx, const kx_packref* packs, int npacks, uint16_t field, uint8_t block_type, const uint32_t* sel,
              const uint64_t* sel_off, void* dst) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (npacks < 0 || (npacks && (!packs || !sel_off))) return fail(ctx, KX_EINVAL, "kx_gather: bad arguments");
    const int eb = type_bits(block_type) / 8;
    if (!eb) return fail(ctx, KX_EINVAL, "kx_gather: unsupported block type");
    if (npacks == 0 || sel_off[npacks] == 0) return KX_OK;
    if (!sel || !dst) return fail(ctx, KX_EINVAL, "kx_gather: null buffers");
    CK(cudaSetDevice(ctx->device));
    const uint64_t total = sel_off[npacks];
    std::vector<ColView> views(static_cast<size_t>(npacks));
    for (int p = 0; p < npacks; ++p) {
        auto it = ctx->store.find(BlockKey{packs[p].pack, packs[p].version, field});
        if (it == ctx->store.end()) return fail(ctx, KX_ENOTFOUND, "kx_gather: block not resident: pack " + std::to_string(packs[p].pack));
        if (it->second.view.type != block_type) return fail(ctx, KX_EINVAL, "kx_gather: block type mismatch");
        if (sel_off[p + 1] < sel_off[p]) return fail(ctx, KX_EINVAL, "kx_gather: sel_off must ascend");
        views[size_t(p)] = it->second.view;
        for (uint64_t i = sel_off[p]; i < sel_off[p + 1]; ++i)   // a row id past the block would be an out-of-bounds device read
            if (sel[i] >= it->second.view.n) return fail(ctx, KX_EINVAL, "kx_gather: row id " + std::to_string(sel[i]) + " outside pack " + std::to_string(packs[p].pack));
    }
    // device layout: views | sel_off | sel | dst
    size_t off_so = round_up(sizeof(ColView) * size_t(npacks), 256), off_sel = off_so + round_up(8 * (size_t(npacks) + 1), 256);
    size_t off_dst = off_sel + round_up(size_t(total) * 4, 256);
    CK(ctx->d_tmp.reserve(off_dst + size_t(total) * eb + 64));
    uint8_t* d = static_cast<uint8_t*>(ctx->d_tmp.p);
    CK(cudaMemcpyAsync(d, views.data(), sizeof(ColView) * size_t(npacks), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + off_so, sel_off, 8 * (size_t(npacks) + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + off_sel, sel, size_t(total) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_gather(reinterpret_cast<const ColView*>(d), reinterpret_cast<const unsigned long long*>(d + off_so), uint32_t(npacks),
                     reinterpret_cast<const uint32_t*>(d + off_sel), total, eb, d + off_dst, ctx->stream));
    CK(cudaMemcpyAsync(dst, d + off_dst, size_t(total) * eb, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // views / sel are host temporaries
    return KX_OK;
    });
}

int kx_gather_bytes(kx_ctx* ctx, const kx_packref* packs, int npacks, uint16_t field, const uint32_t* sel, const uint64_t* sel_off,
                    uint64_t* out_off, uint8_t* out, size_t out_cap) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (npacks < 0 || (npacks && (!packs || !sel_off)) || !out_off) return fail(ctx, KX_EINVAL, "kx_gather_bytes: bad arguments");
    out_off[0] = 0;
    if (npacks == 0 || sel_off[npacks] == 0) return KX_OK;
    if (!sel) return fail(ctx, KX_EINVAL, "kx_gather_bytes: null selection");
    CK(cudaSetDevice(ctx->device));
    const uint64_t total = sel_off[npacks];
    if (total >= 0xffffffffull) return fail(ctx, KX_EUNSUPPORTED, "kx_gather_bytes: more than 2^32 - 1 selected rows in one call");
    std::vector<ColView> views(static_cast<size_t>(npacks));
    for (int p = 0; p < npacks; ++p) {
        auto it = ctx->store.find(BlockKey{packs[p].pack, packs[p].version, field});
        if (it == ctx->store.end()) return fail(ctx, KX_ENOTFOUND, "kx_gather_bytes: block not resident: pack " + std::to_string(packs[p].pack));
        const ColView& v = it->second.view;
        if (v.kind != CK_STR) return fail(ctx, KX_EINVAL, "kx_gather_bytes: not a byte-string block");
        if (sel_off[p + 1] < sel_off[p]) return fail(ctx, KX_EINVAL, "kx_gather_bytes: sel_off must ascend");
        views[size_t(p)] = v;
        for (uint64_t i = sel_off[p]; i < sel_off[p + 1]; ++i)
            if (sel[i] >= v.n) return fail(ctx, KX_EINVAL, "kx_gather_bytes: row id " + std::to_string(sel[i]) + " outside pack " + std::to_string(packs[p].pack));
    }
    // device layout: views | sel_off | sel | offsets (total + 1, 32-bit) | grand total | bytes
    const size_t off_so = round_up(sizeof(ColView) * size_t(npacks), 256), off_sel = off_so + round_up(8 * (size_t(npacks) + 1), 256);
    const size_t off_len = off_sel + round_up(size_t(total) * 4, 256), off_tot = off_len + round_up((size_t(total) + 1) * 4, 256);
    // (off_tot: 64-bit byte total, + 8: the scan's own 32-bit-carried total, unused); the bytes go to a second scratch buffer
    // that is sized once the total is known
    CK(ctx->d_tmp.reserve(off_tot + 256));
    uint8_t* d = static_cast<uint8_t*>(ctx->d_tmp.p);
    const ColView* dviews = reinterpret_cast<const ColView*>(d);
    const unsigned long long* dso = reinterpret_cast<const unsigned long long*>(d + off_so);
    const uint32_t* dsel = reinterpret_cast<const uint32_t*>(d + off_sel);
    uint32_t* dlen = reinterpret_cast<uint32_t*>(d + off_len);
    CK(cudaMemcpyAsync(d, views.data(), sizeof(ColView) * size_t(npacks), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + off_so, sel_off, 8 * (size_t(npacks) + 1), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + off_sel, sel, size_t(total) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_strgather_len(dviews, dso, uint32_t(npacks), dsel, total, dlen, reinterpret_cast<unsigned long long*>(d + off_tot), ctx->stream));
    CK(launch_exclusive_scan(dlen, uint32_t(total), reinterpret_cast<unsigned long long*>(d + off_tot + 8), ctx->stream));
    unsigned long long need = 0;
    CK(cudaMemcpyAsync(&need, d + off_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> offs32(static_cast<size_t>(total));
    if (need > 0xfffffffeull) return fail(ctx, KX_EUNSUPPORTED, "kx_gather_bytes: more than 4 GiB of row bytes in one call");
    if (need > out_cap || (need && !out)) {   // like kx_scan_select: report the capacity the caller must bring
        out_off[total] = need;
        return fail(ctx, KX_ENOMEM, "kx_gather_bytes: the selected rows hold " + std::to_string(need) + " bytes");
    }
    CK(ctx->d_leafbits.reserve(size_t(need) + 64));
    uint8_t* dout = static_cast<uint8_t*>(ctx->d_leafbits.p);
    CK(launch_strgather_copy(dviews, dso, uint32_t(npacks), dsel, total, dlen, dout, ctx->stream));
    CK(cudaMemcpyAsync(offs32.data(), dlen, size_t(total) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (need) CK(cudaMemcpyAsync(out, dout, size_t(need), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < total; ++i) out_off[i] = offs32[size_t(i)];
    out_off[total] = need;
    return KX_OK;
    });
}

int kx_scan_host(kx_ctx* ctx, const kx_prog* prog, int npacks, const uint16_t* fields, const uint8_t* field_types, int nfields,
                 const void* const* blocks, const size_t* block_len, uint8_t* bitsets, const size_t* bitset_off, int64_t* counts,
                 const kx_agg_req* aggs, int naggs, kx_agg_out* agg_out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog, true);
    if (rc) return rc;
    if (npacks < 0 || nfields < 1 || !fields || !field_types || (npacks && (!blocks || !block_len))) return fail(ctx, KX_EINVAL, "bad block table");
    if (naggs && (!aggs || !agg_out)) return fail(ctx, KX_EINVAL, "aggregate buffers missing");
    CK(cudaSetDevice(ctx->device));
    const int nleaves = int(prog->leaves.size());
    auto field_index = [&](uint16_t f) { for (int i = 0; i < nfields; ++i) if (fields[i] == f) return i; return -1; };
    std::vector<int> leaf_fi(static_cast<size_t>(nleaves)), agg_fi(static_cast<size_t>(naggs));
    for (int l = 0; l < nleaves; ++l) {
        leaf_fi[size_t(l)] = field_index(prog->leaves[size_t(l)].field);
        if (leaf_fi[size_t(l)] < 0) return fail(ctx, KX_EINVAL, "leaf field missing from block table");
        if (field_types[leaf_fi[size_t(l)]] != prog->leaves[size_t(l)].type) return fail(ctx, KX_EINVAL, "leaf / block type mismatch");
    }
    for (int j = 0; j < naggs; ++j) {
        agg_fi[size_t(j)] = field_index(aggs[j].field);
        if (agg_fi[size_t(j)] < 0) return fail(ctx, KX_EINVAL, "aggregate field missing from block table");
        if (field_types[agg_fi[size_t(j)]] != aggs[j].block_type) return fail(ctx, KX_EINVAL, "aggregate / block type mismatch");
    }

    // Batches bounded by encoded bytes so that the transient device footprint stays small.  Two slots: the blocks of
    // batch b + 1 are queued on the copy stream into the other staging buffer before batch b is scanned.  (Measured on
    // B200 with 128 MB batches: no gain over one large batch — the scan's small descriptor upload queues behind the next
    // batch's bulk copies on the same H2D copy engine — so batches stay large: one batch for anything below 2 GB.)
    const size_t BATCH_BYTES = 2048ull << 20;
    std::vector<kx_agg_out> part(static_cast<size_t>(naggs));
    std::vector<std::vector<kx_agg_out>> parts(static_cast<size_t>(naggs));
    double kms = 0, tms = 0; int launches = 0;
    uint64_t q_rows = 0, q_packs = 0, q_match = 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0, ctx->stream));
    CK(cudaStreamWaitEvent(ctx->copy_stream, e0, 0));   // uploads do not overtake earlier work on the context's stream

    struct HostBatch { int p0 = 0, p1 = 0; std::vector<BlockLayout> lays; std::vector<std::vector<uint8_t>> cstr; /* values of constant string blocks */ };
    HostBatch slots[2];
    DevBuf* dstage[2] = {&ctx->d_stage, &ctx->d_stage2};
    PinBuf* haux[2] = {&ctx->h_aux, &ctx->h_aux2};

    // parse the blocks of packs [p0, p1), lay them out in slot `sl`'s staging buffer and queue the uploads
    auto prepare = [&](int sl, int p0) -> int {
        HostBatch& B = slots[sl];
        size_t bytes = 0; int p1 = p0;
        while (p1 < npacks) {
            size_t b = 0;
            for (int f = 0; f < nfields; ++f) b += block_len[size_t(p1) * nfields + f];
            if (p1 > p0 && bytes + b > BATCH_BYTES) break;
            bytes += b; ++p1;
        }
        B.p0 = p0; B.p1 = p1;
        const int nb = p1 - p0;
        // pass 1: parse headers, lay the batch out in the staging arena (256 B aligned streams)
        B.lays.clear();
        B.lays.resize(size_t(nb) * nfields);
        B.cstr.assign(size_t(nb) * nfields, {});
        std::vector<BlockLayout>& lays = B.lays;
        std::vector<size_t> off_stream(lays.size(), 0), off_a64(lays.size(), 0), off_a32(lays.size(), 0), off_blob(lays.size(), 0), aux_src(lays.size(), 0);
        size_t dev_bytes = 0, aux_bytes = 0;
        for (int p = 0; p < nb; ++p) {
            for (int f = 0; f < nfields; ++f) {
                size_t bi = size_t(p0 + p) * nfields + f, li = size_t(p) * nfields + f;
                std::string err;
                int rc2;
                if (field_types[f] == KX_BYTES) {
                    // byte-string block: the byte buffer travels verbatim like a packed stream, the flat index array like run ends
                    StrLayout sl2;
                    rc2 = normalize_string_block(static_cast<const uint8_t*>(blocks[bi]), block_len[bi], sl2, err);
                    if (!rc2) {
                        lays[li].view = sl2.view; lays[li].stream = sl2.bytes; lays[li].stream_len = sl2.nbytes; lays[li].aux32 = std::move(sl2.idx);
                        if (sl2.view.is_raw == STR_CONST) B.cstr[li].assign(sl2.bytes, sl2.bytes + sl2.nbytes);
                    }
                } else {
                    rc2 = normalize_block(field_types[f], static_cast<const uint8_t*>(blocks[bi]), block_len[bi], lays[li], err);
                }
                if (rc2) return fail(ctx, rc2, "kx_scan_host: " + err);
                BlockLayout& lay = lays[li];
                if (lay.owned.empty() && lay.stream_len) { off_stream[li] = dev_bytes; dev_bytes += round_up(lay.stream_len + STREAM_PAD, 256); }
                aux_src[li] = aux_bytes;
                aux_bytes += round_up(lay.owned.size(), 16) + round_up(lay.aux64.size() * 8, 16) + round_up(lay.aux32.size() * 4, 16) +
                             round_up(lay.blob.size(), 16);
            }
        }
        // host-built arrays (dictionaries, run values/ends, transcoded streams) go through one pinned buffer
        for (size_t li = 0; li < lays.size(); ++li) {
            BlockLayout& lay = lays[li];
            if (!lay.owned.empty()) { off_stream[li] = dev_bytes; dev_bytes += round_up(lay.owned.size() + STREAM_PAD, 256); }
            if (!lay.aux64.empty()) { off_a64[li] = dev_bytes; dev_bytes += round_up(lay.aux64.size() * 8 + STREAM_PAD, 256); }
            if (!lay.aux32.empty()) { off_a32[li] = dev_bytes; dev_bytes += round_up(lay.aux32.size() * 4 + STREAM_PAD, 256); }
            if (!lay.blob.empty()) { off_blob[li] = dev_bytes; dev_bytes += round_up(lay.blob.size() + STREAM_PAD, 256); }
        }
        CK(dstage[sl]->reserve(dev_bytes + 256));
        CK(haux[sl]->reserve(aux_bytes + 64));
        uint8_t* dbase = static_cast<uint8_t*>(dstage[sl]->p);
        uint8_t* hbase = static_cast<uint8_t*>(haux[sl]->p);
        cudaStream_t cs = ctx->copy_stream;
        // pass 2: queue the copies (asynchronous when the caller's blocks are pinned)
        for (size_t li = 0; li < lays.size(); ++li) {
            BlockLayout& lay = lays[li];
            ColView& v = lay.view;
            uint8_t* hp = hbase + aux_src[li];
            if (lay.owned.empty()) {
                if (lay.stream_len) CK(cudaMemcpyAsync(dbase + off_stream[li], lay.stream, lay.stream_len, cudaMemcpyHostToDevice, cs));
            } else {
                std::memcpy(hp, lay.owned.data(), lay.owned.size());
                CK(cudaMemcpyAsync(dbase + off_stream[li], hp, lay.owned.size(), cudaMemcpyHostToDevice, cs));
                hp += round_up(lay.owned.size(), 16);
            }
            if (!lay.aux64.empty()) {
                std::memcpy(hp, lay.aux64.data(), lay.aux64.size() * 8);
                CK(cudaMemcpyAsync(dbase + off_a64[li], hp, lay.aux64.size() * 8, cudaMemcpyHostToDevice, cs));
                hp += round_up(lay.aux64.size() * 8, 16);
            }
            if (!lay.aux32.empty()) {
                std::memcpy(hp, lay.aux32.data(), lay.aux32.size() * 4);
                CK(cudaMemcpyAsync(dbase + off_a32[li], hp, lay.aux32.size() * 4, cudaMemcpyHostToDevice, cs));
                hp += round_up(lay.aux32.size() * 4, 16);
            }
            if (!lay.blob.empty()) {
                std::memcpy(hp, lay.blob.data(), lay.blob.size());
                CK(cudaMemcpyAsync(dbase + off_blob[li], hp, lay.blob.size(), cudaMemcpyHostToDevice, cs));
            }
            if (v.kind == CK_BITS || v.kind == CK_DICT || ((v.kind == CK_ALP || v.kind == CK_ALPRD) && v.width)) v.data = dbase + off_stream[li];
            if ((v.kind == CK_ALP || v.kind == CK_ALPRD) && !lay.blob.empty()) v.aux = dbase + off_blob[li];
            if (v.kind == CK_DICT) v.aux = dbase + off_a64[li];
            if (v.kind == CK_RUNEND) { v.data = dbase + off_a64[li]; v.aux = dbase + off_a32[li]; }
            if (v.kind == CK_STR) { v.data = dbase + off_stream[li]; v.aux = dbase + off_a32[li]; }
        }
        CK(cudaEventRecord(ctx->ev_copy[sl], cs));
        return KX_OK;
    };

    // scan the batch of slot `sl` once its uploads have landed
    auto scan_slot = [&](int sl) -> int {
        HostBatch& B = slots[sl];
        const int p0 = B.p0, nb = B.p1 - B.p0;
        std::vector<BlockLayout>& lays = B.lays;
        ScanJob job; job.npacks = nb;
        job.nrows.resize(size_t(nb)); job.leaf_views.resize(size_t(nb) * nleaves); job.leaf_dicts.resize(size_t(nb) * nleaves);
        job.agg_views.resize(size_t(nb) * size_t(naggs));
        job.leaf_cstr.assign(size_t(nb) * nleaves, nullptr);
        for (int p = 0; p < nb; ++p) {
            job.nrows[size_t(p)] = lays[size_t(p) * nfields].view.n;
            for (int l = 0; l < nleaves; ++l) {
                const BlockLayout& lay = lays[size_t(p) * nfields + leaf_fi[size_t(l)]];
                job.leaf_views[size_t(p) * nleaves + l] = lay.view;
                if (lay.view.kind == CK_STR && lay.view.is_raw == STR_CONST) job.leaf_cstr[size_t(p) * nleaves + l] = &B.cstr[size_t(p) * nfields + leaf_fi[size_t(l)]];
                job.leaf_dicts[size_t(p) * nleaves + l] = (lay.view.kind == CK_DICT && !lay.aux64.empty()) ? lay.aux64.data() : nullptr;
            }
            for (int j = 0; j < naggs; ++j) job.agg_views[size_t(p) * naggs + j] = lays[size_t(p) * nfields + agg_fi[size_t(j)]].view;
        }
        // bitset offsets of this batch are relative to the caller's buffer start; shift so the
        // device buffer only spans the batch
        std::vector<size_t> offs;
        uint8_t* bdst = nullptr;
        if (bitsets) {
            size_t base = bitset_off[p0];
            for (int p = 0; p < nb; ++p) {
                if (bitset_off[p0 + p] < base) return fail(ctx, KX_EINVAL, "bitset_off must be ascending");
                offs.push_back(bitset_off[p0 + p] - base);
            }
            if (base & 7) return fail(ctx, KX_EINVAL, "bitset_off must be a multiple of 8");
            bdst = bitsets + base;
        }
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[sl], 0));
        int rc2 = run_scan(ctx, prog, job, bdst, bitsets ? offs.data() : nullptr, counts ? counts + p0 : nullptr, aggs, naggs,
                           naggs ? part.data() : nullptr);
        kms += ctx->last_kernel_ms; launches += ctx->last_launches;
        q_rows += ctx->last_rows_scanned; q_packs += ctx->last_packs_scanned; q_match += ctx->last_rows_matched;
        if (rc2) return rc2;
        for (int j = 0; j < naggs; ++j) parts[size_t(j)].push_back(part[size_t(j)]);
        return KX_OK;
    };

    auto drain = [&](int rc2) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->stream); cudaEventDestroy(e0); cudaEventDestroy(e1); return rc2; };
    if (npacks > 0) {
        int sl = 0;
        if ((rc = prepare(0, 0))) return drain(rc);
        for (;;) {
            const int next_p0 = slots[sl].p1;
            if (next_p0 < npacks && (rc = prepare(sl ^ 1, next_p0))) return drain(rc);   // uploads of the next batch start now
            if ((rc = scan_slot(sl))) return drain(rc);
            if (next_p0 >= npacks) break;
            sl ^= 1;
        }
    }
    CK(cudaEventRecord(e1, ctx->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1)); tms = ms;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (int j = 0; j < naggs; ++j) {
        if (parts[size_t(j)].empty()) agg_out[j] = kx_agg_out{};
        else kx_agg_combine(aggs[j].block_type, parts[size_t(j)].data(), int(parts[size_t(j)].size()), &agg_out[j]);
    }
    ctx->last_kernel_ms = kms; ctx->last_total_ms = tms; ctx->last_launches = launches;
    ctx->last_rows_scanned = q_rows; ctx->last_packs_scanned = q_packs; ctx->last_rows_matched = q_match;
    return KX_OK;
    });
}

int kx_agg_combine(uint8_t block_type, const kx_agg_out* parts, int nparts, kx_agg_out* out) {
    if (!out || nparts < 0 || (nparts && !parts)) return KX_EINVAL;
    int t = block_type;
    if (type_bits(t) == 0) return KX_EINVAL;
    kx_agg_out r{};
    double fs = 0, fe = 0;
    for (int i = 0; i < nparts; ++i) {
        const kx_agg_out& p = parts[i];
        if (!p.valid) continue;
        double ps = 0; std::memcpy(&ps, &p.sum_bits, 8);
        if (t == KX_FLOAT32) { uint32_t u = uint32_t(p.sum_bits); float f; std::memcpy(&f, &u, 4); ps = double(f); }   // + sum_err = the float64 partial sum
        if (!r.valid) {
            r = p; fs = ps; fe = p.sum_err;
            continue;
        }
        r.count += p.count;
        if (t == KX_FLOAT64) {
            double tt = fs + ps;
            double c = (std::abs(fs) >= std::abs(ps)) ? ((fs - tt) + ps) : ((ps - tt) + fs);
            fs = tt; fe += p.sum_err + c;
            double a, b; std::memcpy(&a, &r.min_bits, 8); std::memcpy(&b, &p.min_bits, 8); if (b < a) r.min_bits = p.min_bits;
            std::memcpy(&a, &r.max_bits, 8); std::memcpy(&b, &p.max_bits, 8); if (b > a) r.max_bits = p.max_bits;
        } else if (t == KX_FLOAT32) {
            double tt = fs + ps;
            double c = (std::abs(fs) >= std::abs(ps)) ? ((fs - tt) + ps) : ((ps - tt) + fs);
            fs = tt; fe += p.sum_err + c;
            auto f32 = [](uint64_t pat) { uint32_t u = uint32_t(pat); float f; std::memcpy(&f, &u, 4); return f; };
            if (f32(p.min_bits) < f32(r.min_bits)) r.min_bits = p.min_bits;
            if (f32(p.max_bits) > f32(r.max_bits)) r.max_bits = p.max_bits;
        } else {
            r.sum_bits = type_ext(t, r.sum_bits + p.sum_bits);
            bool sg = type_is_signed(t);
            auto lt = [&](uint64_t x, uint64_t y) { return sg ? int64_t(x) < int64_t(y) : x < y; };
            if (lt(p.min_bits, r.min_bits)) r.min_bits = p.min_bits;
            if (lt(r.max_bits, p.max_bits)) r.max_bits = p.max_bits;
        }
    }
    if (r.valid && t == KX_FLOAT64) {
        double s = fs + fe;
        std::memcpy(&r.sum_bits, &s, 8);
        r.sum_err = (fs - s) + fe;
    }
    if (r.valid && t == KX_FLOAT32) {
        const double s = fs + fe;
        const float f = float(s); uint32_t u; std::memcpy(&u, &f, 4);
        r.sum_bits = u; r.sum_err = s - double(f);
    }
    *out = r;
    return KX_OK;
}

int kx_last_scan_stats(kx_ctx* ctx, double* kernel_ms, double* total_ms, int* launches) {
    if (!ctx) return KX_EINVAL;
    if (kernel_ms) *kernel_ms = ctx->last_kernel_ms;
    if (total_ms) *total_ms = ctx->last_total_ms;
    if (launches) *launches = ctx->last_launches;
    return KX_OK;
}

int kx_debug_check_guards(kx_ctx* ctx) {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!guard_mode()) return KX_EUNSUPPORTED;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    int bad = 0;
    std::lock_guard<std::mutex> lk2(guarded_mu());
    for (DevBuf* b : guarded_bufs()) if (!b->guard_intact()) ++bad;
    return bad;
}

int kx_last_query_stats(kx_ctx* ctx, kx_query_stats* out) {
    if (!ctx || !out) return KX_EINVAL;
    out->rows_scanned = ctx->last_rows_scanned; out->packs_scanned = ctx->last_packs_scanned; out->rows_matched = ctx->last_rows_matched;
    out->scan_time_ns = uint64_t(ctx->last_kernel_ms * 1e6); out->total_time_ns = uint64_t(ctx->last_total_ms * 1e6);
    out->kernel_launches = uint32_t(ctx->last_launches); out->reserved = 0;
    return KX_OK;
}

// ------------------------------------------------------------------ narrow drop-ins
int64_t kx_cmp(kx_ctx* ctx, uint8_t block_type, uint8_t mode, const void* src, size_t n, uint64_t a, uint64_t b, uint8_t* bits) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int w = type_bits(block_type);
    if (!w) return fail(ctx, KX_EINVAL, "kx_cmp: unsupported type");
    if (mode == KX_MODE_IN || mode == KX_MODE_NIN || mode < KX_MODE_EQ || mode > KX_MODE_RANGE) return fail(ctx, KX_EINVAL, "kx_cmp: unsupported mode");
    if (n == 0) return 0;
    if (!src || !bits || n > 0xffffffffull) return fail(ctx, KX_EINVAL, "kx_cmp: bad arguments");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    BlockLayout lay;
    lay.view.kind = CK_BITS; lay.view.type = block_type; lay.view.width = uint8_t(w); lay.view.is_raw = 1; lay.view.n = uint32_t(n);
    lay.stream = static_cast<const uint8_t*>(src); lay.stream_len = n * size_t(w / 8);
    TempBlock tb(ctx);
    int rc = upload_block(ctx, lay, tb.sb);
    if (rc) return rc;
    int64_t count = 0;
    rc = single_leaf_scan(ctx, tb.sb, block_type, mode, a, b, nullptr, 0, bits, &count);
    return rc ? rc : count;
    });
}

int64_t kx_bitpack_cmp(kx_ctx* ctx, uint8_t mode, const void* packed, int log2, uint64_t a, uint64_t b, size_t n, uint8_t* bits) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (log2 < 0 || log2 > 64) return fail(ctx, KX_EINVAL, "kx_bitpack_cmp: bad width");
    if (mode == KX_MODE_IN || mode == KX_MODE_NIN || mode < KX_MODE_EQ || mode > KX_MODE_RANGE) return fail(ctx, KX_EINVAL, "kx_bitpack_cmp: unsupported mode");
    if (n == 0) return 0;
    if (!bits || (log2 && !packed) || n > 0xffffffffull) return fail(ctx, KX_EINVAL, "kx_bitpack_cmp: bad arguments");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    BlockLayout lay;
    lay.view.kind = log2 ? CK_BITS : CK_CONST; lay.view.type = KX_UINT64; lay.view.width = uint8_t(log2); lay.view.is_raw = 0;
    lay.view.n = uint32_t(n); lay.view.base = 0;
    lay.stream = static_cast<const uint8_t*>(packed); lay.stream_len = bitpack_bytes(log2, n);
    TempBlock tb(ctx);
    int rc = upload_block(ctx, lay, tb.sb);
    if (rc) return rc;
    int64_t count = 0;
    rc = single_leaf_scan(ctx, tb.sb, KX_UINT64, mode, a, b, nullptr, 0, bits, &count);
    return rc ? rc : count;
    });
}

static int decode_view_to_host(kx_ctx* ctx, const ColView& v, void* dst) {
    size_t bytes = size_t(v.n) * size_t(type_bits(v.type) / 8);
    if (!bytes) return KX_OK;
    CK(ctx->d_tmp.reserve(bytes));
    CK(launch_decode(v, ctx->d_tmp.p, ctx->stream));
    CK(cudaMemcpyAsync(dst, ctx->d_tmp.p, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KX_OK;
}

int kx_bitpack_decode(kx_ctx* ctx, uint8_t block_type, const void* packed, int log2, uint64_t minv, size_t n, void* dst) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (log2 < 0 || log2 > 64 || type_bits(block_type) == 0 || type_is_float(block_type)) return fail(ctx, KX_EINVAL, "kx_bitpack_decode: bad arguments");
    if (n == 0) return KX_OK;
    if (!dst || (log2 && !packed) || n > 0xffffffffull) return fail(ctx, KX_EINVAL, "kx_bitpack_decode: bad arguments");
    CK(cudaSetDevice(ctx->device));
    BlockLayout lay;
    lay.view.kind = log2 ? CK_BITS : CK_CONST; lay.view.type = block_type; lay.view.width = uint8_t(log2); lay.view.n = uint32_t(n);
    lay.view.base = type_ext(block_type, minv);
    lay.stream = static_cast<const uint8_t*>(packed); lay.stream_len = bitpack_bytes(log2, n);
    TempBlock tb(ctx);
    int rc = upload_block(ctx, lay, tb.sb);
    if (rc) return rc;
    return decode_view_to_host(ctx, tb.sb.view, dst);
    });
}

int64_t kx_container_match(kx_ctx* ctx, uint8_t block_type, const void* enc, size_t len, uint8_t mode, uint64_t a, uint64_t b,
                           const uint64_t* set, uint32_t nset, uint8_t* bits) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!enc || !len) return fail(ctx, KX_EINVAL, "kx_container_match: empty block");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    BlockLayout lay; std::string err;
    int rc = normalize_block(block_type, static_cast<const uint8_t*>(enc), len, lay, err);
    if (rc) return fail(ctx, rc, "kx_container_match: " + err);
    if (lay.view.n == 0) return 0;
    if (!bits) return fail(ctx, KX_EINVAL, "kx_container_match: bits missing");
    TempBlock tb(ctx);
    rc = upload_block(ctx, lay, tb.sb);
    if (rc) return rc;
    int64_t count = 0;
    rc = single_leaf_scan(ctx, tb.sb, block_type, mode, a, b, set, nset, bits, &count);
    return rc ? rc : count;
    });
}

int kx_container_decode(kx_ctx* ctx, uint8_t block_type, const void* enc, size_t len, void* dst, size_t dst_cap_rows) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!enc || !len) return fail(ctx, KX_EINVAL, "kx_container_decode: empty block");
    CK(cudaSetDevice(ctx->device));
    BlockLayout lay; std::string err;
    int rc = normalize_block(block_type, static_cast<const uint8_t*>(enc), len, lay, err);
    if (rc) return fail(ctx, rc, "kx_container_decode: " + err);
    if (lay.view.n > dst_cap_rows) return fail(ctx, KX_EINVAL, "kx_container_decode: destination too small");
    if (lay.view.n && !dst) return fail(ctx, KX_EINVAL, "kx_container_decode: dst missing");
    TempBlock tb(ctx);
    rc = upload_block(ctx, lay, tb.sb);
    if (rc) return rc;
    return decode_view_to_host(ctx, tb.sb.view, dst);
    });
}

// ------------------------------------------------------------------ bitsets
static int upload_bits(kx_ctx* ctx, DevBuf& buf, const uint8_t* src, size_t nbits) {
    size_t nbytes = (nbits + 7) / 8, padded = round_up(nbytes, 4) + 8;
    CK(buf.reserve(padded));
    CK(cudaMemsetAsync(static_cast<uint8_t*>(buf.p) + (nbytes & ~size_t(3)), 0, padded - (nbytes & ~size_t(3)), ctx->stream));
    CK(cudaMemcpyAsync(buf.p, src, nbytes, cudaMemcpyHostToDevice, ctx->stream));
    return KX_OK;
}

int kx_bitset_op(kx_ctx* ctx, int op, uint8_t* dst, const uint8_t* src, size_t nbits, int* any, int* all) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (op < KX_BIT_AND || op > KX_BIT_XOR) return fail(ctx, KX_EINVAL, "kx_bitset_op: bad op");
    if (nbits == 0) { if (any) *any = 0; if (all) *all = 1; return KX_OK; }
    if (!dst || !src) return fail(ctx, KX_EINVAL, "kx_bitset_op: null bitset");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = upload_bits(ctx, ctx->d_tmp, dst, nbits)) || (rc = upload_bits(ctx, ctx->d_tmp2, src, nbits))) return rc;
    CK(ctx->d_misc.reserve(64));
    CK(cudaMemsetAsync(ctx->d_misc.p, 0, 8, ctx->stream));
    CK(launch_bitset_op(static_cast<uint32_t*>(ctx->d_tmp.p), static_cast<const uint32_t*>(ctx->d_tmp2.p), nbits, op,
                        static_cast<unsigned int*>(ctx->d_misc.p), ctx->stream));
    unsigned int flags[2] = {0, 0};
    CK(cudaMemcpyAsync(dst, ctx->d_tmp.p, (nbits + 7) / 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(flags, ctx->d_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (any) *any = flags[0] != 0;
    if (all) *all = flags[1] == 0;
    return KX_OK;
    });
}

int kx_bitset_neg(kx_ctx* ctx, uint8_t* buf, size_t nbits) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nbits == 0) return KX_OK;
    if (!buf) return fail(ctx, KX_EINVAL, "kx_bitset_neg: null bitset");
    CK(cudaSetDevice(ctx->device));
    int rc = upload_bits(ctx, ctx->d_tmp, buf, nbits);
    if (rc) return rc;
    CK(launch_bitset_neg(static_cast<uint32_t*>(ctx->d_tmp.p), nbits, ctx->stream));
    CK(cudaMemcpyAsync(buf, ctx->d_tmp.p, (nbits + 7) / 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KX_OK;
    });
}

int64_t kx_bitset_popcount(kx_ctx* ctx, const uint8_t* buf, size_t nbits) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nbits == 0) return 0;
    if (!buf) return fail(ctx, KX_EINVAL, "kx_bitset_popcount: null bitset");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    int rc = upload_bits(ctx, ctx->d_tmp, buf, nbits);
    if (rc) return rc;
    CK(ctx->d_misc.reserve(64));
    CK(cudaMemsetAsync(ctx->d_misc.p, 0, 8, ctx->stream));
    CK(launch_bitset_popcount(static_cast<const uint32_t*>(ctx->d_tmp.p), nbits, static_cast<unsigned long long*>(ctx->d_misc.p), ctx->stream));
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, ctx->d_misc.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return int64_t(c);
    });
}

int64_t kx_bitset_indexes(kx_ctx* ctx, const uint8_t* buf, size_t nbits, uint32_t* dst) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (nbits == 0) return 0;
    if (!buf || !dst || nbits > 0xffffffffull) return fail(ctx, KX_EINVAL, "kx_bitset_indexes: bad arguments");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    int rc = upload_bits(ctx, ctx->d_tmp, buf, nbits);
    if (rc) return rc;
    size_t nwords = (nbits + 31) / 32, nblocks = (nwords + 255) / 256;
    CK(ctx->d_misc.reserve(64 + nblocks * 4));
    CK(ctx->d_tmp2.reserve(nbits * 4 + 64));
    unsigned long long* total = static_cast<unsigned long long*>(ctx->d_misc.p);
    uint32_t* block_tmp = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(ctx->d_misc.p) + 64);
    CK(launch_bitset_indexes(static_cast<const uint32_t*>(ctx->d_tmp.p), nbits, block_tmp, total, static_cast<uint32_t*>(ctx->d_tmp2.p), ctx->stream));
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(&c, total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (c) {
        CK(cudaMemcpyAsync(dst, ctx->d_tmp2.p, size_t(c) * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return int64_t(c);
    });
}

// ------------------------------------------------------------------ pruning
int64_t kx_prune(kx_ctx* ctx, const kx_prog* prog, int npacks, const uint64_t* mins, const uint64_t* maxs, const void* const* blooms,
                 const size_t* bloom_len, const uint64_t* hashes, const uint32_t* hash_off, uint8_t* out) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog);
    if (rc) return rc;
    if (npacks == 0) return 0;
    if (npacks < 0 || !mins || !maxs || !out) return fail(ctx, KX_EINVAL, "kx_prune: bad arguments");
    if (blooms && (!bloom_len || !hashes || !hash_off)) return fail(ctx, KX_EINVAL, "kx_prune: bloom tables incomplete");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    const int nleaves = int(prog->leaves.size());
    const size_t cells = size_t(npacks) * nleaves;

    PruneParams P{};
    bool any_f32 = false;
    P.npacks = uint32_t(npacks); P.nleaves = uint32_t(nleaves); P.npost = uint32_t(prog->postfix.size());
    std::memcpy(P.postfix, prog->postfix.data(), prog->postfix.size());
    for (int l = 0; l < nleaves; ++l) {
        const LeafSpec& s = prog->leaves[size_t(l)];
        PruneLeaf& pl = P.leaves[l];
        pl.a = type_is_float(s.type) ? s.a : type_ext(s.type, s.a);
        pl.b = type_is_float(s.type) ? s.b : type_ext(s.type, s.b);
        pl.flip = type_is_signed(s.type) ? 0x8000000000000000ull : 0;
        pl.mode = s.mode; pl.is_float = type_is_float(s.type);
        pl.set_off = s.set_off; pl.nset = uint32_t(s.set.size());
        if (s.type == KX_FLOAT32) { pl.a = f32_to_f64_bits(s.a); pl.b = f32_to_f64_bits(s.b); any_f32 = true; }
        if (s.type == KX_BYTES) return fail(ctx, KX_EUNSUPPORTED, "kx_prune: byte-string columns need the resident index (kx_prune_stats)");
    }
    std::vector<uint64_t> wmins, wmaxs;   // float32 columns: zone maps widened to float64 patterns
    if (any_f32) {
        wmins.assign(mins, mins + cells); wmaxs.assign(maxs, maxs + cells);
        for (int l = 0; l < nleaves; ++l) {
            if (prog->leaves[size_t(l)].type != KX_FLOAT32) continue;
            for (int p = 0; p < npacks; ++p) { size_t c = size_t(p) * nleaves + l; wmins[c] = f32_to_f64_bits(wmins[c]); wmaxs[c] = f32_to_f64_bits(wmaxs[c]); }
        }
        mins = wmins.data(); maxs = wmaxs.data();
    }
    size_t nhash = 0;
    if (blooms) { for (int l = 0; l <= nleaves; ++l) P.hash_off[l] = hash_off[l]; nhash = hash_off[nleaves]; }

    // device layout: mins | maxs | bloom ptr table | bloom lens | hashes | out words | count | bloom payloads
    size_t off_max = cells * 8, off_bp = off_max + cells * 8, off_bl = off_bp + (blooms ? cells * 8 : 0);
    size_t off_h = off_bl + (blooms ? cells * 8 : 0), off_out = round_up(off_h + nhash * 8, 8);
    size_t out_words = (size_t(npacks) + 31) / 32, off_cnt = round_up(off_out + out_words * 4, 8), off_pay = round_up(off_cnt + 8, 256);
    size_t pay = 0;
    std::vector<size_t> pay_off(blooms ? cells : 0);
    if (blooms) for (size_t i = 0; i < cells; ++i) if (blooms[i] && bloom_len[i] > 1) { pay_off[i] = pay; pay += round_up(bloom_len[i], 32); }
    CK(ctx->d_tmp.reserve(off_pay + pay + 64));
    uint8_t* d = static_cast<uint8_t*>(ctx->d_tmp.p);
    CK(cudaMemcpyAsync(d, mins, cells * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(d + off_max, maxs, cells * 8, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint64_t> ptrs, lens;
    if (blooms) {
        ptrs.resize(cells); lens.resize(cells);
        for (size_t i = 0; i < cells; ++i) {
            bool has = blooms[i] && bloom_len[i] > 1;
            // the buffer must hold k + a power-of-two number of bits (bloom.go:83-100); others are ignored
            if (has) { size_t m = (bloom_len[i] - 1) * 8; if (m & (m - 1)) has = false; }
            ptrs[i] = has ? uint64_t(reinterpret_cast<uintptr_t>(d + off_pay + pay_off[i])) : 0;
            lens[i] = has ? bloom_len[i] : 0;
            if (has) CK(cudaMemcpyAsync(d + off_pay + pay_off[i], blooms[i], bloom_len[i], cudaMemcpyHostToDevice, ctx->stream));
        }
        CK(cudaMemcpyAsync(d + off_bp, ptrs.data(), cells * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(d + off_bl, lens.data(), cells * 8, cudaMemcpyHostToDevice, ctx->stream));
        if (nhash) CK(cudaMemcpyAsync(d + off_h, hashes, nhash * 8, cudaMemcpyHostToDevice, ctx->stream));
    }
    CK(cudaMemsetAsync(d + off_out, 0, off_pay - off_out, ctx->stream));
    P.mins = reinterpret_cast<const uint64_t*>(d); P.maxs = reinterpret_cast<const uint64_t*>(d + off_max);
    P.blooms = blooms ? reinterpret_cast<const uint8_t* const*>(d + off_bp) : nullptr;
    P.bloom_len = reinterpret_cast<const uint64_t*>(d + off_bl);
    P.hashes = reinterpret_cast<const uint64_t*>(d + off_h);
    P.set_vals = prog->dev_sets;
    P.out = reinterpret_cast<uint32_t*>(d + off_out);
    P.count = reinterpret_cast<unsigned long long*>(d + off_cnt);
    CK(launch_prune(P, ctx->stream));
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(out, d + off_out, (size_t(npacks) + 7) / 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&c, d + off_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return int64_t(c);
    });
}

// ------------------------------------------------------------------ resident statistics index
static int bloom_pow2(size_t v) {   // bloom.pow2 (internal/filter/bloom/bloom.go:242-250)
    for (size_t i = 8; i < (size_t(1) << 30); i *= 2) if (i >= v) return int(i);
    return 0;
}

int kx_stats_create(kx_ctx* ctx, int npacks, const uint16_t* fields, const uint8_t* field_types, int nfields,
                    const uint64_t* mins, const uint64_t* maxs, kx_stats** out) {
    return kx_guarded<int>(ctx, [&]() -> int {
    if (!ctx || !out) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    *out = nullptr;
    if (npacks < 1 || nfields < 1 || nfields > 255 || !fields || !field_types || !mins || !maxs) return fail(ctx, KX_EINVAL, "kx_stats_create: bad arguments");
    for (int f = 0; f < nfields; ++f)
        if (type_bits(field_types[f]) == 0 && field_types[f] != KX_BYTES) return fail(ctx, KX_EINVAL, "kx_stats_create: unsupported field type");
    CK(cudaSetDevice(ctx->device));
    auto st = std::make_unique<kx_stats>();
    st->ctx = ctx; st->npacks = npacks; st->nfields = nfields;
    st->fields.assign(fields, fields + nfields); st->types.assign(field_types, field_types + nfields);
    const size_t cells = size_t(npacks) * nfields;
    st->h_bloom_ptr.assign(cells, 0); st->h_bloom_mask.assign(cells, 0); st->h_bloom_k.assign(cells, 0); st->cell_alloc.assign(cells, SlabAlloc{});
    CK(cudaMalloc(&st->d_mins, cells * 8));
    CK(cudaMalloc(&st->d_maxs, cells * 8));
    CK(cudaMalloc(&st->d_tab, cells * 13 + 64));
    std::vector<uint64_t> wmins(mins, mins + cells), wmaxs(maxs, maxs + cells);
    for (int f = 0; f < nfields; ++f)   // float32 columns: zone maps widened to float64 patterns (the prune kernels compare float64)
        if (field_types[f] == KX_FLOAT32)
            for (int p = 0; p < npacks; ++p) { size_t c = size_t(f) * npacks + p; wmins[c] = f32_to_f64_bits(wmins[c]); wmaxs[c] = f32_to_f64_bits(wmaxs[c]); }
    CK(cudaMemcpyAsync(st->d_mins, wmins.data(), cells * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(st->d_maxs, wmaxs.data(), cells * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *out = st.release();
    return KX_OK;
    });
}

void kx_stats_free(kx_stats* st) {
    if (!st) return;
    kx_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& a : st->cell_alloc) slab_release(ctx, a);
    delete st;
}

// allocate (or reuse) the bit array of one cell; m bits
static int stats_bloom_slot(kx_stats* st, size_t cell, size_t m, uint32_t k, uint8_t** bits) {
    kx_ctx* ctx = st->ctx;
    if (st->h_bloom_ptr[cell] && st->h_bloom_mask[cell] == uint32_t(m - 1)) {
        *bits = reinterpret_cast<uint8_t*>(uintptr_t(st->h_bloom_ptr[cell]));
    } else {
        // a filter of another size replaces the old one: allocate first, then give the old extent back
        uint8_t* d = nullptr; SlabAlloc rec;
        int rc = slab_alloc(ctx, m / 8 + 64, &d, &rec);
        if (rc) return rc;
        if (st->h_bloom_ptr[cell]) CK(cudaStreamSynchronize(ctx->stream));   // no probe kernel may still read the old bits
        slab_release(ctx, st->cell_alloc[cell]);
        st->cell_alloc[cell] = rec;
        *bits = d;
    }
    st->h_bloom_ptr[cell] = uint64_t(uintptr_t(*bits)); st->h_bloom_mask[cell] = uint32_t(m - 1); st->h_bloom_k[cell] = uint8_t(k);
    st->dirty = true;
    return KX_OK;
}

int kx_stats_put_bloom(kx_stats* st, int field_index, int pack_index, const void* bloom, size_t len) {
    return kx_guarded<int>(st ? st->ctx : nullptr, [&]() -> int {
    if (!st) return KX_EINVAL;
    kx_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (field_index < 0 || field_index >= st->nfields || pack_index < 0 || pack_index >= st->npacks || !bloom || len < 2)
        return fail(ctx, KX_EINVAL, "kx_stats_put_bloom: bad arguments");
    size_t m = (len - 1) * 8;
    if ((m & (m - 1)) || m < 8) return fail(ctx, KX_EFORMAT, "kx_stats_put_bloom: bit count must be a power of two (bloom.NewFilterBuffer)");
    CK(cudaSetDevice(ctx->device));
    const uint8_t* b = static_cast<const uint8_t*>(bloom);
    uint8_t* bits = nullptr;
    int rc = stats_bloom_slot(st, size_t(field_index) * st->npacks + pack_index, m, b[0], &bits);
    if (rc) return rc;
    CK(cudaMemcpyAsync(bits, b + 1, len - 1, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // the caller's buffer is not retained
    return KX_OK;
    });
}

int kx_stats_build_bloom(kx_stats* st, int field_index, int pack_index, uint8_t block_type, const void* values, const uint32_t* offsets,
                         size_t n, int cardinality, int factor) {
    return kx_guarded<int>(st ? st->ctx : nullptr, [&]() -> int {
    if (!st) return KX_EINVAL;
    kx_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (field_index < 0 || field_index >= st->nfields || pack_index < 0 || pack_index >= st->npacks || (n && !values))
        return fail(ctx, KX_EINVAL, "kx_stats_build_bloom: bad arguments");
    int eb = block_type == KX_BYTES ? 0 : type_bits(block_type) / 8;
    if (block_type != KX_BYTES && eb == 0) return fail(ctx, KX_EINVAL, "kx_stats_build_bloom: unsupported value type");
    if (block_type == KX_BYTES && n && !offsets) return fail(ctx, KX_EINVAL, "kx_stats_build_bloom: byte strings need offsets");
    if (cardinality <= 0 || factor <= 0) return fail(ctx, KX_EINVAL, "kx_stats_build_bloom: cardinality and factor must be positive");   // BuildBloomFilter returns nil
    size_t m = size_t(bloom_pow2(size_t(cardinality) * size_t(factor) * 8));
    if (!m) return fail(ctx, KX_EINVAL, "kx_stats_build_bloom: filter too large");
    CK(cudaSetDevice(ctx->device));
    uint8_t* bits = nullptr;
    int rc = stats_bloom_slot(st, size_t(field_index) * st->npacks + pack_index, m, 4, &bits);
    if (rc) return rc;
    CK(cudaMemsetAsync(bits, 0, m / 8, ctx->stream));
    size_t vbytes = eb ? n * size_t(eb) : (n ? size_t(offsets[n]) : 0), obytes = eb ? 0 : (n + 1) * 4;
    CK(ctx->d_tmp.reserve(round_up(vbytes, 16) + obytes + 64));
    uint8_t* dv = static_cast<uint8_t*>(ctx->d_tmp.p);
    uint32_t* doff = reinterpret_cast<uint32_t*>(dv + round_up(vbytes, 16));
    if (vbytes) CK(cudaMemcpyAsync(dv, values, vbytes, cudaMemcpyHostToDevice, ctx->stream));
    if (obytes) CK(cudaMemcpyAsync(doff, offsets, obytes, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_bloom_build(dv, eb ? nullptr : doff, n, eb, reinterpret_cast<uint32_t*>(bits), uint32_t(m - 1), 4, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));   // host buffers are not retained
    return KX_OK;
    });
}

int kx_stats_get_bloom(kx_stats* st, int field_index, int pack_index, void* out, size_t cap, size_t* len) {
    return kx_guarded<int>(st ? st->ctx : nullptr, [&]() -> int {
    if (!st) return KX_EINVAL;
    kx_ctx* ctx = st->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (field_index < 0 || field_index >= st->nfields || pack_index < 0 || pack_index >= st->npacks || !len)
        return fail(ctx, KX_EINVAL, "kx_stats_get_bloom: bad arguments");
    size_t cell = size_t(field_index) * st->npacks + pack_index;
    if (!st->h_bloom_ptr[cell]) { *len = 0; return KX_OK; }
    size_t need = 1 + (size_t(st->h_bloom_mask[cell]) + 1) / 8;
    *len = need;
    if (!out) return KX_OK;
    if (cap < need) return fail(ctx, KX_EINVAL, "kx_stats_get_bloom: buffer too small");
    CK(cudaSetDevice(ctx->device));
    uint8_t* o = static_cast<uint8_t*>(out);
    o[0] = st->h_bloom_k[cell];
    CK(cudaMemcpyAsync(o + 1, reinterpret_cast<const void*>(uintptr_t(st->h_bloom_ptr[cell])), need - 1, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KX_OK;
    });
}

int64_t kx_prune_stats(kx_ctx* ctx, const kx_prog* prog, kx_stats* st, const uint64_t* hashes, const uint32_t* hash_off, uint8_t* out) {
    return kx_guarded<int64_t>(ctx, [&]() -> int64_t {
    if (!ctx) return KX_EINVAL;
    std::lock_guard<std::mutex> lk(ctx->mu);
    int rc = check_prog_job(ctx, prog);
    if (rc) return rc;
    if (!st || st->ctx != ctx || !out) return fail(ctx, KX_EINVAL, "kx_prune_stats: bad arguments");
    if ((hashes == nullptr) != (hash_off == nullptr)) return fail(ctx, KX_EINVAL, "kx_prune_stats: hashes and hash_off go together");
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, KX_ECUDA, "cudaSetDevice");
    const int nleaves = int(prog->leaves.size()), npacks = st->npacks;
    const size_t cells = size_t(npacks) * st->nfields;

    PruneStatsParams P{};
    P.npacks = uint32_t(npacks); P.nleaves = uint32_t(nleaves); P.npost = uint32_t(prog->postfix.size());
    std::memcpy(P.postfix, prog->postfix.data(), prog->postfix.size());
    std::vector<uint64_t> own_hashes;
    for (int l = 0; l < nleaves; ++l) {
        const LeafSpec& s = prog->leaves[size_t(l)];
        int fi = -1;
        for (int f = 0; f < st->nfields; ++f) if (st->fields[size_t(f)] == s.field) { fi = f; break; }
        if (fi < 0) return fail(ctx, KX_ENOTFOUND, "kx_prune_stats: leaf field has no statistics column");
        if (st->types[size_t(fi)] != s.type) return fail(ctx, KX_EINVAL, "kx_prune_stats: leaf / statistics type mismatch");
        P.leaf_field[l] = uint8_t(fi);
        P.leaf_nozone[l] = s.type == KX_BYTES;
        PruneLeaf& pl = P.leaves[l];
        pl.a = type_is_float(s.type) ? s.a : type_ext(s.type, s.a);
        pl.b = type_is_float(s.type) ? s.b : type_ext(s.type, s.b);
        pl.flip = type_is_signed(s.type) ? 0x8000000000000000ull : 0;
        pl.mode = s.mode; pl.is_float = type_is_float(s.type);
        pl.set_off = s.set_off; pl.nset = uint32_t(s.set.size());
        const uint64_t probe_a = pl.a;   // bloom probes hash the value in its own type
        if (s.type == KX_FLOAT32) { pl.a = f32_to_f64_bits(s.a); pl.b = f32_to_f64_bits(s.b); }
        if (hashes) { P.hash_off[l] = hash_off[l]; continue; }
        // probe hashes of numeric EQ / IN operands: hash.HashT(v) (internal/hash/hash.go:67-92)
        P.hash_off[l] = uint32_t(own_hashes.size());
        if (s.type != KX_BYTES && s.mode == KX_MODE_EQ) own_hashes.push_back(kx_hash_value(s.type, probe_a));
        if (s.type != KX_BYTES && s.mode == KX_MODE_IN) for (uint64_t v : s.set) own_hashes.push_back(kx_hash_value(s.type, v));
    }
    P.hash_off[nleaves] = hashes ? hash_off[nleaves] : uint32_t(own_hashes.size());
    const uint64_t* hsrc = hashes ? hashes : own_hashes.data();
    const size_t nhash = P.hash_off[nleaves];

    if (st->dirty) {   // (re)upload the filter tables after put/build calls
        CK(cudaMemcpyAsync(st->d_tab, st->h_bloom_ptr.data(), cells * 8, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(st->d_tab + cells * 8, st->h_bloom_mask.data(), cells * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(st->d_tab + cells * 12, st->h_bloom_k.data(), cells, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        st->dirty = false;
    }
    size_t out_words = (size_t(npacks) + 31) / 32;
    size_t off_out = round_up(nhash * 8, 8), off_cnt = round_up(off_out + out_words * 4, 8);
    CK(ctx->d_misc.reserve(off_cnt + 64));
    uint8_t* d = static_cast<uint8_t*>(ctx->d_misc.p);
    CK(cudaEventRecord(ctx->ev_start, ctx->stream));
    if (nhash) CK(cudaMemcpyAsync(d, hsrc, nhash * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(d + off_out, 0, off_cnt + 8 - off_out, ctx->stream));
    P.mins = st->d_mins; P.maxs = st->d_maxs;
    P.bloom_ptr = reinterpret_cast<const uint64_t*>(st->d_tab);
    P.bloom_mask = reinterpret_cast<const uint32_t*>(st->d_tab + cells * 8);
    P.bloom_k = st->d_tab + cells * 12;
    P.hashes = reinterpret_cast<const uint64_t*>(d);
    P.set_vals = prog->dev_sets;
    P.out = reinterpret_cast<uint32_t*>(d + off_out);
    P.count = reinterpret_cast<unsigned long long*>(d + off_cnt);
    CK(cudaEventRecord(ctx->ev_k0, ctx->stream));
    CK(launch_prune_stats(P, ctx->stream));
    CK(cudaEventRecord(ctx->ev_k1, ctx->stream));
    unsigned long long c = 0;
    CK(cudaMemcpyAsync(out, d + off_out, (size_t(npacks) + 7) / 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(&c, d + off_cnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventRecord(ctx->ev_end, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev_k0, ctx->ev_k1) == cudaSuccess) ctx->last_kernel_ms = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev_start, ctx->ev_end) == cudaSuccess) ctx->last_total_ms = ms;
    ctx->last_launches = 1;
    return int64_t(c);
    });
}

uint64_t kx_hash_value(uint8_t block_type, uint64_t pattern) {
    switch (type_bits(block_type)) {   // hash.HashT: internal/hash/hash.go:67-92 (floats hash their IEEE bytes)
    case 64: return xxh3_u64(pattern);
    case 32: return xxh3_u32(uint32_t(pattern));
    case 16: return xxh3_u16(uint16_t(pattern));
    case 8: return xxh3_u8(uint8_t(pattern));
    }
    return 0;
}
uint64_t kx_hash_bytes(const void* p, size_t len) { return xxh3_bytes(static_cast<const uint8_t*>(p), len); }

}  // extern "C"
