// kx_general.cu — the multi-leaf filter + fused reduce kernel of libknoxgpu (sm_100a).
//
// Replaces (reference, CPU, one pack at a time): filter.Match / MatchAnd / MatchOr
// (internal/operator/filter/match_core.go:14-215) over the container matchers of internal/encode, followed by
// CountResult / StreamResult + Reducer.Reduce (internal/query/result.go:44-152, internal/reducer/reducer.go:138-314)
// — for a whole batch of packs in one launch, without a decoded column vector or a match bitset ever touching HBM
// (bitsets are written only when the caller asks for them).
//
// Structure (persistent CTAs, 8 consumer warps + 1 TMA producer warp):
//   * Tiles of 256 R rows (R = 32 or 64 groups of 32 rows per consumer warp: one or two passes) are dealt to the CTAs in chunks of
//     `sched_chunk` consecutive tiles, round-robin: a static schedule (results are reproducible run to run) that
//     spreads the expensive and the cheap regions of every pack (e.g. the part of a time-ordered pack a range
//     predicate selects) over all SMs.
//   * The producer streams, per tile, one ring stage per staged leaf column (postfix order) with TMA bulk copies.
//   * Every consumer warp evaluates the leaves of its R groups pass by pass ("bitset word per lane").  Pure AND / pure
//     OR programs keep the running words in registers (MatchAnd / MatchOr early-outs per warp and pass); other trees
//     use a per-warp AND/OR stack in shared memory.
//   * The reduce lags ONE tile behind the filter: the match words of tile t stay in registers while the warp filters
//     tile t + 1, and the value rows of tile t arrive meanwhile — prefetched towards L2 by the lanes that own the
//     matches (sparse tiles, read on demand afterwards) or streamed through the ring by the producer behind the leaf
//     columns of tile t + 1 (dense tiles; the producer decides from the selectivity the consumers report).  No CTA
//     barrier, no shared match words: a warp reduces the rows it filtered.
//   * Lane ↔ row assignment of the reduce (the same whether the values were staged or are read on demand, so a tile
//     gives bit-identical partial sums either way): a pass (32 groups) is cut into KP ∈ {1, 2, 4} chunks of G = 32 / KP
//     groups; in chunk q lane l owns G consecutive rows of group q G + l mod G, starting at row (l / G) G of the
//     group, and visits them in ascending order rotated by l mod G — the rotation makes the shared-memory reads of a
//     staged chunk bank-conflict free (lane-private 64-bit rows, 256 B apart before the rotation).
//   * Per-thread accumulators live in registers (the kernel is instantiated per number of value columns); a fixed
//     shuffle tree and a fixed warp order give the per-CTA partial; the CTA that finishes last combines the partials of
//     all CTAs in CTA order (fixed topology: bit-reproducible), so a query is ONE launch.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_leaf.cuh"

namespace kx {

namespace {

// static tile schedule shared by the producer and the consumers of a CTA: `sched_rounds` full rounds in which CTA b
// takes chunk (round * gridDim + b) of sched_chunk consecutive tiles, then the remaining tiles one by one (round-robin):
// no CTA gets more than one tile above the average, whatever the chunk size
struct TileSched {
    const ScanParams& P;
    uint32_t tile_rows;
    uint32_t t = 0, t_stop = 0, round = 0;
    uint32_t pack = 0, chunk = 0, pack_tiles = 0;
    PackInfo pi{};

    __device__ __forceinline__ TileSched(const ScanParams& p, uint32_t tr) : P(p), tile_rows(tr) {}
    __device__ __forceinline__ bool open_chunk() {
        uint64_t t0;
        if (round < P.sched_rounds) {
            t0 = ((uint64_t)round * gridDim.x + blockIdx.x) * P.sched_chunk;
            t_stop = (uint32_t)(t0 + P.sched_chunk);
        } else {
            t0 = (uint64_t)P.sched_rounds * gridDim.x * P.sched_chunk + (uint64_t)(round - P.sched_rounds) * gridDim.x + blockIdx.x;
            if (t0 >= P.ntiles) return false;
            t_stop = (uint32_t)t0 + 1u;
        }
        t = (uint32_t)t0;
        pack = P.tile_pack ? __ldg(P.tile_pack + t) : t / P.tiles_per_pack;
        pi = P.packs[pack];
        pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
        chunk = t - pi.tile0;
        return true;
    }
    __device__ __forceinline__ bool start() { round = 0; return open_chunk(); }
    __device__ __forceinline__ bool next() {
        if (++t < t_stop) {
            if (++chunk >= pack_tiles) {
                do { ++pack; pi = P.packs[pack]; } while (pi.n == 0);
                pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
                chunk = 0;
            }
            return true;
        }
        ++round;
        return open_chunk();
    }
};

}  // namespace

// NA = value columns reduced by this instantiation (0: filter only; the NA = 4 instantiation also serves 3)
template <int NA, int MINB>
__global__ void __launch_bounds__(SCAN_THREADS, MINB) scan_general_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint8_t* stage_base = smem + 128;
    __shared__ AggAcc warp_acc[CONSUMER_WARPS];
    __shared__ unsigned long long warp_cnt[CONSUMER_WARPS];
    __shared__ unsigned int sm_match[CONSUMER_WARPS], sm_wtiles[CONSUMER_WARPS];   // per warp: matches / tiles finished so far — selectivity feedback for the producer (plain stores)
    // The staging decision of every tile is EXACT: the producer waits until all consumer warps have published the tile's
    // match count (they are already filtering the next tile, whose leaf columns are in flight), decides, records the
    // decision in a small ring and bumps sm_decided; the consumers read it when they get to the tile's reduce.
    __shared__ uint32_t sm_dense[8];
    __shared__ unsigned int sm_decided;
    constexpr bool AGG = NA > 0;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t R = P.R, tile_rows = R * 32u * CONSUMER_WARPS, nstages = P.stages, passes = R >> 5;
    const uint32_t nl = P.nleaves, na = AGG ? P.naggs : 0u;
    uint32_t* code_smem = reinterpret_cast<uint32_t*>(stage_base + (size_t)nstages * P.stage_bytes);   // bitmaps / prefilters / small tables
    // reduce chunks: KP per pass, G groups each; a staged chunk = one slice of G groups (32 G rows) per consumer warp
    const uint32_t KP = AGG ? P.agg_kp : 1u, G = 32u / KP, lgG = 31u - (uint32_t)__clz((int)G);

    if (threadIdx.x == 0) {
        for (int w = 0; w < CONSUMER_WARPS; ++w) { sm_match[w] = 0; sm_wtiles[w] = 0; }
        sm_decided = 0;
        for (uint32_t s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TileSched ts(P, tile_rows);

    if (warp == CONSUMER_WARPS) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0 && ts.start()) {
            uint32_t s = 0, ph = 0;
            auto acquire = [&]() {
                if (P.prod_sleep) mbar_wait_relaxed(&empty_bar[s], ph ^ 1u); else mbar_wait(&empty_bar[s], ph ^ 1u);   // slot released by all consumer warps
            };
            auto advance = [&]() { if (++s == nstages) { s = 0; ph ^= 1u; } };
            auto load_stream = [&](const uint8_t* data, size_t off, uint32_t w, uint32_t rows) {
                if (!data) return;                          // (the consumers test the same pointer)
                acquire();
                const uint32_t bytes = w ? ((((rows * w + 7u) >> 3) + 15u) & ~15u) : 0u;
                mbar_expect_tx(&full_bar[s], bytes);        // arrive (count 1) + expected bytes
                if (bytes) tma_load_1d(stage_base + (size_t)s * P.stage_bytes, data + off, bytes, &full_bar[s]);
                advance();
            };
            // the value columns of one (finished) tile: per column, pass and chunk one stage of eight slices
            auto load_values = [&](uint32_t vpack, uint32_t row0, uint32_t n) {
                for (uint32_t j = 0; j < na; ++j) {
                    const ColView& v = P.views[P.agg_view0 + (size_t)vpack * na + j];
                    if (!agg_stageable(v)) continue;
                    const uint32_t slice_rows = 32u * G, slice_bytes = slice_rows / 8u * v.width;
                    for (uint32_t pass = 0; pass < passes; ++pass) {
                        for (uint32_t q = 0; q < KP; ++q) {
                            acquire();
                            uint32_t total = 0;
                            for (uint32_t w = 0; w < CONSUMER_WARPS; ++w) {
                                const uint32_t r0 = row0 + ((w * R + pass * 32u + q * G) << 5);
                                const uint32_t nr = r0 < n ? min(slice_rows, n - r0) : 0u;
                                total += nr ? ((((nr * v.width + 7u) >> 3) + 15u) & ~15u) : 0u;
                            }
                            mbar_expect_tx(&full_bar[s], total);
                            for (uint32_t w = 0; w < CONSUMER_WARPS; ++w) {
                                const uint32_t r0 = row0 + ((w * R + pass * 32u + q * G) << 5);
                                const uint32_t nr = r0 < n ? min(slice_rows, n - r0) : 0u;
                                if (!nr) continue;
                                const uint32_t bytes = (((nr * v.width + 7u) >> 3) + 15u) & ~15u;
                                tma_load_1d(stage_base + (size_t)s * P.stage_bytes + (size_t)w * slice_bytes, v.data + (size_t)(r0 >> 3) * v.width, bytes, &full_bar[s]);
                            }
                            advance();
                        }
                    }
                }
            };
            // decision for the tile with sequence number `seq` (0-based, in this CTA's schedule): wait for its match count
            uint32_t m_seen = 0;
            auto decide = [&](uint32_t seq, uint32_t rows) -> bool {
                uint32_t m;
                for (;;) {
                    uint32_t done = 0xffffffffu;
                    m = 0;
                    for (int w = 0; w < CONSUMER_WARPS; ++w) {
                        done = min(done, *(volatile unsigned int*)&sm_wtiles[w]);
                        m += *(volatile unsigned int*)&sm_match[w];
                    }
                    if (done > seq) break;
                    __nanosleep(64);
                }
                // (a warp that ran ahead may already have added the next tile's matches: the count is cumulative, the
                // surplus is not counted again next time; the decision only has to be the SAME for all consumers)
                const uint32_t cnt = m - m_seen;
                m_seen = m;
                const bool dense = P.agg_dense_thr == 0u ? true : (P.agg_dense_thr == 0xffffffffu ? false : (uint64_t)cnt * P.agg_dense_thr > rows);
                sm_dense[seq & 7u] = dense ? 1u : 0u;
                __threadfence_block();
                *(volatile unsigned int*)&sm_decided = seq + 1u;
                return dense;
            };
            bool have_prev = false;
            uint32_t prev_pack = 0, prev_row0 = 0, prev_n = 0, prev_rows = 0, seq = 0;
            do {
                const uint32_t row0 = ts.chunk * tile_rows, rows = min(tile_rows, ts.pi.n - row0);
                const PackLeaf* L = P.leaves + (size_t)ts.pack * nl;
                const size_t tile_byte0 = (size_t)ts.chunk * (tile_rows / 8u);   // * width = first byte of the tile in a stream
                for (uint32_t i = 0; i < P.npost; ++i) {
                    const uint32_t op = P.postfix[i];
                    if (op >= 0x80u) continue;
                    load_stream(L[op].data, tile_byte0 * L[op].width, L[op].width, rows);
                    if (L[op].fixmode) load_stream(L[op].fix, tile_byte0, 1u, rows);      // ALP patch correction stream
                }
                if constexpr (AGG) {
                    // the previous tile's value columns follow this tile's leaf columns — if it matched densely
                    if (have_prev && decide(seq - 1u, prev_rows)) load_values(prev_pack, prev_row0, prev_n);
                    prev_pack = ts.pack; prev_row0 = row0; prev_n = ts.pi.n; prev_rows = rows; have_prev = true;
                }
                ++seq;
            } while (ts.next());
            if constexpr (AGG) {
                if (have_prev && decide(seq - 1u, prev_rows)) load_values(prev_pack, prev_row0, prev_n);   // the value columns of the last tile
            }
        }
        return;
    }

    // ===================== consumers: unpack + filter + reduce =====================
    // (passes <= 2: R is 32 or 64, so the per-pass match words are two named registers and every pass loop is unrolled)
    AggAcc acc[AGG ? NA : 1];
#pragma unroll
    for (int j = 0; j < (AGG ? NA : 1); ++j) acc[j] = agg_identity(AGG ? P.agg_type[j] : 0);
    unsigned long long nmatch = 0;   // matches this thread accounted for (per-CTA totals only)
    uint32_t lane_cnt = 0;           // matches of the current pack seen by this lane
    const uint32_t gw0 = warp * R;   // first group (of the tile) of this warp
    const bool two = passes == 2u;
    // shared memory behind the bitmaps: the warp's AND/OR stack (general trees only) and its descriptor cache
    // (the current pack's leaves; the value-column views of the current and the previous pack)
    uint32_t* stk = code_smem + P.stack_off_words + warp * (P.stack_depth * passes * 32u);
    uint32_t* wdesc = code_smem + P.desc_off_words + warp * P.desc_words;
    const PackLeaf* L = reinterpret_cast<const PackLeaf*>(wdesc);
    const ColView* AV = reinterpret_cast<const ColView*>(wdesc + nl * (sizeof(PackLeaf) / 4u));   // [2][na]
    uint32_t cur_pack = 0xffffffffu, desc_sel = 0;

    auto flush_count = [&](uint32_t pk) {
        uint32_t c = __reduce_add_sync(0xffffffffu, lane_cnt);
        if (P.counts && lane == 0 && c) atomicAdd(P.counts + pk, (unsigned long long)c);
        lane_cnt = 0;
    };

    {
        // hash-set leaves: prefilter bitmaps (and small exact tables) are the same for every pack — copy them into shared
        // memory once per CTA
        bool any = false;
        for (uint32_t l = 0; l < nl; ++l) {
            if (!P.pre_log2[l]) continue;
            any = true;
            const uint32_t npre = (1u << P.pre_log2[l]) >> 5;
            for (uint32_t i = threadIdx.x; i < npre; i += CONSUMER_WARPS * 32u) code_smem[P.hs_smem_off[l] + i] = __ldg(P.set_pre + P.pre_off[l] + i);
            if (P.hs_tab_smem_off[l] != 0xffffffffu) {
                const uint32_t nt = 8u << P.tab_log2[l];
                const uint32_t* src = reinterpret_cast<const uint32_t*>(P.set_tabs + P.tab_off[l]);
                for (uint32_t i = threadIdx.x; i < nt; i += CONSUMER_WARPS * 32u) code_smem[P.hs_tab_smem_off[l] + i] = __ldg(src + i);
            }
        }
        if (any) asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
    }

    uint32_t s = 0, ph = 0;
    const uint32_t full0 = smem_addr(full_bar), empty0 = smem_addr(empty_bar);   // 32-bit shared addresses of the ring barriers
    auto release = [&]() {   // this warp is done with ring stage s
        __syncwarp();
        if (lane == 0) mbar_arrive_a(empty0 + s * 8u);
        if (++s == nstages) { s = 0; ph ^= 1u; }
    };

    // ---- the lagging reduce: match words of the previous tile (registers), where its rows are, which views describe it
    uint32_t wp0 = 0u, wp1 = 0u, fb_match = 0u, fb_tiles = 0u, seq = 0u;   // seq: tiles this warp has filtered
    bool have_prev = false, prev_any = false;
    uint32_t prev_row0 = 0, prev_sel = 0;
    auto reduce_prev = [&](bool dense) {
        const uint32_t gl = lane & (G - 1u), sub = lane >> lgG, rot = gl;
#pragma unroll
        for (int j = 0; j < (AGG ? NA : 0); ++j) {
            if ((uint32_t)j >= na) break;
            const ColView& v = AV[prev_sel * na + j];
            const bool staged = dense && agg_stageable(v);
            if (!staged && !prev_any) continue;   // nothing matched in this warp's rows and no ring stage to consume
            const int type = P.agg_type[j];
            const bool raw64 = v.kind == CK_BITS && v.width == 64;
            const uint64_t flip = type_is_signed(type) ? 0x8000000000000000ull : 0ull;
            const uint32_t slice_bytes = 4u * G * v.width;   // 32 G rows of the column
            AggAcc a = acc[j];
#pragma unroll
            for (uint32_t pass = 0; pass < 2u; ++pass) {
                if (pass >= passes) break;
                const uint32_t wp = pass ? wp1 : wp0;
                for (uint32_t q = 0; q < KP; ++q) {
                    if (staged) mbar_wait_a(full0 + s * 8u, ph);
                    const uint32_t word = KP == 1u ? wp : __shfl_sync(0xffffffffu, wp, q * G + gl);
                    const uint32_t r = lane_rows(word, sub, G, rot);
                    const uint32_t srow0 = gl * 32u + sub * G;                                               // first row of the lane's range inside the warp's slice
                    const uint32_t row0 = prev_row0 + ((gw0 + pass * 32u + q * G) << 5) + srow0;            // … and inside the pack
                    const uint8_t* stg = stage_base + (size_t)s * P.stage_bytes + (size_t)warp * slice_bytes;
                    if (raw64) {
                        if (staged) {
                            const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(stg) + srow0;
                            if (type == 9) reduce_staged_raw64<true>(a, vp, r, G, rot, 0ull, 0ull);
                            else reduce_staged_raw64<false>(a, vp, r, G, rot, v.base, flip);
                        } else {
                            const unsigned long long* gp = reinterpret_cast<const unsigned long long*>(v.data) + row0;
                            if (type == 9) reduce_global_raw64<true>(a, gp, r, G, rot, 0ull, 0ull);
                            else reduce_global_raw64<false>(a, gp, r, G, rot, v.base, flip);
                        }
                    } else {
                        reduce_generic(a, v, type, row0, staged ? reinterpret_cast<const uint32_t*>(stg) : nullptr, srow0, r, G, rot);
                    }
                    if (staged) release();
                }
            }
            acc[j] = a;
        }
    };

    if (ts.start()) {
        do {
            const uint32_t pack = ts.pack, n = ts.pi.n;
            const uint32_t pack_row0 = ts.chunk * tile_rows;          // first row of the tile within the pack
            if (pack != cur_pack) {
                // new pack: this warp caches its descriptors; the consumers copy its code bitmaps (built by codeset_kernel)
                if (cur_pack != 0xffffffffu) flush_count(cur_pack);
                __syncwarp();
                desc_sel ^= 1u;
                {
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.leaves + (size_t)pack * nl);
                    for (uint32_t i = lane; i < nl * (uint32_t)(sizeof(PackLeaf) / 4u); i += 32u) wdesc[i] = __ldg(src + i);
                    if (AGG) {
                        const uint32_t* vsrc = reinterpret_cast<const uint32_t*>(P.views + P.agg_view0 + (size_t)pack * na);
                        uint32_t* vdst = wdesc + nl * (uint32_t)(sizeof(PackLeaf) / 4u) + desc_sel * na * (uint32_t)(sizeof(ColView) / 4u);
                        for (uint32_t i = lane; i < na * (uint32_t)(sizeof(ColView) / 4u); i += 32u) vdst[i] = __ldg(vsrc + i);
                    }
                }
                __syncwarp();
                if (P.code_bitmap_words) {
                    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));   // everybody is done with the previous pack's bitmaps
                    for (uint32_t l = 0; l < nl; ++l) {
                        if (L[l].mode != LM_CODESET) continue;
                        const uint32_t nw = (((1u << L[l].width) + (uint32_t)L[l].wm + 31u) >> 5) + 1u;
                        const uint32_t* src = P.code_bits + L[l].a;
                        uint32_t* dst = code_smem + P.code_smem_off[l];
                        for (uint32_t i = threadIdx.x; i < nw; i += CONSUMER_WARPS * 32u) dst[i] = __ldg(src + i);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
                }
                cur_pack = pack;
            }
            const LeafEnv env{P, code_smem, n, pack_row0};
            const uint64_t wr0 = (uint64_t)pack_row0 + (uint64_t)(gw0 + lane) * 32u;   // first pack row of this lane's word, pass 0 (pass 1: + 1024)

            // ---- leaves and the AND/OR program.  Every staged leaf column is one ring stage holding the column's slice of
            // the whole tile; a leaf is evaluated for both passes of the warp before the next one is touched.
            uint32_t w0 = 0u, w1 = 0u;
            if (P.flat_op) {
                // pure AND (1) / pure OR (2) program: running words in registers.  MatchAnd's early-out (match_core.go:44-130),
                // per warp and pass: rows the words so far have ruled out need no work — a pass whose 1024 rows are all ruled
                // out skips the leaf altogether (time-range filters on ordered packs rule out whole tiles); MatchOr's
                // early-out (:132-215) is the mirror image: rows that already matched.
                const bool is_and = P.flat_op == 1u;
                bool first = true;
                for (uint32_t i = 0; i < P.npost; ++i) {
                    const uint32_t op = P.postfix[i];
                    if (op >= 0x80u) continue;
                    const PackLeaf lf = L[op];   // registers: the fields a leaf kind needs are read once per tile
                    const uint32_t* sw = nullptr;
                    const bool staged_leaf = lf.data != nullptr;
                    if (staged_leaf) {
                        mbar_wait_a(full0 + s * 8u, ph);
                        sw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
                    }
                    const uint32_t flipw = lf.neg2 ? 0xffffffffu : 0u;
                    {
                        const uint32_t keep = first ? 0xffffffffu : (is_and ? w0 : ~w0);
                        if (__any_sync(0xffffffffu, keep != 0u)) {
                            const uint32_t word = eval_leaf(env, lf, op, sw, gw0, 32u, lane, wr0, keep) ^ flipw;
                            w0 = first ? word : (is_and ? (w0 & word) : (w0 | word));
                        }
                    }
                    if (two) {
                        const uint32_t keep = first ? 0xffffffffu : (is_and ? w1 : ~w1);
                        if (__any_sync(0xffffffffu, keep != 0u)) {
                            const uint32_t word = eval_leaf(env, lf, op, sw, gw0 + 32u, 32u, lane, wr0 + 1024u, keep) ^ flipw;
                            w1 = first ? word : (is_and ? (w1 & word) : (w1 | word));
                        }
                    }
                    if (staged_leaf) release();
                    first = false;
                }
            } else {
                // general tree: the per-pass words wait on the warp's stack in shared memory (lane-private columns)
                const uint32_t pstride = passes * 32u;
                uint32_t sp = 0;
                for (uint32_t i = 0; i < P.npost; ++i) {
                    const uint32_t op = P.postfix[i];
                    if (op < 0x80u) {
                        const PackLeaf& lf = L[op];
                        uint32_t* dst = stk + sp * pstride + lane;
                        const uint32_t* sw = nullptr;
                        const bool staged_leaf = lf.data != nullptr;
                        if (staged_leaf) {
                            mbar_wait_a(full0 + s * 8u, ph);
                            sw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
                        }
                        const bool inv = lf.neg2 && !lf.fixmode;
                        const bool and_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFEu;
                        const bool or_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFFu;
                        const uint32_t* prev = stk + (sp ? sp - 1u : 0u) * pstride + lane;
                        for (uint32_t pass = 0; pass < passes; ++pass) {
                            const uint32_t g0 = gw0 + pass * 32u;
                            const uint32_t keep = and_next ? prev[pass * 32u] : (or_next ? ~prev[pass * 32u] : 0xffffffffu);
                            uint32_t word = 0;
                            if (__any_sync(0xffffffffu, keep != 0u)) {
                                word = eval_leaf(env, lf, op, sw, g0, 32u, lane, wr0 + pass * 1024u, keep);
                                if (inv) word = ~word;
                            }
                            dst[pass * 32u] = word;
                        }
                        if (staged_leaf) release();
                        if (lf.fixmode) {   // ALP: correct the rows that are patches (1-bit stream in the next stage)
                            mbar_wait_a(full0 + s * 8u, ph);
                            const uint32_t* fw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
                            __builtin_assume(__isShared(fw));
                            for (uint32_t pass = 0; pass < passes; ++pass) {
                                const uint32_t fx = fw[gw0 + pass * 32u + lane];
                                uint32_t word = dst[pass * 32u];
                                word = lf.fixmode == FIX_OR_PRED ? (word | fx) : (word & ~fx);
                                dst[pass * 32u] = lf.neg2 ? ~word : word;
                            }
                            release();
                        }
                        ++sp;
                    } else {
                        --sp;
                        uint32_t* x = stk + (sp - 1u) * pstride + lane;
                        const uint32_t* y = stk + sp * pstride + lane;
                        for (uint32_t pass = 0; pass < passes; ++pass)
                            x[pass * 32u] = (op == 0xFEu) ? (x[pass * 32u] & y[pass * 32u]) : (x[pass * 32u] | y[pass * 32u]);
                    }
                }
                w0 = stk[lane];
                if (two) w1 = stk[32u + lane];
            }

            // ---- outputs of the tile: tail masking (match_core.go semantics: tail bits zero; only the last tile of a pack has
            // a tail), bitset words (coalesced 128 B per warp), per-pack match count
            if (pack_row0 + tile_rows > n) {
                auto tail = [&](uint32_t word, uint64_t wr) -> uint32_t {
                    if (wr >= n) return 0u;
                    const uint32_t left = n - (uint32_t)wr;
                    return left >= 32u ? word : (word & ((1u << left) - 1u));
                };
                w0 = tail(w0, wr0);
                w1 = two ? tail(w1, wr0 + 1024u) : 0u;
            }
            if (P.bitsets) {
                if (wr0 < n) *reinterpret_cast<uint32_t*>(P.bitsets + ts.pi.bitset_off + (wr0 >> 3)) = w0;
                if (two && wr0 + 1024u < n) *reinterpret_cast<uint32_t*>(P.bitsets + ts.pi.bitset_off + ((wr0 + 1024u) >> 3)) = w1;
            }
            const uint32_t tile_cnt = __popc(w0) + __popc(w1);
            lane_cnt += tile_cnt;

            if constexpr (AGG) {
                nmatch += tile_cnt;   // per-CTA totals only: any partition of the matches over threads will do
                const bool any_now = __any_sync(0xffffffffu, tile_cnt != 0u);
                if (any_now) fb_match += __reduce_add_sync(0xffffffffu, tile_cnt);
                ++fb_tiles;
                if (lane == 0) { *(volatile unsigned int*)&sm_match[warp] = fb_match; *(volatile unsigned int*)&sm_wtiles[warp] = fb_tiles; }   // the producer decides on these
                // the previous tile's value rows have had this tile's filter time to arrive: staged by the producer when the
                // tile matched densely (its decision is exact and waits for nothing but this publication), else prefetched
                bool dense = false;
                if (have_prev) {
                    while (*(volatile unsigned int*)&sm_decided < seq) {}
                    dense = sm_dense[(seq - 1u) & 7u] != 0u;
                    if (dense || prev_any) reduce_prev(dense);
                }
                // this tile's turn comes after the next filter: start pulling its matching rows towards L2 now (raw 64-bit
                // columns; skipped while the producer is staging whole tiles anyway)
                if (any_now && tile_cnt <= 8u) {   // (a lane with many matches sits in a dense tile: the producer will stage it)
#pragma unroll
                    for (int j = 0; j < NA; ++j) {
                        if ((uint32_t)j >= na) break;
                        const ColView& v = AV[desc_sel * na + j];
                        if (v.kind != CK_BITS || v.width != 64) continue;
#pragma unroll
                        for (uint32_t pass = 0; pass < 2u; ++pass) {
                            uint32_t word = pass ? w1 : w0;
                            if (!word) continue;
                            const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(v.data) + (wr0 + pass * 1024u);
                            if (__popc(word) <= 4) {
                                while (word) {
                                    const uint32_t b = (uint32_t)__ffs((int)word) - 1u;
                                    word &= word - 1u;
                                    asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + b));
                                }
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; ++k)
                                    if ((word >> (4 * k)) & 0xfu) asm volatile("prefetch.global.L2 [%0];" ::"l"(vp + 4 * k));
                            }
                        }
                    }
                }
                wp0 = w0; wp1 = w1;
                prev_row0 = pack_row0; prev_sel = desc_sel; have_prev = true; prev_any = any_now;
                ++seq;
            }
        } while (ts.next());
        flush_count(cur_pack);
        if constexpr (AGG) {
            // the value columns of the last tile
            while (*(volatile unsigned int*)&sm_decided < seq) {}
            const bool dense = sm_dense[(seq - 1u) & 7u] != 0u;
            if (dense || prev_any) reduce_prev(dense);
        }
    }

    if constexpr (AGG) {
        // ---- per-CTA partial aggregates: fixed-order tree inside the warp, then across warps
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            if ((uint32_t)j >= na) break;
            const int type = P.agg_type[j];
            AggAcc a = acc[j];
            unsigned long long c = nmatch;
            for (int off = 16; off > 0; off >>= 1) {
                AggAcc b;
#pragma unroll
                for (int q = 0; q < 4; ++q) b.s[q] = __shfl_down_sync(0xffffffffu, a.s[q], off);
                unsigned long long cb = __shfl_down_sync(0xffffffffu, c, off);
                agg_merge(a, b, type);
                c += cb;
            }
            if (lane == 0) { warp_acc[warp] = a; warp_cnt[warp] = c; }
            // consumer-only barrier (the producer warp has exited)
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
            if (threadIdx.x == 0) {
                AggAcc r = warp_acc[0];
                unsigned long long rc = warp_cnt[0];
                for (int q = 1; q < CONSUMER_WARPS; ++q) { agg_merge(r, warp_acc[q], type); rc += warp_cnt[q]; }
                AggPartial o;
                o.count = rc; o.valid = rc != 0; o.pad = 0;
                if (type == 9 || type == 10) { o.sum = r.s[0]; o.err = as_f64(r.s[1]); o.mn = r.s[2]; o.mx = r.s[3]; }
                else { o.sum = r.s[0]; o.err = 0.0; o.mn = r.s[1]; o.mx = r.s[2]; }
                P.partials[(size_t)blockIdx.x * na + j] = o;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
        }

        // ---- the CTA that finishes last combines the per-CTA partials: every thread merges a contiguous run of them in
        // index order, a fixed shuffle tree and a fixed warp order do the rest — the topology depends on the grid size
        // only, so the result is bit-reproducible whichever CTA happens to be last (no separate launch, no serial walk)
        __shared__ uint32_t sm_last;
        __shared__ AggPartial warp_part[CONSUMER_WARPS];
        if (threadIdx.x == 0) {
            __threadfence();
            sm_last = atomicAdd(P.done, 1u) == gridDim.x - 1u;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
        if (sm_last) {
            __threadfence();
            const uint32_t nparts = gridDim.x, T = CONSUMER_WARPS * 32u, per = (nparts + T - 1u) / T;
            for (uint32_t j = 0; j < na; ++j) {
                const int type = P.agg_type[j];
                AggPartial r{};
                const uint32_t i1 = min(nparts, (threadIdx.x + 1u) * per);
                for (uint32_t i = threadIdx.x * per; i < i1; ++i) {
                    const ulonglong2* src = reinterpret_cast<const ulonglong2*>(P.partials + (size_t)i * na + j);
                    union { AggPartial p; ulonglong2 q[3]; } u;
                    u.q[0] = __ldcg(src); u.q[1] = __ldcg(src + 1); u.q[2] = __ldcg(src + 2);
                    partial_merge(r, u.p, type);
                }
                for (int off = 1; off < 32; off <<= 1) {   // lane l absorbs lane l + off: partials stay in index order
                    AggPartial o;
                    o.count = __shfl_down_sync(0xffffffffu, r.count, off);
                    o.sum = __shfl_down_sync(0xffffffffu, r.sum, off);
                    o.err = __shfl_down_sync(0xffffffffu, r.err, off);
                    o.mn = __shfl_down_sync(0xffffffffu, r.mn, off);
                    o.mx = __shfl_down_sync(0xffffffffu, r.mx, off);
                    o.valid = __shfl_down_sync(0xffffffffu, r.valid, off);
                    o.pad = 0;
                    if ((lane & (2u * off - 1u)) == 0u) partial_merge(r, o, type);
                }
                if (lane == 0) warp_part[warp] = r;
                asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
                if (threadIdx.x == 0) {
                    AggPartial f = warp_part[0];
                    for (int q = 1; q < CONSUMER_WARPS; ++q) partial_merge(f, warp_part[q], type);
                    P.agg_out[j] = f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
            }
        }
    }
}

cudaError_t launch_scan_general(const ScanParams& P, int grid, size_t smem_bytes, int ctas_per_sm, cudaStream_t stream) {
    int variant;
    void (*kern)(const ScanParams);
    if (P.naggs == 0) { variant = 0; kern = scan_general_kernel<0, 2>; }
    else if (P.naggs == 1 && ctas_per_sm >= 2) { variant = 1; kern = scan_general_kernel<1, 2>; }
    else if (P.naggs == 2 && ctas_per_sm >= 2) { variant = 2; kern = scan_general_kernel<2, 2>; }
    else { variant = 3; kern = scan_general_kernel<4, 1>; }
    // function attributes are per device and sticky: set them once per (device, variant)
    static bool configured[64][4] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev][variant]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SCAN_MAX_DYN_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev][variant] = true;
    }
    kern<<<grid, SCAN_THREADS, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
