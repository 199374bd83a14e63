// kx_warp.cu — the multi-leaf filter + fused reduce kernel of libknoxgpu, warp-autonomous version (sm_100a).
//
// Replaces (reference, CPU, one pack at a time): filter.Match / MatchAnd / MatchOr
// (internal/operator/filter/match_core.go:14-215) over the container matchers of internal/encode, followed by
// CountResult / StreamResult + Reducer.Reduce (internal/query/result.go:44-152, internal/reducer/reducer.go:138-314)
// — for a whole batch of packs in one launch, without a decoded column vector or a match bitset ever touching HBM
// (bitsets are written only when the caller asks for them).
//
// Structure: ONE persistent CTA per SM; every warp of it is a complete, independent pipeline.
//   * A tile is 1024 w_wd rows of one pack (w_wd = 1, 2 or 4 bitset words per lane: lane l owns the 32 consecutive rows
//     of group 32 j + l for j < w_wd).  Tiles are dealt to the warps of the whole grid in chunks of up to `sched_chunk`
//     tiles of one pack, round-robin — a static schedule: results are reproducible run to run.  The tiles of a chunk are
//     spread evenly over their pack, so every chunk samples the expensive and the cheap regions of a time-ordered pack
//     alike (the part a range predicate selects), and a warp changes pack only once per chunk.
//   * Each warp owns a small TMA ring in shared memory.  A ring stage holds the tile's slice of EVERY staged column of
//     the program (leaf streams, ALP patch-correction streams); lane i issues the bulk copy of column i and posts its
//     byte count on the stage's mbarrier, and the warp that issued the copies is the one that waits for them — there
//     is no producer warp, no empty barrier, no CTA-wide handshake, and the per-tile control code is a few dozen
//     instructions.  While a warp evaluates tile t, the columns of tiles t+1 … t+stages-1 are in flight.
//   * Leaves are evaluated leaf by leaf over the words of the tile ("bitset word per lane", kx_leaf.cuh); the running
//     match words live in a per-warp shared-memory array.  Pure AND / pure OR programs skip leaves whose operand rows
//     are already decided (MatchAnd / MatchOr early-outs per warp and word); other trees use a per-warp stack.
//   * The reduce lags ONE tile behind the filter: the match words of tile t wait in shared memory while the warp
//     filters tile t+1, and the value rows of tile t are pulled towards L2 meanwhile — per matching row by the lanes that
//     own the matches (sparse tiles), or as one bulk L2 prefetch of the tile's slice of the column (dense tiles) — and are
//     then read straight from global memory: matching rows only, or 256 contiguous bytes per lane with 128-bit loads.
//     The lane ↔ row assignment is the same either way, so a tile gives bit-identical partial sums on both paths.
//   * Per-thread accumulators live in registers (the kernel is instantiated per number of value columns); a fixed
//     shuffle tree and a fixed warp order give the per-CTA partial; the CTA that finishes last combines the partials of
//     all CTAs in CTA order (fixed topology: bit-reproducible), so a query is ONE launch.
#include <cuda_runtime.h>
#include <stdint.h>

#define KX_IMAD_ROWS 2   // row code: multiply-add + carry chain (kx_leaf.cuh)
#include "kx_leaf.cuh"

namespace kx {

namespace {

constexpr uint32_t END_PACK = 0xffffffffu;
constexpr int WARP_MAX_WARPS = 16;

// static tile schedule of one warp.  The scheduling unit is a CHUNK: up to `sched_chunk` tiles of one pack spread evenly over
// the pack (chunk j of a pack with nch chunks = tiles j, j + nch, j + 2 nch, …).  Chunks are numbered pack by pack and dealt
// round-robin to the warps of the grid (P.ntiles / PackInfo::tile0 / P.tile_pack / P.tiles_per_pack count chunks here).
struct WarpSched {
    uint32_t u = 0;                                   // chunk id
    uint32_t pack = 0, n = 0, tiles = 0, nch = 0;     // the chunk's pack: rows, tiles, chunks
    uint32_t tile = 0, step = 0;                      // current tile of the pack, its position in the chunk
    bool valid = false;

    __device__ __forceinline__ void open(const ScanParams& P, uint32_t tile_rows) {
        if (u >= P.ntiles) { valid = false; return; }
        pack = P.tile_pack ? __ldg(P.tile_pack + u) : u / P.tiles_per_pack;
        const PackInfo pi = P.packs[pack];
        n = pi.n;
        // (tile_rows and sched_chunk are powers of two: shifts instead of two integer divisions per chunk)
        tiles = (n + tile_rows - 1) >> (31 - __clz((int)tile_rows));
        nch = (tiles + P.sched_chunk - 1) >> (31 - __clz((int)P.sched_chunk));
        tile = u - pi.tile0;
        step = 0;
        valid = true;
    }
    __device__ __forceinline__ void next(const ScanParams& P, uint32_t nslots, uint32_t tile_rows) {
        tile += nch;
        if (++step < P.sched_chunk && tile < tiles) return;
        u += nslots;
        open(P, tile_rows);
    }
};

}  // namespace

// NA = value columns reduced by this instantiation (0: filter only; the NA = 4 instantiation also serves 3)
template <int NA, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) scan_warp_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ AggAcc warp_acc[WARP_MAX_WARPS];
    __shared__ unsigned long long warp_cnt[WARP_MAX_WARPS];
    constexpr bool AGG = NA > 0;
    constexpr uint32_t NWARPS = THREADS / 32;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t WPC = P.w_warps, wd = P.w_wd, S = P.stages, tile_rows = wd * 1024u;
    const uint32_t nl = P.nleaves, na = AGG ? P.naggs : 0u;
    uint32_t* code_smem = reinterpret_cast<uint32_t*>(smem + (size_t)WPC * P.w_warp_bytes);   // hash-set prefilters / small tables (per CTA)

    {
        // hash-set leaves: prefilter bitmaps (and small exact tables) are the same for every pack — copied once per CTA
        bool any = false;
        for (uint32_t l = 0; l < nl; ++l) {
            if (!P.pre_log2[l]) continue;
            any = true;
            const uint32_t npre = (1u << P.pre_log2[l]) >> 5;
            for (uint32_t i = threadIdx.x; i < npre; i += THREADS) code_smem[P.hs_smem_off[l] + i] = __ldg(P.set_pre + P.pre_off[l] + i);
            if (P.hs_tab_smem_off[l] != 0xffffffffu) {
                const uint32_t nt = 8u << P.tab_log2[l];
                const uint32_t* src = reinterpret_cast<const uint32_t*>(P.set_tabs + P.tab_off[l]);
                for (uint32_t i = threadIdx.x; i < nt; i += THREADS) code_smem[P.hs_tab_smem_off[l] + i] = __ldg(src + i);
            }
        }
        if (any) __syncthreads();
    }

    AggAcc acc[AGG ? NA : 1];
#pragma unroll
    for (int j = 0; j < (AGG ? NA : 1); ++j) acc[j] = agg_identity(AGG ? P.agg_type[j] : 0);
    unsigned long long nmatch = 0;   // matches this thread accounted for (per-CTA totals only)

    if (warp < WPC) {
        uint8_t* wb = smem + (size_t)warp * P.w_warp_bytes;
        const uint32_t T = P.w_mw_slots;                                 // match-word slots / tile records / value-view sets kept per warp
        uint64_t* full_bar = reinterpret_cast<uint64_t*>(wb);            // [S] (S <= 4)
        uint4* fifo = reinterpret_cast<uint4*>(wb + 64);                 // [S] what each ring stage holds: a tile's leaf columns or a value chunk
        uint4* tmeta = reinterpret_cast<uint4*>(wb + 128);               // [T <= 8] tiles whose reduce is pending: first row, rows of the pack, view set, raw-column mask
        uint32_t* mw = reinterpret_cast<uint32_t*>(wb + 256);            // [T][wd * 32] match words of the last T tiles
        uint32_t* stk = mw + T * wd * 32u;                               // [stack_depth][wd * 32]
        uint32_t* wdesc = stk + P.stack_depth * wd * 32u;                // nl PackLeaf + T na ColView
        uint8_t* stage0 = wb + P.w_stage_off;
        const PackLeaf* L = reinterpret_cast<const PackLeaf*>(wdesc);
        const ColView* AV = reinterpret_cast<const ColView*>(wdesc + nl * (uint32_t)(sizeof(PackLeaf) / 4u));   // [T][na]
        const uint32_t NC = P.w_ncols, stage_bytes = P.stage_bytes, CR = P.w_chunk_rows;
        const uint32_t slot = blockIdx.x * WPC + warp, nslots = gridDim.x * WPC;

        if (lane == 0) {
            for (uint32_t s = 0; s < S; ++s) mbar_init(&full_bar[s], NC);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();

        // ---- issue side.  A free ring stage takes, in this order: the next value chunk of a tile that matched densely (its
        // raw 64-bit value columns are streamed through the ring in chunks of CR rows, behind the leaf columns that were
        // already in flight), else the leaf columns of the warp's next tile (`ic`: the warp's own cursor, ahead of the tile
        // being evaluated; lane i keeps the stream pointer / width of column i for the cursor's pack in registers).
        WarpSched ic;
        ic.u = slot;
        ic.open(P, tile_rows);
        uint32_t iss_pack = END_PACK, iss_w = 0, iss_off = 0;
        const uint8_t* iss_data = nullptr;
        uint32_t iss_sl = 0xffu;
        if (lane < NC) { iss_sl = P.w_slot[lane]; iss_off = (uint32_t)P.w_slot_off[lane] * 16u; }
        // queue of dense tiles and the cursor inside its head tile, packed into one register (the state lives across every
        // leaf call): bits 0-14 slot ids (3 bits each, head lowest), 15-17 entries, 18-19 value column, 20-31 chunk
        uint32_t dqs = 0;
        auto issue = [&](uint32_t st) -> bool {
            if (AGG && (dqs & 0x38000u)) {
                const uint32_t h = dqs & 7u, vq_j = (dqs >> 18) & 3u, vq_c = dqs >> 20;
                const uint4 tm = tmeta[h];                                   // x: first row, y: rows of the pack, z: view set, w: raw-column mask
                const uint32_t rows_tile = min(tile_rows, tm.y - tm.x), c0 = vq_c * CR, rows = min(CR, rows_tile - c0);
                const uint32_t bytes = (rows * 8u + 15u) & ~15u;
                if (lane == 0) fifo[st] = make_uint4(0u, c0, rows, 0x80000000u | (h << 8) | vq_j);
                if (lane < NC) {
                    mbar_expect_tx(&full_bar[st], lane == 0 ? bytes : 0u);
                    if (lane == 0) {
                        const ColView& v = AV[tm.z * na + vq_j];
                        tma_load_1d(stage0 + (size_t)st * stage_bytes, v.data + ((size_t)tm.x + c0) * 8u, bytes, &full_bar[st]);
                    }
                }
                if ((vq_c + 1u) * CR < rows_tile) dqs += 1u << 20;           // next chunk
                else {                                                       // next raw column of the tile, else the next dense tile
                    const uint32_t rest = tm.w >> (vq_j + 1u);
                    if (rest) dqs = (dqs & 0x3ffffu & ~0xc0000u) | ((vq_j + (uint32_t)__ffs((int)rest)) << 18);
                    else {
                        const uint32_t cnt = ((dqs >> 15) & 7u) - 1u, ids = (dqs & 0x7fffu) >> 3;
                        dqs = ids | (cnt << 15);
                        if (cnt) dqs |= ((uint32_t)__ffs((int)tmeta[ids & 7u].w) - 1u) << 18;
                    }
                }
                return true;
            }
            if (!ic.valid) {
                if (lane == 0) fifo[st] = make_uint4(END_PACK, 0u, 0u, 0u);
                return false;
            }
            const uint32_t pack = ic.pack, row0 = ic.tile * tile_rows, n = ic.n;
            if (pack != iss_pack) {
                iss_pack = pack;
                iss_data = nullptr; iss_w = 0;
                if (iss_sl != 0xffu) {
                    const PackLeaf* q = P.leaves + (size_t)pack * nl + (iss_sl & 0x7fu);
                    if (iss_sl & 0x80u) { if (q->fixmode) { iss_data = q->fix; iss_w = 1u; } }
                    else { iss_data = q->data; iss_w = q->width; }
                }
            }
            if (lane == 0) fifo[st] = make_uint4(pack, row0, n, 0u);
            if (lane < NC) {
                const uint32_t rows = min(tile_rows, n - row0);
                const uint32_t bytes = (iss_data && iss_w) ? ((((rows * iss_w + 7u) >> 3) + 15u) & ~15u) : 0u;
                mbar_expect_tx(&full_bar[st], bytes);   // one arrival per column lane + its bytes
                if (bytes) tma_load_1d(stage0 + (size_t)st * stage_bytes + iss_off, iss_data + (size_t)(row0 >> 3) * iss_w, bytes, &full_bar[st]);
            }
            ic.next(P, nslots, tile_rows);
            return true;
        };
        uint32_t live = 0;               // ring stages that hold work
        for (uint32_t st = 0; st < S; ++st) live += issue(st) ? 1u : 0u;
        __syncwarp();

        // ---- evaluate side
        uint32_t lane_cnt = 0;           // matches of the current pack seen by this lane
        uint32_t cur_pack = END_PACK, desc_sel = 0, tslot = 0, cur_rawmask = 0;   // tslot: slot of the tile being evaluated (tile k uses slot k mod T)
        uint8_t* bits_base = nullptr;
        auto flush_count = [&](uint32_t pk) {
            const uint32_t c = __reduce_add_sync(0xffffffffu, lane_cnt);
            if (P.counts && lane == 0 && c) atomicAdd(P.counts + pk, (unsigned long long)c);
            lane_cnt = 0;
        };
        // value columns that are not streamed through the ring — every column of a sparse tile, the bit-packed / dictionary /
        // … columns of a dense one — are read on demand ONE tile behind the filter (their rows were prefetched towards L2)
        bool have_prev = false, prev_any = false, prev_dense = false;
        uint32_t prev_row0 = 0, prev_sel = 0, prev_slot = 0;
        auto reduce_prev = [&]() {
            const uint32_t* mwp = mw + prev_slot * wd * 32u;
#pragma unroll
            for (int j = 0; j < (AGG ? NA : 0); ++j) {
                if ((uint32_t)j >= na) break;
                const ColView& v = AV[prev_sel * na + j];
                const int type = P.agg_type[j];
                const bool raw64 = v.kind == CK_BITS && v.width == 64;
                if (raw64 && prev_dense) continue;   // comes through the ring
                const uint64_t flip = type_is_signed(type) ? 0x8000000000000000ull : 0ull, base = v.base;
                AggAcc a = acc[j];
#pragma unroll 1
                for (uint32_t jw = 0; jw < wd; ++jw) {
                    const uint32_t r = mwp[jw * 32u + lane];
                    if (!__any_sync(0xffffffffu, r != 0u)) continue;
                    const uint32_t row0 = prev_row0 + (jw * 32u + lane) * 32u;   // first pack row of the lane's word
                    if (raw64) {
                        const unsigned long long* gp = reinterpret_cast<const unsigned long long*>(v.data) + row0;
                        if (type == 9) reduce_global_raw64<true>(a, gp, r, 32u, 0u, 0ull, 0ull);
                        else reduce_global_raw64<false>(a, gp, r, 32u, 0u, base, flip);
                    } else {
                        reduce_generic(a, v, type, row0, nullptr, 0u, r, 32u, 0u);
                    }
                }
                acc[j] = a;
            }
        };
        // one value chunk of a dense tile, staged in the ring: rows c0 … c0 + rows of the tile, lane l on rows 32 q + l (whole
        // 256 B shared-memory rows per step: conflict free); the match bit of such a row is bit l of the word of group
        // c0 / 32 + q, read as a broadcast
        auto reduce_chunk = [&](const uint8_t* stage, uint32_t h, uint32_t jcol, uint32_t c0, uint32_t rows) {
            const uint4 tm = tmeta[h];
            const uint32_t* mwt = mw + h * wd * 32u + (c0 >> 5);
            const unsigned long long* sv = reinterpret_cast<const unsigned long long*>(stage) + lane;
            __builtin_assume(__isShared(sv));
            const uint32_t steps = (rows + 31u) >> 5;
#pragma unroll
            for (int j = 0; j < (AGG ? NA : 0); ++j) {
                if ((uint32_t)j != jcol) continue;
                const ColView& v = AV[tm.z * na + j];
                const int type = P.agg_type[j];
                const uint64_t flip = type_is_signed(type) ? 0x8000000000000000ull : 0ull, base = v.base;
                AggAcc a = acc[j];
                if (type == 9) {
#pragma unroll 4
                    for (uint32_t q = 0; q < steps; ++q)
                        if ((mwt[q] >> lane) & 1u) acc_raw64<true>(a, sv[q * 32u], 0ull, 0ull);
                } else {
#pragma unroll 4
                    for (uint32_t q = 0; q < steps; ++q)
                        if ((mwt[q] >> lane) & 1u) acc_raw64<false>(a, sv[q * 32u], base, flip);
                }
                acc[j] = a;
            }
        };

        uint32_t s = 0, phbits = 0;      // phbits: parity of every stage's barrier (a stage that held nothing is not waited on)
        while (live || (AGG && (dqs & 0x38000u))) {   // (a dense tile found after the last refill still has its value chunks to queue)
            const uint4 e = fifo[s];
            const uint32_t pack = e.x, pack_row0 = e.y, n = e.z;
            if (pack == END_PACK) {      // nothing was issued into this stage last time round: something may be pending now
                __syncwarp();
                live += issue(s) ? 1u : 0u;
                __syncwarp();
                if (++s == S) s = 0;
                continue;
            }
            const uint8_t* stage = stage0 + (size_t)s * stage_bytes;
            if (AGG && (e.w & 0x80000000u)) {
                // ---- a value chunk of a dense tile
                mbar_wait(&full_bar[s], (phbits >> s) & 1u);
                phbits ^= 1u << s;
                reduce_chunk(stage, (e.w >> 8) & 15u, e.w & 0xffu, e.y, e.z);
                __syncwarp();
                --live;
                live += issue(s) ? 1u : 0u;
                __syncwarp();
                if (++s == S) s = 0;
                continue;
            }
            if (pack != cur_pack) {
                // new pack: the warp caches its descriptors
                if (cur_pack != END_PACK) flush_count(cur_pack);
                __syncwarp();
                if (++desc_sel == T) desc_sel = 0;
                {
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(P.leaves + (size_t)pack * nl);
                    for (uint32_t i = lane; i < nl * (uint32_t)(sizeof(PackLeaf) / 4u); i += 32u) wdesc[i] = __ldg(src + i);
                    if (AGG) {
                        const uint32_t* vsrc = reinterpret_cast<const uint32_t*>(P.views + P.agg_view0 + (size_t)pack * na);
                        uint32_t* vdst = wdesc + nl * (uint32_t)(sizeof(PackLeaf) / 4u) + desc_sel * na * (uint32_t)(sizeof(ColView) / 4u);
                        for (uint32_t i = lane; i < na * (uint32_t)(sizeof(ColView) / 4u); i += 32u) vdst[i] = __ldg(vsrc + i);
                    }
                }
                if (P.bitsets) bits_base = P.bitsets + P.packs[pack].bitset_off;
                __syncwarp();
                cur_pack = pack;
                cur_rawmask = 0;
                if (AGG)
                    for (uint32_t j = 0; j < na; ++j) {
                        const ColView& v = AV[desc_sel * na + j];
                        if (v.kind == CK_BITS && v.width == 64) cur_rawmask |= 1u << j;
                    }
            }
            const LeafEnv env{P, code_smem, n, pack_row0};
            const uint32_t wr0 = pack_row0 + lane * 32u;   // first pack row of this lane's word 0 (word j: + 1024 j)
            uint32_t* mwc = mw + tslot * wd * 32u;

            mbar_wait(&full_bar[s], (phbits >> s) & 1u);   // every staged column of the tile has landed
            phbits ^= 1u << s;

            if (P.flat_op) {
                // pure AND (1) / pure OR (2) program.  MatchAnd's early-out (match_core.go:44-130), per warp and word: a word
                // whose 1024 rows are all ruled out skips the leaf altogether (time-range filters on ordered packs rule out
                // whole tiles); MatchOr's early-out (:132-215) is the mirror image: rows that already matched.
                const bool is_and = P.flat_op == 1u;
                bool first = true;
                for (uint32_t i = 0; i < P.npost; ++i) {
                    const uint32_t op = P.postfix[i];
                    if (op >= 0x80u) continue;
                    const PackLeaf& lf = L[op];
                    const uint32_t* sw = lf.data ? reinterpret_cast<const uint32_t*>(stage + (uint32_t)P.w_col_off[op] * 16u) : nullptr;
                    const WordsIO io{mwc + lane, first ? nullptr : mwc + lane, is_and ? 0u : 0xffffffffu, first ? 0u : (is_and ? 1u : 2u),
                                     lf.neg2 ? 0xffffffffu : 0u, wd};
                    eval_leaf_words<true>(env, lf, op, sw, lane, wr0, io);
                    first = false;
                }
                if (first)
                    for (uint32_t j = 0; j < wd; ++j) mwc[j * 32u + lane] = 0u;
            } else {
                // general tree: the words wait on the warp's stack in shared memory (lane-private columns)
                const uint32_t pstride = wd * 32u;
                uint32_t sp = 0;
                for (uint32_t i = 0; i < P.npost; ++i) {
                    const uint32_t op = P.postfix[i];
                    if (op < 0x80u) {
                        const PackLeaf& lf = L[op];
                        uint32_t* dst = stk + sp * pstride + lane;
                        const uint32_t* sw = lf.data ? reinterpret_cast<const uint32_t*>(stage + (uint32_t)P.w_col_off[op] * 16u) : nullptr;
                        const bool inv = lf.neg2 && !lf.fixmode;
                        const bool and_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFEu;
                        const bool or_next = sp >= 1u && i + 1u < P.npost && P.postfix[i + 1u] == 0xFFu;
                        const uint32_t* prev = stk + (sp ? sp - 1u : 0u) * pstride + lane;
                        const WordsIO io{dst, (and_next || or_next) ? prev : nullptr, or_next ? 0xffffffffu : 0u, 0u, inv ? 0xffffffffu : 0u, wd};
                        eval_leaf_words<true>(env, lf, op, sw, lane, wr0, io);
                        if (lf.fixmode) {   // ALP: correct the rows that are patches (1-bit stream in its own slot of the stage)
                            const uint32_t* fw = reinterpret_cast<const uint32_t*>(stage + (uint32_t)P.w_fix_off[op] * 16u);
                            __builtin_assume(__isShared(fw));
                            for (uint32_t j = 0; j < wd; ++j) {
                                const uint32_t fx = fw[j * 32u + lane];
                                uint32_t word = dst[j * 32u];
                                word = lf.fixmode == FIX_OR_PRED ? (word | fx) : (word & ~fx);
                                dst[j * 32u] = lf.neg2 ? ~word : word;
                            }
                        }
                        ++sp;
                    } else {
                        --sp;
                        uint32_t* x = stk + (sp - 1u) * pstride + lane;
                        const uint32_t* y = stk + sp * pstride + lane;
                        for (uint32_t j = 0; j < wd; ++j) x[j * 32u] = (op == 0xFEu) ? (x[j * 32u] & y[j * 32u]) : (x[j * 32u] | y[j * 32u]);
                    }
                }
                for (uint32_t j = 0; j < wd; ++j) mwc[j * 32u + lane] = stk[j * 32u + lane];
            }

            // ---- the stage is free: refill it (a value chunk of a dense tile that is waiting, else the leaf columns of the tile
            // `stages` ahead) while this tile is finished
            __syncwarp();
            --live;
            live += issue(s) ? 1u : 0u;

            // ---- outputs of the tile: tail masking (match_core.go semantics: tail bits zero; only the last tile of a pack has
            // a tail), bitset words (coalesced 128 B per warp and word), per-pack match count
            const bool tail_tile = pack_row0 + tile_rows > n;
            uint32_t tile_cnt = 0;
#pragma unroll 1
            for (uint32_t j = 0; j < wd; ++j) {
                uint32_t word = mwc[j * 32u + lane];
                const uint32_t wr = wr0 + j * 1024u;
                if (tail_tile) {
                    if (wr >= n) word = 0u;
                    else if (n - wr < 32u) word &= (1u << (n - wr)) - 1u;
                    mwc[j * 32u + lane] = word;
                }
                if (P.bitsets && wr < n) *reinterpret_cast<uint32_t*>(bits_base + (wr >> 3)) = word;
                tile_cnt += __popc(word);
            }
            lane_cnt += tile_cnt;

            if constexpr (AGG) {
                nmatch += tile_cnt;   // per-CTA totals only: any partition of the matches over threads will do
                const bool any_now = __any_sync(0xffffffffu, tile_cnt != 0u);
                bool dense_now = false;
                if (any_now) {
                    const uint32_t rows = min(tile_rows, n - pack_row0);
                    if (P.agg_dense_thr != 0xffffffffu) dense_now = P.agg_dense_thr == 0u || (uint64_t)__reduce_add_sync(0xffffffffu, tile_cnt) * P.agg_dense_thr > rows;
                    dense_now = dense_now && cur_rawmask != 0u;
                    if (dense_now) {
                        // the tile's raw value columns follow through the ring (from the next free stage on)
                        if (lane == 0) tmeta[tslot] = make_uint4(pack_row0, n, desc_sel, cur_rawmask);
                        const uint32_t cnt = (dqs >> 15) & 7u;
                        if (cnt == 0u) dqs = ((uint32_t)__ffs((int)cur_rawmask) - 1u) << 18;
                        dqs = (dqs & ~0x38000u) | (tslot << (3u * cnt)) | ((cnt + 1u) << 15);
                    } else if (tile_cnt != 0u && tile_cnt <= 8u * wd) {
                        // this tile's turn comes after the next filter: start pulling its matching rows towards L2 now
#pragma unroll
                        for (int j = 0; j < NA; ++j) {
                            if ((uint32_t)j >= na) break;
                            const ColView& v = AV[desc_sel * na + j];
                            if (v.kind != CK_BITS || v.width != 64) continue;
                            for (uint32_t jw = 0; jw < wd; ++jw) {
                                uint32_t word = mwc[jw * 32u + lane];
                                if (!word) continue;
                                const unsigned long long* vp = reinterpret_cast<const unsigned long long*>(v.data) + (wr0 + jw * 1024u);
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
                }
                if (have_prev && prev_any) reduce_prev();
                prev_row0 = pack_row0; prev_sel = desc_sel; prev_slot = tslot; have_prev = true; prev_any = any_now; prev_dense = dense_now;
            }
            if (++tslot == T) tslot = 0;
            __syncwarp();
            if (++s == S) s = 0;
        }
        if (cur_pack != END_PACK) flush_count(cur_pack);
        if constexpr (AGG) {
            if (have_prev && prev_any) reduce_prev();   // the value columns of the last tile
        }
    }

    if constexpr (AGG) {
        // ---- per-CTA partial aggregates: fixed-order tree inside the warp, then across warps
        __shared__ uint32_t sm_last;
        __shared__ AggPartial warp_part[WARP_MAX_WARPS];
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
                unsigned long long cbv = __shfl_down_sync(0xffffffffu, c, off);
                agg_merge(a, b, type);
                c += cbv;
            }
            if (lane == 0) { warp_acc[warp] = a; warp_cnt[warp] = c; }
            __syncthreads();
            if (threadIdx.x == 0) {
                AggAcc r = warp_acc[0];
                unsigned long long rc = warp_cnt[0];
                for (uint32_t q = 1; q < NWARPS; ++q) { agg_merge(r, warp_acc[q], type); rc += warp_cnt[q]; }
                AggPartial o;
                o.count = rc; o.valid = rc != 0; o.pad = 0;
                if (type == 9 || type == 10) { o.sum = r.s[0]; o.err = as_f64(r.s[1]); o.mn = r.s[2]; o.mx = r.s[3]; }
                else { o.sum = r.s[0]; o.err = 0.0; o.mn = r.s[1]; o.mx = r.s[2]; }
                P.partials[(size_t)blockIdx.x * na + j] = o;
            }
            __syncthreads();
        }

        // ---- the CTA that finishes last combines the per-CTA partials: every thread merges a contiguous run of them in
        // index order, a fixed shuffle tree and a fixed warp order do the rest — the topology depends on the grid size
        // only, so the result is bit-reproducible whichever CTA happens to be last (no separate launch, no serial walk)
        if (threadIdx.x == 0) {
            __threadfence();
            sm_last = atomicAdd(P.done, 1u) == gridDim.x - 1u;
        }
        __syncthreads();
        if (sm_last) {
            __threadfence();
            const uint32_t nparts = gridDim.x, T = THREADS, per = (nparts + T - 1u) / T;
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
                __syncthreads();
                if (threadIdx.x == 0) {
                    AggPartial f = warp_part[0];
                    for (uint32_t q = 1; q < NWARPS; ++q) partial_merge(f, warp_part[q], type);
                    P.agg_out[j] = f;
                }
                __syncthreads();
            }
        }
    }
}

cudaError_t launch_scan_warp(const ScanParams& P, int grid, size_t smem_bytes, cudaStream_t stream) {
    int variant, threads = 512;
    void (*kern)(const ScanParams);
    if (P.naggs == 0) { variant = 0; kern = scan_warp_kernel<0, 512>; }
    else if (P.naggs == 1) { variant = 1; kern = scan_warp_kernel<1, 512>; }
    else if (P.naggs == 2) { variant = 2; kern = scan_warp_kernel<2, 512>; }
    else { variant = 3; kern = scan_warp_kernel<4, 256>; threads = 256; }
    // function attributes are per device and sticky: set them once per (device, variant)
    static bool configured[64][4] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64 || !configured[dev][variant]) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WARP_MAX_DYN_SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) configured[dev][variant] = true;
    }
    kern<<<grid, threads, smem_bytes, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
