// kx_scan.cu — the fused decode + filter + reduce kernel of libknoxgpu (sm_100a).
//
// One persistent, warp-specialised kernel scans a whole batch of packs:
//   * a producer warp streams each tile's packed column bytes HBM → shared memory with
//     TMA bulk copies (cp.async.bulk … mbarrier::complete_tx) through a 4-stage ring;
//   * eight consumer warps unpack fields straight out of shared memory with funnel shifts,
//     evaluate every filter leaf as one wrap-around range test, build LSB-first bitset
//     words with __ballot_sync, combine leaves with the AND/OR program in registers,
//     popcount, and (optionally) reduce the matching rows of the value columns;
//   * decoded column vectors never exist in HBM — only bitset words, per-pack counts and
//     per-CTA partial aggregates are written.
//
// Bitset word layout: a little-endian 32-bit word of KnoxDB's bitset (row 8k+i ↔ bit i of
// byte k, internal/bitset/bitset.go:23-29) is exactly the __ballot_sync mask of 32
// consecutive rows.
//
// Replaces (reference, CPU): internal/encode/bitpack/cmp.go:20-130 + cmp_{eq,lt,le,bw}.go,
// internal/cmp/number.go:13-243, float.go:13-242, internal/bitset/generic/bitset.go
// (And/Or/PopCount), internal/operator/filter/match_core.go:44-215 and the reducers of
// internal/reducer/reducer.go:138-314.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"
#include "kx_kernels.h"
#include "kx_decode.cuh"
#include "kx_leaf.cuh"

namespace kx {

// Dictionary-set translation (DictionaryContainer.translateSet, int_dict.go:400-440) on the device: one thread per
// SET value binary-searches the pack's dictionary (sorted, unique, in T order: `flip` maps it to unsigned order) and
// sets the bit of the code it finds.  Runs on the scan stream right before scan_kernel; blockIdx.y = (pack, leaf) job.
__global__ void codeset_kernel(const CodesetJob* __restrict__ jobs, uint32_t njobs, const uint64_t* __restrict__ set_vals, uint32_t* __restrict__ out) {
  for (uint32_t jb = blockIdx.y; jb < njobs; jb += gridDim.y) {   // (gridDim.y is capped at 65535 jobs per launch)
    const CodesetJob J = jobs[jb];
    const unsigned long long* dict = reinterpret_cast<const unsigned long long*>(J.dict);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < J.nset; i += gridDim.x * blockDim.x) {
        const uint64_t val = __ldg(set_vals + J.set_off + i), key = val ^ J.flip;
        uint32_t lo = 0, hi = J.ndict;   // first dictionary entry >= val
        while (lo < hi) {
            uint32_t m = (lo + hi) >> 1;
            if ((__ldg(dict + m) ^ J.flip) < key) lo = m + 1; else hi = m;
        }
        if (lo < J.ndict && __ldg(dict + lo) == val && lo >= J.base) {   // the bitmap is indexed by the FIELD of the code stream: code - base
            const uint32_t f = lo - J.base;
            atomicOr(out + J.out_off + (f >> 5), 1u << (f & 31u));
        }
    }
  }
}

// ALP blocks: rows that are patches carry their true value outside the encoded stream.  One thread per patch
// evaluates the float predicate on it (the loops over `vals, pos` of float_alp.go:238-495) and sets the bit of
// its row in the leaf's correction stream (zeroed before the launch), which the scan ORs in / ANDs out.
__global__ void alpfix_kernel(const AlpFixJob* __restrict__ jobs, uint32_t njobs, uint8_t* __restrict__ out_base) {
  for (uint32_t jb = blockIdx.y; jb < njobs; jb += gridDim.y) {
    const AlpFixJob J = jobs[jb];
    const uint32_t* pos = reinterpret_cast<const uint32_t*>(J.blob);
    const double* vals = reinterpret_cast<const double*>(J.blob + alp_vals_off(J.np));
    uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
    const double a = __longlong_as_double((long long)J.a), b = __longlong_as_double((long long)J.b);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < J.np; k += gridDim.x * blockDim.x) {
        const double x = vals[k];
        bool p;
        switch (J.mode) {
        case 1: p = (a != a) ? (x != x) : (x == a); break;   // MatchEqual: a NaN operand matches NaN patches
        case 3: p = x > a; break;
        case 4: p = x >= a; break;
        case 5: p = x < a; break;
        case 6: p = x <= a; break;
        default: p = x >= a && x <= b; break;                  // MatchBetween
        }
        if (p != (J.invert != 0)) { uint32_t r = pos[k]; atomicOr(out + (r >> 5), 1u << (r & 31u)); }
    }
  }
}

// Run-end blocks: the predicate is evaluated once per RUN by this pre-pass (RunEndContainer.Match* +
// applyMatch, internal/encode/int_runend.go:224-318: match the run values, SetRange(start, end) per
// matching run); the scan kernel then streams the resulting per-leaf bitset like a 1-bit column.
__global__ void runfill_kernel(const RunFillJob* __restrict__ jobs, uint32_t njobs, const uint64_t* __restrict__ set_vals, uint8_t* __restrict__ out_base) {
  for (uint32_t jb = blockIdx.y; jb < njobs; jb += gridDim.y) {
    const RunFillJob J = jobs[jb];
    const uint32_t* ends = reinterpret_cast<const uint32_t*>(J.ends);
    const unsigned long long* vals = reinterpret_cast<const unsigned long long*>(J.vals);
    uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < J.nruns; k += gridDim.x * blockDim.x) {
        uint64_t val = __ldg(vals + k);
        bool p = J.is_set == 2u ? ((uint64_t)k - J.a) <= J.d   // affine run values: the closed form chose a range of runs
                 : J.is_set ? set_has(set_vals + J.a, (uint32_t)J.d, val) : ((val ^ J.wm) - J.a) <= J.d;
        if (!p) continue;
        uint32_t start = k ? __ldg(ends + k - 1) + 1u : 0u, end = __ldg(ends + k);   // inclusive
        if (end >= J.nrows) end = J.nrows - 1u;
        if (start > end) continue;
        uint32_t w0 = start >> 5, w1 = end >> 5;
        uint32_t m0 = 0xffffffffu << (start & 31u), m1 = 0xffffffffu >> (31u - (end & 31u));
        if (w0 == w1) { atomicOr(out + w0, m0 & m1); continue; }
        atomicOr(out + w0, m0);
        for (uint32_t w = w0 + 1; w < w1; ++w) out[w] = 0xffffffffu;   // words owned by this run alone
        atomicOr(out + w1, m1);
    }
  }
}

// Value pre-pass: decode at index + IEEE compare per row, one thread per row, one ballot per 32 rows → the leaf's 1-bit
// column (streamed by the scan like a run-end pre-pass result).  Used for ALP-RD blocks, whose reference matchers decode
// chunk by chunk and run the float compare kernels (float_alprd.go:181-211, cmp/float.go:13-242).
__global__ void valmatch_kernel(const ValJob* __restrict__ jobs, uint32_t njobs, uint8_t* __restrict__ out_base) {
    for (uint32_t jb = blockIdx.y; jb < njobs; jb += gridDim.y) {
        const ValJob& J = jobs[jb];
        const ColView v = J.view;
        uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
        const uint32_t nw = (v.n + 31u) >> 5;
        const bool f32 = v.type == 10;
        const double a = f32 ? (double)__uint_as_float((uint32_t)J.a) : __longlong_as_double((long long)J.a);
        const double b = f32 ? (double)__uint_as_float((uint32_t)J.b) : __longlong_as_double((long long)J.b);
        for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nw; w += (gridDim.x * blockDim.x) >> 5) {
            const uint32_t row = w * 32u + (threadIdx.x & 31u);
            bool p = false;
            if (row < v.n) {
                const uint64_t bits = decode_value(v, row, nullptr, 0);
                const double x = f32 ? (double)__uint_as_float((uint32_t)bits) : __longlong_as_double((long long)bits);   // float32 → float64 is exact
                switch (J.mode) {
                case 1: p = x == a; break;
                case 2: p = x != a; break;
                case 3: p = x > a; break;
                case 4: p = x >= a; break;
                case 5: p = x < a; break;
                case 6: p = x <= a; break;
                default: p = a <= x && x <= b; break;
                }
            }
            const uint32_t word = __ballot_sync(0xffffffffu, p);
            if ((threadIdx.x & 31u) == 0) out[w] = word;
        }
    }
}

// ------------------------------------------------------------------------------ the single-leaf kernel
// One leaf, no aggregates: the hot configuration (fused decode + compare + popcount) — the single-leaf kernel.
// Programs with several leaves, patch corrections or aggregates run scan_general_kernel (kx_general.cu).
// ONLY32 = every pack's leaf is a <= 32-bit packed range test (or all / none): a lean instantiation without the
// other leaf paths (small code footprint, fewer registers), MINB CTAs per SM.
template <bool ONLY32, int MINB>
__global__ void __launch_bounds__(SCAN_THREADS, MINB) scan_kernel(const ScanParams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty_bar = full_bar + MAX_STAGES;
    uint8_t* stage_base = smem + 128;

    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t R = P.R, tile_rows = R * 32u * CONSUMER_WARPS, nstages = P.stages;
    uint32_t* code_smem = reinterpret_cast<uint32_t*>(stage_base + (size_t)nstages * P.stage_bytes);   // LM_CODESET bitmaps of the current pack

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < nstages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CONSUMER_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // contiguous tile range of this CTA: consecutive tiles stay inside one pack (descriptor reuse,
    // sequential DRAM pages) and the split is static, so results are reproducible run to run
    const uint32_t t_begin = (uint32_t)(((uint64_t)blockIdx.x * P.ntiles) / gridDim.x);
    const uint32_t t_end = (uint32_t)(((uint64_t)(blockIdx.x + 1) * P.ntiles) / gridDim.x);
    if (t_begin >= t_end) return;

    uint32_t pack = pack_of_tile(P.packs, P.npacks, t_begin);
    PackInfo pi = P.packs[pack];
    uint32_t pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
    uint32_t chunk = t_begin - pi.tile0;
    auto next_tile = [&]() {   // advance (pack, chunk) to the following tile
        if (++chunk >= pack_tiles) {
            do { ++pack; pi = P.packs[pack]; } while (pi.n == 0);
            pack_tiles = (pi.n + tile_rows - 1) / tile_rows;
            chunk = 0;
        }
    };

    // Ring protocol: one stage per TILE (the single leaf's stream; tiles may hold several passes).
    if (warp == CONSUMER_WARPS) {
        // ===================== TMA producer (one elected lane) =====================
        if (lane == 0) {
            uint32_t s = 0, ph = 0;
            for (uint32_t t = t_begin; t < t_end; ++t) {
                const uint32_t rows = min(tile_rows, pi.n - chunk * tile_rows);
                const PackLeaf& L = P.leaves[pack];
                mbar_wait_relaxed(&empty_bar[s], ph ^ 1u);  // slot released by all consumer warps
                const uint32_t w = L.data ? L.width : 0u;   // (an unstaged leaf still cycles its stage: lockstep)
                const uint32_t bytes = w ? ((((rows * w + 7u) >> 3) + 15u) & ~15u) : 0u;
                mbar_expect_tx(&full_bar[s], bytes);        // arrive (count 1) + expected bytes
                if (bytes) tma_load_1d(stage_base + (size_t)s * P.stage_bytes, L.data + (size_t)chunk * (tile_rows / 8u) * w, bytes, &full_bar[s]);
                if (++s == nstages) { s = 0; ph ^= 1u; }
                if (t + 1 < t_end) next_tile();
            }
        }
        return;
    }

    // ===================== consumers: unpack + filter + popcount =====================
    uint32_t lane_cnt = 0;           // matches of the current pack seen by this lane
    uint32_t bm_pack = 0xffffffffu;  // pack whose code bitmap is cached in shared memory
    const uint32_t passes = (R + 31u) >> 5, Rp = min(R, 32u);

    auto flush_count = [&](uint32_t pk) {
        uint32_t c = __reduce_add_sync(0xffffffffu, lane_cnt);
        if (P.counts && lane == 0 && c) atomicAdd(P.counts + pk, (unsigned long long)c);
        lane_cnt = 0;
    };

    if constexpr (!ONLY32) {
        // hash-set leaf: prefilter bitmap (and a small exact table) are the same for every pack — copy them into shared
        // memory once per CTA
        if (P.pre_log2[0]) {
            const uint32_t npre = (1u << P.pre_log2[0]) >> 5;
            for (uint32_t i = threadIdx.x; i < npre; i += CONSUMER_WARPS * 32u) code_smem[P.hs_smem_off[0] + i] = __ldg(P.set_pre + P.pre_off[0] + i);
            if (P.hs_tab_smem_off[0] != 0xffffffffu) {
                const uint32_t nt = 8u << P.tab_log2[0];
                const uint32_t* src = reinterpret_cast<const uint32_t*>(P.set_tabs + P.tab_off[0]);
                for (uint32_t i = threadIdx.x; i < nt; i += CONSUMER_WARPS * 32u) code_smem[P.hs_tab_smem_off[0] + i] = __ldg(src + i);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
        }
    }

    uint32_t s = 0, ph = 0;
    for (uint32_t t = t_begin; t < t_end; ++t) {
        const PackLeaf& lf = P.leaves[pack];
        const uint32_t pack_row0 = chunk * tile_rows;          // first row of the tile within the pack

        if (!ONLY32 && P.code_bitmap_words && pack != bm_pack) {
            // new pack: the consumers copy its code bitmap (built by codeset_kernel) into shared memory
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));   // everybody is done with the previous pack's bitmap
            if (lf.mode == LM_CODESET) {
                const uint32_t nw = (((1u << lf.width) + (uint32_t)lf.wm + 31u) >> 5) + 1u;
                const uint32_t* src = P.code_bits + lf.a;
                uint32_t* dst = code_smem + P.code_smem_off[0];
                for (uint32_t i = threadIdx.x; i < nw; i += CONSUMER_WARPS * 32u) dst[i] = __ldg(src + i);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(CONSUMER_WARPS * 32));
            bm_pack = pack;
        }

        const uint32_t* sw = reinterpret_cast<const uint32_t*>(stage_base + (size_t)s * P.stage_bytes);
        // the leaf's operands are read ONCE per tile into registers: through the `lf` reference every pass re-reads them from
        // global memory (the bitset stores in between may alias as far as the compiler knows) — six dependent L1 round trips
        // per 32-row step, which is what bounded the narrow widths
        uint32_t t_mode = 0, t_w = 0, t_atop = 0, t_lim = 0, t_flip = 0;
        if constexpr (ONLY32) {
            t_mode = lf.mode; t_w = lf.width;
            const uint32_t k = 32u - t_w, a = (uint32_t)lf.a, d = (uint32_t)lf.d;
            t_atop = a << k; t_lim = (d << k) | ((1u << k) - 1u);   // k == 0: a, d   (leaf_range32's operands)
            t_flip = ((lf.neg != 0) != (lf.neg2 != 0)) ? 0xffffffffu : 0u;
        }
        PackLeaf lfv{};                                        // (the other leaf kinds: the whole descriptor, by value)
        if constexpr (!ONLY32) lfv = lf;
        const bool t_neg2 = ONLY32 ? false : lfv.neg2 != 0;
        mbar_wait(&full_bar[s], ph);                           // TMA bytes have landed
        const bool per_tile = ONLY32 || lfv.mode == LM_CODESET;
        if (per_tile) {
            // per-tile output state shared by the paths that dispatch once per TILE (the lean kernel; dictionary IN / NOT IN): the
            // per-pass output code knows whether the tile has a tail (only the last tile of a pack does)
            const bool own = lane < Rp;
            const uint32_t g_lane = warp * R + lane;               // this lane's group in pass 0 (pass p: + 32 p)
            const bool plain = own && pack_row0 + tile_rows <= pi.n && pack_row0 + tile_rows > pack_row0;   // every row of the lane's words exists
            uint8_t* bt = P.bitsets ? P.bitsets + pi.bitset_off + (size_t)(pack_row0 >> 3) + (size_t)g_lane * 4u : nullptr;
            const uint32_t n = pi.n;
            uint32_t flip = t_flip;
            if constexpr (!ONLY32) flip = ((lfv.neg != 0) != t_neg2) ? 0xffffffffu : 0u;
            auto emit = [&](uint32_t pass, uint32_t word) {
                word ^= flip;
                if (plain) {
                    if (bt) *reinterpret_cast<uint32_t*>(bt + (size_t)pass * 128u) = word;   // coalesced 128 B per warp
                } else {
                    // mask rows past the end of the pack (tail bits must be zero) and lanes that own no group
                    const uint64_t wr = (uint64_t)pack_row0 + (uint64_t)(g_lane + pass * 32u) * 32u;
                    uint32_t valid = 0;
                    if (own && wr < n) {
                        const uint32_t left = n - (uint32_t)wr;
                        valid = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
                    }
                    word &= valid;
                    if (bt && valid) *reinterpret_cast<uint32_t*>(bt + (size_t)pass * 128u) = word;
                }
                lane_cnt += __popc(word);
            };
            if constexpr (ONLY32) {
                if (t_mode == LM_RANGE32) {
                    const uint32_t* seg0 = sw + (size_t)g_lane * t_w;
                    if (t_atop) leaf_b32_passes<true>(seg0, lane, t_w, t_atop, t_lim, passes, own, emit);
                    else leaf_b32_passes<false>(seg0, lane, t_w, 0u, t_lim, passes, own, emit);
                } else {
                    const uint32_t cw = t_mode == LM_ALL ? 0xffffffffu : 0u;
                    for (uint32_t pass = 0; pass < passes; ++pass) emit(pass, cw);
                }
            } else {
                leaf_code32_passes(sw + (size_t)g_lane * lfv.width, lfv.width, code_smem + P.code_smem_off[0], passes, own, emit);
            }
            __syncwarp();                                          // all shared-memory reads of this stage are done
            if (lane == 0) mbar_arrive(&empty_bar[s]);
        } else {
        for (uint32_t pass = 0; pass < passes; ++pass) {
            const uint32_t g0 = warp * R + pass * 32u;         // first group (of the tile) of this pass
            const uint64_t wr = (uint64_t)pack_row0 + (uint64_t)(g0 + lane) * 32u;   // first pack row of this lane's word
            LeafEnv env{P, code_smem, pi.n, pack_row0};
            uint32_t word = eval_leaf(env, lfv, 0u, sw, g0, Rp, lane, wr, 0xffffffffu);
            if (t_neg2) word = ~word;   // float-level NOT of an ALP leaf without patches (with patches: general kernel)
            if (pass + 1 == passes) {    // all shared-memory reads of this stage are done: release it early
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty_bar[s]);
            }
            // mask rows past the end of the pack (tail bits must be zero) and lanes >= Rp
            uint32_t valid = 0;
            if (lane < Rp && wr < pi.n) {
                uint32_t left = pi.n - (uint32_t)wr;
                valid = left >= 32u ? 0xffffffffu : ((1u << left) - 1u);
            }
            word &= valid;
            // outputs: bitset words (coalesced 128 B per warp), per-pack match count
            if (P.bitsets && lane < Rp && wr < pi.n)
                *reinterpret_cast<uint32_t*>(P.bitsets + pi.bitset_off + (wr >> 3)) = word;
            lane_cnt += __popc(word);
        }
        }
        if (++s == nstages) { s = 0; ph ^= 1u; }

        if (t + 1 < t_end) {
            const uint32_t prev = pack;
            next_tile();
            if (pack != prev) flush_count(prev);
        }
    }
    flush_count(pack);
}


// ------------------------------------------------------------------------------ small kernels

// bitset.{And,AndNot,Or,Xor} with any/all flags: internal/bitset/generic/bitset.go:13-295
__global__ void bitset_op_kernel(uint32_t* dst, const uint32_t* src, uint64_t nbits, int op, unsigned int* flags /*[0]=any,[1]=notall*/) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t any = 0, notall = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t d = dst[i], s = src[i], r;
        switch (op) {
        case 0: r = d & s; break;
        case 1: r = d & ~s; break;
        case 2: r = d | s; break;
        default: r = d ^ s; break;
        }
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        r &= mask;
        dst[i] = r;
        any |= r;
        notall |= (r ^ mask);
    }
    if (__any_sync(0xffffffffu, any != 0) && (threadIdx.x & 31) == 0) atomicOr(&flags[0], 1u);
    if (__any_sync(0xffffffffu, notall != 0) && (threadIdx.x & 31) == 0) atomicOr(&flags[1], 1u);
}

__global__ void bitset_neg_kernel(uint32_t* buf, uint64_t nbits) {
    uint64_t nwords = (nbits + 31) >> 5;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        buf[i] = ~buf[i] & mask;
    }
}

__global__ void bitset_popcount_kernel(const uint32_t* buf, uint64_t nbits, unsigned long long* out) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t c = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nwords; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        c += __popc(buf[i] & mask);
    }
    c = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, (unsigned long long)c);
}

// Bitset.Indexes (internal/bitset/iterator.go:269-290): pass 1 = per-block popcounts,
// pass 2 (after an exclusive scan on the host side of the stream) = ordered scatter.
__global__ void bitset_block_counts_kernel(const uint32_t* buf, uint64_t nbits, uint32_t* block_counts) {
    // one block handles 256 consecutive words
    uint64_t nwords = (nbits + 31) >> 5;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t c = 0;
    if (i < nwords) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        c = __popc(buf[i] & mask);
    }
    __shared__ uint32_t ws[8];
    uint32_t s = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; ++q) t += ws[q]; block_counts[blockIdx.x] = t; }
}

__global__ void exclusive_scan_kernel(uint32_t* v, uint32_t n, unsigned long long* total) {
    // single-block scan (n = #256-word blocks; ≤ 2^32/8192 entries)
    __shared__ uint32_t carry;
    __shared__ uint32_t ws[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t x = i < n ? v[i] : 0, incl = x;
        for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint32_t w = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0, wi = w;
            for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, wi, off); if (threadIdx.x >= (uint32_t)off) wi += y; }
            ws[threadIdx.x] = wi - w;
        }
        __syncthreads();
        uint32_t excl = carry + ws[threadIdx.x >> 5] + incl - x;
        if (i < n) v[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// ------------------------------------------------------------------------------ Simple8b (legacy container 6) on the device
// s8b.Decode (internal/encode/s8b/generic/decode.go:15-81) as a two-pass transcode at registration: pass 1 reads every
// 64-bit codeword's selector (top 4 bits; {128,128,60,30,20,15,12,10,8,7,6,5,4,3,2,1} values of
// {0,0,1,2,3,4,5,6,7,8,10,12,15,20,30,60} bits, encode.go:32-42) and records its value count and the widest value it holds;
// an exclusive scan turns the counts into row offsets; pass 2 writes every value into the fixed-width LSB-first bit stream
// the scan kernels read (64-bit atomic ORs into a zeroed buffer: neighbouring codewords share output words).
__device__ __constant__ uint8_t S8B_COUNT[16] = {128, 128, 60, 30, 20, 15, 12, 10, 8, 7, 6, 5, 4, 3, 2, 1};
__device__ __constant__ uint8_t S8B_BITS[16] = {0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 15, 20, 30, 60};

__global__ void s8b_count_kernel(const unsigned long long* __restrict__ words, uint32_t nwords, uint32_t* __restrict__ counts, uint32_t* __restrict__ maxbits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t mb = 0;
    if (i < nwords) {
        const unsigned long long w = words[i];
        const uint32_t sel = (uint32_t)(w >> 60), cnt = S8B_COUNT[sel], bits = S8B_BITS[sel];
        counts[i] = cnt;
        if (sel == 1u) mb = 1u;
        else if (sel > 1u) {
            unsigned long long any = 0;   // OR of the fields: its bit length is the widest value of the word
            const unsigned long long m = bits >= 64u ? ~0ull : ((1ull << bits) - 1ull);
            for (uint32_t q = 0; q < cnt; ++q) any |= (w >> (q * bits)) & m;
            mb = any ? 64u - (uint32_t)__clzll((long long)any) : 0u;
        }
    }
    mb = __reduce_max_sync(0xffffffffu, mb);
    if ((threadIdx.x & 31u) == 0u && mb) atomicMax(maxbits, mb);
}

__global__ void s8b_pack_kernel(const unsigned long long* __restrict__ words, uint32_t nwords, const uint32_t* __restrict__ offs, uint32_t nrows, uint32_t width,
                                unsigned long long* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const unsigned long long w = words[i];
    const uint32_t sel = (uint32_t)(w >> 60), cnt = S8B_COUNT[sel], bits = S8B_BITS[sel];
    const unsigned long long m = bits >= 64u ? ~0ull : ((1ull << bits) - 1ull);
    uint32_t row = offs[i];
    for (uint32_t q = 0; q < cnt && row < nrows; ++q, ++row) {
        const unsigned long long f = sel == 0u ? 0ull : (sel == 1u ? 1ull : ((w >> (q * bits)) & m));
        if (!f) continue;
        const unsigned long long bit = (unsigned long long)row * width;
        const uint32_t sh = (uint32_t)(bit & 63ull);
        atomicOr(out + (bit >> 6), f << sh);
        if (sh + width > 64u) atomicOr(out + (bit >> 6) + 1, f >> (64u - sh));
    }
}

__global__ void bitset_scatter_kernel(const uint32_t* buf, uint64_t nbits, const uint32_t* block_offs, uint32_t* dst) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t wv = 0;
    if (i < nwords) {
        uint32_t mask = 0xffffffffu;
        if (i == nwords - 1 && (nbits & 31)) mask = (1u << (nbits & 31)) - 1u;
        wv = buf[i] & mask;
    }
    uint32_t c = __popc(wv), incl = c;
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t q = 0; q < (threadIdx.x >> 5); ++q) woff += ws[q];
    uint32_t pos = block_offs[blockIdx.x] + woff + incl - c;
    uint32_t base = (uint32_t)(i << 5);
    while (wv) { uint32_t b = __ffs(wv) - 1; dst[pos++] = base + b; wv &= wv - 1; }
}

// ---- Bitset.Indexes for a whole batch of packs (reader.go:432-436: sel := bits.Indexes(hits) per pack).
// The packs' bitsets sit at ascending offsets of one device buffer; a thread owns one 32-bit word of that
// buffer, finds its pack by binary search over the offsets and ignores words that lie in the gaps.
__device__ __forceinline__ uint32_t sel_word(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t w,
                                             uint64_t total_words, uint32_t* local_word) {
    if (w >= total_words) return 0;
    const uint64_t byte = w * 4;
    uint32_t lo = 0, hi = npacks;   // last pack with bitset_off <= byte
    while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (packs[m].bitset_off <= byte) lo = m; else hi = m; }
    const uint64_t lw = (byte - packs[lo].bitset_off) >> 2;
    if (lw >= ((uint64_t)packs[lo].n + 31) >> 5) return 0;   // gap between two packs
    *local_word = (uint32_t)lw;
    return *reinterpret_cast<const uint32_t*>(bits + byte);   // tail bits past n are already zero
}

__global__ void select_counts_kernel(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t total_words,
                                     uint32_t* __restrict__ block_counts) {
    uint32_t lw;
    uint32_t c = __popc(sel_word(packs, npacks, bits, (uint64_t)blockIdx.x * 256 + threadIdx.x, total_words, &lw));
    __shared__ uint32_t ws[8];
    uint32_t s = __reduce_add_sync(0xffffffffu, c);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) { uint32_t t = 0; for (int q = 0; q < 8; ++q) t += ws[q]; block_counts[blockIdx.x] = t; }
}

__global__ void select_scatter_kernel(const PackInfo* __restrict__ packs, uint32_t npacks, const uint8_t* __restrict__ bits, uint64_t total_words,
                                      const uint32_t* __restrict__ block_offs, uint32_t* __restrict__ dst) {
    uint32_t lw = 0;
    uint32_t wv = sel_word(packs, npacks, bits, (uint64_t)blockIdx.x * 256 + threadIdx.x, total_words, &lw);
    uint32_t c = __popc(wv), incl = c;
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, incl, off); if ((threadIdx.x & 31) >= (uint32_t)off) incl += y; }
    __shared__ uint32_t ws[8];
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t q = 0; q < (threadIdx.x >> 5); ++q) woff += ws[q];
    uint32_t pos = block_offs[blockIdx.x] + woff + incl - c;
    const uint32_t base = lw << 5;   // row id relative to the pack
    while (wv) { uint32_t b = __ffs(wv) - 1; dst[pos++] = base + b; wv &= wv - 1; }
}

// NumberContainer.AppendTo(dst, sel) for a batch of packs (internal/encode/int_*.go AppendTo with a selection;
// query/result.go:196-264 copies the selected rows): one thread per selected row, decode at index.
__global__ void gather_kernel(const ColView* __restrict__ views, const unsigned long long* __restrict__ sel_off, uint32_t npacks,
                              const uint32_t* __restrict__ sel, uint64_t total, int elem_bytes, uint8_t* __restrict__ dst) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = npacks;   // last pack with sel_off <= i
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sel_off[m] <= i) lo = m; else hi = m; }
        uint64_t x = decode_value(views[lo], sel[i], nullptr, 0);
        switch (elem_bytes) {
        case 8: reinterpret_cast<uint64_t*>(dst)[i] = x; break;
        case 4: reinterpret_cast<uint32_t*>(dst)[i] = (uint32_t)x; break;
        case 2: reinterpret_cast<uint16_t*>(dst)[i] = (uint16_t)x; break;
        default: dst[i] = (uint8_t)x; break;
        }
    }
}

// NumberContainer.AppendTo(dst, nil) / bitpack.Decode: one thread per row
__global__ void decode_kernel(ColView v, uint8_t* dst) {
    int nb = type_bits(v.type) / 8;
    for (uint64_t row = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; row < v.n; row += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t x = decode_value(v, (uint32_t)row, nullptr, 0);
        switch (nb) {
        case 8: reinterpret_cast<uint64_t*>(dst)[row] = x; break;
        case 4: reinterpret_cast<uint32_t*>(dst)[row] = (uint32_t)x; break;
        case 2: reinterpret_cast<uint16_t*>(dst)[row] = (uint16_t)x; break;
        default: dst[row] = (uint8_t)x; break;
        }
    }
}

// Zone-map + bloom pruning: one thread per pack (stats.matchVector, internal/pack/stats/match.go:92-195)
__global__ void prune_kernel(PruneParams P) {
    uint32_t pack = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (pack < P.npacks) {
        uint32_t stack = 0; int sp = 0;   // bit stack (depth <= 32)
        for (uint32_t i = 0; i < P.npost; ++i) {
            uint32_t op = P.postfix[i];
            if (op < 0x80u) {
                const PruneLeaf& L = P.leaves[op];
                uint64_t mn = P.mins[(size_t)pack * P.nleaves + op], mx = P.maxs[(size_t)pack * P.nleaves + op];
                uint64_t f = L.flip;
                uint64_t kmn = mn ^ f, kmx = mx ^ f, ka = L.a ^ f, kb = L.b ^ f;
                bool m = true;
                if (L.is_float) {
                    double dmn = __longlong_as_double((long long)mn), dmx = __longlong_as_double((long long)mx);
                    double da = __longlong_as_double((long long)L.a), db = __longlong_as_double((long long)L.b);
                    switch (L.mode) {
                    case 1: m = dmn <= da && dmx >= da; break;
                    case 3: m = dmx > da; break;
                    case 4: m = dmx >= da; break;
                    case 5: m = dmn < da; break;
                    case 6: m = dmn <= da; break;
                    case 9: m = dmn <= db && dmx >= da; break;
                    default: m = true;
                    }
                } else {
                    switch (L.mode) {   // MatchRangeVectors, internal/operator/filter/match_num.go
                    case 1: m = kmn <= ka && kmx >= ka; break;            // EQ  :357-371
                    case 3: m = kmx > ka; break;                           // GT  :429-434
                    case 4: m = kmx >= ka; break;                          // GE  :460-465
                    case 5: m = kmn < ka; break;                           // LT  :491-496
                    case 6: m = kmn <= ka; break;                          // LE  :522-527
                    case 9: m = kmn <= kb && kmx >= ka; break;             // RG  :573-588
                    case 7: {                                              // IN  :693-735 set.ContainsRange(min,max) in uint64 order
                        uint64_t lo = mn, hi = mx;
                        if (lo > hi) { uint64_t t2 = lo; lo = hi; hi = t2; }
                        const uint64_t* s = P.set_vals + L.set_off;
                        uint32_t l2 = 0, h2 = L.nset;
                        while (l2 < h2) { uint32_t mid = (l2 + h2) >> 1; if (s[mid] < lo) l2 = mid + 1; else h2 = mid; }
                        m = l2 < L.nset && s[l2] <= hi;
                        break;
                    }
                    default: m = true;                                     // NE :394-401, NIN :810-817 (undecided)
                    }
                }
                // bloom probe for EQ / IN when the pack carries a filter (match.go:141-192)
                if (m && P.blooms && (L.mode == 1 || L.mode == 7)) {
                    const uint8_t* bf = P.blooms[(size_t)pack * P.nleaves + op];
                    if (bf) {
                        uint32_t mbits = (uint32_t)((P.bloom_len[(size_t)pack * P.nleaves + op] - 1) * 8);
                        uint32_t mask = mbits - 1u, kk = bf[0];
                        const uint8_t* bits = bf + 1;
                        bool anyhit = false;
                        for (uint32_t h = P.hash_off[op]; h < P.hash_off[op + 1] && !anyhit; ++h) {
                            uint64_t hv = P.hashes[h];
                            uint32_t h0 = (uint32_t)hv, h1 = (uint32_t)(hv >> 32);
                            bool hit = true;
                            for (uint32_t q = 0; q < kk && hit; ++q) {      // bloom.go:136-150
                                hit = (bits[(h0 & mask) >> 3] >> (h0 & 7u)) & 1u;
                                h0 += h1;
                            }
                            anyhit = hit;
                        }
                        if (P.hash_off[op + 1] > P.hash_off[op]) m = anyhit;
                    }
                }
                stack = (stack << 1) | (m ? 1u : 0u); ++sp;
            } else {
                uint32_t y = stack & 1u; stack >>= 1; --sp;
                uint32_t x = stack & 1u;
                stack = (stack & ~1u) | (op == 0xFEu ? (x & y) : (x | y));
            }
        }
        alive = (stack & 1u) != 0;
    }
    uint32_t b = __ballot_sync(0xffffffffu, alive);
    if ((threadIdx.x & 31) == 0 && pack < P.npacks) P.out[pack >> 5] = b;
    uint32_t c = __popc(b);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(P.count, (unsigned long long)c);
}

// ------------------------------------------------------------------------------ launchers
static int grid_for(uint64_t items, uint64_t cap) { uint64_t g = (items + 255) / 256; return (int)(g < cap ? g : cap); }

cudaError_t launch_scan(const ScanParams& P, int grid, size_t smem_bytes, bool only32, int ctas_per_sm, cudaStream_t stream) {
    int variant = !only32 ? 0 : (ctas_per_sm >= 3 ? 2 : 1);
    void (*kern)(const ScanParams) = scan_kernel<false, 2>;
    if (variant == 1) kern = scan_kernel<true, 2>;
    if (variant == 2) kern = scan_kernel<true, 3>;
    // function attributes are per device and sticky: set them once per (device, variant)
    static bool configured[64][3] = {};
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

cudaError_t launch_valmatch(const ValJob* jobs, uint32_t njobs, uint32_t max_rows, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_rows == 0) return cudaSuccess;
    uint32_t gx = (max_rows + 255u) / 256u;
    if (gx > 148u * 8u) gx = 148u * 8u;
    valmatch_kernel<<<dim3(gx, njobs < 65535u ? njobs : 65535u), 256, 0, stream>>>(jobs, njobs, out_base);
    return cudaGetLastError();
}

cudaError_t launch_alpfix(const AlpFixJob* jobs, uint32_t njobs, uint32_t max_patches, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_patches == 0) return cudaSuccess;
    uint32_t gx = (max_patches + 255u) / 256u;
    if (gx > 148u) gx = 148u;
    alpfix_kernel<<<dim3(gx, njobs < 65535u ? njobs : 65535u), 256, 0, stream>>>(jobs, njobs, out_base);
    return cudaGetLastError();
}

cudaError_t launch_runfill(const RunFillJob* jobs, uint32_t njobs, uint32_t max_runs, const uint64_t* set_vals, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_runs == 0) return cudaSuccess;
    uint32_t gx = (max_runs + 255u) / 256u;
    if (gx > 148u * 4u) gx = 148u * 4u;
    runfill_kernel<<<dim3(gx, njobs < 65535u ? njobs : 65535u), 256, 0, stream>>>(jobs, njobs, set_vals, out_base);
    return cudaGetLastError();
}

cudaError_t launch_codeset(const CodesetJob* jobs, uint32_t njobs, uint32_t max_set, const uint64_t* set_vals, uint32_t* out, cudaStream_t stream) {
    if (njobs == 0 || max_set == 0) return cudaSuccess;
    uint32_t gx = (max_set + 127u) / 128u;
    if (gx > 32u) gx = 32u;
    codeset_kernel<<<dim3(gx, njobs < 65535u ? njobs : 65535u), 128, 0, stream>>>(jobs, njobs, set_vals, out);
    return cudaGetLastError();
}

cudaError_t launch_bitset_op(uint32_t* dst, const uint32_t* src, uint64_t nbits, int op, unsigned int* flags, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_op_kernel<<<grid ? grid : 1, 256, 0, stream>>>(dst, src, nbits, op, flags);
    return cudaGetLastError();
}
cudaError_t launch_bitset_neg(uint32_t* buf, uint64_t nbits, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_neg_kernel<<<grid ? grid : 1, 256, 0, stream>>>(buf, nbits);
    return cudaGetLastError();
}
cudaError_t launch_bitset_popcount(const uint32_t* buf, uint64_t nbits, unsigned long long* out, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    int grid = grid_for(nwords, 148 * 8);
    bitset_popcount_kernel<<<grid ? grid : 1, 256, 0, stream>>>(buf, nbits, out);
    return cudaGetLastError();
}
cudaError_t launch_bitset_indexes(const uint32_t* buf, uint64_t nbits, uint32_t* block_tmp, unsigned long long* total, uint32_t* dst, cudaStream_t stream) {
    uint64_t nwords = (nbits + 31) >> 5;
    uint32_t nblocks = (uint32_t)((nwords + 255) / 256);
    if (nblocks == 0) return cudaMemsetAsync(total, 0, 8, stream);
    bitset_block_counts_kernel<<<nblocks, 256, 0, stream>>>(buf, nbits, block_tmp);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(block_tmp, nblocks, total);
    bitset_scatter_kernel<<<nblocks, 256, 0, stream>>>(buf, nbits, block_tmp, dst);
    return cudaGetLastError();
}
cudaError_t launch_select(const PackInfo* packs, uint32_t npacks, const uint8_t* bits, uint64_t total_words, uint32_t* block_tmp,
                          unsigned long long* total, uint32_t* dst, cudaStream_t stream) {
    uint32_t nblocks = (uint32_t)((total_words + 255) / 256);
    if (nblocks == 0) return cudaMemsetAsync(total, 0, 8, stream);
    select_counts_kernel<<<nblocks, 256, 0, stream>>>(packs, npacks, bits, total_words, block_tmp);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(block_tmp, nblocks, total);
    select_scatter_kernel<<<nblocks, 256, 0, stream>>>(packs, npacks, bits, total_words, block_tmp, dst);
    return cudaGetLastError();
}
cudaError_t launch_gather(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                          int elem_bytes, void* dst, cudaStream_t stream) {
    if (total == 0) return cudaSuccess;
    int grid = grid_for(total, 148 * 16);
    gather_kernel<<<grid ? grid : 1, 256, 0, stream>>>(views, sel_off, npacks, sel, total, elem_bytes, reinterpret_cast<uint8_t*>(dst));
    return cudaGetLastError();
}
cudaError_t launch_s8b_count(const void* words, uint32_t nwords, uint32_t* counts, uint32_t* maxbits, unsigned long long* total, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(maxbits, 0, 4, stream);
    if (e != cudaSuccess) return e;
    if (nwords == 0) return cudaMemsetAsync(total, 0, 8, stream);
    s8b_count_kernel<<<(nwords + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(words), nwords, counts, maxbits);
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(counts, nwords, total);
    return cudaGetLastError();
}
cudaError_t launch_s8b_pack(const void* words, uint32_t nwords, const uint32_t* offs, uint32_t nrows, uint32_t width, void* out, cudaStream_t stream) {
    if (nwords == 0 || width == 0) return cudaSuccess;
    s8b_pack_kernel<<<(nwords + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(words), nwords, offs, nrows, width,
                                                            reinterpret_cast<unsigned long long*>(out));
    return cudaGetLastError();
}
cudaError_t launch_exclusive_scan(uint32_t* v, uint32_t n, unsigned long long* total, cudaStream_t stream) {
    exclusive_scan_kernel<<<1, 1024, 0, stream>>>(v, n, total);
    return cudaGetLastError();
}
cudaError_t launch_decode(const ColView& v, void* dst, cudaStream_t stream) {
    int grid = grid_for(v.n, 148 * 16);
    decode_kernel<<<grid ? grid : 1, 256, 0, stream>>>(v, reinterpret_cast<uint8_t*>(dst));
    return cudaGetLastError();
}
cudaError_t launch_prune(const PruneParams& P, cudaStream_t stream) {
    int grid = (int)((P.npacks + 255) / 256);
    prune_kernel<<<grid ? grid : 1, 256, 0, stream>>>(P);
    return cudaGetLastError();
}

}  // namespace kx
