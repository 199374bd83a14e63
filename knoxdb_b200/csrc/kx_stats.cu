// kx_stats.cu — kernels over the device-resident statistics index (kx_stats): zone-map + bloom pruning
// of candidate packs and bloom-filter construction from column values.
//
// Replaces (reference, CPU): stats.matchVector / matchFilterVector → Matcher.MatchRangeVectors on the
// min/max columns of a statistics pack + bloom.Filter.Contains per surviving pack
// (internal/pack/stats/match.go:92-195, internal/operator/filter/match_num.go MatchRangeVectors,
// internal/filter/bloom/bloom.go:136-150,182-184) and stats.BuildBloomFilter → hash.Vec64/Vec32/… +
// bloom.Filter.Add (internal/pack/stats/filter.go:296-367, internal/filter/bloom/bloom.go:168-199).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_kernels.h"
#include "kx_types.h"
#include "kx_xxh3.h"

namespace kx {

// zone-map test of one leaf: MatchRange / MatchRangeVectors semantics (match_num.go), operands and
// statistics as 64-bit patterns of T; `flip` makes unsigned order equal T order
__device__ __forceinline__ bool zone_match(const PruneLeaf& L, uint64_t mn, uint64_t mx, const uint64_t* __restrict__ set_vals) {
    if (L.is_float) {
        double dmn = __longlong_as_double((long long)mn), dmx = __longlong_as_double((long long)mx);
        double da = __longlong_as_double((long long)L.a), db = __longlong_as_double((long long)L.b);
        switch (L.mode) {
        case 1: return dmn <= da && dmx >= da;
        case 3: return dmx > da;
        case 4: return dmx >= da;
        case 5: return dmn < da;
        case 6: return dmn <= da;
        case 9: return dmn <= db && dmx >= da;
        }
        return true;
    }
    const uint64_t f = L.flip, kmn = mn ^ f, kmx = mx ^ f, ka = L.a ^ f, kb = L.b ^ f;
    switch (L.mode) {
    case 1: return kmn <= ka && kmx >= ka;            // EQ  match_num.go:357-371
    case 3: return kmx > ka;                           // GT  :429-434
    case 4: return kmx >= ka;                          // GE  :460-465
    case 5: return kmn < ka;                           // LT  :491-496
    case 6: return kmn <= ka;                          // LE  :522-527
    case 9: return kmn <= kb && kmx >= ka;             // RG  :573-588
    case 7: {                                          // IN  :693-735 set.ContainsRange(min, max) in uint64 order
        uint64_t lo = mn, hi = mx;
        if (lo > hi) { uint64_t t = lo; lo = hi; hi = t; }
        const uint64_t* s = set_vals + L.set_off;
        uint32_t l2 = 0, h2 = L.nset;
        while (l2 < h2) { uint32_t mid = (l2 + h2) >> 1; if (s[mid] < lo) l2 = mid + 1; else h2 = mid; }
        return l2 < L.nset && s[l2] <= hi;
    }
    }
    return true;                                       // NE :394-401, NIN :810-817: undecided → keep
}

// bloom.Filter.Contains: bit (h0 + i*h1) & mask for i < k; the k loads are independent (no early exit)
__device__ __forceinline__ bool bloom_contains(const uint8_t* __restrict__ bits, uint32_t mask, uint32_t k, uint64_t hv) {
    uint32_t h0 = (uint32_t)hv, h1 = (uint32_t)(hv >> 32);
    if (k == 4) {
        uint32_t l0 = h0 & mask, l1 = (h0 + h1) & mask, l2 = (h0 + 2u * h1) & mask, l3 = (h0 + 3u * h1) & mask;
        uint32_t b0 = bits[l0 >> 3], b1 = bits[l1 >> 3], b2 = bits[l2 >> 3], b3 = bits[l3 >> 3];
        return ((b0 >> (l0 & 7u)) & (b1 >> (l1 & 7u)) & (b2 >> (l2 & 7u)) & (b3 >> (l3 & 7u)) & 1u) != 0;
    }
    bool hit = true;
    for (uint32_t q = 0; q < k; ++q) { uint32_t l = (h0 + q * h1) & mask; hit = hit && ((bits[l >> 3] >> (l & 7u)) & 1u); }
    return hit;
}

// one thread per data pack over the resident index (column-major: stat[field][pack] → coalesced)
__global__ void prune_stats_kernel(PruneStatsParams P) {
    uint32_t pack = blockIdx.x * blockDim.x + threadIdx.x;
    bool alive = false;
    if (pack < P.npacks) {
        uint32_t stack = 0;
        for (uint32_t i = 0; i < P.npost; ++i) {
            uint32_t op = P.postfix[i];
            if (op < 0x80u) {
                const PruneLeaf& L = P.leaves[op];
                const size_t cell = (size_t)P.leaf_field[op] * P.npacks + pack;
                bool m = true;
                if (!P.leaf_nozone[op]) m = zone_match(L, P.mins[cell], P.maxs[cell], P.set_vals);
                // filters only serve EQ / IN (match.go: filterType), and only packs that carry one
                if (m && (L.mode == 1 || L.mode == 7) && P.hash_off[op + 1] > P.hash_off[op]) {
                    const uint8_t* bits = reinterpret_cast<const uint8_t*>(P.bloom_ptr[cell]);
                    if (bits) {
                        const uint32_t mask = P.bloom_mask[cell], k = P.bloom_k[cell];
                        bool anyhit = false;
                        for (uint32_t h = P.hash_off[op]; h < P.hash_off[op + 1]; ++h) anyhit = anyhit || bloom_contains(bits, mask, k, P.hashes[h]);
                        m = anyhit;
                    }
                }
                stack = (stack << 1) | (m ? 1u : 0u);
            } else {
                uint32_t y = stack & 1u; stack >>= 1;
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

// BuildBloomFilter: hash every value (fixed width: hash.Vec64/32/16/8 = XXH3 of the value's LE bytes;
// byte strings: hash.Hash) and set its k bits.  bits: m/8 bytes, 4-byte aligned, zeroed by the caller.
__global__ void bloom_build_kernel(const uint8_t* __restrict__ values, const uint32_t* __restrict__ offsets, uint64_t n, int elem_bytes,
                                   uint32_t* __restrict__ bits, uint32_t mask, uint32_t k) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t hv;
        switch (elem_bytes) {
        case 8: hv = xxh3::u64(reinterpret_cast<const unsigned long long*>(values)[i]); break;
        case 4: hv = xxh3::u32(reinterpret_cast<const uint32_t*>(values)[i]); break;
        case 2: hv = xxh3::u16(reinterpret_cast<const uint16_t*>(values)[i]); break;
        case 1: hv = xxh3::u8(values[i]); break;
        default: { uint32_t o0 = offsets[i], o1 = offsets[i + 1]; hv = xxh3::bytes(values + o0, o1 - o0); break; }
        }
        uint32_t h0 = (uint32_t)hv, h1 = (uint32_t)(hv >> 32);
        for (uint32_t q = 0; q < k; ++q) {
            uint32_t l = h0 & mask;
            atomicOr(bits + (l >> 5), 1u << (l & 31u));   // little-endian words: byte l>>3, bit l&7
            h0 += h1;
        }
    }
}

cudaError_t launch_prune_stats(const PruneStatsParams& P, cudaStream_t stream) {
    int grid = (int)((P.npacks + 127) / 128);
    prune_stats_kernel<<<grid ? grid : 1, 128, 0, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t launch_bloom_build(const uint8_t* values, const uint32_t* offsets, uint64_t n, int elem_bytes, uint32_t* bits, uint32_t mask,
                               uint32_t k, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    uint64_t g = (n + 255) / 256;
    int grid = (int)(g < 148 * 8 ? g : 148 * 8);
    bloom_build_kernel<<<grid, 256, 0, stream>>>(values, offsets, n, elem_bytes, bits, mask, k);
    return cudaGetLastError();
}

}  // namespace kx
