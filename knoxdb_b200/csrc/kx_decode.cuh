// kx_decode.cuh — decode-at-index of one row of a column block (device code shared by the scan kernel, the
// gather / decode kernels and the time-bucketed reduce).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"

namespace kx {

// ------------------------------------------------------------------------------ field access
// `w`-bit field starting at bit offset `bit` of a little-endian bit string (32-bit words).
__device__ __forceinline__ uint64_t load_field(const uint32_t* __restrict__ words, uint64_t bit, uint32_t w) {
    uint64_t idx = bit >> 5;
    uint32_t sh = (uint32_t)bit & 31u;
    uint32_t w0 = words[idx], w1 = words[idx + 1];
    uint32_t lo = __funnelshift_r(w0, w1, sh);
    if (w <= 32) return lo & (uint32_t)width_mask((int)w);
    uint32_t w2 = words[idx + 2];
    uint32_t hi = __funnelshift_r(w1, w2, sh);
    return (((uint64_t)hi << 32) | lo) & width_mask((int)w);
}

__device__ __forceinline__ uint32_t run_of_row(const uint32_t* __restrict__ ends, uint32_t nruns, uint32_t row) {
    uint32_t lo = 0, hi = nruns;   // first run with ends[k] >= row (ends are inclusive)
    while (lo < hi) {
        uint32_t m = (lo + hi) >> 1;
        if (__ldg(ends + m) >= row) hi = m; else lo = m + 1;
    }
    return lo;
}

// ALP powers of ten, the exact literals of internal/encode/alp/constants.go:88-150
static __device__ const double ALP_F10[24] = {
    1.0, 10.0, 100.0, 1000.0, 10000.0, 100000.0, 1000000.0, 10000000.0, 100000000.0, 1000000000.0, 10000000000.0,
    100000000000.0, 1000000000000.0, 10000000000000.0, 100000000000000.0, 1000000000000000.0, 10000000000000000.0,
    100000000000000000.0, 1000000000000000000.0, 10000000000000000000.0, 100000000000000000000.0,
    1000000000000000000000.0, 10000000000000000000000.0, 100000000000000000000000.0};
static __device__ const double ALP_IF10[21] = {
    1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001, 0.00000001, 0.000000001, 0.0000000001, 0.00000000001,
    0.000000000001, 0.0000000000001, 0.00000000000001, 0.000000000000001, 0.0000000000000001, 0.00000000000000001,
    0.000000000000000001, 0.0000000000000000001, 0.00000000000000000001};

// value of row `row` of a block as the sign-/zero-extended 64-bit pattern of T (IEEE bits for
// floats).  `staged`: the tile's bit stream in shared memory (row index relative to the tile),
// or nullptr to read the block's stream from global memory.
__device__ __forceinline__ uint64_t decode_value(const ColView& v, uint32_t row, const uint32_t* staged, uint32_t row_in_tile) {
    switch (v.kind) {
    case CK_CONST: return v.base;
    case CK_DELTA: return type_ext(v.type, (uint64_t)row * v.delta + v.base);
    case CK_BITS: {
        uint64_t f;
        if (staged) f = load_field(staged, (uint64_t)row_in_tile * v.width, v.width);
        else if (v.width == 64) f = __ldg(reinterpret_cast<const unsigned long long*>(v.data) + row);
        else f = load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width);
        if (type_is_float(v.type)) return f;
        return type_ext(v.type, f + v.base);
    }
    case CK_DICT: {
        uint64_t code = staged ? load_field(staged, (uint64_t)row_in_tile * v.width, v.width)
                               : load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width);
        code += v.delta;
        if (code >= v.naux) code = v.naux ? v.naux - 1u : 0u;   // (a corrupt code stream must not read past the dictionary)
        return __ldg(reinterpret_cast<const unsigned long long*>(v.aux) + code);
    }
    case CK_RUNEND: {
        uint32_t k = run_of_row(reinterpret_cast<const uint32_t*>(v.aux), v.naux, row);
        return __ldg(reinterpret_cast<const unsigned long long*>(v.data) + k);
    }
    case CK_ALPRD: {   // DecoderRD.DecodeValue (internal/encode/alp/rd.go:212-219): bits = uint64(left) << shift | right
        const uint32_t lw = v.naux & 0xffu, shift = (v.naux >> 16) & 0xffu;
        uint64_t l = (lw ? load_field(reinterpret_cast<const uint32_t*>(v.aux), (uint64_t)row * lw, lw) : 0ull) + v.pad;
        if ((v.naux >> 8) & 0xffu) l = ((l < 4u ? v.delta >> (16u * (uint32_t)l) : v.extra >> (16u * ((uint32_t)l - 4u))) & 0xffffull);
        const uint64_t r = (v.width ? load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width) : 0ull) + v.base;
        const uint64_t bits = ((l & 0xffffull) << shift) | r;
        return v.type == 10 ? (uint64_t)(uint32_t)bits : bits;
    }
    case CK_ALP: {   // Decoder.DecodeValue (internal/encode/alp/decoder.go:97-124): patch or T(val) * F10[f] * IF10[e]
        if (v.naux) {
            const uint32_t* pm = reinterpret_cast<const uint32_t*>(v.aux + alp_mask_off(v.naux));
            if ((__ldg(pm + (row >> 5)) >> (row & 31u)) & 1u) {
                const uint32_t* pos = reinterpret_cast<const uint32_t*>(v.aux);
                uint32_t lo = 0, hi = v.naux;
                while (lo < hi) { uint32_t m = (lo + hi) >> 1; if (__ldg(pos + m) < row) lo = m + 1; else hi = m; }
                return __ldg(reinterpret_cast<const unsigned long long*>(v.aux + alp_vals_off(v.naux)) + lo);
            }
        }
        uint64_t f = 0;
        if (v.width) f = staged ? load_field(staged, (uint64_t)row_in_tile * v.width, v.width)
                                : load_field(reinterpret_cast<const uint32_t*>(v.data), (uint64_t)row * v.width, v.width);
        double d = __dmul_rn(__dmul_rn(__ll2double_rn((long long)(f + v.base)), ALP_F10[v.delta & 0xffu]), ALP_IF10[(v.delta >> 8) & 0xffu]);
        return (uint64_t)__double_as_longlong(d);
    }
    }
    return 0;
}

}  // namespace kx
