// kx_string.cu — byte-string leaves: FixedString / CompactString / DictString containers
// (internal/encode/string_{fixed,compact,dict}.go) matched row by row like matchStringEqual … matchStringBetween
// (internal/encode/string_match.go:13-188: bytes.Equal / bytes.Compare of every row against the operand), and set
// membership like bytesInSetMatcher / bytesNotInSetMatcher (internal/operator/filter/match_bytes.go:392-520).
//
// Runs as a pre-pass on the scan stream (like runfill_kernel): one thread per row resolves its row's (offset, length)
// from the block's flat index array, compares the bytes with the operand(s) from the program's byte pool and the warp's
// ballot becomes one word of the leaf's bitset, which the scan kernel then streams as a 1-bit column (LM_BITS) and
// combines with the other leaves.  Dictionary blocks evaluate the predicate once per dictionary ENTRY into a shared-
// memory bitmap when the dictionary is small, rows then test one bit of it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kx_types.h"
#include "kx_kernels.h"

namespace kx {

// bytes.Compare(v, a): -1 / 0 / +1, lexicographic on unsigned bytes, a proper prefix is smaller
__device__ __forceinline__ int bytes_cmp(const uint8_t* __restrict__ v, uint32_t vl, const uint8_t* __restrict__ a, uint32_t al) {
    const uint32_t m = min(vl, al);
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t x = v[i], y = __ldg(a + i);
        if (x != y) return x < y ? -1 : 1;
    }
    return vl < al ? -1 : (vl > al ? 1 : 0);
}

// IN / NOT IN (bytesInSetMatcher, internal/operator/filter/match_bytes.go:392-520): binary search in the sorted set;
// a = [al x (offset, length)] + bytes
static __device__ __noinline__ bool str_in_set(uint32_t mode, const uint8_t* __restrict__ v, uint32_t vl, const uint8_t* __restrict__ a, uint32_t al) {
    const uint32_t* tab = reinterpret_cast<const uint32_t*>(a);
    uint32_t lo = 0, hi = al;
    bool found = false;
    while (lo < hi) {
        const uint32_t m = (lo + hi) >> 1;
        const int c = bytes_cmp(v, vl, a + __ldg(tab + 2 * m), __ldg(tab + 2 * m + 1));
        if (c == 0) { found = true; break; }
        if (c < 0) hi = m; else lo = m + 1;
    }
    return mode == 7u ? found : !found;
}

__device__ __forceinline__ bool str_pred(uint32_t mode, const uint8_t* __restrict__ v, uint32_t vl, const uint8_t* __restrict__ a, uint32_t al,
                                         const uint8_t* __restrict__ b, uint32_t bl) {
    if (mode == 7u || mode == 8u) return str_in_set(mode, v, vl, a, al);   // IN / NOT IN: out of line, the scalar modes keep their code
    switch (mode) {
    case 1: return vl == al && bytes_cmp(v, vl, a, al) == 0;          // MatchEqual (length first: most rows differ there or in byte 0)
    case 2: return !(vl == al && bytes_cmp(v, vl, a, al) == 0);       // MatchNotEqual
    case 3: return bytes_cmp(v, vl, a, al) > 0;
    case 4: return bytes_cmp(v, vl, a, al) >= 0;
    case 5: return bytes_cmp(v, vl, a, al) < 0;
    case 6: return bytes_cmp(v, vl, a, al) <= 0;
    default: return bytes_cmp(v, vl, a, al) >= 0 && bytes_cmp(v, vl, b, bl) <= 0;   // MatchBetween
    }
}

constexpr uint32_t STR_DICT_SMEM_ENTRIES = 32768;   // dictionary predicate bitmap in shared memory: 4 KB

__global__ void __launch_bounds__(256, 8) strmatch_kernel(const StrJob* __restrict__ jobs, uint32_t njobs, const uint8_t* __restrict__ pool, uint8_t* __restrict__ out_base) {
    __shared__ uint32_t dict_bits[STR_DICT_SMEM_ENTRIES / 32];
  for (uint32_t jb = blockIdx.y; jb < njobs; jb += gridDim.y) {   // (gridDim.y is capped at 65535 jobs per launch)
    __syncthreads();   // the previous job's dictionary bitmap is no longer read
    const StrJob& J = jobs[jb];
    const ColView& v = J.view;
    const uint32_t n = v.n, layout = v.is_raw;
    const uint8_t* a = pool + J.a_off;
    const uint8_t* b = pool + J.b_off;
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(v.aux);
    uint32_t* out = reinterpret_cast<uint32_t*>(out_base + J.out_off);
    const uint32_t lane = threadIdx.x & 31u;

    const bool dict_map = layout == STR_DICT && v.naux <= STR_DICT_SMEM_ENTRIES;
    if (dict_map) {   // predicate per dictionary entry (every block of the job builds its own copy: the dictionary is small)
        for (uint32_t base = (threadIdx.x >> 5) * 32u; base < v.naux; base += (blockDim.x >> 5) * 32u) {
            const uint32_t c = base + lane;
            bool p = false;
            if (c < v.naux) p = str_pred(J.mode, v.data + __ldg(idx + n + c), __ldg(idx + n + v.naux + c), a, J.a_len, b, J.b_len);
            const uint32_t w = __ballot_sync(0xffffffffu, p);
            if (lane == 0) dict_bits[base >> 5] = w;
        }
        __syncthreads();
    }
    // rows: whole warps walk 32-row groups, grid-stride over the block's rows
    const uint32_t ngroups = (n + 31u) >> 5, warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < ngroups; g += warps) {
        const uint32_t row = g * 32u + lane;
        bool p = false;
        if (row < n) {
            if (dict_map) {
                const uint32_t c = __ldg(idx + row);
                p = (dict_bits[c >> 5] >> (c & 31u)) & 1u;
            } else {
                uint32_t ofs, len;
                switch (layout) {
                case STR_FIXED: len = (uint32_t)v.delta; ofs = row * len; break;
                case STR_COMPACT: ofs = __ldg(idx + row); len = __ldg(idx + n + row); break;
                case STR_DICT: { const uint32_t c = __ldg(idx + row); ofs = __ldg(idx + n + c); len = __ldg(idx + n + v.naux + c); break; }
                default: ofs = 0; len = (uint32_t)v.delta; break;   // STR_CONST (normally decided on the host)
                }
                p = str_pred(J.mode, v.data + ofs, len, a, J.a_len, b, J.b_len);
            }
        }
        const uint32_t w = __ballot_sync(0xffffffffu, p);
        if (lane == 0) out[g] = w;   // rows past n are zero: the tail bits of the bitset stay clear
    }
  }
}

// (offset, length) of row `row` of a string block inside its byte buffer
__device__ __forceinline__ void str_row(const ColView& v, uint32_t row, uint32_t& ofs, uint32_t& len) {
    const uint32_t* idx = reinterpret_cast<const uint32_t*>(v.aux);
    switch (v.is_raw) {
    case STR_FIXED: len = (uint32_t)v.delta; ofs = row * len; break;
    case STR_COMPACT: ofs = __ldg(idx + row); len = __ldg(idx + v.n + row); break;
    case STR_DICT: { const uint32_t c = __ldg(idx + row); ofs = __ldg(idx + v.n + c); len = __ldg(idx + v.n + v.naux + c); break; }
    default: ofs = 0; len = (uint32_t)v.delta; break;   // STR_CONST
    }
}

// StringContainer.AppendTo(dst, sel) for a batch of packs (internal/encode/string_{const,fixed,compact,dict}.go AppendTo with
// a selection; query/result.go:196-264 copies the selected rows of bytes columns like any other result column).
// Pass 1: one thread per selected row writes the row's length; an exclusive scan turns the lengths into offsets;
// pass 2: one WARP per selected row copies its bytes (lane-strided: coalesced for long rows, one transaction for short ones).
__global__ void strgather_len_kernel(const ColView* __restrict__ views, const unsigned long long* __restrict__ sel_off, uint32_t npacks,
                                     const uint32_t* __restrict__ sel, uint64_t total, uint32_t* __restrict__ lens,
                                     unsigned long long* __restrict__ nbytes) {
    unsigned long long mine = 0;   // 64-bit grand total beside the 32-bit offsets (a selection of 4 GiB or more is refused, not wrapped)
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = npacks;   // last pack with sel_off <= i
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sel_off[m] <= i) lo = m; else hi = m; }
        uint32_t ofs, len;
        str_row(views[lo], sel[i], ofs, len);
        lens[i] = len;
        mine += len;
    }
    for (int off = 16; off > 0; off >>= 1) mine += __shfl_down_sync(0xffffffffu, mine, off);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(nbytes, mine);
}

__global__ void strgather_copy_kernel(const ColView* __restrict__ views, const unsigned long long* __restrict__ sel_off, uint32_t npacks,
                                      const uint32_t* __restrict__ sel, uint64_t total, const uint32_t* __restrict__ offs, uint8_t* __restrict__ out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t i = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5; i < total; i += warps) {
        uint32_t lo = 0, hi = npacks;
        while (hi - lo > 1) { uint32_t m = (lo + hi) >> 1; if (sel_off[m] <= i) lo = m; else hi = m; }
        const ColView& v = views[lo];
        uint32_t ofs, len;
        str_row(v, sel[i], ofs, len);
        const uint8_t* src = v.data + ofs;
        uint8_t* dst = out + offs[i];
        for (uint32_t b = lane; b < len; b += 32u) dst[b] = src[b];
    }
}

cudaError_t launch_strgather_len(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                                 uint32_t* lens, unsigned long long* nbytes, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(nbytes, 0, 8, stream);
    if (e != cudaSuccess || total == 0) return e;
    uint64_t g = (total + 255u) / 256u;
    if (g > 148u * 16u) g = 148u * 16u;
    strgather_len_kernel<<<(unsigned)g, 256, 0, stream>>>(views, sel_off, npacks, sel, total, lens, nbytes);
    return cudaGetLastError();
}

cudaError_t launch_strgather_copy(const ColView* views, const unsigned long long* sel_off, uint32_t npacks, const uint32_t* sel, uint64_t total,
                                  const uint32_t* offs, uint8_t* out, cudaStream_t stream) {
    if (total == 0) return cudaSuccess;
    uint64_t g = (total + 7u) / 8u;   // 8 warps per block, one row per warp and step
    if (g > 148u * 16u) g = 148u * 16u;
    strgather_copy_kernel<<<(unsigned)g, 256, 0, stream>>>(views, sel_off, npacks, sel, total, offs, out);
    return cudaGetLastError();
}

cudaError_t launch_strmatch(const StrJob* jobs, uint32_t njobs, uint32_t max_rows, const uint8_t* pool, uint8_t* out_base, cudaStream_t stream) {
    if (njobs == 0 || max_rows == 0) return cudaSuccess;
    uint32_t gx = (max_rows + 256u * 8u - 1u) / (256u * 8u);   // ~8 groups per warp
    if (gx > 148u * 4u) gx = 148u * 4u;
    if (gx < 1u) gx = 1u;
    strmatch_kernel<<<dim3(gx, njobs < 65535u ? njobs : 65535u), 256, 0, stream>>>(jobs, njobs, pool, out_base);
    return cudaGetLastError();
}

}  // namespace kx
