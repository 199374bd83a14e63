// kx_comm.cu — the one small collective of a pack-sharded scan, inside the library.
//
// Packs are independent units (internal/pack/table/reader.go:299-449 carries no cross-pack state), so a table is
// sharded by pack key over the GPUs of a box, one kx_ctx per GPU, with NO data-path collective.  What the ranks
// exchange per query is one fixed-size record each — the match count and the per-rank partial aggregates — with ONE
// ncclAllGather enqueued on the scan stream right behind the scan kernel; a one-warp kernel then combines the records
// in RANK order (fixed topology: every rank computes the bit-identical result) and a single D2H copy returns it.
// No host round trip between scan, collective and combine.
//
// NCCL is bound at run time (dlopen "libnccl.so.2", KX_NCCL_LIB overrides): a process that already carries an NCCL
// (e.g. PyTorch's) shares it, a single-GPU host needs none.  Only the C types of nccl.h are used at build time.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "kx_comm.h"

namespace kx {

namespace {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("KX_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) { api.error = std::string("NCCL not found: ") + (dlerror() ? dlerror() : "dlopen failed"); return; }
        auto sym = [&](const char* s) -> void* {
            void* p = dlsym(api.handle, s);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + s;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return &api;
}

std::string nccl_err(NcclApi* a, ncclResult_t r, const char* what) {
    return std::string(what) + ": " + (a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
}

}  // namespace

struct CommState {
    ncclComm_t comm = nullptr;
    int nranks = 1, rank = 0;
};

static_assert(KX_COMM_ID_BYTES_INTERNAL == NCCL_UNIQUE_ID_BYTES, "unique id size");

int comm_unique_id(void* out, std::string& err) {
    NcclApi* a = nccl_api();
    if (!a->error.empty()) { err = a->error; return -1; }
    ncclUniqueId id;
    ncclResult_t r = a->GetUniqueId(&id);
    if (r != ncclSuccess) { err = nccl_err(a, r, "ncclGetUniqueId"); return -1; }
    std::memcpy(out, id.internal, NCCL_UNIQUE_ID_BYTES);
    return 0;
}

int comm_create(int nranks, int rank, const void* id_bytes, CommState** out, std::string& err) {
    *out = nullptr;
    auto* c = new CommState();
    c->nranks = nranks; c->rank = rank;
    if (nranks > 1) {
        NcclApi* a = nccl_api();
        if (!a->error.empty()) { err = a->error; delete c; return -1; }
        ncclUniqueId id;
        std::memcpy(id.internal, id_bytes, NCCL_UNIQUE_ID_BYTES);
        ncclResult_t r = a->CommInitRank(&c->comm, nranks, id, rank);
        if (r != ncclSuccess) { err = nccl_err(a, r, "ncclCommInitRank"); delete c; return -1; }
    }
    *out = c;
    return 0;
}

void comm_destroy(CommState* c) {
    if (!c) return;
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    delete c;
}

int comm_nranks(const CommState* c) { return c ? c->nranks : 1; }
int comm_rank(const CommState* c) { return c ? c->rank : 0; }
int comm_nccl_version() {
    NcclApi* a = nccl_api();
    int v = 0;
    if (a->error.empty() && a->GetVersion) a->GetVersion(&v);
    return v;
}

int comm_allgather(CommState* c, const void* send, void* recv, size_t bytes, cudaStream_t stream, std::string& err) {
    if (!c || c->nranks <= 1) {
        if (send != recv) {
            cudaError_t e = cudaMemcpyAsync(recv, send, bytes, cudaMemcpyDeviceToDevice, stream);
            if (e != cudaSuccess) { err = cudaGetErrorString(e); return -1; }
        }
        return 0;
    }
    NcclApi* a = nccl_api();
    ncclResult_t r = a->AllGather(send, recv, bytes, ncclUint8, c->comm, stream);
    if (r != ncclSuccess) { err = nccl_err(a, r, "ncclAllGather"); return -1; }
    return 0;
}

// ------------------------------------------------------------------------------ the two one-block kernels around it
namespace {

__device__ __forceinline__ double as_f64(uint64_t b) { return __longlong_as_double((long long)b); }
__device__ __forceinline__ uint64_t as_u64(double d) { return (uint64_t)__double_as_longlong(d); }

// this rank's record: total match count (sum of the per-pack counts) + the combined aggregates of its packs
__global__ void xchg_pack_kernel(const unsigned long long* __restrict__ counts, uint32_t npacks, const AggPartial* __restrict__ agg, uint32_t naggs,
                                 RankPartial* __restrict__ out) {
    __shared__ unsigned long long ws[32];
    unsigned long long c = 0;
    for (uint32_t i = threadIdx.x; i < npacks; i += blockDim.x) c += counts[i];
    for (int off = 16; off > 0; off >>= 1) c += __shfl_down_sync(0xffffffffu, c, off);
    if ((threadIdx.x & 31u) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t += ws[w];
        out->total_count = t; out->pad = 0;
    }
    if (threadIdx.x < MAX_AGGS) {
        AggPartial z{};
        out->agg[threadIdx.x] = threadIdx.x < naggs ? agg[threadIdx.x] : z;
    }
}

// all ranks' records, combined in rank order (the same arithmetic as the in-kernel combine of the per-CTA partials)
__global__ void xchg_combine_kernel(const RankPartial* __restrict__ recs, uint32_t nranks, uint32_t naggs, uint32_t agg_types /* 4 x u8 */,
                                    RankPartial* __restrict__ out) {
    const uint32_t j = threadIdx.x;
    if (j == 31u) {
        unsigned long long t = 0;
        for (uint32_t r = 0; r < nranks; ++r) t += recs[r].total_count;
        out->total_count = t; out->pad = 0;
    }
    if (j >= MAX_AGGS) return;
    AggPartial acc{};
    if (j < naggs) {
        const int type = (agg_types >> (8 * j)) & 0xffu;
        for (uint32_t r = 0; r < nranks; ++r) {
            const AggPartial p = recs[r].agg[j];
            if (!p.valid) continue;
            if (!acc.valid) { acc = p; continue; }
            acc.count += p.count;
            if (type == 9 || type == 10) {
                double s = as_f64(acc.sum), e = acc.err, s2 = as_f64(p.sum);
                double t = s + s2;
                double c = (fabs(s) >= fabs(s2)) ? ((s - t) + s2) : ((s2 - t) + s);
                acc.sum = as_u64(t); acc.err = e + p.err + c;
                if (as_f64(p.mn) < as_f64(acc.mn)) acc.mn = p.mn;
                if (as_f64(p.mx) > as_f64(acc.mx)) acc.mx = p.mx;
            } else {
                acc.sum += p.sum;
                if (p.mn < acc.mn) acc.mn = p.mn;
                if (p.mx > acc.mx) acc.mx = p.mx;
            }
        }
    }
    out->agg[j] = acc;
}

}  // namespace

cudaError_t launch_xchg_pack(const unsigned long long* counts, uint32_t npacks, const AggPartial* agg, uint32_t naggs, RankPartial* out, cudaStream_t stream) {
    xchg_pack_kernel<<<1, 256, 0, stream>>>(counts, npacks, agg, naggs, out);
    return cudaGetLastError();
}
cudaError_t launch_xchg_combine(const RankPartial* recs, uint32_t nranks, uint32_t naggs, const uint8_t* agg_type, RankPartial* out, cudaStream_t stream) {
    uint32_t types = 0;
    for (uint32_t j = 0; j < naggs && j < MAX_AGGS; ++j) types |= uint32_t(agg_type[j]) << (8 * j);
    xchg_combine_kernel<<<1, 32, 0, stream>>>(recs, nranks, naggs, types, out);
    return cudaGetLastError();
}

}  // namespace kx
