// kx_comm.h — the communicator of a pack-sharded scan (kx_comm.cu): run-time bound NCCL, the per-rank record that is
// all-gathered once per query, and the two one-block kernels that pack / combine it on the scan stream.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "kx_types.h"

namespace kx {

constexpr int KX_COMM_ID_BYTES_INTERNAL = 128;   // NCCL_UNIQUE_ID_BYTES

// one rank's contribution to a query: 16 + MAX_AGGS * 48 = 208 bytes
struct RankPartial {
    unsigned long long total_count;   // matches over all packs of the rank
    uint64_t pad;
    AggPartial agg[MAX_AGGS];
};

struct CommState;
int  comm_unique_id(void* out128, std::string& err);
int  comm_create(int nranks, int rank, const void* id128, CommState** out, std::string& err);   // the context's device must be current
void comm_destroy(CommState* c);
int  comm_nranks(const CommState* c);
int  comm_rank(const CommState* c);
int  comm_nccl_version();
// enqueue on `stream`: every rank contributes `bytes` bytes at `send`; recv receives nranks * bytes in rank order
int  comm_allgather(CommState* c, const void* send, void* recv, size_t bytes, cudaStream_t stream, std::string& err);

cudaError_t launch_xchg_pack(const unsigned long long* counts, uint32_t npacks, const AggPartial* agg, uint32_t naggs, RankPartial* out, cudaStream_t stream);
cudaError_t launch_xchg_combine(const RankPartial* recs, uint32_t nranks, uint32_t naggs, const uint8_t* agg_type, RankPartial* out, cudaStream_t stream);

}  // namespace kx
