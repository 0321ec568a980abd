// comm.cuh -- multi-GPU whole-genome mode: one job spread over the GPUs of a node (SURVEY 8e rows 2-3,
// BASELINE configs[4]: whole-genome / -t semantics, where regions on different chromosomes are linked).
//
// Decomposition. Pass 1 (K1: classify, statistics, ordered compaction) is the only stage that touches the
// 25 B/record columns; it is purely per record plus prefix counts, so every rank runs it on its own
// CONTIGUOUS SLICE of the globally (tid, pos)-sorted record stream (rank order = stream order; the cut may
// fall anywhere, also inside a chromosome). Everything after works on the anomalous 1-3 %:
//   exchange 1  all-gather of the per-rank totals (anomalous reads, kept proper pairs per key, records);
//               all-reduce of the pass-1 accumulators (counts: sum, first / last record per (bam, chromosome):
//               min / max of (global record index, pos)); every rank rebases its anomalous-read stream
//               (global record index, global inclusive proper-pair counts) straight into its slot of the
//               global stream, and the slots are all-gathered over NVLink (one grouped NCCL call);
//   replicated  K2 (regions) and K3 (mate join, edges) run on the global stream on every rank:
//               region indices and flush windows are therefore global and identical;
//   replicated  K4 as well (the table of deletion windows, the calls, the scores: a few hundred microseconds), so every
//               rank returns the complete SV table and there is no second exchange.
// NCCL is loaded with dlopen at the first use (a process that never attaches a communicator needs no NCCL).
#pragma once
#include <dlfcn.h>
#include <mutex>
#include <string>
#include <nccl.h>   // types and prototypes only; nothing is linked

#include "common.cuh"
#include "k234_regions_links_sv.cuh"

namespace bdk {

struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string error;
};

// Resolves the NCCL entry points once per process. A library that is already loaded under the soname
// libnccl.so.2 (e.g. the one PyTorch ships) is the one dlopen returns, so the process never holds two NCCLs.
inline NcclApi& nccl_api_storage() { static NcclApi api; return api; }
inline NcclApi* nccl_api() {
    static std::once_flag once;
    NcclApi& api = nccl_api_storage();
    std::call_once(once, [&api] {
        api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the process already has one (e.g. PyTorch's): use it
        const char* names[] = {getenv("BDK_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            if (api.handle) break;
            if (!n || !*n) continue;
            api.handle = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        }
        if (!api.handle) { const char* e = dlerror(); api.error = std::string("cannot load NCCL (libnccl.so.2): ") + (e ? e : "not found"); return; }
        bool ok = true;
        auto sym = [&](const char* s) -> void* { void* p = dlsym(api.handle, s); if (!p) { ok = false; api.error = std::string("NCCL symbol missing: ") + s; } return p; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return api.handle ? &api : nullptr;
}
inline const char* nccl_api_error() { return nccl_api_storage().error.c_str(); }

// ---- exchange 1 -------------------------------------------------------------------------------------------
// first / last record keys are (record index << 32 | pos): make the index global before the min / max all-reduce
__global__ void __launch_bounds__(256) comm_rebase_span_kernel(unsigned long long* __restrict__ first, unsigned long long* __restrict__ last,
                                                               uint32_t n, uint32_t rec_off) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (first[i] == ~0ull) continue;                 // this rank saw no record of (bam, chromosome) i: identity of min / max
        first[i] += (unsigned long long)rec_off << 32;
        last[i] += (unsigned long long)rec_off << 32;
    }
}

// local anomalous-read stream -> this rank's slot of the global stream: global record index, global inclusive
// proper-pair counts
__global__ void __launch_bounds__(GS_THREADS) comm_rebase_stream_kernel(const bdk_aread* __restrict__ ar, const uint32_t* __restrict__ P, uint32_t A_local,
        int nkey, uint32_t rec_off, const uint32_t* __restrict__ koff, bdk_aread* __restrict__ ar_g, uint32_t* __restrict__ P_g) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < A_local; i += gridDim.x * blockDim.x) {
        const int4* src = reinterpret_cast<const int4*>(ar + i);
        int4 lo = src[0], hi = src[1];                   // hi = (meta, record, qid)
        hi.y = (int)((uint32_t)hi.y + rec_off);
        int4* dst = reinterpret_cast<int4*>(ar_g + i);
        dst[0] = lo; dst[1] = hi;
        for (int k = 0; k < nkey; ++k) P_g[(size_t)i * nkey + k] = P[(size_t)i * nkey + k] + koff[k];
    }
}

}  // namespace bdk
