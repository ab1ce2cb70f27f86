// See comm.cuh.
#include "comm.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <mutex>

namespace {
NcclApi g_api;
const char* g_why = nullptr;
std::once_flag g_once;

void loadNccl() {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_api.lib = dlopen(n, RTLD_NOW|RTLD_GLOBAL);
        if (g_api.lib) break;
    }
    if (!g_api.lib) {
        g_why = "libnccl.so.2 not found (dlopen)";
        return;
    }
#define SNB_SYM(field, name) g_api.field = (decltype(g_api.field)) dlsym(g_api.lib, name); \
    if (!g_api.field) { g_why = "symbol " name " missing from libnccl"; return; }
    SNB_SYM(GetUniqueId, "ncclGetUniqueId")
    SNB_SYM(CommInitRank, "ncclCommInitRank")
    SNB_SYM(CommDestroy, "ncclCommDestroy")
    SNB_SYM(Reduce, "ncclReduce")
    SNB_SYM(AllReduce, "ncclAllReduce")
    SNB_SYM(Broadcast, "ncclBroadcast")
    SNB_SYM(GroupStart, "ncclGroupStart")
    SNB_SYM(GroupEnd, "ncclGroupEnd")
    SNB_SYM(GetErrorString, "ncclGetErrorString")
#undef SNB_SYM
    g_api.CommGetAsyncError = (decltype(g_api.CommGetAsyncError)) dlsym(g_api.lib, "ncclCommGetAsyncError");
}
}  // namespace

const NcclApi* ncclApi(const char** why) {
    std::call_once(g_once, loadNccl);
    if (g_why) {
        if (why) *why = g_why;
        return nullptr;
    }
    return &g_api;
}

void partitionRange(int64_t n, int world, int rank, double rank0_share, int64_t* lo, int64_t* hi) {
    if (world <= 1) {
        *lo = 0; *hi = n;
        return;
    }
    rank0_share = std::max(0.0, rank0_share);
    const double total = rank0_share+(world-1);
    auto edge = [&](int r) -> int64_t {     // first unit of rank r
        if (r <= 0) return 0;
        if (r >= world) return n;
        const double w = rank0_share+(r-1);
        return std::min<int64_t>(n, (int64_t) std::llround((double) n*w/total));
    };
    *lo = edge(rank);
    *hi = edge(rank+1);
}
