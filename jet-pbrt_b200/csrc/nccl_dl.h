// nccl_dl.h -- NCCL bound at RUN time (dlopen), so that libjetpbrt_b200.so carries no link-time dependency on it:
// a single-GPU caller needs no NCCL at all, and inside a process that already loaded NCCL (torch.distributed) the very
// same library instance is reused (dlopen by SONAME returns the loaded object) -- two NCCL copies in one process
// would each bring their own proxy threads and their own view of the NVLink / NVSwitch topology.
//
// Only the handful of entry points the film reduce needs (ONE ncclReduce of the float32 film).
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>

#include <string>

namespace jpbrt {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;

    bool Load() {
        if (handle) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        // JPBRT_NCCL_LIB=<path>: load exactly this NCCL.  (Whichever libnccl.so.2 is loaded FIRST is the one the whole
        // process gets: a program that also uses another NCCL client -- torch.distributed brings its own, newer copy --
        // must let that client load first, or point this at the same file.  The Python glue imports torch before the
        // first communicator call for that reason.)
        if (const char* forced = getenv("JPBRT_NCCL_LIB")) handle = dlopen(forced, RTLD_NOW | RTLD_GLOBAL);
        for (const char* n : names) {
            if (handle) break;
            handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD);  // already in the process (e.g. torch's)?
            if (handle) break;
        }
        for (int i = 0; !handle && i < 2; ++i) handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!handle) {
            const char* e = dlerror();
            error = std::string("NCCL is not available (dlopen libnccl.so.2: ") + (e ? e : "?") + ")";
            return false;
        }
        bool ok = true;
        auto sym = [&](const char* name) { void* p = dlsym(handle, name); if (!p) { ok = false; error = std::string("NCCL symbol missing: ") + name; } return p; };
        GetVersion = (decltype(GetVersion))sym("ncclGetVersion");
        GetUniqueId = (decltype(GetUniqueId))sym("ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))sym("ncclCommInitRank");
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Reduce = (decltype(Reduce))sym("ncclReduce");
        AllReduce = (decltype(AllReduce))sym("ncclAllReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        if (!ok) { handle = nullptr; return false; }
        return true;
    }
};

inline NcclApi& nccl() {
    static NcclApi api;
    return api;
}

}  // namespace jpbrt
