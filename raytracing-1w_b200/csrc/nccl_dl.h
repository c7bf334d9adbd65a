// nccl_dl.h — NCCL bound at run time (dlopen), so that single-GPU users of librt1w.so need no NCCL at all and a host
// process that already carries one (PyTorch bundles its own libnccl.so.2) shares it instead of loading a second copy.
// Only the calls the radiance reduce needs (one ncclReduce(sum, fp32, 3 W H) per render).
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>
#include <string>

namespace rt1w {

struct NcclApi {
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclReduce) Reduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    void *handle = nullptr;
    std::string error;   // why loading failed
    std::string origin;  // which library was bound
};

// Search order: $RT1W_NCCL_LIB, a libnccl.so.2 the process has already loaded, then the dynamic linker's search path.
inline const NcclApi &nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = std::getenv("RT1W_NCCL_LIB");
        if (env && *env) {
            api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
            api.origin = env;
        }
        if (!api.handle) {
            api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
            api.origin = "libnccl.so.2 (already loaded by the host process)";
        }
        if (!api.handle) {
            api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
            api.origin = "libnccl.so.2";
        }
        if (!api.handle) {
            const char *why = dlerror();
            api.error = std::string("NCCL is not available (multi-GPU rendering needs libnccl.so.2): ") + (why ? why : "dlopen failed");
            return;
        }
        auto sym = [&](const char *name) -> void * {
            void *p = dlsym(api.handle, name);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return api;
}

} // namespace rt1w
