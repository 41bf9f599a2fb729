// nccl_dyn.h — NCCL bound at run time with dlopen("libnccl.so.2").
//
// The library is only needed by sharded lattices (chemsim_lbm_create_slab with
// nranks > 1).  Binding it lazily keeps libchemsim_lbm.so loadable on a box
// without NCCL, and inside a PyTorch process it resolves to the libnccl that
// torch already loaded, so both share one NCCL.
#pragma once

#include <dlfcn.h>
#include <nccl.h>   // types and enums only; no symbol is linked

#include <string>

namespace chemsim {

struct NcclDyn {
    bool ok = false;
    std::string error;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;

    NcclDyn()
    {
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { error = dlerror(); return; }
#define CHEMSIM_NCCL_SYM(name)                                                   \
        name = reinterpret_cast<decltype(name)>(dlsym(lib, "nccl" #name));       \
        if (!name) { error = "missing symbol nccl" #name; return; }
        CHEMSIM_NCCL_SYM(GetUniqueId)
        CHEMSIM_NCCL_SYM(CommInitRank)
        CHEMSIM_NCCL_SYM(CommDestroy)
        CHEMSIM_NCCL_SYM(GroupStart)
        CHEMSIM_NCCL_SYM(GroupEnd)
        CHEMSIM_NCCL_SYM(Send)
        CHEMSIM_NCCL_SYM(Recv)
        CHEMSIM_NCCL_SYM(AllReduce)
        CHEMSIM_NCCL_SYM(AllGather)
        CHEMSIM_NCCL_SYM(GetErrorString)
#undef CHEMSIM_NCCL_SYM
        ok = true;
    }
};

inline const NcclDyn &nccl_dyn()
{
    static const NcclDyn instance;
    return instance;
}

}  // namespace chemsim
