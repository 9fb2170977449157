// gmg_nccl.h -- the handful of NCCL entry points the z-slab sharding uses, bound at run time with dlopen so that the
// single-GPU library has no link-time dependency on NCCL and a process that already loaded torch's bundled
// libnccl.so.2 shares that copy (same SONAME).  Types restated from NCCL's public ABI (stable since 2.x).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace gmg
{
struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm *NcclComm;
enum { NCCL_SUCCESS = 0, NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
};

// nullptr (with the reason in *why) when no libnccl.so.2 can be loaded
const NcclApi *ncclApi(const char **why);
} // namespace gmg
