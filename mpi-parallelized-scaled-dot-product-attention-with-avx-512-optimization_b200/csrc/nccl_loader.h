// nccl_loader.h -- NCCL bound at run time with dlopen, so that the single-GPU path has
// no NCCL dependency and the library loads on a box without it.  Only the entry points
// the merge needs (allreduce MAX / SUM, reduce SUM: attention-mpi.c:342,354,380) plus
// broadcast / send / recv for the rank-0 scatter of the drop-in (mpi.c:196,232-264,305).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace sdpa {

// Mirrors of the public NCCL ABI (nccl.h); stable since NCCL 2.x.
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclInt32 = 2, ncclFloat32 = 7, ncclFloat64 = 8, ncclBfloat16 = 9, ncclUint8 = 1 } ncclDataType_t;
typedef enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 } ncclRedOp_t;

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*GetVersion)(int*);
};

// Returns nullptr (and sets the error string) when libnccl cannot be loaded.
const NcclApi* nccl_api();

}  // namespace sdpa
