// Multi-GPU plumbing of the core: NCCL resolved at run time (dlopen of libnccl.so.2 -- the copy the
// host process already carries, e.g. torch's, or the system one), and the partition rule shared by
// the list build, the bonded kernel and the host-side tests.
//
// Reference behaviour being replaced: one kernel object per device, tiles partitioned by context
// index, exceptions split in contiguous ranges, reciprocal space / self energy / dispersion correction
// on context 0 only (platforms/cuda/src/CudaParallelOpenMMLabKernels.cpp:19-53;
// platforms/cuda/src/CudaOpenMMLabKernels.cpp:379,405,466,680-683,762-765).  The reference has no
// collective of its own (OpenMM sums the per-device force buffers); here it is one ncclBroadcast of
// the positions in and one ncclReduce(sum) of 2^32 fixed-point forces + FP64 slice energies out.
#ifndef SNB_COMM_CUH_
#define SNB_COMM_CUH_

#include <cuda_runtime.h>
#include <nccl.h>
#include <stdint.h>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;     // optional
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

// Loads NCCL on first use; returns null (and fills `why`) when no libnccl.so.2 can be found.
const NcclApi* ncclApi(const char** why);

// Contiguous share [lo, hi) of `n` work units for `rank` of `world`, rank 0 weighted by
// `rank0_share` relative to 1.0 for every other rank (it also runs reciprocal space).
void partitionRange(int64_t n, int world, int rank, double rank0_share, int64_t* lo, int64_t* hi);

#endif
