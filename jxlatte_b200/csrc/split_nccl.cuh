// split_nccl.cuh -- one frame split by group rows across the GPUs of a box, entirely behind the C ABI:
// a Java host (one thread and one context per GPU) calls jxlb200_comm_init once and jxlb200_vardct_reconstruct_split_dev per
// frame; the halo rows travel with ncclSend / ncclRecv over NVLink on a side stream and overlap the slab's own work.
//
// Order of work in one call (slab of R rows, neighbours above and / or below):
//   1. stage 1 of the slab                                                                               main stream
//   2. send my 8 boundary rows up / down, receive the neighbours' (3 planes each, one NCCL group)        comm stream, after 1
//   3. stage 2 of the rows that need no halo [8, R - 8)                                                  main stream, beside 2
//   4. stage 2 of the top 8 and bottom 8 rows once the halos are in                                     main stream, after 2
// Stage 2 of a sub-range with has_top / has_bottom set reads its neighbour rows from the slab itself.  Bit-identical to the whole frame
// (tests/test_baseline_sizes_gpu.py::test_nccl_group_row_split_is_bit_identical).
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy already in the process -- PyTorch's, a JVM shim's -- or the
// system's), so the library loads and every other entry point works on a box without NCCL.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi &nccl_api() {
    static NcclApi a;
    static bool tried = false;
    if (tried) return a;
    tried = true;
    a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);        // the copy this process already uses, if any
    if (!a.h) a.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) a.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!a.h) return a;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(a.h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(a.h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.h, "ncclCommDestroy");
    a.Send = (decltype(a.Send))dlsym(a.h, "ncclSend");
    a.Recv = (decltype(a.Recv))dlsym(a.h, "ncclRecv");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.h, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.h, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.h, "ncclGetErrorString");
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.Send && a.Recv && a.GroupStart && a.GroupEnd && a.GetErrorString;
    return a;
}
