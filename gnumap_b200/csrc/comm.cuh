// comm.cuh -- several GPUs behind one process: the accumulator reduce of the path, inside the C ABI.
//
// The reference's in-process model is `-c N` pthreads sharing one set of accumulators under a mutex (reference
// src/Driver.cpp:1527-1554); its multi-node model sums the per-node accumulators at the end with MPI_Allreduce
// (amount_genome, src/Driver.cpp:1660-1672) and MPI_Reduce to rank 0 (the five read planes, :1719-1767).  Here every
// GPU owns a context with its own accumulators (device atomics, no lock) and the final sum is one of
//
//   GMX_COMM_PEER  own kernels over peer memory: GPU g sums slice g of every context's accumulators with direct NVLink
//                  loads (fixed order: deterministic) and stores the result straight into the root's memory (and into
//                  every peer's for an all-reduce).  One launch per GPU, all concurrent; every GPU moves (n-1)/n of the
//                  array in and 1/n out, so the root's links are never the bottleneck.  Also serves contexts that share
//                  a device.
//   GMX_COMM_NCCL  ncclCommInitAll over the contexts' devices + ncclReduce / ncclAllReduce (libnccl.so.2 is opened at
//                  run time; the library does not link against it).
#pragma once

#include <dlfcn.h>
#include <nccl.h>          // types and enums only: every entry point is resolved with dlsym

#define GMX_COMM_MAX 16

struct NcclApi {
    void *lib = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool load()
    {
        if (lib) return true;
        lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        Reduce = (decltype(Reduce))dlsym(lib, "ncclReduce");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && Reduce && AllReduce && GroupStart && GroupEnd && GetErrorString;
    }
};

struct gmx_comm {
    int n = 0;
    int backend = GMX_COMM_PEER;
    gmx_ctx *ctx[GMX_COMM_MAX];
    NcclApi nccl;
    ncclComm_t comms[GMX_COMM_MAX];
    bool have_comms = false;
    cudaEvent_t ev[2] = {nullptr, nullptr};      // on the root's stream, around the reduce
    float last_ms = 0;
    uint64_t last_bytes = 0;
    std::string err;
};

struct PeerBufs {
    float *buf[GMX_COMM_MAX];     // the same array in every context
    int n;
};

// sum over the contexts of elements [lo, hi), in context order; the result goes to context 0, or to all of them
__global__ void __launch_bounds__(256) k_reduce_slice(PeerBufs B, uint64_t lo, uint64_t hi, int all)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    // head up to a 16-byte boundary (cudaMalloc'd bases are aligned alike, so one test serves all buffers), float4 body, tail
    uint64_t a = (lo + 3) & ~(uint64_t)3; if (a > hi) a = hi;
    uint64_t b = hi & ~(uint64_t)3; if (b < a) b = a;
    auto one = [&](uint64_t i) {
        float s = B.buf[0][i];
        for (int p = 1; p < B.n; ++p) s = __fadd_rn(s, B.buf[p][i]);
        B.buf[0][i] = s;
        if (all) for (int p = 1; p < B.n; ++p) B.buf[p][i] = s;
    };
    for (uint64_t i = lo + tid; i < a; i += nth) one(i);
    for (uint64_t i = b + tid; i < hi; i += nth) one(i);
    for (uint64_t v = (a >> 2) + tid; v < (b >> 2); v += nth) {
        float4 s = reinterpret_cast<const float4 *>(B.buf[0])[v];
#pragma unroll 4
        for (int p = 1; p < B.n; ++p) {
            const float4 t = reinterpret_cast<const float4 *>(B.buf[p])[v];
            s.x = __fadd_rn(s.x, t.x); s.y = __fadd_rn(s.y, t.y); s.z = __fadd_rn(s.z, t.z); s.w = __fadd_rn(s.w, t.w);
        }
        reinterpret_cast<float4 *>(B.buf[0])[v] = s;
        if (all) for (int p = 1; p < B.n; ++p) reinterpret_cast<float4 *>(B.buf[p])[v] = s;
    }
}
