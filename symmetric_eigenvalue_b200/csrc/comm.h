// Collective layer: the communicator that ships the CUDA IPC handles of the peer-memory back end (p2p.h, the default
// on several GPUs), the all-gathers of the rarely used collective calls (eigenvector file, orthogonality check,
// selected vectors), and the FALLBACK exchange of the O(n) vectors of a cooperative merge (child eigenvalues,
// boundary rows -> z, secular root slices, residual partial sums) when peer memory is not available
// (CUPPEN_P2P=0).  Replaces the blocking MPI_Send/Recv/Bcast traffic of the reference
// (/root/reference/src/main.c:397-417,504-542, /root/reference/src/filehandling.c:347-348,415-437).
//
// Two back ends: NCCL over NVLink (resolved with dlopen so that single-GPU users need no NCCL),
// or caller-supplied callbacks (the CPU tests drive those with torch.distributed/gloo).
#ifndef CUPPEN_COMM_H
#define CUPPEN_COMM_H

#include "../../include/cuppen_b200.h"
#include "platform.h"

#if CUPPEN_CUDA
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace cuppen {

#if CUPPEN_CUDA
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    void load() {
        if (lib) return;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* nm : names) { lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) CUPPEN_THROW(CUPPEN_ERR_COMM, "cannot load libnccl.so.2: %s", dlerror());
#define CUPPEN_NCCL_SYM(field, name)                                                   \
        *(void**)(&field) = dlsym(lib, name);                                          \
        if (!field) CUPPEN_THROW(CUPPEN_ERR_COMM, "libnccl: missing symbol %s", name);
        CUPPEN_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        CUPPEN_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        CUPPEN_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        CUPPEN_NCCL_SYM(GroupStart, "ncclGroupStart")
        CUPPEN_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        CUPPEN_NCCL_SYM(Send, "ncclSend")
        CUPPEN_NCCL_SYM(Recv, "ncclRecv")
        CUPPEN_NCCL_SYM(AllReduce, "ncclAllReduce")
        CUPPEN_NCCL_SYM(AllGather, "ncclAllGather")
        CUPPEN_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef CUPPEN_NCCL_SYM
    }
};
inline NcclApi& nccl_api() { static NcclApi api; api.load(); return api; }
#define NCCL_CHECK(expr)                                                                         \
    do {                                                                                         \
        ncclResult_t r_ = (expr);                                                                \
        if (r_ != ncclSuccess)                                                                   \
            CUPPEN_THROW(CUPPEN_ERR_COMM, "NCCL error at %s:%d: %s", __FILE__, __LINE__,          \
                         nccl_api().GetErrorString(r_));                                         \
    } while (0)
#endif

inline void nccl_unique_id(unsigned char* id) {
#if CUPPEN_CUDA
    static_assert(sizeof(ncclUniqueId) == CUPPEN_NCCL_ID_BYTES, "NCCL unique id size");
    ncclUniqueId u;
    NCCL_CHECK(nccl_api().GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
#else
    (void)id;
    CUPPEN_THROW(CUPPEN_ERR_COMM, "NCCL is not available in the host test build");
#endif
}

struct Comm {
    int rank = 0, world = 1;
    bool use_cb = false;
    cuppen_comm_callbacks cb{};
#if CUPPEN_CUDA
    ncclComm_t nccl = nullptr;
#endif

    void init_nccl(const unsigned char* id) {
#if CUPPEN_CUDA
        ncclUniqueId u;
        memcpy(&u, id, sizeof u);
        NCCL_CHECK(nccl_api().CommInitRank(&nccl, world, u, rank));
#else
        (void)id;
        CUPPEN_THROW(CUPPEN_ERR_COMM, "NCCL is not available in the host test build");
#endif
    }
    void destroy() {
#if CUPPEN_CUDA
        if (nccl) { nccl_api().CommDestroy(nccl); nccl = nullptr; }
#endif
    }
    void check_cb(int rc, const char* what) {
        if (rc != 0) CUPPEN_THROW(CUPPEN_ERR_COMM, "communication callback %s failed (%d)", what, rc);
    }
    // broadcast `bytes` from `root` to the contiguous rank group [lo, lo+cnt); ranks outside: no-op
    void group_bcast(void* buf, size_t bytes, int root, int lo, int cnt, Stream s) {
        if (cnt <= 1 || bytes == 0 || rank < lo || rank >= lo + cnt) return;
        if (use_cb) { dev_sync(s); check_cb(cb.group_bcast(cb.user, buf, bytes, root, lo, cnt), "group_bcast"); return; }
#if CUPPEN_CUDA
        NcclApi& api = nccl_api();
        NCCL_CHECK(api.GroupStart());
        if (rank == root) {
            for (int r = lo; r < lo + cnt; ++r)
                if (r != root) NCCL_CHECK(api.Send(buf, bytes, ncclInt8, r, nccl, s));
        } else {
            NCCL_CHECK(api.Recv(buf, bytes, ncclInt8, root, nccl, s));
        }
        NCCL_CHECK(api.GroupEnd());
#else
        CUPPEN_THROW(CUPPEN_ERR_COMM, "no communicator");
#endif
    }
    void allreduce_sum(double* buf, size_t count, Stream s) {
        if (world <= 1) return;
        if (use_cb) { dev_sync(s); check_cb(cb.allreduce_sum_f64(cb.user, buf, count), "allreduce"); return; }
#if CUPPEN_CUDA
        NCCL_CHECK(nccl_api().AllReduce(buf, buf, count, ncclFloat64, ncclSum, nccl, s));
#else
        CUPPEN_THROW(CUPPEN_ERR_COMM, "no communicator");
#endif
    }
    void allreduce_sum_i32(int* buf, size_t count, Stream s) {
        if (world <= 1) return;
        if (use_cb) { dev_sync(s); check_cb(cb.allreduce_sum_i32(cb.user, buf, count), "allreduce_i32"); return; }
#if CUPPEN_CUDA
        NCCL_CHECK(nccl_api().AllReduce(buf, buf, count, ncclInt32, ncclSum, nccl, s));
#else
        CUPPEN_THROW(CUPPEN_ERR_COMM, "no communicator");
#endif
    }
    // personalised exchange of byte buffers (entry `rank` is ignored)
    void alltoallv(const std::vector<const void*>& send, const std::vector<size_t>& sbytes, const std::vector<void*>& recv,
                   const std::vector<size_t>& rbytes, Stream s) {
        if (world <= 1) return;
        if (use_cb) {
            dev_sync(s);
            check_cb(cb.alltoallv(cb.user, send.data(), sbytes.data(), recv.data(), rbytes.data()), "alltoallv");
            return;
        }
#if CUPPEN_CUDA
        NcclApi& api = nccl_api();
        NCCL_CHECK(api.GroupStart());
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            if (sbytes[r]) NCCL_CHECK(api.Send(send[r], sbytes[r], ncclInt8, r, nccl, s));
            if (rbytes[r]) NCCL_CHECK(api.Recv(recv[r], rbytes[r], ncclInt8, r, nccl, s));
        }
        NCCL_CHECK(api.GroupEnd());
#else
        CUPPEN_THROW(CUPPEN_ERR_COMM, "no communicator");
#endif
    }
    void allgather(const void* send, void* recv, size_t bytes, Stream s) {
        if (world <= 1) { dev_d2d(recv, send, bytes, s); return; }
        if (use_cb) { dev_sync(s); check_cb(cb.allgather(cb.user, send, recv, bytes), "allgather"); return; }
#if CUPPEN_CUDA
        NCCL_CHECK(nccl_api().AllGather(send, recv, bytes, ncclInt8, nccl, s));
#else
        CUPPEN_THROW(CUPPEN_ERR_COMM, "no communicator");
#endif
    }
};

}  // namespace cuppen
#endif
