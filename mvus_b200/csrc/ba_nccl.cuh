// Multi-GPU plumbing: one process per GPU (torch.distributed launches them);
// this library joins an NCCL communicator and sums the per-rank normal equations and costs.
// NCCL is dlopen'ed (libnccl.so.2: the copy torch already loaded in-process, else the system
// one) so that the library itself loads on machines without NCCL and single-GPU use never
// touches it.
//
// Detections are sharded across ranks (the caller's choice of shard; mvus_b200/shard.py cuts by
// time), every rank accumulates its partial A, bc, D, E, W~; the camera blocks are all-reduced,
// the spline-side arrays are reduced to the rank that owns the block range in the SHARDED exact
// solve (ba_solve.cuh, solve_damped: local cyclic-reduction levels + a small top system), and
// the trial costs are all-reduced so that every rank takes the same accept / reject decision.
#pragma once
#include <dlfcn.h>
#include "ba_ctx.cuh"

namespace mvus {

typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*fn_uid)(nccl_uid_t*);
typedef int (*fn_init)(void**, int, nccl_uid_t, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef const char* (*fn_errstr)(int);
typedef int (*fn_reduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);

struct NcclApi {
    void* lib = nullptr;
    fn_uid GetUniqueId = nullptr;
    fn_init CommInitRank = nullptr;
    fn_destroy CommDestroy = nullptr;
    fn_allreduce AllReduce = nullptr;
    fn_errstr GetErrorString = nullptr;
    fn_reduce Reduce = nullptr;
    fn_group GroupStart = nullptr, GroupEnd = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
        GetUniqueId = (fn_uid)dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (fn_init)dlsym(lib, "ncclCommInitRank");
        CommDestroy = (fn_destroy)dlsym(lib, "ncclCommDestroy");
        AllReduce = (fn_allreduce)dlsym(lib, "ncclAllReduce");
        GetErrorString = (fn_errstr)dlsym(lib, "ncclGetErrorString");
        Reduce = (fn_reduce)dlsym(lib, "ncclReduce");
        GroupStart = (fn_group)dlsym(lib, "ncclGroupStart");
        GroupEnd = (fn_group)dlsym(lib, "ncclGroupEnd");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
inline NcclApi& nccl_api() { static NcclApi a; return a; }
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;   // ncclDouble, ncclSum (stable ABI values)

// One communicator per process and unique id (an id can initialise only one communicator);
// handles share it and it lives until the process exits.
struct CommCache { std::string uid; void* comm = nullptr; int world = 0, rank = 0; };
inline CommCache& comm_cache() { static CommCache c; return c; }

inline void nccl_destroy(mvus_ba_ctx* h) { h->nccl_comm = nullptr; }

inline int nccl_sum(mvus_ba_ctx* h, double* buf, size_t count) {
    if (h->world <= 1 || count == 0) return MVUS_OK;
    const int rc = nccl_api().AllReduce(buf, buf, count, NCCL_FLOAT64, NCCL_SUM, h->nccl_comm, h->st);
    if (rc != 0) return fail(h, MVUS_ERR_NCCL, std::string("ncclAllReduce: ") +
                             (nccl_api().GetErrorString ? nccl_api().GetErrorString(rc) : "error"));
    return MVUS_OK;
}

// block range [lo, hi) of super-blocks owned by rank r in the sharded solve
inline void owner_range(const mvus_ba_ctx* h, int r, int64_t* lo, int64_t* hi) {
    const int64_t nchunks = h->nb / h->Bc;
    *lo = nchunks * r / h->world * h->Bc;
    *hi = nchunks * (r + 1) / h->world * h->Bc;
}

// Sum the per-rank normal equations.  Camera blocks (small) are all-reduced.  The spline-side
// arrays D, E, W~ (2.8 GB at config 4) are only needed by the rank that owns the block range in
// the sharded solve, so each range is REDUCED TO ITS OWNER (grouped ncclReduce: half the wire
// traffic of an all-reduce); `full` forces an all-reduce (diagnostic entry point).
inline int reduce_normal_equations(mvus_ba_ctx* h, bool full) {
    if (h->world <= 1) return MVUS_OK;
    const size_t qq = (size_t)h->q * h->q, wn = (size_t)h->q * h->ldw;
    int rc = nccl_sum(h, h->A.p, (size_t)h->nc * h->Pc * h->Pc + h->ncP);
    if (rc) return rc;
    NcclApi& api = nccl_api();
    if (full || !api.Reduce || !api.GroupStart || !api.GroupEnd) {
        rc = nccl_sum(h, h->D.p, h->nb * qq);
        if (!rc) rc = nccl_sum(h, h->E.p, h->nb * qq);
        if (!rc) rc = nccl_sum(h, h->Wp(), (size_t)h->nb * wn);
        return rc;
    }
    int e = api.GroupStart();
    for (int r = 0; r < h->world && e == 0; ++r) {
        int64_t lo, hi;
        owner_range(h, r, &lo, &hi);
        if (hi <= lo) continue;
        const size_t nblk = (size_t)(hi - lo);
        e = api.Reduce(h->D.p + lo * qq, h->D.p + lo * qq, nblk * qq, NCCL_FLOAT64, NCCL_SUM, r, h->nccl_comm, h->st);
        if (!e) e = api.Reduce(h->E.p + lo * qq, h->E.p + lo * qq, nblk * qq, NCCL_FLOAT64, NCCL_SUM, r, h->nccl_comm, h->st);
        if (!e) e = api.Reduce(h->Wp() + lo * wn, h->Wp() + lo * wn, nblk * wn, NCCL_FLOAT64, NCCL_SUM, r, h->nccl_comm, h->st);
    }
    const int e2 = api.GroupEnd();
    if (e || e2) return fail(h, MVUS_ERR_NCCL, "grouped ncclReduce of the normal equations failed");
    return MVUS_OK;
}

// Make rank 0's copy of a buffer the copy of every rank (zero elsewhere + sum): the Cholesky of
// the reduced system is computed redundantly from split-K partial sums whose order is not fixed,
// so it could differ in the last bits between ranks; broadcasting keeps x and every
// accept/reject decision identical.
inline int nccl_bcast0(mvus_ba_ctx* h, double* buf, size_t count) {
    if (h->world <= 1 || count == 0) return MVUS_OK;
    if (h->rank != 0) {
        cudaError_t e = cudaMemsetAsync(buf, 0, count * sizeof(double), h->st);
        if (e != cudaSuccess) return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e));
    }
    return nccl_sum(h, buf, count);
}
inline int nccl_max_flag(mvus_ba_ctx* h, int* flag) {
    if (h->world <= 1) return MVUS_OK;
    const int rc = nccl_api().AllReduce(flag, flag, 1, 2 /*ncclInt32*/, 2 /*ncclMax*/, h->nccl_comm, h->st);
    if (rc != 0) return fail(h, MVUS_ERR_NCCL, "ncclAllReduce(flag) failed");
    return MVUS_OK;
}

// sum of squares lives at partial[cost_slot]; make it the global sum
inline int allreduce_cost_slot(mvus_ba_ctx* h) {
    return nccl_sum(h, h->partial.p + h->cost_slot, 1);
}

}  // namespace mvus

extern "C" int mvus_ba_nccl_unique_id(char id_out[128]) {
    std::string err;
    if (!mvus::nccl_api().load(err)) return MVUS_ERR_NCCL;
    mvus::nccl_uid_t id;
    if (mvus::nccl_api().GetUniqueId(&id) != 0) return MVUS_ERR_NCCL;
    memcpy(id_out, id.internal, 128);
    return MVUS_OK;
}

extern "C" int mvus_ba_comm_init(mvus_ba_handle h, int32_t world_size, int32_t rank, const char id[128]) {
    if (!h || !id || world_size < 1 || rank < 0 || rank >= world_size) return mvus::fail(h, MVUS_ERR_ARG, "bad argument");
    if (world_size != h->world || rank != h->rank) mvus::invalidate_solver(h);   // block padding depends on the world size
    if (world_size == 1) { h->world = 1; h->rank = 0; return MVUS_OK; }
    std::string err;
    if (!mvus::nccl_api().load(err)) return mvus::fail(h, MVUS_ERR_NCCL, err);
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    mvus::CommCache& cc = mvus::comm_cache();
    const std::string key(id, 128);
    if (!(cc.comm && cc.uid == key && cc.world == world_size && cc.rank == rank)) {
        mvus::nccl_uid_t uid;
        memcpy(uid.internal, id, 128);
        void* comm = nullptr;
        const int rc = mvus::nccl_api().CommInitRank(&comm, world_size, uid, rank);
        if (rc != 0) return mvus::fail(h, MVUS_ERR_NCCL, std::string("ncclCommInitRank failed: ") +
                                       (mvus::nccl_api().GetErrorString ? mvus::nccl_api().GetErrorString(rc) : ""));
        cc.uid = key; cc.comm = comm; cc.world = world_size; cc.rank = rank;
    }
    h->nccl_comm = cc.comm;
    h->world = world_size;
    h->rank = rank;
    return MVUS_OK;
}
