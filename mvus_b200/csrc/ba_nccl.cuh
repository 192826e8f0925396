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
#include <climits>
#include <chrono>
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
// the sharded solve, and a rank only has non-zero rows where its own detections (and the motion
// rows it owns) put them: K2 records the smallest / largest block of four spans it saw, the ranks
// exchange those TOUCHED ranges (one tiny all-reduce), and for every owner s only the hull of
// "rows touched by another rank inside s's range" is reduced to s (grouped ncclReduce; ranks
// without a contribution there add zeros).  With detections sharded by time along the owner
// ranges (mvus_b200/shard.py, mvus_ba_shard_bounds) that hull is the 3-control-point halo at the
// range boundaries; with any other sharding it grows up to the whole range and stays correct.
// `full` forces an all-reduce of everything (diagnostic entry point).
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
    // ---- touched super-block range of every rank: [tlo, thi] (inclusive), empty if tlo > thi
    const int W = h->world;
    const auto t_start = std::chrono::steady_clock::now();
    auto ms_since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    };
    std::vector<int> mine(2 * W, INT_MIN), all(2 * W, INT_MIN);
    {
        int kq[3] = {0, 0x7fffffff, -1};
        if (h->n_chunks > 0 && h->k2_queue.p)
            MV_CUDA(h, cudaMemcpyAsync(kq, h->k2_queue.p, sizeof(kq), cudaMemcpyDeviceToHost, h->st));
        MV_CUDA(h, cudaStreamSynchronize(h->st));
        if (h->verbose) fprintf(stderr, "[mvus_ba] rank %d: K2 + camera all-reduce done after %.3f ms (host wait)\n", h->rank, ms_since(t_start));
        int64_t tlo = h->nb, thi = -1;
        if (kq[2] >= 0) {                                  // K2: control points 4 bmin - 3 .. 4 bmax + 3
            tlo = std::max<int64_t>(0, (4 * (int64_t)kq[1] - 3)) / h->bw;
            thi = std::min<int64_t>(h->nb - 1, (4 * (int64_t)kq[2] + 3) / h->bw);
        }
        if (h->M > 0) {                                    // motion rows of the own range reach <= 6 control points further
            int64_t lo, hi;
            owner_range(h, h->rank, &lo, &hi);
            if (hi > lo) { tlo = std::min(tlo, lo); thi = std::max(thi, std::min<int64_t>(h->nb - 1, hi + 6 / h->bw + 1)); }
        }
        mine[2 * h->rank] = -(int)tlo;
        mine[2 * h->rank + 1] = (int)thi;
    }
    MV_CUDA(h, h->touch.alloc(2 * W));
    MV_CUDA(h, cudaMemcpyAsync(h->touch.p, mine.data(), 2 * W * sizeof(int), cudaMemcpyHostToDevice, h->st));
    if (api.AllReduce(h->touch.p, h->touch.p, 2 * W, 2 /*ncclInt32*/, 2 /*ncclMax*/, h->nccl_comm, h->st) != 0)
        return fail(h, MVUS_ERR_NCCL, "ncclAllReduce(touched ranges) failed");
    MV_CUDA(h, cudaMemcpyAsync(all.data(), h->touch.p, 2 * W * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    int e = api.GroupStart();
    int64_t moved = 0;
    for (int s = 0; s < W && e == 0; ++s) {
        int64_t lo, hi;
        owner_range(h, s, &lo, &hi);
        if (s == W - 1) hi = h->nb;
        // hulls of the foreign touched blocks inside [lo, hi): one from the ranks before s (they reach into the
        // low end), one from the ranks after s (high end); merged if they meet
        int64_t seg[2][2] = {{hi, lo}, {hi, lo}};
        for (int r = 0; r < W; ++r) {
            if (r == s) continue;
            const int64_t tlo = -(int64_t)all[2 * r], thi = all[2 * r + 1];
            if (tlo > thi) continue;
            const int64_t a = std::max(tlo, lo), b = std::min(thi + 1, hi);
            int64_t* g = seg[r < s ? 0 : 1];
            if (a < b) { g[0] = std::min(g[0], a); g[1] = std::max(g[1], b); }
        }
        if (seg[0][0] < seg[0][1] && seg[1][0] < seg[1][1] && seg[1][0] <= seg[0][1]) {
            seg[0][0] = std::min(seg[0][0], seg[1][0]); seg[0][1] = std::max(seg[0][1], seg[1][1]);
            seg[1][0] = hi; seg[1][1] = lo;
        }
        for (int k = 0; k < 2 && e == 0; ++k) {
            const int64_t hlo = seg[k][0], hhi = seg[k][1];
            if (hhi <= hlo) continue;
            const size_t nblk = (size_t)(hhi - hlo);
            moved += (int64_t)nblk;
            e = api.Reduce(h->D.p + hlo * qq, h->D.p + hlo * qq, nblk * qq, NCCL_FLOAT64, NCCL_SUM, s, h->nccl_comm, h->st);
            if (!e) e = api.Reduce(h->E.p + hlo * qq, h->E.p + hlo * qq, nblk * qq, NCCL_FLOAT64, NCCL_SUM, s, h->nccl_comm, h->st);
            if (!e) e = api.Reduce(h->Wp() + hlo * wn, h->Wp() + hlo * wn, nblk * wn, NCCL_FLOAT64, NCCL_SUM, s, h->nccl_comm, h->st);
        }
    }
    const int e2 = api.GroupEnd();
    h->halo_blocks = moved;
    if (h->verbose) {
        cudaStreamSynchronize(h->st);
        fprintf(stderr, "[mvus_ba] rank %d: halo exchange of %lld of %lld super-blocks done after %.3f ms; touched [%d, %d]\n",
                h->rank, (long long)moved, (long long)h->nb, ms_since(t_start), -all[2 * h->rank], all[2 * h->rank + 1]);
    }
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


// Control-point bounds of the block ranges the ranks of a `world_size`-GPU solve own (bounds[world_size + 1],
// bounds[0] = 0, bounds[world_size] = number of control points): detections whose knot span falls in
// [bounds[r], bounds[r+1]) belong on rank r (mvus_b200/shard.py), so that the normal-equation exchange only moves
// the halo rows at the range boundaries.  Needs mvus_ba_set_splines only.
extern "C" int mvus_ba_shard_bounds(mvus_ba_handle h, int32_t world_size, int64_t* bounds) {
    if (!h || !bounds || world_size < 1) return mvus::fail(h, MVUS_ERR_ARG, "bad argument");
    if (!h->have_spl) return mvus::fail(h, MVUS_ERR_ARG, "set_splines first");
    int bw = 3;
    int64_t nb = 0, Bc = 1;
    const int rc = mvus::solver_dims(h, world_size, &bw, &nb, &Bc);
    if (rc) return rc;
    const int64_t nchunks = nb / Bc;
    for (int r = 0; r <= world_size; ++r)
        bounds[r] = std::min<int64_t>(h->n_ctrl, nchunks * r / world_size * Bc * bw);
    bounds[world_size] = h->n_ctrl;
    return MVUS_OK;
}
