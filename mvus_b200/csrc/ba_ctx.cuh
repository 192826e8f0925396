// Context (handle) of the BA library, device-buffer helpers and error plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/mvus_ba.h"
#include "ba_math.cuh"
#include "ba_tables.hpp"

namespace mvus {

constexpr int TILE_DET = 128;      // detections per CTA tile (one camera per tile)

// Compact block-row Jacobian in HBM: one BLOCK per 32 detections (= one warp of a K1 tile, 4 per tile):
//   [2P+2 planes x 32 doubles | 32 span ints]; planes: u columns (P), v columns (P), r_u, r_v.
// Element (plane p, detection t) sits at jblk_off(p, t): the XOR permutes the 32-byte sectors of a
// 256-byte plane row (stores stay coalesced) so that K2's MMA fragment loads are bank-conflict free.
template <int P>
struct JBlk {
    static constexpr int NPL = 2 * P + 2;
    static constexpr int BLK_D = NPL * 32 + 16;       // doubles per block
    static constexpr int SPAN_OFF = NPL * 32;         // (in doubles) start of the 32 span ints
};
__host__ __device__ __forceinline__ int jblk_off(int plane, int t) { return plane * 32 + (t ^ ((plane & 3) << 2)); }
inline int jblk_doubles(int P) { return (2 * P + 2) * 32 + 16; }

// Device buffer on the stream-ordered allocator, from the library's OWN memory pool (one per device,
// created by mvus_ba_create; the process-wide default pool, which torch or any other cudaMallocAsync
// user shares, is left alone).  The pool keeps freed memory (release threshold = max), so the ~35 GB a
// config-4 handle needs are obtained from the driver once per process and re-used by every later BA call
// (the reference's main.py makes two BA calls per camera); plain cudaMalloc/cudaFree of that much memory
// costs several hundred ms per call.  mvus_ba_trim() hands the cached memory back to the driver.
cudaMemPool_t library_pool(int device);      // mvus_ba.cu

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) return cudaSuccess;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPool_t pool = library_pool(dev);
        cudaError_t e = pool ? cudaMallocFromPoolAsync((void**)&p, count * sizeof(T), pool, cudaStreamPerThread)
                             : cudaMallocAsync((void**)&p, count * sizeof(T), cudaStreamPerThread);
        if (e == cudaSuccess) e = cudaStreamSynchronize(cudaStreamPerThread);   // usable on any stream afterwards
        if (e == cudaSuccess) n = count; else p = nullptr;
        return e;
    }
    // sync = false: the caller has already made sure that nothing on any stream still uses the memory
    // (mvus_ba_destroy synchronises once instead of once per buffer)
    void release(bool sync = true) {
        if (p) {
            if (sync) cudaDeviceSynchronize();       // nothing on any stream may still use it
            cudaFreeAsync(p, cudaStreamPerThread);
        }
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
};

// Page-locked scratch for the scalars a handle reads back (costs, step norms): cudaMallocHost / cudaFreeHost take
// milliseconds and serialise every thread of the process on the driver, which dominated the life of a small
// problem's handle (22 of 36 ms per 7 x 5000 problem: tests/cfg5_probe.py); blocks are recycled instead.
double* pin_scratch_acquire();               // mvus_ba.cu: 64 doubles
void pin_scratch_release(double* p);

struct NcclApi;   // ba_nccl.cuh
struct LmHooks;   // mvus_ba.cu: evaluation / accumulation overrides of the LM driver (points mode)

}  // namespace mvus

struct mvus_ba_ctx {
    mvus_ba_desc desc;
    std::string err;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t evs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // sub-phase timers of the solve / accumulate
    double ms_syrk = 0.0, ms_bcr = 0.0, ms_reduce = 0.0, ms_k2 = 0.0;
    int sm_count = 148;

    // sizes
    int nc = 0, C = 6, Pc = 9, P = 21;
    int64_t N = 0, M = 0, n = 0, m = 0, n_other = 0, n_ctrl = 0;
    bool have_det = false, have_spl = false;
    int motion_spread = 4;        // consecutive control points a motion-prior row touches (ba_tables.hpp)

    // detections
    std::vector<int64_t> cam_ptr;
    mvus::DevBuf<double> frame, xr, yr, obs_u, obs_v, calib, height;
    mvus::DevBuf<int64_t> row_off;
    int n_tiles = 0;
    mvus::DevBuf<int> tile_cam, tile_cnt;
    mvus::DevBuf<int64_t> tile_start;

    // splines (static tables)
    mvus::HostSplineTables T;
    mvus::DevBuf<double> int_a, int_b, knots, spanpoly, span_t0, lut_t0, lut_invh;
    mvus::DevBuf<int64_t> knot_off, ctrl_off, xoff, lut_off;
    mvus::DevBuf<int> ncoef, deg, lut_n, lut;
    mvus::SplineView sv;

    // motion samples
    mvus::DevBuf<double> tau;
    mvus::DevBuf<int> tau_spl;
    mvus::DevBuf<unsigned char> tau_flag;

    // per-evaluation state
    mvus::DevBuf<double> x, x_trial, camprep, r, J, mJ, partial, scratch, gt_out;
    mvus::DevBuf<int> span, mbase, flag, frozen;
    // K2 work list: chunks of up to 4 consecutive tiles of one camera, visited in time order
    int n_chunks = 0;
    bool chunk_sorted = false;          // chunk_perm is valid for the current inputs / start point (accumulate)
    mvus::DevBuf<int> chunk_tile0, chunk_nt, chunk_key, chunk_key2, chunk_id, chunk_perm, k2_queue;
    mvus::DevBuf<unsigned char> sort_tmp;
    double* h_pin = nullptr;      // pinned scratch for scalars
    size_t h_pin_n = 0;

    // normal equations / solver (ba_solve.cuh)
    int bw = 3, q = 9;            // control points per super-block, unknowns per super-block
    int64_t nb = 0;               // super-blocks
    int ncP = 0, ldw = 0;         // camera unknowns, leading dimension of W~ (ncP + 1 rhs column)
    mvus::DevBuf<double> A, D, E, W, Dw, Ew, Ww, ZL, Sd, dlt_c, dlt_s, diag_c, diag_s, gvec, xs;
    mvus::DevBuf<double> Hb;             // K2's control x control band array (ba_k2.cuh, band_to_blocks_kernel)
    mvus::DevBuf<double> bs;             // spline right-hand side (-J^T r) as a contiguous vector
    mvus::DevBuf<double> Dt, ZLt, dst;   // top-level system of the sharded solve
    int64_t Bc = 1;                      // chunk size (super-blocks) of the sharded solve
    // chunk pre-reduction (ba_chunk.cuh): Lc super-blocks per chunk, nh = ceil(nb / Lc) heads (+1 ghost)
    int Lc = 1, Lc_req = 0;              // Lc_req: desc.solver_chunk (0 = choose from nb)
    int64_t nh = 0;
    mvus::DevBuf<double> Dh, Eh, Wh, ZLh, dsh, DhR, Gh, Linv;
    size_t w_guard = 0;                  // doubles in front of W~ inside the allocation W (K2's guard rows)
    double* Wp() const { return W.p ? W.p + w_guard : nullptr; }
    int launches = 0;
    // CUDA graph of the linear solve's launch sequence (small single-GPU systems, ba_solve.cuh: solve_damped):
    // 0 = not tried, 1 = ran once directly, 2 = captured, -1 = capture failed / not used
    int graph_state = 0, graph_launches = 0;
    cudaGraphExec_t solve_graph = nullptr;
    bool verbose = false;                // MVUS_BA_VERBOSE, read once at mvus_ba_create
    double band_lo = 0.5, band_hi = 1.5; // trust-region band (MVUS_BA_BAND_LO / _HI at create: experiments only)
    int64_t cost_slot = 0;        // index in `partial` where the last evaluation left sum r^2

    // Scene.BA(motion_prior=True) (ba_points.cuh): point meta data on the points handle, LM hooks, dense
    // cross-camera addend of the reduced camera system
    int64_t ptG = 0;
    mvus::DevBuf<int> pt_cam, pt_idx;
    mvus::DevBuf<double> pt_frame, pt_yH, pt_r, pt_J, Ax;
    mvus::LmHooks* hooks = nullptr;

    // multi-GPU
    int world = 1, rank = 0;
    mvus::DevBuf<int> touch;             // touched super-block ranges of all ranks (reduce_normal_equations)
    int64_t halo_blocks = 0;             // super-blocks moved by the last normal-equation exchange (diagnostic)
    void* nccl_comm = nullptr;
};

namespace mvus {

void invalidate_solver(mvus_ba_ctx* h);     // mvus_ba.cu

inline int fail(mvus_ba_ctx* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}

#define MV_CUDA(h, call)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess)                                                                \
            return mvus::fail(h, MVUS_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
    } while (0)

template <class T>
inline cudaError_t upload(DevBuf<T>& b, const T* src, size_t n, cudaStream_t st) {
    cudaError_t e = b.alloc(n > 0 ? n : 1);
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}
template <class T>
inline cudaError_t upload(DevBuf<T>& b, const std::vector<T>& v, cudaStream_t st) {
    return upload(b, v.data(), v.size(), st);
}

}  // namespace mvus
