// C ABI of the B200-native mvus bundle adjustment (include/mvus_ba.h).
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo.
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <cstdio>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <algorithm>

#include "ba_ctx.cuh"
#include "ba_kernels.cuh"
#include "ba_chunk.cuh"
#include "ba_solve.cuh"
#include "ba_nccl.cuh"
#include "ba_bookkeeping.cuh"
#include "spl_fit.cuh"
#include "align.cuh"
#include "ba_points.cuh"
#include <functional>

using namespace mvus;

static std::string g_create_err;

static inline double __longlong_as_double_host(double bits_as_double) {
    // step_dots_kernel stores max|g| with an integer atomicMax on the bit pattern; the slot is
    // read back as a double holding those same bits, so this is the identity.
    return bits_as_double;
}

// The library's private stream-ordered memory pool (one per device): freed buffers stay cached for the next
// handle, the process-wide default pool is not touched.
cudaMemPool_t mvus::library_pool(int device) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    if (device < 0 || device >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (!pools[device]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        pools[device] = pool;
    }
    return pools[device];
}

// Hand the device memory cached by destroyed handles back to the driver (keeps `keep_bytes`).
extern "C" int mvus_ba_trim(int32_t device, uint64_t keep_bytes) {
    cudaMemPool_t pool = mvus::library_pool(device);
    if (!pool) return MVUS_ERR_CUDA;
    cudaDeviceSynchronize();
    return cudaMemPoolTrimTo(pool, (size_t)keep_bytes) == cudaSuccess ? MVUS_OK : MVUS_ERR_CUDA;
}

extern "C" const char* mvus_ba_version(void) { return "mvus-b200 0.1 (sm_100a, fp64)"; }

extern "C" const char* mvus_ba_last_error(mvus_ba_handle h) {
    return h ? h->err.c_str() : g_create_err.c_str();
}

extern "C" int mvus_ba_create(const mvus_ba_desc* desc, mvus_ba_handle* out) {
    if (!desc || !out) { g_create_err = "null argument"; return MVUS_ERR_ARG; }
    *out = nullptr;
    if (desc->num_cams < 1) { g_create_err = "num_cams must be >= 1"; return MVUS_ERR_ARG; }
    if (desc->motion_type < 0 || desc->motion_type > 2) { g_create_err = "bad motion_type"; return MVUS_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("no usable CUDA device (there is no CPU fallback): ") +
                       (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return MVUS_ERR_CUDA;
    }
    if (desc->device < 0 || desc->device >= ndev) { g_create_err = "bad device ordinal"; return MVUS_ERR_ARG; }
    e = cudaSetDevice(desc->device);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); return MVUS_ERR_CUDA; }
    if (!mvus::library_pool(desc->device)) cudaGetLastError();      // (falls back to the default pool)
    mvus_ba_ctx* h = new mvus_ba_ctx();
    h->verbose = getenv("MVUS_BA_VERBOSE") != nullptr;              // diagnostics: read here, once, not inside the solve
    if (getenv("MVUS_BA_BAND_LO")) h->band_lo = atof(getenv("MVUS_BA_BAND_LO"));
    if (getenv("MVUS_BA_BAND_HI")) h->band_hi = atof(getenv("MVUS_BA_BAND_HI"));
    h->Lc_req = desc->solver_chunk;
    h->desc = *desc;
    h->nc = desc->num_cams;
    h->C = desc->opt_calib ? 15 : 6;
    h->Pc = 3 + h->C;
    h->P = h->Pc + 12;
    h->n_other = (int64_t)h->nc * h->Pc;
    e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e == cudaSuccess)
    {
        for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreate(&h->ev[k]);
        for (int k = 0; k < 6 && e == cudaSuccess; ++k) e = cudaEventCreate(&h->evs[k]);
    }
    if (e == cudaSuccess) { h->h_pin_n = 64; h->h_pin = pin_scratch_acquire(); if (!h->h_pin) e = cudaErrorMemoryAllocation; }
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, desc->device);
    if (e != cudaSuccess) { g_create_err = cudaGetErrorString(e); delete h; return MVUS_ERR_CUDA; }
    *out = h;
    return MVUS_OK;
}

namespace {
std::mutex g_pin_mu;
std::vector<double*> g_pin_free;
}
double* mvus::pin_scratch_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pin_mu);
        if (!g_pin_free.empty()) { double* p = g_pin_free.back(); g_pin_free.pop_back(); return p; }
    }
    double* p = nullptr;
    if (cudaMallocHost((void**)&p, 64 * sizeof(double)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void mvus::pin_scratch_release(double* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_pin_mu);
    g_pin_free.push_back(p);
}

extern "C" void mvus_ba_destroy(mvus_ba_handle h) {
    if (!h) return;
    cudaSetDevice(h->desc.device);
    nccl_destroy(h);
    // everything that used this handle's buffers ran on its own stream (uploads and the points mode join it before
    // they return), so waiting for that stream -- not for the whole device, which would stall every other handle
    // of the process -- is enough before the buffers go back to the pool
    if (h->st) cudaStreamSynchronize(h->st);
    cudaStreamSynchronize(cudaStreamPerThread);
    for (auto* b : {&h->frame, &h->xr, &h->yr, &h->obs_u, &h->obs_v, &h->calib, &h->height, &h->int_a,
                    &h->int_b, &h->knots, &h->spanpoly, &h->span_t0, &h->lut_t0, &h->lut_invh, &h->tau,
                    &h->x, &h->x_trial, &h->camprep, &h->r, &h->J, &h->mJ, &h->partial, &h->A, &h->D, &h->E,
                    &h->W, &h->Dw, &h->Ew, &h->Ww, &h->ZL, &h->Sd, &h->dlt_c, &h->dlt_s, &h->diag_c,
                    &h->diag_s, &h->gvec, &h->xs, &h->scratch, &h->gt_out, &h->Dt, &h->ZLt, &h->dst, &h->bs, &h->Hb,
                    &h->Dh, &h->Eh, &h->Wh, &h->ZLh, &h->dsh, &h->DhR, &h->Gh, &h->Linv,
                    &h->pt_frame, &h->pt_yH, &h->pt_r, &h->pt_J, &h->Ax})
        b->release(false);
    for (auto* b : {&h->row_off, &h->tile_start, &h->knot_off, &h->ctrl_off, &h->xoff, &h->lut_off}) b->release(false);
    for (auto* b : {&h->tile_cam, &h->tile_cnt, &h->ncoef, &h->deg, &h->lut_n, &h->lut, &h->tau_spl, &h->span,
                    &h->mbase, &h->flag, &h->frozen, &h->chunk_tile0, &h->chunk_nt, &h->chunk_key, &h->chunk_key2,
                    &h->chunk_id, &h->chunk_perm, &h->k2_queue, &h->touch, &h->pt_cam, &h->pt_idx})
        b->release(false);
    h->tau_flag.release(false);
    h->sort_tmp.release(false);
    if (h->solve_graph) cudaGraphExecDestroy(h->solve_graph);
    pin_scratch_release(h->h_pin);
    for (int k = 0; k < 8; ++k) if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    for (int k = 0; k < 6; ++k) if (h->evs[k]) cudaEventDestroy(h->evs[k]);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

// ------------------------------------------------------------------------------------------
// Host -> device copy of the detections.  The caller's arrays are pageable NumPy memory: one
// cudaMemcpyAsync stream moves them at ~9 GB/s (the driver stages through one bounce buffer), which
// was 167 ms of a 0.77 s config-4 call.  Large uploads are therefore cut into 8 MB pieces and staged by
// UP_THREADS host threads, each with its own stream and two page-locked buffers (allocated once per
// process): memcpy into pinned memory in parallel, DMA out of it.
struct UploadSeg { double* dst; const double* src; size_t bytes; };
namespace {
constexpr int UP_THREADS = 4;
constexpr size_t UP_CHUNK = 8u << 20;
struct UploadPool {
    void* pin[UP_THREADS][2] = {};
    cudaStream_t st[UP_THREADS] = {};
    cudaEvent_t ev[UP_THREADS][2] = {};
    int device = -1;
    bool ok = false;
    bool init(int dev) {
        if (ok && device == dev) return true;
        if (ok) return false;                             // one device per process (one process per GPU)
        for (int t = 0; t < UP_THREADS; ++t) {
            if (cudaStreamCreateWithFlags(&st[t], cudaStreamNonBlocking) != cudaSuccess) return false;
            for (int k = 0; k < 2; ++k) {
                if (cudaHostAlloc(&pin[t][k], UP_CHUNK, cudaHostAllocDefault) != cudaSuccess) return false;
                if (cudaEventCreateWithFlags(&ev[t][k], cudaEventDisableTiming) != cudaSuccess) return false;
            }
        }
        device = dev; ok = true;
        return true;
    }
};
UploadPool g_upload;
std::mutex g_upload_mutex;
}  // namespace

static cudaError_t upload_segments(int device, const std::vector<UploadSeg>& segs, cudaStream_t st) {
    size_t total = 0;
    for (const auto& s : segs) total += s.bytes;
    std::unique_lock<std::mutex> lock(g_upload_mutex, std::defer_lock);
    const bool threaded = total >= (64u << 20) && lock.try_lock() && g_upload.init(device);
    if (!threaded) {
        cudaGetLastError();
        for (const auto& s : segs) {
            cudaError_t e = cudaMemcpyAsync(s.dst, s.src, s.bytes, cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
    // pieces of at most UP_CHUNK bytes, dealt round-robin to the threads
    struct Piece { char* dst; const char* src; size_t bytes; };
    std::vector<Piece> pieces;
    for (const auto& s : segs)
        for (size_t o = 0; o < s.bytes; o += UP_CHUNK)
            pieces.push_back({(char*)s.dst + o, (const char*)s.src + o, std::min(UP_CHUNK, s.bytes - o)});
    cudaError_t err[UP_THREADS];
    std::vector<std::thread> th;
    for (int t = 0; t < UP_THREADS; ++t)
        th.emplace_back([&, t]() {
            cudaError_t e = cudaSetDevice(device);
            int k = 0;
            for (size_t i = t; i < pieces.size() && e == cudaSuccess; i += UP_THREADS, k ^= 1) {
                e = cudaEventSynchronize(g_upload.ev[t][k]);              // the DMA out of this buffer is done
                if (e != cudaSuccess) break;
                memcpy(g_upload.pin[t][k], pieces[i].src, pieces[i].bytes);
                e = cudaMemcpyAsync(pieces[i].dst, g_upload.pin[t][k], pieces[i].bytes, cudaMemcpyHostToDevice, g_upload.st[t]);
                if (e == cudaSuccess) e = cudaEventRecord(g_upload.ev[t][k], g_upload.st[t]);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(g_upload.st[t]);
            err[t] = e;
        });
    for (auto& t : th) t.join();
    for (int t = 0; t < UP_THREADS; ++t)
        if (err[t] != cudaSuccess) return err[t];
    return cudaSuccess;
}

// ------------------------------------------------------------------------------------------
// The solver arrays (bw, q, nb, Bc, ldw and every D/E/W~ buffer) are sized from n_ctrl, M and the
// world size at the first solve.  Any later mvus_ba_set_detections / set_splines / comm_init must
// make solver_alloc run again, or K2 and the cyclic-reduction levels would use stale block counts.
void mvus::invalidate_solver(mvus_ba_ctx* h) {
    h->chunk_sorted = false;
    if (h->solve_graph) { cudaGraphExecDestroy(h->solve_graph); h->solve_graph = nullptr; }
    if (h->graph_state > 0) h->graph_state = 0;
    if (h->st) cudaStreamSynchronize(h->st);      // (the handle's own stream is the only user of these)
    h->A.release(false);
    h->W.release(false);
    h->J.release(false);
}

static int finish_dims(mvus_ba_ctx* h) {
    if (!(h->have_det && h->have_spl)) return MVUS_OK;
    h->n = h->n_other + 3 * h->n_ctrl;
    h->m = 2 * h->N + h->M;
    MV_CUDA(h, h->x.alloc(h->n));
    MV_CUDA(h, h->x_trial.alloc(h->n));
    MV_CUDA(h, h->camprep.alloc((size_t)h->nc * CAMPREP_DOUBLES));
    MV_CUDA(h, h->r.alloc(h->m > 0 ? h->m : 1));
    MV_CUDA(h, h->span.alloc(h->N > 0 ? h->N : 1));
    MV_CUDA(h, h->partial.alloc((size_t)h->n_tiles + (h->M + 127) / 128 + 8));
    MV_CUDA(h, h->flag.alloc(8));
    MV_CUDA(h, cudaMemsetAsync(h->flag.p, 0, 8 * sizeof(int), h->st));
    return MVUS_OK;
}

static int set_detections_core(mvus_ba_ctx* h, const int64_t* count, const double* const* frame,
                               const double* const* x, const double* const* y, const double* height,
                               const double* calib) {
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    const int nc = h->nc;
    std::vector<int64_t> cam_ptr(nc + 1, 0);
    for (int i = 0; i < nc; ++i) {
        if (count[i] < 0) return fail(h, MVUS_ERR_ARG, "negative detection count");
        cam_ptr[i + 1] = cam_ptr[i] + count[i];
    }
    h->cam_ptr = cam_ptr;
    h->N = cam_ptr[nc];
    if (h->N >= (int64_t)1 << 31) return fail(h, MVUS_ERR_UNSUPPORTED, "more than 2^31 detections per handle");
    const size_t na = h->N > 0 ? (size_t)h->N : 1;
    MV_CUDA(h, h->frame.alloc(na));
    MV_CUDA(h, h->xr.alloc(na));
    MV_CUDA(h, h->yr.alloc(na));
    {
        std::vector<UploadSeg> segs;
        for (int i = 0; i < nc; ++i) {
            if (count[i] == 0) continue;
            if (!frame[i] || !x[i] || !y[i]) return fail(h, MVUS_ERR_ARG, "null detection arrays");
            const size_t nb = (size_t)count[i] * sizeof(double);
            segs.push_back({h->frame.p + cam_ptr[i], frame[i], nb});
            segs.push_back({h->xr.p + cam_ptr[i], x[i], nb});
            segs.push_back({h->yr.p + cam_ptr[i], y[i], nb});
        }
        MV_CUDA(h, upload_segments(h->desc.device, segs, h->st));
    }
    MV_CUDA(h, upload(h->height, height, (size_t)nc, h->st));
    MV_CUDA(h, upload(h->calib, calib, (size_t)nc * 9, h->st));
    std::vector<int64_t> row_off(nc + 1);
    for (int i = 0; i <= nc; ++i) row_off[i] = 2 * cam_ptr[i];
    MV_CUDA(h, upload(h->row_off, row_off, h->st));
    // tiles: TILE_DET detections of one camera each
    std::vector<int> tcam, tcnt;
    std::vector<int64_t> tstart;
    for (int i = 0; i < nc; ++i)
        for (int64_t s = cam_ptr[i]; s < cam_ptr[i + 1]; s += TILE_DET) {
            tcam.push_back(i);
            tstart.push_back(s);
            tcnt.push_back((int)std::min<int64_t>(TILE_DET, cam_ptr[i + 1] - s));
        }
    h->n_tiles = (int)tcam.size();
    {   // K2 chunks: up to K2_CHUNK_TILES consecutive tiles of one camera
        std::vector<int> c0, cn;
        for (int t = 0; t < h->n_tiles;) {
            int e = t + 1;
            while (e < h->n_tiles && e - t < K2_CHUNK_TILES && tcam[e] == tcam[t]) ++e;
            c0.push_back(t); cn.push_back(e - t);
            t = e;
        }
        h->n_chunks = (int)c0.size();
        MV_CUDA(h, upload(h->chunk_tile0, c0, h->st));
        MV_CUDA(h, upload(h->chunk_nt, cn, h->st));
    }
    MV_CUDA(h, upload(h->tile_cam, tcam, h->st));
    MV_CUDA(h, upload(h->tile_start, tstart, h->st));
    MV_CUDA(h, upload(h->tile_cnt, tcnt, h->st));
    MV_CUDA(h, h->obs_u.alloc(na));
    MV_CUDA(h, h->obs_v.alloc(na));
    if (!h->desc.opt_calib && h->n_tiles > 0) {
        observe_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->tile_cam.p, h->tile_start.p, h->tile_cnt.p,
                                                          h->calib.p, h->desc.undist_points, h->xr.p, h->yr.p,
                                                          h->obs_u.p, h->obs_v.p);
        MV_CUDA(h, cudaGetLastError());
    }
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    h->have_det = true;
    invalidate_solver(h);
    return finish_dims(h);
}

extern "C" int mvus_ba_set_detections_rows(mvus_ba_handle h, const int64_t* count, const double* const* frame,
                                           const double* const* x, const double* const* y,
                                           const double* height, const double* calib) {
    if (!h || !count || !frame || !x || !y || !height || !calib) return fail(h, MVUS_ERR_ARG, "null argument");
    return set_detections_core(h, count, frame, x, y, height, calib);
}

extern "C" int mvus_ba_set_detections(mvus_ba_handle h, const int64_t* cam_ptr, const double* frame,
                                      const double* x, const double* y, const double* height,
                                      const double* calib) {
    if (!h || !cam_ptr || !height || !calib) return fail(h, MVUS_ERR_ARG, "null argument");
    const int nc = h->nc;
    if (cam_ptr[0] != 0) return fail(h, MVUS_ERR_ARG, "cam_ptr[0] must be 0");
    std::vector<int64_t> count(nc);
    std::vector<const double*> pf(nc), px(nc), py(nc);
    for (int i = 0; i < nc; ++i) {
        if (cam_ptr[i + 1] < cam_ptr[i]) return fail(h, MVUS_ERR_ARG, "cam_ptr must be non-decreasing");
        count[i] = cam_ptr[i + 1] - cam_ptr[i];
        if (count[i] > 0 && (!frame || !x || !y)) return fail(h, MVUS_ERR_ARG, "null detection arrays");
        pf[i] = frame ? frame + cam_ptr[i] : nullptr;
        px[i] = x ? x + cam_ptr[i] : nullptr;
        py[i] = y ? y + cam_ptr[i] : nullptr;
    }
    return set_detections_core(h, count.data(), pf.data(), px.data(), py.data(), height, calib);
}

extern "C" int mvus_ba_set_splines(mvus_ba_handle h, int32_t S, const double* interval,
                                   const int64_t* knot_ptr, const double* knots, const int32_t* degree) {
    if (!h || S < 1 || !interval || !knot_ptr || !knots || !degree) return fail(h, MVUS_ERR_ARG, "null argument");
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    if (!build_spline_tables(S, interval, knot_ptr, knots, degree, h->n_other, h->T))
        return fail(h, MVUS_ERR_ARG, "malformed spline (degree must be 1 or 3, >= degree+1 coefficients)");
    HostSplineTables& T = h->T;
    h->n_ctrl = T.n_ctrl;
    MV_CUDA(h, upload(h->int_a, T.int_a, h->st));
    MV_CUDA(h, upload(h->int_b, T.int_b, h->st));
    MV_CUDA(h, upload(h->knots, T.knots, h->st));
    MV_CUDA(h, upload(h->knot_off, T.knot_off, h->st));
    MV_CUDA(h, upload(h->ncoef, T.ncoef, h->st));
    MV_CUDA(h, upload(h->deg, T.deg, h->st));
    MV_CUDA(h, upload(h->ctrl_off, T.ctrl_off, h->st));
    MV_CUDA(h, upload(h->xoff, T.xoff, h->st));
    MV_CUDA(h, upload(h->spanpoly, T.spanpoly, h->st));
    MV_CUDA(h, upload(h->span_t0, T.span_t0, h->st));
    MV_CUDA(h, upload(h->lut_off, T.lut_off, h->st));
    MV_CUDA(h, upload(h->lut_n, T.lut_n, h->st));
    MV_CUDA(h, upload(h->lut_t0, T.lut_t0, h->st));
    MV_CUDA(h, upload(h->lut_invh, T.lut_invh, h->st));
    MV_CUDA(h, upload(h->lut, T.lut, h->st));
    h->sv = SplineView{S, h->int_a.p, h->int_b.p, h->knots.p, h->knot_off.p, h->ncoef.p, h->deg.p,
                       h->ctrl_off.p, h->xoff.p, h->spanpoly.p, h->span_t0.p, h->lut_off.p, h->lut_n.p,
                       h->lut_t0.p, h->lut_invh.p, h->lut.p};
    h->M = 0;
    h->motion_spread = 4;
    if (h->desc.motion_type != MVUS_MOTION_NONE) {
        std::vector<double> tau;
        std::vector<int> spl;
        std::vector<unsigned char> fl;
        build_motion_samples(T, tau, spl, fl);
        h->M = (int64_t)tau.size();
        h->motion_spread = motion_spread(T, tau, spl, fl, h->desc.motion_type == MVUS_MOTION_F);
        MV_CUDA(h, upload(h->tau, tau, h->st));
        MV_CUDA(h, upload(h->tau_spl, spl, h->st));
        MV_CUDA(h, upload(h->tau_flag, fl, h->st));
    }
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    h->have_spl = true;
    invalidate_solver(h);
    return finish_dims(h);
}

extern "C" int mvus_ba_dims(mvus_ba_handle h, int64_t* n, int64_t* m, int64_t* N, int64_t* M, int32_t* P) {
    if (!h) return MVUS_ERR_ARG;
    if (!(h->have_det && h->have_spl)) return fail(h, MVUS_ERR_ARG, "set_detections and set_splines first");
    if (n) *n = h->n;
    if (m) *m = h->m;
    if (N) *N = h->N;
    if (M) *M = h->M;
    if (P) *P = h->P;
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
// Evaluate residuals (and the Jacobian) at device vector xd.  2*cost lands in h->partial[np].
int mvus::evaluate(mvus_ba_ctx* h, const double* xd, bool want_j) {
    const int nc = h->nc;
    cam_prep_kernel<<<(nc + 63) / 64, 64, 0, h->st>>>(xd, nc, h->C, h->desc.opt_calib, h->calib.p,
                                                      h->height.p, h->camprep.p);
    h->launches++;
    if (want_j) {
        MV_CUDA(h, h->J.alloc((size_t)(h->n_tiles > 0 ? h->n_tiles : 1) * (TILE_DET / 32) * jblk_doubles(h->P)));
        if (h->M > 0) {
            MV_CUDA(h, h->mJ.alloc((size_t)10 * h->M));
            MV_CUDA(h, h->mbase.alloc((size_t)h->M));
        }
    }
    if (h->n_tiles > 0) {
#define MV_LAUNCH_K1(CAL, WJ)                                                                           \
    resjac_kernel<CAL, WJ><<<h->n_tiles, TILE_DET, 0, h->st>>>(                                         \
        h->sv, xd, h->camprep.p, h->tile_cam.p, h->tile_start.p, h->tile_cnt.p, h->row_off.p,           \
        h->frame.p, h->xr.p, h->yr.p, h->obs_u.p, h->obs_v.p, h->desc.undist_points, h->desc.opt_sync,  \
        h->desc.opt_rs, h->N, h->r.p, h->J.p, h->partial.p)
#define MV_LAUNCH_R2(CAL)                                                                               \
    residual2_kernel<CAL><<<(h->n_tiles + 1) / 2, TILE_DET, 0, h->st>>>(                                \
        h->sv, xd, h->camprep.p, h->tile_cam.p, h->tile_start.p, h->tile_cnt.p, h->row_off.p,           \
        h->frame.p, h->xr.p, h->yr.p, h->obs_u.p, h->obs_v.p, h->desc.undist_points, h->desc.opt_sync,  \
        h->desc.opt_rs, h->n_tiles, h->r.p, h->partial.p)
        if (h->desc.opt_calib) { if (want_j) MV_LAUNCH_K1(true, true); else MV_LAUNCH_R2(true); }
        else { if (want_j) MV_LAUNCH_K1(false, true); else MV_LAUNCH_R2(false); }
#undef MV_LAUNCH_R2
#undef MV_LAUNCH_K1
        h->launches++;
    }
    int64_t np = h->n_tiles;
    if (h->M > 0) {
        const int gb = (int)((h->M + 127) / 128);
        if (want_j)
            motion_kernel<true><<<gb, 128, 0, h->st>>>(h->sv, xd, h->desc.motion_type, h->desc.motion_weight,
                                                       h->tau.p, h->tau_spl.p, h->tau_flag.p, h->M,
                                                       h->r.p + 2 * h->N, h->mbase.p, h->mJ.p,
                                                       h->partial.p + np, h->flag.p);
        else
            motion_kernel<false><<<gb, 128, 0, h->st>>>(h->sv, xd, h->desc.motion_type, h->desc.motion_weight,
                                                        h->tau.p, h->tau_spl.p, h->tau_flag.p, h->M,
                                                        h->r.p + 2 * h->N, nullptr, nullptr,
                                                        h->partial.p + np, h->flag.p);
        if (h->world > 1 && h->rank != 0)     // motion rows count once in the all-reduced cost
            MV_CUDA(h, cudaMemsetAsync(h->partial.p + np, 0, (size_t)gb * sizeof(double), h->st));
        np += gb;
        h->launches++;
    }
    reduce_partial_kernel<<<1, 1024, 0, h->st>>>(h->partial.p, np, h->partial.p + np);
    h->launches++;
    h->cost_slot = np;
    MV_CUDA(h, cudaGetLastError());
    return MVUS_OK;
}

static int check_ready(mvus_ba_ctx* h) {
    if (!h) return MVUS_ERR_ARG;
    if (!(h->have_det && h->have_spl)) return fail(h, MVUS_ERR_ARG, "set_detections and set_splines first");
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    return MVUS_OK;
}

static int check_motion_flag(mvus_ba_ctx* h) {
    int f = 0;
    MV_CUDA(h, cudaMemcpyAsync(&f, h->flag.p, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    if (f) return fail(h, MVUS_ERR_UNSUPPORTED,
                       "a motion-prior row touches more than 7 consecutive control points (knots denser than the unit sample grid)");
    return MVUS_OK;
}

extern "C" int mvus_ba_residual(mvus_ba_handle h, const double* x, double* r) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x || !r) return fail(h, MVUS_ERR_ARG, "null argument");
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    rc = evaluate(h, h->x.p, false);
    if (rc) return rc;
    MV_CUDA(h, cudaMemcpyAsync(r, h->r.p, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    return MVUS_OK;
}

extern "C" int mvus_ba_residual_jacobian(mvus_ba_handle h, const double* x, double* r, int32_t* span,
                                         double* J, int32_t* mbase, double* mJ) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x) return fail(h, MVUS_ERR_ARG, "null argument");
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    rc = evaluate(h, h->x.p, true);
    if (rc) return rc;
    if (r) MV_CUDA(h, cudaMemcpyAsync(r, h->r.p, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if ((span || J) && h->N) {          // de-block into the documented layout (column planes J[2P][N])
        MV_CUDA(h, h->scratch.alloc((size_t)2 * h->P * h->N));
        deblock_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->J.p, h->P, h->tile_start.p, h->tile_cnt.p, h->N,
                                                          h->scratch.p, h->span.p);
        MV_CUDA(h, cudaGetLastError());
        if (span) MV_CUDA(h, cudaMemcpyAsync(span, h->span.p, h->N * sizeof(int), cudaMemcpyDeviceToHost, h->st));
        if (J) MV_CUDA(h, cudaMemcpyAsync(J, h->scratch.p, (size_t)2 * h->P * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    }
    if (mbase && h->M) MV_CUDA(h, cudaMemcpyAsync(mbase, h->mbase.p, h->M * sizeof(int), cudaMemcpyDeviceToHost, h->st));
    if (mJ && h->M) MV_CUDA(h, cudaMemcpyAsync(mJ, h->mJ.p, (size_t)10 * h->M * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    return check_motion_flag(h);
}

extern "C" int mvus_ba_detections_global(mvus_ba_handle h, const double* x, double* out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x || !out) return fail(h, MVUS_ERR_ARG, "null argument");
    if (h->N == 0) return MVUS_OK;
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    cam_prep_kernel<<<(h->nc + 63) / 64, 64, 0, h->st>>>(h->x.p, h->nc, h->C, h->desc.opt_calib, h->calib.p,
                                                         h->height.p, h->camprep.p);
    MV_CUDA(h, h->scratch.alloc((size_t)3 * h->N));
    det_global_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->camprep.p, h->tile_cam.p, h->tile_start.p,
                                                         h->tile_cnt.p, h->frame.p, h->xr.p, h->yr.p,
                                                         h->obs_u.p, h->obs_v.p, h->desc.opt_calib,
                                                         h->desc.undist_points, h->row_off.p, h->scratch.p);
    MV_CUDA(h, cudaGetLastError());
    MV_CUDA(h, cudaMemcpyAsync(out, h->scratch.p, (size_t)3 * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    return MVUS_OK;
}

extern "C" int mvus_ba_time_resjac(mvus_ba_handle h, const double* x, int32_t reps, double* ms_mean) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x || !ms_mean || reps < 1) return fail(h, MVUS_ERR_ARG, "bad argument");
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    rc = evaluate(h, h->x.p, true);     // warm-up + allocation
    if (rc) return rc;
    MV_CUDA(h, cudaEventRecord(h->ev[0], h->st));
    for (int k = 0; k < reps; ++k) {
        rc = evaluate(h, h->x.p, true);
        if (rc) return rc;
    }
    MV_CUDA(h, cudaEventRecord(h->ev[1], h->st));
    MV_CUDA(h, cudaEventSynchronize(h->ev[1]));
    float ms = 0.f;
    MV_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    *ms_mean = (double)ms / reps;
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
static int ensure_solver(mvus_ba_ctx* h) {
    if (h->A.p && h->W.p) return MVUS_OK;
    return solver_alloc(h);
}

// sums[] of step_dots_kernel -> host
static int step_scalars(mvus_ba_ctx* h, const double* xd, double out[5]) {
    MV_CUDA(h, cudaMemsetAsync(h->xs.p, 0, 8 * sizeof(double), h->st));
    const int64_t cnt = std::max<int64_t>(h->n, std::max<int64_t>(3 * h->n_ctrl, h->ncP));
    double* bc = h->A.p + (size_t)h->nc * h->Pc * h->Pc;
    step_dots_kernel<<<(int)((cnt + 255) / 256), 256, 0, h->st>>>(h->dlt_c.p, h->dlt_s.p, h->diag_c.p,
                                                                  h->diag_s.p, bc, h->bs.p, h->ncP,
                                                                  3 * h->n_ctrl, xd, h->n, h->xs.p);
    h->launches++;
    // multi-GPU: the sums come from atomics in a non-deterministic order; every rank must take the SAME
    // accept / reject / lambda decisions (or the ranks would issue different numbers of collectives), so rank
    // 0's values are the values
    if (h->world > 1) { const int e = nccl_bcast0(h, h->xs.p, 5); if (e) return e; }
    MV_CUDA(h, cudaMemcpyAsync(h->h_pin, h->xs.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    for (int k = 0; k < 5; ++k) out[k] = h->h_pin[k];
    return MVUS_OK;
}

// Overrides of the LM driver's two problem-specific steps (points mode, ba_points.cuh): residual (+ Jacobian)
// evaluation leaving sum r^2 in h->partial[h->cost_slot], and the assembly of the normal equations.
struct mvus::LmHooks {
    std::function<int(const double*, bool)> eval;
    std::function<int()> accum;
    std::function<int(double*)> copy_r;
};

static int eval_cost(mvus_ba_ctx* h, const double* xd, bool want_j, double* cost) {
    int rc = h->hooks ? h->hooks->eval(xd, want_j) : evaluate(h, xd, want_j);
    if (rc) return rc;
    rc = allreduce_cost_slot(h);
    if (rc) return rc;
    return read_cost(h, cost);
}

struct PhaseTimer {
    mvus_ba_ctx* h; cudaEvent_t a, b; double* acc;
    PhaseTimer(mvus_ba_ctx* h_, int slot, double* acc_) : h(h_), a(h_->ev[slot]), b(h_->ev[slot + 1]), acc(acc_) {
        cudaEventRecord(a, h->st);
    }
    void stop() {
        cudaEventRecord(b, h->st);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        *acc += ms;
    }
};

extern "C" int mvus_ba_solve(mvus_ba_handle h, const double* x0, double* x_out, double* r_out,
                             mvus_ba_stats* stats) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x0 || !x_out) return fail(h, MVUS_ERR_ARG, "null argument");
    rc = ensure_solver(h);
    if (rc) return rc;
    mvus_ba_stats st;
    memset(&st, 0, sizeof(st));
    h->launches = 0;
    h->chunk_sorted = false;
    h->ms_syrk = h->ms_bcr = h->ms_reduce = h->ms_k2 = 0.0;
    const double ftol = h->desc.ftol, xtol = h->desc.xtol, gtol = h->desc.gtol;
    const int max_nfev = h->desc.max_nfev > 0 ? h->desc.max_nfev : 100 * (int)std::min<int64_t>(h->n, 1000);
    MV_CUDA(h, cudaEventRecord(h->ev[6], h->st));
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x0, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    double F = 0.0;
    {
        PhaseTimer t(h, 0, &st.ms_resjac);
        rc = eval_cost(h, h->x.p, true, &F);
        t.stop();
        st.n_resjac++;
    }
    if (rc) return rc;
    rc = check_motion_flag(h);
    if (rc) return rc;
    if (!std::isfinite(F)) return fail(h, MVUS_ERR_NONFINITE, "Residuals are not finite in the initial point.");
    st.cost0 = F;
    st.nfev = 1; st.njev = 1;
    // Levenberg-Marquardt with trust-region control of the step length (More, as in MINPACK's
    // lmpar, and the radius update of scipy/optimize/_lsq/trf.py:526-541): lambda is chosen so
    // that the scaled step norm |delta|_D is within [0.5, 1.5] of the radius Delta; a poor step shrinks
    // Delta to |delta|_D / 4, a very good one at the boundary doubles it.  Re-solving for a new
    // lambda costs linear solves but no residual evaluations (nfev is what max_iter caps).
    const double lam_min = 1e-10;
    const bool verbose = h->verbose;
    const double band_lo = h->band_lo, band_hi = h->band_hi;
    double lam = 1e-4, Delta = -1.0;
    double pexp = 2.0 / 3.0;          // running estimate of p in |delta|_D ~ lambda^-p
    double last_l = -1.0, last_n = 0.0;
    int status = 0;
    bool r_is_current = true, need_accum = true;
    double sc[5] = {0, 0, 0, 0, 0};
    auto solve_norm = [&](double l, int* ok, double* nrm) -> int {
        PhaseTimer t(h, 4, &st.ms_solve);
        int e = solve_damped(h, l, ok);
        if (!e && *ok) e = step_scalars(h, h->x.p, sc);
        t.stop();
        st.lm_iterations++;
        *nrm = *ok ? std::sqrt(sc[0]) : 1e300;
        if (verbose) fprintf(stderr, "[mvus_ba]   solve lam %.3e ok %d |d|_D %.4e Delta %.4e p %.3f\n", l, *ok, *nrm, Delta, pexp);
        return e;
    };
    while (st.nfev < max_nfev && status == 0) {
        if (need_accum) {
            PhaseTimer t(h, 2, &st.ms_accum);
            cudaEventRecord(h->evs[3], h->st);
            rc = h->hooks ? h->hooks->accum() : accumulate(h);
            cudaEventRecord(h->evs[4], h->st);
            if (!rc) rc = reduce_normal_equations(h, false);
            cudaEventRecord(h->evs[5], h->st);
            if (!rc) rc = compute_diag(h);
            if (!rc && h->desc.rs_bounds) {
                active_rho_kernel<<<(h->ncP + 127) / 128, 128, 0, h->st>>>(
                    h->x.p, h->A.p + (size_t)h->nc * h->Pc * h->Pc, h->nc, h->Pc, 1, h->frozen.p);
                h->launches++;
            }
            t.stop();
            if (rc) return rc;
            {
                float a = 0.f, b = 0.f;
                cudaEventElapsedTime(&a, h->evs[3], h->evs[4]);
                cudaEventElapsedTime(&b, h->evs[4], h->evs[5]);
                h->ms_k2 += a; h->ms_reduce += b;
            }
            need_accum = false;
        }
        int ok = 0;
        double nrm = 0.0;
        if (Delta > 0.0 && last_l > 0.0 && last_n > 0.0 && last_n < 1e299)   // aim the first solve at Delta
            lam = std::min(std::max(last_l * std::pow(last_n / Delta, 1.0 / pexp), lam_min), 1e30);
        rc = solve_norm(lam, &ok, &nrm);
        if (rc) return rc;
        if (ok && Delta < 0.0) Delta = nrm;
        double prev_l = ok ? lam : -1.0, prev_n = nrm;
        double good_l = ok ? lam : -1.0;      // last lambda whose solve succeeded (dlt_* / sc belong to the LAST solve)
        // bracket / secant search on lambda (log scale) for |delta|_D ~ Delta
        double lo_l = -1, lo_n = 0, hi_l = -1, hi_n = 0;
        for (int its = 0; its < 10; ++its) {
            if (!ok || nrm > band_hi * Delta) { lo_l = lam; lo_n = nrm; }
            else if (nrm < band_lo * Delta && lam > lam_min) { hi_l = lam; hi_n = nrm; }
            else break;
            if (lo_l > 0 && hi_l > 0 && lo_n < 1e299) {
                // bracketed: |delta|_D(lambda) is monotone but has plateaus and cliffs, so interpolate
                // in log-log and stay within the middle half of the bracket (bisection-like progress)
                const double a = std::log(lo_l), b = std::log(hi_l);
                double w = (std::log(lo_n) - std::log(Delta)) / (std::log(lo_n) - std::log(hi_n));
                w = std::min(std::max(w, 0.25), 0.75);
                lam = std::exp(a + w * (b - a));
            } else if (lo_l > 0 && hi_l > 0) {
                lam = std::sqrt(lo_l * hi_l);
            } else if (lo_l > 0) {
                const double f = ok ? std::pow(nrm / Delta, 1.0 / pexp) : 10.0;
                if (std::max(lam * f, lam * 2.0) > 1e30) break;      // (lam stays the one that was solved for)
                lam = std::max(lam * f, lam * 2.0);
            } else {
                lam = std::max(std::min(lam * std::pow(nrm / Delta, 1.0 / pexp), lam * 0.5), lam_min);
            }
            rc = solve_norm(lam, &ok, &nrm);
            if (rc) return rc;
            if (ok && Delta < 0.0) Delta = nrm;
            if (ok && prev_l > 0.0 && prev_n < 1e299 && lam != prev_l && nrm > 0.0 && prev_n > 0.0) {
                const double pe = -std::log(nrm / prev_n) / std::log(lam / prev_l);
                if (std::isfinite(pe)) pexp = std::min(std::max(pe, 0.15), 1.0);
            }
            if (ok) { prev_l = lam; prev_n = nrm; good_l = lam; }
        }
        if (!ok && good_l > 0.0) {            // the search ended on a failed factorisation: go back to the last good step
            lam = good_l;
            rc = solve_norm(lam, &ok, &nrm);
            if (rc) return rc;
        }
        if (!ok) { status = -1; break; }
        last_l = lam; last_n = nrm;
        st.optimality = __longlong_as_double_host(sc[4]);
        if (r_is_current && st.optimality < gtol) { status = 1; break; }
        const double pred = 0.5 * (lam * sc[0] + sc[1]);
        const double step_norm = std::sqrt(sc[2]), x_norm = std::sqrt(sc[3]);
        apply_step_kernel<<<(int)((std::max<int64_t>(h->n_other, h->n_ctrl) + 255) / 256), 256, 0, h->st>>>(
            h->x.p, h->dlt_c.p, h->dlt_s.p, h->nc, h->C, h->Pc, h->n_other, h->sv, h->n_ctrl,
            h->desc.rs_bounds, h->x_trial.p);
        h->launches++;
        double Fn = 0.0;
        {
            PhaseTimer t(h, 0, &st.ms_trial);
            rc = eval_cost(h, h->x_trial.p, false, &Fn);
            t.stop();
        }
        if (rc) return rc;
        st.nfev++;
        r_is_current = false;
        const double actual = F - Fn;
        const double ratio = (pred > 0.0 && std::isfinite(Fn)) ? actual / pred : -1.0;
        if (verbose) fprintf(stderr, "[mvus_ba] nfev %d F %.8e Fn %.8e ratio %.3f lam %.3e |d|_D %.3e Delta %.3e |g|inf %.3e\n",
                             st.nfev, F, Fn, ratio, lam, nrm, Delta, st.optimality);
        if (ratio < 0.25) Delta = 0.25 * nrm;
        else if (ratio > 0.75 && nrm > 0.7 * Delta) Delta *= 2.0;
        const bool x_small = step_norm < xtol * (xtol + x_norm);
        if (std::isfinite(Fn) && actual > 0.0) {
            std::swap(h->x.p, h->x_trial.p);
            const bool f_small = actual < ftol * Fn && ratio > 0.25;
            F = Fn;
            if (f_small && x_small) status = 4;
            else if (f_small) status = 2;
            else if (x_small) status = 3;
            if (status || st.nfev >= max_nfev) {
                PhaseTimer t(h, 0, &st.ms_trial);      // leave r consistent with the accepted x
                rc = eval_cost(h, h->x.p, false, &F);
                t.stop();
                r_is_current = true;
                if (rc) return rc;
                break;
            }
            {
                PhaseTimer t(h, 0, &st.ms_resjac);
                rc = eval_cost(h, h->x.p, true, &F);
                t.stop();
                st.n_resjac++;
            }
            if (rc) return rc;
            st.njev++;
            r_is_current = true;
            need_accum = true;
        } else if (x_small) {
            status = 3;
        }
    }
    if (!r_is_current) {
        PhaseTimer t(h, 0, &st.ms_trial);
        rc = eval_cost(h, h->x.p, false, &F);
        t.stop();
        if (rc) return rc;
    }
    st.cost = F;
    st.lambda = lam;
    st.status = status;
    MV_CUDA(h, cudaMemcpyAsync(x_out, h->x.p, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (r_out && h->hooks) { rc = h->hooks->copy_r(r_out); if (rc) return rc; }
    else if (r_out) MV_CUDA(h, cudaMemcpyAsync(r_out, h->r.p, h->m * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaEventRecord(h->ev[7], h->st));
    MV_CUDA(h, cudaEventSynchronize(h->ev[7]));
    float ms = 0.f;
    MV_CUDA(h, cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]));
    st.ms_total = ms;
    st.launches = h->launches;
    st.ms_syrk = h->ms_syrk; st.ms_bcr = h->ms_bcr; st.ms_reduce = h->ms_reduce; st.ms_k2 = h->ms_k2;
    if (stats) *stats = st;
    return MVUS_OK;
}

extern "C" int mvus_ba_normal_equations(mvus_ba_handle h, const double* x, double* A, double* g,
                                        double* Hss, int32_t* band_ctrl, double* Hcs, double* cost) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x) return fail(h, MVUS_ERR_ARG, "null argument");
    rc = ensure_solver(h);
    if (rc) return rc;
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    double F = 0.0;
    rc = eval_cost(h, h->x.p, true, &F);
    if (rc) return rc;
    rc = check_motion_flag(h);
    if (rc) return rc;
    h->chunk_sorted = false;
    rc = accumulate(h);
    if (!rc) rc = reduce_normal_equations(h, true);      // diagnostics: every rank gets everything
    if (!rc) rc = compute_diag(h, true);
    if (rc) return rc;
    if (cost) *cost = F;
    const int q = h->q, bw = h->bw, ldw = h->ldw, Pc = h->Pc;
    const int64_t nb = h->nb;
    if (A) MV_CUDA(h, cudaMemcpyAsync(A, h->A.p, (size_t)h->nc * Pc * Pc * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (g) {
        double* bc = h->A.p + (size_t)h->nc * Pc * Pc;
        gradient_kernel<<<(int)((std::max<int64_t>(h->n_other, h->n_ctrl) + 255) / 256), 256, 0, h->st>>>(
            bc, h->bs.p, h->nc, h->C, Pc, h->n_other, h->sv, h->n_ctrl, h->gvec.p);
        MV_CUDA(h, cudaMemcpyAsync(g, h->gvec.p, h->n * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    }
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    if (band_ctrl) *band_ctrl = 2 * bw;
    if (Hss || Hcs) {
        std::vector<double> D((size_t)nb * q * q), E((size_t)nb * q * q), W((size_t)nb * q * ldw);
        MV_CUDA(h, cudaMemcpy(D.data(), h->D.p, D.size() * sizeof(double), cudaMemcpyDeviceToHost));
        MV_CUDA(h, cudaMemcpy(E.data(), h->E.p, E.size() * sizeof(double), cudaMemcpyDeviceToHost));
        MV_CUDA(h, cudaMemcpy(W.data(), h->Wp(), W.size() * sizeof(double), cudaMemcpyDeviceToHost));
        const int64_t nctrl = h->n_ctrl;
        if (Hss) {
            const int band = 2 * bw;
            memset(Hss, 0, (size_t)nctrl * band * 9 * sizeof(double));
            for (int64_t i = 0; i < nctrl; ++i)
                for (int dj = 0; dj < band && i + dj < nctrl; ++dj) {
                    const int64_t j = i + dj, ki = i / bw, kj = j / bw;
                    if (kj > ki + 1) continue;
                    for (int a = 0; a < 3; ++a)
                        for (int b = 0; b < 3; ++b) {
                            const int la = (int)(i - ki * bw) * 3 + a, lb = (int)(j - kj * bw) * 3 + b;
                            const double v = (ki == kj) ? D[((size_t)ki * q + std::min(la, lb)) * q + std::max(la, lb)]   // upper triangle stored
                                                    : E[((size_t)ki * q + la) * q + lb];
                            Hss[((size_t)i * band + dj) * 9 + a * 3 + b] = v;
                        }
                }
        }
        if (Hcs) {
            const int ncP = h->ncP;
            for (int c = 0; c < ncP; ++c)
                for (int64_t k = 0; k < 3 * nctrl; ++k) Hcs[(size_t)c * 3 * nctrl + k] = W[(size_t)k * ldw + c];
        }
    }
    return MVUS_OK;
}

extern "C" int mvus_ba_time_accumulate(mvus_ba_handle h, int32_t reps, double* ms_mean) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!ms_mean || reps < 1) return fail(h, MVUS_ERR_ARG, "bad argument");
    if (!h->J.p) return fail(h, MVUS_ERR_ARG, "evaluate a Jacobian first (mvus_ba_time_resjac)");
    rc = ensure_solver(h);
    if (rc) return rc;
    h->chunk_sorted = false;
    rc = accumulate(h);                  // (sorts the chunks; the timed repetitions reuse the order, as the LM loop does)
    if (rc) return rc;
    MV_CUDA(h, cudaEventRecord(h->ev[0], h->st));
    for (int k = 0; k < reps; ++k) {
        rc = accumulate(h);
        if (rc) return rc;
    }
    MV_CUDA(h, cudaEventRecord(h->ev[1], h->st));
    MV_CUDA(h, cudaEventSynchronize(h->ev[1]));
    float ms = 0.f;
    MV_CUDA(h, cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]));
    *ms_mean = (double)ms / reps;
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int mvus_ba_global_traj(mvus_ba_handle h, const double* x, const int32_t* cam_ids,
                                   int64_t* n_out, double* out, double* gd_out) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x || !cam_ids || !n_out || !out) return fail(h, MVUS_ERR_ARG, "null argument");
    *n_out = 0;
    const int64_t N = h->N;
    if (N == 0) return MVUS_OK;
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    cam_prep_kernel<<<(h->nc + 63) / 64, 64, 0, h->st>>>(h->x.p, h->nc, h->C, h->desc.opt_calib, h->calib.p,
                                                         h->height.p, h->camprep.p);
    DevBuf<double> ts, ts_s;
    DevBuf<int> idx, idx_s, flag, pos, cams;
    DevBuf<unsigned char> tmp;
    cudaError_t e = ts.alloc(N);
    if (e == cudaSuccess) e = ts_s.alloc(N);
    if (e == cudaSuccess) e = idx.alloc(N);
    if (e == cudaSuccess) e = idx_s.alloc(N);
    if (e == cudaSuccess) e = flag.alloc(N);
    if (e == cudaSuccess) e = pos.alloc(N);
    if (e == cudaSuccess) e = upload(cams, cam_ids, (size_t)h->nc, h->st);
    size_t tb1 = 0, tb2 = 0;
    if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(nullptr, tb1, ts.p, ts_s.p, idx.p, idx_s.p, (int)N, 0, 64, h->st);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tb2, flag.p, pos.p, (int)N, h->st);
    size_t tb = tb1 > tb2 ? tb1 : tb2;
    if (e == cudaSuccess) e = tmp.alloc(tb);
    auto cleanup = [&]() { cudaStreamSynchronize(h->st); ts.release(false); ts_s.release(false); idx.release(false); idx_s.release(false); flag.release(false); pos.release(false); cams.release(false); tmp.release(false); };
    if (e != cudaSuccess) { cleanup(); return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e)); }
    if (gd_out) {          // global_detections: 3 x N (camera id, frame id, time stamp), concatenation order
        e = h->scratch.alloc((size_t)3 * N);
        if (e == cudaSuccess) {
            gd_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->camprep.p, h->tile_cam.p, h->tile_start.p, h->tile_cnt.p,
                                                         h->frame.p, h->yr.p, cams.p, N, h->scratch.p);
            e = cudaMemcpyAsync(gd_out, h->scratch.p, (size_t)3 * N * sizeof(double), cudaMemcpyDeviceToHost, h->st);
        }
        if (e != cudaSuccess) { cleanup(); return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e)); }
    }
    gt_times_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->camprep.p, h->tile_cam.p, h->tile_start.p, h->tile_cnt.p,
                                                       h->frame.p, h->yr.p, ts.p, idx.p);
    e = cub::DeviceRadixSort::SortPairs(tmp.p, tb, ts.p, ts_s.p, idx.p, idx_s.p, (int)N, 0, 64, h->st);
    const int gb = (int)((N + 255) / 256);
    if (e == cudaSuccess) {
        gt_flag_kernel<<<gb, 256, 0, h->st>>>(h->sv, ts_s.p, N, flag.p);
        e = cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, pos.p, (int)N, h->st);
    }
    int last_pos = 0, last_flag = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&last_pos, pos.p + (N - 1), sizeof(int), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&last_flag, flag.p + (N - 1), sizeof(int), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) { cleanup(); return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e)); }
    const int64_t n = (int64_t)last_pos + last_flag;
    *n_out = n;
    if (n > 0) {
        e = h->gt_out.alloc((size_t)7 * N);                // sized for the worst case: allocated once
        if (e == cudaSuccess) {
            gt_gather_kernel<<<gb, 256, 0, h->st>>>(h->sv, h->x.p, ts_s.p, idx_s.p, flag.p, pos.p, N, n, h->row_off.p,
                                                   h->nc, cams.p, h->frame.p, h->gt_out.p);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(out, h->gt_out.p, (size_t)7 * n * sizeof(double), cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    }
    cleanup();
    if (e != cudaSuccess) return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e));
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
// Pinned host memory for the large outputs (detections_global, residuals, global_traj): pageable
// device->host copies measured ~4 GB/s on the GPU box, pinned ones run at PCIe speed.
extern "C" void* mvus_ba_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
extern "C" void mvus_ba_host_free(void* p) { if (p) cudaFreeHost(p); }

// visible (Scene.compute_visibility, common.py:427-438): 1-based interval id per detection at x,
// 0 = covered by no interval -- util.sampling(..., belong=True) (util.py:103-106).
extern "C" int mvus_ba_visibility(mvus_ba_handle h, const double* x, int64_t* visible) {
    int rc = check_ready(h);
    if (rc) return rc;
    if (!x || !visible) return fail(h, MVUS_ERR_ARG, "null argument");
    if (h->N == 0) return MVUS_OK;
    MV_CUDA(h, cudaMemcpyAsync(h->x.p, x, h->n * sizeof(double), cudaMemcpyHostToDevice, h->st));
    cam_prep_kernel<<<(h->nc + 63) / 64, 64, 0, h->st>>>(h->x.p, h->nc, h->C, h->desc.opt_calib, h->calib.p,
                                                         h->height.p, h->camprep.p);
    MV_CUDA(h, h->scratch.alloc((size_t)3 * h->N));      // sized for detections_global: allocated once
    visibility_kernel<<<h->n_tiles, TILE_DET, 0, h->st>>>(h->sv, h->camprep.p, h->tile_cam.p, h->tile_start.p,
                                                         h->tile_cnt.p, h->frame.p, h->yr.p,
                                                         reinterpret_cast<long long*>(h->scratch.p));
    MV_CUDA(h, cudaGetLastError());
    MV_CUDA(h, cudaMemcpyAsync(visible, h->scratch.p, (size_t)h->N * sizeof(int64_t), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int mvus_ba_spline_to_traj(mvus_ba_handle h, const double* x, const double* t, int64_t n,
                                      int64_t* n_out, double* out) {
    if (!h) return MVUS_ERR_ARG;
    if (!h->have_spl) return fail(h, MVUS_ERR_ARG, "set_splines first");
    if (!x || !n_out || !out || (n > 0 && !t)) return fail(h, MVUS_ERR_ARG, "null argument");
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    *n_out = 0;
    if (n == 0) return MVUS_OK;
    if (n >= (int64_t)1 << 31) return fail(h, MVUS_ERR_UNSUPPORTED, "more than 2^31 sample times");
    const int64_t nx = h->n_other + 3 * h->n_ctrl;
    DevBuf<double> xd, td, od;
    DevBuf<int> flag, pos;
    DevBuf<unsigned char> tmp;
    auto cleanup = [&]() { cudaStreamSynchronize(h->st); xd.release(false); td.release(false); od.release(false); flag.release(false); pos.release(false); tmp.release(false); };
    cudaError_t e = upload(xd, x, (size_t)nx, h->st);
    if (e == cudaSuccess) e = upload(td, t, (size_t)n, h->st);
    if (e == cudaSuccess) e = flag.alloc(n);
    if (e == cudaSuccess) e = pos.alloc(n);
    size_t tb = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, pos.p, (int)n, h->st);
    if (e == cudaSuccess) e = tmp.alloc(tb);
    const int gb = (int)((n + 255) / 256);
    if (e == cudaSuccess) {
        s2t_flag_kernel<<<gb, 256, 0, h->st>>>(h->sv, td.p, n, flag.p);
        e = cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, pos.p, (int)n, h->st);
    }
    int lp = 0, lf = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&lp, pos.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&lf, flag.p + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) { cleanup(); return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e)); }
    const int64_t m = (int64_t)lp + lf;
    *n_out = m;
    if (m > 0) {
        e = od.alloc((size_t)4 * m);
        if (e == cudaSuccess) {
            s2t_gather_kernel<<<gb, 256, 0, h->st>>>(h->sv, xd.p, td.p, flag.p, pos.p, n, m, od.p);
            e = cudaMemcpyAsync(out, od.p, (size_t)4 * m * sizeof(double), cudaMemcpyDeviceToHost, h->st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    }
    cleanup();
    if (e != cudaSuccess) return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e));
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
extern "C" int mvus_ba_align(mvus_ba_handle h, const double* x, int64_t n, const double* tau, const double* pts,
                             int32_t nshift, const double* shift, int32_t spline_is_src, int32_t want,
                             double* mean_err, int64_t* count, double* M, double* err) {
    if (!h) return MVUS_ERR_ARG;
    if (!h->have_spl) return fail(h, MVUS_ERR_ARG, "set_splines first");
    if (!x || !tau || !pts || !shift || !mean_err || !count || !M || n < 1 || nshift < 1)
        return fail(h, MVUS_ERR_ARG, "null or empty argument");
    MV_CUDA(h, cudaSetDevice(h->desc.device));
    const int64_t nx = h->n_other + 3 * h->n_ctrl;
    DevBuf<double> xd, td, pd, sd, me, Md, ed;
    DevBuf<int64_t> cd;
    auto cleanup = [&]() { cudaStreamSynchronize(h->st); xd.release(false); td.release(false); pd.release(false); sd.release(false); me.release(false); Md.release(false); ed.release(false); cd.release(false); };
    cudaError_t e = upload(xd, x, (size_t)nx, h->st);
    if (e == cudaSuccess) e = upload(td, tau, (size_t)n, h->st);
    if (e == cudaSuccess) e = upload(pd, pts, (size_t)3 * n, h->st);
    if (e == cudaSuccess) e = upload(sd, shift, (size_t)nshift, h->st);
    if (e == cudaSuccess) e = me.alloc(nshift);
    if (e == cudaSuccess) e = Md.alloc((size_t)16 * nshift);
    if (e == cudaSuccess) e = cd.alloc(nshift);
    if (e == cudaSuccess && err) e = ed.alloc(n);
    if (e == cudaSuccess) {
        align_fit_kernel<<<nshift, AL_T, 0, h->st>>>(h->sv, xd.p, td.p, pd.p, n, sd.p, spline_is_src, err ? want : -1,
                                                     me.p, cd.p, Md.p, err ? ed.p : nullptr);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(mean_err, me.p, nshift * sizeof(double), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(count, cd.p, nshift * sizeof(int64_t), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(M, Md.p, (size_t)16 * nshift * sizeof(double), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess && err) e = cudaMemcpyAsync(err, ed.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    cleanup();
    if (e != cudaSuccess) return fail(h, MVUS_ERR_CUDA, cudaGetErrorString(e));
    return MVUS_OK;
}

// ------------------------------------------------------------------------------------------
// Scene.BA(motion_prior=True) (ba_points.cuh)
extern "C" int mvus_ba_points_set(mvus_ba_handle hp, int64_t G, const int32_t* cam, const double* frame, const double* yH) {
    if (!hp) return MVUS_ERR_ARG;
    if (!hp->have_spl) return fail(hp, MVUS_ERR_ARG, "set_splines first (one pseudo-spline with G coefficients)");
    if (G < 1 || !cam || !frame || !yH) return fail(hp, MVUS_ERR_ARG, "null or empty argument");
    if (G != hp->n_ctrl) return fail(hp, MVUS_ERR_ARG, "G must equal the number of coefficients of the pseudo-spline");
    for (int64_t j = 0; j < G; ++j)
        if (cam[j] < 0 || cam[j] >= hp->nc) return fail(hp, MVUS_ERR_ARG, "camera slot out of range");
    MV_CUDA(hp, cudaSetDevice(hp->desc.device));
    MV_CUDA(hp, upload(hp->pt_cam, cam, (size_t)G, hp->st));
    MV_CUDA(hp, upload(hp->pt_frame, frame, (size_t)G, hp->st));
    MV_CUDA(hp, upload(hp->pt_yH, yH, (size_t)G, hp->st));
    MV_CUDA(hp, hp->pt_r.alloc(G));
    MV_CUDA(hp, hp->pt_J.alloc((size_t)9 * G));
    MV_CUDA(hp, hp->pt_idx.alloc((size_t)3 * G));
    MV_CUDA(hp, cudaStreamSynchronize(hp->st));
    hp->ptG = G;
    return MVUS_OK;
}

namespace {
struct PointsRun {
    mvus_ba_ctx* hs; mvus_ba_ctx* hp;
    PointsView pv;
    int nblk;
    int check() const {
        if (!hs || !hp) return MVUS_ERR_ARG;
        if (!(hs->have_det && hs->have_spl && hp->have_det && hp->have_spl)) return fail(hp, MVUS_ERR_ARG, "both handles need detections and splines set");
        if (hp->ptG < 1) return fail(hp, MVUS_ERR_ARG, "mvus_ba_points_set first");
        if (hs->nc != hp->nc || hs->C != hp->C || hs->desc.device != hp->desc.device || hp->N != 0 || hp->M != 0 || hs->M != 0)
            return fail(hp, MVUS_ERR_ARG, "handles do not describe the same cameras (or the points handle has detections / motion samples)");
        if (hs->world > 1 || hp->world > 1) return fail(hp, MVUS_ERR_UNSUPPORTED, "the discrete-trajectory mode runs on one GPU");
        return MVUS_OK;
    }
    void init(mvus_ba_ctx* s, mvus_ba_ctx* p, int motion_type, double weight) {
        hs = s; hp = p;
        pv = PointsView{p->ptG, p->pt_cam.p, p->pt_frame.p, p->pt_yH.p, p->nc, (int)p->n_other, motion_type, weight,
                        s->desc.opt_sync, s->desc.opt_rs};
        nblk = (int)((p->ptG + 127) / 128);
    }
    // residuals (and Jacobians) at xd (hp layout); sum r^2 -> hp->partial[hp->cost_slot]
    int eval(const double* xd, bool want_j) {
        cudaStreamSynchronize(hp->st);                 // xd was produced on hp's stream; hs works on its own
        cudaError_t e = cudaMemcpyAsync(hs->x.p, xd, hs->n_other * sizeof(double), cudaMemcpyDeviceToDevice, hs->st);
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        int rc = evaluate(hs, hs->x.p, want_j);
        if (rc) return fail(hp, rc, hs->err);
        e = hp->partial.alloc((size_t)nblk + 8);
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        cudaStreamSynchronize(hs->st);
        if (want_j)
            points_motion_kernel<true><<<nblk, 128, 0, hp->st>>>(pv, hs->sv, xd, hp->pt_r.p, hp->pt_idx.p, hp->pt_J.p,
                                                                 hp->partial.p, hp->flag.p);
        else
            points_motion_kernel<false><<<nblk, 128, 0, hp->st>>>(pv, hs->sv, xd, hp->pt_r.p, nullptr, nullptr,
                                                                  hp->partial.p, hp->flag.p);
        points_cost_kernel<<<1, 256, 0, hp->st>>>(hs->partial.p + hs->cost_slot, hp->partial.p, nblk, hp->partial.p + nblk);
        hp->cost_slot = nblk;
        hp->launches += hs->launches + 2;
        hs->launches = 0;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        return MVUS_OK;
    }
    int accum() {
        const size_t qq = (size_t)hp->q * hp->q;
        const int Pc = hp->Pc, ncP = hp->ncP;
        double* bc = hp->A.p + (size_t)hp->nc * Pc * Pc;
        cudaError_t e = hp->Ax.alloc((size_t)ncP * ncP);
        if (e == cudaSuccess) e = hs->scratch.alloc((size_t)2 * hs->P * std::max<int64_t>(hs->N, 1));
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        cudaMemsetAsync(hp->A.p, 0, hp->A.bytes(), hp->st);
        cudaMemsetAsync(hp->Ax.p, 0, hp->Ax.bytes(), hp->st);
        cudaMemsetAsync(hp->D.p, 0, hp->nb * qq * sizeof(double), hp->st);
        cudaMemsetAsync(hp->E.p, 0, hp->nb * qq * sizeof(double), hp->st);
        cudaMemsetAsync(hp->Wp(), 0, (size_t)hp->nb * hp->q * hp->ldw * sizeof(double), hp->st);
        if (hs->N > 0) {
            deblock_kernel<<<hs->n_tiles, TILE_DET, 0, hp->st>>>(hs->J.p, hs->P, hs->tile_start.p, hs->tile_cnt.p, hs->N,
                                                              hs->scratch.p, hs->span.p);
            cam_blocks_kernel<<<dim3(hp->nc, Pc * (Pc + 1)), 256, 0, hp->st>>>(hs->scratch.p, hs->r.p, hs->P, Pc, hs->N,
                                                                             hs->row_off.p, hp->A.p, bc);
        }
        points_accum_kernel<<<nblk, 128, 0, hp->st>>>(pv, hp->x.p, hp->pt_r.p, hp->pt_idx.p, hp->pt_J.p, hp->bw, Pc,
                                                     hp->ldw, hp->A.p, bc, hp->Ax.p, hp->D.p, hp->E.p, hp->Wp());
        hp->launches += 3;
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        return MVUS_OK;
    }
    int copy_r(double* r_out) {
        cudaError_t e = cudaMemcpyAsync(r_out, hs->r.p, 2 * hs->N * sizeof(double), cudaMemcpyDeviceToHost, hp->st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(r_out + 2 * hs->N, hp->pt_r.p, hp->ptG * sizeof(double), cudaMemcpyDeviceToHost, hp->st);
        if (e != cudaSuccess) return fail(hp, MVUS_ERR_CUDA, cudaGetErrorString(e));
        return MVUS_OK;
    }
};

int points_prepare(PointsRun& run, mvus_ba_ctx* hs, mvus_ba_ctx* hp, int32_t motion_type, double weight, const double* xs0) {
    run.hs = hs; run.hp = hp;
    int rc = run.check();
    if (rc) return rc;
    if (motion_type != MVUS_MOTION_F && motion_type != MVUS_MOTION_KE) return fail(hp, MVUS_ERR_ARG, "motion_type must be F or KE");
    if (!xs0) return fail(hp, MVUS_ERR_ARG, "null argument");
    MV_CUDA(hp, cudaSetDevice(hp->desc.device));
    MV_CUDA(hp, cudaMemcpyAsync(hs->x.p, xs0, hs->n * sizeof(double), cudaMemcpyHostToDevice, hs->st));   // the constant splines
    MV_CUDA(hp, cudaStreamSynchronize(hs->st));
    run.init(hs, hp, motion_type, weight);
    return MVUS_OK;
}
}  // namespace

// The mode's residual vector (reference order: reprojection rows of hs, then G motion rows in global_traj order),
// its gradient J^T r in hp's x layout and the cost, at x.  Diagnostics for the parity tests.
extern "C" int mvus_ba_points_eval(mvus_ba_handle hs, mvus_ba_handle hp, int32_t motion_type, double motion_weight,
                                   const double* xs0, const double* x, double* r_out, double* g_out, double* cost) {
    PointsRun run;
    int rc = points_prepare(run, hs, hp, motion_type, motion_weight, xs0);
    if (rc) return rc;
    if (!x) return fail(hp, MVUS_ERR_ARG, "null argument");
    rc = ensure_solver(hp);
    if (rc) return rc;
    MV_CUDA(hp, cudaMemsetAsync(hp->flag.p, 0, sizeof(int), hp->st));
    MV_CUDA(hp, cudaMemcpyAsync(hp->x.p, x, hp->n * sizeof(double), cudaMemcpyHostToDevice, hp->st));
    rc = run.eval(hp->x.p, true);
    if (rc) return rc;
    double F = 0.0;
    rc = read_cost(hp, &F);
    if (rc) return rc;
    if (cost) *cost = F;
    int f = 0;
    MV_CUDA(hp, cudaMemcpyAsync(&f, hp->flag.p, sizeof(int), cudaMemcpyDeviceToHost, hp->st));
    MV_CUDA(hp, cudaStreamSynchronize(hp->st));
    if (f) return fail(hp, MVUS_ERR_UNSUPPORTED, "a motion row spans more than 4 consecutive trajectory points");
    if (r_out) { rc = run.copy_r(r_out); if (rc) return rc; }
    if (g_out) {
        rc = run.accum();
        if (!rc) rc = compute_diag(hp, true);
        if (rc) return rc;
        double* bc = hp->A.p + (size_t)hp->nc * hp->Pc * hp->Pc;
        gradient_kernel<<<(int)((std::max<int64_t>(hp->n_other, hp->n_ctrl) + 255) / 256), 256, 0, hp->st>>>(
            bc, hp->bs.p, hp->nc, hp->C, hp->Pc, hp->n_other, hp->sv, hp->n_ctrl, hp->gvec.p);
        MV_CUDA(hp, cudaMemcpyAsync(g_out, hp->gvec.p, hp->n * sizeof(double), cudaMemcpyDeviceToHost, hp->st));
    }
    MV_CUDA(hp, cudaStreamSynchronize(hp->st));
    return MVUS_OK;
}

extern "C" int mvus_ba_solve_points(mvus_ba_handle hs, mvus_ba_handle hp, int32_t motion_type, double motion_weight,
                                    const double* xs0, const double* x0, double* x_out, double* r_out,
                                    mvus_ba_stats* stats) {
    PointsRun run;
    int rc = points_prepare(run, hs, hp, motion_type, motion_weight, xs0);
    if (rc) return rc;
    MV_CUDA(hp, cudaMemsetAsync(hp->flag.p, 0, sizeof(int), hp->st));
    LmHooks hk;
    hk.eval = [&](const double* xd, bool wj) { return run.eval(xd, wj); };
    hk.accum = [&]() { return run.accum(); };
    hk.copy_r = [&](double* r) { return run.copy_r(r); };
    hp->hooks = &hk;
    rc = mvus_ba_solve(hp, x0, x_out, r_out, stats);
    hp->hooks = nullptr;
    if (rc == MVUS_ERR_UNSUPPORTED) hp->err = "a motion row spans more than 4 consecutive trajectory points";
    if (!rc) {
        int f = 0;
        MV_CUDA(hp, cudaMemcpyAsync(&f, hp->flag.p, sizeof(int), cudaMemcpyDeviceToHost, hp->st));
        MV_CUDA(hp, cudaStreamSynchronize(hp->st));
        if (f) return fail(hp, MVUS_ERR_UNSUPPORTED, "a motion row spans more than 4 consecutive trajectory points");
    }
    return rc;
}
