// Device version of the bookkeeping Scene.all_detect_to_traj leaves behind (common.py:887-944):
// `global_traj` = every detection of the optimised cameras whose global time stamp lies inside
// a spline interval (closed, as spline_to_traj selects them, common.py:292), sorted by time
// stamp, with the spline position at that time.  The reference rebuilds it with NumPy sorts at
// EVERY error_BA evaluation (common.py:462-464) although only the final state is kept (it is a
// pickled output, README.md:216-221); here it is produced once after the solve: one radix sort
// (CUB), one flag/scan and one gather kernel.  Not part of the residual path.
#pragma once
#include <cub/cub.cuh>
#include "ba_ctx.cuh"

namespace mvus {

__global__ void gt_times_kernel(const double* __restrict__ camprep, const int* __restrict__ tile_cam,
                                const int64_t* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                                const double* __restrict__ frame, const double* __restrict__ yr,
                                double* __restrict__ ts, int* __restrict__ idx) {
    const int tl = blockIdx.x;
    if ((int)threadIdx.x >= tile_cnt[tl]) return;
    const CamPrep& c = *reinterpret_cast<const CamPrep*>(camprep + (size_t)tile_cam[tl] * CAMPREP_DOUBLES);
    const int64_t d = tile_start[tl] + threadIdx.x;
    ts[d] = c.alpha * (frame[d] + c.rho * (yr[d] * c.invH)) + c.beta;
    idx[d] = (int)d;
}

// global_detections (common.py:927): rows camera id, frame id, global time stamp, in the
// concatenation order of the optimised cameras.  out[3][N]
__global__ void gd_kernel(const double* __restrict__ camprep, const int* __restrict__ tile_cam,
                          const int64_t* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                          const double* __restrict__ frame, const double* __restrict__ yr,
                          const int* __restrict__ cam_ids, int64_t N, double* __restrict__ out) {
    const int tl = blockIdx.x;
    if ((int)threadIdx.x >= tile_cnt[tl]) return;
    const int cam = tile_cam[tl];
    const CamPrep& c = *reinterpret_cast<const CamPrep*>(camprep + (size_t)cam * CAMPREP_DOUBLES);
    const int64_t d = tile_start[tl] + threadIdx.x;
    out[d] = (double)cam_ids[cam];
    out[N + d] = frame[d];
    out[2 * N + d] = c.alpha * (frame[d] + c.rho * (yr[d] * c.invH)) + c.beta;
}

__device__ __forceinline__ int closed_interval(const SplineView& sp, double t) {
    for (int s = 0; s < sp.S; ++s)
        if (t >= sp.int_a[s] && t <= sp.int_b[s]) return s;
    return -1;
}

__global__ void gt_flag_kernel(SplineView sp, const double* __restrict__ ts_sorted, int64_t N,
                               int* __restrict__ flag) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < N) flag[k] = closed_interval(sp, ts_sorted[k]) >= 0 ? 1 : 0;
}

__global__ void gt_gather_kernel(SplineView sp, const double* __restrict__ x, const double* __restrict__ ts_sorted,
                                 const int* __restrict__ idx_sorted, const int* __restrict__ flag,
                                 const int* __restrict__ pos, int64_t N, int64_t n_out,
                                 const int64_t* __restrict__ row_off, int nc, const int* __restrict__ cam_ids,
                                 const double* __restrict__ frame, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= N || !flag[k]) return;
    const int64_t o = pos[k];
    const int d = idx_sorted[k];
    const double t = ts_sorted[k];
    int lo = 0, hi = nc - 1;                       // camera of detection d: row_off[c]/2 <= d
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if ((row_off[mid] >> 1) <= d) lo = mid; else hi = mid - 1; }
    const int s = closed_interval(sp, t);
    double X[3], dX[3], B[4];
    spline_eval<false>(sp, x, s, t, X, dX, B);
    out[0 * n_out + o] = (double)o;
    out[1 * n_out + o] = (double)cam_ids[lo];
    out[2 * n_out + o] = frame[d];
    out[3 * n_out + o] = t;
    out[4 * n_out + o] = X[0];
    out[5 * n_out + o] = X[1];
    out[6 * n_out + o] = X[2];
}

// Scene.spline_to_traj (common.py:273-301): evaluate the splines at ascending times t; a time is
// kept iff it lies inside a spline interval (closed ends, common.py:292).
__global__ void s2t_flag_kernel(SplineView sp, const double* __restrict__ t, int64_t n, int* __restrict__ flag) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) flag[k] = closed_interval(sp, t[k]) >= 0 ? 1 : 0;
}
__global__ void s2t_gather_kernel(SplineView sp, const double* __restrict__ x, const double* __restrict__ t,
                                  const int* __restrict__ flag, const int* __restrict__ pos, int64_t n,
                                  int64_t n_out, double* __restrict__ out) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || !flag[k]) return;
    const int64_t o = pos[k];
    const double tt = t[k];
    double X[3], dX[3], B[4];
    spline_eval<false>(sp, x, closed_interval(sp, tt), tt, X, dX, B);
    out[o] = tt;
    out[n_out + o] = X[0];
    out[2 * n_out + o] = X[1];
    out[3 * n_out + o] = X[2];
}

}  // namespace mvus
