// K0 / K1 / K1m: camera preparation, fused residual + analytic Jacobian per detection,
// motion-prior rows.  One thread per detection (per motion sample); detections are tiled so
// that one CTA only sees one camera and keeps its prepared parameters in shared memory.
//
// HBM traffic per detection (algorithmic): read frame,x,y (24 B; +16 B of
// pre-undistorted observation when the calibration is fixed), write r_u,r_v (16 B), the span
// index (4 B) and the compact block row 2*P*8 B  ->  380 B/det (P=21), 524 B/det (P=30).
// All loads/stores are unit-stride across the warp (SoA detections, column-plane Jacobian);
// knot/coefficient/camera tables are small and stay in L2/L1.
#pragma once
#include "ba_ctx.cuh"

namespace mvus {

__global__ void cam_prep_kernel(const double* __restrict__ x, int nc, int C, int calib,
                                const double* __restrict__ calib9, const double* __restrict__ height,
                                double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc) return;
    CamPrep c;
    cam_prep_one(x, i, nc, C, calib != 0, calib9, height[i], c);
    double* o = out + (size_t)i * CAMPREP_DOUBLES;
    const double* src = reinterpret_cast<const double*>(&c);
    for (int k = 0; k < CAMPREP_DOUBLES; ++k) o[k] = src[k];
}

// Observation with fixed calibration: K * undistortPoints(raw) once per BA instead of once
// per evaluation (the reference recomputes it every call, common.py:126).
__global__ void observe_kernel(const int* __restrict__ tile_cam, const int64_t* __restrict__ tile_start,
                               const int* __restrict__ tile_cnt, const double* __restrict__ calib9,
                               int undist, const double* __restrict__ xr, const double* __restrict__ yr,
                               double* __restrict__ ou, double* __restrict__ ov) {
    const int tl = blockIdx.x;
    if ((int)threadIdx.x >= tile_cnt[tl]) return;
    const int64_t d = tile_start[tl] + threadIdx.x;
    const double* c = calib9 + tile_cam[tl] * 9;
    if (undist) {
        double xn, yn;
        undistort5(xr[d], yr[d], c, c + 4, xn, yn);
        ou[d] = c[0] * xn + c[2];
        ov[d] = c[1] * yn + c[3];
    } else {
        ou[d] = xr[d];
        ov[d] = yr[d];
    }
}

// Jacobian block of the warp (ba_ctx.cuh, JBlk): column p of the u row -> plane p, of the v row -> plane P + p
struct BlockSink {
    double* blk;
    int lt, P;
    __device__ __forceinline__ void put(int p, double a, double b) {
        __stcs(blk + jblk_off(p, lt), a);
        __stcs(blk + jblk_off(P + p, lt), b);
    }
};
struct NullSink {
    __device__ __forceinline__ void put(int, double, double) {}
};

__device__ __forceinline__ double block_sum_128(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += sh[k];
    return s;
}

// K1.  grid = tiles, block = TILE_DET.
template <bool CALIB, bool WANTJ>
__global__ void __launch_bounds__(TILE_DET)
resjac_kernel(SplineView sp, const double* __restrict__ x, const double* __restrict__ camprep,
              const int* __restrict__ tile_cam, const int64_t* __restrict__ tile_start,
              const int* __restrict__ tile_cnt, const int64_t* __restrict__ row_off,
              const double* __restrict__ frame, const double* __restrict__ xr,
              const double* __restrict__ yr, const double* __restrict__ obs_u,
              const double* __restrict__ obs_v, int undist, int opt_sync, int opt_rs, int64_t N,
              double* __restrict__ r, double* __restrict__ Jb, double* __restrict__ partial) {
    __shared__ double s_cam[CAMPREP_DOUBLES];
    __shared__ double s_red[TILE_DET / 32];
    const int tl = blockIdx.x;
    const int cam = tile_cam[tl];
    for (int k = threadIdx.x; k < CAMPREP_DOUBLES; k += blockDim.x)
        s_cam[k] = camprep[(size_t)cam * CAMPREP_DOUBLES + k];
    __syncthreads();
    const CamPrep& c = *reinterpret_cast<const CamPrep*>(s_cam);
    const int cnt = tile_cnt[tl];
    double sq = 0.0;
    if ((int)threadIdx.x < cnt) {
        const int64_t d = tile_start[tl] + threadIdx.x;
        const double f = __ldcs(frame + d), yy = __ldcs(yr + d);
        double xx = 0.0, ou = 0.0, ov = 0.0;
        if (CALIB) xx = __ldcs(xr + d);
        else { ou = __ldcs(obs_u + d); ov = __ldcs(obs_v + d); }
        double ru, rv;
        int sp_out;
        FreeMask fm{opt_sync != 0, opt_rs != 0};
        if (WANTJ) {
            constexpr int P = 3 + (CALIB ? 15 : 6) + 12;
            const int lt = threadIdx.x & 31;
            double* blk = Jb + ((size_t)tl * (TILE_DET / 32) + (threadIdx.x >> 5)) * JBlk<P>::BLK_D;
            BlockSink sink{blk, lt, P};
            resjac_one<CALIB, true>(c, undist != 0, fm, f, xx, yy, ou, ov, sp, x, ru, rv, sp_out, sink);
            __stcs(blk + jblk_off(2 * P, lt), ru);
            __stcs(blk + jblk_off(2 * P + 1, lt), rv);
            reinterpret_cast<int*>(blk + JBlk<P>::SPAN_OFF)[lt] = sp_out;
        } else {
            NullSink sink;
            resjac_one<CALIB, false>(c, undist != 0, fm, f, xx, yy, ou, ov, sp, x, ru, rv, sp_out, sink);
        }
        const int64_t r0 = row_off[cam], ncam = (row_off[cam + 1] - r0) >> 1;
        const int64_t local = d - (r0 >> 1);
        __stcs(r + r0 + local, ru);
        __stcs(r + r0 + ncam + local, rv);
        sq = ru * ru + rv * rv;
    }
    const double s = block_sum_128(sq, s_red);
    if (threadIdx.x == 0) partial[tl] = s;
}

// K1, residual only (trial evaluations of the LM loop): TWO tiles per CTA, each thread carries one detection of
// each tile.  The per-detection work is one dependent chain of loads (frame -> time -> bucket -> span ->
// polynomial row -> coefficients); at 40 B of HBM traffic per detection the one-detection-per-thread kernel ran
// at 0.21 of the HBM roofline on that latency.  Two independent chains per thread double the loads in flight.
// grid = ceil(n_tiles / 2), block = TILE_DET.  partial[] keeps one slot per tile (the second one is zeroed).
template <bool CALIB>
__global__ void __launch_bounds__(TILE_DET)
residual2_kernel(SplineView sp, const double* __restrict__ x, const double* __restrict__ camprep,
                 const int* __restrict__ tile_cam, const int64_t* __restrict__ tile_start,
                 const int* __restrict__ tile_cnt, const int64_t* __restrict__ row_off,
                 const double* __restrict__ frame, const double* __restrict__ xr,
                 const double* __restrict__ yr, const double* __restrict__ obs_u,
                 const double* __restrict__ obs_v, int undist, int opt_sync, int opt_rs, int n_tiles,
                 double* __restrict__ r, double* __restrict__ partial) {
    __shared__ double s_cam[2][CAMPREP_DOUBLES];
    __shared__ double s_red[TILE_DET / 32];
    const int ta = 2 * blockIdx.x, tb = ta + 1;
    const bool hasb = tb < n_tiles;
    const int cama = tile_cam[ta], camb = hasb ? tile_cam[tb] : cama;
    for (int k = threadIdx.x; k < CAMPREP_DOUBLES; k += blockDim.x) {
        s_cam[0][k] = camprep[(size_t)cama * CAMPREP_DOUBLES + k];
        s_cam[1][k] = camprep[(size_t)camb * CAMPREP_DOUBLES + k];
    }
    __syncthreads();
    const CamPrep& ca = *reinterpret_cast<const CamPrep*>(s_cam[0]);
    const CamPrep& cb = *reinterpret_cast<const CamPrep*>(s_cam[1]);
    const bool ina = (int)threadIdx.x < tile_cnt[ta], inb = hasb && (int)threadIdx.x < tile_cnt[tb];
    const int64_t da = tile_start[ta] + threadIdx.x, db = hasb ? tile_start[tb] + threadIdx.x : da;
    // all inputs of both detections first
    double fa = 0.0, ya = 0.0, xa = 0.0, oua = 0.0, ova = 0.0, fb = 0.0, yb = 0.0, xb = 0.0, oub = 0.0, ovb = 0.0;
    if (ina) {
        fa = __ldcs(frame + da); ya = __ldcs(yr + da);
        if (CALIB) xa = __ldcs(xr + da); else { oua = __ldcs(obs_u + da); ova = __ldcs(obs_v + da); }
    }
    if (inb) {
        fb = __ldcs(frame + db); yb = __ldcs(yr + db);
        if (CALIB) xb = __ldcs(xr + db); else { oub = __ldcs(obs_u + db); ovb = __ldcs(obs_v + db); }
    }
    FreeMask fm{opt_sync != 0, opt_rs != 0};
    NullSink sink;
    double rua = 0.0, rva = 0.0, rub = 0.0, rvb = 0.0;
    int spa, spb;
    if (ina) resjac_one<CALIB, false>(ca, undist != 0, fm, fa, xa, ya, oua, ova, sp, x, rua, rva, spa, sink);
    if (inb) resjac_one<CALIB, false>(cb, undist != 0, fm, fb, xb, yb, oub, ovb, sp, x, rub, rvb, spb, sink);
    if (ina) {
        const int64_t r0 = row_off[cama], ncam = (row_off[cama + 1] - r0) >> 1, local = da - (r0 >> 1);
        __stcs(r + r0 + local, rua);
        __stcs(r + r0 + ncam + local, rva);
    }
    if (inb) {
        const int64_t r0 = row_off[camb], ncam = (row_off[camb + 1] - r0) >> 1, local = db - (r0 >> 1);
        __stcs(r + r0 + local, rub);
        __stcs(r + r0 + ncam + local, rvb);
    }
    const double s = block_sum_128(rua * rua + rva * rva + rub * rub + rvb * rvb, s_red);
    if (threadIdx.x == 0) { partial[ta] = s; if (hasb) partial[tb] = 0.0; }
}

// K1m.  grid = ceil(M / 128), block = 128.
template <bool WANTJ>
__global__ void __launch_bounds__(128)
motion_kernel(SplineView sp, const double* __restrict__ x, int type, double w,
              const double* __restrict__ tau, const int* __restrict__ tau_spl,
              const unsigned char* __restrict__ flags, int64_t M, double* __restrict__ r_motion,
              int* __restrict__ mbase, double* __restrict__ mJ, double* __restrict__ partial,
              int* __restrict__ err_flag) {
    __shared__ double s_red[4];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double sq = 0.0;
    if (j < M) {
        double rr, fa[3], fc[7];
        int base;
        const bool ok = motion_one<WANTJ>(type, w, sp, x, tau, tau_spl, flags, j, rr, base, fa, fc);
        if (!ok) atomicExch(err_flag, 1);
        r_motion[j] = rr;
        if (WANTJ) {
            mbase[j] = base;
#pragma unroll
            for (int k = 0; k < 3; ++k) mJ[(int64_t)k * M + j] = fa[k];
#pragma unroll
            for (int k = 0; k < 7; ++k) mJ[(int64_t)(3 + k) * M + j] = fc[k];
        }
        sq = rr * rr;
    }
    const double s = block_sum_128(sq, s_red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// Deterministic final reduction of per-CTA partial sums (single CTA): out[0] = sum.
__global__ void reduce_partial_kernel(const double* __restrict__ partial, int64_t n, double* __restrict__ out) {
    __shared__ double sh[1024];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// Diagnostic getter (mvus_ba_residual_jacobian): Jacobian blocks -> column planes J[2P][N] and span[N].
__global__ void deblock_kernel(const double* __restrict__ Jb, int P, const int64_t* __restrict__ tile_start,
                               const int* __restrict__ tile_cnt, int64_t N, double* __restrict__ J,
                               int* __restrict__ span) {
    const int tl = blockIdx.x, t = threadIdx.x;
    if (t >= tile_cnt[tl]) return;
    const int blk_d = (2 * P + 2) * 32 + 16;
    const double* blk = Jb + ((size_t)tl * (TILE_DET / 32) + (t >> 5)) * blk_d;
    const int64_t d = tile_start[tl] + t;
    for (int p = 0; p < 2 * P; ++p) J[(int64_t)p * N + d] = blk[jblk_off(p, t & 31)];
    span[d] = reinterpret_cast<const int*>(blk + (2 * P + 2) * 32)[t & 31];
}

// detections_global (common.py:105-127): time stamp and observation per detection.
__global__ void det_global_kernel(const double* __restrict__ camprep, const int* __restrict__ tile_cam,
                                  const int64_t* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                                  const double* __restrict__ frame, const double* __restrict__ xr,
                                  const double* __restrict__ yr, const double* __restrict__ obs_u,
                                  const double* __restrict__ obs_v, int calib, int undist,
                                  const int64_t* __restrict__ row_off, double* __restrict__ out) {
    const int tl = blockIdx.x;
    if ((int)threadIdx.x >= tile_cnt[tl]) return;
    const int cam = tile_cam[tl];
    const CamPrep& c = *reinterpret_cast<const CamPrep*>(camprep + (size_t)cam * CAMPREP_DOUBLES);
    const int64_t d = tile_start[tl] + threadIdx.x;
    const int64_t c0 = row_off[cam] >> 1, ncam = (row_off[cam + 1] >> 1) - c0;
    double* t = out + 3 * c0 + (d - c0);       // camera block: [t (ncam) | u (ncam) | v (ncam)]
    double* u = t + ncam;
    double* v = u + ncam;
    *t = c.alpha * (frame[d] + c.rho * (yr[d] * c.invH)) + c.beta;
    if (calib) {
        if (undist) {
            double xn, yn;
            undistort5(xr[d], yr[d], c.K4, c.d, xn, yn);
            *u = c.K4[0] * xn + c.K4[2];
            *v = c.K4[1] * yn + c.K4[3];
        } else { *u = xr[d]; *v = yr[d]; }
    } else { *u = obs_u[d]; *v = obs_v[d]; }
}

// visible[d] = 1-based interval id of detection d's time stamp, 0 if none (util.py:103-106)
__global__ void visibility_kernel(SplineView sp, const double* __restrict__ camprep, const int* __restrict__ tile_cam,
                                  const int64_t* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                                  const double* __restrict__ frame, const double* __restrict__ yr,
                                  long long* __restrict__ visible) {
    const int tl = blockIdx.x;
    if ((int)threadIdx.x >= tile_cnt[tl]) return;
    const CamPrep& c = *reinterpret_cast<const CamPrep*>(camprep + (size_t)tile_cam[tl] * CAMPREP_DOUBLES);
    const int64_t d = tile_start[tl] + threadIdx.x;
    const double t = c.alpha * (frame[d] + c.rho * (yr[d] * c.invH)) + c.beta;
    visible[d] = (long long)(find_interval(sp, t) + 1);
}

}  // namespace mvus
