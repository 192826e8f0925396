// Scene.BA(motion_prior=True): the discrete-trajectory mode of the reference (reconstruction/common.py:
// 466-467, 527-550, 587-605, 631-634, 681-687).  CPU restatement: oracle/points_oracle.py.
//
// Unknowns: camera side (alpha, beta, rho, camera vectors) + the G points of global_traj; the splines are
// constants.  Two handles carry it:
//   hs  the ordinary handle of the flight (detections + splines): reprojection rows r, J and the camera blocks
//       of the normal equations come from its K1 with the spline coefficients held fixed;
//   hp  a handle whose "spline" is the point list (one pseudo-spline with G coefficients, no detections): its
//       block-tridiagonal solver arrays D / E / W~ hold the point-point and point-camera blocks of the motion
//       rows, so the exact solve of ba_solve.cuh (chunk pre-reduction, cyclic reduction, Schur SYRK, dense
//       Cholesky) and the LM driver run unchanged.  x of hp = [camera side | X plane | Y plane | Z plane].
// Motion rows (error_motion(motion_prior=True), common.py:386-403): the time stamp of point j follows the
// camera that saw it, t_j = alpha_c (f_j + rho_c y_j / H_c) + beta_c (common.py:128-148); points are grouped by
// the spline interval t_j falls in (a <= t < b) in global_traj order; Scene.motion_prior (959-1001) on each
// group; the row sits at the middle point (F) or the later point (KE).  Each row therefore couples up to three
// points AND the time parameters of up to three cameras: the camera-camera part of J^T J is no longer block
// diagonal; cross-camera entries go to the dense addend Ax of form_schur_kernel.
#pragma once
#include "ba_ctx.cuh"

namespace mvus {

constexpr int PT_SCAN = 64;        // how far the previous / next member of the same interval group is searched

struct PointsView {
    int64_t G;
    const int* cam;                // camera slot of each point
    const double* frame;           // frame id of its detection
    const double* yH;              // raw y / image height of its detection
    int nc, n_other;
    int motion_type;               // 1 = F, 2 = KE
    double w;
    int free_sync, free_rs;
};

__device__ __forceinline__ double pt_time(const PointsView& pv, const double* x, int64_t j) {
    const int c = pv.cam[j];
    return x[c] * (pv.frame[j] + x[2 * pv.nc + c] * pv.yH[j]) + x[pv.nc + c];
}

// One thread per point j: residual of the row that sits at j, and (WANTJ) its factors.
//   pidx[3][G]  the row's points p < j < n (-1 = absent)          pJ[9][G]: 0-2 fa (axis factors),
//   3-5 fc (point factors of p, j, n: d r / d P_{k,ax} = fa[ax] fc[k]), 6-8 d r / d t of p, j, n
template <bool WANTJ>
__global__ void points_motion_kernel(PointsView pv, SplineView sp, const double* __restrict__ x,
                                     double* __restrict__ r, int* __restrict__ pidx, double* __restrict__ pJ,
                                     double* __restrict__ partial, int* __restrict__ flag) {
    __shared__ double red[4];
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t G = pv.G;
    double rj = 0.0;
    if (j < G) {
        const double eps = 1e-20;
        const double tj = pt_time(pv, x, j);
        const int gj = find_interval(sp, tj);
        int64_t p = -1, n = -1;
        bool far = false;
        // previous / next member of the same interval group in global_traj order.  Groups follow each other in
        // time and the list was sorted by time, so a point of an EARLIER (later) interval ends the backward
        // (forward) search; points outside every interval are skipped.
        if (gj >= 0) {
            for (int64_t k = j - 1; k >= 0; --k) {
                if (j - k > PT_SCAN) { far = true; break; }
                const int g = find_interval(sp, pt_time(pv, x, k));
                if (g == gj) { p = k; break; }
                if (g >= 0 && g < gj) break;
            }
            if (pv.motion_type == 1)
                for (int64_t k = j + 1; k < G; ++k) {
                    if (k - j > PT_SCAN) { far = true; break; }
                    const int g = find_interval(sp, pt_time(pv, x, k));
                    if (g == gj) { n = k; break; }
                    if (g > gj) break;
                }
        }
        const bool row = gj >= 0 && p >= 0 && (pv.motion_type == 2 || n >= 0);
        // the solver's super-blocks hold 3 points: a row may span at most 4 consecutive points
        if (far || (row && ((pv.motion_type == 1 ? n : j) - p > 3))) atomicExch(flag, 1);
        double fa[3] = {0, 0, 0}, fc[3] = {0, 0, 0}, gt[3] = {0, 0, 0};
        if (row) {
            const double* X = x + pv.n_other;
            double Pj[3], Pp[3];
            for (int a = 0; a < 3; ++a) { Pj[a] = X[a * G + j]; Pp[a] = X[a * G + p]; }
            const double tp = pt_time(pv, x, p);
            if (pv.motion_type == 2) {
                const double dt = tj - tp, idt = 1.0 / (dt + eps), aw = fabs(pv.w), sg = dt < 0 ? -1.0 : 1.0;
                double s2 = 0.0, v[3];
                for (int a = 0; a < 3; ++a) { v[a] = (Pj[a] - Pp[a]) * idt; s2 += v[a] * v[a]; }
                rj = 0.5 * aw * sg * dt * s2;
                if (WANTJ) {
                    for (int a = 0; a < 3; ++a) fa[a] = aw * v[a] * dt * idt * sg;
                    fc[0] = -1.0; fc[1] = 1.0;
                    const double ddt = 0.5 * aw * sg * s2 * (1.0 - 2.0 * dt * idt);
                    gt[0] = -ddt; gt[1] = ddt;
                }
            } else {
                const double tn = pt_time(pv, x, n);
                double Pn[3];
                for (int a = 0; a < 3; ++a) Pn[a] = X[a * G + n];
                const double dt1 = tj - tp, dt2 = tn - tj, dt3 = dt1 + dt2;
                const double i1 = 1.0 / (dt1 + eps), i2 = 1.0 / (dt2 + eps), i3 = 1.0 / (dt3 + eps);
                const double kap = dt3 * i3, dkap = i3 * (1.0 - kap);
                double d1 = 0.0, d2 = 0.0;
                for (int a = 0; a < 3; ++a) {
                    const double v1 = (Pj[a] - Pp[a]) * i1, v2 = (Pn[a] - Pj[a]) * i2;
                    const double g = pv.w * (v2 - v1) * kap;
                    const double s = g < 0.0 ? -1.0 : 1.0;
                    rj += s * g;
                    if (WANTJ) {
                        fa[a] = s * pv.w * kap;
                        d1 += s * (pv.w * v1 * i1 * kap + pv.w * (v2 - v1) * dkap);
                        d2 += s * (-pv.w * v2 * i2 * kap + pv.w * (v2 - v1) * dkap);
                    }
                }
                if (WANTJ) {
                    fc[0] = i1; fc[1] = -(i1 + i2); fc[2] = i2;
                    gt[0] = -d1; gt[1] = d1 - d2; gt[2] = d2;
                }
            }
        }
        r[j] = rj;
        if (WANTJ) {
            pidx[j] = row ? (int)p : -1;
            pidx[G + j] = row ? (int)j : -1;
            pidx[2 * G + j] = (row && pv.motion_type == 1) ? (int)n : -1;
            for (int k = 0; k < 3; ++k) { pJ[k * G + j] = fa[k]; pJ[(3 + k) * G + j] = fc[k]; pJ[(6 + k) * G + j] = gt[k]; }
        }
    }
    double s = rj * rj;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) partial[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

// out[0] = a[0] + sum b[0..nb)
__global__ void points_cost_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t nb,
                                   double* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nb; i += blockDim.x) s += b[i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = a[0];
        for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
        out[0] = t;
    }
}

// Camera blocks of the reprojection rows: A[c][a][b] = sum_det (Ju_a Ju_b + Jv_a Jv_b), bc[c][a] = -sum (Ju_a r_u
// + Jv_a r_v) from the de-blocked Jacobian planes J[2P][N] (deblock_kernel).  grid (nc, Pc * (Pc + 1)).
__global__ void cam_blocks_kernel(const double* __restrict__ J, const double* __restrict__ r, int P, int Pc, int64_t N,
                                  const int64_t* __restrict__ row_off, double* __restrict__ A, double* __restrict__ bc) {
    __shared__ double red[8];
    const int c = blockIdx.x, e = blockIdx.y;
    const int a = e / (Pc + 1), b = e - a * (Pc + 1);          // b == Pc: right-hand side entry
    if (b < Pc && b > a) return;                               // lower triangle, mirrored below
    const int64_t d0 = row_off[c] >> 1, nd = (row_off[c + 1] >> 1) - d0;
    const double* ru = r + row_off[c];
    const double* rv = ru + nd;
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < nd; i += blockDim.x) {
        const double ua = J[(int64_t)a * N + d0 + i], va = J[(int64_t)(P + a) * N + d0 + i];
        if (b < Pc) s += ua * J[(int64_t)b * N + d0 + i] + va * J[(int64_t)(P + b) * N + d0 + i];
        else s -= ua * ru[i] + va * rv[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)blockDim.x / 32; ++w) t += red[w];
        if (b < Pc) {
            atomicAdd(A + ((int64_t)c * Pc + a) * Pc + b, t);
            if (a != b) atomicAdd(A + ((int64_t)c * Pc + b) * Pc + a, t);
        } else {
            atomicAdd(bc + (int64_t)c * Pc + a, t);
        }
    }
}

// Motion rows -> normal equations of hp.  One thread per row (= per point j).
__global__ void points_accum_kernel(PointsView pv, const double* __restrict__ x, const double* __restrict__ r,
                                    const int* __restrict__ pidx, const double* __restrict__ pJ, int bw, int Pc,
                                    int ldw, double* __restrict__ A, double* __restrict__ bc, double* __restrict__ Ax,
                                    double* __restrict__ D, double* __restrict__ E, double* __restrict__ W) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t G = pv.G;
    if (j >= G || pidx[G + j] < 0) return;
    const int q = 3 * bw, ncP = pv.nc * Pc;
    const double rr = r[j];
    int pt[3]; double fa[3], fc[3], gt[3];
    for (int k = 0; k < 3; ++k) { pt[k] = pidx[k * G + j]; fa[k] = pJ[k * G + j]; fc[k] = pJ[(3 + k) * G + j]; gt[k] = pJ[(6 + k) * G + j]; }
    // camera-side entries: (column, value), up to 3 points x (alpha, beta, rho)
    int ccol[9]; double cval[9]; int ne = 0;
    for (int k = 0; k < 3; ++k) {
        if (pt[k] < 0 || gt[k] == 0.0) continue;
        const int c = pv.cam[pt[k]];
        const double rho = x[2 * pv.nc + c], al = x[c], yH = pv.yH[pt[k]], f = pv.frame[pt[k]];
        if (pv.free_sync) {
            ccol[ne] = c * Pc; cval[ne++] = gt[k] * (f + rho * yH);
            ccol[ne] = c * Pc + 1; cval[ne++] = gt[k];
        }
        if (pv.free_rs) { ccol[ne] = c * Pc + 2; cval[ne++] = gt[k] * al * yH; }
    }
    for (int e1 = 0; e1 < ne; ++e1) {
        atomicAdd(bc + ccol[e1], -cval[e1] * rr);
        for (int e2 = 0; e2 < ne; ++e2) {
            const int c1 = ccol[e1] / Pc, c2 = ccol[e2] / Pc;
            const double v = cval[e1] * cval[e2];
            if (c1 == c2) atomicAdd(A + ((int64_t)c1 * Pc + (ccol[e1] - c1 * Pc)) * Pc + (ccol[e2] - c2 * Pc), v);
            else atomicAdd(Ax + (int64_t)ccol[e1] * ncP + ccol[e2], v);
        }
    }
    for (int k1 = 0; k1 < 3; ++k1) {
        if (pt[k1] < 0 || fc[k1] == 0.0) continue;
        const int j1 = pt[k1], kb1 = j1 / bw;
        for (int a1 = 0; a1 < 3; ++a1) {
            const double v1 = fc[k1] * fa[a1];
            const int l1 = (j1 - kb1 * bw) * 3 + a1;
            double* wrow = W + ((int64_t)kb1 * q + l1) * ldw;
            atomicAdd(wrow + (ldw - 1), -v1 * rr);
            for (int e = 0; e < ne; ++e) atomicAdd(wrow + ccol[e], v1 * cval[e]);
            for (int k2 = k1; k2 < 3; ++k2) {
                if (pt[k2] < 0 || fc[k2] == 0.0) continue;
                const int j2 = pt[k2], kb2 = j2 / bw;
                for (int a2 = (k2 == k1 ? a1 : 0); a2 < 3; ++a2) {
                    const double v = v1 * fc[k2] * fa[a2];
                    const int l2 = (j2 - kb2 * bw) * 3 + a2;
                    if (kb1 == kb2) atomicAdd(D + ((int64_t)kb1 * q + l1) * q + l2, v);      // l1 <= l2: upper triangle
                    else atomicAdd(E + ((int64_t)kb1 * q + l1) * q + l2, v);
                }
            }
        }
    }
}

}  // namespace mvus
