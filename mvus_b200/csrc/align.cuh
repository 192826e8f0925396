// Ground-truth alignment (analysis/compare_gt.py:73-151 of the reference) on the device: batched similarity
// fits between a spline evaluated at shifted times and a fixed point set.
//
//   coarse search (compare_gt.py:112-126): the spline is the INTERPOLATING spline of the ground truth
//       (util.match_overlap, util.py:119-135), the points are the reconstruction sampled at the GT rate, one
//       fit per integer shift -- hundreds of independent fits, one CTA each;
//   fine stage (compare_gt.py:35-70): the spline is the reconstruction, the points are the GT samples at
//       t = alpha * t_gt + beta, one fit per residual evaluation of the 2-parameter least-squares problem.
//
// One fit = thirdparty/transformation.py:869-975 affine_matrix_from_points(shear=False, scale=True): centroids,
// rotation that maximises trace(R H) for H = sum v1 v0^T, scale = sqrt(sum |v1|^2 / sum |v0|^2).  The reference
// takes the rotation from an SVD of H (Kabsch) with the reflection fix; here it is Horn's unit quaternion (largest
// eigenvector of the 4x4 matrix built from H, cyclic Jacobi) -- the same maximiser, always a proper rotation.
// A point takes part iff its time lies in a spline interval by the rule of util.sampling ((t>=a) xor (t>=b)).
#pragma once
#include "ba_ctx.cuh"

namespace mvus {

constexpr int AL_T = 256;

__device__ __forceinline__ double al_block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < AL_T / 32; ++w) s += red[w];
    return s;
}

// symmetric 4x4 eigen-decomposition by cyclic Jacobi; returns the eigenvector of the largest eigenvalue
__device__ inline void al_top_eigvec4(double A[4][4], double q[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, dia = 0.0;
        for (int i = 0; i < 4; ++i) { dia += A[i][i] * A[i][i]; for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j]; }
        if (off <= 1e-32 * dia || off == 0.0) break;
        for (int p = 0; p < 3; ++p)
            for (int r = p + 1; r < 4; ++r) {
                if (A[p][r] == 0.0) continue;
                const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
                const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akr = A[k][r];
                    A[k][p] = c * akp - s * akr; A[k][r] = s * akp + c * akr;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], ark = A[r][k];
                    A[p][k] = c * apk - s * ark; A[r][k] = s * apk + c * ark;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkr = V[k][r];
                    V[k][p] = c * vkp - s * vkr; V[k][r] = s * vkp + c * vkr;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i) if (A[i][i] > A[best][best]) best = i;
    double nrm = 0.0;
    for (int k = 0; k < 4; ++k) { q[k] = V[k][best]; nrm += q[k] * q[k]; }
    nrm = 1.0 / sqrt(nrm);
    for (int k = 0; k < 4; ++k) q[k] *= nrm;
}

// grid = shifts.  tau[n], pts = 3 planes of n.  spline_is_src: the fit maps the spline points onto `pts`
// (fine stage) or `pts` onto the spline points (coarse search).  err (may be null): per-point distance for
// shift `want`, 0 where the point takes no part.
__global__ void __launch_bounds__(AL_T)
align_fit_kernel(SplineView sp, const double* __restrict__ x, const double* __restrict__ tau,
                 const double* __restrict__ pts, int64_t n, const double* __restrict__ shift, int spline_is_src,
                 int want, double* __restrict__ mean_err, int64_t* __restrict__ count, double* __restrict__ Mout,
                 double* __restrict__ err) {
    __shared__ double red[AL_T / 32];
    __shared__ double Ms[12];
    const int b = blockIdx.x, tid = threadIdx.x;
    const double sh = shift[b];
    // pass 1: members and centroids
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int64_t j = tid; j < n; j += AL_T) {
        const double t = tau[j] + sh;
        const int s = find_interval(sp, t);
        if (s < 0) continue;
        double X[3], dX[3], B[4];
        spline_eval<false>(sp, x, s, t, X, dX, B);
        acc[0] += 1.0;
        for (int a = 0; a < 3; ++a) { acc[1 + a] += X[a]; acc[4 + a] += pts[(int64_t)a * n + j]; }
    }
    double tot[7];
    for (int k = 0; k < 7; ++k) tot[k] = al_block_sum(acc[k], red);
    const double cnt = tot[0];
    if (cnt < 3.0) {                                       // affine_matrix_from_points raises for fewer than ndims points
        if (tid == 0) {
            mean_err[b] = __longlong_as_double(0x7ff0000000000000LL);
            count[b] = (int64_t)cnt;
            for (int k = 0; k < 16; ++k) Mout[(int64_t)b * 16 + k] = (k % 5 == 0) ? 1.0 : 0.0;
        }
        if (err && b == want) for (int64_t j = tid; j < n; j += AL_T) err[j] = 0.0;
        return;
    }
    double cs[3], cp[3];
    for (int a = 0; a < 3; ++a) { cs[a] = tot[1 + a] / cnt; cp[a] = tot[4 + a] / cnt; }
    // pass 2: centred moments.  v0 = source - its centroid, v1 = target - its centroid, H[i][j] = sum v0_i v1_j
    double h[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t j = tid; j < n; j += AL_T) {
        const double t = tau[j] + sh;
        const int s = find_interval(sp, t);
        if (s < 0) continue;
        double X[3], dX[3], B[4], v0[3], v1[3];
        spline_eval<false>(sp, x, s, t, X, dX, B);
        for (int a = 0; a < 3; ++a) {
            const double vs = X[a] - cs[a], vp = pts[(int64_t)a * n + j] - cp[a];
            v0[a] = spline_is_src ? vs : vp;
            v1[a] = spline_is_src ? vp : vs;
        }
        for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) h[i * 3 + k] += v0[i] * v1[k];
        h[9] += v0[0] * v0[0] + v0[1] * v0[1] + v0[2] * v0[2];
        h[10] += v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
    }
    double H[11];
    for (int k = 0; k < 11; ++k) H[k] = al_block_sum(h[k], red);
    if (tid == 0) {
        // transformation.py:952-962 (Horn): N from the sums xx = sum v0_x v1_x, xy = sum v0_x v1_y, ...
        const double xx = H[0], xy = H[1], xz = H[2], yx = H[3], yy = H[4], yz = H[5], zx = H[6], zy = H[7], zz = H[8];
        double N[4][4] = {{xx + yy + zz, yz - zy, zx - xz, xy - yx},
                          {yz - zy, xx - yy - zz, xy + yx, zx + xz},
                          {zx - xz, xy + yx, yy - xx - zz, yz + zy},
                          {xy - yx, zx + xz, yz + zy, zz - xx - yy}};
        double q[4];
        al_top_eigvec4(N, q);
        const double w = q[0], qx = q[1], qy = q[2], qz = q[3];
        double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * w), 2 * (qx * qz + qy * w),
                       2 * (qx * qy + qz * w), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * w),
                       2 * (qx * qz - qy * w), 2 * (qy * qz + qx * w), 1 - 2 * (qx * qx + qy * qy)};
        const double sc = sqrt(H[10] / H[9]);
        const double* c0 = spline_is_src ? cs : cp;
        const double* c1 = spline_is_src ? cp : cs;
        for (int i = 0; i < 3; ++i) {
            double tr = c1[i];
            for (int k = 0; k < 3; ++k) { Ms[i * 4 + k] = sc * R[i * 3 + k]; tr -= sc * R[i * 3 + k] * c0[k]; }
            Ms[i * 4 + 3] = tr;
        }
        for (int k = 0; k < 12; ++k) Mout[(int64_t)b * 16 + k] = Ms[k];
        Mout[(int64_t)b * 16 + 12] = 0.0; Mout[(int64_t)b * 16 + 13] = 0.0; Mout[(int64_t)b * 16 + 14] = 0.0;
        Mout[(int64_t)b * 16 + 15] = 1.0;
        count[b] = (int64_t)cnt;
    }
    __syncthreads();
    // pass 3: distances |target - M source|
    double es = 0.0;
    const bool keep = err && b == want;
    for (int64_t j = tid; j < n; j += AL_T) {
        const double t = tau[j] + sh;
        const int s = find_interval(sp, t);
        double e = 0.0;
        if (s >= 0) {
            double X[3], dX[3], B[4], P[3];
            spline_eval<false>(sp, x, s, t, X, dX, B);
            for (int a = 0; a < 3; ++a) P[a] = pts[(int64_t)a * n + j];
            const double* src = spline_is_src ? X : P;
            const double* dst = spline_is_src ? P : X;
            double d2 = 0.0;
            for (int i = 0; i < 3; ++i) {
                const double ti = Ms[i * 4] * src[0] + Ms[i * 4 + 1] * src[1] + Ms[i * 4 + 2] * src[2] + Ms[i * 4 + 3];
                d2 += (dst[i] - ti) * (dst[i] - ti);
            }
            e = sqrt(d2);
            es += e;
        }
        if (keep) err[j] = e;
    }
    const double esum = al_block_sum(es, red);
    if (tid == 0) mean_err[b] = esum / cnt;
}

}  // namespace mvus
