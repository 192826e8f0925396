// Smoothing-spline fit on the device: the arithmetic of scipy.interpolate.splprep (FITPACK parcur / fppara)
// as Scene.traj_to_spline calls it (reconstruction/common.py:224-270, 247: splprep(part[1:], u=part[0], s=s, k=3);
// :267 the k = 1 fallback), and therefore of Scene.triangulate's refit (:754-815).
//
// Split of the work: FITPACK's knot-placement strategy and the search for the smoothing parameter are a
// few hundred scalar decisions per fit -- they stay on the host (mvus_b200/splfit.py, the product's mirror
// of fppara's control flow).  Everything that touches the m data points runs here, per fit iteration:
//   spl_normal_kernel   one thread per data point: knot interval (bisection over the knots), the k+1
//                       non-zero B-splines (de Boor / fpbspl), FP64 REDs into the banded normal matrix
//                       G (upper band, k+1 diagonals) and the idim right-hand sides
//   spl_chol_kernel     banded Cholesky of G (+ the smoothing penalty, half-bandwidth k+1) and the idim
//                       triangular solves; one CTA (the system has n-k-1 unknowns, a few thousand)
//   spl_resid_kernel    one thread per data point: spline value from the new coefficients, squared residual,
//                       block-reduced fp and FITPACK's per-knot-interval sums fpint (a data point ON an interior
//                       knot is split half / half between the two intervals, fppara label 140)
// FITPACK triangularises the observation matrix row by row with Givens rotations; forming G = A^T A is the same
// least-squares problem (conditioning is benign: B-splines of degree <= 3), and fp is summed from the actual
// residuals.  oracle/fitpack_oracle.py is the CPU restatement the tests pin against the installed splprep.
#pragma once
#include "ba_ctx.cuh"

struct mvus_spl_ctx {
    int device = 0, k = 3, idim = 3;
    int64_t m = 0;
    cudaStream_t st = nullptr;
    std::string err;
    mvus::DevBuf<double> u, x, t, G, rhs, c, fpint, term, scal, pen;
    double* h_pin = nullptr;
};

namespace mvus {

constexpr int SPL_KMAX = 3;

// t[l] <= u < t[l+1] with k <= l <= n-k-2 (the last data point belongs to the last interval)
__device__ __forceinline__ int spl_interval(const double* __restrict__ t, int n, int k, double u) {
    int lo = k, hi = n - k - 2;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid] <= u) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// FITPACK fpbspl: the k+1 non-zero B-splines of degree k at u, t[l] <= u < t[l+1]
__device__ __forceinline__ void spl_basis(const double* __restrict__ t, int k, double u, int l, double h[SPL_KMAX + 1]) {
    double hh[SPL_KMAX + 1];
    h[0] = 1.0;
    for (int j = 1; j <= k; ++j) {
        for (int i = 0; i < j; ++i) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 0; i < j; ++i) {
            const int li = l + i + 1, lj = li - j;
            const double f = hh[i] / (t[li] - t[lj]);
            h[i] += f * (t[li] - u);
            h[i + 1] = f * (u - t[lj]);
        }
    }
}

// G: band storage G[d * nk1 + j] = A[j - d][j] (d = 0 main diagonal .. hb), rhs[dim * nk1 + j]
__global__ void spl_normal_kernel(const double* __restrict__ u, const double* __restrict__ x, int64_t m, int idim,
                                  const double* __restrict__ t, int n, int k, double* __restrict__ G,
                                  double* __restrict__ rhs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int nk1 = n - k - 1;
    const double ui = u[i];
    const int l = spl_interval(t, n, k, ui);
    double h[SPL_KMAX + 1];
    spl_basis(t, k, ui, l, h);
    for (int a = 0; a <= k; ++a) {
        const int ia = l - k + a;
        for (int b = a; b <= k; ++b) atomicAdd(G + (size_t)(b - a) * nk1 + (l - k + b), h[a] * h[b]);
        for (int d = 0; d < idim; ++d) atomicAdd(rhs + (size_t)d * nk1 + ia, h[a] * x[(size_t)d * m + i]);
    }
}

// One CTA.  Adds pen (same band layout, hbp diagonals; may be null) scaled by pscale to G, factors the band
// (upper storage: U^T U with U[j-d][j] at band d), solves for the idim right-hand sides (thread d), returns the
// sum of the diagonal of the triangular factor in scal[1] (FITPACK's first guess of p) and a failure flag in scal[2].
__global__ void spl_chol_kernel(double* __restrict__ G, int nk1, int hb, const double* __restrict__ pen, double pscale,
                                const double* __restrict__ rhs, int idim, double* __restrict__ c,
                                double* __restrict__ scal) {
    if (pen)
        for (size_t i = threadIdx.x; i < (size_t)(hb + 1) * nk1; i += blockDim.x) G[i] += pscale * pen[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        double dsum = 0.0, bad = 0.0;
        for (int j = 0; j < nk1; ++j) {
            // column j of U: entries U[i][j], i = j-hb .. j, stored at band d = j - i
            for (int d = hb; d >= 0; --d) {
                const int i = j - d;
                if (i < 0) continue;
                double v = G[(size_t)d * nk1 + j];
                for (int e = d + 1; e <= hb; ++e) {           // sum over rows r = j - e < i of U[r][i] U[r][j]
                    const int r = j - e;
                    if (r < 0) break;
                    v -= G[(size_t)(e - d) * nk1 + i] * G[(size_t)e * nk1 + j];
                }
                if (d == 0) {
                    if (!(v > 0.0)) { bad = 1.0; v = 1.0; }
                    v = sqrt(v);
                    dsum += v;
                    G[j] = v;
                } else {
                    G[(size_t)d * nk1 + j] = v / G[i];
                }
            }
        }
        scal[1] = dsum;
        scal[2] = bad;
    }
    __syncthreads();
    if ((int)threadIdx.x < idim) {
        const double* b = rhs + (size_t)threadIdx.x * nk1;
        double* y = c + (size_t)threadIdx.x * nk1;
        for (int j = 0; j < nk1; ++j) {                     // U^T y = b
            double v = b[j];
            for (int d = 1; d <= hb && j - d >= 0; ++d) v -= G[(size_t)d * nk1 + j] * y[j - d];
            y[j] = v / G[j];
        }
        for (int j = nk1 - 1; j >= 0; --j) {                // U c = y
            double v = y[j];
            for (int d = 1; d <= hb && j + d < nk1; ++d) v -= G[(size_t)d * nk1 + j + d] * y[j + d];
            y[j] = v / G[j];
        }
    }
}

__global__ void __launch_bounds__(256)
spl_resid_kernel(const double* __restrict__ u, const double* __restrict__ x, int64_t m, int idim,
                 const double* __restrict__ t, int n, int k, const double* __restrict__ c,
                 double* __restrict__ term_out, double* __restrict__ fpint, double* __restrict__ scal) {
    __shared__ double sh[8];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nk1 = n - k - 1;
    double term = 0.0;
    if (i < m) {
        const double ui = u[i];
        const int l = spl_interval(t, n, k, ui);
        double h[SPL_KMAX + 1];
        spl_basis(t, k, ui, l, h);
        for (int d = 0; d < idim; ++d) {
            double s = 0.0;
            for (int a = 0; a <= k; ++a) s += h[a] * c[(size_t)d * nk1 + l - k + a];
            const double e = x[(size_t)d * m + i] - s;
            term += e * e;
        }
        if (term_out) term_out[i] = term;
        const int iv = l - k;                              // knot interval of the point
        if (ui == t[l] && l > k) {                         // on an interior knot: half to each side
            atomicAdd(fpint + iv - 1, 0.5 * term);
            atomicAdd(fpint + iv, 0.5 * term);
        } else {
            atomicAdd(fpint + iv, term);
        }
    }
    double v = term;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
        if (s != 0.0) atomicAdd(scal, s);
    }
}

}  // namespace mvus

static std::string g_spl_err;
#define SPL_CUDA(h, call)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (call);                                                            \
        if (_e != cudaSuccess) { (h)->err = std::string(#call) + ": " + cudaGetErrorString(_e); return MVUS_ERR_CUDA; } \
    } while (0)

extern "C" const char* mvus_ba_spl_last_error(mvus_spl_handle h) { return h ? h->err.c_str() : g_spl_err.c_str(); }

extern "C" int mvus_ba_spl_create(int32_t device, int64_t m, int32_t idim, int32_t k, const double* u, const double* x,
                                  mvus_spl_handle* out) {
    if (!out || !u || !x || m < 2 || idim < 1 || idim > 8 || k < 1 || k > mvus::SPL_KMAX || m <= k) {
        g_spl_err = "bad argument (m > k must hold)";
        return MVUS_ERR_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        cudaGetLastError();
        g_spl_err = "no usable CUDA device (there is no CPU fallback)";
        return MVUS_ERR_CUDA;
    }
    cudaSetDevice(device);
    mvus::library_pool(device);
    mvus_spl_ctx* h = new mvus_spl_ctx();
    h->device = device; h->m = m; h->idim = idim; h->k = k;
    cudaError_t e = cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) { h->h_pin = mvus::pin_scratch_acquire(); if (!h->h_pin) e = cudaErrorMemoryAllocation; }
    if (e == cudaSuccess) e = mvus::upload(h->u, u, (size_t)m, h->st);
    if (e == cudaSuccess) e = mvus::upload(h->x, x, (size_t)m * idim, h->st);
    if (e == cudaSuccess) e = h->term.alloc((size_t)m);
    if (e == cudaSuccess) e = h->scal.alloc(8);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) { g_spl_err = cudaGetErrorString(e); delete h; return MVUS_ERR_CUDA; }
    *out = h;
    return MVUS_OK;
}

extern "C" void mvus_ba_spl_destroy(mvus_spl_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->st) cudaStreamSynchronize(h->st);
    cudaStreamSynchronize(cudaStreamPerThread);
    for (auto* b : {&h->u, &h->x, &h->t, &h->G, &h->rhs, &h->c, &h->fpint, &h->term, &h->scal, &h->pen}) b->release(false);
    mvus::pin_scratch_release(h->h_pin);
    if (h->st) cudaStreamDestroy(h->st);
    delete h;
}

// One least-squares / smoothing solve on the knots t[n]:  min |x - s(u)|^2 + pscale * c^T P c  with the
// penalty band P = pen (k+2 diagonals of B^T B, fpdisc jumps; NULL for the plain least-squares spline).
// Outputs (any may be NULL): c[idim][n-k-1], fp, fpint[n-2k-1], diag_sum (sum of the diagonal of the
// triangular factor: FITPACK's first guess p = (n-k-1) / diag_sum).
extern "C" int mvus_ba_spl_solve(mvus_spl_handle h, int32_t n, const double* t, const double* pen, double pscale,
                                 double* c, double* fp, double* fpint, double* diag_sum) {
    if (!h || !t || n < 2 * (h->k + 1)) { if (h) h->err = "bad argument"; return MVUS_ERR_ARG; }
    SPL_CUDA(h, cudaSetDevice(h->device));
    const int k = h->k, nk1 = n - k - 1, nrint = nk1 - k, hb = pen ? k + 1 : k, idim = h->idim;
    SPL_CUDA(h, mvus::upload(h->t, t, (size_t)n, h->st));
    SPL_CUDA(h, h->G.alloc((size_t)(k + 2) * nk1));
    SPL_CUDA(h, h->rhs.alloc((size_t)idim * nk1));
    SPL_CUDA(h, h->c.alloc((size_t)idim * nk1));
    SPL_CUDA(h, h->fpint.alloc((size_t)nrint));
    SPL_CUDA(h, cudaMemsetAsync(h->G.p, 0, (size_t)(k + 2) * nk1 * sizeof(double), h->st));
    SPL_CUDA(h, cudaMemsetAsync(h->rhs.p, 0, (size_t)idim * nk1 * sizeof(double), h->st));
    SPL_CUDA(h, cudaMemsetAsync(h->fpint.p, 0, (size_t)nrint * sizeof(double), h->st));
    SPL_CUDA(h, cudaMemsetAsync(h->scal.p, 0, 8 * sizeof(double), h->st));
    if (pen) SPL_CUDA(h, mvus::upload(h->pen, pen, (size_t)(k + 2) * nk1, h->st));
    const int gb = (int)((h->m + 255) / 256);
    mvus::spl_normal_kernel<<<gb, 256, 0, h->st>>>(h->u.p, h->x.p, h->m, idim, h->t.p, n, k, h->G.p, h->rhs.p);
    mvus::spl_chol_kernel<<<1, 128, 0, h->st>>>(h->G.p, nk1, hb, pen ? h->pen.p : nullptr, pscale, h->rhs.p, idim,
                                               h->c.p, h->scal.p);
    mvus::spl_resid_kernel<<<gb, 256, 0, h->st>>>(h->u.p, h->x.p, h->m, idim, h->t.p, n, k, h->c.p, h->term.p,
                                                 h->fpint.p, h->scal.p);
    SPL_CUDA(h, cudaGetLastError());
    if (c) SPL_CUDA(h, cudaMemcpyAsync(c, h->c.p, (size_t)idim * nk1 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    if (fpint) SPL_CUDA(h, cudaMemcpyAsync(fpint, h->fpint.p, (size_t)nrint * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SPL_CUDA(h, cudaMemcpyAsync(h->h_pin, h->scal.p, 4 * sizeof(double), cudaMemcpyDeviceToHost, h->st));
    SPL_CUDA(h, cudaStreamSynchronize(h->st));
    if (h->h_pin[2] != 0.0) { h->err = "normal matrix of the spline fit is not positive definite"; return MVUS_ERR_NONFINITE; }
    if (fp) *fp = h->h_pin[0];
    if (diag_sum) *diag_sum = h->h_pin[1];
    return MVUS_OK;
}
