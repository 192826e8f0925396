// Chunk pre-reduction of the block-tridiagonal spline system (in front of the cyclic reduction of
// ba_solve.cuh).  NumPy model: tests/proto/bcr_proto.py::prereduce / solve_chunked.
//
// Why: cyclic reduction touches W~ (nb*q rows x ldw columns, 2.8 GB at config 4) once per LEVEL for every
// block that is still alive plus its two eliminated neighbours: ~5 |W~| of HBM traffic per solve, one
// kernel per level.  Here the super-blocks are cut into chunks of Lc consecutive blocks; the first block of
// a chunk is its HEAD, the other Lc-1 are eliminated one after the other in ascending order.  Eliminating
// block k only needs the previous block's result, so ONE thread per W~ column streams down the chunk with
// 2q doubles of state: every W~ row is read once and written once (2 |W~|, the algorithmic minimum), and the
// cyclic reduction that follows only sees the nb/Lc heads.
//
// Per eliminated block k (chunk head j0, D'_k = D_k - ZR_{k-1}^T ZR_{k-1}):
//   L_k L_k^T = D'_k,   ZR_k = L_k^-1 E_k          (coupling to k+1; for the last block: to the next head)
//                       ZH_k = L_k^-1 F_k          (fill-in coupling to the own head; F_{j0+1} = E_{j0}^T,
//                                                   F_{k+1} = -ZR_k^T ZH_k)
//   W~_k = L_k^-1 (W_k - ZR_{k-1}^T W~_{k-1})
// Head system (block tridiagonal again):
//   D^h_c = D_{j0} - sum_k ZH_k^T ZH_k - [ZR^T ZR of the last block of chunk c-1]
//   E^h_c = -ZH_last^T ZR_last,     W^h_c = W_{j0} - sum_k ZH_k^T W~_k - [ZR^T W~ of the last block of chunk c-1]
// Back substitution (descending k): ds_k = L_k^-T (v_k - ZH_k ds_head - ZR_k ds_{k+1}).
// Storage: L_k -> Dw[k], ZR_k -> Ew[k], ZH_k -> ZL[k] (the arrays cyclic reduction uses for the same roles).
#pragma once
#include "ba_ctx.cuh"

namespace mvus {

// ---- small-matrix part: one warp per chunk ------------------------------------------------
template <int Q>
__global__ void __launch_bounds__(32)
chunk_factor_kernel(int64_t nb, int Lc, int64_t c_first, double* __restrict__ Dw, double* __restrict__ Ew,
                    double* __restrict__ ZL, double* __restrict__ Linv, double* __restrict__ Dh,
                    double* __restrict__ Eh, double* __restrict__ DhR, int* __restrict__ fail_flag) {
    constexpr int QQ = Q * Q;
    __shared__ double Dk[QQ], F[QQ], R[QQ], Hd[QQ], T1[QQ], T2[QQ];
    const int lane = threadIdx.x;
    const int64_t c = c_first + blockIdx.x;
    const int64_t j0 = c * Lc, j1 = (j0 + Lc < nb) ? j0 + Lc : nb;
    if (j0 >= nb) return;
    bool bad = false;
    for (int i = lane; i < QQ; i += 32) Hd[i] = Dw[j0 * QQ + i];
    if (j1 - j0 == 1) {                                   // a head without followers keeps its original coupling
        __syncwarp();
        for (int i = lane; i < QQ; i += 32) {
            Dh[c * QQ + i] = Hd[i];
            Eh[c * QQ + i] = (j1 < nb) ? Ew[j0 * QQ + i] : 0.0;
            if (j1 < nb) DhR[(c + 1) * QQ + i] = 0.0;
        }
        return;
    }
    for (int i = lane; i < QQ; i += 32) {
        const int a = i / Q, b = i - a * Q;
        Dk[i] = Dw[(j0 + 1) * QQ + i];
        F[i] = Ew[j0 * QQ + b * Q + a];                   // E_{j0}^T : rows j0+1, columns head
    }
    __syncwarp();
    for (int64_t k = j0 + 1; k < j1; ++k) {
        const bool has_r = k + 1 < nb;
        for (int i = lane; i < QQ; i += 32) R[i] = has_r ? Ew[k * QQ + i] : 0.0;
        // Cholesky of Dk (lower, in place), column by column
#pragma unroll 1
        for (int cc = 0; cc < Q; ++cc) {
            double d = Dk[cc * Q + cc];
            if (!(d > 0.0)) { bad = true; d = 1.0; }
            d = sqrt(d);
            __syncwarp();
            if (lane == 0) Dk[cc * Q + cc] = d;
            for (int i = cc + 1 + lane; i < Q; i += 32) Dk[i * Q + cc] /= d;
            __syncwarp();
            // trailing update: entry (i, kk), cc < kk <= i, spread over the lanes
            const int nrem = Q - cc - 1;
            for (int e = lane; e < nrem * nrem; e += 32) {
                const int i = cc + 1 + e / nrem, kk = cc + 1 + e % nrem;
                if (kk <= i) Dk[i * Q + kk] -= Dk[i * Q + cc] * Dk[kk * Q + cc];
            }
            __syncwarp();
        }
        // ZR = L^-1 R, ZH = L^-1 F: one lane per column of [R | F]
        for (int col = lane; col < 2 * Q; col += 32) {
            double* Mx = col < Q ? R : F;
            const int cx = col < Q ? col : col - Q;
#pragma unroll 1
            for (int i = 0; i < Q; ++i) {
                double v = Mx[i * Q + cx];
#pragma unroll 1
                for (int kk = 0; kk < i; ++kk) v -= Dk[i * Q + kk] * Mx[kk * Q + cx];
                Mx[i * Q + cx] = v / Dk[i * Q + i];
            }
        }
        __syncwarp();
        const bool more = k + 1 < j1;
        for (int i = lane; i < QQ; i += 32) {
            const int a = i / Q, b = i - a * Q;
            double hh = 0.0, rr = 0.0, rf = 0.0;
#pragma unroll 1
            for (int kk = 0; kk < Q; ++kk) {
                const double fa = F[kk * Q + a], ra = R[kk * Q + a];
                hh += fa * F[kk * Q + b];
                rr += ra * R[kk * Q + b];
                rf += more ? ra * F[kk * Q + b] : fa * R[kk * Q + b];      // R^T F (next fill-in) or F^T R (head coupling)
            }
            Hd[i] -= hh;
            T1[i] = rr; T2[i] = rf;
            Dw[k * QQ + i] = Dk[i];
            Ew[k * QQ + i] = R[i];
            ZL[k * QQ + i] = F[i];
            if (a == b) Linv[k * Q + a] = 1.0 / Dk[i];
        }
        __syncwarp();
        if (more) {
            for (int i = lane; i < QQ; i += 32) { Dk[i] = Dw[(k + 1) * QQ + i] - T1[i]; F[i] = -T2[i]; }
        } else {
            for (int i = lane; i < QQ; i += 32) {
                Eh[c * QQ + i] = has_r ? -T2[i] : 0.0;
                if (has_r) DhR[(c + 1) * QQ + i] = T1[i];
            }
        }
        __syncwarp();
    }
    for (int i = lane; i < QQ; i += 32) Dh[c * QQ + i] = Hd[i];
    if (__any_sync(0xffffffffu, bad) && lane == 0) atomicExch(fail_flag, 1);
}

// ---- W~ part: one thread per column streams down the chunk --------------------------------
// grid (chunks, column groups of CW_T), block CW_T.  Wsrc: where a block's original rows are read (may be Ww
// itself: every element is read before the same thread overwrites it).  Writes W~_k into Ww (zero rows for
// the head), the head's rows into Wh[c] and the contribution to the NEXT head into Gh[c+1].
// Per block k the CTA stages [L_k | ZH_k | ZR_k | 1/diag L_k] in shared memory (cp.async, double-buffered: the
// matrices of block k+1 arrive while block k is computed; every thread reads them as broadcasts) and each
// thread keeps three q-vectors: w (this block's rows of its column), wnx (the next block's rows, prefetched
// one block ahead and already carrying -ZR_k^T W~_k when its turn comes) and wh (the head's accumulator).
constexpr int CW_T = 128;
__device__ __forceinline__ void cw_cp8(double* dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
template <int Q>
__global__ void __launch_bounds__(CW_T, 4)
chunk_w_kernel(int64_t nb, int Lc, int64_t c_first, int ldw, const double* __restrict__ Wsrc,
               const double* __restrict__ Dw, const double* __restrict__ Ew, const double* __restrict__ ZL,
               const double* __restrict__ Linv, double* __restrict__ Ww, double* __restrict__ Wh,
               double* __restrict__ Gh) {
    constexpr int QQ = Q * Q;
    constexpr int NM = 3 * QQ + Q;                        // doubles per stage
    constexpr int NMP = (NM + 1) & ~1;
    __shared__ __align__(16) double sm[2][NMP];           // [stage][L | ZH | ZR | 1/diag]
    const int tid = threadIdx.x;
    const int64_t c = c_first + blockIdx.x;
    const int64_t j0 = c * Lc, j1 = (j0 + Lc < nb) ? j0 + Lc : nb;
    if (j0 >= nb) return;
    const int col = blockIdx.y * CW_T + tid;
    const bool act = col < ldw;
    const int64_t wn = (int64_t)Q * ldw;
    const int cl = act ? col : 0;
    double wh[Q], w[Q], wnx[Q];
#pragma unroll
    for (int a = 0; a < Q; ++a) { wh[a] = Wsrc[j0 * wn + (int64_t)a * ldw + cl]; wnx[a] = 0.0; }
    if (j1 - j0 > 1) {
        auto stage_mats = [&](int64_t k, int st) {
            for (int e = tid; e < NM; e += CW_T) {
                const double* src = e < QQ ? Dw + k * QQ + e : e < 2 * QQ ? ZL + k * QQ + (e - QQ)
                                  : e < 3 * QQ ? Ew + k * QQ + (e - 2 * QQ) : Linv + k * Q + (e - 3 * QQ);
                cw_cp8(&sm[st][e], src);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        stage_mats(j0 + 1, 0);
#pragma unroll
        for (int a = 0; a < Q; ++a) wnx[a] = Wsrc[(j0 + 1) * wn + (int64_t)a * ldw + cl];
        for (int64_t k = j0 + 1; k < j1; ++k) {
            const int st = (int)((k - j0 - 1) & 1);
            const bool more = k + 1 < j1;
#pragma unroll
            for (int a = 0; a < Q; ++a) { w[a] = wnx[a]; wnx[a] = 0.0; }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                              // block k's matrices are there; stage st^1 is free again
            if (more) {
                stage_mats(k + 1, st ^ 1);
#pragma unroll
                for (int a = 0; a < Q; ++a) wnx[a] = Wsrc[(k + 1) * wn + (int64_t)a * ldw + cl];
            }
            const double* Lm = sm[st];
            const double* Zh = sm[st] + QQ;
            const double* Zr = sm[st] + 2 * QQ;
            const double* Li = sm[st] + 3 * QQ;
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                double v = w[i];
#pragma unroll
                for (int kk = 0; kk < i; ++kk) v -= Lm[i * Q + kk] * w[kk];
                w[i] = v * Li[i];
            }
            if (act) {
#pragma unroll
                for (int a = 0; a < Q; ++a) Ww[k * wn + (int64_t)a * ldw + col] = w[a];
            }
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                const double x = w[b];
#pragma unroll
                for (int a = 0; a < Q; ++a) wh[a] -= Zh[b * Q + a] * x;
            }
            // -ZR_k^T W~_k goes to the next block's rows (or, after the last block, to the next head)
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                const double x = w[b];
#pragma unroll
                for (int a = 0; a < Q; ++a) wnx[a] -= Zr[b * Q + a] * x;
            }
        }
    }
    if (!act) return;
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        Wh[c * wn + (int64_t)a * ldw + col] = wh[a];
        Ww[j0 * wn + (int64_t)a * ldw + col] = 0.0;
    }
    if (j1 < nb) {
        // contribution to the next head: ZR_last^T W~_last = -wnx (zero for a lone head, whose coupling stays in E^h)
#pragma unroll
        for (int a = 0; a < Q; ++a) Gh[(c + 1) * wn + (int64_t)a * ldw + col] = (j1 - j0 > 1) ? -wnx[a] : 0.0;
    }
}

// Heads c in [c_lo, c_hi]: D^h -= DhR, W^h -= Gh (the left chunk's contributions).  grid (c_hi - c_lo + 1).
__global__ void head_fix_kernel(int64_t c_lo, int q, int ldw, double* __restrict__ Dh, const double* __restrict__ DhR,
                                double* __restrict__ Wh, const double* __restrict__ Gh) {
    const int64_t c = c_lo + blockIdx.x;
    const int qq = q * q;
    const int64_t wn = (int64_t)q * ldw;
    for (int i = threadIdx.x; i < qq; i += blockDim.x) Dh[c * qq + i] -= DhR[c * qq + i];
    for (int64_t i = threadIdx.x; i < wn; i += blockDim.x) Wh[c * wn + i] -= Gh[c * wn + i];
}

// Heads back into their rows of Ww (the Schur SYRK runs over all rows of Ww).  grid (heads).
__global__ void head_rows_kernel(int64_t c_lo, int Lc, int q, int ldw, const double* __restrict__ Wh,
                                 double* __restrict__ Ww) {
    const int64_t c = c_lo + blockIdx.x;
    const int64_t wn = (int64_t)q * ldw;
    for (int64_t i = threadIdx.x; i < wn; i += blockDim.x) Ww[c * Lc * wn + i] = Wh[c * wn + i];
}

// Back substitution inside the chunks: ds holds v_k = W~_k[rhs] - W~_k[0:ncP] . dc for every row (wdc_kernel),
// dsh the solved heads.  One warp per chunk, descending k.
template <int Q>
__global__ void __launch_bounds__(32)
chunk_back_kernel(int64_t nb, int Lc, int64_t c_first, const double* __restrict__ Dw, const double* __restrict__ Ew,
                  const double* __restrict__ ZL, const double* __restrict__ dsh, double* __restrict__ ds) {
    constexpr int QQ = Q * Q;
    const int lane = threadIdx.x;
    const int64_t c = c_first + blockIdx.x;
    const int64_t j0 = c * Lc, j1 = (j0 + Lc < nb) ? j0 + Lc : nb;
    if (j0 >= nb) return;
    const int a = lane < Q ? lane : 0;
    const double xh = dsh[c * Q + a];                       // lane a: component a of the head's solution
    double xn = (j1 < nb) ? dsh[(c + 1) * Q + a] : 0.0;     // solution of block k+1 (starts with the next head)
    if (lane < Q) ds[j0 * Q + lane] = xh;
    for (int64_t k = j1 - 1; k > j0; --k) {
        double v = ds[k * Q + a];
        const double* zh = ZL + k * QQ + a * Q;
        const double* zr = Ew + k * QQ + a * Q;
#pragma unroll
        for (int b = 0; b < Q; ++b) {
            v -= zh[b] * __shfl_sync(0xffffffffu, xh, b);
            v -= zr[b] * __shfl_sync(0xffffffffu, xn, b);
        }
        // x = L^-T v : x_i = (v_i - sum_{cc > i} L[cc][i] x_cc) / L[i][i]; lane a keeps the column L[.][a]
        const double* L = Dw + k * QQ;
        double lcol[Q];
#pragma unroll
        for (int i = 0; i < Q; ++i) lcol[i] = L[i * Q + a];
        double x = 0.0;
#pragma unroll
        for (int i = Q - 1; i >= 0; --i) {
            const double xi = __shfl_sync(0xffffffffu, v, i) / __shfl_sync(0xffffffffu, lcol[i], i);
            if (a == i) x = xi;
            if (a < i) v -= lcol[i] * xi;
        }
        if (lane < Q) ds[k * Q + lane] = x;
        xn = x;
    }
}

}  // namespace mvus
