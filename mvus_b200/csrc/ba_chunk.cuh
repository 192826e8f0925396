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

// Register / shuffle version of chunk_factor_kernel for Q <= 15 (2Q <= 32 lanes): the shared-memory version above
// needs ~18 us per block (a dozen __syncwarp-separated phases with shared-memory round trips), and that chain of
// Lc - 1 dependent blocks is pure latency that neither a bigger problem nor more GPUs hide.  Here
//   lane i <  Q ("row lane")   holds row i of the current diagonal block (dh) and column i of E_k / ZR_k (m);
//   lane Q + b ("F lane")      holds column b of the fill-in F_k / ZH_k (m) and column b of the head's block (dh).
// Cholesky: right-looking, column j's pivot and multipliers travel by shuffles; the triangular solves read
// L[i][kk] from row lane i; ONE shuffle stream of ZR[kk][a] serves both ZR^T ZR (row lanes -> next diagonal block)
// and ZR^T ZH (F lanes -> next fill-in), a second one of ZH[kk][a] serves ZH^T ZH (head) and, for the last block,
// ZH^T ZR (head coupling).  The next block's D and E are prefetched at the top of the iteration.
template <int Q>
__global__ void __launch_bounds__(32)
chunk_factor_fast_kernel(int64_t nb, int Lc, int64_t c_first, double* __restrict__ Dw, double* __restrict__ Ew,
                         double* __restrict__ ZL, double* __restrict__ Linv, double* __restrict__ Dh,
                         double* __restrict__ Eh, double* __restrict__ DhR, int* __restrict__ fail_flag) {
    static_assert(2 * Q <= 32, "one lane per column of [E | F]");
    constexpr int QQ = Q * Q;
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const int64_t c = c_first + blockIdx.x;
    const int64_t j0 = c * Lc, j1 = (j0 + Lc < nb) ? j0 + Lc : nb;
    if (j0 >= nb) return;
    if (j1 - j0 == 1) {                                   // a head without followers keeps its original coupling
        for (int i = lane; i < QQ; i += 32) {
            Dh[c * QQ + i] = Dw[j0 * QQ + i];
            Eh[c * QQ + i] = (j1 < nb) ? Ew[j0 * QQ + i] : 0.0;
            if (j1 < nb) DhR[(c + 1) * QQ + i] = 0.0;
        }
        return;
    }
    const bool isrow = lane < Q, isf = lane >= Q && lane < 2 * Q;
    const int li = isrow ? lane : 0, lb = isf ? lane - Q : 0;
    double m[Q], dh[Q], nxt[Q], rn[Q];
    double rinv = 1.0;                                    // row lane i: 1 / L[i][i] of the current block
    bool bad = false;
#pragma unroll
    for (int a = 0; a < Q; ++a) {
        // row lanes: row li of D_{j0+1}, column li of E_{j0+1}; F lanes: column lb of E_{j0}^T and of the head's block
        dh[a] = isrow ? Dw[(j0 + 1) * QQ + li * Q + a] : Dw[j0 * QQ + a * Q + lb];
        m[a] = isrow ? ((j0 + 2 < nb) ? Ew[(j0 + 1) * QQ + a * Q + li] : 0.0) : Ew[j0 * QQ + lb * Q + a];
    }
    for (int64_t k = j0 + 1; k < j1; ++k) {
        const bool more = k + 1 < j1, has_r = k + 1 < nb;
        // prefetch block k + 1 (row lanes): its diagonal block row and its coupling column
#pragma unroll
        for (int a = 0; a < Q; ++a) {
            nxt[a] = (more && isrow) ? Dw[(k + 1) * QQ + li * Q + a] : 0.0;
            rn[a] = (more && isrow && k + 2 < nb) ? Ew[(k + 1) * QQ + a * Q + li] : 0.0;
        }
        // Cholesky of the diagonal block (rows in the row lanes)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double djj = __shfl_sync(FULL, dh[j], j);
            if (!(djj > 0.0)) { bad = true; djj = 1.0; }
            const double inv = 1.0 / sqrt(djj);
            if (lane == j) rinv = inv;
            const double lij = dh[j] * inv;                // row lane i >= j: L[i][j]
            if (isrow) dh[j] = lij;
#pragma unroll
            for (int k2 = j + 1; k2 < Q; ++k2) {
                const double lkj = __shfl_sync(FULL, lij, k2);
                if (isrow) dh[k2] -= lij * lkj;            // entry (i, k2); meaningful for k2 <= i
            }
        }
        // ZR = L^-1 E, ZH = L^-1 F: every lane solves its column
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            double v = m[i];
#pragma unroll
            for (int kk = 0; kk < i; ++kk) v -= __shfl_sync(FULL, dh[kk], i) * m[kk];
            m[i] = v * __shfl_sync(FULL, rinv, i);
        }
        // store L (rows), 1/diag, ZR (row lanes' columns), ZH (F lanes' columns)
#pragma unroll
        for (int a = 0; a < Q; ++a) {
            if (isrow) { Dw[k * QQ + li * Q + a] = dh[a]; Ew[k * QQ + a * Q + li] = m[a]; }
            if (isf) ZL[k * QQ + a * Q + lb] = m[a];
        }
        if (isrow) Linv[k * Q + li] = rinv;
        // nxt -= ZR[:, a] . own column: row lanes -> D_{k+1} - ZR^T ZR, F lanes -> -ZR^T ZH = next fill-in
#pragma unroll
        for (int kk = 0; kk < Q; ++kk)
#pragma unroll
            for (int a = 0; a < Q; ++a) nxt[a] -= __shfl_sync(FULL, m[kk], a) * m[kk];
        if (!more && has_r && isrow) {
#pragma unroll
            for (int a = 0; a < Q; ++a) DhR[(c + 1) * QQ + li * Q + a] = -nxt[a];       // (ZR^T ZR)[li][a]
        }
        // head: Hd[:, b] -= ZH[:, a] . ZH[:, b] (F lanes).  In the last block the row lanes collect -ZH^T ZR
        // (the head coupling) from the same shuffle stream, in rn (no next block to prefetch then)
#pragma unroll
        for (int kk = 0; kk < Q; ++kk)
#pragma unroll
            for (int a = 0; a < Q; ++a) {
                const double bf = __shfl_sync(FULL, m[kk], Q + a) * m[kk];
                dh[a] -= isf ? bf : 0.0;
                rn[a] -= (isrow && !more) ? bf : 0.0;
            }
        if (more) {
#pragma unroll
            for (int a = 0; a < Q; ++a) {
                if (isrow) dh[a] = nxt[a];
                m[a] = isrow ? rn[a] : nxt[a];
            }
        } else if (isrow) {
#pragma unroll
            for (int a = 0; a < Q; ++a) Eh[c * QQ + a * Q + li] = has_r ? rn[a] : 0.0;  // (-ZH^T ZR)[a][li]
        }
    }
    if (isf) {
#pragma unroll
        for (int a = 0; a < Q; ++a) Dh[c * QQ + a * Q + lb] = dh[a];
    }
    if (bad && lane == 0) atomicExch(fail_flag, 1);
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
            // (the matrices are read as shared-memory broadcasts: one LDS per FMA made the kernel L1TEX-bound at
            //  87 % -- for even Q two neighbouring entries come with one 16-byte load)
            constexpr bool V2 = (Q % 2 == 0);
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                double v = w[i];
                if (V2) {
#pragma unroll
                    for (int kk = 0; kk + 1 < i; kk += 2) {
                        const double2 l2 = *reinterpret_cast<const double2*>(Lm + i * Q + kk);
                        v -= l2.x * w[kk];
                        v -= l2.y * w[kk + 1];
                    }
                    if (i & 1) v -= Lm[i * Q + i - 1] * w[i - 1];
                } else {
#pragma unroll
                    for (int kk = 0; kk < i; ++kk) v -= Lm[i * Q + kk] * w[kk];
                }
                w[i] = v * Li[i];
            }
            if (act) {
#pragma unroll
                for (int a = 0; a < Q; ++a) Ww[k * wn + (int64_t)a * ldw + col] = w[a];
            }
#pragma unroll
            for (int b = 0; b < Q; ++b) {
                const double x = w[b];
                if (V2) {
#pragma unroll
                    for (int a = 0; a < Q; a += 2) {
                        const double2 h2 = *reinterpret_cast<const double2*>(Zh + b * Q + a);
                        const double2 r2 = *reinterpret_cast<const double2*>(Zr + b * Q + a);
                        wh[a] -= h2.x * x; wh[a + 1] -= h2.y * x;
                        wnx[a] -= r2.x * x; wnx[a + 1] -= r2.y * x;     // -ZR_k^T W~_k: next block's rows / next head
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < Q; ++a) { wh[a] -= Zh[b * Q + a] * x; wnx[a] -= Zr[b * Q + a] * x; }
                }
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
