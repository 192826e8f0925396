// K2 (normal-equation accumulation) and K3 (damped solve + Levenberg-Marquardt driver).
//
// Replaces, for the BA of common.py:670, SciPy's TRF inner loop: LSMR on [J; sqrt(reg) I]
// (scipy/optimize/_lsq/trf.py:488-500), the 2-D subspace trust-region solve (:508) and the
// radius update (:526-541).  Here each LM iteration solves the damped normal equations
//      [ A + l dA    W^T     ] [dc]   [bc]          A : camera blocks (nc x Pc x Pc, block diagonal)
//      [ W           B + l dB] [ds] = [bs]          B : spline block, block-banded
// EXACTLY: the spline unknowns are grouped into super-blocks of `bw` control points so that B is
// block tridiagonal, eliminated by block cyclic reduction (log2(nb) fully parallel levels; a
// sequential banded Cholesky would be 2e5 dependent steps at config 4), the Schur complement
// S = A - sum_k (L_k^-1 W_k)^T (L_k^-1 W_k) is formed by an FP64 SYRK and factorised densely.
// tests/proto/bcr_proto.py is the NumPy model of exactly this sequence.
//
// Why direct and not PCG (profiles/r2_pcg_experiment.txt, tests/proto/pcg_experiment.py): J^T J has 7
// gauge zeros and weak time-warp modes (1e-3 .. 1 against a largest eigenvalue ~1e9); the Marquardt-
// damped, Jacobi-scaled matrix the solver actually sees has condition ~6e4 at lambda = 1e-4, and
// block-Jacobi PCG needs 100-300 iterations for a step accurate to 1e-3 -- it does converge, but at
// config 4 an iteration is at least one pass over W~ (2.8 GB, 0.45 ms) and ~200 of them cost 5x the
// exact solve (18 ms), which is cheap here because only nc*Pc <= 1152 unknowns survive the Schur
// complement.
#pragma once
#include "ba_ctx.cuh"
#include "ba_k2.cuh"

namespace mvus {

int evaluate(mvus_ba_ctx* h, const double* xd, bool want_j);
inline void owner_range(const mvus_ba_ctx* h, int r, int64_t* lo, int64_t* hi);    // ba_nccl.cuh

constexpr int QMAX = 18;          // 3 * max control points per super-block (bw <= 6)
constexpr double DIAG_MIN = 1e-6, DIAG_MAX = 1e32, DIAG_FLOOR_FRAC = 1e-2;

// K2m: motion rows -> spline block only.  One thread per sample.
__global__ void accumulate_motion_kernel(const double* __restrict__ r_motion, const int* __restrict__ mbase,
                                         const double* __restrict__ mJ, int64_t M, int bw, int ldw,
                                         int64_t c_lo, int64_t c_hi, double* __restrict__ D,
                                         double* __restrict__ E, double* __restrict__ W) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const int base = mbase[j];
    if (base < 0 || base < c_lo || base >= c_hi) return;    // (multi-GPU: rows of another rank's control points)
    const int q = 3 * bw;
    double fa[3], fc[7];
    for (int k = 0; k < 3; ++k) fa[k] = mJ[(int64_t)k * M + j];
    for (int k = 0; k < 7; ++k) fc[k] = mJ[(int64_t)(3 + k) * M + j];
    const double rr = r_motion[j];
    for (int k1 = 0; k1 < 7; ++k1) {
        if (fc[k1] == 0.0) continue;
        const int j1 = base + k1, kb1 = j1 / bw;
        for (int a1 = 0; a1 < 3; ++a1) {
            const double v1 = fc[k1] * fa[a1];
            const int l1 = (j1 - kb1 * bw) * 3 + a1;
            atomicAdd(W + ((int64_t)kb1 * q + l1) * ldw + (ldw - 1), -v1 * rr);
            for (int k2 = k1; k2 < 7; ++k2) {
                if (fc[k2] == 0.0) continue;
                const int j2 = base + k2, kb2 = j2 / bw;
                for (int a2 = (k2 == k1 ? a1 : 0); a2 < 3; ++a2) {
                    const double v = v1 * fc[k2] * fa[a2];
                    const int l2 = (j2 - kb2 * bw) * 3 + a2;
                    if (kb1 == kb2) {
                        atomicAdd(D + ((int64_t)kb1 * q + l1) * q + l2, v);      // l1 <= l2: upper triangle
                    } else {
                        atomicAdd(E + ((int64_t)kb1 * q + l1) * q + l2, v);
                    }
                }
            }
        }
    }
}

// Marquardt diagonals (clamped) and damped working copies.
// Marquardt diagonals (clamped) and the spline right-hand side b_s = W~[:, last] as a vector.
// Rows outside [row_lo, row_hi) (another rank's block range) are written as 0 and filled in by
// the all-reduce that follows.
__global__ void diag_kernel(const double* __restrict__ A, const double* __restrict__ D,
                            const double* __restrict__ W, int nc, int Pc, int64_t nbq, int q, int ldw,
                            int64_t n_ctrl3, int64_t row_lo, int64_t row_hi, double* __restrict__ diag_c,
                            double* __restrict__ diag_s, double* __restrict__ bs) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < (int64_t)nc * Pc) {
        const int cam = (int)(i / Pc), p = (int)(i - (int64_t)cam * Pc);
        diag_c[i] = fmin(fmax(A[((int64_t)cam * Pc + p) * Pc + p], DIAG_MIN), DIAG_MAX);
    }
    if (i < nbq) {
        const int64_t kb = i / q;
        const int l = (int)(i - kb * q);
        const bool mine = i >= row_lo && i < row_hi;
        // padding unknowns (beyond the last control point) get a unit diagonal, no damping
        diag_s[i] = (mine && i < n_ctrl3) ? fmin(fmax(D[(kb * q + l) * q + l], DIAG_MIN), DIAG_MAX) : 0.0;
        bs[i] = mine ? W[i * ldw + (ldw - 1)] : 0.0;
    }
}

__global__ void damp_copy_kernel(const double* __restrict__ D, const double* __restrict__ diag_s, const double* __restrict__ lam_p,
                                 int64_t nbq, int q, int64_t n_ctrl3, double* __restrict__ Dw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over nb*q*q
    if (i >= nbq * q) return;
    const int64_t row = i / q;
    const int col = (int)(i - row * q), l = (int)(row % q);
    double v = l <= col ? D[i] : D[(row - l + col) * q + l];      // D holds its upper triangle (K2, K2m)
    if (l == col) v = row < n_ctrl3 ? v + lam_p[0] * diag_s[row] : 1.0;
    Dw[i] = v;
}

// Floor on the control-point scaling: d_i = max(d_i, frac * mean(d)).  Control points all carry
// the same unit (metres); ones with almost no data (ends of an interval) would otherwise get
// almost no damping and jump by metres in one step while the quartic KE prior explodes
// (measured on the 'covered' test flight).
__global__ void sum_kernel(const double* __restrict__ v, int64_t n, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s = i < n ? v[i] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s != 0.0) atomicAdd(out, s);
}
__global__ void floor_kernel(double* __restrict__ v, int64_t n, const double* __restrict__ sum, double frac) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = fmax(v[i], frac * sum[0] / (double)n);
}

// ------------------------------------------------------------------------------------------
// One level of block cyclic reduction.  grid = blocks active at this level (indices j = i*s),
// block = 128 threads.  sp = s/2 is the previous level's stride (0 at level 0).
//   - apply the pending Schur updates of level-1 neighbours j -+ sp (if they were eliminated),
//   - if j is an odd multiple of s (or the root pass): factor D_j, form ZL/ZR/W~ rows.
// ZR is stored in E[j]; root = final pass on block 0.
template <int Q>
__global__ void __launch_bounds__(256, 4)
bcr_level_kernel(int64_t nb, int ldw, int64_t s, int64_t sp, int root, int odd_only,
                 const double* __restrict__ Wsrc, double* __restrict__ Dw,
                 double* __restrict__ Ew, double* __restrict__ Ww, double* __restrict__ ZL,
                 int* __restrict__ fail_flag) {
    constexpr int q = Q;
    __shared__ double zl_m[Q * Q], zr_m[Q * Q];   // ZL, ZR of the LEFT eliminated neighbour (j - sp)
    __shared__ double zl_p[Q * Q], zr_p[Q * Q];   // ZL, ZR of the RIGHT eliminated neighbour (j + sp)
    __shared__ double Dj[Q * Q], El[Q * Q], Er[Q * Q];
    __shared__ int s_bad;
    // root: 0 = regular level, 1 = final pass on block 0, 2 = update-only pass (sharded solve:
    // apply the pending updates of the last local level and leave the bridged coupling in Ew[j])
    const bool upd_only = (root == 2);
    if (upd_only) root = 0;
    // odd_only (level 0): only the blocks that are eliminated are launched; survivors have nothing to do
    const int64_t j = root ? 0 : (odd_only ? (2 * (int64_t)blockIdx.x + 1) * s : (int64_t)blockIdx.x * s);
    const bool elim = !upd_only && (root || (((j / s) & 1) == 1));
    const int tid = threadIdx.x, nt = blockDim.x;
    const int qq = q * q;
    const bool has_m = sp > 0 && j - sp >= 0 && (((j - sp) / sp) & 1);
    const bool has_p = sp > 0 && j + sp < nb && (((j + sp) / sp) & 1);
#pragma unroll 1
    for (int i = tid; i < qq; i += nt) {
        Dj[i] = Dw[j * qq + i];
        zl_m[i] = has_m ? ZL[(j - sp) * qq + i] : 0.0;
        zr_m[i] = has_m ? Ew[(j - sp) * qq + i] : 0.0;
        zl_p[i] = has_p ? ZL[(j + sp) * qq + i] : 0.0;
        zr_p[i] = has_p ? Ew[(j + sp) * qq + i] : 0.0;
    }
    if (tid == 0) s_bad = 0;
    __syncthreads();
    // D_j -= ZR_m^T ZR_m + ZL_p^T ZL_p ; couplings of the eliminated block
#pragma unroll 1
    for (int i = tid; i < qq; i += nt) {
        const int a = i / q, b = i - a * q;
        double acc = 0.0, el = 0.0, er = 0.0;
#pragma unroll 1
        for (int k = 0; k < q; ++k) {
            acc += zr_m[k * q + a] * zr_m[k * q + b] + zl_p[k * q + a] * zl_p[k * q + b];
            el -= zr_m[k * q + a] * zl_m[k * q + b];      // rows j, cols j - s   (bridge through j - sp)
            er -= zl_p[k * q + a] * zr_p[k * q + b];      // rows j, cols j + s   (bridge through j + sp)
        }
        if (sp > 0) { Dj[i] -= acc; El[i] = el; Er[i] = er; }
    }
    if (sp == 0 && elim && !root) {
        // level 0: original couplings.  E[j-1] has rows j-1, cols j -> transpose; E[j] rows j, cols j+1
#pragma unroll 1
        for (int i = tid; i < qq; i += nt) {
            const int a = i / q, b = i - a * q;
            El[i] = Ew[(j - 1) * qq + b * q + a];
            Er[i] = (j + 1 < nb) ? Ew[j * qq + i] : 0.0;
        }
    }
    __syncthreads();
    if (elim) {
        // Cholesky of Dj (lower, in place), q <= 18: one warp, rows in registers (lane = row), pivot and
        // multipliers by shuffles -- no shared-memory round trips inside the q dependent steps
        if (tid < 32) {
            double row[Q];
#pragma unroll
            for (int c = 0; c < Q; ++c) row[c] = tid < Q ? Dj[tid * Q + c] : 0.0;
            bool bad = false;
#pragma unroll
            for (int c = 0; c < Q; ++c) {
                double d = __shfl_sync(0xffffffffu, row[c], c);
                if (!(d > 0.0)) { bad = true; d = 1.0; }
                const double inv = 1.0 / sqrt(d);
                const double lic = row[c] * inv;
                row[c] = lic;
#pragma unroll
                for (int k2 = c + 1; k2 < Q; ++k2) row[k2] -= lic * __shfl_sync(0xffffffffu, lic, k2);
            }
            if (tid < Q) {
#pragma unroll
                for (int c = 0; c < Q; ++c) Dj[tid * Q + c] = row[c];
            }
            if (bad && tid == 0) s_bad = 1;
        }
        __syncthreads();
        // ZL = L^-1 El, ZR = L^-1 Er : thread per column of [El | Er]
#pragma unroll 1
        for (int c = tid; c < 2 * q; c += nt) {
            double* Mx = c < q ? El : Er;
            const int cc = c < q ? c : c - q;
#pragma unroll 1
            for (int i = 0; i < q; ++i) {
                double v = Mx[i * q + cc];
#pragma unroll 1
                for (int k = 0; k < i; ++k) v -= Dj[i * q + k] * Mx[k * q + cc];
                Mx[i * q + cc] = v / Dj[i * q + i];
            }
        }
        __syncthreads();
        if (!root)
#pragma unroll 1
            for (int i = tid; i < qq; i += nt) { ZL[j * qq + i] = El[i]; Ew[j * qq + i] = Er[i]; }
    }
#pragma unroll 1
    for (int i = tid; i < qq; i += nt) Dw[j * qq + i] = Dj[i];
    if (upd_only && sp > 0)
        for (int i = tid; i < qq; i += nt) Ew[j * qq + i] = (j + s < nb) ? Er[i] : 0.0;
    // W~ columns: pending update, then forward substitution if eliminated
    const double* Wm = Ww + (j - sp) * (int64_t)q * ldw;
    const double* Wp = Ww + (j + sp) * (int64_t)q * ldw;
    double* Wj = Ww + j * (int64_t)q * ldw;
    // first touch of a block reads the undamped original W~ (no separate working copy pass)
    const double* Wj_in = Wsrc ? Wsrc + j * (int64_t)q * ldw : Wj;
    for (int c = tid; c < ldw; c += nt) {
        double w[Q];
#pragma unroll
        for (int a = 0; a < Q; ++a) w[a] = Wj_in[(int64_t)a * ldw + c];
        if (has_m)
#pragma unroll 3
            for (int k = 0; k < Q; ++k) {
                const double x = Wm[(int64_t)k * ldw + c];
#pragma unroll
                for (int a = 0; a < Q; ++a) w[a] -= zr_m[k * Q + a] * x;
            }
        if (has_p)
#pragma unroll 3
            for (int k = 0; k < Q; ++k) {
                const double x = Wp[(int64_t)k * ldw + c];
#pragma unroll
                for (int a = 0; a < Q; ++a) w[a] -= zl_p[k * Q + a] * x;
            }
        if (elim) {
#pragma unroll
            for (int i = 0; i < Q; ++i) {
                asm volatile("" ::: "memory");      // keep row i's L loads inside iteration i (registers)
                double v = w[i];
#pragma unroll
                for (int k = 0; k < i; ++k) v -= Dj[i * Q + k] * w[k];
                w[i] = v / Dj[i * Q + i];
            }
        }
#pragma unroll
        for (int a = 0; a < Q; ++a) Wj[(int64_t)a * ldw + c] = w[a];
    }
    if (tid == 0 && s_bad) atomicExch(fail_flag, 1);
}

// S~ = sum over rows of W~^T W~ (lower-triangle tiles): the Schur complement's dense FP64
// GEMM (ncP^2 * 3 n_ctrl flops = 200 GFLOP at config 4).  FP64 MMA (mma.sync m8n8k4 f64 -- the
// only FP64 tensor instruction; tcgen05 has no f64 kind), 128x128 output tile per CTA, 16 warps
// each owning a 32x32 warp tile = 4x4 fragments (32 accumulator doubles per thread).  K chunks of
// 16 rows go through a 4-stage cp.async ring in shared memory (+4 padded leading dimension: the
// fragment loads are bank-conflict free, lane -> (k = lane&3, m = lane>>2) -> 4*k + m distinct
// 8-byte banks per half warp) with ONE __syncthreads per chunk; rows of W~ are 8-byte aligned only
// (ldw = ncP + 1 is odd), hence 8-byte copies, zero-filled past the slab / the last column.
// Split-K over row slabs, partial tiles reduced with FP64 RED.  Only the ncP camera columns enter
// the GEMM (576 = 18 x 32 columns need no padding strip): the right-hand-side column (W~^T w_rhs, one row
// of S~) is accumulated on the scalar FP64 pipe by an otherwise idle warp of the diagonal tile pairs.
// History: v1 4x4 scalar micro-tiles 4.4 TFLOP/s, v2 8x8 scalar 6.2 TFLOP/s, v3 FP64 MMA with a
// single-buffered register-prefetch pipeline 23 TFLOP/s issued (DMMA pipe 60 %; profiles/r1_notes.md).
constexpr int SY_T = 128, SY_K = 32, SY_LD = SY_T + 4, SY_STAGES = 3, SY_NQ = SY_K / 4;
constexpr size_t SY_SMEM = (size_t)SY_STAGES * 2 * SY_K * SY_LD * sizeof(double);

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void sy_cp8(unsigned dst, const double* src, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(ok ? 8u : 0u) : "memory");
}

__global__ void __launch_bounds__(512)
syrk_kernel(const double* __restrict__ Ww, int64_t R, int ldw, int n, int slab, int nbig, int small,
            double* __restrict__ Sfull) {
    extern __shared__ __align__(16) double sy_smem[];
    int p = blockIdx.x, ti = 0;
    while (p >= ti + 1) { p -= ti + 1; ++ti; }
    const int tj = p;
    const bool diag = (ti == tj);
    // row slab: nbig slabs of `slab` rows, then slabs of `small` rows (launch_syrk)
    const int by = (int)blockIdx.y;
    const int64_t r0 = by < nbig ? (int64_t)by * slab : (int64_t)nbig * slab + (int64_t)(by - nbig) * small;
    const int64_t r1e = r0 + (by < nbig ? slab : small);
    const int64_t r1 = r1e < R ? r1e : R;
    const int nchunk = (int)((r1 - r0 + SY_K - 1) / SY_K);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // 4 x 4 warps of 32 x 32.  Warp w issues on scheduler w & 3: in a diagonal tile pair only the warp
    // tiles wx <= wy work, so they are handed out in an order that spreads them over the four schedulers
    // (with wx = w & 3 scheduler 0 had 4 working warps and scheduler 3 one: the CTA took as long as a full tile).
    int wy = warp >> 2, wx = warp & 3;
    if (diag) {
        if (warp < 10) {                                   // the 10 lower warp tiles, row by row
            wy = warp >= 6 ? 3 : warp >= 3 ? 2 : warp >= 1 ? 1 : 0;
            wx = warp - wy * (wy + 1) / 2;
        } else {                                           // the 6 idle ones (wx > wy)
            const int idx = warp - 10;
            wy = idx < 3 ? 0 : idx < 5 ? 1 : 2;
            wx = idx < 3 ? idx + 1 : idx < 5 ? idx - 1 : 3;
        }
    }
    const int fk = lane & 3, fm = lane >> 2;               // fragment coordinates
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
    // warp tiles that only hold padding columns (last tile) or lie strictly above the diagonal of a
    // diagonal tile pair contribute nothing: they still help staging, but skip the MMAs
    const bool warp_active = (ti * SY_T + wy * 32 < n) && (tj * SY_T + wx * 32 < n) && !(diag && wx > wy);
    // The right-hand-side row of S~ (W~^T w_rhs) is accumulated by one of the six warps a diagonal tile pair leaves
    // idle: the five diagonal pairs cover every column once per row slab, the data are in shared memory anyway,
    // and the separate GEMV pass over W~ (0.45 ms, 2.8 GB at config 4) is gone.
    const bool rhs_warp = diag && warp == 15;
    double racc[4] = {0.0, 0.0, 0.0, 0.0};
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sy_smem);
    // Staging: each thread copies 4 elements of each panel per chunk -- column cc = tid & 127 of the rows
    // kq, kq + 4, kq + 8, kq + 12 (kq = tid >> 7).  Everything that does not change from chunk to chunk is
    // computed once (the generic form cost ~220 instructions per warp and chunk, issued right after the
    // barrier while the FP64-MMA pipe ran dry: profiles/r2_notes.md): two running row pointers, three row
    // offsets, the per-thread zero-fill sizes of padding columns.
    constexpr unsigned STAGE_B = 2u * SY_K * SY_LD * 8u, PANEL_B = SY_K * SY_LD * 8u, ROW4_B = 4u * SY_LD * 8u;
    const int cc = tid & 127, kq = tid >> 7;
    const bool cola = ti * SY_T + cc < n, colb = tj * SY_T + cc < n;
    const unsigned sza = cola ? 8u : 0u, szb = colb ? 8u : 0u;
    const double* pa = Ww + (r0 + kq) * ldw + (cola ? ti * SY_T + cc : 0);
    const double* pb = Ww + (r0 + kq) * ldw + (colb ? tj * SY_T + cc : 0);
    const unsigned soff = (unsigned)(kq * SY_LD + cc) * 8u;
    const int64_t ld4 = 4 * (int64_t)ldw, ld16 = SY_K * (int64_t)ldw;
    const int nfull = (int)((r1 - r0) / SY_K);             // chunks whose 16 rows all exist
    // Pipeline synchronisation without a CTA-wide rendezvous: full[st] completes when all 512 threads' copies
    // of the chunk in stage st have landed (cp.async.mbarrier.arrive.noinc), empty[st] when all 16 warps have
    // finished reading it.  Warps drift apart by up to a chunk, so the FP64-MMA pipe of a scheduler always has
    // a warp with queued work (with __syncthreads per chunk it ran dry at every barrier: 7.7 ms -> see notes).
    __shared__ __align__(8) unsigned long long sy_bar[2 * SY_STAGES];
    const unsigned bar0 = (unsigned)__cvta_generic_to_shared(sy_bar);
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < SY_STAGES; ++st) { k2_mbar_init(bar0 + 8u * st, 512u); k2_mbar_init(bar0 + 8u * (SY_STAGES + st), 16u); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // quarter k of the copies of chunk c (row kq + 4k of both panels)
    auto issue_q = [&](int c, int k) {
        if (c < nchunk) {
            const int st = c % SY_STAGES;
            if (k == 0 && c >= SY_STAGES) k2_mbar_wait(bar0 + 8u * (SY_STAGES + st), (unsigned)((c / SY_STAGES - 1) & 1));
            const unsigned sa = sbase + (unsigned)st * STAGE_B + soff + k * ROW4_B;
            const bool inr = c < nfull || r0 + (int64_t)c * SY_K + kq + 4 * k < r1;   // rows past r1 are zero-filled
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sa), "l"(inr ? pa + k * ld4 : Ww), "r"(inr ? sza : 0u) : "memory");
            if (!diag)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sa + PANEL_B), "l"(inr ? pb + k * ld4 : Ww), "r"(inr ? szb : 0u) : "memory");
            else if (k == 0 && tid < SY_K) {
                // diagonal tile pairs do not use the B panel: its first SY_K slots take the right-hand-side column
                // W~[row][n] of the chunk's rows for the fused rhs row (see the rhs warp below)
                const int64_t row = r0 + (int64_t)c * SY_K + tid;
                const bool rin = row < r1;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(sbase + (unsigned)st * STAGE_B + PANEL_B + (unsigned)tid * 8u),
                             "l"(rin ? Ww + row * ldw + n : Ww), "r"(rin ? 8u : 0u) : "memory");
            }
            if (k == SY_NQ - 1) {
                pa += ld16; pb += ld16;
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(bar0 + 8u * st) : "memory");
            }
        }
    };
    auto issue = [&](int c) {
#pragma unroll
        for (int k = 0; k < SY_NQ; ++k) issue_q(c, k);
    };
    auto mma_step = [&](const double* As, const double* Bp, int ks) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            a[i] = As[(ks + fk) * SY_LD + wy * 32 + i * 8 + fm];
            b[i] = Bp[(ks + fk) * SY_LD + wx * 32 + i * 8 + fm];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    };
#pragma unroll
    for (int c = 0; c < SY_STAGES - 1; ++c) issue(c);
    for (int c = 0; c < nchunk; ++c) {
        const int st = c % SY_STAGES;
        k2_mbar_wait(bar0 + 8u * st, (unsigned)((c / SY_STAGES) & 1));        // chunk c has landed
        const double* As = sy_smem + (size_t)st * 2 * SY_K * SY_LD;
        const double* Bp = diag ? As : As + SY_K * SY_LD;
        if (rhs_warp) {                        // S~[n][ti*128 + col] += W~[row][n] * W~[row][col] on the scalar FP64 pipe
            const double* rv = As + SY_K * SY_LD;
#pragma unroll 8
            for (int row = 0; row < SY_K; ++row) {
                const double g = rv[row];
#pragma unroll
                for (int k = 0; k < 4; ++k) racc[k] += g * As[row * SY_LD + lane + 32 * k];
            }
        }
        // each quarter of the chunk's MMAs is queued BEFORE a quarter of the copies of chunk c + STAGES - 1 is
        // issued, so the FP64-MMA pipe has work while this warp runs through its staging instructions
#pragma unroll
        for (int k = 0; k < SY_NQ; ++k) {
            if (warp_active) mma_step(As, Bp, 4 * k);
            if (k == SY_NQ - 1) {              // this warp is done with stage st
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar0 + 8u * (SY_STAGES + st)) : "memory");
            }
            issue_q(c + SY_STAGES - 1, k);
        }
    }
    if (rhs_warp) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int col = ti * SY_T + lane + 32 * k;
            if (col < n && racc[k] != 0.0) atomicAdd(Sfull + (int64_t)n * ldw + col, racc[k]);
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int row = ti * SY_T + wy * 32 + i * 8 + fm;
                const int col = tj * SY_T + wx * 32 + j * 8 + fk * 2 + e;
                const double v = acc[i][j][e];
                if (row < n && col <= row && v != 0.0) atomicAdd(Sfull + (int64_t)row * ldw + col, v);
            }
}

// rs_bounds (common.py:655-660, scipy trf_bounds): active-set treatment of the box 0 <= rho <= 1.
// A rho that sits on a bound while the descent direction -g pushes it outward is frozen for this
// linear solve (its row/column leave the system); free ones are stepped and projected afterwards.
__global__ void active_rho_kernel(const double* __restrict__ x, const double* __restrict__ bc, int nc, int Pc,
                                  int rs_bounds, int* __restrict__ frozen) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nc * Pc) return;
    int f = 0;
    if (rs_bounds && (i % Pc) == 2) {
        const int cam = i / Pc;
        const double rho = x[2 * nc + cam], b = bc[i];        // b = -gradient
        f = (rho <= 0.0 && b < 0.0) || (rho >= 1.0 && b > 0.0);
    }
    frozen[i] = f;
}

__global__ void freeze_cols_kernel(double* __restrict__ Ww, int64_t rows, int ldw, int nc, int Pc,
                                   const int* __restrict__ frozen) {
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    for (int cam = 0; cam < nc; ++cam)
        if (frozen[cam * Pc + 2]) Ww[row * ldw + cam * Pc + 2] = 0.0;
}

// S = blockdiag(A) [+ Ax] + lam * diag - S~ (lower triangle), rhs = bc - S~[ncP][:]
__global__ void form_schur_kernel(const double* __restrict__ A, const double* __restrict__ bc,
                                  const double* __restrict__ diag_c, const double* __restrict__ lam_p, int nc, int Pc, int ldw,
                                  const int* __restrict__ frozen, const double* __restrict__ Ax,
                                  double* __restrict__ Sfull, double* __restrict__ rhs) {
    const int ncP = nc * Pc;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)ncP * ncP) return;
    const int row = (int)(i / ncP), col = (int)(i - (int64_t)row * ncP);
    if (col > row) return;
    double v = -Sfull[(int64_t)row * ldw + col];
    const int cr = row / Pc, cc = col / Pc;
    if (cr == cc) v += A[((int64_t)cr * Pc + (row - cr * Pc)) * Pc + (col - cc * Pc)];
    else if (Ax) v += Ax[(int64_t)row * ncP + col];          // cross-camera entries (points mode, ba_points.cuh)
    if (row == col) {
        v += lam_p[0] * diag_c[row];
        rhs[row] = frozen[row] ? 0.0 : bc[row] - Sfull[(int64_t)ncP * ldw + row];
    }
    if (frozen[row] || frozen[col]) v = (row == col) ? 1.0 : 0.0;
    Sfull[(int64_t)row * ldw + col] = v;
}

// Dense Cholesky + solve of the reduced camera system (n = nc*Pc <= 1152): right-looking blocked
// factorisation, panel width 32.  Lower triangle, leading dimension lds >= n + 1, in place.
constexpr int CH_B = 32;
// Panel step of the right-looking Cholesky, two launches per 32-wide panel (round 1 used three plus a
// forward/backward solve on one warp: 55 dependent launches, 1.3 ms at n = 576, replicated on every rank):
//   chol_panel_kernel   EVERY CTA factors the 32 x 32 diagonal block itself in shared memory (redundant, but
//                       it removes a dependent launch; CTA 0 writes L_pp back) and then solves its 32 rows below
//                       the panel, X L_pp^T = S_ip;
//   chol_update_kernel  trailing update S_ij -= L_ip L_jp^T.
// The right-hand side rides along as ROW n of S: after the last panel it holds y = L^-1 rhs (the forward
// substitution), so only the backward substitution is left for chol_back_kernel.
__global__ void __launch_bounds__(256)
chol_panel_kernel(double* __restrict__ S, int lds, int n, int p0, int* __restrict__ fail_flag) {
    __shared__ double L[CH_B][CH_B + 1];
    const int nb = min(CH_B, n - p0), tid = threadIdx.x;
    for (int i = tid; i < CH_B * CH_B; i += 256) {
        const int r = i / CH_B, c = i - r * CH_B;
        L[r][c] = (r < nb && c <= r) ? S[(int64_t)(p0 + r) * lds + p0 + c] : (r == c ? 1.0 : 0.0);
    }
    __syncthreads();
    // 32 x 32 Cholesky on ONE warp, rows in registers (lane = row): pivot and multipliers travel by shuffles, no
    // barrier inside the 32 dependent steps (the 256-thread version with three __syncthreads per step took ~15 us)
    if (tid < 32) {
        double row[CH_B];
#pragma unroll
        for (int c = 0; c < CH_B; ++c) row[c] = L[tid][c];
        bool bad = false;
#pragma unroll
        for (int j = 0; j < CH_B; ++j) {
            double d = __shfl_sync(0xffffffffu, row[j], j);
            if (!(d > 0.0)) { bad = bad || j < nb; d = 1.0; }
            const double inv = 1.0 / sqrt(d);
            const double lij = row[j] * inv;               // lane i >= j: L[i][j]
            row[j] = lij;
#pragma unroll
            for (int k2 = j + 1; k2 < CH_B; ++k2) {
                const double lkj = __shfl_sync(0xffffffffu, lij, k2);
                row[k2] -= lij * lkj;                      // entry (i, k2), meaningful for k2 <= i
            }
        }
#pragma unroll
        for (int c = 0; c < CH_B; ++c) L[tid][c] = row[c];
        if (bad && tid == 0 && blockIdx.x == 0) atomicExch(fail_flag, 1);
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int i = tid; i < CH_B * CH_B; i += 256) {
            const int r = i / CH_B, c = i - r * CH_B;
            if (r < nb && c <= r) S[(int64_t)(p0 + r) * lds + p0 + c] = L[r][c];
        }
    // rows below the panel, the right-hand-side row n included: thread = row
    const int row = p0 + nb + blockIdx.x * CH_B + tid;
    if (tid >= CH_B || row > n) return;
    double x[CH_B];
    double* sr = S + (int64_t)row * lds + p0;
#pragma unroll
    for (int c = 0; c < CH_B; ++c) x[c] = c < nb ? sr[c] : 0.0;
#pragma unroll
    for (int c = 0; c < CH_B; ++c) {
        if (c < nb) {
            double v = x[c];
#pragma unroll
            for (int k = 0; k < c; ++k) v -= x[k] * L[c][k];
            x[c] = v / L[c][c];
        }
    }
#pragma unroll
    for (int c = 0; c < CH_B; ++c) if (c < nb) sr[c] = x[c];
}

// trailing update S_ij -= L_ip L_jp^T for 32x32 tiles i >= j below the panel (rows up to and including n)
__global__ void __launch_bounds__(CH_B * 8)
chol_update_kernel(double* __restrict__ S, int lds, int n, int p0) {
    __shared__ double Li[CH_B][CH_B + 1], Lj[CH_B][CH_B + 1];
    const int nb = min(CH_B, n - p0);
    int p = blockIdx.x, ti = 0;
    while (p >= ti + 1) { p -= ti + 1; ++ti; }
    const int tj = p;
    const int i0 = p0 + nb + ti * CH_B, j0 = p0 + nb + tj * CH_B;
    for (int i = threadIdx.x; i < CH_B * CH_B; i += blockDim.x) {
        const int r = i / CH_B, c = i - r * CH_B;
        Li[r][c] = (i0 + r <= n && c < nb) ? S[(int64_t)(i0 + r) * lds + p0 + c] : 0.0;
        Lj[r][c] = (j0 + r < n && c < nb) ? S[(int64_t)(j0 + r) * lds + p0 + c] : 0.0;
    }
    __syncthreads();
    const int c = threadIdx.x & 31, rq = threadIdx.x >> 5;      // 8 row groups of 4 rows
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = rq * 4 + k;
        const int gi = i0 + r, gj = j0 + c;
        if (gi <= n && gj < n && gj <= gi) {
            double acc = 0.0;
#pragma unroll
            for (int kk = 0; kk < CH_B; ++kk) acc += Li[r][kk] * Lj[c][kk];
            S[(int64_t)gi * lds + gj] -= acc;
        }
    }
}

// x = L^-T y with y = row n of S; blocked by 32 on one CTA, the diagonal block of every step staged in shared
// memory first (the 32 dependent steps then run on shared memory instead of 32 global round trips)
__global__ void __launch_bounds__(1024)
chol_back_kernel(const double* __restrict__ S, int lds, int n, double* __restrict__ xout) {
    __shared__ double y[1152];
    __shared__ double Ld[CH_B][CH_B + 1];
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n; i += nt) y[i] = S[(int64_t)n * lds + i];
    __syncthreads();
    for (int p0 = ((n - 1) / CH_B) * CH_B; p0 >= 0; p0 -= CH_B) {
        const int nb = min(CH_B, n - p0);
        {
            const int r = tid / CH_B, c = tid - r * CH_B;      // 1024 threads = one 32 x 32 block
            Ld[r][c] = (r < nb && c <= r) ? S[(int64_t)(p0 + r) * lds + p0 + c] : 0.0;
        }
        __syncthreads();
        if (tid < 32) {
            for (int j = nb - 1; j >= 0; --j) {
                const double xj = y[p0 + j] / Ld[j][j];
                __syncwarp();
                if (tid == 0) y[p0 + j] = xj;
                if (tid < j) y[p0 + tid] -= Ld[j][tid] * xj;
                __syncwarp();
            }
        }
        __syncthreads();
        for (int i = tid; i < p0; i += nt) {
            double acc = 0.0;
            for (int k = 0; k < nb; ++k) acc += S[(int64_t)(p0 + k) * lds + i] * y[p0 + k];
            y[i] -= acc;
        }
        __syncthreads();
    }
    for (int i = tid; i < n; i += nt) xout[i] = y[i];
}

inline void dense_chol_solve(mvus_ba_ctx* h, double* S, int lds, int n, double* rhs, double* xout, int* fail_flag) {
    cudaMemcpyAsync(S + (int64_t)n * lds, rhs, n * sizeof(double), cudaMemcpyDeviceToDevice, h->st);   // rhs = row n
    for (int p0 = 0; p0 < n; p0 += CH_B) {
        const int nb = std::min(CH_B, n - p0);
        const int rem = n + 1 - p0 - nb;                   // rows below the panel, rhs row included (>= 1)
        const int nblk = (rem + CH_B - 1) / CH_B;
        chol_panel_kernel<<<nblk, 256, 0, h->st>>>(S, lds, n, p0, fail_flag);
        h->launches++;
        if (n - p0 - nb > 0) {
            chol_update_kernel<<<nblk * (nblk + 1) / 2, CH_B * 8, 0, h->st>>>(S, lds, n, p0);
            h->launches++;
        }
    }
    chol_back_kernel<<<1, 1024, 0, h->st>>>(S, lds, n, xout);
    h->launches++;
}

// v[k][a] = W~[k][a][ncP] - W~[k][a][0:ncP] . dc     (one warp per row)
__global__ void wdc_kernel(const double* __restrict__ Ww, const double* __restrict__ dc, int64_t rows, int ldw,
                           double* __restrict__ v) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const double* w = Ww + row * ldw;
    double acc = 0.0;
    for (int c = lane; c < ldw - 1; c += 32) acc += w[c] * dc[c];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) v[row] = w[ldw - 1] - acc;
}

// Back substitution of one level: ds_k = L_k^-T ( v_k - ZL_k ds_{k-s} - ZR_k ds_{k+s} ).
// grid = eliminated blocks of this level, block = 32.
__global__ void bcr_back_kernel(int64_t nb, int q, int64_t s, int root, const double* __restrict__ Dw,
                                const double* __restrict__ Ew, const double* __restrict__ ZL,
                                double* __restrict__ ds) {
    const int64_t k = root ? 0 : ((int64_t)blockIdx.x * 2 + 1) * s;
    if (k >= nb) return;
    const int qq = q * q, lane = threadIdx.x;
    const int a = lane < q ? lane : 0;
    double v = ds[k * q + a];
    if (!root) {
        const double* zl = ZL + k * qq + a * q;
        const double* dl = ds + (k - s) * q;
        for (int b = 0; b < q; ++b) v -= zl[b] * dl[b];
        if (k + s < nb) {
            const double* zr = Ew + k * qq + a * q;
            const double* dr = ds + (k + s) * q;
            for (int b = 0; b < q; ++b) v -= zr[b] * dr[b];
        }
    }
    // x = L^-T v on one warp: lane a keeps column a of L (L[i][a], i >= a); x_i leaves lane i by shuffle and the
    // lanes below it update their v (the version with one thread walking the triangle took ~13 us per level)
    const double* L = Dw + k * qq;
    double lcol[QMAX];
#pragma unroll
    for (int i = 0; i < QMAX; ++i) lcol[i] = i < q ? L[i * q + a] : 0.0;
    double x = 0.0;
#pragma unroll
    for (int i = QMAX - 1; i >= 0; --i) {
        if (i < q) {
            const double xi = __shfl_sync(0xffffffffu, v, i) / __shfl_sync(0xffffffffu, lcol[i], i);
            if (a == i) x = xi;
            if (a < i) v -= lcol[i] * xi;
        }
    }
    if (lane < q) ds[k * q + lane] = x;
}

// ------------------------------------------------------------------------------------------
// Step bookkeeping.
// sums[0] = sum diag*delta^2, sums[1] = b . delta, sums[2] = |delta|^2, sums[3] = |x|^2,
// sums[4] = max |g| (as ordered-int atomicMax on the bits of a non-negative double)
__global__ void step_dots_kernel(const double* __restrict__ dc, const double* __restrict__ ds,
                                 const double* __restrict__ diag_c, const double* __restrict__ diag_s,
                                 const double* __restrict__ bc, const double* __restrict__ bs, int ncP,
                                 int64_t n_ctrl3, const double* __restrict__ x, int64_t n,
                                 double* __restrict__ sums) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, gm = 0;
    if (i < ncP) {
        const double d = dc[i];
        s0 += diag_c[i] * d * d; s1 += bc[i] * d; s2 += d * d; gm = fabs(bc[i]);
    }
    if (i < n_ctrl3) {
        const double d = ds[i], b = bs[i];
        s0 += diag_s[i] * d * d; s1 += b * d; s2 += d * d; gm = fmax(gm, fabs(b));
    }
    if (i < n) s3 = x[i] * x[i];
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
        gm = fmax(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sums + 0, s0); atomicAdd(sums + 1, s1); atomicAdd(sums + 2, s2); atomicAdd(sums + 3, s3);
        atomicMax(reinterpret_cast<unsigned long long*>(sums + 4), (unsigned long long)__double_as_longlong(gm));
    }
}

// x_trial = x + delta in the reference layout; rho clamped to [0,1] under rs_bounds.
__global__ void apply_step_kernel(const double* __restrict__ x, const double* __restrict__ dc,
                                  const double* __restrict__ ds, int nc, int C, int Pc, int64_t n_other,
                                  SplineView sp, int64_t n_ctrl, int rs_bounds, double* __restrict__ xt) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_other) {
        int cam, p;
        if (i < 3 * (int64_t)nc) { p = (int)(i / nc); cam = (int)(i - (int64_t)p * nc); }
        else { const int64_t k = i - 3 * (int64_t)nc; cam = (int)(k / C); p = 3 + (int)(k - (int64_t)cam * C); }
        double v = x[i] + dc[cam * Pc + p];
        if (rs_bounds && p == 2) v = fmin(fmax(v, 0.0), 1.0);
        xt[i] = v;
    }
    if (i < n_ctrl) {
        int s = 0;
        while (s + 1 < sp.S && i >= sp.ctrl_off[s + 1]) ++s;
        const int64_t l = i - sp.ctrl_off[s];
        const int nco = sp.ncoef[s];
        for (int ax = 0; ax < 3; ++ax) {
            const int64_t xi = sp.xoff[s] + (int64_t)ax * nco + l;
            xt[xi] = x[xi] + ds[i * 3 + ax];
        }
    }
}

// gradient in the reference layout: g = -b
__global__ void gradient_kernel(const double* __restrict__ bc, const double* __restrict__ bs, int nc, int C,
                                int Pc, int64_t n_other, SplineView sp, int64_t n_ctrl,
                                double* __restrict__ g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_other) {
        int cam, p;
        if (i < 3 * (int64_t)nc) { p = (int)(i / nc); cam = (int)(i - (int64_t)p * nc); }
        else { const int64_t k = i - 3 * (int64_t)nc; cam = (int)(k / C); p = 3 + (int)(k - (int64_t)cam * C); }
        g[i] = -bc[cam * Pc + p];
    }
    if (i < n_ctrl) {
        int s = 0;
        while (s + 1 < sp.S && i >= sp.ctrl_off[s + 1]) ++s;
        const int64_t l = i - sp.ctrl_off[s];
        const int nco = sp.ncoef[s];
        for (int ax = 0; ax < 3; ++ax)
            g[sp.xoff[s] + (int64_t)ax * nco + l] = -bs[i * 3 + ax];
    }
}

// ------------------------------------------------------------------------------------------
// Host orchestration
// Super-block structure of the spline block for a given world size: bw control points per super-block
// (reprojection rows touch 4 consecutive control points -> bw >= 3; a motion row touches up to `spread`
// consecutive ones -> bw >= spread - 1), nb super-blocks, chunk size Bc of the sharded solve.
inline int solver_dims(mvus_ba_ctx* h, int world, int* bw_out, int64_t* nb_out, int64_t* Bc_out) {
    const int spread = h->M > 0 ? h->motion_spread : 4;      // computed once in mvus_ba_set_splines (ba_tables.hpp)
    if (spread > 7) return fail(h, MVUS_ERR_UNSUPPORTED,
                                "a motion-prior row touches more than 7 consecutive control points");
    const int bw = std::max(3, spread - 1);
    int64_t nb = (h->n_ctrl + bw - 1) / bw, Bc = 1;
    if (world > 1) {
        // chunk size for the sharded solve: power of two, ~8 chunks per rank, >= 1
        while (Bc * 2 * 8 * world <= nb) Bc *= 2;
        nb = (nb + Bc - 1) / Bc * Bc;                       // padded with decoupled identity blocks
    }
    *bw_out = bw; *nb_out = nb; *Bc_out = Bc;
    return MVUS_OK;
}

inline int solver_alloc(mvus_ba_ctx* h) {
    int rc = solver_dims(h, h->world, &h->bw, &h->nb, &h->Bc);
    if (rc) return rc;
    h->q = 3 * h->bw;
    h->ncP = h->nc * h->Pc;
    h->ldw = h->ncP + 1;
    if (h->ncP > 1152) return fail(h, MVUS_ERR_UNSUPPORTED, "more than 1152 camera unknowns");
    const size_t qq = (size_t)h->q * h->q;
    MV_CUDA(h, h->A.alloc((size_t)h->nc * h->Pc * h->Pc + h->ncP));      // A then bc
    const size_t nba = (size_t)h->nb + 1;                      // +1: ghost block of the sharded solve
    if ((int64_t)nba * h->q >= ((int64_t)1 << 26))             // K2's run table packs (row << 5 | local column)
        return fail(h, MVUS_ERR_UNSUPPORTED, "more than 2^26 spline unknowns per handle");
    MV_CUDA(h, h->D.alloc(nba * qq));
    MV_CUDA(h, h->E.alloc(nba * qq));
    MV_CUDA(h, h->W.alloc((nba * h->q + 2 * K2_GUARD) * h->ldw));       // K2_GUARD rows in front of W~ (K2's window of
    h->w_guard = (size_t)K2_GUARD * h->ldw;                             //   block 0 starts at control point -3): see Wp()
    MV_CUDA(h, h->Dw.alloc(nba * qq));
    MV_CUDA(h, h->Ew.alloc(nba * qq));
    MV_CUDA(h, h->Ww.alloc(nba * h->q * h->ldw));
    MV_CUDA(h, h->ZL.alloc(nba * qq));
    // chunk pre-reduction (ba_chunk.cuh): long systems only -- below a few hundred super-blocks the cyclic
    // reduction's handful of tiny launches is cheaper than a sequential sweep over a chunk
    int Lc = h->Lc_req > 0 ? h->Lc_req : (h->nb >= 4096 ? 32 : h->nb >= 1024 ? 16 : h->nb >= 256 ? 8 : 1);
    if (h->world > 1) {                                        // rank ranges are multiples of Bc (a power of two)
        int p2 = 1;
        while (p2 * 2 <= Lc && p2 * 2 <= h->Bc) p2 *= 2;
        Lc = p2;
    }
    if (Lc >= h->nb) Lc = 1;
    h->Lc = Lc;
    h->nh = (h->nb + Lc - 1) / Lc;
    if (Lc > 1) {
        const size_t nha = (size_t)h->nh + 1;                  // +1: ghost head of the sharded solve
        MV_CUDA(h, h->Dh.alloc(nha * qq));
        MV_CUDA(h, h->Eh.alloc(nha * qq));
        MV_CUDA(h, h->ZLh.alloc(nha * qq));
        MV_CUDA(h, h->DhR.alloc(nha * qq));
        MV_CUDA(h, h->Wh.alloc(nha * h->q * h->ldw));
        MV_CUDA(h, h->Gh.alloc(nha * h->q * h->ldw));
        MV_CUDA(h, h->dsh.alloc(nha * h->q));
        MV_CUDA(h, h->Linv.alloc(nba * h->q));
    }
    if (h->world > 1) {
        const size_t nch = (size_t)(h->nb / h->Bc);
        MV_CUDA(h, h->Dt.alloc(nch * (2 * qq + (size_t)h->q * h->ldw)));
        MV_CUDA(h, h->ZLt.alloc(nch * qq));
        MV_CUDA(h, h->dst.alloc(nch * h->q));
    }
    MV_CUDA(h, h->Sd.alloc((size_t)h->ldw * h->ldw + h->ldw));           // S~ then rhs
    MV_CUDA(h, h->dlt_c.alloc(h->ncP));
    MV_CUDA(h, h->dlt_s.alloc(((size_t)h->nb + 1) * h->q));
    MV_CUDA(h, h->diag_c.alloc(h->ncP));
    MV_CUDA(h, h->frozen.alloc(h->ncP));
    MV_CUDA(h, cudaMemsetAsync(h->frozen.p, 0, h->ncP * sizeof(int), h->st));
    MV_CUDA(h, h->diag_s.alloc((size_t)h->nb * h->q));
    MV_CUDA(h, h->bs.alloc(((size_t)h->nb + 1) * h->q));
    MV_CUDA(h, h->gvec.alloc(h->n));
    MV_CUDA(h, h->xs.alloc(16));
    return MVUS_OK;
}

// K2 + K2m at the current J / r.  Fills A, bc, D, E, W (W's last column = -g_s) and the diagonals.
inline int accumulate(mvus_ba_ctx* h) {
    const size_t qq = (size_t)h->q * h->q;
    double* bc = h->A.p + (size_t)h->nc * h->Pc * h->Pc;
    MV_CUDA(h, cudaMemsetAsync(h->A.p, 0, h->A.bytes(), h->st));
    MV_CUDA(h, cudaMemsetAsync(h->D.p, 0, h->nb * qq * sizeof(double), h->st));
    MV_CUDA(h, cudaMemsetAsync(h->E.p, 0, h->nb * qq * sizeof(double), h->st));
    MV_CUDA(h, cudaMemsetAsync(h->Wp(), 0, (size_t)h->nb * h->q * h->ldw * sizeof(double), h->st));
    if (h->n_chunks > 0) {
        // chunk order by time (see chunk_key_kernel); spans move little between evaluations, but the sort is cheap
        const int nch = h->n_chunks;
        MV_CUDA(h, h->chunk_key.alloc(nch));
        MV_CUDA(h, h->chunk_key2.alloc(nch));
        MV_CUDA(h, h->chunk_id.alloc(nch));
        MV_CUDA(h, h->chunk_perm.alloc(nch));
        MV_CUDA(h, h->k2_queue.alloc(4));
        // (the order only serves L2 locality of the RED targets, and the spans move little between the evaluations
        //  of one solve: sorted at the first accumulation after the inputs / the start point changed, reused afterwards)
        if (!h->chunk_sorted) {
            size_t tb = 0;
            MV_CUDA(h, cub::DeviceRadixSort::SortPairs(nullptr, tb, h->chunk_key.p, h->chunk_key2.p, h->chunk_id.p,
                                                       h->chunk_perm.p, nch, 0, 32, h->st));
            MV_CUDA(h, h->sort_tmp.alloc(tb));
            const int blk_d = jblk_doubles(h->P);
            chunk_key_kernel<<<(nch + 255) / 256, 256, 0, h->st>>>(h->J.p, blk_d, blk_d - 16, h->chunk_tile0.p, nch,
                                                                  h->chunk_key.p, h->chunk_id.p);
            MV_CUDA(h, cub::DeviceRadixSort::SortPairs(h->sort_tmp.p, tb, h->chunk_key.p, h->chunk_key2.p, h->chunk_id.p,
                                                       h->chunk_perm.p, nch, 0, 32, h->st));
            h->chunk_sorted = true;
            h->launches += 2;
        }
        {   // [0] work queue, [1] smallest, [2] largest block-of-four-spans any warp saw (the touched rows)
            const int init[3] = {0, 0x7fffffff, -1};
            MV_CUDA(h, cudaMemcpyAsync(h->k2_queue.p, init, sizeof(init), cudaMemcpyHostToDevice, h->st));
        }
        h->launches += 1;
        const int64_t n_rows = (int64_t)(h->nb + 1) * h->q;
        if ((n_rows + 2 * K2_GUARD) * (int64_t)h->ldw >= ((int64_t)1 << 32))        // K2 flushes with 32-bit offsets
            return fail(h, MVUS_ERR_UNSUPPORTED, "W~ larger than 2^32 entries per handle");
        const size_t hb_n = (size_t)(n_rows + 2 * K2_GUARD) * K2_BAND;
        MV_CUDA(h, h->Hb.alloc(hb_n));
        MV_CUDA(h, cudaMemsetAsync(h->Hb.p, 0, hb_n * sizeof(double), h->st));
        const int grid = h->sm_count;
#define MV_K2(PP)                                                                                            \
    do {                                                                                                     \
        MV_CUDA(h, cudaFuncSetAttribute(accumulate_kernel<PP>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                        (int)K2Cfg<PP>::SMEM));                                              \
        accumulate_kernel<PP><<<grid, K2Cfg<PP>::THREADS, K2Cfg<PP>::SMEM, h->st>>>(                         \
            h->J.p, h->chunk_perm.p, h->chunk_tile0.p, h->chunk_nt.p, nch, h->k2_queue.p, h->tile_cam.p,     \
            h->tile_cnt.p, h->ldw, h->A.p, bc, h->Hb.p, h->W.p);                                                      \
    } while (0)
        if (h->P == 21) MV_K2(21); else MV_K2(30);
#undef MV_K2
        band_to_blocks_kernel<<<(int)((n_rows * (K2_HALF + 1) + 255) / 256), 256, 0, h->st>>>(
            h->Hb.p + (size_t)K2_GUARD * K2_BAND, n_rows, h->q, h->D.p, h->E.p);
        h->launches += 2;
    }
    if (h->M > 0) {      // parameter-only rows: every rank takes the rows whose first control point it owns
        int64_t lo = 0, hi = h->nb;
        if (h->world > 1) owner_range(h, h->rank, &lo, &hi);
        const int64_t c_lo = h->rank == 0 ? -1 : lo * h->bw;
        const int64_t c_hi = h->rank == h->world - 1 ? ((int64_t)1 << 40) : hi * h->bw;
        accumulate_motion_kernel<<<(int)((h->M + 127) / 128), 128, 0, h->st>>>(
            h->r.p + 2 * h->N, h->mbase.p, h->mJ.p, h->M, h->bw, h->ldw, c_lo, c_hi, h->D.p, h->E.p, h->Wp());
        h->launches++;
    }
    MV_CUDA(h, cudaGetLastError());
    return MVUS_OK;
}

inline void owner_range(const mvus_ba_ctx* h, int r, int64_t* lo, int64_t* hi);
inline int reduce_normal_equations(mvus_ba_ctx* h, bool full);   // ba_nccl.cuh
inline void owner_range(const mvus_ba_ctx* h, int r, int64_t* lo, int64_t* hi);
inline int nccl_bcast0(mvus_ba_ctx* h, double* buf, size_t count);
inline int nccl_max_flag(mvus_ba_ctx* h, int* flag);
inline int nccl_sum(mvus_ba_ctx* h, double* buf, size_t count);

inline int compute_diag(mvus_ba_ctx* h, bool full_everywhere = false) {
    const int64_t nbq = h->nb * h->q;
    const int64_t cnt = std::max<int64_t>(nbq, h->ncP);
    int64_t lo = 0, hi = h->nb;
    if (h->world > 1 && !full_everywhere) owner_range(h, h->rank, &lo, &hi);
    diag_kernel<<<(int)((cnt + 255) / 256), 256, 0, h->st>>>(h->A.p, h->D.p, h->Wp(), h->nc, h->Pc, nbq, h->q,
                                                            h->ldw, 3 * h->n_ctrl, lo * h->q, hi * h->q,
                                                            h->diag_c.p, h->diag_s.p, h->bs.p);
    h->launches++;
    if (h->world > 1 && !full_everywhere) {
        int e = nccl_sum(h, h->diag_s.p, (size_t)nbq);
        if (!e) e = nccl_sum(h, h->bs.p, (size_t)nbq);
        if (e) return e;
    }
    const int64_t n3 = 3 * h->n_ctrl;
    if (n3 > 0) {
        MV_CUDA(h, cudaMemsetAsync(h->xs.p + 8, 0, sizeof(double), h->st));
        sum_kernel<<<(int)((n3 + 255) / 256), 256, 0, h->st>>>(h->diag_s.p, n3, h->xs.p + 8);
        if (h->world > 1) { const int e = nccl_bcast0(h, h->xs.p + 8, 1); if (e) return e; }   // same floor on every rank
        floor_kernel<<<(int)((n3 + 255) / 256), 256, 0, h->st>>>(h->diag_s.p, n3, h->xs.p + 8, DIAG_FLOOR_FRAC);
        h->launches += 2;
    }
    MV_CUDA(h, cudaGetLastError());
    return MVUS_OK;
}

struct BcrView {          // a block-tridiagonal system (possibly a sub-range of the handle's arrays)
    double* Dw; double* Ew; double* Ww; double* ZL; double* ds; int64_t nb;
    const double* Worig = nullptr;   // if set: Ww has NOT been initialised, a block's first touch reads from here
};

inline void launch_level(mvus_ba_ctx* h, const BcrView& v, int grid, int64_t s, int64_t sp, int root, int* fail_flag,
                         int odd_only = 0, const double* wsrc = nullptr) {
    if (grid <= 0) return;
    const int nthr = h->ldw > 192 ? 256 : 128;      // one W~ column per thread and pass
#define MV_LVL(QQ) bcr_level_kernel<QQ><<<grid, nthr, 0, h->st>>>(v.nb, h->ldw, s, sp, root, odd_only, wsrc, v.Dw, v.Ew, v.Ww, v.ZL, fail_flag)
    switch (h->q) {
        case 9: MV_LVL(9); break;
        case 12: MV_LVL(12); break;
        case 15: MV_LVL(15); break;
        default: MV_LVL(18); break;
    }
#undef MV_LVL
    h->launches++;
}

// Elimination levels with strides 1, 2, .. < s_end (s_end = nb: all levels), optional root pass.
inline std::vector<int64_t> bcr_eliminate(mvus_ba_ctx* h, const BcrView& v, int64_t s_end, bool with_root, int* fail_flag) {
    std::vector<int64_t> levels;
    for (int64_t s = 1; s < s_end && s < v.nb; s <<= 1) levels.push_back(s);
    for (size_t lv = 0; lv < levels.size(); ++lv) {
        const int64_t s = levels[lv], sp = lv ? levels[lv - 1] : 0;
        // level 0: only the odd blocks work (no pending updates exist yet); blocks first touched at
        // level 0 (odd) or 1 (even) read their W~ rows from the original array when no copy was made
        if (lv == 0) launch_level(h, v, (int)(v.nb / 2), s, sp, 0, fail_flag, 1, v.Worig);
        else launch_level(h, v, (int)((v.nb + s - 1) / s), s, sp, 0, fail_flag, 0, lv == 1 ? v.Worig : nullptr);
    }
    if (with_root)
        launch_level(h, v, 1, levels.empty() ? 1 : levels.back() * 2, levels.empty() ? 0 : levels.back(), 1, fail_flag,
                     0, levels.size() <= 1 ? v.Worig : nullptr);
    return levels;
}

inline void bcr_back(mvus_ba_ctx* h, const BcrView& v, const std::vector<int64_t>& levels, bool with_root) {
    if (with_root) { bcr_back_kernel<<<1, 32, 0, h->st>>>(v.nb, h->q, 0, 1, v.Dw, v.Ew, v.ZL, v.ds); h->launches++; }
    for (int lv = (int)levels.size() - 1; lv >= 0; --lv) {
        const int64_t s = levels[lv];
        const int64_t nel = (v.nb / s + 1) / 2;     // odd multiples of s below nb (kernel guards the rest)
        if (nel <= 0) continue;
        bcr_back_kernel<<<(int)nel, 32, 0, h->st>>>(v.nb, h->q, s, 0, v.Dw, v.Ew, v.ZL, v.ds);
        h->launches++;
    }
}

inline void launch_syrk(mvus_ba_ctx* h, const double* Wrows, int64_t R, double* Sfull) {
    if (R <= 0) return;
    const int ldw = h->ldw, n = h->ncP;
    const int nts = (n + SY_T - 1) / SY_T, npairs = nts * (nts + 1) / 2;
    // one 512-thread CTA per SM; CTA durations differ by up to 16:3 active warp tiles; slabs are multiples of the K chunk
    // CTAs are dispatched in order as SMs free up, so the launch ends with a tail of up to one CTA duration:
    // the first 3/4 of the rows go in big slabs (~6 waves), the last quarter in slabs a quarter as long
    int64_t nbig = std::max<int64_t>(1, (6 * h->sm_count + npairs - 1) / npairs);
    int slab = (int)std::max<int64_t>(256, ((R * 3 / 4 + nbig - 1) / nbig + SY_K - 1) / SY_K * SY_K);
    nbig = std::min<int64_t>(nbig, R / slab);
    const int small = std::max(64, slab / 4 / SY_K * SY_K);
    const int64_t rest = R - nbig * slab;
    const int64_t nsmall = (rest + small - 1) / small;
    dim3 g(npairs, (unsigned)(nbig + nsmall));
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SY_SMEM); attr_set = true; }
    syrk_kernel<<<g, 512, SY_SMEM, h->st>>>(Wrows, R, ldw, n, slab, (int)nbig, small, Sfull);
    h->launches += 1;
}

// copy the chunk-head blocks of this rank (and its ghost) into the compact top-level arrays
__global__ void top_gather_kernel(const double* __restrict__ Dw, const double* __restrict__ Ew,
                                  const double* __restrict__ Ww, int q, int ldw, int64_t Bc, int64_t c0,
                                  int64_t c1, int64_t nchunks, double* __restrict__ Dt,
                                  double* __restrict__ Et, double* __restrict__ Wt) {
    const int64_t c = c0 + blockIdx.x;                 // blockIdx.x in [0, c1 - c0]: last one = ghost
    if (c >= nchunks) return;
    const bool ghost = (c == c1);
    const int64_t k = c * Bc;
    const int qq = q * q;
    for (int i = threadIdx.x; i < qq; i += blockDim.x) {
        Dt[c * qq + i] = Dw[k * qq + i];
        if (!ghost) Et[c * qq + i] = Ew[k * qq + i];
    }
    const int64_t wn = (int64_t)q * ldw;
    for (int64_t i = threadIdx.x; i < wn; i += blockDim.x) Wt[c * wn + i] = Ww[k * wn + i];
}

__global__ void top_scatter_kernel(const double* __restrict__ dst, int q, int64_t Bc, int64_t c0, int64_t c1,
                                   int64_t nchunks, double* __restrict__ ds) {
    const int64_t c = c0 + blockIdx.x;                 // includes the ghost c1
    const int a = threadIdx.x;
    if (a >= q) return;
    ds[c * Bc * q + a] = c < nchunks ? dst[c * q + a] : 0.0;
}

__global__ void zero_outside_kernel(double* __restrict__ v, int64_t n, int64_t lo, int64_t hi) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && (i < lo || i >= hi)) v[i] = 0.0;
}

// Chunk pre-reduction of the super-blocks [lo, hi) (multiples of Lc, or hi = nb): L/ZR/ZH of the eliminated
// blocks into Dw/Ew/ZL, their W~ rows into Ww, the head system of chunks [lo/Lc, ceil(hi/Lc)) into Dh/Eh/Wh.
// The head after the range (a ghost when another rank owns it) receives the last chunk's contribution.
inline void prereduce_range(mvus_ba_ctx* h, int64_t lo, int64_t hi, const double* wsrc, int* fail_flag) {
    const int Lc = h->Lc, q = h->q, ldw = h->ldw;
    const size_t qq = (size_t)q * q;
    const int64_t wn = (int64_t)q * ldw;
    const int64_t c0 = lo / Lc, c1 = (hi + Lc - 1) / Lc;
    if (c1 <= c0) return;
    const bool has_next = hi < h->nb || h->world > 1;           // a head (or ghost slot) exists after the range
    // the first head has no chunk on its left inside the range; the ghost starts from zero
    cudaMemsetAsync(h->DhR.p + c0 * qq, 0, qq * sizeof(double), h->st);
    cudaMemsetAsync(h->Gh.p + c0 * wn, 0, wn * sizeof(double), h->st);
    if (has_next) {
        cudaMemsetAsync(h->Dh.p + c1 * qq, 0, qq * sizeof(double), h->st);
        cudaMemsetAsync(h->Eh.p + c1 * qq, 0, qq * sizeof(double), h->st);
        cudaMemsetAsync(h->Wh.p + c1 * wn, 0, wn * sizeof(double), h->st);
        if (hi >= h->nb) {                                      // last rank: nothing flows into the ghost
            cudaMemsetAsync(h->DhR.p + c1 * qq, 0, qq * sizeof(double), h->st);
            cudaMemsetAsync(h->Gh.p + c1 * wn, 0, wn * sizeof(double), h->st);
        }
    }
    const int nchunk = (int)(c1 - c0);
    // the kernels see the blocks [0, nbv): a block couples to its right neighbour iff that neighbour exists
    const int64_t nbv = (hi < h->nb) ? hi + 1 : h->nb;
    const dim3 gw(nchunk, (ldw + CW_T - 1) / CW_T);
#define MV_CHUNK(QQ, FACTOR)                                                                                       \
    FACTOR<QQ><<<nchunk, 32, 0, h->st>>>(nbv, Lc, c0, h->Dw.p, h->Ew.p, h->ZL.p, h->Linv.p, h->Dh.p, h->Eh.p,        \
                                         h->DhR.p, fail_flag);                                                    \
    chunk_w_kernel<QQ><<<gw, CW_T, 0, h->st>>>(nbv, Lc, c0, ldw, wsrc, h->Dw.p, h->Ew.p, h->ZL.p, h->Linv.p, h->Ww.p, h->Wh.p, h->Gh.p)
    switch (q) {
        case 9: MV_CHUNK(9, chunk_factor_fast_kernel); break;
        case 12: MV_CHUNK(12, chunk_factor_fast_kernel); break;
        case 15: MV_CHUNK(15, chunk_factor_fast_kernel); break;
        default: MV_CHUNK(18, chunk_factor_kernel); break;
    }
#undef MV_CHUNK
    head_fix_kernel<<<(int)(c1 - c0 + (has_next ? 1 : 0)), 256, 0, h->st>>>(c0, q, ldw, h->Dh.p, h->DhR.p, h->Wh.p, h->Gh.p);
    h->launches += 3;
}

inline void chunk_back_range(mvus_ba_ctx* h, int64_t lo, int64_t hi) {
    const int Lc = h->Lc;
    const int64_t c0 = lo / Lc, c1 = (hi + Lc - 1) / Lc;
    if (c1 <= c0) return;
    const int64_t nbv = (hi < h->nb) ? hi + 1 : h->nb;
#define MV_CB(QQ) chunk_back_kernel<QQ><<<(int)(c1 - c0), 32, 0, h->st>>>(nbv, Lc, c0, h->Dw.p, h->Ew.p, h->ZL.p, h->dsh.p, h->dlt_s.p)
    switch (h->q) {
        case 9: MV_CB(9); break;
        case 12: MV_CB(12); break;
        case 15: MV_CB(15); break;
        default: MV_CB(18); break;
    }
#undef MV_CB
    h->launches++;
}

// Solve the damped system for the current normal equations; delta -> dlt_c / dlt_s.
// *ok = 0 if a Cholesky pivot was not positive.
inline int solve_damped(mvus_ba_ctx* h, double lam, int* ok) {
    const int q = h->q, ldw = h->ldw, Lc = h->Lc;
    const int64_t nb = h->nb, nbq = nb * q;
    const size_t qq = (size_t)q * q;
    const int64_t wn = (int64_t)q * ldw;
    double* bc = h->A.p + (size_t)h->nc * h->Pc * h->Pc;
    double* rhs = h->Sd.p + (size_t)ldw * ldw;
    int* fail_flag = h->flag.p + 1;
    bool timed = false;
    const bool pre = Lc > 1;
    // the system the cyclic reduction works on: all super-blocks, or the chunk heads after the pre-reduction
    double* sD = pre ? h->Dh.p : h->Dw.p;
    double* sE = pre ? h->Eh.p : h->Ew.p;
    double* sW = pre ? h->Wh.p : h->Ww.p;
    double* sZ = pre ? h->ZLh.p : h->ZL.p;
    double* sd = pre ? h->dsh.p : h->dlt_s.p;
    // lambda travels through device memory (xs[12]) so that the launch sequence below does not depend on it
    const double* lam_p = h->xs.p + 12;
    h->h_pin[16] = lam;
    MV_CUDA(h, cudaMemcpyAsync(h->xs.p + 12, h->h_pin + 16, sizeof(double), cudaMemcpyHostToDevice, h->st));
    if (h->world <= 1) {
        // ---------------- single GPU: (chunk pre-reduction +) full cyclic reduction + root ----------------
        // The launch sequence is the same for every solve of a handle.  For SMALL systems, where the ~25-60
        // launches cost more than the kernels (config 5: 370 launches per problem, throughput bound by the launch
        // rate; config 2), it is captured into a CUDA graph at the handle's second solve and replayed afterwards.
        auto single_body = [&](bool timing) -> int {
            MV_CUDA(h, cudaMemsetAsync(fail_flag, 0, sizeof(int), h->st));
            MV_CUDA(h, cudaMemsetAsync(h->Sd.p, 0, h->Sd.bytes(), h->st));
            damp_copy_kernel<<<(int)((nbq * q + 255) / 256), 256, 0, h->st>>>(h->D.p, h->diag_s.p, lam_p, nbq, q,
                                                                              3 * h->n_ctrl, h->Dw.p);
            MV_CUDA(h, cudaMemcpyAsync(h->Ew.p, h->E.p, nb * qq * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
            h->launches += 1;
            const double* wsrc = h->Wp();          // first touch of every block reads W~ directly: no 2|W~| copy pass
            if (h->desc.rs_bounds) {               // frozen columns must be zeroed in a private copy
                MV_CUDA(h, cudaMemcpyAsync(h->Ww.p, h->Wp(), (size_t)nbq * ldw * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
                freeze_cols_kernel<<<(int)((nbq + 255) / 256), 256, 0, h->st>>>(h->Ww.p, nbq, ldw, h->nc, h->Pc, h->frozen.p);
                h->launches++;
                wsrc = h->Ww.p;
            }
            if (timing) cudaEventRecord(h->evs[0], h->st);
            BcrView v{sD, sE, sW, sZ, sd, pre ? h->nh : nb};
            if (pre) prereduce_range(h, 0, nb, wsrc, fail_flag);
            else if (!h->desc.rs_bounds) v.Worig = wsrc;
            std::vector<int64_t> levels = bcr_eliminate(h, v, v.nb, true, fail_flag);
            if (pre) {                             // the heads' rows join the eliminated ones for the SYRK
                head_rows_kernel<<<(int)h->nh, 256, 0, h->st>>>(0, Lc, q, ldw, h->Wh.p, h->Ww.p);
                h->launches++;
            }
            if (timing) cudaEventRecord(h->evs[1], h->st);
            launch_syrk(h, h->Ww.p, nbq, h->Sd.p);
            if (timing) { cudaEventRecord(h->evs[2], h->st); timed = true; }
            form_schur_kernel<<<(int)(((int64_t)h->ncP * h->ncP + 255) / 256), 256, 0, h->st>>>(
                h->A.p, bc, h->diag_c.p, lam_p, h->nc, h->Pc, ldw, h->frozen.p, h->Ax.p, h->Sd.p, rhs);
            h->launches++;
            dense_chol_solve(h, h->Sd.p, ldw, h->ncP, rhs, h->dlt_c.p, fail_flag);
            wdc_kernel<<<(int)((nbq * 32 + 255) / 256), 256, 0, h->st>>>(h->Ww.p, h->dlt_c.p, nbq, ldw, h->dlt_s.p);
            h->launches++;
            if (pre) {
                wdc_kernel<<<(int)((h->nh * q * 32 + 255) / 256), 256, 0, h->st>>>(h->Wh.p, h->dlt_c.p, h->nh * q, ldw, h->dsh.p);
                h->launches++;
            }
            bcr_back(h, v, levels, true);
            if (pre) chunk_back_range(h, 0, nb);
            MV_CUDA(h, cudaGetLastError());
            return MVUS_OK;
        };
        const bool small = (int64_t)nbq * ldw <= ((int64_t)8 << 20);          // W~ up to 64 MB
        if (!small || h->graph_state < 0) {
            const int rc = single_body(true);
            if (rc) return rc;
        } else if (h->graph_state == 0) {                 // first solve: direct (function attributes, allocations)
            const int rc = single_body(false);
            if (rc) return rc;
            h->graph_state = 1;
        } else {
            if (h->graph_state == 1) {
                cudaGraph_t g = nullptr;
                cudaError_t ce = cudaStreamBeginCapture(h->st, cudaStreamCaptureModeThreadLocal);
                int rc = MVUS_OK;
                if (ce == cudaSuccess) {
                    const int l0 = h->launches;
                    rc = single_body(false);
                    h->graph_launches = h->launches - l0;
                    ce = cudaStreamEndCapture(h->st, &g);
                }
                if (rc == MVUS_OK && ce == cudaSuccess && g) ce = cudaGraphInstantiate(&h->solve_graph, g, 0);
                if (g) cudaGraphDestroy(g);
                if (rc != MVUS_OK || ce != cudaSuccess || !h->solve_graph) {
                    cudaGetLastError();
                    h->solve_graph = nullptr;
                    h->graph_state = -1;                  // capture not possible here: direct launches from now on
                    const int rc2 = single_body(true);
                    if (rc2) return rc2;
                } else {
                    h->graph_state = 2;
                    h->launches -= h->graph_launches;     // (counted again by the replay below)
                }
            }
            if (h->graph_state == 2) {
                MV_CUDA(h, cudaGraphLaunch(h->solve_graph, h->st));
                h->launches += h->graph_launches;
            }
        }
    } else {
        // ---------------- sharded solve ----------------
        MV_CUDA(h, cudaMemsetAsync(fail_flag, 0, sizeof(int), h->st));
        MV_CUDA(h, cudaMemsetAsync(h->Sd.p, 0, h->Sd.bytes(), h->st));
        // Super-blocks are cut into chunks of Bc (power of two); this rank owns chunks [c0, c1).  With the
        // pre-reduction (Lc > 1, Lc divides Bc) the own range is first reduced to its chunk heads; "blocks"
        // below are then heads and Bc counts heads.
        // Local levels (strides < Bc) run on the view [lo, hi] whose last block is a zeroed GHOST of
        // the next rank's first block: it collects the left-side Schur updates.  The Bc-chunk heads
        // form the top system (nchunks blocks), summed over ranks and eliminated redundantly.
        const int64_t Bc = h->Bc, nchunks = nb / Bc;
        const int64_t c0 = nchunks * h->rank / h->world, c1 = nchunks * (h->rank + 1) / h->world;
        const int64_t lo = c0 * Bc, hi = c1 * Bc, nloc = hi - lo;
        const int64_t Bs = Bc / Lc, slo = lo / Lc, shi = hi / Lc, sloc = shi - slo;    // the same in units of system blocks
        // damped working copies of the own range; ghost slot zero
        const double* wsrc = h->Wp();
        if (nloc > 0) {
            damp_copy_kernel<<<(int)((nloc * q * q + 255) / 256), 256, 0, h->st>>>(
                h->D.p + lo * qq, h->diag_s.p + lo * q, lam_p, nloc * q, q,
                std::max<int64_t>(0, 3 * h->n_ctrl - lo * q), h->Dw.p + lo * qq);
            MV_CUDA(h, cudaMemcpyAsync(h->Ew.p + lo * qq, h->E.p + lo * qq, nloc * qq * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
            h->launches++;
            if (!pre || h->desc.rs_bounds) {
                MV_CUDA(h, cudaMemcpyAsync(h->Ww.p + lo * wn, h->Wp() + lo * wn, (size_t)nloc * wn * sizeof(double), cudaMemcpyDeviceToDevice, h->st));
                wsrc = h->Ww.p;
            }
            if (h->desc.rs_bounds) {
                freeze_cols_kernel<<<(int)((nloc * q + 255) / 256), 256, 0, h->st>>>(h->Ww.p + lo * wn, nloc * q, ldw, h->nc, h->Pc, h->frozen.p);
                h->launches++;
            }
        }
        cudaEventRecord(h->evs[0], h->st);
        if (pre) {
            if (nloc > 0) prereduce_range(h, lo, hi, wsrc, fail_flag);      // zeroes the ghost head itself
            else {
                MV_CUDA(h, cudaMemsetAsync(sD + shi * qq, 0, qq * sizeof(double), h->st));
                MV_CUDA(h, cudaMemsetAsync(sE + shi * qq, 0, qq * sizeof(double), h->st));
                MV_CUDA(h, cudaMemsetAsync(sW + shi * wn, 0, wn * sizeof(double), h->st));
            }
        } else {
            MV_CUDA(h, cudaMemsetAsync(sD + shi * qq, 0, qq * sizeof(double), h->st));
            MV_CUDA(h, cudaMemsetAsync(sE + shi * qq, 0, qq * sizeof(double), h->st));
            MV_CUDA(h, cudaMemsetAsync(sW + shi * wn, 0, wn * sizeof(double), h->st));
        }
        BcrView lv{sD + slo * qq, sE + slo * qq, sW + slo * wn, sZ + slo * qq, sd + slo * q, sloc + 1};
        std::vector<int64_t> llev;
        if (sloc > 0 && Bs > 1) {
            llev = bcr_eliminate(h, lv, Bs, false, fail_flag);
            launch_level(h, lv, (int)((lv.nb + Bs - 1) / Bs), Bs, Bs / 2, 2, fail_flag);     // pending updates of the last local level
        }
        // top system
        const size_t tD = (size_t)nchunks * qq, tW = (size_t)nchunks * wn;
        MV_CUDA(h, cudaMemsetAsync(h->Dt.p, 0, (2 * tD + tW) * sizeof(double), h->st));
        double* Dt = h->Dt.p; double* Et = Dt + tD; double* Wt = Et + tD;
        if (sloc > 0) {
            top_gather_kernel<<<(int)(c1 - c0 + 1), 128, 0, h->st>>>(sD, sE, sW, q, ldw, Bs, c0, c1, nchunks, Dt, Et, Wt);
            h->launches++;
            // chunk heads leave the local system: their rows must not enter the local SYRK
            MV_CUDA(h, cudaMemset2DAsync(sW + slo * wn, (size_t)Bs * wn * sizeof(double), 0, wn * sizeof(double), (size_t)(c1 - c0), h->st));
            if (pre) {                         // the local heads' rows join the eliminated ones (Ww) for the SYRK
                head_rows_kernel<<<(int)sloc, 256, 0, h->st>>>(slo, Lc, q, ldw, h->Wh.p, h->Ww.p);
                h->launches++;
            }
        }
        int e = nccl_sum(h, Dt, 2 * tD + tW);
        if (e) return e;
        BcrView tv{Dt, Et, Wt, h->ZLt.p, h->dst.p, nchunks};
        std::vector<int64_t> tlev = bcr_eliminate(h, tv, nchunks, true, fail_flag);
        // Schur complement: local rows (+ the replicated top rows once, on rank 0), summed over ranks
        cudaEventRecord(h->evs[1], h->st);
        launch_syrk(h, h->Ww.p + lo * wn, nloc * q, h->Sd.p);
        if (h->rank == 0) launch_syrk(h, Wt, nchunks * q, h->Sd.p);
        cudaEventRecord(h->evs[2], h->st);
        timed = true;
        e = nccl_sum(h, h->Sd.p, (size_t)ldw * ldw);
        if (e) return e;
        form_schur_kernel<<<(int)(((int64_t)h->ncP * h->ncP + 255) / 256), 256, 0, h->st>>>(
            h->A.p, bc, h->diag_c.p, lam_p, h->nc, h->Pc, ldw, h->frozen.p, h->Ax.p, h->Sd.p, rhs);
        h->launches++;
        dense_chol_solve(h, h->Sd.p, ldw, h->ncP, rhs, h->dlt_c.p, fail_flag);
        e = nccl_bcast0(h, h->dlt_c.p, (size_t)h->ncP);
        if (e) return e;
        // back substitution: top system (redundant), then the local levels, then inside the chunks
        wdc_kernel<<<(int)((nchunks * q * 32 + 255) / 256), 256, 0, h->st>>>(Wt, h->dlt_c.p, nchunks * q, ldw, h->dst.p);
        h->launches++;
        bcr_back(h, tv, tlev, true);
        if (nloc > 0) {
            wdc_kernel<<<(int)((nloc * q * 32 + 255) / 256), 256, 0, h->st>>>(h->Ww.p + lo * wn, h->dlt_c.p, nloc * q, ldw, h->dlt_s.p + lo * q);
            h->launches++;
            if (pre) {
                wdc_kernel<<<(int)((sloc * q * 32 + 255) / 256), 256, 0, h->st>>>(sW + slo * wn, h->dlt_c.p, sloc * q, ldw, sd + slo * q);
                h->launches++;
            }
            top_scatter_kernel<<<(int)(c1 - c0 + 1), 32, 0, h->st>>>(h->dst.p, q, Bs, c0, c1, nchunks, sd);
            h->launches++;
            bcr_back(h, lv, llev, false);
            if (pre) chunk_back_range(h, lo, hi);
        }
        zero_outside_kernel<<<(int)(((nbq + q) + 255) / 256), 256, 0, h->st>>>(h->dlt_s.p, nbq + q, lo * q, hi * q);
        h->launches++;
        e = nccl_sum(h, h->dlt_s.p, (size_t)nbq);
        if (!e) e = nccl_max_flag(h, fail_flag);
        if (e) return e;
    }
    MV_CUDA(h, cudaGetLastError());
    int f = 0;
    MV_CUDA(h, cudaMemcpyAsync(&f, fail_flag, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    if (timed) {
        float a = 0.f, b = 0.f;
        cudaEventElapsedTime(&a, h->evs[0], h->evs[1]);
        cudaEventElapsedTime(&b, h->evs[1], h->evs[2]);
        h->ms_bcr += a; h->ms_syrk += b;
    }
    *ok = f ? 0 : 1;
    return MVUS_OK;
}

inline int read_cost(mvus_ba_ctx* h, double* cost) {
    MV_CUDA(h, cudaMemcpyAsync(h->h_pin, h->partial.p + h->cost_slot, sizeof(double), cudaMemcpyDeviceToHost, h->st));
    MV_CUDA(h, cudaStreamSynchronize(h->st));
    *cost = 0.5 * h->h_pin[0];
    return MVUS_OK;
}


}  // namespace mvus
