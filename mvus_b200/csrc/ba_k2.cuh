// K2 (FP64-MMA version): J^T J / J^T r accumulation, one WARP per run of detections that share a
// camera and a knot span.  The per-run update  S_run = sum_t (u_t u_t^T + v_t v_t^T)  over the
// ~5-15 detections of a run is a tiny SYRK; it is done with mma.sync.m8n8k4.f64 fragments:
//   * the P+1 columns are packed into NT = ceil((P+1)/8) tiles in the order
//        [camera 0..PC-1 | residual | control 0..11 | padding],
//   * the k dimension packs the run's u rows followed by its v rows (2 n rows, 4 per MMA step),
//   * every tile pair (i <= j) is one 8x8 accumulator fragment; lane (fm = lane>>2, fk = lane&3)
//     holds C[8 i + fm][8 j + 2 fk + e], e = 0, 1.
// Entries whose column is a control point are flushed at the end of the run with predicated FP64
// REDs (camera x control -> W~ with the 8 lanes of equal fk writing 8 consecutive columns of one
// row; control x control -> E, or the UPPER triangle of D: damp_copy_kernel mirrors it);
// camera-only entries stay in registers for the whole tile, are summed over the CTA's warps
// through shared memory and written once per tile.
// Against the scalar version (one thread per entry, 4 shared loads per 2 FMAs, bound by
// shared-memory bandwidth) a run of 5 detections costs 9 shared loads + 18 MMAs per warp.
// HBM traffic per detection (algorithmic): read r (16 B) + span (4 B) + J (16 P B)
//   -> 356 B (P=21) / 500 B (P=30); writes are O(runs), not O(detections).
#pragma once
#include "ba_ctx.cuh"

namespace mvus {

template <int P>
struct K2Cfg {
    static constexpr int PC = P - 12;                  // camera unknowns (9 or 18)
    static constexpr int NSLOT = P + 1;                // camera | residual | 12 control columns
    static constexpr int NT = (NSLOT + 7) / 8;
    static constexpr int NPAIR = NT * (NT + 1) / 2;
    static constexpr int TR = PC / 8;                  // tile of the residual slot (last camera-side slot)
    static constexpr int TC = (PC + 1) / 8;            // first tile with a control slot
    static constexpr int NCT = NT - TC;                // tiles with control slots
    static constexpr int CPAD = PC + 1 - 8 * TC;       // camera-side slots at the start of tile TC
    static constexpr int NKEEP = (TR + 1) * (TR + 2) / 2;   // tile pairs with camera-only entries
    static constexpr int THREADS = 256, WARPS = THREADS / 32;
    static constexpr int LDT = TILE_DET + 4;           // +4: conflict-free fragment loads (4 fm + fk pattern)
    static constexpr int SPLIT = 16;                   // forced run split when a tile has few runs
    static constexpr int CT = 8 * NCT;                 // run-table entries per run: position in the control tiles
    // staged planes + span + run start + run span + per-run control-column table + misc
    static constexpr size_t SMEM = (size_t)(2 * (P + 1)) * LDT * sizeof(double) +
                                   (size_t)(TILE_DET + (TILE_DET + 1) + TILE_DET + CT * TILE_DET + 8) * sizeof(int);
    static_assert(8 * TC <= PC, "row tiles below TC must hold camera columns only");
    static_assert(TC == TR, "the residual slot must sit in the first control tile");
    static_assert((size_t)WARPS * NKEEP * 64 <= (size_t)(2 * (P + 1)) * LDT, "partial sums must fit the staging area");
};

__device__ __forceinline__ void k2_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void k2_cp_async8(unsigned dst, const double* src, unsigned src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

// Predicated FP64 reduction (a flush is 10-20 of these with lane-dependent predicates).
__device__ __forceinline__ void k2_red(double* ptr, double v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}"
                 :: "l"(ptr), "d"(v), "r"((int)pred) : "memory");
}

// Accumulator bits (2 k + e, k = index of tile pair (i <= j)) that use column (TC + jj, e) / row tile TC + ii.
template <int NT, int TC>
__device__ __forceinline__ constexpr unsigned k2_colbits(int jj, int e) {
    unsigned bits = 0;
    int k = 0;
    for (int i = 0; i < NT; ++i)
        for (int j = i; j < NT; ++j, ++k)
            if (j == TC + jj) bits |= 1u << (2 * k + e);
    return bits;
}
template <int NT, int TC>
__device__ __forceinline__ constexpr unsigned k2_rowbits(int ii) {
    unsigned bits = 0;
    int k = 0;
    for (int i = 0; i < NT; ++i)
        for (int j = i; j < NT; ++j, ++k)
            if (i == TC + ii && j >= TC) bits |= 3u << (2 * k);
    return bits;
}

template <int P>
__global__ void __launch_bounds__(256, (P == 21 ? 4 : 3))
accumulate_kernel(const double* __restrict__ J, const double* __restrict__ r, const int* __restrict__ span,
                  const int* __restrict__ tile_cam, const int64_t* __restrict__ tile_start,
                  const int* __restrict__ tile_cnt, const int64_t* __restrict__ row_off, int64_t N,
                  int bw, int ldw, double* __restrict__ A, double* __restrict__ bc,
                  double* __restrict__ D, double* __restrict__ E, double* __restrict__ W) {
    using Cfg = K2Cfg<P>;
    constexpr int NT = Cfg::NT, PC = Cfg::PC, LDT = Cfg::LDT, TR = Cfg::TR, TC = Cfg::TC;
    constexpr int NCT = Cfg::NCT, CPAD = Cfg::CPAD, CT = Cfg::CT;
    constexpr int VOFF = (P + 1) * LDT;
    extern __shared__ double s_mem[];
    double* s_J = s_mem;                                        // [2*(P+1)][LDT]: u planes (P = r_u), then v planes
    int* s_span = reinterpret_cast<int*>(s_mem + (size_t)2 * (P + 1) * LDT);
    int* s_rstart = s_span + TILE_DET;                          // [TILE_DET + 1]
    int* s_rg = s_rstart + TILE_DET + 1;                        // [TILE_DET] span index of the run
    int* s_ctab = s_rg + TILE_DET;                              // [TILE_DET][CT] (global row << 5 | local column) of
                                                                //   the control column at position x of the control
                                                                //   tiles, -1 if that position is no control point
    int* s_misc = s_ctab + CT * TILE_DET;                       // [0] = number of runs, [1..4] warp counts, [5] next run
    const int tl = blockIdx.x, cam = tile_cam[tl], cnt = tile_cnt[tl];
    const int64_t d0 = tile_start[tl];
    const int q = 3 * bw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- stage the tile with 8-byte cp.async (all 2 (P+1) planes of a thread in flight at once;
    //      zero fill past the end of the tile); the run table is built while the copies land
    {
        const int t = tid & (TILE_DET - 1), p0 = tid >> 7;
        const unsigned sz = t < cnt ? 8u : 0u;
        const int64_t r0 = row_off[cam], ncam = (row_off[cam + 1] - r0) >> 1;
        const int64_t loc = d0 + t - (r0 >> 1);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const double* src = J + (int64_t)(half * P + p0) * N + (sz ? d0 + t : 0);
            unsigned dst = (unsigned)__cvta_generic_to_shared(s_J + (half * (P + 1) + p0) * LDT + t);
#pragma unroll
            for (int p = p0; p < P; p += 2) {
                k2_cp_async8(dst, src, sz);
                src += 2 * N;
                dst += 2 * LDT * 8;
            }
            if (p0 == (P & 1))
                k2_cp_async8((unsigned)__cvta_generic_to_shared(s_J + (half * (P + 1) + P) * LDT + t),
                             r + (sz ? r0 + half * ncam + loc : 0), sz);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    if (tid < TILE_DET) s_span[tid] = tid < cnt ? span[d0 + tid] : -2;
    if (tid == 0) s_misc[5] = Cfg::WARPS;                       // runs 0..WARPS-1 are taken statically
    __syncthreads();
    // ---- run table: maximal runs of equal span index; a tile with few runs is additionally cut
    //      every SPLIT slots so that all warps have work (the partial sums simply add up)
    bool head = tid < cnt && (tid == 0 || s_span[tid] != s_span[tid - 1]);
    const int natural = __syncthreads_count(head);
    if (natural < 12) head = head || (tid < cnt && (tid & (Cfg::SPLIT - 1)) == 0);
    if (tid < TILE_DET) {
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_misc[1 + warp] = __popc(bal);
        s_rg[tid] = head ? (int)(__popc(bal & ((1u << lane) - 1u))) : -1;          // rank inside the warp
    }
    __syncthreads();
    {
        int rk = -1, g = 0, base = 0;
        if (tid < TILE_DET) {
            for (int w = 0; w < warp; ++w) base += s_misc[1 + w];
            rk = s_rg[tid];
            g = s_span[tid];
        }
        __syncthreads();                                        // every rank is read before s_rg is rewritten
        if (tid == 0) s_misc[0] = s_misc[1] + s_misc[2] + s_misc[3] + s_misc[4];
        if (rk >= 0) {
            s_rstart[base + rk] = tid;
            s_rg[base + rk] = g;
        }
    }
    __syncthreads();
    const int nruns = s_misc[0];
    if (tid == 0) s_rstart[nruns] = cnt;
    for (int x = tid; x < nruns * CT; x += Cfg::THREADS) {
        const int g = s_rg[x / CT], cb = x % CT - CPAD;         // control column 0..11 of the run's 4 x 3 window
        int packed = -1;
        if (g >= 0 && cb >= 0 && cb < 12) {
            const int m = cb / 3, j = g - 3 + m;
            if (j >= 0) {
                const int kb = j / bw, lc = (j - kb * bw) * 3 + (cb - 3 * m);
                packed = ((kb * q + lc) << 5) | lc;
            }
        }
        s_ctab[x] = packed;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- one warp per run
    const int fk = lane & 3, fm = lane >> 2;
    int pl[NT];                                   // staged plane (in doubles) of slot 8 i + fm, -1 = padding
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int s = 8 * i + fm;
        pl[i] = s < PC ? s * LDT : (s == PC ? P * LDT : (s <= PC + 12 ? (s - 1) * LDT : -1));
    }
    // lane constants of the flush: which accumulator entries (bit 2 k + e) this lane can ever flush
    unsigned vmask = 0;
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = i; j < NT; ++j, ++k) {
                if (j < TC) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a = 8 * i + fm, b = 8 * j + 2 * fk + e, cb = b - PC - 1;
                    if (cb >= 0 && cb < 12 && a <= b) vmask |= 1u << (2 * k + e);   // a <= b: never a padding row
                }
            }
    }
    const bool rW0 = fm < CPAD;                   // this lane's row of tile TC is a camera / residual row (target W~)
    double* const Wc0 = W + cam * PC + fm;        // rows of the camera-only tiles: column cam*PC + 8 i + fm
    double* const Wr0 = W + (fm == CPAD - 1 ? ldw - 1 : cam * PC + 8 * TC + fm);   // the same for a W~ row of tile TC
    const int flip0 = fm == CPAD - 1 ? (int)0x80000000 : 0;   // residual row: W~'s last column = -J^T r
    double acc[Cfg::NPAIR][2];
#pragma unroll
    for (int k = 0; k < Cfg::NPAIR; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }

    int rr = warp;
    while (rr < nruns) {
        int rnext = 0;                               // next run: taken from the CTA's counter (balances the warps)
        if (lane == 0) rnext = atomicAdd(&s_misc[5], 1);
        const int g = s_rg[rr];
        if (g >= 0) {                                // g < 0: uncovered detections, zero rows
            const int t0 = s_rstart[rr], n = s_rstart[rr + 1] - t0;
            const int nsteps = (2 * n + 3) >> 2;
            double fc[NT], fn[NT];
            {
                const int rho = fk;
                const bool hv = rho >= n;
                const int off = t0 + rho + (hv ? VOFF - n : 0);
                const bool ok = rho < 2 * n;
#pragma unroll
                for (int i = 0; i < NT; ++i) fc[i] = (ok && pl[i] >= 0) ? s_J[pl[i] + off] : 0.0;
            }
#pragma unroll 1
            for (int s = 0; s < nsteps; ++s) {
                {
                    const int rho = 4 * (s + 1) + fk;
                    const bool hv = rho >= n;
                    const int off = t0 + rho + (hv ? VOFF - n : 0);
                    const bool ok = rho < 2 * n;
#pragma unroll
                    for (int i = 0; i < NT; ++i) fn[i] = (ok && pl[i] >= 0) ? s_J[pl[i] + off] : 0.0;
                }
                int k = 0;
#pragma unroll
                for (int i = 0; i < NT; ++i)
#pragma unroll
                    for (int j = i; j < NT; ++j, ++k) k2_dmma(acc[k][0], acc[k][1], fc[i], fc[j]);
#pragma unroll
                for (int i = 0; i < NT; ++i) fc[i] = fn[i];
            }
            // ---- flush the entries whose column is a control point (slots PC+1 .. PC+12).  Addresses are
            //      (row pointer) + (column offset): both are set up once per run, an entry costs a select,
            //      an address add and the predicated RED.  D receives its upper triangle only.
            const int* ct = s_ctab + rr * CT;
            unsigned m = vmask;
            int64_t cw[NCT][2];                      // W~: row offset (row * ldw) of the column's control row
            int ccl[NCT][2], cblk[NCT][2];           // D / E: local column, first row of the super-block
#pragma unroll
            for (int jj = 0; jj < NCT; ++jj)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int pk = ct[8 * jj + 2 * fk + e];
                    const int rb = pk >> 5;
                    cw[jj][e] = (int64_t)rb * ldw;
                    ccl[jj][e] = pk & 31;
                    cblk[jj][e] = rb - (pk & 31);
                    if (pk < 0) m &= ~k2_colbits<NT, TC>(jj, e);
                }
            double* pd[NCT];                         // row pointers: D / E row of a control row, W~ column of a camera row
            double* pe[NCT];
            int ablk[NCT];
#pragma unroll
            for (int ii = 0; ii < NCT; ++ii) {
                const int pk = ct[8 * ii + fm];
                const int ra = pk >> 5;
                const bool rw = ii == 0 && rW0;
                ablk[ii] = ra - (pk & 31);
                pd[ii] = rw ? Wr0 : D + (int64_t)ra * q;
                pe[ii] = rw ? Wr0 : E + (int64_t)ra * q;
                if (!rw && pk < 0) m &= ~k2_rowbits<NT, TC>(ii);
            }
            int k = 0;
#pragma unroll
            for (int i = 0; i < NT; ++i) {
#pragma unroll
                for (int j = i; j < NT; ++j, ++k) {
                    if (j < TC) continue;            // camera-only tile pair: stays in registers
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double v = acc[k][e];
                        if (j > TR) acc[k][e] = 0.0;
                        else if (2 * fk + e >= CPAD) acc[k][e] = 0.0;    // j == TR == TC: control column
                        const bool pred = ((m >> (2 * k + e)) & 1u) && v != 0.0;
                        if (i < TC) {                // camera rows only
                            k2_red(Wc0 + 8 * i + cw[j - TC][e], v, pred);
                        } else {
                            const int ii = i - TC;
                            const bool rw = ii == 0 && rW0;
                            double* base = ablk[ii] == cblk[j - TC][e] ? pd[ii] : pe[ii];
                            const int64_t off = rw ? cw[j - TC][e] : (int64_t)ccl[j - TC][e];
                            const double val = ii == 0 ? __hiloint2double(__double2hiint(v) ^ flip0, __double2loint(v)) : v;
                            k2_red(base + off, val, pred);
                        }
                    }
                }
            }
        }
        rr = __shfl_sync(0xffffffffu, rnext, 0);
    }

    // ---- camera-only entries: sum the warps' fragments through shared memory, one RED per tile
    __syncthreads();                                  // every warp is done with the staged tile
    double* s_part = s_J;                             // [WARPS][NKEEP][2][32]
    {
        int k = 0, kk = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = i; j < NT; ++j, ++k) {
                if (j > TR) continue;
                s_part[((warp * Cfg::NKEEP + kk) * 2 + 0) * 32 + lane] = acc[k][0];
                s_part[((warp * Cfg::NKEEP + kk) * 2 + 1) * 32 + lane] = acc[k][1];
                ++kk;
            }
    }
    __syncthreads();
    for (int x = tid; x < Cfg::NKEEP * 64; x += Cfg::THREADS) {
        const int kk = x >> 6, e = (x >> 5) & 1, ln = x & 31;
        int i = 0, j = 0, c = kk;                     // kept pairs are enumerated (i, j), i <= j <= TR
        while (c >= TR + 1 - i) { c -= TR + 1 - i; ++i; }
        j = i + c;
        const int a = 8 * i + (ln >> 2), b = 8 * j + 2 * (ln & 3) + e;
        if (a > b || a >= PC || b > PC) continue;
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < Cfg::WARPS; ++w) v += s_part[((w * Cfg::NKEEP + kk) * 2 + e) * 32 + ln];
        if (v == 0.0) continue;
        if (b == PC) atomicAdd(bc + cam * PC + a, -v);
        else {
            atomicAdd(A + ((int64_t)cam * PC + a) * PC + b, v);
            if (a != b) atomicAdd(A + ((int64_t)cam * PC + b) * PC + a, v);
        }
    }
}

}  // namespace mvus
