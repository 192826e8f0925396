// K2 (FP64-MMA version): J^T J / J^T r accumulation, one WARP per run of detections that share a
// camera and a knot span.  The per-run update  S_run = sum_t (u_t u_t^T + v_t v_t^T)  over the
// ~5-15 detections of a run is a tiny SYRK; it is done with mma.sync.m8n8k4.f64 fragments:
//   * the P+1 columns are packed into NT = ceil((P+1)/8) tiles in the order
//        [camera 0..PC-1 | residual | control 0..11 | padding],
//   * the k dimension packs the run's u rows followed by its v rows (2 n rows, 4 per MMA step),
//   * every tile pair (i <= j) is one 8x8 accumulator fragment; lane (fm = lane>>2, fk = lane&3)
//     holds C[8 i + fm][8 j + 2 fk + e], e = 0, 1.
// A warp owns a contiguous chunk of the tile (16 detections, ~3 runs at config 4) and keeps a
// SLIDING WINDOW over the control points: control point j always lives in column slot j mod 4,
// so consecutive runs (spans g, g+1, ...) share three of their four slots and only the slot
// whose control point leaves the window is flushed (predicated FP64 REDs: camera x control ->
// W~, the 8 lanes of equal fk writing 8 consecutive columns of one row; control x control -> E
// or the UPPER triangle of D, damp_copy_kernel mirrors it).  Camera-only entries stay in
// registers for the whole tile, are summed over the CTA's warps through shared memory and written
// once per tile.  Tiles are visited in TIME order across cameras (tile_key_kernel + CUB sort) so
// that the RED targets of concurrently running CTAs stay in L2: in camera-major order the DRAM
// round trip of D, E and W~ cost 11 of 22 ms at config 4.  tests/proto/k2_window_proto.py is the
// NumPy model of this bookkeeping (checked against the dense J^T J on the CPU).
// Against the scalar version (one thread per entry, 4 shared loads per 2 FMAs, bound by
// shared-memory bandwidth: 29.6 ms) a run of 5 detections costs 9 shared loads + 18 MMAs per warp;
// this version takes 15.4 ms and is bound by issue slots (flush bookkeeping), profiles/r1_notes.md.
// HBM traffic per detection (algorithmic): read r (16 B) + span (4 B) + J (16 P B)
//   -> 356 B (P=21) / 500 B (P=30); writes are O(runs), not O(detections).
#pragma once
#include <cub/cub.cuh>
#include "ba_ctx.cuh"

namespace mvus {

constexpr int K2_TILE = TILE_DET;      // 64 measured the same (15.5 vs 15.4 ms at config 4)

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
    static constexpr int KT = K2_TILE;                 // detections per CTA (a K1 tile is cut into TILE_DET / KT parts)
    static constexpr int SPLITS = TILE_DET / KT;
    static constexpr int THREADS = 2 * KT, WARPS = THREADS / 32;
    static constexpr int LDT = KT + 4;           // +4: conflict-free fragment loads (4 fm + fk pattern)
    static constexpr int CHUNK = KT / WARPS;     // detections per warp; runs are cut at chunk boundaries
    static constexpr int CT = 8 * NCT;                 // positions in the control tiles
    // staged planes + span + run start + run span + per-run control-column table + misc
    static constexpr size_t SMEM = (size_t)(2 * (P + 1)) * LDT * sizeof(double) +
                                   (size_t)(KT + (KT + 1) + KT + 12 * KT + 16 + 8) * sizeof(int);
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

// Tile order of K2: by the span index of the tile's first detection, i.e. by TIME across all cameras.
// Tiles that run concurrently then update the same few hundred control points: the D / E blocks and
// W~ rows they RED into stay in L2 (in camera-major order every camera's sweep re-fetched all of D, E
// and its 72-byte segments of every W~ row from DRAM: 14 GB read + 10 GB written at config 4).
__global__ void tile_key_kernel(const int* __restrict__ span, const int64_t* __restrict__ tile_start,
                                int n_tiles, int* __restrict__ key, int* __restrict__ id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_tiles) { key[i] = span[tile_start[i]]; id[i] = i; }
}

template <int P>
__global__ void __launch_bounds__(K2Cfg<P>::THREADS, (P == 21 ? 4 : 3) * (256 / K2Cfg<P>::THREADS))
accumulate_kernel(const double* __restrict__ J, const double* __restrict__ r, const int* __restrict__ span,
                  const int* __restrict__ tile_perm, const int* __restrict__ tile_cam,
                  const int64_t* __restrict__ tile_start, const int* __restrict__ tile_cnt,
                  const int64_t* __restrict__ row_off, int64_t N,
                  int bw, int ldw, double* __restrict__ A, double* __restrict__ bc,
                  double* __restrict__ D, double* __restrict__ E, double* __restrict__ W) {
    using Cfg = K2Cfg<P>;
    constexpr int NT = Cfg::NT, PC = Cfg::PC, LDT = Cfg::LDT, TR = Cfg::TR, TC = Cfg::TC;
    constexpr int NCT = Cfg::NCT, CPAD = Cfg::CPAD, CT = Cfg::CT, KT = Cfg::KT;
    constexpr int VOFF = (P + 1) * LDT;
    extern __shared__ double s_mem[];
    double* s_J = s_mem;                                        // [2*(P+1)][LDT]: u planes (P = r_u), then v planes
    int* s_span = reinterpret_cast<int*>(s_mem + (size_t)2 * (P + 1) * LDT);
    int* s_rstart = s_span + KT;                                // [KT + 1]
    int* s_rg = s_rstart + KT + 1;                              // [KT] span index of the run
    int* s_ctab = s_rg + KT;                                    // [KT][12] (global row << 5 | local column) of the
                                                                //   run's control column slot*3 + axis, -1 if none
    int* s_leave = s_span;                                      // [KT] slots (bit mask) to flush after the run
                                                                //   (aliases s_span, dead once the runs are known)
    int* s_first = s_ctab + 12 * KT;                          // [WARPS + 1] first run of each warp's chunk
    int* s_misc = s_first + 16;                                 // [0] = number of runs, [1..4] warp counts
    const int tl = tile_perm[blockIdx.x / Cfg::SPLITS], part = blockIdx.x % Cfg::SPLITS, cam = tile_cam[tl];
    const int cnt = min(tile_cnt[tl] - part * KT, KT);
    if (cnt <= 0) return;
    const int64_t d0 = tile_start[tl] + part * KT;
    const int q = 3 * bw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // ---- stage the tile with 8-byte cp.async (all 2 (P+1) planes of a thread in flight at once;
    //      zero fill past the end of the tile); the run table is built while the copies land
    {
        const int t = tid & (KT - 1), p0 = tid / KT;
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
    if (tid < KT) s_span[tid] = tid < cnt ? span[d0 + tid] : -2;
    __syncthreads();
    // ---- run table: runs of equal span index, cut at the boundaries of the warps' chunks
    const bool head = tid < cnt && (tid == 0 || s_span[tid] != s_span[tid - 1] || (tid & (Cfg::CHUNK - 1)) == 0);
    if (tid < KT) {
        const unsigned bal = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_misc[1 + warp] = __popc(bal);
        s_rg[tid] = head ? (int)(__popc(bal & ((1u << lane) - 1u))) : -1;          // rank inside the warp
    }
    __syncthreads();
    {
        int rk = -1, g = 0, base = 0;
        if (tid < KT) {
            for (int w = 0; w < warp; ++w) base += s_misc[1 + w];
            rk = s_rg[tid];
            g = s_span[tid];
        }
        __syncthreads();                                        // every rank is read before s_rg is rewritten
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < KT / 32; ++w) tot += s_misc[1 + w];
            s_misc[0] = tot;
        }
        if (rk >= 0) {
            s_rstart[base + rk] = tid;
            s_rg[base + rk] = g;
        }
    }
    __syncthreads();
    const int nruns = s_misc[0];
    if (tid == 0) s_rstart[nruns] = cnt;
    // control point j lives in slot j & 3: position x = CPAD + 3 slot + axis holds, for a run of span g,
    // the control point j = g - ((g - slot) & 3) of its window g-3 .. g
    for (int x = tid; x < nruns * 12; x += Cfg::THREADS) {
        const int g = s_rg[x / 12], cb = x % 12;
        int packed = -1;
        if (g >= 0) {
            const int sl = cb / 3, j = g - ((g - sl) & 3);
            if (j >= 0) {
                const int kb = j / bw, lc = (j - kb * bw) * 3 + (cb - 3 * sl);
                packed = ((kb * q + lc) << 5) | lc;
            }
        }
        s_ctab[x] = packed;
    }
    __syncthreads();                                            // s_rstart complete
    // slots to flush after a run: those whose control point differs in the next run of the same warp
    // (all four at the end of the warp's chunk or before uncovered detections)
    if (tid < nruns) {
        const int g = s_rg[tid], nx = tid + 1;
        int leave = 0xF;
        if (nx < nruns && (s_rstart[nx] / Cfg::CHUNK) == (s_rstart[tid] / Cfg::CHUNK) && s_rg[nx] >= 0 && g >= 0) {
            const int gn = s_rg[nx];
            leave = 0;
#pragma unroll
            for (int sl = 0; sl < 4; ++sl)
                if (g - ((g - sl) & 3) != gn - ((gn - sl) & 3)) leave |= 1 << sl;
        }
        s_leave[tid] = leave;
    }
    // first run of each warp's chunk = number of run heads before slot CHUNK * w
    if (tid <= Cfg::WARPS) {
        int lo = 0, hi = nruns;                                 // first run with start >= CHUNK * tid
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s_rstart[mid] < Cfg::CHUNK * tid) lo = mid + 1; else hi = mid;
        }
        s_first[tid] = tid == Cfg::WARPS ? nruns : lo;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- one warp per chunk of CHUNK detections, run by run
    const int fk = lane & 3, fm = lane >> 2;
    int pl[NT];                                   // staged plane (in doubles) of slot 8 i + fm, -1 = padding;
                                                  //   control slots: plane of axis ax in window position 0
    int psl[NCT];                                 // control slots: slot id 0..3 (else -1)
#pragma unroll
    for (int i = 0; i < NT; ++i) {
        const int s = 8 * i + fm;
        pl[i] = s < PC ? s * LDT : (s == PC ? P * LDT : -1);
        if (i >= TC) {
            const int cb = s - PC - 1;
            const bool ok = cb >= 0 && cb < 12;
            psl[i - TC] = ok ? cb / 3 : -1;
            if (ok) pl[i] = (PC + cb - 3 * (cb / 3)) * LDT;
        }
    }
    // lane constants of the flush: which accumulator entries (bit 2 k + e) this lane can ever flush, and the
    // slots they belong to (4-bit sets: columns (jj, e) at bits 4 (2 jj + e), row tiles at bits 16 + 4 ii)
    unsigned vmask = 0, sbits = 0;
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
#pragma unroll
        for (int jj = 0; jj < NCT; ++jj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cb = 8 * jj + 2 * fk + e - CPAD;
                if (cb >= 0 && cb < 12) sbits |= (1u << (cb / 3)) << (4 * (2 * jj + e));
            }
#pragma unroll
        for (int ii = 0; ii < NCT; ++ii) {
            const int ca = 8 * ii + fm - CPAD;
            if (ca >= 0 && ca < 12) sbits |= (1u << (ca / 3)) << (16 + 4 * ii);
        }
    }
    const bool rW0 = fm < CPAD;                   // this lane's row of tile TC is a camera / residual row (target W~)
    double* const Wc0 = W + cam * PC + fm;        // rows of the camera-only tiles: column cam*PC + 8 i + fm
    double* const Wr0 = W + (fm == CPAD - 1 ? ldw - 1 : cam * PC + 8 * TC + fm);   // the same for a W~ row of tile TC
    const int flip0 = fm == CPAD - 1 ? (int)0x80000000 : 0;   // residual row: W~'s last column = -J^T r
    double acc[Cfg::NPAIR][2];
#pragma unroll
    for (int k = 0; k < Cfg::NPAIR; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }

    const int rr_end = s_first[warp + 1];
    for (int rr = s_first[warp]; rr < rr_end; ++rr) {
        const int g = s_rg[rr];
        if (g < 0) continue;                         // uncovered detections: zero rows
        const int t0 = s_rstart[rr], n = s_rstart[rr + 1] - t0;
        const int nsteps = (2 * n + 3) >> 2;
        int plr[NT];                                 // planes of this run: window position of slot sl is (sl - g - 1) & 3
#pragma unroll
        for (int i = 0; i < NT; ++i) {
            plr[i] = pl[i];
            if (i >= TC && psl[i - TC] >= 0) plr[i] = pl[i] + ((psl[i - TC] - g - 1) & 3) * (3 * LDT);
        }
        double fc[NT], fn[NT];
        {
            const int rho = fk;
            const bool hv = rho >= n;
            const int off = t0 + rho + (hv ? VOFF - n : 0);
            const bool ok = rho < 2 * n;
#pragma unroll
            for (int i = 0; i < NT; ++i) fc[i] = (ok && plr[i] >= 0) ? s_J[plr[i] + off] : 0.0;
        }
#pragma unroll 1
        for (int s = 0; s < nsteps; ++s) {
            {
                const int rho = 4 * (s + 1) + fk;
                const bool hv = rho >= n;
                const int off = t0 + rho + (hv ? VOFF - n : 0);
                const bool ok = rho < 2 * n;
#pragma unroll
                for (int i = 0; i < NT; ++i) fn[i] = (ok && plr[i] >= 0) ? s_J[plr[i] + off] : 0.0;
            }
            int k = 0;
#pragma unroll
            for (int i = 0; i < NT; ++i)
#pragma unroll
                for (int j = i; j < NT; ++j, ++k) k2_dmma(acc[k][0], acc[k][1], fc[i], fc[j]);
#pragma unroll
            for (int i = 0; i < NT; ++i) fc[i] = fn[i];
        }
        // ---- flush the entries of the slots that leave the window (all at the end of the chunk).
        //      Addresses are (row pointer) + (column offset), set up once per flush; an entry costs a few
        //      selects, an address add and the predicated RED.  D receives its upper triangle only.
        const unsigned lm = (unsigned)s_leave[rr];
        if (lm == 0) continue;
        const int* ct = s_ctab + rr * 12 - CPAD;     // indexed by position in the control tiles; positions that are
                                                     //   no control column read a neighbour's entry and are masked by vmask
        unsigned m = vmask;
        int cpk[NCT][2];                             // packed (row << 5 | local column) of the lane's columns
#pragma unroll
        for (int jj = 0; jj < NCT; ++jj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                cpk[jj][e] = ct[8 * jj + 2 * fk + e];
                if (cpk[jj][e] < 0) m &= ~k2_colbits<NT, TC>(jj, e);
            }
        int apk[NCT];                                // the same for the lane's rows in the control tiles
#pragma unroll
        for (int ii = 0; ii < NCT; ++ii) {
            apk[ii] = ct[8 * ii + fm];
            if (!(ii == 0 && rW0) && apk[ii] < 0) m &= ~k2_rowbits<NT, TC>(ii);
        }
        int k = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
#pragma unroll
            for (int j = i; j < NT; ++j, ++k) {
                if (j < TC) continue;                // camera-only tile pair: stays in registers
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int jj = j - TC;
                    unsigned em = sbits >> (4 * (2 * jj + e));
                    if (i >= TC) em |= sbits >> (16 + 4 * (i - TC));
                    const bool go = (em & lm & 0xFu) != 0;        // one of the entry's slots leaves
                    const double v = acc[k][e];
                    if (go) acc[k][e] = 0.0;                       // (camera-only entries have no slot: kept)
                    const bool pred = go && ((m >> (2 * k + e)) & 1u) && v != 0.0;
                    const int pb = cpk[jj][e];
                    if (i < TC) {                    // camera rows only
                        k2_red(Wc0 + 8 * i + (int64_t)(pb >> 5) * ldw, v, pred);
                    } else {
                        const int ii = i - TC;
                        const bool rw = ii == 0 && rW0;
                        const int pa = apk[ii];
                        // control x control: (row, column) ordered by control point (slots rotate)
                        const int pr = pa <= pb ? pa : pb, pc = pa <= pb ? pb : pa;
                        const bool same = (pr >> 5) - (pr & 31) == (pc >> 5) - (pc & 31);
                        double* ptr = rw ? Wr0 + (int64_t)(pb >> 5) * ldw
                                         : (same ? D : E) + (int64_t)(pr >> 5) * q + (pc & 31);
                        const double val = ii == 0 ? __hiloint2double(__double2hiint(v) ^ flip0, __double2loint(v)) : v;
                        k2_red(ptr, val, pred);
                    }
                }
            }
        }
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
