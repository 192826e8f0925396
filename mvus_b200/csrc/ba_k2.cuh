// K2 (round 2): J^T J / J^T r accumulation by STREAMING WARPS.
//
// Data.  K1 writes the compact block-row Jacobian in 32-detection BLOCKS (one per warp of a K1 tile):
//   block = [2P+2 planes x 32 doubles | 32 span ints], planes = u columns (P), v columns (P), r_u, r_v,
//   element (plane p, detection t) at p*32 + (t ^ ((p & 3) << 2))   (jblk_off, ba_ctx.cuh)
// -- 11 392 B (P = 21) / 16 000 B (P = 30), contiguous and 16-byte aligned, so a block is ONE
// cp.async.bulk (TMA 1-D) into shared memory, completing on the warp's own mbarrier.  The XOR swizzle
// keeps K1's stores coalesced (it permutes 32-byte sectors inside a 256-byte plane row) and makes the
// MMA fragment loads below bank-conflict free (lanes of one quarter-warp read 4 consecutive planes x
// 4 consecutive detections).
//
// Work.  A warp owns a CHUNK = up to 4 consecutive K1 tiles (512 detections) of one camera, fetched from
// a global queue in TIME order across cameras (chunk_key_kernel + CUB sort: concurrently running warps
// then RED into the same few thousand control points and D / E / W~ stay in L2).  No CTA-wide phase, no
// __syncthreads after start-up: every warp is its own load -> MMA -> flush pipeline.
//
// Arithmetic.  Detections are grouped into BLOCK-RUNS: consecutive detections whose knot span g lies in
// the same block of four spans b = g >> 2 (about 20 detections at config 4).  All of them touch only
// the 7 control points 4b-3 .. 4b+3, so the run's update  S = sum_t (u_t u_t^T + v_t v_t^T)  is one
// symmetric rank-2n update over the columns
//      [ X: 3 points | Y: 1 point | Z: 3 points | residual | camera 0..PC-1 ]      (31 / 40 columns)
// done with mma.sync.m8n8k4.f64 fragments (NT = 4 / 5 column tiles, every tile pair (i <= j) one 8x8
// accumulator; u rows then v rows of the run packed along k, 4 rows per MMA step).  A detection with
// span g = 4b + o contributes its 4 active control points to window positions o .. o+3, the rest are
// zeros (the fragment load maps column -> plane  PC + 3 (w - o) + axis  and predicates it).
// PING-PONG: in phase A the window is X | Y | Z = points (4b-3..4b-1 | 4b | 4b+1..4b+3); when the next
// run is block b+1 only the entries that involve X or Y are flushed (they are complete for this camera),
// the Z entries stay in their registers and Z becomes the OLD half of the next window (phase B:
// Z | Y | X = 4b+1..4b+3 | 4b+4 | 4b+5..4b+7), and vice versa.  So every (camera, control point) entry of
// W~ and every control-point pair of D / E is flushed ONCE per camera sweep with lane-uniform code (no
// slot tables); any other transition (gap, span going backwards, end of chunk) flushes everything, which
// keeps the kernel correct for arbitrary span sequences.  Flushes are FP64 REDs: camera x control -> W~
// (residual x control -> its last column, negated), control x control -> the upper triangle of D or E
// ordered by global row.  Camera-only entries stay in registers for the whole chunk.
//
// HBM traffic per detection (algorithmic): r (16 B) + span (4 B) + J (16 P B)
//   -> 356 B (P = 21) / 500 B (P = 30); writes are O(runs).
// tests/proto/k2_stream_proto.py is the NumPy model of exactly this bookkeeping (lane constants, plane
// mapping, masks, flush addresses), checked against the dense J^T J on the CPU.
#pragma once
#include <cub/cub.cuh>
#include "ba_ctx.cuh"

namespace mvus {

constexpr int K2_CHUNK_TILES = 4;     // tiles (of TILE_DET detections) a warp streams per chunk

constexpr int K2_BAND = 23;           // control x control entries of reprojection rows: |row difference| <= 11; the band
                                      //   array holds entry (Ra, Rb) at Ra*(K2_BAND-1) + Rb + K2_HALF on BOTH sides of the diagonal
constexpr int K2_HALF = 11;
constexpr int K2_GUARD = 9;           // rows in front of the band array and of W~: the window of block 0 starts at point -3

template <int P>
struct K2Cfg {
    static constexpr int PC = P - 12;                 // camera unknowns (9 or 18)
    static constexpr int NCTRL = 21;                  // X (9) | Y (3) | Z (9)
    static constexpr int CR = NCTRL;                  // residual column
    static constexpr int CC = NCTRL + 1;              // first camera column
    static constexpr int NCOL = CC + PC;              // 31 / 40
    static constexpr int NT = (NCOL + 7) / 8;         // 4 / 5 column tiles
    static constexpr int NPAIR = NT * (NT + 1) / 2;
    static constexpr int TM = 2;                      // the one tile that mixes control, residual and camera columns
    static constexpr int NPL = JBlk<P>::NPL;
    static constexpr int BLK_D = JBlk<P>::BLK_D;
    static constexpr int BLK_BYTES = BLK_D * 8;
    static constexpr int ZERO_D = BLK_D;              // a row of 32 zeros behind the block: the target of every masked load
    static constexpr int RO_D = BLK_D + 32;           // 32 ints: load-table row (byte offset) of each detection of the block
    static constexpr int STAGE_D = BLK_D + 32 + 16;
    static constexpr int WARPS = P == 21 ? 16 : 12;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int ROWS = 9;                    // load-table rows: (o, h) = 4 x 2, + "no row"
    static constexpr int RS = NT * 8 + 4;             // entries per table row (== 4 or 12 mod 16: rows of different o
                                                      //   fall into different shared-memory banks)
    static constexpr int TAB = 2 * ROWS * RS;         // entries (int2), both phases
    static constexpr size_t SMEM = (size_t)WARPS * STAGE_D * 8 + (size_t)TAB * 8 + WARPS * 8;
    static_assert(NCTRL > 8 * TM && NCTRL < 8 * (TM + 1), "control columns must end inside tile TM");
    static_assert(2 * NPAIR <= 32, "accumulator registers are indexed by a 32-bit mask");
    static_assert(BLK_BYTES % 16 == 0 && (STAGE_D * 8) % 16 == 0, "bulk copies move multiples of 16 bytes");
};

enum { K2_GX = 0, K2_GY = 1, K2_GZ = 2, K2_GR = 3, K2_GC = 4, K2_GPAD = 5 };

template <int P>
__host__ __device__ __forceinline__ int k2_group(int c) {
    return c < 9 ? K2_GX : c < 12 ? K2_GY : c < 21 ? K2_GZ : c == 21 ? K2_GR : c < K2Cfg<P>::NCOL ? K2_GC : K2_GPAD;
}
// window position (0..6) of control column c in phase ph (0 = A: X old, 1 = B: Z old), and its axis
__host__ __device__ __forceinline__ int k2_wpos(int c, int ph) {
    return c < 9 ? c / 3 + 4 * ph : c < 12 ? 3 : (c - 12) / 3 + 4 * (1 - ph);
}
__host__ __device__ __forceinline__ int k2_axis(int c) { return c < 12 ? c % 3 : (c - 12) % 3; }

__device__ __forceinline__ void k2_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// Predicated FP64 reduction (a flush is ~20 of these with lane-dependent predicates, no branches).
__device__ __forceinline__ void k2_red(double* ptr, double v, bool pred) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\t@p red.global.add.f64 [%0], %1;\n\t}"
                 :: "l"(ptr), "d"(v), "r"((int)pred) : "memory");
}

__device__ __forceinline__ double k2_lds_f64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int2 k2_lds_v2(unsigned addr) {
    int2 v;
    asm volatile("ld.shared.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ int k2_lds_s32(unsigned addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void k2_mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void k2_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void k2_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tK2_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra K2_DONE;\n\tbra K2_WAIT;\n\tK2_DONE:\n\t}" :: "r"(bar), "r"(parity) : "memory");
}

// Chunk order of K2: by the span index of the chunk's first detection, i.e. by TIME across all cameras.
__global__ void chunk_key_kernel(const double* __restrict__ Jb, int blk_d, int span_off, const int* __restrict__ chunk_tile0,
                                 int n_chunks, int* __restrict__ key, int* __restrict__ id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_chunks) return;
    const int* sp = reinterpret_cast<const int*>(Jb + (size_t)chunk_tile0[i] * 4 * blk_d + span_off);
    key[i] = sp[0] + 1;          // uncovered (-1) first
    id[i] = i;
}

// K2 leaves the control x control entries in a BAND array Hb (entry (Ra, Rb), |Ra - Rb| <= 11, at
// Ra*(K2_BAND-1) + Rb + K2_HALF, whichever of the two rows the lane held as its fragment row: no ordering, no
// super-block arithmetic inside the flush).  This kernel folds the two sides and moves the entries into the
// solver's block-tridiagonal form: D (upper triangle of the diagonal blocks) and E (coupling to the next
// block), both zeroed before.  Hb points at row 0; K2_GUARD rows in front of it absorb the window of block 0.
__global__ void band_to_blocks_kernel(const double* __restrict__ Hb, int64_t n_rows, int q,
                                      double* __restrict__ D, double* __restrict__ E) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rows * (K2_HALF + 1)) return;
    const int64_t R = i / (K2_HALF + 1);
    const int d = (int)(i - R * (K2_HALF + 1));
    if (R + d >= n_rows) return;
    double v = Hb[R * (K2_BAND - 1) + (R + d) + K2_HALF];
    if (d > 0) v += Hb[(R + d) * (K2_BAND - 1) + R + K2_HALF];
    if (v == 0.0) return;
    const int64_t kb = R / q;
    const int lh = (int)(R - kb * q) + d;
    if (lh < q) D[R * q + lh] = v;
    else E[R * q + (lh - q)] = v;
}

template <int P>
__global__ void __launch_bounds__(K2Cfg<P>::THREADS, 1)
accumulate_kernel(const double* __restrict__ Jb, const int* __restrict__ chunk_perm, const int* __restrict__ chunk_tile0,
                  const int* __restrict__ chunk_nt, int n_chunks, int* __restrict__ queue,
                  const int* __restrict__ tile_cam, const int* __restrict__ tile_cnt,
                  int ldw, double* __restrict__ A, double* __restrict__ bc,
                  double* __restrict__ Hb, double* __restrict__ W) {
    using Cfg = K2Cfg<P>;
    constexpr int NT = Cfg::NT, PC = Cfg::PC, TM = Cfg::TM, CR = Cfg::CR, CC = Cfg::CC, NCOL = Cfg::NCOL;
    constexpr int RS = Cfg::RS, ROWS = Cfg::ROWS;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char k2_smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int fk = lane & 3, fm = lane >> 2;
    double* stage = reinterpret_cast<double*>(k2_smem) + (size_t)warp * Cfg::STAGE_D;
    const int* stage_span = reinterpret_cast<const int*>(stage + Cfg::NPL * 32);
    int2* tab = reinterpret_cast<int2*>(k2_smem + (size_t)Cfg::WARPS * Cfg::STAGE_D * 8);
    const unsigned stage_u = (unsigned)__cvta_generic_to_shared(stage);
    const unsigned tab_u = (unsigned)__cvta_generic_to_shared(tab) + fm * 8;
    const unsigned bar = (unsigned)__cvta_generic_to_shared(k2_smem + (size_t)Cfg::WARPS * Cfg::STAGE_D * 8 +
                                                            (size_t)Cfg::TAB * 8 + warp * 8);
    // ---- fragment-load table: (phase, row kind (o, h) or "none", tile, fm) -> (BYTE offset of the plane in the
    //      stage, XOR swizzle of that plane in bytes); masked entries point at the row of zeros behind the block
    for (int x = tid; x < Cfg::TAB; x += Cfg::THREADS) {
        const int ph = x / (ROWS * RS), rr = (x / RS) % ROWS, e = x % RS;
        int p = -1;
        if (rr < 8 && e < NT * 8) {
            const int o = rr >> 1, h = rr & 1, c = e, g = k2_group<P>(c);
            if (g <= K2_GZ) {
                const int m = k2_wpos(c, ph) - o;
                if (m >= 0 && m < 4) p = PC + 3 * m + k2_axis(c) + h * P;
            } else if (g == K2_GR) p = 2 * P + h;
            else if (g == K2_GC) p = c - CC + h * P;
        }
        tab[x] = p >= 0 ? make_int2(p * 256, (p & 3) << 5) : make_int2(Cfg::ZERO_D * 8, 0);
    }
    if (lane == 0) k2_mbar_init(bar, 1);
    stage[Cfg::ZERO_D + lane] = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    unsigned parity = 0;

    // ---- lane constants of the flush.  Row slot i = column 8 i + fm, column slot (j, e) = column 8 j + 2 fk + e;
    //      crX[ph] = 3 w + axis of a control column (its global row is 3 (4 b - 3) + that); 0 for the others
    //      (their entries are never selected together with a control partner).
    int crA[TM + 1], crB[TM + 1], ccA[TM + 1][2], ccB[TM + 1][2];
#pragma unroll
    for (int i = 0; i <= TM; ++i) {
        const int c = 8 * i + fm;
        const bool ct = c < Cfg::NCTRL;
        crA[i] = ct ? 3 * k2_wpos(c, 0) + k2_axis(c) : 0;
        crB[i] = ct ? 3 * k2_wpos(c, 1) + k2_axis(c) : 0;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int cb = 8 * i + 2 * fk + e;
            const bool cbt = cb < Cfg::NCTRL;
            ccA[i][e] = cbt ? 3 * k2_wpos(cb, 0) + k2_axis(cb) : 0;
            ccB[i][e] = cbt ? 3 * k2_wpos(cb, 1) + k2_axis(cb) : 0;
        }
    }
    // which accumulator registers (bit 2 k + e, k = index of tile pair (i <= j)) hold an entry that involves
    // group X / Y / Z, and which hold camera-only entries (kept for the whole chunk)
    unsigned mX = 0, mY = 0, mZ = 0, mK = 0;
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = i; j < NT; ++j, ++k)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int a = 8 * i + fm, b = 8 * j + 2 * fk + e;
                    if (a > b || b >= NCOL) continue;
                    const int ga = k2_group<P>(a), gb = k2_group<P>(b);
                    const unsigned bit = 1u << (2 * k + e);
                    if (ga == K2_GX || gb == K2_GX) mX |= bit;
                    if (ga == K2_GY || gb == K2_GY) mY |= bit;
                    if (ga == K2_GZ || gb == K2_GZ) mZ |= bit;
                    if (ga >= K2_GR && gb == K2_GC) mK |= bit;      // residual x camera, camera x camera
                }
    }

    double acc[Cfg::NPAIR][2];
#pragma unroll
    for (int k = 0; k < Cfg::NPAIR; ++k) { acc[k][0] = 0.0; acc[k][1] = 0.0; }

    int camoff = 0;      // cam * PC of the current chunk
    int b_min = 0x7fffffff, b_max = -1;      // blocks of four spans this warp has seen (multi-GPU: the touched rows)
    // ---- flush of the entries selected by `m` (bits as above) for the window of block `blk` in phase `ph`.
    //      Rows 3 (4 blk - 3) + cr >= -K2_GUARD always, and <= n_rows - 1 for every span K1 can emit, so there are
    //      no bounds checks: an entry is its bit, a != 0 test, one add for the offset and the RED.
    auto flush = [&](unsigned m, int blk, int ph) {
        const int R0 = 3 * (4 * blk - 3);               // global row of window position 0, axis 0 (>= -9)
        unsigned hrow[TM + 1], wrow[TM + 1], hcol[TM + 1][2];
#pragma unroll
        for (int i = 0; i <= TM; ++i) {
            const int Rr = R0 + (ph ? crB[i] : crA[i]) + K2_GUARD;
            hrow[i] = (unsigned)(Rr * (K2_BAND - 1) + K2_HALF);
            wrow[i] = (unsigned)(Rr * ldw + camoff);
#pragma unroll
            for (int e = 0; e < 2; ++e) hcol[i][e] = (unsigned)(R0 + (ph ? ccB[i][e] : ccA[i][e]) + K2_GUARD);
        }
        int k = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i) {
#pragma unroll
            for (int j = i; j < NT; ++j, ++k) {
                if (i > TM) continue;                     // camera-only tile pairs: nothing to flush here
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double v = acc[k][e];
                    if (!((m >> (2 * k + e)) & 1u)) continue;
                    acc[k][e] = 0.0;
                    if (v == 0.0) continue;
                    const int c = 8 * j + 2 * fk + e;
                    if (j < TM || c < Cfg::NCTRL) atomicAdd(Hb + (size_t)(hrow[i] + hcol[j][e]), v);     // control x control
                    else if (c == CR) atomicAdd(W + (size_t)(wrow[i] - camoff + (ldw - 1)), -v);     // -> -J^T r
                    else atomicAdd(W + (size_t)(wrow[i] + (c - CC)), v);                             // camera x control
                }
            }
        }
    };
    // ---- camera-only entries: once per chunk
    auto flush_camera = [&]() {
        int k = 0;
#pragma unroll
        for (int i = 0; i < NT; ++i)
#pragma unroll
            for (int j = i; j < NT; ++j, ++k) {
                if (j < TM) continue;
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    if (!((mK >> (2 * k + e)) & 1u)) continue;
                    const double v = acc[k][e];
                    acc[k][e] = 0.0;
                    if (v == 0.0) continue;
                    const int a = 8 * i + fm, b = 8 * j + 2 * fk + e;
                    if (a == CR) atomicAdd(bc + camoff + (b - CC), -v);
                    else {
                        atomicAdd(A + ((int64_t)camoff + (a - CC)) * PC + (b - CC), v);
                        if (a != b) atomicAdd(A + ((int64_t)camoff + (b - CC)) * PC + (a - CC), v);
                    }
                }
            }
    };

    for (;;) {
        int ci = 0;
        if (lane == 0) ci = atomicAdd(queue, 1);
        ci = __shfl_sync(FULL, ci, 0);
        if (ci >= n_chunks) break;
        const int ch = chunk_perm[ci];
        const int tile0 = chunk_tile0[ch], ntile = chunk_nt[ch];
        camoff = tile_cam[tile0] * PC;
        int cur_b = -1, ph = 0;

        for (int sb = 0; sb < 4 * ntile; ++sb) {
            const int tl = tile0 + (sb >> 2);
            const int nvalid = min(32, tile_cnt[tl] - 32 * (sb & 3));
            if (nvalid <= 0) continue;
            __syncwarp();                                  // every lane is done with the previous block
            if (lane == 0)
                k2_bulk_load(stage_u, Jb + ((size_t)tl * 4 + (sb & 3)) * Cfg::BLK_D, Cfg::BLK_BYTES, bar);
            k2_mbar_wait(bar, parity);
            parity ^= 1u;
            // ---- block-runs inside this block: uncovered detections (span -1, all-zero rows) join the run
            //      they follow, so they never force a flush
            const int g = lane < nvalid ? stage_span[lane] : -1;
            const unsigned cov = __ballot_sync(FULL, g >= 0);
            const unsigned prevm = cov & (FULL >> (31 - lane));
            const int src = prevm ? 31 - __clz(prevm) : 0;
            const int b_src = __shfl_sync(FULL, g >> 2, src);
            const int b_eff = prevm ? b_src : cur_b;
            reinterpret_cast<int*>(stage + Cfg::RO_D)[lane] = g >= 0 ? 2 * (g & 3) * RS * 8 : 0;
            __syncwarp();
            int b_prev = __shfl_up_sync(FULL, b_eff, 1);
            if (lane == 0) b_prev = cur_b;
            const unsigned heads = __ballot_sync(FULL, lane < nvalid && b_eff != b_prev);
            int start = 0;
            while (start < nvalid) {
                const unsigned rest = heads & ~((2u << start) - 1u);
                const int end = rest ? min(__ffs(rest) - 1, nvalid) : nvalid;
                const int bseg = __shfl_sync(FULL, b_eff, start);
                if (bseg != cur_b) {
                    if (cur_b >= 0) {
                        const bool partial = bseg == cur_b + 1;
                        flush(partial ? (mY | (ph ? mZ : mX)) : (mX | mY | mZ), cur_b, ph);
                        ph = partial ? (ph ^ 1) : 0;
                    } else ph = 0;
                    cur_b = bseg;
                    if (bseg >= 0) { b_min = min(b_min, bseg); b_max = max(b_max, bseg); }
                }
                if (bseg >= 0) {
                    // ---- rank-2n update of the segment [start, end): u rows then v rows along k
                    const int n = end - start, n2 = 2 * n, nsteps = (n2 + 3) >> 2;
                    const unsigned tph = tab_u + ph * (ROWS * RS * 8);
                    double fa[NT], fb[NT];
                    auto load = [&](int s, double* f) {
                        const int rho = 4 * s + fk;
                        const int h = rho >= n ? 1 : 0;
                        const int t = (start + rho - (h ? n : 0)) & 31;
                        const int ro = k2_lds_s32(stage_u + Cfg::RO_D * 8 + t * 4);
                        const unsigned tr = tph + (rho < n2 ? ro + h * (RS * 8) : 8 * RS * 8);
                        const unsigned t8 = stage_u + t * 8;         // (the swizzle only touches bits 5..6 of t * 8)
#pragma unroll
                        for (int i = 0; i < NT; ++i) {
                            const int2 e = k2_lds_v2(tr + i * 64);
                            f[i] = k2_lds_f64(e.x + (t8 ^ e.y));
                        }
                    };
                    auto mma = [&](const double* f) {
                        int k = 0;
#pragma unroll
                        for (int i = 0; i < NT; ++i)
#pragma unroll
                            for (int j = i; j < NT; ++j, ++k) k2_dmma(acc[k][0], acc[k][1], f[i], f[j]);
                    };
                    load(0, fa);
#pragma unroll 1
                    for (int s = 0; s < nsteps; s += 2) {
                        load(s + 1, fb);               // (rows past the segment read the zero row)
                        mma(fa);
                        if (s + 1 < nsteps) {
                            load(s + 2, fa);
                            mma(fb);
                        }
                    }
                }
                start = end;
            }
        }
        // ---- end of the chunk: everything leaves
        if (cur_b >= 0) flush(mX | mY | mZ, cur_b, ph);
        flush_camera();
    }
    if (lane == 0 && b_max >= 0) { atomicMin(queue + 1, b_min); atomicMax(queue + 2, b_max); }
}

}  // namespace mvus
