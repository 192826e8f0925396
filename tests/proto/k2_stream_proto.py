"""NumPy model of K2's bookkeeping, round 2 (TEST INFRASTRUCTURE): mvus_b200/csrc/ba_k2.cuh restated
lane by lane -- the 32-detection Jacobian blocks with their XOR swizzle (ba_ctx.cuh, jblk_off), the
column layout  X | Y | Z | residual | camera, the fragment-load table (phase, row kind (o, h),
column) -> (plane offset, swizzle) with masked entries reading a row of zeros, the ping-pong phases, the per-lane flush masks and the
flush addresses (control x control -> upper triangle of D or E ordered by global row, camera x control
-> W~, residual x control -> -last column, camera-only entries once per chunk; control x control
entries first go to the band array Hb[row][12] and band_to_blocks moves them to D / E).  The MMA itself is a
plain outer-product sum over the fragment values.  tests/test_host.py compares the result with the
dense J^T J of the oracle-checked Jacobian, for time-ordered and scrambled detection orders."""
import numpy as np

TILE, CHUNK_TILES, BAND, HALF, GUARD = 128, 4, 23, 11, 9
GX, GY, GZ, GR, GC, GPAD = range(6)


def jblk_off(p, t):
    return p * 32 + (t ^ ((p & 3) << 2))


class Cfg:
    def __init__(self, P):
        self.P, self.PC = P, P - 12
        self.NCTRL, self.CR, self.CC = 21, 21, 22
        self.NCOL = self.CC + self.PC
        self.NT = (self.NCOL + 7) // 8
        self.TM = 2
        self.NPL = 2 * P + 2
        self.BLK_D = self.NPL * 32 + 16

    def group(self, c):
        return GX if c < 9 else GY if c < 12 else GZ if c < 21 else GR if c == 21 else GC if c < self.NCOL else GPAD


def wpos(c, ph):
    return c // 3 + 4 * ph if c < 9 else 3 if c < 12 else (c - 12) // 3 + 4 * (1 - ph)


def axis(c):
    return c % 3 if c < 12 else (c - 12) % 3


def write_blocks(cfg, Ju, Jv, ru, rv, span):
    """K1's output for one camera: list of tiles, each 4 blocks (float arrays of BLK_D; spans kept as a
    separate int array per block for readability)."""
    n = len(span)
    tiles = []
    for t0 in range(0, n, TILE):
        cnt = min(TILE, n - t0)
        blocks = []
        for sub in range(4):
            blk = np.full(cfg.NPL * 32, np.nan)          # lanes past the end stay garbage (never read)
            sp = np.full(32, -12345, dtype=np.int64)
            for lt in range(32):
                d = t0 + 32 * sub + lt
                if 32 * sub + lt >= cnt:
                    continue
                for p in range(cfg.P):
                    blk[jblk_off(p, lt)] = Ju[d, p]
                    blk[jblk_off(cfg.P + p, lt)] = Jv[d, p]
                blk[jblk_off(2 * cfg.P, lt)] = ru[d]
                blk[jblk_off(2 * cfg.P + 1, lt)] = rv[d]
                sp[lt] = span[d]
            blocks.append((blk, sp))
        tiles.append((cnt, blocks))
    return tiles


class Warp:
    """One streaming warp of accumulate_kernel<P>."""

    def __init__(self, cfg, bw, nb, nc, out):
        self.cfg, self.bw, self.q = cfg, bw, 3 * bw
        self.ldw = nc * cfg.PC + 1
        self.n_rows = (nb + 1) * self.q
        self.A, self.bc, self.D, self.E, self.W = out
        NT = cfg.NT
        self.lanes = [(lane >> 2, lane & 3) for lane in range(32)]      # (fm, fk)
        # masks over accumulator registers (pair index k, e)
        self.pairs = [(i, j) for i in range(NT) for j in range(i, NT)]
        self.mX = np.zeros((32, len(self.pairs), 2), bool)
        self.mY = np.zeros_like(self.mX); self.mZ = np.zeros_like(self.mX); self.mK = np.zeros_like(self.mX)
        for lane, (fm, fk) in enumerate(self.lanes):
            for k, (i, j) in enumerate(self.pairs):
                for e in range(2):
                    a, b = 8 * i + fm, 8 * j + 2 * fk + e
                    if a > b or b >= cfg.NCOL:
                        continue
                    ga, gb = cfg.group(a), cfg.group(b)
                    self.mX[lane, k, e] = ga == GX or gb == GX
                    self.mY[lane, k, e] = ga == GY or gb == GY
                    self.mZ[lane, k, e] = ga == GZ or gb == GZ
                    self.mK[lane, k, e] = ga >= GR and gb == GC
        self.acc = np.zeros((32, len(self.pairs), 2))
        # fragment-load table (phase, row kind, column) -> (plane offset, swizzle); masked -> zero row
        self.tab = [[[None] * (NT * 8) for _ in range(9)] for _ in range(2)]
        for ph in (0, 1):
            for rr in range(9):
                for c in range(NT * 8):
                    p = -1
                    if rr < 8:
                        o, h = rr >> 1, rr & 1
                        g = cfg.group(c)
                        if g <= GZ:
                            mm = wpos(c, ph) - o
                            if 0 <= mm < 4:
                                p = cfg.PC + 3 * mm + axis(c) + h * cfg.P
                        elif g == GR:
                            p = 2 * cfg.P + h
                        elif g == GC:
                            p = c - cfg.CC + h * cfg.P
                    self.tab[ph][rr][c] = (p * 256, (p & 3) << 5) if p >= 0 else (cfg.BLK_D * 8, 0)   # byte offsets
        self.Hb = np.zeros((self.n_rows + 2 * GUARD) * BAND)
        self.Wg = np.zeros((self.n_rows + 2 * GUARD) * self.ldw)
        self.n_flush_entries = 0

    # ---- flush (lambda `flush` of the kernel): no bounds checks (guard rows), no ordering (two-sided band)
    def flush(self, m, blk, ph, cam):
        cfg = self.cfg
        R0 = 3 * (4 * blk - 3)
        camoff = cam * cfg.PC

        def cr(c):
            return 3 * wpos(c, ph) + axis(c) if c < cfg.NCTRL else 0
        for lane, (fm, fk) in enumerate(self.lanes):
            hrow, wrow, hcol = [], [], []
            for i in range(cfg.TM + 1):
                Rr = R0 + cr(8 * i + fm) + GUARD
                hrow.append(Rr * (BAND - 1) + HALF)
                wrow.append(Rr * self.ldw + camoff)
                hcol.append([R0 + cr(8 * i + 2 * fk + e) + GUARD for e in range(2)])
            for k, (i, j) in enumerate(self.pairs):
                if i > cfg.TM:
                    continue
                for e in range(2):
                    v = self.acc[lane, k, e]
                    if not m[lane, k, e]:
                        continue
                    self.acc[lane, k, e] = 0.0
                    if v == 0.0:
                        continue
                    self.n_flush_entries += 1
                    c = 8 * j + 2 * fk + e
                    if j < cfg.TM or c < cfg.NCTRL:
                        self.Hb[hrow[i] + hcol[j][e]] += v
                    elif c == cfg.CR:
                        self.Wg[wrow[i] - camoff + self.ldw - 1] += -v
                    else:
                        self.Wg[wrow[i] + (c - cfg.CC)] += v

    def band_to_blocks(self):
        q, n_rows = self.q, self.n_rows
        Hb = self.Hb[GUARD * BAND:]
        for R in range(n_rows):
            for d in range(HALF + 1):
                if R + d >= n_rows:
                    continue
                v = Hb[R * (BAND - 1) + R + d + HALF]
                if d > 0:
                    v += Hb[(R + d) * (BAND - 1) + R + HALF]
                if v == 0.0:
                    continue
                kb = R // q
                lh = R - kb * q + d
                if lh < q:
                    self.D.reshape(-1)[R * q + lh] += v
                else:
                    self.E.reshape(-1)[R * q + lh - q] += v
        assert np.abs(self.Wg[:GUARD * self.ldw]).max() == 0.0       # nothing non-zero lands in the guard rows
        self.W += self.Wg[GUARD * self.ldw:GUARD * self.ldw + self.W.size].reshape(self.W.shape)
        self.Hb[:] = 0.0
        self.Wg[:] = 0.0

    def flush_camera(self, cam):
        cfg = self.cfg
        for lane, (fm, fk) in enumerate(self.lanes):
            for k, (i, j) in enumerate(self.pairs):
                if j < cfg.TM:
                    continue
                for e in range(2):
                    if not self.mK[lane, k, e]:
                        continue
                    v = self.acc[lane, k, e]
                    self.acc[lane, k, e] = 0.0
                    if v == 0.0:
                        continue
                    a, b = 8 * i + fm, 8 * j + 2 * fk + e
                    if a == cfg.CR:
                        self.bc[cam, b - cfg.CC] += -v
                    else:
                        self.A[cam, a - cfg.CC, b - cfg.CC] += v
                        if a != b:
                            self.A[cam, b - cfg.CC, a - cfg.CC] += v

    # ---- one chunk (body of the kernel's for(;;) loop)
    def run_chunk(self, tiles, cam):
        cfg = self.cfg
        NT = cfg.NT
        cur_b, ph = -1, 0
        for cnt, blocks in tiles:
            for sub in range(4):
                nvalid = min(32, cnt - 32 * sub)
                if nvalid <= 0:
                    continue
                stage, sp = blocks[sub]
                stage = np.concatenate((stage, np.full(16, np.nan), np.zeros(32)))   # span ints, then the row of zeros
                g = np.array([sp[l] if l < nvalid else -1 for l in range(32)])
                cov = g >= 0
                b_eff = np.zeros(32, int)
                for l in range(32):
                    prev = [k for k in range(l + 1) if cov[k]]
                    b_eff[l] = (g[prev[-1]] >> 2) if prev else cur_b
                o_reg = np.where(cov, g & 3, 0)
                b_prev = np.concatenate(([cur_b], b_eff[:-1]))
                heads = [(l < nvalid) and (b_eff[l] != b_prev[l]) for l in range(32)]
                start = 0
                while start < nvalid:
                    rest = [l for l in range(start + 1, 32) if heads[l]]
                    end = min(rest[0], nvalid) if rest else nvalid
                    bseg = b_eff[start]
                    if bseg != cur_b:
                        if cur_b >= 0:
                            partial = bseg == cur_b + 1
                            m = (self.mY | (self.mZ if ph else self.mX)) if partial else (self.mX | self.mY | self.mZ)
                            self.flush(m, cur_b, ph, cam)
                            ph = (ph ^ 1) if partial else 0
                        else:
                            ph = 0
                        cur_b = bseg
                    if bseg >= 0:
                        n = end - start
                        n2 = 2 * n
                        for s in range((n2 + 3) >> 2):
                            f = np.zeros((32, NT))
                            for lane, (fm, fk) in enumerate(self.lanes):
                                rho = 4 * s + fk
                                h = 1 if rho >= n else 0
                                t = (start + rho - (n if h else 0)) & 31
                                row = (2 * o_reg[t] + h) if rho < n2 else 8
                                for i in range(NT):
                                    off, sw = self.tab[ph][row][8 * i + fm]
                                    f[lane, i] = stage[(off + ((t * 8) ^ sw)) // 8]
                            # mma.m8n8k4: C[fm][2 fk + e] += sum_k A[k][fm] * B[k][2 fk + e]
                            for k, (i, j) in enumerate(self.pairs):
                                Ai = f[:, i].reshape(8, 4)           # [fm][k]
                                Bj = f[:, j].reshape(8, 4)           # [n][k]
                                C = Ai @ Bj.T                        # [fm][n]
                                for lane, (fm, fk) in enumerate(self.lanes):
                                    self.acc[lane, k, 0] += C[fm, 2 * fk]
                                    self.acc[lane, k, 1] += C[fm, 2 * fk + 1]
                    start = end
        if cur_b >= 0:
            self.flush(self.mX | self.mY | self.mZ, cur_b, ph, cam)
        self.flush_camera(cam)


def accumulate_camera(Ju, Jv, ru, rv, span, Pc, cam, nc, bw, nb, out=None):
    """Same interface as k2_window_proto.accumulate_camera."""
    cfg = Cfg(Pc + 12)
    q = 3 * bw
    ldw = nc * Pc + 1
    if out is None:
        out = (np.zeros((nc, Pc, Pc)), np.zeros((nc, Pc)), np.zeros((nb + 1, q, q)), np.zeros((nb + 1, q, q)),
               np.zeros(((nb + 1) * q, ldw)))
    tiles = write_blocks(cfg, Ju, Jv, ru, rv, span)
    w = Warp(cfg, bw, nb, nc, out)
    for c0 in range(0, len(tiles), CHUNK_TILES):
        w.run_chunk(tiles[c0:c0 + CHUNK_TILES], cam)
    w.band_to_blocks()
    return out
