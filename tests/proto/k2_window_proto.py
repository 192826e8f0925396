"""NumPy model of K2's bookkeeping (TEST INFRASTRUCTURE): the sliding window over control-point
slots that mvus_b200/csrc/ba_k2.cuh uses to accumulate J^T J / J^T r for one camera.

Per detection the compact block row has Pc camera columns, 12 control columns (window position
m = 0..3 of the control points g-3 .. g, three axes each) and the residual.  The kernel keeps, per
warp, one symmetric accumulator over the packed columns

    [camera 0..Pc-1 | residual | slot 0 (3 axes) | slot 1 | slot 2 | slot 3]

where control point j always lives in slot j & 3.  A warp owns a chunk of CHUNK consecutive
detections of a tile; runs (equal span index, cut at chunk ends) are processed in order and after
each run only the slots whose control point differs in the next run are flushed: camera x slot
entries into W~ (residual x slot with a minus sign into its last column), slot x slot entries into
the upper triangle of D (same super-block) or into E (next super-block), ordered by control point
because the slots rotate.  This file restates exactly that and nothing else; the test compares it
with the dense J^T J."""
import numpy as np

TILE, CHUNK = 128, 16


def slot_cp(g, sl):
    """control point held by slot sl for a run of span g (ba_k2.cuh: j = g - ((g - sl) & 3))"""
    return g - ((g - sl) & 3)


def leave_mask(g, g_next):
    """slots to flush after a run of span g when the next run of the same chunk has span g_next
    (None: end of chunk / uncovered detections follow -> everything)"""
    if g_next is None or g_next < 0:
        return 0xF
    return sum(1 << sl for sl in range(4) if slot_cp(g, sl) != slot_cp(g_next, sl))


def accumulate_camera(Ju, Jv, ru, rv, span, Pc, cam, nc, bw, nb, out=None):
    """Ju, Jv: (n, Pc + 12) block rows of one camera's detections (time order), ru/rv residuals,
    span: (n,) last active control point (-1 = uncovered).  Adds into out = (A, bc, D, E, W) with
    A [nc,Pc,Pc], bc [nc,Pc], D/E [nb,q,q] (D upper triangle only), W [nb*q, nc*Pc+1]."""
    q = 3 * bw
    ldw = nc * Pc + 1
    if out is None:
        out = (np.zeros((nc, Pc, Pc)), np.zeros((nc, Pc)), np.zeros((nb, q, q)), np.zeros((nb, q, q)),
               np.zeros((nb * q, ldw)))
    A, bc, D, E, W = out
    n = len(span)
    NS = Pc + 1 + 12

    def row_of(j, ax):                       # global row of control point j, axis ax
        kb = j // bw
        return kb * q + (j - kb * bw) * 3 + ax

    for t0 in range(0, n, TILE):
        cnt = min(TILE, n - t0)
        keep = np.zeros((Pc + 1, Pc + 1))    # camera-only entries: once per tile
        for c0 in range(0, cnt, CHUNK):      # one warp per chunk
            c1 = min(c0 + CHUNK, cnt)
            heads = [c0] + [t for t in range(c0 + 1, c1) if span[t0 + t] != span[t0 + t - 1]]
            runs = [(heads[k], heads[k + 1] if k + 1 < len(heads) else c1) for k in range(len(heads))]
            acc = np.zeros((NS, NS))
            for k, (a, b) in enumerate(runs):
                g = int(span[t0 + a])
                if g < 0:
                    continue
                # packed rows of the run: slot sl takes window position m = (sl - g - 1) & 3
                for t in range(a, b):
                    for Jr, rr in ((Ju[t0 + t], ru[t0 + t]), (Jv[t0 + t], rv[t0 + t])):
                        v = np.zeros(NS)
                        v[:Pc] = Jr[:Pc]
                        v[Pc] = rr
                        for sl in range(4):
                            m = (sl - g - 1) & 3
                            v[Pc + 1 + 3 * sl:Pc + 4 + 3 * sl] = Jr[Pc + 3 * m:Pc + 3 * m + 3]
                        acc += np.outer(v, v)
                g_next = int(span[t0 + runs[k + 1][0]]) if k + 1 < len(runs) else None
                lm = leave_mask(g, g_next)
                for sl in range(4):
                    if not (lm >> sl) & 1:
                        continue
                    j1 = slot_cp(g, sl)
                    for ax in range(3):
                        x1 = Pc + 1 + 3 * sl + ax
                        if j1 >= 0:
                            r1 = row_of(j1, ax)
                            W[r1, cam * Pc:cam * Pc + Pc] += acc[:Pc, x1]
                            W[r1, ldw - 1] -= acc[Pc, x1]
                        acc[:Pc + 1, x1] = 0.0
                        for s2 in range(4):
                            j2 = slot_cp(g, s2)
                            for ay in range(3):
                                x2 = Pc + 1 + 3 * s2 + ay
                                lo, hi = min(x1, x2), max(x1, x2)      # the kernel keeps position (lo, hi)
                                val = acc[lo, hi]
                                acc[lo, hi] = 0.0
                                if lo != hi:
                                    acc[hi, lo] = 0.0                  # (mirror entry: never flushed, reset)
                                if val == 0.0 or j1 < 0 or j2 < 0:
                                    continue
                                ra, rb = row_of(j1, ax), row_of(j2, ay)
                                rlo, rhi = min(ra, rb), max(ra, rb)    # ordered by control point
                                if rlo // q == rhi // q:
                                    D[rlo // q, rlo % q, rhi % q] += val
                                else:
                                    assert rhi // q == rlo // q + 1
                                    E[rlo // q, rlo % q, rhi % q] += val
            keep += acc[:Pc + 1, :Pc + 1]
        A[cam] += keep[:Pc, :Pc]
        bc[cam] -= keep[:Pc, Pc]
    return out
