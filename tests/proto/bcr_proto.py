"""NumPy model of the device solver's linear algebra (TEST INFRASTRUCTURE): block cyclic
reduction over the super-blocked spline normal matrix + dense Schur complement on the camera
block.  mvus_b200/csrc/ba_solve.cuh mirrors this step for step; tests compare both with a
dense solve of the same damped normal equations."""
import numpy as np


def assemble(J, r, n_other, n_ctrl, ctrl_cols, bw):
    """J dense (m x n) in reference layout, ctrl_cols[j, ax] = column of control point j axis ax.
    Returns A (n_other x n_other), D [nb,q,q], E [nb,q,q], Wt [nb,q,n_other+1] with b=-g in the
    last column, bc (n_other)."""
    q = 3 * bw
    nb = (n_ctrl + bw - 1) // bw
    H = J.T @ J
    g = J.T @ r
    A = H[:n_other, :n_other].copy()
    bc = -g[:n_other].copy()
    perm = np.full(nb * q, -1)
    for j in range(n_ctrl):
        for ax in range(3):
            perm[j * 3 + ax] = ctrl_cols[j, ax]
    D = np.zeros((nb, q, q)); E = np.zeros((nb, q, q)); Wt = np.zeros((nb, q, n_other + 1))
    for k in range(nb):
        pk = perm[k * q:(k + 1) * q]
        ok = pk >= 0
        D[k][np.ix_(ok, ok)] = H[np.ix_(pk[ok], pk[ok])]
        D[k][~ok, ~ok] = 1.0
        Wt[k][ok, :n_other] = H[np.ix_(pk[ok], np.arange(n_other))]
        Wt[k][ok, n_other] = -g[pk[ok]]
        if k + 1 < nb:
            pn = perm[(k + 1) * q:(k + 2) * q]
            okn = pn >= 0
            E[k][np.ix_(ok, okn)] = H[np.ix_(pk[ok], pn[okn])]
    return A, bc, D, E, Wt, perm


def solve_bcr(A, bc, D, E, Wt, lam, dfloor=(1e-6, 1e32)):
    """Solve [[A+lam*dA, W^T],[W, B+lam*dB]] [dc; ds] = [bc; bs] by cyclic reduction."""
    nb, q, _ = D.shape
    ncp = A.shape[0]
    dA = np.clip(np.diag(A), *dfloor)
    Dw = D.copy(); Ew = E.copy(); Ww = Wt.copy()
    for k in range(nb):
        Dw[k] += lam * np.diag(np.clip(np.diag(D[k]), *dfloor))
    ZL = np.zeros_like(D)
    L = np.zeros_like(D)
    levels = []
    s = 1
    while s < nb:
        levels.append(s)
        s *= 2
    order = []
    for lev, s in enumerate(levels):
        if lev > 0:
            sp = s // 2     # apply pending updates from level lev-1 to all blocks multiple of s
            for j in range(0, nb, s):
                for nbk, Zn in ((j - sp, 'R'), (j + sp, 'L')):
                    if 0 <= nbk < nb and (nbk // sp) % 2 == 1:
                        Z = Ew[nbk] if Zn == 'R' else ZL[nbk]
                        Dw[j] -= Z.T @ Z
                        Ww[j] -= Z.T @ Ww[nbk]
                # new right coupling (j -> j+s) bridged by eliminated j+sp
                if j + s < nb:
                    Ew[j] = -ZL[j + sp].T @ EwR[j + sp]
        # eliminate odd multiples of s
        EwR = {}
        for k in range(s, nb, 2 * s):
            L[k] = np.linalg.cholesky(Dw[k])
            El = Ew[k - s].T                      # coupling rows k, cols k-s
            ZL[k] = np.linalg.solve(L[k], El)
            Er = Ew[k] if k + s < nb else np.zeros((q, q))
            Ew[k] = np.linalg.solve(L[k], Er)     # ZR stored in place
            EwR[k] = Ew[k]
            Ww[k] = np.linalg.solve(L[k], Ww[k])
            order.append((k, s))
    # pending updates for root block 0
    if levels:
        s = levels[-1]
        Z = ZL[s]
        Dw[0] -= Z.T @ Z
        Ww[0] -= Z.T @ Ww[s]
    L[0] = np.linalg.cholesky(Dw[0])
    Ww[0] = np.linalg.solve(L[0], Ww[0])
    order.append((0, 0))
    # camera system
    Sfull = np.zeros((ncp + 1, ncp + 1))
    Sfull[:ncp, :ncp] = A + lam * np.diag(dA)
    Sfull[:ncp, ncp] = bc
    Sfull[ncp, :ncp] = bc
    for k in range(nb):
        Sfull -= Ww[k].T @ Ww[k]
    S = Sfull[:ncp, :ncp]
    rhs = Sfull[:ncp, ncp]
    dc = np.linalg.solve(S, rhs)
    # back substitution
    ds = np.zeros((nb, q))
    for k, s in reversed(order):
        v = Ww[k][:, ncp] - Ww[k][:, :ncp] @ dc
        if s > 0:
            v -= ZL[k] @ ds[k - s]
            if k + s < nb:
                v -= Ew[k] @ ds[k + s]
        ds[k] = np.linalg.solve(L[k].T, v)
    return dc, ds


def prereduce(Dw, Ew, Ww, Lc):
    """Chunk pre-reduction (ba_solve.cuh: chunk_factor_kernel / chunk_w_kernel): blocks are cut into chunks
    of Lc; the first block of a chunk is its HEAD, the others are eliminated one after the other in ascending
    order.  Eliminated block k keeps L_k, ZR_k (coupling to k+1, or to the next head for the last block of the
    chunk) and ZH_k (fill-in coupling to its own head).  Returns the head system (Dt, Et, Wt) -- block
    tridiagonal again, nchunk blocks -- and the factors (Lk, ZR, ZH, Ww with the eliminated rows W~_k)."""
    nb, q, _ = Dw.shape
    nch = (nb + Lc - 1) // Lc
    Lk = np.zeros_like(Dw); ZR = np.zeros_like(Dw); ZH = np.zeros_like(Dw)
    Wn = Ww.copy()
    Dt = np.zeros((nch, q, q)); Et = np.zeros((nch, q, q)); Wt = np.zeros((nch,) + Ww.shape[1:])
    DtR = np.zeros_like(Dt); G = np.zeros_like(Wt)          # contributions of the chunk on the LEFT
    for c in range(nch):
        j0, j1 = c * Lc, min(nb, (c + 1) * Lc)
        Dh = Dw[j0].copy(); Wh = Ww[j0].copy()
        if j1 - j0 == 1:                                     # head only: original coupling to the next head
            Dt[c], Wt[c] = Dh, Wh
            if j1 < nb:
                Et[c] = Ew[j0]
            Wn[j0] = 0.0
            continue
        F = Ew[j0].T.copy()                                  # rows j0+1, cols head
        Dk = Dw[j0 + 1].copy(); wk = Ww[j0 + 1].copy()
        for k in range(j0 + 1, j1):
            Lk[k] = np.linalg.cholesky(Dk)
            Er = Ew[k] if k + 1 < nb else np.zeros((q, q))   # rows k, cols k+1
            ZR[k] = np.linalg.solve(Lk[k], Er)
            ZH[k] = np.linalg.solve(Lk[k], F)
            Wn[k] = np.linalg.solve(Lk[k], wk)
            Dh -= ZH[k].T @ ZH[k]
            Wh -= ZH[k].T @ Wn[k]
            if k + 1 < j1:
                Dk = Dw[k + 1] - ZR[k].T @ ZR[k]
                wk = Ww[k + 1] - ZR[k].T @ Wn[k]
                F = -ZR[k].T @ ZH[k]
            elif k + 1 < nb:                                 # next head
                DtR[c + 1] = ZR[k].T @ ZR[k]
                G[c + 1] = ZR[k].T @ Wn[k]
                Et[c] = -ZH[k].T @ ZR[k]                     # rows head c, cols head c+1
        Dt[c], Wt[c] = Dh, Wh
        Wn[j0] = 0.0                                         # head rows leave the eliminated set
    Dt -= DtR
    Wt -= G
    return Dt, Et, Wt, Lk, ZR, ZH, Wn


def solve_chunked(A, bc, D, E, Wt_in, lam, Lc, dfloor=(1e-6, 1e32)):
    """Same system as solve_bcr, with the chunk pre-reduction in front of the cyclic reduction."""
    nb, q, _ = D.shape
    ncp = A.shape[0]
    Dw = D.copy()
    for k in range(nb):
        Dw[k] += lam * np.diag(np.clip(np.diag(D[k]), *dfloor))
    Dt, Et, Wt, Lk, ZR, ZH, Wn = prereduce(Dw, E, Wt_in, Lc)
    # head system: cyclic reduction with lam = 0 on the already damped blocks.  The camera block must see the
    # eliminated rows too: fold them into A, bc first (what the SYRK over all rows does on the device)
    Sx = np.zeros((ncp + 1, ncp + 1))
    for k in range(nb):
        Sx += Wn[k].T @ Wn[k]
    dA = np.clip(np.diag(A), *dfloor)
    A2 = A + lam * np.diag(dA) - Sx[:ncp, :ncp]
    bc2 = bc - Sx[:ncp, ncp]
    dc, dst = solve_bcr(A2, bc2, Dt, Et, Wt, 0.0, dfloor=(0.0, 0.0))
    ds = np.zeros((nb, q))
    nch = Dt.shape[0]
    for c in range(nch):
        j0, j1 = c * Lc, min(nb, (c + 1) * Lc)
        ds[j0] = dst[c]
        for k in range(j1 - 1, j0, -1):
            v = Wn[k][:, ncp] - Wn[k][:, :ncp] @ dc - ZH[k] @ ds[j0]
            if k + 1 < nb:
                v -= ZR[k] @ (ds[k + 1] if k + 1 < j1 else dst[c + 1])
            ds[k] = np.linalg.solve(Lk[k].T, v)
    return dc, ds


if __name__ == '__main__':
    import sys
    sys.path.insert(0, '/root/repo')
    from mvus_b200 import synth
    from oracle import ba_oracle
    for bw, kw, bakw in [(3, dict(nc=4, det_per_cam=300, gaps=[(0.4, 0.5)]), dict()),
                         (4, dict(nc=3, det_per_cam=300, rolling_shutter=True, motion_type='F'), dict(rs=True, motion_reg=True, motion_weights=1e2)),
                         (5, dict(nc=3, det_per_cam=200, rolling_shutter=True, motion_type='KE', frames_per_knot=3.0), dict(rs=True, motion_reg=True, motion_weights=1e2))]:
        fl, _ = synth.make_flight(**kw)
        prob = ba_oracle.Problem(fl, fl.numCam, **bakw)
        x = prob.x0
        J = prob.jacobian(x).toarray()
        free = prob.free_mask()
        J[:, ~free] = 0
        r = prob.residual(x)
        n_ctrl = sum(prob.ncoef)
        cols = np.zeros((n_ctrl, 3), int)
        j = 0
        for s in range(prob.S):
            for l in range(prob.ncoef[s]):
                for ax in range(3):
                    cols[j, ax] = prob.coef_off[s] + ax * prob.ncoef[s] + l
                j += 1
        A, bc, D, E, Wt, perm = assemble(J, r, prob.n_other, n_ctrl, cols, bw)
        # check band assumption
        H = J.T @ J
        lam = 1e-4
        dc, ds = solve_bcr(A, bc, D, E, Wt, lam)
        dd = np.clip(np.diag(H), 1e-6, 1e32)
        ref = np.linalg.solve(H + lam * np.diag(dd), -J.T @ r)
        mine = np.zeros(prob.n)
        mine[:prob.n_other] = dc
        flat = ds.reshape(-1)
        ok = perm >= 0
        mine[perm[ok]] = flat[ok]
        print('bw', bw, 'n', prob.n, 'nb', D.shape[0], 'rel err', np.abs(mine - ref).max() / np.abs(ref).max())
