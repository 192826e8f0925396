"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the smoothing-spline
fit behind ``Scene.traj_to_spline`` (reconstruction/common.py:224-270), i.e. of
``scipy.interpolate.splprep(x, u=u, s=s, k=k)`` with unit weights.

The algorithm is NOT in /root/reference: it lives in SciPy's bundled FITPACK (Dierckx), routine
``parcur`` -> ``fppara`` with ``fpbspl``, ``fpgivs/fprota`` (Givens QR of the banded observation matrix),
``fpback``, ``fpknot`` (knot placement), ``fpdisc`` (derivative jumps) and ``fprati`` (rational
interpolation for the smoothing parameter); installed version: SciPy 1.18.1.  Restated here from the
published algorithm (P. Dierckx, "Curve and Surface Fitting with Splines", 1993, ch. 5 and 9):

  1. start with no interior knots; repeat: weighted least-squares spline on the current knots; fp = residual
     sum of squares; if |fp - s| < 0.001 s stop; if fp < s go to 3; otherwise
  2. add `nplus` knots, one at a time, each in the knot interval with the largest residual sum, at the data
     point in the middle of that interval (fpknot); nplus follows the decrease of fp
     (nplus <- min(2 nplus, max(int(nplus (fp - s) / (fpold - fp)), nplus / 2, 1)));
  3. smoothing: with the knots fixed, find p with F(p) = s where F(p) is the residual sum of squares of
     min sum (x - s(u))^2 + (1/p^2) sum (jumps of the k-th derivative)^2 (FITPACK appends the jump rows
     divided by p to the observation matrix), by rational interpolation (<= 20 iterations).

This restatement forms the banded normal equations (what the CUDA path does) instead of FITPACK's row-by-row
Givens rotations: identical in exact arithmetic; fp is summed from the actual residuals.  It is pinned
against the installed ``splprep`` in tests/test_splfit.py (identical knots, coefficients <= 1e-9 relative).
"""
import numpy as np
from scipy.linalg import cholesky_banded, cho_solve_banded

TOL, MAXIT = 0.001, 20
CON1, CON9, CON4 = 0.1, 0.9, 0.04


def bspl_basis(t, k, u, l):
    """Non-zero B-splines of degree k at u with t[l] <= u < t[l+1] (FITPACK fpbspl), l 0-based: h[0..k]
    belong to coefficients l-k .. l."""
    h = np.zeros((k + 1,) + np.shape(u))
    h[0] = 1.0
    for j in range(1, k + 1):
        hh = h[:j].copy()
        h[0] = 0.0
        for i in range(j):
            li = l + i + 1
            lj = li - j
            f = hh[i] / (t[li] - t[lj])
            h[i] = h[i] + f * (t[li] - u)
            h[i + 1] = f * (u - t[lj])
    return h


def knot_intervals(t, k, u):
    """Index l (0-based) with t[l] <= u < t[l+1], k <= l <= n-k-2; the last data point falls in the last
    interval (fppara's search loop)."""
    n = len(t)
    return np.clip(np.searchsorted(t, u, side='right') - 1, k, n - k - 2)


def normal_equations(t, k, u, x):
    """Banded normal equations of the least-squares spline: G (upper band storage, k+1 rows as
    scipy's cholesky_banded wants with lower=False), rhs (idim x nk1)."""
    n = len(t)
    nk1 = n - k - 1
    l = knot_intervals(t, k, u)
    h = bspl_basis(t, k, u, l)                         # (k+1, m)
    G = np.zeros((k + 1, nk1))
    rhs = np.zeros((x.shape[0], nk1))
    for a in range(k + 1):
        ia = l - k + a
        for b in range(a, k + 1):
            ib = l - k + b
            np.add.at(G[k - (b - a)], ib, h[a] * h[b])           # G[k + i - j, j] = A[i, j], i <= j
        for d in range(x.shape[0]):
            np.add.at(rhs[d], ia, h[a] * x[d])
    return G, rhs, l, h


def disc_jumps(t, k):
    """FITPACK fpdisc: rows = jumps of the k-th derivative of the B-splines at the interior knots
    (n - 2k - 2 rows, k+2 non-zeros each, row r touches coefficients r .. r+k+1)."""
    n = len(t)
    k1, k2 = k + 1, k + 2
    nk1 = n - k1
    nrint = nk1 - k
    fac = nrint / (t[nk1] - t[k])
    b = np.zeros((nk1 - k1, k2))
    for l in range(k2, nk1 + 1):                         # 1-based l as in FITPACK
        lmk = l - k1
        hv = np.zeros(2 * k1)
        for j in range(1, k1 + 1):
            ik = j + k1
            lj = l + j
            lk = lj - k2
            hv[j - 1] = t[l - 1] - t[lk - 1]
            hv[ik - 1] = t[l - 1] - t[lj - 1]
        lp = lmk
        for j in range(1, k2 + 1):
            jk = j
            prod = hv[j - 1]
            for i in range(1, k + 1):
                jk += 1
                prod = prod * hv[jk - 1] * fac
            lk = lp + k1
            b[lmk - 1, j - 1] = (t[lk - 1] - t[lp - 1]) / prod
            lp += 1
    return b


def solve_banded(G, rhs):
    cb = cholesky_banded(G, lower=False)
    return np.array([cho_solve_banded((cb, False), r) for r in rhs]), cb


def residual_sums(t, k, u, x, c, l, h):
    """fp and FITPACK's per-knot-interval bookkeeping (fpint, nrdata): the residual of a data point that
    sits exactly on an interior knot is split half/half between the two intervals (fppara, label 140)."""
    n = len(t)
    nk1 = n - k - 1
    m = len(u)
    s = np.zeros_like(x)
    for a in range(k + 1):
        s += h[a] * c[:, l - k + a]
    term = np.sum((x - s) ** 2, axis=0)
    fp = float(term.sum())
    nrint = nk1 - k
    fpint = np.zeros(nrint)
    nrdata = np.zeros(nrint, dtype=int)
    fpart = 0.0
    i = 0                                     # 0-based interval index
    ll = k + 1                                # 0-based index of the next interior knot t[ll]
    new = 0
    for it in range(m):
        if u[it] >= t[ll] and ll <= nk1 - 1:
            new = 1
            ll += 1
        fpart += term[it]
        if new:
            store = term[it] * 0.5
            fpint[i] = fpart - store
            i += 1
            fpart = store
            new = 0
    fpint[nrint - 1] = fpart
    return fp, fpint, term


def count_data(t, k, u):
    """nrdata of FITPACK: number of data points strictly inside each knot interval."""
    n = len(t)
    nk1 = n - k - 1
    out = []
    for j in range(k, nk1):
        out.append(int(np.sum((u > t[j]) & (u < t[j + 1]))))
    return np.array(out, dtype=int)


def add_knot(u, t, k, fpint, nrdata):
    """FITPACK fpknot: new knot at the middle data point of the interval with the largest fpint."""
    nrint = len(fpint)
    fpmax, number, maxpt, maxbeg = 0.0, -1, 0, 0
    jbegin = 1                                           # 1-based index of the first data point inside (istart = 1)
    for j in range(nrint):
        jpoint = nrdata[j]
        if fpmax < fpint[j] and jpoint != 0:
            fpmax, number, maxpt, maxbeg = fpint[j], j, jpoint, jbegin
        jbegin += jpoint + 1
    if number < 0:
        return None
    ihalf = maxpt // 2 + 1
    nrx = maxbeg + ihalf                                 # 1-based data index
    fpint = list(fpint)
    nrdata = list(nrdata)
    an_lo, an_hi = ihalf - 1, maxpt - ihalf
    fpint[number:number + 1] = [fpmax * an_lo / maxpt, fpmax * an_hi / maxpt]
    nrdata[number:number + 1] = [an_lo, an_hi]
    pos = number + k + 1                                 # new knot goes after t[number + k]
    t = np.insert(t, pos, u[nrx - 1])
    return t, np.array(fpint), np.array(nrdata, dtype=int)


def fprati(p1, f1, p2, f2, p3, f3):
    if p3 > 0.0:
        h1, h2, h3 = f1 * (f2 - f3), f2 * (f3 - f1), f3 * (f1 - f2)
        p = -(p1 * p2 * h3 + p2 * p3 * h1 + p3 * p1 * h2) / (p1 * h1 + p2 * h2 + p3 * h3)
    else:
        p = (p1 * (f1 - f3) * f2 - p2 * (f2 - f3) * f1) / ((f1 - f2) * f3)
    if f2 < 0.0:
        p3, f3 = p2, f2
    else:
        p1, f1 = p2, f2
    return p, p1, f1, p3, f3


def interpolation_knots(u, k):
    m = len(u)
    k1 = k + 1
    t = np.concatenate((np.full(k1, u[0]), np.zeros(m - k1), np.full(k1, u[-1])))
    k3 = k // 2
    j = k3 + 2                                           # 1-based
    for i in range(k1, m):                               # 0-based target positions k1 .. m-1
        t[i] = u[j - 1] if 2 * k3 != k else 0.5 * (u[j - 1] + u[j - 2])
        j += 1
    return t


def parcur_fit(u, x, s, k=3, nest=None, solver=None):
    """splprep(x, u=u, s=s, k=k) with unit weights -> (t, c (idim x (n-k-1)), fp, ier).
    `solver`: optional object providing normal_equations / solve / residual_sums with the signatures used
    below (the CUDA path plugs in here in the product's own host loop; the oracle uses NumPy)."""
    u = np.asarray(u, dtype=np.float64)
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    m = len(u)
    k1 = k + 1
    nmin = 2 * k1
    nmax = m + k1
    nest = m + 2 * k if nest is None else nest
    acc = TOL * s
    if s == 0.0:
        t = interpolation_knots(u, k)
        G, rhs, l, h = normal_equations(t, k, u, x)
        c, _ = solve_banded(G, rhs)
        fp, _, _ = residual_sums(t, k, u, x, c, l, h)
        return t, c, 0.0, -1
    t = np.concatenate((np.full(k1, u[0]), np.full(k1, u[-1])))
    fpold, nplus = 0.0, 0
    nrdata = np.array([m - 2], dtype=int)
    fpint = np.zeros(1)
    fp0 = None
    ier = -2
    for _ in range(m):
        n = len(t)
        if n == nmin:
            ier = -2
        G, rhs, l, h = normal_equations(t, k, u, x)
        c, cb = solve_banded(G, rhs)
        fp, fpint_new, term = residual_sums(t, k, u, x, c, l, h)
        if ier == -2:
            fp0 = fp
        fpms = fp - s
        if abs(fpms) < acc:
            return t, c, fp, (ier if ier == -2 else 0)
        if fpms < 0.0:
            break                                        # -> smoothing with these knots
        if n == nmax:
            return t, c, fp, -1
        if n == nest:
            return t, c, fp, 1
        if ier == 0:
            npl1 = nplus * 2
            rn = nplus
            if fpold - fp > acc:
                npl1 = int(rn * fpms / (fpold - fp))
            nplus = min(nplus * 2, max(npl1, nplus // 2, 1))
        else:
            nplus = 1
            ier = 0
        fpold = fp
        fpint = fpint_new
        nrdata = count_data(t, k, u)
        for _l in range(nplus):
            out = add_knot(u, t, k, fpint, nrdata)
            if out is None:
                break
            t, fpint, nrdata = out
            if len(t) == nmax:
                t = interpolation_knots(u, k)
                break
            if len(t) == nest:
                break
    # ---- smoothing: F(p) = s
    n = len(t)
    nk1 = n - k1
    if n == nmin:                                        # the polynomial already satisfies fp <= s
        return t, c, fp, -2
    b = disc_jumps(t, k)
    BtB = np.zeros((k + 2, nk1))                         # upper band of B^T B (half-bandwidth k+1)
    for r in range(b.shape[0]):
        for a in range(k + 2):
            for bb in range(a, k + 2):
                BtB[k + 1 - (bb - a), r + bb] += b[r, a] * b[r, bb]
    Gp = np.zeros((k + 2, nk1))
    Gp[1:] = G
    p1, f1, p3, f3 = 0.0, fp0 - s, -1.0, fpms
    p = nk1 / float(np.sum(np.sqrt(cb[-1] ** 2)))        # sum of the diagonal of the triangular factor
    ich1 = ich3 = 0
    ier = 0
    for it in range(1, MAXIT + 1):
        pinv = 1.0 / p
        cbp = cholesky_banded(Gp + pinv * pinv * BtB, lower=False)     # (FITPACK appends the rows B / p)
        c = np.array([cho_solve_banded((cbp, False), r) for r in rhs])
        fp, _, _ = residual_sums(t, k, u, x, c, l, h)
        fpms = fp - s
        if abs(fpms) < acc:
            return t, c, fp, 0
        if it == MAXIT:
            return t, c, fp, 3
        p2, f2 = p, fpms
        done = False
        if ich3 == 0:
            if f2 - f3 > acc:
                if f2 < 0.0:
                    ich3 = 1
            else:                                        # our initial choice of p is too large
                p3, f3 = p2, f2
                p = p * CON4
                if p <= p1:
                    p = p1 * CON9 + p2 * CON1
                done = True
        if not done and ich1 == 0:
            if f1 - f2 > acc:
                if f2 > 0.0:
                    ich1 = 1
            else:                                        # our initial choice of p is too small
                p1, f1 = p2, f2
                p = p / CON4
                if p3 >= 0.0 and p >= p3:
                    p = p2 * CON1 + p3 * CON9
                done = True
        if done:
            continue
        if f2 >= f1 or f2 <= f3:
            return t, c, fp, 2
        p, p1, f1, p3, f3 = fprati(p1, f1, p2, f2, p3, f3)
    return t, c, fp, ier
