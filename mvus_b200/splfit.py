"""Scene.traj_to_spline (reconstruction/common.py:224-270) with the spline fits on the GPU.

``splprep(x, u, s, k)`` below has the call signature and return value of
``scipy.interpolate.splprep(x, u=u, s=s, k=k)`` as common.py:247/267 use it (unit weights, returns
``([t, [cx, cy, cz], k], u)``).  FITPACK's algorithm (parcur / fppara) is split in two:

* everything that touches the m data points -- B-spline basis, banded normal equations, their Cholesky
  solve, residuals and per-knot-interval residual sums -- runs in CUDA (mvus_b200/csrc/spl_fit.cuh through
  ``_cabi.SplHandle``); the data are uploaded once per interval and stay in HBM for all fit iterations;
* the scalar strategy of fppara -- how many knots to add (nplus), where (fpknot: middle data point of the
  interval with the largest residual sum), when to stop, and the rational-interpolation search for the
  smoothing parameter p (fprati) -- is host control flow here, a few hundred decisions per fit.

There is no CPU fallback: without the CUDA library ``_cabi`` raises.  The oracle
(oracle/fitpack_oracle.py, test infrastructure) is a separate restatement that the tests pin against the
installed SciPy and then compare with this path.
"""
import numpy as np

from . import _cabi

TOL, MAXIT = 0.001, 20          # fppara: relative accuracy of fp = s, iterations of the p search
CON1, CON9, CON4 = 0.1, 0.9, 0.04
DEVICE = 0


def find_intervals(x, gap=5, idx=False):
    """util.find_intervals (tools/util.py:58-87): start / end of every continuous part of ascending
    time stamps (interruptions of `gap` or more split; parts shorter than `gap` are dropped)."""
    x = np.asarray(x)
    assert len(x.shape) == 1 and (x[1:] > x[:-1]).all(), 'Input must be an ascending 1D-array'
    x_s, x_e = np.append(-np.inf, x), np.append(x, np.inf)
    start = x_s[1:] - x_s[:-1] >= gap
    end = x_e[:-1] - x_e[1:] <= -gap
    interval = np.array([x[start], x[end]])
    int_idx = np.array([np.where(start)[0], np.where(end)[0]])
    mask = interval[1] - interval[0] >= gap
    interval, int_idx = interval[:, mask], int_idx[:, mask]
    assert (interval[0, 1:] > interval[1, :-1]).all()
    return (interval, int_idx) if idx else interval


def _interpolation_knots(u, k):
    """fppara label 10: knots of the interpolating spline (n = m + k + 1)."""
    m, k1 = len(u), k + 1
    k3 = k // 2
    inner = u[k3 + 1:k3 + 1 + m - k1] if 2 * k3 != k else 0.5 * (u[k3 + 1:k3 + 1 + m - k1] + u[k3:k3 + m - k1])
    return np.concatenate((np.full(k1, u[0]), inner, np.full(k1, u[-1])))


def _count_inside(t, k, u):
    """nrdata: data points strictly inside each knot interval (u ascending)."""
    kn = t[k:len(t) - k]
    lo = np.searchsorted(u, kn[:-1], side='right')
    hi = np.searchsorted(u, kn[1:], side='left')
    return np.maximum(hi - lo, 0)


def _add_knot(u, t, k, fpint, nrdata):
    """fpknot: one new knot at the middle data point of the interval with the largest residual sum."""
    cand = np.where(nrdata != 0, fpint, -1.0)
    number = int(np.argmax(cand))                   # first maximum, as the strict `fpmax < fpint(j)` scan finds it
    fpmax = cand[number]
    if not fpmax > 0.0:
        return None
    maxpt = int(nrdata[number])
    maxbeg = 1 + int(np.sum(nrdata[:number])) + number
    ihalf = maxpt // 2 + 1
    nrx = maxbeg + ihalf                            # 1-based index of the data point that becomes a knot
    lo, hi = ihalf - 1, maxpt - ihalf
    fpint = np.concatenate((fpint[:number], [fpmax * lo / maxpt, fpmax * hi / maxpt], fpint[number + 1:]))
    nrdata = np.concatenate((nrdata[:number], [lo, hi], nrdata[number + 1:]))
    return np.insert(t, number + k + 1, u[nrx - 1]), fpint, nrdata


def _jump_penalty(t, k):
    """Upper band of B^T B, B = fpdisc's matrix of the jumps of the k-th derivative at the interior knots
    (row r touches coefficients r .. r+k+1).  pen[d][j] = (B^T B)[j-d][j], d = 0 .. k+1."""
    n = len(t)
    k1, k2 = k + 1, k + 2
    nk1 = n - k1
    nrint = nk1 - k
    fac = nrint / (t[nk1] - t[k])
    rows = nk1 - k1
    ls = np.arange(k2, nk1 + 1)                     # FITPACK's 1-based l
    hv = np.empty((rows, 2 * k1))
    for j in range(1, k1 + 1):
        hv[:, j - 1] = t[ls - 1] - t[ls + j - k2 - 1]
        hv[:, j + k1 - 1] = t[ls - 1] - t[ls + j - 1]
    b = np.empty((rows, k2))
    lp = ls - k1
    for j in range(1, k2 + 1):
        prod = hv[:, j - 1].copy()
        for i in range(1, k + 1):
            prod = prod * hv[:, j + i - 1] * fac
        b[:, j - 1] = (t[lp + j - 1 + k1 - 1] - t[lp + j - 1 - 1]) / prod
    pen = np.zeros((k2, nk1))
    r = np.arange(rows)
    for a in range(k2):
        for bb in range(a, k2):
            np.add.at(pen[bb - a], r + bb, b[:, a] * b[:, bb])
    return pen


def fit(u, x, s, k=3, device=None):
    """FITPACK fppara on the device data: returns (t, c [idim x (n-k-1)], fp, ier)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
    m, k1 = len(u), k + 1
    if m <= k:
        raise TypeError('m > k must hold')          # what splprep raises (and traj_to_spline catches)
    if not (u[1:] > u[:-1]).all():
        raise ValueError('Invalid inputs.')
    hd = _cabi.SplHandle(u, x, k=k, device=DEVICE if device is None else device)
    try:
        nmin, nmax, nest = 2 * k1, m + k1, m + 2 * k
        acc = TOL * s
        if s == 0.0:
            t = _interpolation_knots(u, k)
            c, fp, _, _ = hd.solve(t)
            return t, c, 0.0, -1
        t = np.concatenate((np.full(k1, u[0]), np.full(k1, u[-1])))
        fpold, nplus, ier, fp0 = 0.0, 0, -2, None
        for _ in range(m):
            n = len(t)
            if n == nmin:
                ier = -2
            c, fp, fpint, dsum = hd.solve(t)
            if ier == -2:
                fp0 = fp
            fpms = fp - s
            if abs(fpms) < acc:
                return t, c, fp, (-2 if ier == -2 else 0)
            if fpms < 0.0:
                break
            if n == nmax:
                return t, c, fp, -1
            if n == nest:
                return t, c, fp, 1
            if ier == 0:
                npl1 = nplus * 2
                if fpold - fp > acc:
                    npl1 = int(nplus * fpms / (fpold - fp))
                nplus = min(nplus * 2, max(npl1, nplus // 2, 1))
            else:
                nplus, ier = 1, 0
            fpold = fp
            nrdata = _count_inside(t, k, u)
            for _l in range(nplus):
                out = _add_knot(u, t, k, fpint, nrdata)
                if out is None:
                    break
                t, fpint, nrdata = out
                if len(t) == nmax:
                    t = _interpolation_knots(u, k)
                    break
                if len(t) == nest:
                    break
        # ---- smoothing spline on the final knots: F(p) = s
        n = len(t)
        nk1 = n - k1
        if n == nmin:
            return t, c, fp, -2
        pen = _jump_penalty(t, k)
        p1, f1, p3, f3 = 0.0, fp0 - s, -1.0, fpms
        p = nk1 / dsum
        ich1 = ich3 = 0
        for it in range(1, MAXIT + 1):
            c, fp, _, _ = hd.solve(t, pen=pen, pscale=1.0 / (p * p))      # (FITPACK appends the rows B / p)
            fpms = fp - s
            if abs(fpms) < acc:
                return t, c, fp, 0
            if it == MAXIT:
                return t, c, fp, 3
            p2, f2 = p, fpms
            moved = False
            if ich3 == 0:
                if f2 - f3 > acc:
                    if f2 < 0.0:
                        ich3 = 1
                else:                                # the initial choice of p was too large
                    p3, f3 = p2, f2
                    p = p * CON4
                    if p <= p1:
                        p = p1 * CON9 + p2 * CON1
                    moved = True
            if not moved and ich1 == 0:
                if f1 - f2 > acc:
                    if f2 > 0.0:
                        ich1 = 1
                else:                                # the initial choice of p was too small
                    p1, f1 = p2, f2
                    p = p / CON4
                    if p3 >= 0.0 and p >= p3:
                        p = p2 * CON1 + p3 * CON9
                    moved = True
            if moved:
                continue
            if f2 >= f1 or f2 <= f3:
                return t, c, fp, 2
            # fprati: rational interpolation through (p1, f1), (p2, f2), (p3, f3)
            if p3 > 0.0:
                h1, h2, h3 = f1 * (f2 - f3), f2 * (f3 - f1), f3 * (f1 - f2)
                p = -(p1 * p2 * h3 + p2 * p3 * h1 + p3 * p1 * h2) / (p1 * h1 + p2 * h2 + p3 * h3)
            else:
                p = (p1 * (f1 - f3) * f2 - p2 * (f2 - f3) * f1) / ((f1 - f2) * f3)
            if f2 < 0.0:
                p3, f3 = p2, f2
            else:
                p1, f1 = p2, f2
        return t, c, fp, 0
    finally:
        hd.close()


def splprep(x, u, s, k=3):
    """scipy.interpolate.splprep(x, u=u, s=s, k=k) -> ([t, [c_0, ..], k], u)."""
    t, c, fp, ier = fit(u, x, s, k)
    nk1 = len(t) - k - 1
    # scipy returns coefficient arrays of length n - k - 1 per dimension
    return [t, [np.ascontiguousarray(c[d][:nk1]) for d in range(c.shape[0])], k], np.asarray(u, dtype=np.float64)


def traj_to_spline(scene, smooth_factor):
    """Scene.traj_to_spline (common.py:224-270): one spline per continuous interval of ``scene.traj``; the
    smoothing factor is adapted until the ratio interval length / number of knots lies inside
    [min(smooth_factor), max(smooth_factor)]; degree 1 when the cubic fit raises."""
    assert len(smooth_factor) == 2, 'Smoothness should be defined by two parameters (min, max)'
    timestamp = scene.traj[0]
    interval, idx = find_intervals(timestamp, idx=True)
    tck = [None] * interval.shape[1]
    for i in range(interval.shape[1]):
        part = scene.traj[:, idx[0, i]:idx[1, i] + 1]
        measure = part[0, -1] - part[0, 0]
        s = (1e-3) ** 2 * measure
        thres_min, thres_max = min(smooth_factor), max(smooth_factor)
        prev, t = 0, 0
        try:
            while True:
                tck[i], u = splprep(part[1:], part[0], s, k=3)
                numKnot = len(tck[i][0]) - 4
                if numKnot == prev and numKnot == 4 and t == 2:
                    break
                else:
                    prev = numKnot
                if measure / numKnot > thres_max:
                    s /= 1.5
                    t = 1
                elif measure / numKnot < thres_min:
                    s *= 2
                    t = 2
                else:
                    break
        except _cabi.MvusError:
            raise                                    # a missing GPU / library is never hidden behind the k = 1 fallback
        except Exception:
            tck[i], u = splprep(part[1:], part[0], s, k=1)
    scene.spline['tck'], scene.spline['int'] = tck, interval
    return scene.spline
