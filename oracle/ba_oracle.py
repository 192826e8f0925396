"""CPU oracle for the mvus bundle-adjustment hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference's BA error function and of the SciPy
call it is minimised with.  It is imported only by ``tests/``, by
``__graft_entry__.smoke()`` and by ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs -- never by the product package ``mvus_b200`` (which fails loudly without its CUDA
library).

Parity status: the reference has no tests or golden vectors, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in the build container by
``tests/golden/make_golden.py`` (imports /root/reference through ``oracle/ref_shim.py``)
and committed as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` re-checks the
restatement against them (<=1e-9 px) and, when /root/reference is present, against the
live reference.  Third-party arithmetic the reference calls but does not vendor:
SciPy 1.18.1 FITPACK ``splev`` (de Boor recurrence, restated in ``bspline_basis``),
OpenCV 4.13.0 ``undistortPoints`` (5 fixed-point iterations, restated in ``undistort5``)
and ``Rodrigues`` (restated in ``rodrigues``), SciPy ``least_squares`` (called, not
restated).

What follows which reference lines:
  residual            common.py:448-487 (error_BA), 304-359 (error_cam 'each'),
                      105-127 (detection_to_global), util.py:90-116 (sampling)
  motion rows         common.py:362-424 (error_motion, motion_reg branch), 273-301
                      (spline_to_traj), 959-1001 (motion_prior)
  pack / layout       common.py:612-651
  pattern_near3       common.py:490-610 (jac_BA)
  shipped_solve       common.py:655-670
The analytic Jacobian (``jacobian``) is the derivative of THAT error function (abs
included, i.e. sign-applied), checked against a complex-step derivative of the same
restatement (``jacobian_cs``); it is not the reference's finite-difference ``res.jac``
(SURVEY.md H1).
"""
import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------- geometry
def rodrigues(w):
    """Rotation vector -> matrix; complex-safe (cv2.Rodrigues, common.py:1136,1140)."""
    w = np.asarray(w)
    th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2]
    if abs(th2) < np.finfo(float).eps ** 2:
        return np.eye(3, dtype=w.dtype) + skew(w)
    th = np.sqrt(th2)
    k = w / th
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(k, k) + np.sin(th) * skew(k)


def skew(v):
    return np.array([[0 * v[0], -v[2], v[1]], [v[2], 0 * v[0], -v[0]], [-v[1], v[0], 0 * v[0]]])


def rodrigues_jac(w):
    """dR/dw_k, k=0..2 (3 matrices).  Closed form (Gallego & Yezzi 2015):
    dR/dw_k = (w_k [w]x + [w x (I - R) e_k]x) R / |w|^2 ; [e_k]x at w = 0."""
    w = np.asarray(w, dtype=float)
    th2 = w @ w
    R = rodrigues(w)
    out = []
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        if th2 < 1e-24:
            out.append(skew(e))
        else:
            out.append((w[k] * skew(w) + skew(np.cross(w, (np.eye(3) - R) @ e))) @ R / th2)
    return out


def undistort5(x, y, K4, d, iters=5):
    """cv2.undistortPoints(src, K, d) for d = (k1,k2,p1,p2,k3): 5 fixed-point iterations
    on normalised coordinates (common.py:1147-1157).  Complex-safe."""
    fx, fy, cx, cy = K4
    k1, k2, p1, p2, k3 = d
    x0 = (x - cx) / fx
    y0 = (y - cy) / fy
    xn, yn = x0, y0
    for _ in range(iters):
        r2 = xn * xn + yn * yn
        icdist = 1.0 / (1.0 + ((k3 * r2 + k2) * r2 + k1) * r2)
        dx = 2.0 * p1 * xn * yn + p2 * (r2 + 2.0 * xn * xn)
        dy = p1 * (r2 + 2.0 * yn * yn) + 2.0 * p2 * xn * yn
        xn = (x0 - dx) * icdist
        yn = (y0 - dy) * icdist
    return xn, yn


def bspline_basis(t, knots, k, x, nder=0):
    """FITPACK fpbspl/splev restated: for each x return the knot span l (t[l] <= x < t[l+1],
    clamped to [k, n-1]) and the k+1 non-zero B-spline values B_{l-k..l}(x) (and first
    derivatives if nder=1).  Complex-safe in x (span is taken from the real part)."""
    n = len(knots) - k - 1
    xr = np.real(x)
    l = np.searchsorted(knots, xr, side='right') - 1
    l = np.clip(l, k, n - 1)
    tk = np.asarray(knots)

    def basis(deg):
        h = [np.ones_like(x)]
        for j in range(1, deg + 1):
            hh = list(h)
            h = [np.zeros_like(x) for _ in range(j + 1)]
            for i in range(1, j + 1):
                li = l + i
                lj = li - j
                f = hh[i - 1] / (tk[li] - tk[lj])
                h[i - 1] = h[i - 1] + f * (tk[li] - x)
                h[i] = f * (x - tk[lj])
        return h

    B = basis(k)
    if not nder:
        return l, B
    # derivative: B'_{j,k} = k ( B_{j,k-1}/(t_{j+k}-t_j) - B_{j+1,k-1}/(t_{j+k+1}-t_{j+1}) )
    Bm = basis(k - 1)                       # values of B_{l-k+1..l, k-1}
    dB = []
    for m in range(k + 1):                  # basis index j = l-k+m
        j = l - k + m
        term = np.zeros_like(x)
        if m >= 1:
            term = term + Bm[m - 1] / (tk[j + k] - tk[j])
        if m <= k - 1:
            term = term - Bm[m] / (tk[j + k + 1] - tk[j + 1])
        dB.append(k * term)
    return l, B, dB


# ------------------------------------------------------------------------------ problem
class Problem:
    """Flat description of one ``Scene.BA(numCam, ...)`` call (what common.py:612-665 reads
    from the Scene).  Works on a reference Scene or on the mirror Scene (duck-typed)."""

    def __init__(self, scene, numCam, rs=False, motion_reg=False, motion_weights=1,
                 rs_bounds=False):
        st = scene.settings
        self.nc = int(numCam)
        self.seq = list(scene.sequence[:numCam])
        self.opt_calib = bool(st['opt_calib'])
        self.undist = bool(st['undist_points'])
        self.opt_rs = bool(rs)
        try:
            self.opt_sync = bool(st['opt_sync'])
        except (KeyError, TypeError):
            self.opt_sync = True
        self.rs_bounds = bool(rs_bounds)
        self.motion_type = st['motion_type'] if motion_reg else None
        self.motion_weight = float(motion_weights)
        self.C = 15 if self.opt_calib else 6
        self.det = [np.array(scene.detections[i], dtype=float) for i in self.seq]
        self.height = np.array([scene.cameras[i].resolution[1] for i in self.seq], dtype=float)
        self.K4 = np.array([[scene.cameras[i].K[0, 0], scene.cameras[i].K[1, 1],
                             scene.cameras[i].K[0, 2], scene.cameras[i].K[1, 2]] for i in self.seq])
        self.dist = np.array([np.asarray(scene.cameras[i].d, dtype=float).reshape(5)
                              for i in self.seq])
        self.interval = np.array(scene.spline['int'], dtype=float)
        self.knots = [np.array(t[0], dtype=float) for t in scene.spline['tck']]
        self.degree = [int(t[2]) for t in scene.spline['tck']]
        self.ncoef = [len(t[1][0]) for t in scene.spline['tck']]
        self.S = len(self.knots)
        self.N = [d.shape[1] for d in self.det]
        self.n_other = self.nc * (3 + self.C)
        self.coef_off = self.n_other + 3 * np.concatenate(([0], np.cumsum(self.ncoef)))
        self.n = int(self.coef_off[-1])
        self.row_off = np.concatenate(([0], np.cumsum([2 * n for n in self.N]))).astype(int)
        # motion samples (spline_to_traj, common.py:289-297): global unit grid, closed intervals
        self.tau = np.zeros(0)
        self.tau_spl = np.zeros(0, dtype=int)
        if self.motion_type is not None:
            grid = np.arange(self.interval[0, 0], self.interval[1, -1], 1.0)
            tau, spl = [], []
            for s in range(self.S):
                part = grid[(grid >= self.interval[0, s]) & (grid <= self.interval[1, s])]
                tau.append(part)
                spl.append(np.full(len(part), s))
            self.tau = np.concatenate(tau)
            self.tau_spl = np.concatenate(spl).astype(int)
        self.M = len(self.tau)
        self.m = int(self.row_off[-1]) + self.M
        self.x0 = self.pack(scene)

    # ---- parameter vector (common.py:616-650) ----------------------------------------
    def pack(self, scene):
        import cv2
        cams = []
        for i in self.seq:
            c = scene.cameras[i]
            r = cv2.Rodrigues(np.asarray(c.R, dtype=float))[0].reshape(-1)
            if self.opt_calib:
                cams.append(np.concatenate(([c.K[0, 0], c.K[1, 1], c.K[0, 2], c.K[1, 2]], r,
                                            np.asarray(c.t, float).reshape(3),
                                            np.asarray(c.d, float).reshape(5))))
            else:
                cams.append(np.concatenate((r, np.asarray(c.t, float).reshape(3))))
        spl = [np.ravel(np.asarray(t[1], dtype=float)) for t in scene.spline['tck']]
        return np.concatenate([np.asarray(scene.alpha, float)[self.seq],
                               np.asarray(scene.beta, float)[self.seq],
                               np.asarray(scene.rs, float)[self.seq]] + cams + spl)

    def unpack(self, x):
        nc, C = self.nc, self.C
        a, b, r = x[:nc], x[nc:2 * nc], x[2 * nc:3 * nc]
        cams = x[3 * nc:3 * nc + nc * C].reshape(nc, C)
        coefs = [x[self.coef_off[s]:self.coef_off[s + 1]].reshape(3, -1) for s in range(self.S)]
        return a, b, r, cams, coefs

    def cam_parts(self, cam, i):
        if self.opt_calib:
            return cam[:4], cam[4:7], cam[7:10], cam[10:15]
        return self.K4[i], cam[:3], cam[3:6], self.dist[i]

    def membership(self, t):
        """util.sampling(..., belong=True): 1-based interval id, 0 = none (util.py:103-106)."""
        tr = np.real(t)
        idx = np.zeros(len(tr), dtype=int)
        for s in range(self.S):
            mask = np.logical_xor(tr - self.interval[0, s] >= 0, tr - self.interval[1, s] >= 0)
            idx[mask] = s + 1
        return idx

    def bounds(self):
        lo = np.full(self.n, -np.inf)
        hi = np.full(self.n, np.inf)
        if self.rs_bounds:
            lo[2 * self.nc:3 * self.nc] = 0.0
            hi[2 * self.nc:3 * self.nc] = 1.0
        return lo, hi

    # ---- residual ------------------------------------------------------------------
    def _cam_terms(self, x, i):
        """Signed reprojection errors of camera i and everything the Jacobian needs."""
        a, b, rho, cams, coefs = self.unpack(x)
        f, xr, yr = self.det[i]
        K4, w, T, d = self.cam_parts(cams[i], i)
        tau = f + rho[i] * yr / self.height[i]
        t = a[i] * tau + b[i]
        if self.undist:
            xn, yn = undistort5(xr, yr, K4, d)
            uo, vo = K4[0] * xn + K4[2], K4[1] * yn + K4[3]
        else:
            uo, vo = xr + 0 * t, yr + 0 * t
        idx = self.membership(t)
        R = rodrigues(w)
        return dict(f=f, xr=xr, yr=yr, K4=K4, w=w, T=T, d=d, tau=tau, t=t, uo=uo, vo=vo, idx=idx,
                    R=R, a=a[i], b=b[i], rho=rho[i], coefs=coefs)

    def residual(self, x, signed=False):
        """error_BA(x) (common.py:448-487).  ``signed=True`` drops the abs() (same cost)."""
        x = np.asarray(x)
        out = np.zeros(self.m, dtype=x.dtype)
        for i in range(self.nc):
            c = self._cam_terms(x, i)
            N = self.N[i]
            eu = np.zeros(N, dtype=x.dtype)
            ev = np.zeros(N, dtype=x.dtype)
            for s in range(self.S):
                m = c['idx'] == s + 1
                if not m.any():
                    continue
                t = c['t'][m]
                l, B = bspline_basis(None, self.knots[s], self.degree[s], t)
                k = self.degree[s]
                X = [sum(B[q] * c['coefs'][s][ax][l - k + q] for q in range(k + 1)) for ax in range(3)]
                R, T, K4 = c['R'], c['T'], c['K4']
                Xc = [R[r, 0] * X[0] + R[r, 1] * X[1] + R[r, 2] * X[2] + T[r] for r in range(3)]
                eu[m] = K4[0] * Xc[0] / Xc[2] + K4[2] - c['uo'][m]
                ev[m] = K4[1] * Xc[1] / Xc[2] + K4[3] - c['vo'][m]
            r0 = self.row_off[i]
            out[r0:r0 + N] = eu
            out[r0 + N:r0 + 2 * N] = ev
        if self.M:
            out[self.row_off[-1]:] = self._motion(x)[0]
        if signed:
            return out
        if np.iscomplexobj(out):
            return out * np.where(np.real(out) < 0, -1.0, 1.0)
        return np.abs(out)

    def _motion_groups(self):
        """For every interval: indices (into the sample list) of the samples that
        util.sampling assigns to it (a <= tau < b), as error_motion does (common.py:380,411)."""
        idx = self.membership(self.tau)
        return [np.nonzero(idx == s + 1)[0] for s in range(self.S)]

    def _motion(self, x, want_jac=False):
        """Motion-prior rows (error_motion motion_reg branch + motion_prior).  Returns the
        SIGNED row values (abs applied by the caller) and, optionally, COO Jacobian pieces of
        the sign-applied rows."""
        _, _, _, _, coefs = self.unpack(x)
        w = self.motion_weight
        eps = 1e-20
        out = np.zeros(self.M, dtype=np.asarray(x).dtype)
        rows, cols, vals = [], [], []
        for s, g in enumerate(self._motion_groups()):
            if len(g) == 0:
                continue
            ts = self.tau[g]
            k = self.degree[s]
            l, B = bspline_basis(None, self.knots[s], k, ts)
            P = np.array([sum(B[q] * coefs[s][ax][l - k + q] for q in range(k + 1))
                          for ax in range(3)])                       # 3 x ng
            if self.motion_type == 'KE':
                if len(g) < 2:
                    continue
                dt = ts[1:] - ts[:-1]
                vel = (P[:, 1:] - P[:, :-1]) / (dt + eps)
                out[g[1:]] = np.sum(np.abs(w * 0.5 * (vel ** 2 * dt)) if not np.iscomplexobj(vel)
                                    else w * 0.5 * (vel ** 2 * dt) * np.sign(w), axis=0)
                if want_jac:
                    # d r_j / d C_{c,ax} = |w| v_ax dt/(dt+eps) (B_c(tau_j) - B_c(tau_{j-1}))
                    fac = abs(w) * vel * (dt / (dt + eps))               # 3 x (ng-1)
                    for ax in range(3):
                        for q in range(k + 1):
                            col = self.coef_off[s] + ax * self.ncoef[s] + (l - k + q)
                            rows += [g[1:], g[1:]]
                            cols += [col[1:], col[:-1]]
                            vals += [fac[ax] * B[q][1:], -fac[ax] * B[q][:-1]]
            else:  # 'F'
                if len(g) < 3:
                    continue
                dt1 = ts[1:-1] - ts[:-2]
                dt2 = ts[2:] - ts[1:-1]
                dt3 = dt1 + dt2
                v1 = (P[:, 1:-1] - P[:, :-2]) / (dt1 + eps)
                v2 = (P[:, 2:] - P[:, 1:-1]) / (dt2 + eps)
                acc = w * ((v2 - v1) / (dt3 + eps) * dt3)               # 3 x (ng-2), signed
                sg = np.where(np.real(acc) < 0, -1.0, 1.0)
                out[g[1:-1]] = np.sum(acc * sg, axis=0)
                if want_jac:
                    sc = w * sg * (dt3 / (dt3 + eps))                    # 3 x (ng-2)
                    for ax in range(3):
                        for q in range(k + 1):
                            col = self.coef_off[s] + ax * self.ncoef[s] + (l - k + q)
                            rows += [g[1:-1]] * 3
                            cols += [col[2:], col[1:-1], col[:-2]]
                            vals += [sc[ax] * B[q][2:] / (dt2 + eps),
                                     -sc[ax] * B[q][1:-1] * (1 / (dt2 + eps) + 1 / (dt1 + eps)),
                                     sc[ax] * B[q][:-2] / (dt1 + eps)]
        if want_jac:
            if rows:
                return out, np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)
            return out, np.zeros(0, int), np.zeros(0, int), np.zeros(0)
        return (out,)

    def cost(self, x):
        r = self.residual(x)
        return 0.5 * float(r @ r)

    # ---- Jacobians -----------------------------------------------------------------
    def jacobian_cs(self, x, cols=None, h=1e-30):
        """Complex-step derivative of the restated error function, sign-applied so that it is
        the derivative of abs(residual) away from kinks.  Dense m x len(cols)."""
        x = np.asarray(x, dtype=float)
        cols = np.arange(self.n) if cols is None else np.asarray(cols)
        sgn = np.where(self.residual(x, signed=True) < 0, -1.0, 1.0)
        sgn[self.row_off[-1]:] = 1.0          # motion rows are already sign-applied
        J = np.zeros((self.m, len(cols)))
        for q, c in enumerate(cols):
            xc = x.astype(complex)
            xc[c] += 1j * h
            J[:, q] = np.imag(self._signed_for_cs(xc)) / h * sgn
        return J

    def _signed_for_cs(self, xc):
        out = self.residual(xc, signed=True)
        return out

    def jacobian(self, x):
        """Analytic Jacobian of abs(error_BA) (CSR, m x n): 4 active control points per
        detection, all motion control points; fixed columns (opt_sync / rs off) are kept
        (they are the true derivative; masking is the solver's business)."""
        x = np.asarray(x, dtype=float)
        nc, C = self.nc, self.C
        rows, cols, vals = [], [], []

        def add(r, c, v):
            rows.append(np.asarray(r))
            cols.append(np.broadcast_to(np.asarray(c), np.shape(r)))
            vals.append(np.asarray(v, dtype=float))

        for i in range(nc):
            c = self._cam_terms(x, i)
            N = self.N[i]
            R, T, K4, d = c['R'], c['T'], c['K4'], c['d']
            dR = rodrigues_jac(c['w'])
            cam0 = 3 * nc + i * C
            if self.opt_calib and self.undist:
                dobs = _undistort5_jac(c['xr'], c['yr'], K4, d)   # (du_obs/dp, dv_obs/dp), p = 9 params
            for s in range(self.S):
                m = np.nonzero(c['idx'] == s + 1)[0]
                if len(m) == 0:
                    continue
                t = c['t'][m]
                k = self.degree[s]
                l, B, dB = bspline_basis(None, self.knots[s], k, t, nder=1)
                cf = c['coefs'][s]
                X = np.array([sum(B[q] * cf[ax][l - k + q] for q in range(k + 1)) for ax in range(3)])
                dX = np.array([sum(dB[q] * cf[ax][l - k + q] for q in range(k + 1)) for ax in range(3)])
                Xc = R @ X + T.reshape(3, 1)
                iz = 1.0 / Xc[2]
                eu = K4[0] * Xc[0] * iz + K4[2] - c['uo'][m]
                ev = K4[1] * Xc[1] * iz + K4[3] - c['vo'][m]
                su = np.where(eu < 0, -1.0, 1.0)
                sv = np.where(ev < 0, -1.0, 1.0)
                z0 = np.zeros_like(iz)
                Gu = np.array([K4[0] * iz, z0, -K4[0] * Xc[0] * iz * iz]) * su     # 3 x n
                Gv = np.array([z0, K4[1] * iz, -K4[1] * Xc[1] * iz * iz]) * sv
                GRu = R.T @ Gu                                                      # (G R)^T rows
                GRv = R.T @ Gv
                vu = np.sum(GRu * dX, axis=0)
                vv = np.sum(GRv * dX, axis=0)
                ru = self.row_off[i] + m
                rv = ru + N
                yH = c['yr'][m] / self.height[i]
                for r_, v_ in ((ru, vu), (rv, vv)):
                    add(r_, i, v_ * c['tau'][m])
                    add(r_, nc + i, v_)
                    add(r_, 2 * nc + i, v_ * c['a'] * yH)
                ro = 4 if self.opt_calib else 0
                for q in range(3):
                    dq = dR[q] @ X
                    add(ru, cam0 + ro + q, np.sum(Gu * dq, axis=0))
                    add(rv, cam0 + ro + q, np.sum(Gv * dq, axis=0))
                    add(ru, cam0 + ro + 3 + q, Gu[q])
                    add(rv, cam0 + ro + 3 + q, Gv[q])
                if self.opt_calib:
                    du = np.zeros((9, len(m)))
                    dv = np.zeros((9, len(m)))
                    du[0] = Xc[0] * iz
                    du[2] = 1.0
                    dv[1] = Xc[1] * iz
                    dv[3] = 1.0
                    if self.undist:
                        du -= dobs[0][:, m]
                        dv -= dobs[1][:, m]
                    pc = [0, 1, 2, 3, 10, 11, 12, 13, 14]
                    for q in range(9):
                        add(ru, cam0 + pc[q], du[q] * su)
                        add(rv, cam0 + pc[q], dv[q] * sv)
                for ax in range(3):
                    for q in range(k + 1):
                        col = self.coef_off[s] + ax * self.ncoef[s] + (l - k + q)
                        rows.append(ru); cols.append(col); vals.append(GRu[ax] * B[q])
                        rows.append(rv); cols.append(col); vals.append(GRv[ax] * B[q])
        if self.M:
            _, mr, mc, mv = self._motion(x, want_jac=True)
            rows.append(self.row_off[-1] + mr)
            cols.append(mc)
            vals.append(mv)
        rows = np.concatenate([np.ravel(r) for r in rows])
        cols = np.concatenate([np.ravel(c) for c in cols])
        vals = np.concatenate([np.ravel(v) for v in vals])
        J = sp.coo_matrix((vals, (rows, cols)), shape=(self.m, self.n)).tocsr()
        J.sum_duplicates()
        return J

    def free_mask(self):
        """Columns the reference actually optimises: alpha/beta only if opt_sync, rho only if
        rs (their pattern columns are zeroed otherwise, common.py:512-521)."""
        free = np.ones(self.n, dtype=bool)
        if not self.opt_sync:
            free[:2 * self.nc] = False
        if not self.opt_rs:
            free[2 * self.nc:3 * self.nc] = False
        return free

    # ---- the reference's sparsity pattern (jac_BA, common.py:490-610) ------------------
    def pattern_near3(self, x=None, near=3):
        """0/1 pattern of jac_BA: per covered detection alpha_i, beta_i (if opt_sync), rho_i (if
        rs), the camera block and the ``near`` knots closest to its time stamp x 3 axes; an
        uncovered detection has an empty row (common.py:565-566); the camera block is stacked
        twice, for the u and the v rows (568); motion rows get the 3 nearest knots (571-585)."""
        x = self.x0 if x is None else x
        nc, C = self.nc, self.C
        R_all, C_all = [], []
        for i in range(nc):
            c = self._cam_terms(x, i)
            N = self.N[i]
            r0 = self.row_off[i]
            rows, cols = [], []
            vis = np.nonzero(c['idx'] > 0)[0]
            fixed = list(range(3 * nc + i * C, 3 * nc + (i + 1) * C))
            if self.opt_sync:
                fixed += [i, nc + i]
            if self.opt_rs:
                fixed += [2 * nc + i]
            for col in fixed:
                rows.append(vis)
                cols.append(np.full(len(vis), col))
            for s in range(self.S):
                m = np.nonzero(c['idx'] == s + 1)[0]
                if len(m) == 0:
                    continue
                knot = self.knots[s][2:-2]
                # default (unstable) sort on purpose: clamped end knots tie and the reference's
                # own np.argsort call breaks the tie the same way only with the same algorithm
                kk = np.argsort(np.abs(knot[None, :] - c['t'][m][:, None]), axis=1)[:, :near]
                for ax in range(3):
                    rows.append(np.repeat(m, kk.shape[1]))
                    cols.append((self.coef_off[s] + ax * len(knot) + kk).ravel())
            blk_r = np.concatenate(rows)
            blk_c = np.concatenate(cols)
            R_all += [r0 + blk_r, r0 + N + blk_r]
            C_all += [blk_c, blk_c]
        for j in range(self.M):
            s = self.membership(self.tau[j:j + 1])[0] - 1    # 0 -> -1 = last spline, as the reference
            knot = self.knots[s][2:-2]
            kk = np.argsort(np.abs(knot - self.tau[j]))[:near]
            for ax in range(3):
                R_all.append(np.full(len(kk), self.row_off[-1] + j))
                C_all.append(self.coef_off[s] + ax * len(knot) + kk)
        R_all = np.concatenate(R_all)
        C_all = np.concatenate(C_all)
        A = sp.coo_matrix((np.ones(len(R_all), dtype=np.int8), (R_all, C_all)),
                          shape=(self.m, self.n)).tocsr()
        A.data[:] = 1
        return A

    # ---- solver oracles --------------------------------------------------------------
    def shipped_solve(self, x0=None, max_nfev=10, pattern=None):
        """Oracle A: the reference's own call (common.py:670) on the restated error function:
        2-point finite differences on the near=3 pattern, TRF + LSMR, xtol=1e-12."""
        from scipy.optimize import least_squares
        x0 = self.x0 if x0 is None else x0
        A = self.pattern_near3(x0) if pattern is None else pattern
        lo, hi = self.bounds()
        bounds = (lo, hi) if self.rs_bounds else (-np.inf, np.inf)
        return least_squares(self.residual, x0, jac_sparsity=A, tr_solver='lsmr', xtol=1e-12,
                             max_nfev=max_nfev, verbose=0, bounds=bounds)

    def exact_solve(self, x0=None, max_nfev=200, ftol=1e-12, xtol=1e-12, gtol=1e-12, dense=None):
        """Oracle B: the same error function minimised by SciPy TRF with the exact (analytic)
        Jacobian until a tolerance stops it."""
        from scipy.optimize import least_squares
        x0 = self.x0 if x0 is None else x0
        free = self.free_mask()
        lo, hi = self.bounds()
        bounds = (lo[free], hi[free]) if self.rs_bounds else (-np.inf, np.inf)
        dense = (self.n <= 1500) if dense is None else dense

        def expand(z):
            xx = np.array(x0, dtype=float)
            xx[free] = z
            return xx

        def fun(z):
            return self.residual(expand(z))

        def jac(z):
            J = self.jacobian(expand(z))[:, np.nonzero(free)[0]]
            return J.toarray() if dense else J

        res = least_squares(fun, np.asarray(x0, float)[free], jac=jac,
                            tr_solver='exact' if dense else 'lsmr', ftol=ftol, xtol=xtol, gtol=gtol,
                            max_nfev=max_nfev, bounds=bounds,
                            tr_options={} if dense else {'atol': 1e-12, 'btol': 1e-12})
        res.x = expand(res.x)
        return res


def _undistort5_jac(x, y, K4, d, iters=5):
    """Forward-mode derivative of (u_obs, v_obs) = K * undistort5(x, y) w.r.t.
    p = (fx, fy, cx, cy, k1, k2, p1, p2, k3).  Returns (du 9xN, dv 9xN)."""
    fx, fy, cx, cy = K4
    k1, k2, p1, p2, k3 = d
    N = len(x)
    x0 = (x - cx) / fx
    y0 = (y - cy) / fy
    dx0 = np.zeros((9, N)); dy0 = np.zeros((9, N))
    dx0[0] = -x0 / fx
    dx0[2] = -1.0 / fx
    dy0[1] = -y0 / fy
    dy0[3] = -1.0 / fy
    xn, yn = x0, y0
    dxn, dyn = dx0.copy(), dy0.copy()
    for _ in range(iters):
        r2 = xn * xn + yn * yn
        dr2 = 2 * (xn * dxn + yn * dyn)
        poly = ((k3 * r2 + k2) * r2 + k1) * r2
        dpoly_dr2 = (3 * k3 * r2 + 2 * k2) * r2 + k1
        den = 1.0 + poly
        dden = dpoly_dr2 * dr2
        dden[4] += r2
        dden[5] += r2 * r2
        dden[8] += r2 * r2 * r2
        ic = 1.0 / den
        dic = -dden * ic * ic
        dX = 2 * p1 * xn * yn + p2 * (r2 + 2 * xn * xn)
        dY = p1 * (r2 + 2 * yn * yn) + 2 * p2 * xn * yn
        ddX = 2 * p1 * (dxn * yn + xn * dyn) + p2 * (dr2 + 4 * xn * dxn)
        ddX[6] += 2 * xn * yn
        ddX[7] += r2 + 2 * xn * xn
        ddY = p1 * (dr2 + 4 * yn * dyn) + 2 * p2 * (dxn * yn + xn * dyn)
        ddY[6] += r2 + 2 * yn * yn
        ddY[7] += 2 * xn * yn
        nx = (x0 - dX) * ic
        ny = (y0 - dY) * ic
        dnx = (dx0 - ddX) * ic + (x0 - dX) * dic
        dny = (dy0 - ddY) * ic + (y0 - dY) * dic
        xn, yn, dxn, dyn = nx, ny, dnx, dny
    du = fx * dxn
    dv = fy * dyn
    du[0] += xn
    du[2] += 1.0
    dv[1] += yn
    dv[3] += 1.0
    return du, dv
