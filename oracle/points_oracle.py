"""CPU restatement of ``Scene.BA(motion_prior=True)`` -- the discrete-trajectory mode of the reference
(reconstruction/common.py:466-467, 527-550, 587-605, 631-634, 681-687).  TEST INFRASTRUCTURE: only tests/ import
this.  Pinned against the reference's own ``error_BA`` closure captured from inside ``Scene.BA`` (golden
fixtures tests/golden/points_*.npz made by tests/golden/make_golden_points.py, and live when the reference tree
is present).

What the mode is, read off the reference:
  * unknowns  x = [alpha, beta, rho, camera vectors, P_0x P_0y P_0z P_1x ...]: the G points of
    ``global_traj`` (one per in-interval detection, sorted by the time stamps BEFORE the BA,
    common.py:887-944) replace the spline coefficients (common.py:631-634);
  * reprojection rows: ``error_cam(cam, 'each')`` WITHOUT motion_prior (common.py:478) -- still measured
    against the splines, which are constants in this mode, so these rows depend on the camera side only;
  * motion rows (common.py:480-482 -> error_motion(motion_prior=True), 386-403): the time stamps of the
    points follow alpha/beta/rho of the camera that saw them (detection_to_global(motion_prior=True),
    common.py:128-148); points are grouped by the spline interval their CURRENT time stamp falls in
    (util.sampling, a <= t < b), in ``global_traj`` order; Scene.motion_prior (959-1001) runs on each group;
    the result lands at the middle point (F) / the later point (KE) of each triple / pair.
    (The reference scatters by np.intersect1d on time stamps, i.e. sorted by time -- `placement` below restates
    that; it only differs from the positional order once time stamps of different cameras have crossed.)
"""
import numpy as np

from .ba_oracle import Problem, bspline_basis


class PointsProblem:
    def __init__(self, scene, numCam, rs=False, motion_weights=1, rs_bounds=False):
        self.base = Problem(scene, numCam, rs=rs, motion_reg=False, rs_bounds=rs_bounds)
        b = self.base
        self.motion_type = scene.settings['motion_type']
        self.w = float(motion_weights)
        self.nc, self.C, self.n_other = b.nc, b.C, b.n_other
        self.coefs0 = b.x0[b.n_other:].copy()
        # all_detect_to_traj (common.py:887-944) at the Scene's parameters
        a, be, rho = b.x0[:b.nc], b.x0[b.nc:2 * b.nc], b.x0[2 * b.nc:3 * b.nc]
        cam, frame, yH, ts = [], [], [], []
        for i in range(b.nc):
            f, _, yr = b.det[i]
            cam.append(np.full(len(f), i))
            frame.append(f)
            yH.append(yr / b.height[i])
            ts.append(a[i] * (f + rho[i] * yr / b.height[i]) + be[i])
        cam, frame, yH, ts = (np.concatenate(v) for v in (cam, frame, yH, ts))
        order = np.argsort(ts)
        cam, frame, yH, ts = cam[order], frame[order], yH[order], ts[order]
        keep = np.zeros(len(ts), dtype=bool)          # spline_to_traj(t=...): closed intervals (common.py:292)
        pos = np.zeros((3, len(ts)))
        for s in range(b.S):
            m = (ts >= b.interval[0, s]) & (ts <= b.interval[1, s])
            keep |= m
            if m.any():
                l, B = bspline_basis(None, b.knots[s], b.degree[s], ts[m])
                k = b.degree[s]
                co = self.coefs0[b.coef_off[s] - b.n_other:b.coef_off[s + 1] - b.n_other].reshape(3, -1)
                pos[:, m] = [sum(B[q] * co[ax][l - k + q] for q in range(k + 1)) for ax in range(3)]
        self.pt_cam, self.pt_frame, self.pt_yH = cam[keep].astype(int), frame[keep], yH[keep]
        self.ts0 = ts[keep]
        self.G = int(keep.sum())
        self.n = self.n_other + 3 * self.G
        self.m = b.m + self.G
        self.x0 = np.concatenate((b.x0[:b.n_other], np.ravel(pos[:, keep].T)))

    # ---- pieces ------------------------------------------------------------------------
    def _std(self, x):
        return np.concatenate((x[:self.n_other], self.coefs0.astype(x.dtype)))

    def timestamps(self, x):
        nc = self.nc
        a, be, rho = x[:nc], x[nc:2 * nc], x[2 * nc:3 * nc]
        c = self.pt_cam
        return a[c] * (self.pt_frame + rho[c] * self.pt_yH) + be[c]

    def neighbours(self, ts):
        """prev / next point of the same interval group in global_traj order (-1 = none), and the group id."""
        gid = self.base.membership(ts)
        prev = np.full(self.G, -1)
        nxt = np.full(self.G, -1)
        for g in range(1, self.base.S + 1):
            idx = np.nonzero(gid == g)[0]
            prev[idx[1:]] = idx[:-1]
            nxt[idx[:-1]] = idx[1:]
        return gid, prev, nxt

    def placement(self, ts, mid, gid):
        """Where the reference puts the residual of the k-th triple / pair: np.intersect1d(global_traj[3],
        global_traj_ts, assume_unique=True, return_indices=True) (common.py:401-403) returns the indices sorted
        by TIME STAMP, while the residuals are listed interval by interval in global_traj order; both orders
        agree unless time stamps of different cameras have crossed since global_traj was sorted."""
        key = np.lexsort((mid, gid[mid]))                       # interval by interval, positional inside
        listed = mid[key]
        by_time = mid[np.argsort(np.real(ts[mid]), kind='stable')]
        out = np.empty_like(mid)
        out[key] = by_time
        return out                                              # residual of middle point mid[k] lands at out[k]

    def motion(self, x, signed=False, reference_order=True):
        """error_motion(motion_prior=True) (common.py:386-403) -> length-G vector."""
        eps = 1e-20
        ts = self.timestamps(x)
        P = x[self.n_other:].reshape(-1, 3).T
        gid, prev, nxt = self.neighbours(ts)
        out = np.zeros(self.G, dtype=x.dtype)

        def absum(v):
            if signed and np.iscomplexobj(v):
                return np.sum(v * np.where(np.real(v) < 0, -1.0, 1.0), axis=0)
            return np.sum(np.abs(v), axis=0)
        if self.motion_type == 'F':
            j = np.nonzero((gid > 0) & (prev >= 0) & (nxt >= 0))[0]
            p, n = prev[j], nxt[j]
            dt1, dt2 = ts[j] - ts[p], ts[n] - ts[j]
            dt3 = dt1 + dt2
            v1 = (P[:, j] - P[:, p]) / (dt1 + eps)
            v2 = (P[:, n] - P[:, j]) / (dt2 + eps)
            out[self.placement(ts, j, gid) if reference_order else j] = absum(self.w * ((v2 - v1) / (dt3 + eps) * dt3))
        else:
            j = np.nonzero((gid > 0) & (prev >= 0))[0]
            p = prev[j]
            dt = ts[j] - ts[p]
            v = (P[:, j] - P[:, p]) / (dt + eps)
            out[self.placement(ts, j, gid) if reference_order else j] = absum(self.w * 0.5 * (v ** 2 * dt))
        return out

    def residual(self, x, signed=False, reference_order=True):
        x = np.asarray(x)
        rr = self.base.residual(self._std(x), signed=signed)
        return np.concatenate((rr, self.motion(x, signed=signed, reference_order=reference_order)))

    def cost(self, x):
        r = self.residual(x)
        return 0.5 * float(np.real(r) @ np.real(r))

    def jacobian_cs(self, x, cols=None, h=1e-30):
        """Complex-step derivative of abs(error_BA) (dense m x len(cols)); motion rows in global_traj order
        (row j = the triple centred at / the pair ending at point j)."""
        x = np.asarray(x, dtype=float)
        cols = np.arange(self.n) if cols is None else np.asarray(cols)
        sgn = np.ones(self.m)
        sgn[:self.base.m] = np.where(self.base.residual(self._std(x), signed=True) < 0, -1.0, 1.0)
        J = np.zeros((self.m, len(cols)))
        for q, c in enumerate(cols):
            xc = x.astype(complex)
            xc[c] += 1j * h
            J[:, q] = np.imag(self.residual(xc, signed=True, reference_order=False)) / h * sgn
        return J

    def free_mask(self):
        m = np.ones(self.n, dtype=bool)
        nc = self.nc
        if not self.base.opt_sync:
            m[:2 * nc] = False
        if not self.base.opt_rs:
            m[2 * nc:3 * nc] = False
        return m
