"""ctypes binding of libmvus_ba.so (include/mvus_ba.h).  There is NO fallback: if the CUDA
library is missing or no GPU is usable, every call raises."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libmvus_ba.so')

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)


class BADesc(ctypes.Structure):
    _fields_ = [('num_cams', ctypes.c_int32), ('opt_calib', ctypes.c_int32),
                ('undist_points', ctypes.c_int32), ('opt_sync', ctypes.c_int32),
                ('opt_rs', ctypes.c_int32), ('rs_bounds', ctypes.c_int32),
                ('motion_type', ctypes.c_int32), ('device', ctypes.c_int32),
                ('motion_weight', ctypes.c_double), ('max_nfev', ctypes.c_int32),
                ('solver_chunk', ctypes.c_int32), ('ftol', ctypes.c_double), ('xtol', ctypes.c_double),
                ('gtol', ctypes.c_double)]


class BAStats(ctypes.Structure):
    _fields_ = [('cost0', ctypes.c_double), ('cost', ctypes.c_double), ('optimality', ctypes.c_double),
                ('lam', ctypes.c_double), ('nfev', ctypes.c_int32), ('njev', ctypes.c_int32),
                ('status', ctypes.c_int32), ('lm_iterations', ctypes.c_int32),
                ('ms_total', ctypes.c_double), ('ms_resjac', ctypes.c_double),
                ('ms_accum', ctypes.c_double), ('ms_solve', ctypes.c_double),
                ('ms_trial', ctypes.c_double), ('launches', ctypes.c_int32), ('n_resjac', ctypes.c_int32),
                ('ms_syrk', ctypes.c_double), ('ms_bcr', ctypes.c_double), ('ms_reduce', ctypes.c_double),
                ('ms_k2', ctypes.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


EXPORTS = ['mvus_ba_version', 'mvus_ba_create', 'mvus_ba_destroy', 'mvus_ba_last_error',
           'mvus_ba_set_detections', 'mvus_ba_set_detections_rows', 'mvus_ba_set_splines', 'mvus_ba_dims', 'mvus_ba_residual',
           'mvus_ba_residual_jacobian', 'mvus_ba_solve', 'mvus_ba_detections_global',
           'mvus_ba_normal_equations', 'mvus_ba_global_traj', 'mvus_ba_spline_to_traj', 'mvus_ba_visibility', 'mvus_ba_host_alloc',
           'mvus_ba_host_free', 'mvus_ba_trim', 'mvus_ba_nccl_unique_id', 'mvus_ba_comm_init', 'mvus_ba_shard_bounds',
           'mvus_ba_time_resjac', 'mvus_ba_time_accumulate', 'mvus_ba_spl_create', 'mvus_ba_spl_destroy',
           'mvus_ba_spl_last_error', 'mvus_ba_spl_solve', 'mvus_ba_align', 'mvus_ba_points_set',
           'mvus_ba_solve_points', 'mvus_ba_points_eval']

_lib = None


class MvusError(RuntimeError):
    pass


def load():
    """Load the CUDA library; raises if it has not been built (no CPU path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MvusError('%s not built: run `python -m mvus_b200.build` (the BA path has no CPU fallback)'
                        % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.mvus_ba_version.restype = ctypes.c_char_p
    lib.mvus_ba_last_error.restype = ctypes.c_char_p
    lib.mvus_ba_last_error.argtypes = [ctypes.c_void_p]
    lib.mvus_ba_create.argtypes = [ctypes.POINTER(BADesc), ctypes.POINTER(ctypes.c_void_p)]
    lib.mvus_ba_destroy.argtypes = [ctypes.c_void_p]
    lib.mvus_ba_destroy.restype = None
    lib.mvus_ba_set_detections.argtypes = [ctypes.c_void_p, _lp, _dp, _dp, _dp, _dp, _dp]
    _pp = ctypes.POINTER(ctypes.c_void_p)
    lib.mvus_ba_set_detections_rows.argtypes = [ctypes.c_void_p, _lp, _pp, _pp, _pp, _dp, _dp]
    lib.mvus_ba_set_splines.argtypes = [ctypes.c_void_p, ctypes.c_int32, _dp, _lp, _dp, _ip]
    lib.mvus_ba_dims.argtypes = [ctypes.c_void_p, _lp, _lp, _lp, _lp, _ip]
    lib.mvus_ba_residual.argtypes = [ctypes.c_void_p, _dp, _dp]
    lib.mvus_ba_residual_jacobian.argtypes = [ctypes.c_void_p, _dp, _dp, _ip, _dp, _ip, _dp]
    lib.mvus_ba_solve.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, ctypes.POINTER(BAStats)]
    lib.mvus_ba_detections_global.argtypes = [ctypes.c_void_p, _dp, _dp]
    lib.mvus_ba_normal_equations.argtypes = [ctypes.c_void_p, _dp, _dp, _dp, _dp, _ip, _dp, _dp]
    lib.mvus_ba_global_traj.argtypes = [ctypes.c_void_p, _dp, _ip, _lp, _dp, _dp]
    lib.mvus_ba_spline_to_traj.argtypes = [ctypes.c_void_p, _dp, _dp, ctypes.c_int64, _lp, _dp]
    lib.mvus_ba_align.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int64, _dp, _dp, ctypes.c_int32, _dp, ctypes.c_int32,
                                  ctypes.c_int32, _dp, _lp, _dp, _dp]
    lib.mvus_ba_visibility.argtypes = [ctypes.c_void_p, _dp, _lp]
    lib.mvus_ba_host_alloc.argtypes = [ctypes.c_size_t]
    lib.mvus_ba_host_alloc.restype = ctypes.c_void_p
    lib.mvus_ba_host_free.argtypes = [ctypes.c_void_p]
    lib.mvus_ba_host_free.restype = None
    lib.mvus_ba_nccl_unique_id.argtypes = [ctypes.c_char_p]
    lib.mvus_ba_comm_init.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_char_p]
    lib.mvus_ba_trim.argtypes = [ctypes.c_int32, ctypes.c_uint64]
    lib.mvus_ba_shard_bounds.argtypes = [ctypes.c_void_p, ctypes.c_int32, _lp]
    lib.mvus_ba_spl_create.argtypes = [ctypes.c_int32, ctypes.c_int64, ctypes.c_int32, ctypes.c_int32, _dp, _dp,
                                       ctypes.POINTER(ctypes.c_void_p)]
    lib.mvus_ba_spl_destroy.argtypes = [ctypes.c_void_p]
    lib.mvus_ba_spl_destroy.restype = None
    lib.mvus_ba_spl_last_error.argtypes = [ctypes.c_void_p]
    lib.mvus_ba_spl_last_error.restype = ctypes.c_char_p
    lib.mvus_ba_spl_solve.argtypes = [ctypes.c_void_p, ctypes.c_int32, _dp, _dp, ctypes.c_double, _dp, _dp, _dp, _dp]
    lib.mvus_ba_time_resjac.argtypes = [ctypes.c_void_p, _dp, ctypes.c_int32, _dp]
    lib.mvus_ba_time_accumulate.argtypes = [ctypes.c_void_p, ctypes.c_int32, _dp]
    _lib = lib
    return lib


class PinnedPool:
    """Grow-only pool of page-locked host blocks for the big outputs.  `empty(n, dtype)` returns a
    NumPy array living in pinned memory; when the array (and every view of it) is garbage
    collected the block returns to the pool, so repeated BA calls (main.py makes 2 per camera)
    reuse the same pages.  Falls back to pageable memory if pinning fails or the cap is hit."""
    CAP_BYTES = 24 << 30

    MIN_BYTES = 16 << 20

    def __init__(self):
        import threading
        self.lock = threading.Lock()     # (batch.solve_scenes runs several BA calls from host threads)
        self.free = []          # (nbytes, ptr)
        self.total = 0
        self.new_bytes = 0      # bytes newly pinned (diagnostic)
        self.reused_bytes = 0

    def empty(self, n, dtype=np.float64):
        import weakref
        nbytes = max(int(n) * np.dtype(dtype).itemsize, 8)
        if nbytes < self.MIN_BYTES:         # pinning costs milliseconds: only worth it for large transfers
            return np.empty(int(n), dtype=dtype)
        lib = load()
        ptr = None
        with self.lock:
            for k, (sz, p) in enumerate(self.free):
                if nbytes <= sz <= 2 * nbytes + (1 << 20):
                    ptr, cap = p, sz
                    del self.free[k]
                    break
        if ptr is None:
            if self.total + nbytes > self.CAP_BYTES:
                return np.empty(int(n), dtype=dtype)
            ptr = lib.mvus_ba_host_alloc(nbytes)
            if not ptr:
                return np.empty(int(n), dtype=dtype)
            cap = nbytes
            self.total += nbytes
            self.new_bytes += nbytes
        else:
            self.reused_bytes += nbytes
        buf = (ctypes.c_char * cap).from_address(ptr)
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, self.free.append, (cap, ptr))
        return arr


POOL = PinnedPool()


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _l(a):
    return a.ctypes.data_as(_lp)


class Handle:
    """One BA problem on one GPU (wraps mvus_ba_handle)."""

    def __init__(self, fp, device=0, ftol=1e-8, xtol=1e-12, gtol=1e-8, max_nfev=None, solver_chunk=0):
        self.lib = load()
        self.fp = fp
        desc = BADesc(num_cams=fp.nc, opt_calib=int(fp.opt_calib), undist_points=int(fp.undist),
                      opt_sync=int(fp.opt_sync), opt_rs=int(fp.opt_rs), rs_bounds=int(fp.rs_bounds),
                      motion_type=int(fp.motion_type), device=int(device),
                      motion_weight=float(fp.motion_weight),
                      max_nfev=int(fp.max_nfev if max_nfev is None else max_nfev), solver_chunk=int(solver_chunk),
                      ftol=ftol, xtol=xtol, gtol=gtol)
        self.h = ctypes.c_void_p()
        rc = self.lib.mvus_ba_create(ctypes.byref(desc), ctypes.byref(self.h))
        if rc != 0:
            raise MvusError('mvus_ba_create: %s' % self.lib.mvus_ba_last_error(None).decode())
        self.reset_inputs(fp)

    def reset_inputs(self, fp):
        """(Re)load detections and splines of a FlatProblem with the same cameras / flags."""
        self.fp = fp
        nc = fp.nc
        rows = [(ctypes.c_void_p * nc)(*[int(d[k].ctypes.data) for d in fp.dets]) for k in range(3)]
        self._check(self.lib.mvus_ba_set_detections_rows(self.h, _l(fp.N_cam), rows[0], rows[1], rows[2],
                                                         _d(fp.height), _d(fp.calib)))
        interval = np.ascontiguousarray(fp.interval.reshape(-1))
        self._check(self.lib.mvus_ba_set_splines(self.h, fp.S, _d(interval), _l(fp.knot_ptr), _d(fp.knots),
                                                 _i(fp.degree)))
        n, m, N, M = (ctypes.c_int64() for _ in range(4))
        P = ctypes.c_int32()
        self._check(self.lib.mvus_ba_dims(self.h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(N),
                                          ctypes.byref(M), ctypes.byref(P)))
        self.n, self.m, self.N, self.M, self.P = n.value, m.value, N.value, M.value, P.value
        assert self.n == fp.n, (self.n, fp.n)

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.mvus_ba_last_error(self.h).decode()
            if rc == -3:
                raise ValueError(msg)       # scipy raises ValueError here too (least_squares.py:945)
            raise MvusError('mvus_ba error %d: %s' % (rc, msg))

    def close(self):
        if self.h:
            self.lib.mvus_ba_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def comm_init(self, world, rank, uid):
        self._check(self.lib.mvus_ba_comm_init(self.h, world, rank, uid))

    def shard_bounds(self, world):
        """Control-point bounds of the ranks' block ranges (mvus_ba_shard_bounds)."""
        b = np.zeros(world + 1, dtype=np.int64)
        self._check(self.lib.mvus_ba_shard_bounds(self.h, world, _l(b)))
        return b

    def residual(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.empty(self.m)
        self._check(self.lib.mvus_ba_residual(self.h, _d(x), _d(r)))
        return r

    def residual_jacobian(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.empty(self.m)
        span = np.empty(self.N, dtype=np.int32)
        J = np.empty(2 * self.P * self.N)
        mbase = np.empty(self.M, dtype=np.int32)
        mJ = np.empty(10 * self.M)
        self._check(self.lib.mvus_ba_residual_jacobian(self.h, _d(x), _d(r), _i(span), _d(J), _i(mbase), _d(mJ)))
        return r, span, J, mbase, mJ

    def solve(self, x0, want_r=True):
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.empty(self.n)
        r = POOL.empty(self.m) if want_r else None
        st = BAStats()
        self._check(self.lib.mvus_ba_solve(self.h, _d(x0), _d(x), _d(r), ctypes.byref(st)))
        return x, r, st

    def detections_global(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = POOL.empty(3 * self.N)
        self._check(self.lib.mvus_ba_detections_global(self.h, _d(x), _d(out)))
        # zero-copy 3 x N_i views, one per camera: the arrays Scene.detections_global holds
        cp = self.fp.cam_ptr
        return [out[3 * cp[k]:3 * cp[k + 1]].reshape(3, -1) for k in range(self.fp.nc)]

    def visibility(self, x):
        """Per camera: int64 array, 1-based interval id of each detection, 0 = none."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = POOL.empty(self.N, dtype=np.int64)
        self._check(self.lib.mvus_ba_visibility(self.h, _d(x), _l(out)))
        cp = self.fp.cam_ptr
        return [out[cp[k]:cp[k + 1]] for k in range(self.fp.nc)]

    def spline_to_traj(self, x, t):
        """4 x n' array [t; X; Y; Z] of the splines in x at the ascending times t inside the intervals."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        t = np.ascontiguousarray(t, dtype=np.float64)
        out = np.empty(4 * max(len(t), 1))
        n = ctypes.c_int64()
        self._check(self.lib.mvus_ba_spline_to_traj(self.h, _d(x), _d(t), len(t), ctypes.byref(n), _d(out)))
        return out[:4 * n.value].reshape(4, n.value)

    def global_traj(self, x, cam_ids):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ids = np.ascontiguousarray(cam_ids, dtype=np.int32)
        out = POOL.empty(7 * max(self.N, 1))
        gd = POOL.empty(3 * max(self.N, 1))
        n = ctypes.c_int64()
        self._check(self.lib.mvus_ba_global_traj(self.h, _d(x), _i(ids), ctypes.byref(n), _d(out), _d(gd)))
        return out[:7 * n.value].reshape(7, n.value), gd[:3 * self.N].reshape(3, self.N)

    def align(self, x, tau, pts, shifts, spline_is_src, want=-1):
        """mvus_ba_align: one similarity fit per shift between the handle's splines at tau + shift and the 3 x n
        points.  -> (mean_err [nshift], count [nshift], M [nshift, 4, 4], err [n] or None)."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        tau = np.ascontiguousarray(tau, dtype=np.float64)
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        shifts = np.ascontiguousarray(np.atleast_1d(shifts), dtype=np.float64)
        n, ns = len(tau), len(shifts)
        assert pts.shape == (3, n)
        mean_err, count, M = np.empty(ns), np.empty(ns, dtype=np.int64), np.empty((ns, 4, 4))
        err = np.empty(n) if want >= 0 else None
        self._check(self.lib.mvus_ba_align(self.h, _d(x), n, _d(tau), _d(pts), ns, _d(shifts), int(bool(spline_is_src)),
                                           int(want), _d(mean_err), _l(count), _d(M), _d(err)))
        return mean_err, count, M, err

    # ---- Scene.BA(motion_prior=True): this handle is the POINTS handle (see include/mvus_ba.h) ----
    def points_set(self, cam_slot, frame, y_over_height):
        cam_slot = np.ascontiguousarray(cam_slot, dtype=np.int32)
        frame = np.ascontiguousarray(frame, dtype=np.float64)
        yh = np.ascontiguousarray(y_over_height, dtype=np.float64)
        self.lib.mvus_ba_points_set.argtypes = [ctypes.c_void_p, ctypes.c_int64, _ip, _dp, _dp]
        self._check(self.lib.mvus_ba_points_set(self.h, len(cam_slot), _i(cam_slot), _d(frame), _d(yh)))
        self.G = len(cam_slot)

    def solve_points(self, hs, motion_type, weight, xs0, x0):
        xs0 = np.ascontiguousarray(xs0, dtype=np.float64)
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        x = np.empty(self.n)
        r = np.empty(2 * hs.N + self.G)
        st = BAStats()
        self.lib.mvus_ba_solve_points.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double,
                                                  _dp, _dp, _dp, _dp, ctypes.POINTER(BAStats)]
        self._check(self.lib.mvus_ba_solve_points(hs.h, self.h, int(motion_type), float(weight), _d(xs0), _d(x0),
                                                  _d(x), _d(r), ctypes.byref(st)))
        return x, r, st

    def points_eval(self, hs, motion_type, weight, xs0, x, want_g=True):
        xs0 = np.ascontiguousarray(xs0, dtype=np.float64)
        x = np.ascontiguousarray(x, dtype=np.float64)
        r = np.empty(2 * hs.N + self.G)
        g = np.empty(self.n) if want_g else None
        cost = ctypes.c_double()
        self.lib.mvus_ba_points_eval.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_double,
                                                 _dp, _dp, _dp, _dp, ctypes.POINTER(ctypes.c_double)]
        self._check(self.lib.mvus_ba_points_eval(hs.h, self.h, int(motion_type), float(weight), _d(xs0), _d(x),
                                                 _d(r), _d(g), ctypes.byref(cost)))
        return r, g, cost.value

    def normal_equations(self, x, want_dense=True):
        fp = self.fp
        x = np.ascontiguousarray(x, dtype=np.float64)
        A = np.empty(fp.nc * fp.Pc * fp.Pc)
        g = np.empty(self.n)
        band = ctypes.c_int32()
        cost = ctypes.c_double()
        # first call to learn the band width
        self._check(self.lib.mvus_ba_normal_equations(self.h, _d(x), _d(A), _d(g), None, ctypes.byref(band),
                                                      None, ctypes.byref(cost)))
        Hss = Hcs = None
        if want_dense:
            Hss = np.empty(fp.n_ctrl * band.value * 9)
            Hcs = np.empty(fp.nc * fp.Pc * 3 * fp.n_ctrl)
            self._check(self.lib.mvus_ba_normal_equations(self.h, _d(x), _d(A), _d(g), _d(Hss),
                                                          ctypes.byref(band), _d(Hcs), ctypes.byref(cost)))
            Hss = Hss.reshape(fp.n_ctrl, band.value, 3, 3)
            Hcs = Hcs.reshape(fp.nc * fp.Pc, 3 * fp.n_ctrl)
        return A.reshape(fp.nc, fp.Pc, fp.Pc), g, Hss, Hcs, cost.value

    def time_resjac(self, x, reps=5):
        x = np.ascontiguousarray(x, dtype=np.float64)
        ms = ctypes.c_double()
        self._check(self.lib.mvus_ba_time_resjac(self.h, _d(x), reps, ctypes.byref(ms)))
        return ms.value

    def time_accumulate(self, reps=5):
        ms = ctypes.c_double()
        self._check(self.lib.mvus_ba_time_accumulate(self.h, reps, ctypes.byref(ms)))
        return ms.value


class SplHandle:
    """Data of one spline interval on the GPU (wraps mvus_spl_handle): u[m] ascending, x[idim][m]."""

    def __init__(self, u, x, k=3, device=0):
        self.lib = load()
        self.u = np.ascontiguousarray(u, dtype=np.float64)
        self.x = np.ascontiguousarray(np.atleast_2d(x), dtype=np.float64)
        self.k, self.idim, self.m = int(k), self.x.shape[0], len(self.u)
        self.h = ctypes.c_void_p()
        rc = self.lib.mvus_ba_spl_create(int(device), self.m, self.idim, self.k, _d(self.u), _d(self.x), ctypes.byref(self.h))
        if rc != 0:
            msg = self.lib.mvus_ba_spl_last_error(None).decode()
            if rc == -1:
                raise TypeError(msg)               # splprep raises TypeError('m > k must hold') here
            raise MvusError('mvus_ba_spl_create: %s' % msg)

    def solve(self, t, pen=None, pscale=0.0):
        """-> (c [idim x (n-k-1)], fp, fpint [n-2k-1], diag_sum) on the knots t."""
        t = np.ascontiguousarray(t, dtype=np.float64)
        n, k = len(t), self.k
        c = np.empty((self.idim, n - k - 1))
        fpint = np.empty(n - 2 * k - 1)
        fp, ds = ctypes.c_double(), ctypes.c_double()
        if pen is not None:
            pen = np.ascontiguousarray(pen, dtype=np.float64)
            assert pen.shape == (k + 2, n - k - 1)
        rc = self.lib.mvus_ba_spl_solve(self.h, n, _d(t), _d(pen), float(pscale), _d(c), ctypes.byref(fp), _d(fpint),
                                        ctypes.byref(ds))
        if rc != 0:
            raise MvusError('mvus_ba_spl_solve %d: %s' % (rc, self.lib.mvus_ba_spl_last_error(self.h).decode()))
        return c, fp.value, fpint, ds.value

    def close(self):
        if self.h:
            self.lib.mvus_ba_spl_destroy(self.h)
            self.h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
