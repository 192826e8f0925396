"""Scene -> flat struct-of-arrays problem description for the C ABI (include/mvus_ba.h) and
back.  This is the packing half of ``Scene.BA`` (reconstruction/common.py:612-651, 672-692):
the same parameter-vector layout, the same camera order (``sequence[:numCam]``).

Works on the mirror ``mvus_b200.scene.Scene`` and, duck-typed, on the reference's own
``reconstruction.common.Scene`` (attribute names are identical).
"""
import numpy as np

from . import hostmath


class FlatProblem:
    """Everything ``mvus_ba_create / set_detections / set_splines`` need, as contiguous
    float64 / int arrays, plus x0 in the reference layout."""

    def __init__(self, scene, numCam, rs=False, motion_reg=False, motion_weights=1, rs_bounds=False,
                 max_iter=10):
        st = scene.settings
        self.nc = int(numCam)
        self.seq = [int(i) for i in scene.sequence[:numCam]]
        self.opt_calib = bool(st['opt_calib'])
        self.undist = bool(st['undist_points'])
        try:                                   # common.py:512-515: missing key -> optimise
            self.opt_sync = bool(st['opt_sync'])
        except (KeyError, TypeError):
            self.opt_sync = True
        self.opt_rs = bool(rs)
        self.rs_bounds = bool(rs_bounds)
        if motion_reg:
            mt = st['motion_type']             # KeyError if absent, as in the reference (common.py:414)
            assert mt == 'F' or mt == 'KE', 'Motion type must be either F or KE'
            self.motion_type = 1 if mt == 'F' else 2
        else:
            self.motion_type = 0
        self.motion_weight = float(motion_weights)
        self.max_nfev = int(max_iter)
        self.C = 15 if self.opt_calib else 6
        self.Pc = 3 + self.C
        self.P = self.Pc + 12

        # per-camera 3 x N_i arrays with contiguous rows (no copy when the Scene already holds
        # C-contiguous float64 arrays; np.loadtxt(...).T in the reference gives strided rows -> copy)
        self.dets = [np.ascontiguousarray(scene.detections[i], dtype=np.float64) for i in self.seq]
        self.N_cam = np.array([d.shape[1] for d in self.dets], dtype=np.int64)
        self.cam_ptr = np.concatenate(([0], np.cumsum(self.N_cam))).astype(np.int64)
        self.N = int(self.cam_ptr[-1])
        cams = [scene.cameras[i] for i in self.seq]
        for c in cams:
            # the device carries (fx, fy, cx, cy) only, as Camera.vector2P does (common.py:1127-1144);
            # a K with skew or a non-unit last row would silently project differently from P = K[R|t]
            K = np.asarray(c.K, dtype=np.float64)
            if abs(K[0, 1]) > 1e-12 * abs(K[0, 0]) or abs(K[1, 0]) > 0 or not np.array_equal(K[2], [0.0, 0.0, 1.0]):
                raise ValueError('camera matrix with skew / non-unit last row is not supported on the device path')
        self.height = np.array([c.resolution[1] for c in cams], dtype=np.float64)
        self.calib = np.array([[c.K[0, 0], c.K[1, 1], c.K[0, 2], c.K[1, 2]] +
                               list(np.asarray(c.d, dtype=np.float64).reshape(5)) for c in cams],
                              dtype=np.float64).reshape(self.nc, 9)

        tck = scene.spline['tck']
        self.S = len(tck)
        self.interval = np.ascontiguousarray(np.asarray(scene.spline['int'], dtype=np.float64).reshape(2, self.S))
        self.knots_list = [np.asarray(t[0], dtype=np.float64) for t in tck]
        self.degree = np.array([int(t[2]) for t in tck], dtype=np.int32)
        self.knot_ptr = np.concatenate(([0], np.cumsum([len(k) for k in self.knots_list]))).astype(np.int64)
        self.knots = np.ascontiguousarray(np.concatenate(self.knots_list)) if self.S else np.zeros(0)
        self.ncoef = np.array([len(t[1][0]) for t in tck], dtype=np.int64)
        self.ctrl_off = np.concatenate(([0], np.cumsum(self.ncoef))).astype(np.int64)
        self.n_ctrl = int(self.ctrl_off[-1])
        self.n_other = self.nc * self.Pc
        self.n = self.n_other + 3 * self.n_ctrl
        self.x0 = self.pack(scene)
        if self.rs_bounds:
            rho = self.x0[2 * self.nc:3 * self.nc]
            if (rho < 0.0).any() or (rho > 1.0).any():
                # scipy: least_squares raises for an infeasible start (least_squares.py, "`x0` is infeasible.")
                raise ValueError('`x0` is infeasible.')

    def _cat(self, row):
        return np.ascontiguousarray(np.concatenate([d[row] for d in self.dets])) if self.N else np.zeros(0)

    @property
    def frame(self):
        return self._cat(0)

    @property
    def x_raw(self):
        return self._cat(1)

    @property
    def y_raw(self):
        return self._cat(2)

    # -- common.py:616-650 -------------------------------------------------------------
    def pack(self, scene):
        seq = self.seq
        parts = [np.asarray(scene.alpha, dtype=np.float64)[seq], np.asarray(scene.beta, dtype=np.float64)[seq],
                 np.asarray(scene.rs, dtype=np.float64)[seq]]
        for i in seq:
            c = scene.cameras[i]
            if c.R is None or c.t is None:
                # a camera without a pose yet (main.py adds cameras one by one): only its time stamps and
                # undistorted observations can be asked for (detection_to_global, common.py:105-127)
                r, t = np.zeros(3), np.zeros(3)
            else:
                r = hostmath.matrix_to_rodrigues(c.R)
                t = np.asarray(c.t, dtype=np.float64).reshape(3)
            if self.opt_calib:
                parts.append(np.concatenate(([c.K[0, 0], c.K[1, 1], c.K[0, 2], c.K[1, 2]], r, t,
                                             np.asarray(c.d, dtype=np.float64).reshape(5))))
            else:
                parts.append(np.concatenate((r, t)))
        for t in scene.spline['tck']:
            parts.append(np.ravel(np.asarray(t[1], dtype=np.float64)))
        return np.ascontiguousarray(np.concatenate(parts))

    # -- common.py:672-692 -------------------------------------------------------------
    def unpack_into(self, scene, x):
        nc, C = self.nc, self.C
        seq = self.seq
        x = np.asarray(x, dtype=np.float64)
        for name, blk in (('alpha', x[:nc]), ('beta', x[nc:2 * nc]), ('rs', x[2 * nc:3 * nc])):
            arr = np.array(getattr(scene, name), dtype=np.float64)
            arr[seq] = blk
            setattr(scene, name, arr)
        cams = x[3 * nc:3 * nc + nc * C].reshape(nc, C)
        for k, i in enumerate(seq):
            scene.cameras[i].vector2P(cams[k].copy(), calib=self.opt_calib)
        off = self.n_other
        for s, t in enumerate(scene.spline['tck']):
            nco = int(self.ncoef[s])
            blk = x[off:off + 3 * nco].reshape(3, nco)
            t[1] = [blk[0].copy(), blk[1].copy(), blk[2].copy()]
            off += 3 * nco
