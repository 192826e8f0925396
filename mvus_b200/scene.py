"""Host-side mirror of the reference's flight/camera containers.

Only what the bundle-adjustment path reads or must leave behind is mirrored
(SURVEY.md section 8b): attribute names, array shapes and the ``BA`` signature are
those of ``reconstruction/common.py`` (``Scene`` at :22-62, ``Camera`` at
:1040-1069) so that a reference ``Scene`` and a mirror ``Scene`` are interchangeable
for ``mvus_b200.ba.bundle_adjust`` and a result pickles to the same field names.

The per-detection arithmetic (time stamps, undistortion, spline evaluation,
projection) is NOT done here: ``detection_to_global`` / ``error_cam`` /
``compute_visibility`` call the CUDA library through ``mvus_b200.ba``.
"""
import numpy as np

from . import hostmath


class Camera:
    """One camera: K (3x3), R (3x3), t (3,), d (5,), P = K [R|t], fps, resolution.
    Mirrors reconstruction/common.py:1040-1168 (fields and parameter-vector layout)."""

    def __init__(self, **kwargs):
        self.P = kwargs.get('P')
        self.K = kwargs.get('K')
        self.R = kwargs.get('R')
        self.t = kwargs.get('t')
        self.d = kwargs.get('d')
        self.c = kwargs.get('c')
        self.fps = kwargs.get('fps')
        self.resolution = kwargs.get('resolution')

    def compose(self):
        """P = K [R | t]   (common.py:1082-1083)."""
        self.P = self.K @ np.hstack((self.R, np.reshape(self.t, (3, 1))))
        return self.P

    def projectPoint(self, X):
        """Pinhole projection x = P X / (P X)_z of 3xN or 4xN points (common.py:1072-1079)."""
        assert self.P is not None, 'The projection matrix P has not been calculated yet'
        X = np.asarray(X, dtype=np.float64)
        if X.shape[0] == 3:
            X = np.vstack((X, np.ones(X.shape[1])))
        x = self.P @ X
        return x / x[2]

    def P2vector(self, calib=False):
        """[rvec, t] (6) or [fx, fy, cx, cy, rvec, t, d] (15)   (common.py:1113-1124)."""
        r = hostmath.matrix_to_rodrigues(self.R)
        t = np.asarray(self.t, dtype=np.float64).reshape(3)
        if calib:
            k = np.array([self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]])
            return np.concatenate((k, r, t, np.asarray(self.d, dtype=np.float64).reshape(5)))
        return np.concatenate((r, t))

    def vector2P(self, vector, calib=False):
        """Inverse of P2vector; recomposes P   (common.py:1127-1144)."""
        vector = np.asarray(vector, dtype=np.float64)
        if calib:
            self.K = np.eye(3)
            self.K[0, 0], self.K[1, 1] = vector[0], vector[1]
            self.K[:2, -1] = vector[2:4]
            self.R = hostmath.rodrigues_to_matrix(vector[4:7])
            self.t = vector[7:10].copy()
            self.d = vector[10:15].copy()
        else:
            self.R = hostmath.rodrigues_to_matrix(vector[:3])
            self.t = vector[3:6].copy()
        return self.compose()


class Scene:
    """Flight container (common.py:22-62).  ``BA`` keeps the reference signature
    (common.py:441) and post-conditions."""

    def __init__(self):
        self.numCam = 0
        self.cameras = []
        self.detections = []
        self.detections_raw = []
        self.detections_global = []
        self.alpha = []
        self.beta = []
        self.beta_after_Fbeta = []
        self.cf = []
        self.traj = []
        self.traj_len = []
        self.sequence = []
        self.visible = []
        self.settings = []
        self.gt = []
        self.out = {}
        self.spline = {'tck': [], 'int': []}
        self.rs = []
        self.ref_cam = 0
        self.find_order = True

    def addCamera(self, *camera):
        for c in camera:
            assert type(c) is Camera, "camera is not an instance of Camera"
            self.cameras.append(c)

    def addDetection(self, *detection):
        for d in detection:
            assert d.shape[0] == 3, "Detection must in form of (x,y,frameId)*N"
            self.detections.append(d)

    def init_alpha(self, *prior):
        """alpha_i = fps_ref / fps_i   (common.py:92-102)."""
        if len(prior):
            assert len(prior) == self.numCam
            self.alpha = prior
        else:
            fps_ref = self.cameras[self.ref_cam].fps
            self.alpha = np.array([fps_ref / c.fps for c in self.cameras], dtype=np.float64)

    # --- the BA path and its satellites: all device-side, see mvus_b200/ba.py ---------
    def BA(self, numCam, max_iter=10, rs=False, motion_prior=False, motion_reg=False,
           motion_weights=1, norm=False, rs_bounds=False):
        from . import ba
        return ba.bundle_adjust(self, numCam, max_iter=max_iter, rs=rs, motion_prior=motion_prior,
                                motion_reg=motion_reg, motion_weights=motion_weights, norm=norm,
                                rs_bounds=rs_bounds)

    def detection_to_global(self, *cam, motion_prior=False):
        from . import ba
        return ba.detection_to_global(self, *cam, motion_prior=motion_prior)

    def error_cam(self, cam_id, mode='dist', motion_prior=False, norm=False):
        from . import ba
        return ba.error_cam(self, cam_id, mode=mode, motion_prior=motion_prior, norm=norm)

    def compute_visibility(self):
        from . import ba
        return ba.compute_visibility(self)

    def remove_outliers(self, cams, thres=30, verbose=False):
        from . import ba
        return ba.remove_outliers(self, cams, thres=thres, verbose=verbose)

    def spline_to_traj(self, sampling_rate=1, t=None):
        from . import ba
        return ba.spline_to_traj(self, sampling_rate=sampling_rate, t=t)

    def traj_to_spline(self, smooth_factor):
        from . import splfit
        return splfit.traj_to_spline(self, smooth_factor)
