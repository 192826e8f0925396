"""Drop-in bundle adjustment: the reference's ``Scene.BA`` (reconstruction/common.py:441-697)
with the whole inner loop on the GPU.

``bundle_adjust(scene, numCam, ...)`` keeps the reference signature, reads the same Scene
attributes and leaves the same post-conditions:
  * alpha/beta/rs[sequence[:nc]] updated (common.py:676)
  * cameras: R, t (and K, d under opt_calib) set, P recomposed (common.py:678-680, 1143)
  * spline['tck'][s][1] = [cx, cy, cz] (common.py:689-692)
  * detections_global refreshed for ALL cameras (common.py:695)
  * visible set from the PRE-BA parameters (common.py:493 -> compute_visibility)
  * under motion_reg: traj = unit-step samples and the global_* bookkeeping arrays
    (common.py:464, 630, 887-944) -- they are pickled outputs (README.md:216-221)
and returns a scipy-style OptimizeResult.

Every per-detection computation goes through the C ABI (mvus_b200/_cabi.py ->
libmvus_ba.so).  If that library is missing or there is no GPU this module raises; there is
no CPU path.
"""
import numpy as np

from . import _cabi
from .problem import FlatProblem

try:                                       # scipy's result container, when available
    from scipy.optimize import OptimizeResult
except Exception:                          # pragma: no cover
    class OptimizeResult(dict):
        __getattr__ = dict.get
        __setattr__ = dict.__setitem__

_MESSAGES = {-1: 'Linear solve failed (normal matrix not positive definite at maximum damping).',
             0: 'The maximum number of function evaluations is exceeded.',
             1: '`gtol` termination condition is satisfied.',
             2: '`ftol` termination condition is satisfied.',
             3: '`xtol` termination condition is satisfied.',
             4: 'Both `ftol` and `xtol` termination conditions are satisfied.'}

DEVICE = 0
_COMM = None          # (world, rank, uid bytes) set by mvus_b200.shard.init_comm


def _all_cams_problem(scene):
    """A FlatProblem over ALL cameras (used for detections_global / visibility refresh)."""
    class _S:
        pass
    s = _S()
    s.__dict__.update({k: getattr(scene, k) for k in ('settings', 'cameras', 'detections', 'alpha', 'beta',
                                                     'rs', 'spline')})
    s.sequence = list(range(scene.numCam))
    return FlatProblem(s, scene.numCam)


def _interval_membership(t, interval):
    """util.sampling(..., belong=True) index (util.py:103-106) for already-computed times."""
    idx = np.zeros(len(t), dtype=int)
    for s in range(interval.shape[1]):
        mask = np.logical_xor(t - interval[0, s] >= 0, t - interval[1, s] >= 0)
        idx[mask] = s + 1
    return idx


def detection_to_global(scene, *cam, motion_prior=False):
    """Scene.detection_to_global (common.py:105-127): global time stamps and (undistorted)
    observations, computed by the library's det_global kernel."""
    if motion_prior:
        raise NotImplementedError('motion_prior=True (discrete-trajectory mode) is not part of the BA path '
                                  'main.py uses; see SURVEY.md 8b / 8f')
    assert len(scene.alpha) == scene.numCam and len(scene.beta) == scene.numCam, \
        'The Number of alpha and beta is wrong'
    if len(cam):
        cams = cam
        if not isinstance(cams[0], (int, np.integer)):
            cams = cams[0]
        cams = [int(c) for c in cams]
    else:
        cams = list(range(scene.numCam))
        scene.detections_global = [[] for _ in cams]

    class _S:
        pass
    s = _S()
    s.__dict__.update({k: getattr(scene, k) for k in ('settings', 'cameras', 'detections', 'alpha', 'beta',
                                                     'rs', 'spline')})
    s.sequence = cams
    fp = FlatProblem(s, len(cams))
    hd = _cabi.Handle(fp, device=DEVICE)
    try:
        dg = hd.detections_global(fp.x0)
    finally:
        hd.close()
    while len(scene.detections_global) < scene.numCam:
        scene.detections_global.append([])
    for k, i in enumerate(cams):
        scene.detections_global[i] = dg[k]


def compute_visibility(scene):
    """Scene.compute_visibility (common.py:427-438)."""
    detection_to_global(scene)
    interval = np.asarray(scene.spline['int'], dtype=np.float64)
    scene.visible = [_interval_membership(scene.detections_global[i][0], interval)
                     for i in range(scene.numCam)]


def error_cam(scene, cam_id, mode='dist', motion_prior=False, norm=False):
    """Scene.error_cam (common.py:304-359) through the residual kernel.  Modes as in the
    reference; 'each' keeps zeros for uncovered detections, the others drop them."""
    if motion_prior or norm:
        raise NotImplementedError('error_cam(motion_prior/norm=True) is outside the BA path (SURVEY.md 8f)')

    class _S:
        pass
    s = _S()
    s.__dict__.update({k: getattr(scene, k) for k in ('settings', 'cameras', 'detections', 'alpha', 'beta',
                                                     'rs', 'spline')})
    s.sequence = [int(cam_id)]
    fp = FlatProblem(s, 1)
    hd = _cabi.Handle(fp, device=DEVICE)
    try:
        r = hd.residual(fp.x0)                      # residual-only K1: 40 B per detection, no Jacobian
        cov = hd.visibility(fp.x0)[0] > 0           # covered by a spline interval (util.py:103-106)
        dg = hd.detections_global(fp.x0)
    finally:
        hd.close()
    if len(scene.detections_global) < scene.numCam:       # fresh Scene: detection_to_global not called yet
        scene.detections_global = list(scene.detections_global) + [[] for _ in range(scene.numCam - len(scene.detections_global))]
    scene.detections_global[cam_id] = dg[0]
    N = fp.N
    eu, ev = r[:N], r[N:2 * N]
    if mode == 'each':
        return np.concatenate((eu, ev))
    # the reference concatenates interval by interval; detections are time-sorted so this is the same order
    if mode == 'dist':
        return np.sqrt(eu[cov] ** 2 + ev[cov] ** 2)
    if mode == 'xy_1D':
        return np.concatenate((eu[cov], ev[cov]))
    if mode == 'xy_2D':
        return np.vstack((eu[cov], ev[cov]))
    raise ValueError('unknown mode %r' % (mode,))


def remove_outliers(scene, cams, thres=30, verbose=False):
    """Scene.remove_outliers (common.py:700-717)."""
    if not thres:
        return
    for i in cams:
        e = error_cam(scene, i, mode='each')
        ex, ey = np.split(e, 2)
        err = np.sqrt(ex ** 2 + ey ** 2)
        scene.detections[i] = scene.detections[i][:, err < thres]
        detection_to_global(scene, i)
        if verbose:
            print('{} out of {} detections are removed for camera {}'.format(
                int(np.sum(err >= thres)), int(np.sum(err != 0)), i))


def _spline_only_problem(scene):
    """A FlatProblem with the Scene's splines and one camera without detections: enough for the
    entry points that only evaluate splines (also on a Scene that has no cameras yet)."""
    class _S:
        pass
    s = _S()
    s.spline = scene.spline
    s.settings = {'opt_calib': False, 'undist_points': False}
    s.alpha, s.beta, s.rs = np.ones(1), np.zeros(1), np.zeros(1)
    cam = _S()
    cam.K, cam.R, cam.t, cam.d, cam.resolution = np.eye(3), np.eye(3), np.zeros(3), np.zeros(5), [1, 1]
    s.cameras, s.detections, s.sequence = [cam], [np.zeros((3, 0))], [0]
    return FlatProblem(s, 1)


def spline_to_traj(scene, sampling_rate=1, t=None, hd=None, x=None):
    """Scene.spline_to_traj (common.py:273-301) through the library (mvus_ba_spline_to_traj):
    samples the splines at a constant rate or at the given ascending time stamps and stores the
    4 x n result in ``scene.traj``."""
    interval = np.asarray(scene.spline['int'])
    if t is not None:
        assert len(t.shape) == 1, 'Input timestamps must be a 1D array'
        timestamp = np.asarray(t, dtype=np.float64)
    else:
        timestamp = np.arange(interval[0, 0], interval[1, -1], sampling_rate)
    assert (timestamp[1:] >= timestamp[:-1]).all(), 'time stamps must be ascending'
    own = hd is None
    if own:
        fp = _spline_only_problem(scene)
        hd = _cabi.Handle(fp, device=DEVICE)
        x = fp.x0
    try:
        scene.traj = hd.spline_to_traj(x, timestamp)
    finally:
        if own:
            hd.close()
    return scene.traj


def _all_detect_to_traj(scene, fp, hd, x):
    """Bookkeeping of Scene.all_detect_to_traj (common.py:887-944): global_time_stamps_all,
    frame_id_all, global_detections from the refreshed detections_global; global_traj (sorted,
    in-interval detections with their spline positions) from the device (mvus_ba_global_traj)."""
    cams = fp.seq
    # drop the previous call's arrays first so that their pinned blocks are reused
    scene.global_detections = scene.frame_id_all = scene.global_time_stamps_all = scene.global_traj = None
    scene.global_traj, gd = hd.global_traj(x, cams)
    scene.global_detections = gd
    scene.frame_id_all = gd[1]                 # views of global_detections (same values, no copy)
    scene.global_time_stamps_all = gd[2]


def bundle_adjust(scene, numCam, max_iter=10, rs=False, motion_prior=False, motion_reg=False,
                  motion_weights=1, norm=False, rs_bounds=False, ftol=1e-8, xtol=1e-12, gtol=1e-8,
                  return_handle=False, bookkeeping=True):
    """Scene.BA(numCam, max_iter, rs, motion_prior, motion_reg, motion_weights, norm, rs_bounds)
    -- reference signature (common.py:441); ``max_iter`` is scipy's max_nfev (common.py:670),
    xtol = 1e-12 as the reference passes, ftol / gtol = scipy defaults."""
    if motion_prior:                        # discrete-trajectory mode (common.py:466-467 ...): mvus_b200/points.py
        from . import points
        return points.bundle_adjust_points(scene, numCam, max_iter=max_iter, rs=rs, motion_weights=motion_weights,
                                           rs_bounds=rs_bounds, ftol=ftol, xtol=xtol, gtol=gtol)
    assert len(scene.alpha) == scene.numCam and len(scene.beta) == scene.numCam, \
        'The Number of alpha and beta is wrong'
    import time as _time
    _t = [_time.perf_counter()]
    host = {}

    def lap(name):
        _t.append(_time.perf_counter())
        host[name] = host.get(name, 0.0) + (_t[-1] - _t[-2]) * 1e3

    fp = FlatProblem(scene, numCam, rs=rs, motion_reg=motion_reg, motion_weights=motion_weights,
                     rs_bounds=rs_bounds, max_iter=max_iter)
    lap('pack_ms')
    print('Number of BA parameters is {}'.format(fp.n))
    interval = np.asarray(scene.spline['int'], dtype=np.float64)
    others = [i for i in range(scene.numCam) if i not in fp.seq]

    def refresh(hd, x):
        """detections_global of the optimised cameras from the BA handle itself (one upload of
        the detections serves visibility, solve and refresh); cameras outside sequence[:numCam]
        go through a second, small handle."""
        dg = list(scene.detections_global) if len(scene.detections_global) == scene.numCam \
            else [[] for _ in range(scene.numCam)]
        for i in fp.seq:                    # drop the stale arrays first: their pinned block is reused
            dg[i] = None
        scene.detections_global = dg
        new = hd.detections_global(x)
        for k, i in enumerate(fp.seq):
            dg[i] = new[k]
        scene.detections_global = dg
        if others:
            detection_to_global(scene, others)

    def visibility(hd, x):
        """Scene.visible with the PRE-BA parameters (common.py:493): interval ids from the device
        for the optimised cameras, host membership test for the (few) remaining ones."""
        vis = [None] * scene.numCam
        for k, v in zip(fp.seq, hd.visibility(x)):
            vis[k] = v
        if others:
            detection_to_global(scene, others)
            for i in others:
                vis[i] = _interval_membership(scene.detections_global[i][0], interval)
        scene.visible = vis

    print('Doing BA with {} cameras...\n'.format(numCam))
    hd = _cabi.Handle(fp, device=DEVICE, ftol=ftol, xtol=xtol, gtol=gtol)
    lap('upload_ms')
    try:
        if _COMM is not None and _COMM[0] > 1:
            hd.comm_init(*_COMM)
        scene.visible = None                # (stale array dropped first: pinned block reuse)
        visibility(hd, fp.x0)               # common.py:493
        lap('visibility_ms')
        x, r, st = hd.solve(fp.x0)
        lap('solve_ms')
        fp.unpack_into(scene, x)            # common.py:672-692
        refresh(hd, x)                      # common.py:695
        lap('refresh_ms')
        if motion_reg and bookkeeping:
            _all_detect_to_traj(scene, fp, hd, x)
            spline_to_traj(scene, hd=hd, x=x)   # common.py:379 leaves traj = unit-step samples
            lap('bookkeeping_ms')
    finally:
        if not return_handle:
            hd.close()
            lap('close_ms')

    res = OptimizeResult(x=x, cost=st.cost, fun=r, jac=None, grad=None, optimality=st.optimality,
                         active_mask=np.zeros(fp.n, dtype=int), nfev=st.nfev, njev=st.njev,
                         status=st.status, message=_MESSAGES.get(st.status, ''), success=st.status > 0)
    res.stats = st.as_dict()
    res.stats['host'] = host
    if return_handle:
        res.handle = hd
    return res
