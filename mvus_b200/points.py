"""``Scene.BA(motion_prior=True)``: the discrete-trajectory mode of the reference (reconstruction/common.py:
466-467, 527-550, 587-605, 631-634, 681-687) on the GPU (csrc/ba_points.cuh through ``mvus_ba_solve_points``).

The unknowns are the camera side and the G points of ``global_traj``; the splines are constants.  Host side here:
the pre- and post-conditions of the reference's ``BA`` around the solve (global_traj bookkeeping, packing of the
interleaved point vector, the spline refit ``traj_to_spline`` the reference runs afterwards).  One deliberate
reading: the reference decides between its two refit branches (common.py:683-687) with the time stamps its LAST
residual evaluation left in ``global_traj[3]`` -- a finite-difference probe or a rejected trial; here the time
stamps at the returned solution are used.
"""
import numpy as np

from . import _cabi, ba, splfit
from .problem import FlatProblem


class _PointScene:
    """The Scene as the points handle sees it: same cameras, no detections, one pseudo-spline of G coefficients."""

    def __init__(self, scene, G):
        self.settings, self.sequence, self.cameras = scene.settings, scene.sequence, scene.cameras
        self.alpha, self.beta, self.rs = scene.alpha, scene.beta, scene.rs
        self.detections = [np.zeros((3, 0)) for _ in scene.detections]
        knots = np.concatenate(([0.0], np.arange(G, dtype=np.float64), [G - 1.0]))
        self.spline = {'tck': [[knots, [np.zeros(G), np.zeros(G), np.zeros(G)], 1]], 'int': np.array([[0.0], [G - 1.0]])}


def point_meta(scene, fp):
    """Per point of global_traj: camera slot (position in sequence[:numCam]), frame id, raw y / image height."""
    gt = scene.global_traj
    slot_of = {cam: k for k, cam in enumerate(fp.seq)}
    slot = np.array([slot_of[int(c)] for c in gt[1]], dtype=np.int32)
    frame = np.ascontiguousarray(gt[2], dtype=np.float64)
    yh = np.empty(gt.shape[1])
    for k, cam in enumerate(fp.seq):
        m = slot == k
        det = fp.dets[k]
        idx = np.searchsorted(det[0], frame[m])
        assert np.array_equal(det[0][idx], frame[m]), 'global_traj and detections disagree'
        yh[m] = det[2][idx] / fp.height[k]
    return slot, frame, yh


def reference_placement(ts, rows, gid):
    """The reference scatters the motion residuals with np.intersect1d(global_traj[3], ..., return_indices=True)
    (common.py:401-403): sorted by TIME STAMP, while they were listed interval by interval in global_traj order.
    -> destination index of the residual of every row (identity unless time stamps of cameras have crossed)."""
    key = np.lexsort((rows, gid[rows]))
    out = np.empty_like(rows)
    out[key] = rows[np.argsort(ts[rows], kind='stable')]
    return out


def bundle_adjust_points(scene, numCam, max_iter=10, rs=False, motion_weights=1, rs_bounds=False, ftol=1e-8,
                         xtol=1e-12, gtol=1e-8):
    from scipy.optimize import OptimizeResult
    assert len(scene.alpha) == scene.numCam and len(scene.beta) == scene.numCam, \
        'The Number of alpha and beta is wrong'
    mt = scene.settings['motion_type']
    assert mt == 'F' or mt == 'KE', 'Motion type must be either F or KE'
    fps = FlatProblem(scene, numCam, rs=rs, motion_reg=False, rs_bounds=rs_bounds, max_iter=max_iter)
    interval0 = np.array(scene.spline['int'], dtype=np.float64)
    hs = _cabi.Handle(fps, device=ba.DEVICE)
    hp = None
    try:
        # common.py:638-639: interpolate 3-D points for the detections of all cameras (also sets traj)
        ba._all_detect_to_traj(scene, fps, hs, fps.x0)
        dg = list(scene.detections_global) if len(scene.detections_global) == scene.numCam \
            else [[] for _ in range(scene.numCam)]
        for k, new in zip(fps.seq, hs.detections_global(fps.x0)):
            dg[k] = new
        scene.detections_global = dg
        scene.traj = np.array(scene.global_traj[3:])
        G = scene.global_traj.shape[1]
        n_other = fps.n_other
        print('Number of BA parameters is {}'.format(n_other + 3 * G))
        # common.py:493 compute_visibility with the pre-BA parameters
        vis = [None] * scene.numCam
        for k, v in zip(fps.seq, hs.visibility(fps.x0)):
            vis[k] = v
        others = [i for i in range(scene.numCam) if i not in fps.seq]
        if others:
            ba.detection_to_global(scene, others)
            for i in others:
                vis[i] = ba._interval_membership(scene.detections_global[i][0], np.asarray(scene.spline['int']))
        scene.visible = vis
        print('Doing BA with {} cameras...\n'.format(numCam))
        slot, frame, yh = point_meta(scene, fps)
        fpp = FlatProblem(_PointScene(scene, G), numCam, rs=rs, motion_reg=False, rs_bounds=rs_bounds, max_iter=max_iter)
        x0 = np.concatenate((fps.x0[:n_other], np.ravel(scene.global_traj[4:7])))          # planes X | Y | Z
        hp = _cabi.Handle(fpp, device=ba.DEVICE, ftol=ftol, xtol=xtol, gtol=gtol)
        hp.points_set(slot, frame, yh)
        x, r, st = hp.solve_points(hs, 1 if mt == 'F' else 2, motion_weights, fps.x0, x0)
        # ---- after BA (common.py:672-695) ----
        fps.unpack_into(scene, np.concatenate((x[:n_other], fps.x0[n_other:])))             # alpha, beta, rs, cameras
        P = x[n_other:].reshape(3, G)
        scene.global_traj[4:7] = P
        nc = fps.nc
        ts = x[slot] * (frame + x[2 * nc + slot] * yh) + x[nc + slot]
        scene.global_traj[3] = ts                      # detection_to_global(motion_prior=True), common.py:128-148
        if not (ts[1:] > ts[:-1]).all():
            scene.traj = np.array(scene.global_traj[3:, np.argsort(ts)])
        splfit.traj_to_spline(scene, scene.settings['smooth_factor'])
        ba.detection_to_global(scene)
    finally:
        if hp is not None:
            hp.close()
        hs.close()
    # res.x / res.fun in the reference's layout: points interleaved, motion rows where the reference puts them
    x_ref = np.concatenate((x[:n_other], np.ravel(P.T)))
    fun = r.copy()
    gid = ba._interval_membership(ts, interval0)
    rows = []
    for g in range(1, interval0.shape[1] + 1):
        idx = np.nonzero(gid == g)[0]
        rows.append(idx[1:-1] if mt == 'F' else idx[1:])
    rows = np.concatenate(rows) if rows else np.zeros(0, dtype=int)
    rm = r[2 * fps.N:]
    fun[2 * fps.N:] = 0.0
    fun[2 * fps.N + reference_placement(ts, rows, gid)] = rm[rows]
    res = OptimizeResult(x=x_ref, cost=st.cost, fun=fun, jac=None, grad=None, optimality=st.optimality,
                         active_mask=np.zeros(len(x_ref), dtype=int), nfev=st.nfev, njev=st.njev, status=st.status,
                         message=ba._MESSAGES.get(st.status, ''), success=st.status > 0)
    res.stats = st.as_dict()
    return res
